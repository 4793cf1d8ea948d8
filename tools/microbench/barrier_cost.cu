// Cost of (a) scattered 4-byte global stores incl. their drain and (b) two grid-barrier implementations, at the launch shape of the
// on-chip kernel (148 CTAs x 192 threads, cooperative).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o barrier_cost barrier_cost.cu && ./barrier_cost
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void barrier_counter(unsigned* bar, unsigned n_ctas)
{
    __syncthreads();
    if(threadIdx.x == 0)
    {
        unsigned gen, old;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
        if(old == n_ctas - 1)
        {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(bar), "r"(0u) : "memory");
            asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(bar + 1), "r"(gen + 1u) : "memory");
        }
        else { unsigned seen; do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar + 1) : "memory"); } while(seen == gen); }
    }
    __syncthreads();
}
// all-to-all flags: every CTA publishes its epoch in its own slot, thread t of every CTA waits for slot t
__device__ __forceinline__ void barrier_flags(unsigned* flags, unsigned n_ctas, unsigned epoch)
{
    __syncthreads();
    if(threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flags + blockIdx.x), "r"(epoch) : "memory");
    for(unsigned t = threadIdx.x; t < n_ctas; t += blockDim.x)
    {
        unsigned seen;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + t) : "memory"); } while((int)(seen - epoch) < 0);
    }
    __syncthreads();
}
// same with the flags 128 bytes apart (one L2 line per CTA)
__device__ __forceinline__ void barrier_flags_wide(unsigned* flags, unsigned n_ctas, unsigned epoch)
{
    __syncthreads();
    if(threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flags + 32 * blockIdx.x), "r"(epoch) : "memory");
    for(unsigned t = threadIdx.x; t < n_ctas; t += blockDim.x)
    {
        unsigned seen;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + 32 * t) : "memory"); } while((int)(seen - epoch) < 0);
    }
    __syncthreads();
}

template<int KIND>
__global__ void kb(unsigned* bar, unsigned* flags, unsigned epoch0, int n, long long* cyc)
{
    __syncthreads();
    const long long t0 = clock64();
    for(int i = 0; i < n; ++i)
    {
        if(KIND == 0) barrier_counter(bar, gridDim.x);
        if(KIND == 1) barrier_flags(flags, gridDim.x, epoch0 + i + 1);
        if(KIND == 2) barrier_flags_wide(flags, gridDim.x, epoch0 + i + 1);
    }
    const long long t1 = clock64();
    if(threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int KIND>   // 0: scattered st.global.f32 + fence, 1: scattered st.global.v2.f32 + fence, 2: red + fence
__global__ void ks(const unsigned* __restrict__ idx, float* buf, int H, long long* cyc)
{
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const unsigned* my = idx + (size_t)warp * H * 32 + lane;
    __syncthreads();
    const long long t0 = clock64();
    for(int h = 0; h < H; ++h)
    {
        const unsigned v = my[h * 32];
        if(KIND == 0) buf[v] = (float)h;
        if(KIND == 1) reinterpret_cast<float2*>(buf)[v] = make_float2((float)h, 1.f);
        if(KIND == 2) atomicAdd(buf + v, 1.0f);
    }
    __threadfence();
    const long long t1 = clock64();
    if(lane == 0) cyc[warp] = t1 - t0;
}

int main()
{
    const int CTAS = 148, WPC = 6, H = 21, N = 525000;
    unsigned *bar, *flags, *idx; long long* cyc; float* buf;
    cudaMalloc(&bar, 64); cudaMalloc(&flags, 148 * 128 + 128); cudaMalloc(&cyc, 148 * 24 * 8);
    cudaMemset(bar, 0, 64); cudaMemset(flags, 0, 148 * 128 + 128);
    std::vector<unsigned> h_idx((size_t)CTAS * 24 * H * 32);
    srand(2);
    for(auto& x : h_idx) x = rand() % N;
    cudaMalloc(&idx, h_idx.size() * 4); cudaMemcpy(idx, h_idx.data(), h_idx.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&buf, 2 * N * 4); cudaMemset(buf, 0, 2 * N * 4);
    std::vector<long long> h(148 * 24);
    unsigned epoch = 0;
    const int n = 64;
    for(int kind = 0; kind < 3; ++kind)
        for(int rep = 0; rep < 3; ++rep)
        {
            void* args[] = { &bar, &flags, &epoch, (void*)&n, &cyc };
            const void* f = kind == 0 ? (const void*)kb<0> : kind == 1 ? (const void*)kb<1> : (const void*)kb<2>;
            cudaError_t e = cudaLaunchCooperativeKernel(f, dim3(CTAS), dim3(WPC * 32), args, 0, 0);
            cudaDeviceSynchronize();
            epoch += n;
            cudaMemcpy(h.data(), cyc, CTAS * 8, cudaMemcpyDeviceToHost);
            double a = 0; for(int i = 0; i < CTAS; ++i) a += h[i];
            printf("barrier %-24s: %.0f cycles per barrier (%d in a row, 148 CTAs x %d threads) %s\n",
                   kind == 0 ? "counter + generation" : kind == 1 ? "all-to-all flags" : "all-to-all flags, 128 B", a / CTAS / n, n, WPC * 32, cudaGetErrorString(e));
        }
    for(int wpc : {6, 12})
        for(int kind = 0; kind < 3; ++kind)
        {
            for(int rep = 0; rep < 3; ++rep)
            {
                if(kind == 0) ks<0><<<CTAS, wpc * 32>>>(idx, buf, H, cyc);
                if(kind == 1) ks<1><<<CTAS, wpc * 32>>>(idx, buf, H, cyc);
                if(kind == 2) ks<2><<<CTAS, wpc * 32>>>(idx, buf, H, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(h.data(), cyc, CTAS * wpc * 8, cudaMemcpyDeviceToHost);
            double a = 0; for(int i = 0; i < CTAS * wpc; ++i) a += h[i];
            a /= CTAS * wpc;
            printf("scattered %-22s + fence, warps/SM %2d: %.0f cycles per warp for %d rounds -> %.2f cycles per lane-access per SM\n",
                   kind == 0 ? "st.f32 (525k slots)" : kind == 1 ? "st.v2.f32" : "red.add.f32 (525k slots)", wpc, a, H, a / ((double)H * 32 * wpc));
        }
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
