// What bounds the scattered per-variable accesses?  Sweeps warps per SM, loads in flight per lane and the load flavour.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scatter_cost2 scatter_cost2.cu && ./scatter_cost2
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ld_cg(const float* p) { return __ldcg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ld_ca(const float* p) { float2 v; asm volatile("ld.global.ca.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v; }
__device__ __forceinline__ float2 ld_nc(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ld_lu(const float* p) { float2 v; asm volatile("ld.global.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); return v; }

template<int KIND, int DEPTH>   // KIND 0 cg, 1 ca, 2 nc, 3 no_allocate, 4 red.f32, 5 one 4-byte cg load
__global__ void k(const unsigned* __restrict__ idx, float* buf, float* out, int H, long long* cyc)
{
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const unsigned* my = idx + (size_t)warp * H * 32 + lane;
    float acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if(KIND == 4)
    {
        for(int h = 0; h < H; ++h) atomicAdd(buf + 2 * (size_t)my[h * 32] + (h & 1), 1.0f);
        __threadfence();
    }
    else
        for(int h0 = 0; h0 < H; h0 += DEPTH)
        {
            float2 v[DEPTH];
#pragma unroll
            for(int q = 0; q < DEPTH; ++q)
            {
                const float* p = buf + 2 * (size_t)my[min(h0 + q, H - 1) * 32];
                if(KIND == 0) v[q] = ld_cg(p);
                if(KIND == 1) v[q] = ld_ca(p);
                if(KIND == 2) v[q] = ld_nc(p);
                if(KIND == 3) v[q] = ld_lu(p);
                if(KIND == 5) { v[q].x = __ldcg(p); v[q].y = 0; }
            }
#pragma unroll
            for(int q = 0; q < DEPTH; ++q) acc += v[q].x + v[q].y;
        }
    const long long t1 = clock64();
    if(acc == 12345.f) out[0] = acc;
    if(lane == 0) cyc[warp] = t1 - t0;
}

template<int KIND, int DEPTH>
void run(const char* name, int wpc, const unsigned* idx, float* buf, float* out, long long* cyc, int H)
{
    const int CTAS = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for(int rep = 0; rep < 10; ++rep)
    {
        cudaEventRecord(e0);
        k<KIND, DEPTH><<<CTAS, wpc * 32>>>(idx, buf, out, H, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if(ms < best) best = ms;
    }
    std::vector<long long> h(CTAS * wpc);
    cudaMemcpy(h.data(), cyc, h.size() * 8, cudaMemcpyDeviceToHost);
    double a = 0; for(auto x : h) a += x;
    a /= h.size();
    printf("%-28s warps/SM %2d depth %2d: %7.0f cycles per warp for %d rounds -> %.2f cycles per lane-access per SM, launch %.2f us\n",
           name, wpc, DEPTH, a, H, a / ((double)H * 32 * wpc), best * 1e3);
}

int main()
{
    const int V = 50000, H = 21, MAXW = 148 * 24;
    std::vector<unsigned> h_idx((size_t)MAXW * H * 32);
    srand(1);
    for(auto& x : h_idx) x = rand() % V;
    unsigned* idx; float* buf; float* out; long long* cyc;
    cudaMalloc(&idx, h_idx.size() * 4); cudaMalloc(&buf, 2 * V * 4); cudaMalloc(&out, 4); cudaMalloc(&cyc, MAXW * 8);
    cudaMemcpy(idx, h_idx.data(), h_idx.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(buf, 0, 2 * V * 4);
    for(int wpc : {1, 2, 6, 12, 24})
    {
        run<0, 8>("ld.cg.v2 (STRONG.GPU)", wpc, idx, buf, out, cyc, H);
        run<0, 21>("ld.cg.v2 (STRONG.GPU)", wpc, idx, buf, out, cyc, H);
        run<1, 8>("ld.ca.v2", wpc, idx, buf, out, cyc, H);
        run<1, 21>("ld.ca.v2", wpc, idx, buf, out, cyc, H);
        run<2, 21>("ld.nc.v2", wpc, idx, buf, out, cyc, H);
        run<3, 21>("ld.L1::no_allocate.v2", wpc, idx, buf, out, cyc, H);
        run<5, 21>("ld.cg.f32", wpc, idx, buf, out, cyc, H);
        run<4, 1>("red.add.f32 + fence", wpc, idx, buf, out, cyc, H);
    }
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
