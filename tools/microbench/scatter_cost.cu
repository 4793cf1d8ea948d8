// Throughput of the scattered per-variable accesses of one pass: every warp does H rounds of one 8-byte load (ld.global.cg.v2.f32)
// and/or one red.global.add.f32 per lane at random pair addresses of a 2V-float buffer (the shape of the 1 M-node set-cover
// instance: 148 CTAs x 6 warps, H = 21, V = 50 000).  Prints cycles per warp and microseconds per launch.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scatter_cost scatter_cost.cu && ./scatter_cost
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template<int MODE>   // 1 loads, 2 reds, 3 both, 4 reds that add to per-lane-private addresses (no sharing)
__global__ void k(const unsigned* __restrict__ idx, float* buf, float* out, int H, long long* cyc)
{
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const unsigned* my = idx + (size_t)warp * H * 32 + lane;
    float acc = 0;
    const long long t0 = clock64();
    if(MODE & 1)
    {
        float2 v[8];
        for(int h0 = 0; h0 < H; h0 += 8)
        {
#pragma unroll
            for(int k2 = 0; k2 < 8; ++k2) if(h0 + k2 < H) v[k2] = __ldcg(reinterpret_cast<const float2*>(buf + 2 * (size_t)my[(h0 + k2) * 32])); else v[k2] = make_float2(0, 0);
#pragma unroll
            for(int k2 = 0; k2 < 8; ++k2) acc += v[k2].x + v[k2].y;
        }
    }
    const long long t1 = clock64();
    if(MODE & 6)
        for(int h = 0; h < H; ++h)
        {
            const unsigned v = (MODE & 4) ? (unsigned)(warp * 32 + lane) : my[h * 32];
            atomicAdd(buf + 2 * (size_t)v + (h & 1), 1.0f);
        }
    const long long t2 = clock64();
    if(acc == 12345.f) out[0] = acc;
    if(lane == 0) { cyc[2 * warp] = t1 - t0; cyc[2 * warp + 1] = t2 - t1; }
}

int main()
{
    const int V = 50000, H = 21, CTAS = 148, WPC = 6, NW = CTAS * WPC;
    std::vector<unsigned> h_idx((size_t)NW * H * 32);
    srand(1);
    for(auto& x : h_idx) x = rand() % V;
    unsigned* idx; float* buf; float* out; long long* cyc;
    cudaMalloc(&idx, h_idx.size() * 4); cudaMalloc(&buf, (2 * V + NW * 64) * 4); cudaMalloc(&out, 4); cudaMalloc(&cyc, NW * 16);
    cudaMemcpy(idx, h_idx.data(), h_idx.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(buf, 0, (2 * V + NW * 64) * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<long long> h_cyc(2 * NW);
    for(int mode = 1; mode <= 4; ++mode)
    {
        float best = 1e9f;
        for(int rep = 0; rep < 20; ++rep)
        {
            cudaEventRecord(e0);
            if(mode == 1) k<1><<<CTAS, WPC * 32>>>(idx, buf, out, H, cyc);
            if(mode == 2) k<2><<<CTAS, WPC * 32>>>(idx, buf, out, H, cyc);
            if(mode == 3) k<3><<<CTAS, WPC * 32>>>(idx, buf, out, H, cyc);
            if(mode == 4) k<4><<<CTAS, WPC * 32>>>(idx, buf, out, H, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if(ms < best) best = ms;
        }
        cudaMemcpy(h_cyc.data(), cyc, NW * 16, cudaMemcpyDeviceToHost);
        double a = 0, b = 0; for(int w = 0; w < NW; ++w) { a += h_cyc[2 * w]; b += h_cyc[2 * w + 1]; }
        printf("mode %d (%s): launch %.2f us; per warp: load phase %.0f cycles, red phase %.0f cycles (H = %d rounds of 32 lanes)\n", mode,
               mode == 1 ? "loads" : mode == 2 ? "reds" : mode == 3 ? "loads then reds" : "reds to private addresses", best * 1e3, a / NW, b / NW, H);
    }
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
