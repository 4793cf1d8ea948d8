// Micro-benchmark: warp-blocking issue cost (clock64 deltas, one warp per SM) of the instructions the lane kernel's
// staging uses: UBLKCP (cp.async.bulk), scattered / coalesced LDGSTS (cp.async), scattered LDG, mbarrier expect_tx.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_cost issue_cost.cu && ./issue_cost
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void bench(const float* src, const int* idx, long long* out, float* sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
    unsigned char* buf = sm + 128;
    const int lane = threadIdx.x & 31;
    if(lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar))); asm volatile("fence.proxy.async.shared::cta;"); }
    __syncwarp();
    long long t[12];
    const float* base = src + (size_t)blockIdx.x * 65536;
    t[0] = clock64();
    // (a) 4 bulk copies of 1 KiB issued by lane 0 sequentially
    if(lane == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(4096));
        for(int k = 0; k < 4; ++k)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(buf + k * 1024)), "l"(__cvta_generic_to_global(base + k * 4096)), "r"(1024), "r"(smem_u32(bar)) : "memory");
    }
    __syncwarp();
    t[1] = clock64();
    // wait for them
    { uint32_t ok = 0; while(!ok) asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory"); }
    t[2] = clock64();
    // (b) 16 scattered 8-byte cp.async per lane
    int id[16];
#pragma unroll
    for(int k = 0; k < 16; ++k) id[k] = idx[(blockIdx.x * 16 + k) * 32 + lane];
    t[3] = clock64();
#pragma unroll
    for(int k = 0; k < 16; ++k)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(buf + 8192 + (k * 32 + lane) * 8)), "l"(__cvta_generic_to_global(src + 2 * (size_t)id[k])) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    t[4] = clock64();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    t[5] = clock64();
    // (c) 16 coalesced 8-byte cp.async per lane
#pragma unroll
    for(int k = 0; k < 16; ++k)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(buf + 8192 + (k * 32 + lane) * 8)), "l"(__cvta_generic_to_global(base + 32768 + (k * 32 + lane) * 2)) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    t[6] = clock64();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    t[7] = clock64();
    // (d) 16 scattered 8-byte LDG per lane (different indices)
    float2 v[16];
#pragma unroll
    for(int k = 0; k < 16; ++k) v[k] = *reinterpret_cast<const float2*>(src + 2 * (size_t)(id[k] ^ 0x5555));
    t[8] = clock64();
    float acc = 0;
#pragma unroll
    for(int k = 0; k < 16; ++k) acc += v[k].x + v[k].y;
    t[9] = clock64();
    // (e) one bulk copy of 4 KiB by lane 0
    if(lane == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(4096));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(buf)), "l"(__cvta_generic_to_global(base + 16384)), "r"(4096), "r"(smem_u32(bar)) : "memory");
    }
    __syncwarp();
    t[10] = clock64();
    { uint32_t ok = 0; while(!ok) asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(1) : "memory"); }
    t[11] = clock64();
    // (f) 16 scattered 16-byte cp.async.cg per lane (L1 bypass)
    long long u[6];
    u[0] = clock64();
#pragma unroll
    for(int k = 0; k < 16; ++k)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(buf + 16384 + (k * 32 + lane) * 16)), "l"(__cvta_generic_to_global(src + 4 * (size_t)((id[k] ^ 0x3333) >> 1))) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    u[1] = clock64();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    u[2] = clock64();
    // (g) 16 scattered 4-byte cp.async.ca per lane
#pragma unroll
    for(int k = 0; k < 16; ++k)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(buf + 8192 + (k * 32 + lane) * 4)), "l"(__cvta_generic_to_global(src + (size_t)(id[k] ^ 0x1111))) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    u[3] = clock64();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    u[4] = clock64();
    // (h) 4 bulk copies of 1 KiB by lane 0, each followed by ~300 cycles of dependent ALU work; (i) the ALU work alone
    long long w[4];
    float z = acc;
    w[0] = clock64();
    if(lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(4096));
    for(int k = 0; k < 4; ++k)
    {
        if(lane == 0)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(buf + k * 1024)), "l"(__cvta_generic_to_global(base + 20480 + k * 4096)), "r"(1024), "r"(smem_u32(bar)) : "memory");
#pragma unroll 1
        for(int j = 0; j < 75; ++j) z = z * 1.0001f + 0.5f;
    }
    w[1] = clock64();
#pragma unroll 1
    for(int k = 0; k < 4; ++k)
    {
#pragma unroll 1
        for(int j = 0; j < 75; ++j) z = z * 1.0001f + 0.5f;
    }
    w[2] = clock64();
    { uint32_t ok = 0; while(!ok) asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory"); }
    acc += z;
    if(lane == 0) { out[148 * 17 + blockIdx.x * 2] = w[1] - w[0]; out[148 * 17 + blockIdx.x * 2 + 1] = w[2] - w[1]; }
    if(lane == 0) for(int k = 0; k < 12; ++k) out[blockIdx.x * 12 + k] = t[k] - t[0];
    if(lane == 0) for(int k = 0; k < 5; ++k) out[148 * 12 + blockIdx.x * 5 + k] = u[k] - u[0];
    sink[blockIdx.x * 32 + lane] = acc + reinterpret_cast<float*>(buf)[lane] + reinterpret_cast<float*>(buf + 8192)[lane];
}

int main()
{
    const int B = 148;
    float* src; int* idx; long long* out; float* sink;
    cudaMalloc(&src, (size_t)B * 65536 * 4 + (1 << 24)); cudaMemset(src, 0, (size_t)B * 65536 * 4 + (1 << 24));
    cudaMalloc(&idx, B * 16 * 32 * 4); cudaMalloc(&out, B * 19 * 8); cudaMalloc(&sink, B * 32 * 4);
    int* h = new int[B * 16 * 32];
    uint32_t s = 12345;
    for(int i = 0; i < B * 16 * 32; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % 1000000; }
    cudaMemcpy(idx, h, B * 16 * 32 * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for(int rep = 0; rep < 3; ++rep)
    {
        bench<<<B, 32, 65536>>>(src, idx, out, sink);
        cudaDeviceSynchronize();
        long long ho[B * 12];
        cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
        double m[12] = {0};
        for(int b = 0; b < B; ++b) for(int k = 0; k < 12; ++k) m[k] += (double)ho[b * 12 + k] / B;
        printf("rep %d (mean cycles over %d SMs)\n", rep, B);
        printf("  4 x UBLKCP 1KiB + expect_tx issue: %.0f   land: %.0f\n", m[1] - m[0], m[2] - m[1]);
        printf("  16 idx loads (coalesced LDG, dependent): %.0f\n", m[3] - m[2]);
        printf("  16 x scattered LDGSTS.64 issue: %.0f   wait: %.0f\n", m[4] - m[3], m[5] - m[4]);
        printf("  16 x coalesced LDGSTS.64 issue: %.0f   wait: %.0f\n", m[6] - m[5], m[7] - m[6]);
        printf("  16 x scattered LDG.64 issue: %.0f   use: %.0f\n", m[8] - m[7], m[9] - m[8]);
        printf("  1 x UBLKCP 4KiB + expect_tx issue: %.0f   land: %.0f\n", m[10] - m[9], m[11] - m[10]);
        long long hu[B * 5];
        cudaMemcpy(hu, out + B * 12, sizeof(hu), cudaMemcpyDeviceToHost);
        double mu[5] = {0};
        for(int b = 0; b < B; ++b) for(int k = 0; k < 5; ++k) mu[k] += (double)hu[b * 5 + k] / B;
        printf("  16 x scattered LDGSTS.128 .cg issue: %.0f   wait: %.0f\n", mu[1] - mu[0], mu[2] - mu[1]);
        printf("  16 x scattered LDGSTS.32 .ca issue: %.0f   wait: %.0f\n", mu[3] - mu[2], mu[4] - mu[3]);
        long long hw[B * 2];
        cudaMemcpy(hw, out + B * 17, sizeof(hw), cudaMemcpyDeviceToHost);
        double mw[2] = {0};
        for(int b = 0; b < B; ++b) for(int k = 0; k < 2; ++k) mw[k] += (double)hw[b * 2 + k] / B;
        printf("  4 x (UBLKCP 1KiB + ALU work): %.0f   the ALU work alone: %.0f   => copies cost %.0f\n", mw[0], mw[1], mw[0] - mw[1]);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
