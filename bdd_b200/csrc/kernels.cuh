// bdd_b200/csrc/kernels.cuh -- sm_100a kernels of the deferred min-marginal-averaging sweep.
//
// Replaces (does not port) the reference's per-hop kernels
//   min_marginals_from_directional_costs_cuda   bdd_cuda_parallel_mma.cu:59-86
//   compute_mm_diff_flush_mm_lo                 bdd_cuda_parallel_mma.cu:29-42
//   forward_step_with_solve / backward_step_with_solve   :164-205 / :259-299
//   compute_delta_atomic, normalize_delta_st    :358-376, :410-430
//   forward_step / backward_step(_with_path_costs)       bdd_cuda_base.cu:560-667
// which are launched 3x per hop (6H+10 launches per iteration) and use CAS-loop atomicMin on
// global memory for every arc.
//
// Design: ONE launch per pass.  A warp owns a bundle of 32/P BDDs and walks it hop by hop
// (see layout.hpp).  The chain-dependent frontier (cost_from_root going forward,
// cost_from_terminal going backward) lives in shared memory; with P == 1 a BDD never leaves
// its lane, so the walk needs no synchronisation at all, with P > 1 one __syncwarp per hop.
// Per-layer min-marginals are lane-local minima followed by log2(P) xor-shuffles; the
// cross-BDD sum of min-marginal differences is a red.global.add per layer (or, in
// deterministic mode, a fixed-order segmented sum in delta_segsum_kernel); the division by
// the number of BDDs per variable is folded into the read of delta_in; the buffer of the
// pass after next is zeroed by the same launch.  Everything else is coalesced streaming.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstdint>

#include "layout.hpp"

namespace bddb200 {

enum SweepMode { MODE_MMA = 0, MODE_PLAIN = 1, MODE_MM = 2 };

template<typename REAL> struct real2;
template<> struct real2<float> { using type = float2; };
template<> struct real2<double> { using type = double2; };

template<typename REAL> __device__ __forceinline__ REAL real_inf();
template<> __device__ __forceinline__ float real_inf<float>() { return CUDART_INF_F; }
template<> __device__ __forceinline__ double real_inf<double>() { return CUDART_INF; }

// shared-memory atomic min for floating point through the order-preserving integer view
// (non-negative values compare like signed ints, negative ones like reversed unsigned ints).
// v + 0 turns -0.0 into +0.0 so that it cannot win against negative numbers.
__device__ __forceinline__ void smem_atomic_min(float* addr, float v)
{
    v += 0.0f;
    if(v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void smem_atomic_min(double* addr, double v)
{
    v += 0.0;
    if(v >= 0.0) atomicMin(reinterpret_cast<long long*>(addr), __double_as_longlong(v));
    else atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

template<typename REAL>
struct SweepArgs {
    const BundleDesc* bundles;
    const HopRec* hops;
    const uint32_t* topo;
    const int2* lay_vn;        // per layer entry {variable or -1, nr_bdds(variable)}
    const int32_t* bundle_bdd;
    REAL* cfr;                 // cost from root, per slot
    REAL* cft;                 // cost from terminal, per slot
    const REAL* lo_in;
    const REAL* hi_in;
    REAL* lo_out;
    REAL* hi_out;
    REAL* mmd;                 // deferred min-marginal difference per layer entry
    const REAL* delta_in;      // 2V
    REAL* delta_out;           // 2V (zeroed before the launch or by the previous pass)
    REAL* zero_buf;            // 2V buffer to clear for the pass after next (may be null)
    REAL* mm_lo_out;           // MODE_MM
    REAL* mm_hi_out;
    REAL* bdd_lb;              // backward: cost_from_terminal of every BDD's root
    REAL omega;
    uint32_t n_zero;
    uint32_t bundle_first, bundle_count;
    uint32_t tile_slots;       // capacity of one shared-memory frontier buffer, in slots
    int normalize_in;          // divide delta_in by nr_bdds(var) while reading
    int accumulate;            // add |mm_diff| to delta_out with atomics
};

template<int P, typename REAL>
__device__ __forceinline__ REAL group_min(REAL v)
{
#pragma unroll
    for(int o = P / 2; o > 0; o >>= 1)
    {
        const REAL w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}

// Layer record + the damped min-marginal update shared by both directions.
//   mm_diff = omega * (mm_hi - mm_lo), 0 if either is infinite  (bdd_cuda_parallel_mma.cu:29-42)
//   lo' = lo + min(mm_diff, 0) + delta[2v];  hi' = hi + min(-mm_diff, 0) + delta[2v+1]   (:185-193, :280-281)
template<typename REAL>
struct LayerState {
    int var;
    REAL lo_c, hi_c, d0, d1;
};

template<typename REAL, int MODE>
__device__ __forceinline__ LayerState<REAL> load_layer(const SweepArgs<REAL>& a, uint32_t lay)
{
    LayerState<REAL> s;
    const int2 vn = __ldg(a.lay_vn + lay);
    s.var = vn.x;
    s.lo_c = 0; s.hi_c = 0; s.d0 = 0; s.d1 = 0;
    if(s.var >= 0)
    {
        s.lo_c = a.lo_in[lay];
        s.hi_c = a.hi_in[lay];
        if(MODE == MODE_MMA)
        {
            using R2 = typename real2<REAL>::type;
            const R2 d = *reinterpret_cast<const R2*>(a.delta_in + 2 * (size_t)s.var);
            s.d0 = d.x; s.d1 = d.y;
            if(a.normalize_in)
            {
                const REAL n = (REAL)vn.y;
                s.d0 /= n; s.d1 /= n;
            }
        }
    }
    return s;
}

template<typename REAL>
__device__ __forceinline__ void store_layer(const SweepArgs<REAL>& a, uint32_t lay, int var, REAL lo_n, REAL hi_n, REAL diff)
{
    a.lo_out[lay] = lo_n;
    a.hi_out[lay] = hi_n;
    a.mmd[lay] = diff;
    if(a.accumulate)
    {
        // compute_delta_atomic, bdd_cuda_parallel_mma.cu:358-376
        if(diff > 0) atomicAdd(a.delta_out + 2 * (size_t)var + 1, diff);
        else if(diff < 0) atomicAdd(a.delta_out + 2 * (size_t)var, -diff);
    }
}

// ------------------------------------------------------------------ forward ------------
// forward_mm (bdd_cuda_parallel_mma.cu:207-257) for MODE_MMA, forward_run
// (bdd_cuda_base.cu:588-612) for MODE_PLAIN.
template<typename REAL, int LOGP, int MODE>
__device__ __forceinline__ void sweep_forward(const SweepArgs<REAL>& a, const BundleDesc& bd, REAL* tiles, const int lane)
{
    constexpr int P = 1 << LOGP;
    constexpr int BPW = 32 >> LOGP;
    const int bl = lane >> LOGP;
    const int p = lane & (P - 1);
    const REAL INF = real_inf<REAL>();
    REAL* cur = tiles;
    REAL* nxt = tiles + a.tile_slots;
    REAL* spare = tiles + 2 * (size_t)a.tile_slots;
    const HopRec* hops = a.hops + bd.hop_base;
    const uint32_t n_hops = bd.n_hops;

    HopRec h = hops[0];
    HopRec hn = n_hops > 1 ? hops[1] : HopRec{0u, 0u};
    for(uint32_t j = 0; j < h.J; ++j) cur[j * 32 + lane] = INF;
    if(p == 0 && a.topo[h.node_off + lane] != TOPO_PAD) cur[lane] = 0;   // flush_costs_from_root, bdd_cuda_base.cu:1438-1445
    for(uint32_t j = 0; j < hn.J; ++j) nxt[j * 32 + lane] = INF;
    if(P > 1) __syncwarp();

    for(uint32_t k = 0; k < n_hops; ++k)
    {
        const HopRec hnn = (k + 2 < n_hops) ? hops[k + 2] : HopRec{0u, 0u};
        for(uint32_t j = 0; j < hnn.J; ++j) spare[j * 32 + lane] = INF;

        const uint32_t lay = bd.layer_base + k * BPW + bl;
        const LayerState<REAL> ls = load_layer<REAL, MODE>(a, lay);
        REAL lo_n = ls.lo_c, hi_n = ls.hi_c, diff = 0;
        if(MODE == MODE_MMA)
        {
            REAL mm0 = INF, mm1 = INF;
            for(uint32_t j = 0; j < h.J; ++j)
            {
                const uint32_t t = __ldg(a.topo + h.node_off + j * 32 + lane);
                if(t < TOPO_TOP)
                {
                    const REAL c = cur[j * 32 + lane];
                    const uint32_t lo = t & 0xFFFFu, hi = t >> 16;
                    const REAL ta = lo == CHILD_BOT ? INF : a.cft[hn.node_off + lo];
                    const REAL tb = hi == CHILD_BOT ? INF : a.cft[hn.node_off + hi];
                    const REAL m0 = c + ls.lo_c + ta;     // same association as bdd_cuda_parallel_mma.cu:83-84
                    const REAL m1 = c + ls.hi_c + tb;
                    mm0 = m0 < mm0 ? m0 : mm0;
                    mm1 = m1 < mm1 ? m1 : mm1;
                }
            }
            mm0 = group_min<P>(mm0);
            mm1 = group_min<P>(mm1);
            if(isfinite(mm0) && isfinite(mm1)) diff = a.omega * (mm1 - mm0);
            lo_n = ls.lo_c + (diff < 0 ? diff : (REAL)0) + ls.d0;
            hi_n = ls.hi_c + (-diff < 0 ? -diff : (REAL)0) + ls.d1;
        }
        for(uint32_t j = 0; j < h.J; ++j)
        {
            const uint32_t s = h.node_off + j * 32 + lane;
            const uint32_t t = __ldg(a.topo + s);
            const REAL c = cur[j * 32 + lane];
            a.cfr[s] = c;
            if(t < TOPO_TOP)
            {
                const uint32_t lo = t & 0xFFFFu, hi = t >> 16;
                if(lo != CHILD_BOT)
                {
                    const REAL v = c + lo_n;
                    if(P == 1) { if(v < nxt[lo]) nxt[lo] = v; } else smem_atomic_min(nxt + lo, v);
                }
                if(hi != CHILD_BOT)
                {
                    const REAL v = c + hi_n;
                    if(P == 1) { if(v < nxt[hi]) nxt[hi] = v; } else smem_atomic_min(nxt + hi, v);
                }
            }
        }
        if(MODE == MODE_MMA && p == 0 && ls.var >= 0)
            store_layer(a, lay, ls.var, lo_n, hi_n, diff);
        if(P > 1) __syncwarp();
        REAL* tmp = cur; cur = nxt; nxt = spare; spare = tmp;
        h = hn; hn = hnn;
    }
}

// ------------------------------------------------------------------ backward -----------
// backward_mm (bdd_cuda_parallel_mma.cu:301-346) for MODE_MMA, backward_run(false)
// (bdd_cuda_base.cu:670-713) for MODE_PLAIN, backward_run(true) + the per-layer min
// reduction of min_marginals_cuda (bdd_cuda_base.cu:716-736) for MODE_MM.
template<typename REAL, int LOGP, int MODE>
__device__ __forceinline__ void sweep_backward(const SweepArgs<REAL>& a, const BundleDesc& bd, REAL* tiles, const int lane)
{
    constexpr int P = 1 << LOGP;
    constexpr int BPW = 32 >> LOGP;
    const int bl = lane >> LOGP;
    const int p = lane & (P - 1);
    const REAL INF = real_inf<REAL>();
    REAL* cur = tiles;
    REAL* nxt = tiles + a.tile_slots;
    const HopRec* hops = a.hops + bd.hop_base;
    const uint32_t n_hops = bd.n_hops;

    HopRec hn = HopRec{0u, 0u};
    for(int k = (int)n_hops - 1; k >= 0; --k)
    {
        const HopRec h = hops[k];
        const uint32_t lay = bd.layer_base + (uint32_t)k * BPW + bl;
        const LayerState<REAL> ls = load_layer<REAL, MODE>(a, lay);
        REAL lo_n = ls.lo_c, hi_n = ls.hi_c, diff = 0;
        if(MODE != MODE_PLAIN)
        {
            REAL mm0 = INF, mm1 = INF;
            for(uint32_t j = 0; j < h.J; ++j)
            {
                const uint32_t s = h.node_off + j * 32 + lane;
                const uint32_t t = __ldg(a.topo + s);
                if(t < TOPO_TOP)
                {
                    const REAL c = a.cfr[s];
                    const uint32_t lo = t & 0xFFFFu, hi = t >> 16;
                    const REAL ta = lo == CHILD_BOT ? INF : nxt[lo];
                    const REAL tb = hi == CHILD_BOT ? INF : nxt[hi];
                    REAL m0, m1;
                    if(MODE == MODE_MMA) { m0 = c + ls.lo_c + ta; m1 = c + ls.hi_c + tb; }
                    else { m0 = c + (ta + ls.lo_c); m1 = c + (tb + ls.hi_c); }   // path costs, bdd_cuda_base.cu:636-641
                    mm0 = m0 < mm0 ? m0 : mm0;
                    mm1 = m1 < mm1 ? m1 : mm1;
                }
            }
            mm0 = group_min<P>(mm0);
            mm1 = group_min<P>(mm1);
            if(MODE == MODE_MMA)
            {
                if(isfinite(mm0) && isfinite(mm1)) diff = a.omega * (mm1 - mm0);
                lo_n = ls.lo_c + (diff < 0 ? diff : (REAL)0) + ls.d0;
                hi_n = ls.hi_c + (-diff < 0 ? -diff : (REAL)0) + ls.d1;
            }
            else if(p == 0 && ls.var >= 0)
            {
                a.mm_lo_out[lay] = mm0;
                a.mm_hi_out[lay] = mm1;
            }
        }
        for(uint32_t j = 0; j < h.J; ++j)
        {
            const uint32_t s = h.node_off + j * 32 + lane;
            const uint32_t t = __ldg(a.topo + s);
            REAL val = t == TOPO_TOP ? (REAL)0 : INF;          // set_special_nodes_costs, bdd_cuda_base.cu:217-227
            if(t < TOPO_TOP)
            {
                const uint32_t lo = t & 0xFFFFu, hi = t >> 16;
                const REAL ta = lo == CHILD_BOT ? INF : nxt[lo];
                const REAL tb = hi == CHILD_BOT ? INF : nxt[hi];
                const REAL vh = hi_n + tb, vl = lo_n + ta;        // bdd_cuda_parallel_mma.cu:286
                val = vh < vl ? vh : vl;
            }
            cur[j * 32 + lane] = val;
            a.cft[s] = val;
        }
        if(MODE == MODE_MMA && p == 0 && ls.var >= 0)
            store_layer(a, lay, ls.var, lo_n, hi_n, diff);
        if(P > 1) __syncwarp();
        REAL* tmp = cur; cur = nxt; nxt = tmp;
        hn = h;
    }
    (void)hn;
    if(p == 0)
    {
        const int32_t bi = a.bundle_bdd[bd.bdd_base + bl];
        if(bi >= 0) a.bdd_lb[bi] = nxt[lane];   // root = node 0 of hop 0
    }
}

template<typename REAL, int MODE, bool FORWARD>
__global__ void __launch_bounds__(256) sweep_kernel(const SweepArgs<REAL> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    if(MODE == MODE_MMA && a.zero_buf != nullptr)
        for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_zero; i += gridDim.x * blockDim.x)
            a.zero_buf[i] = 0;
    uint32_t g = blockIdx.x * wpc + warp;
    if(g >= a.bundle_count) return;
    g += a.bundle_first;
    constexpr int NBUF = FORWARD ? 3 : 2;
    REAL* tiles = reinterpret_cast<REAL*>(smem_raw) + (size_t)warp * NBUF * a.tile_slots;
    const BundleDesc bd = a.bundles[g];
#define BDDB200_DISPATCH(LP) \
    case LP: if(FORWARD) sweep_forward<REAL, LP, (MODE == MODE_MM ? MODE_PLAIN : MODE)>(a, bd, tiles, lane); \
             else sweep_backward<REAL, LP, MODE>(a, bd, tiles, lane); break;
    switch(bd.logP)
    {
        BDDB200_DISPATCH(0) BDDB200_DISPATCH(1) BDDB200_DISPATCH(2)
        BDDB200_DISPATCH(3) BDDB200_DISPATCH(4) BDDB200_DISPATCH(5)
        default: break;
    }
#undef BDDB200_DISPATCH
}

// ------------------------------------------------------------------ small kernels ------

// deterministic replacement of compute_delta (bdd_cuda_parallel_mma.cu:379-393): per variable,
// sum its layers' mm differences in BDD order (the order the single-threaded CPU solver uses).
template<typename REAL>
__global__ void delta_segsum_kernel(const uint32_t* __restrict__ var_lay_begin, const uint32_t* __restrict__ var_lay,
                                    const REAL* __restrict__ mmd, REAL* __restrict__ delta_out, uint32_t n_vars)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if(v >= n_vars) return;
    REAL lo = 0, hi = 0;
    for(uint32_t e = var_lay_begin[v]; e < var_lay_begin[v + 1]; ++e)
    {
        const REAL d = mmd[var_lay[e]];
        if(d > 0) hi += d; else if(d < 0) lo += -d;
    }
    delta_out[2 * (size_t)v] = lo;
    delta_out[2 * (size_t)v + 1] = hi;
}

// normalize_delta_st, bdd_cuda_parallel_mma.cu:410-419 (guarded for variables without BDD)
template<typename REAL>
__global__ void normalize_kernel(const REAL* __restrict__ in, REAL* __restrict__ out, const int32_t* __restrict__ nr_bdds, uint32_t n2)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n2) return;
    const int n = nr_bdds[i >> 1];
    out[i] = n > 0 ? in[i] / (REAL)n : in[i];
}

// set_vars_costs_func, bdd_cuda_base.cu:454-474
template<typename REAL>
__global__ void update_costs_kernel(const int2* __restrict__ lay_vn, REAL* __restrict__ cost, const REAL* __restrict__ c, uint32_t n_c, uint32_t n_lay)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_lay) return;
    const int2 vn = lay_vn[i];
    if(vn.x < 0) return;
    if((uint32_t)vn.x >= n_c) { cost[i] = 0; return; }
    cost[i] += c[vn.x] / (REAL)vn.y;
}

// set_var_cost_func, bdd_cuda_base.cu:425-452
template<typename REAL>
__global__ void set_cost_kernel(const int2* __restrict__ lay_vn, REAL* __restrict__ hi, int var, REAL add, uint32_t n_lay)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n_lay && lay_vn[i].x == var) hi[i] += add;
}

// distribute_deffered_mm_diff_func, bdd_cuda_base.cu:1396-1414
template<typename REAL>
__global__ void distribute_kernel(const int2* __restrict__ lay_vn, REAL* __restrict__ lo, REAL* __restrict__ hi, REAL* __restrict__ mmd, uint32_t n_lay)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_lay || lay_vn[i].x < 0) return;
    const REAL d = mmd[i];
    if(d > 0) hi[i] += d; else lo[i] -= d;
    mmd[i] = 0;
}

// out[e] = src[ext2lay[e]] for inner layers, `fill` for terminal layers
template<typename T>
__global__ void gather_ext_kernel(const uint32_t* __restrict__ ext2lay, const int32_t* __restrict__ ext_var, const T* __restrict__ src, T* __restrict__ out, T fill, uint32_t n_ext)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n_ext) return;
    out[e] = ext_var[e] == INT_MAX ? fill : src[ext2lay[e]];
}

template<typename T>
__global__ void scatter_ext_kernel(const uint32_t* __restrict__ ext2lay, const int32_t* __restrict__ ext_var, const T* __restrict__ in, T* __restrict__ dst, uint32_t n_ext)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n_ext || ext_var[e] == INT_MAX) return;
    dst[ext2lay[e]] = in[e];
}

// out[i] = src[perm[i]]
template<typename T>
__global__ void permute_kernel(const uint32_t* __restrict__ perm, const T* __restrict__ src, T* __restrict__ out, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) out[i] = src[perm[i]];
}

// compute_net_costs_func, bdd_cuda_parallel_mma.cu:432-447 (layer order, terminals 0)
template<typename REAL>
__global__ void net_costs_kernel(const uint32_t* __restrict__ ext2lay, const int32_t* __restrict__ ext_var, const REAL* __restrict__ lo, const REAL* __restrict__ hi,
                                 const REAL* __restrict__ mmd, REAL* __restrict__ out, uint32_t n_ext)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n_ext) return;
    if(ext_var[e] == INT_MAX) { out[e] = 0; return; }
    const uint32_t l = ext2lay[e];
    out[e] = hi[l] - lo[l] + mmd[l];
}

// add_scaled_product_func, bdd_cuda_parallel_mma.h:53-60
template<typename REAL>
__global__ void gradient_step_kernel(const uint32_t* __restrict__ ext2lay, const int32_t* __restrict__ ext_var, REAL* __restrict__ hi, const REAL* __restrict__ g, REAL step, uint32_t n_ext)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n_ext || ext_var[e] == INT_MAX) return;
    const uint32_t l = ext2lay[e];
    hi[l] = hi[l] + step * g[e];
}

// make_dual_feasible, bdd_cuda_base.cu:1261-1303: subtract the per-variable mean; terminals 0.
template<typename REAL>
__global__ void dual_feasible_kernel(const uint32_t* __restrict__ var_lay_begin, const uint32_t* __restrict__ sorted_ext, const int32_t* __restrict__ nr_bdds,
                                     REAL* __restrict__ d, uint32_t n_vars)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if(v >= n_vars) return;
    const uint32_t b = var_lay_begin[v], e = var_lay_begin[v + 1];
    if(b == e) return;
    REAL sum = 0;
    for(uint32_t i = b; i < e; ++i) sum += d[sorted_ext[i]];
    const REAL mean = sum / (REAL)nr_bdds[v];
    for(uint32_t i = b; i < e; ++i) d[sorted_ext[i]] -= mean;
}

template<typename REAL>
__global__ void zero_terminals_kernel(const int32_t* __restrict__ ext_var, REAL* __restrict__ d, uint32_t n_ext)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e < n_ext && ext_var[e] == INT_MAX) d[e] = 0;
}

// compute_primal_objective_vec, bdd_cuda_base.cu:1352-1362
template<typename REAL>
__global__ void primal_objective_kernel(const uint32_t* __restrict__ var_lay_begin, const uint32_t* __restrict__ var_lay, const REAL* __restrict__ lo, const REAL* __restrict__ hi,
                                        double* __restrict__ out, uint32_t n_vars)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if(v >= n_vars) return;
    REAL s = 0;
    for(uint32_t i = var_lay_begin[v]; i < var_lay_begin[v + 1]; ++i) s += hi[var_lay[i]] - lo[var_lay[i]];
    out[v] = (double)s;
}

// lower_bound, bdd_cuda_base.cu:1243-1251: sum of the roots' cost_from_terminal in double,
// as a fixed-shape two-stage tree (bit-reproducible).
template<typename REAL>
__global__ void lb_partial_kernel(const REAL* __restrict__ bdd_lb, double* __restrict__ partial, uint32_t n_bdds)
{
    __shared__ double sh[256];
    const uint32_t chunk = (n_bdds + gridDim.x - 1) / gridDim.x;
    const uint32_t b0 = blockIdx.x * chunk, b1 = min(n_bdds, b0 + chunk);
    double s = 0;
    for(uint32_t b = b0 + threadIdx.x; b < b1; b += blockDim.x) s += (double)bdd_lb[b];
    sh[threadIdx.x] = s;
    __syncthreads();
    for(int o = 128; o > 0; o >>= 1)
    {
        if((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if(threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void lb_final_kernel(const double* __restrict__ partial, double* __restrict__ out, uint32_t n)
{
    __shared__ double sh[256];
    double s = 0;
    for(uint32_t i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for(int o = 128; o > 0; o >>= 1)
    {
        if((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if(threadIdx.x == 0) out[0] = sh[0];
}

// compute_bdd_sol_func, bdd_cuda_base.cu:1103-1135: per BDD follow the cheaper arc from the
// root (tie -> hi), comparing hi_path - lo_path > 0 with path = cfr + (cft[child] + cost).
template<typename REAL>
__global__ void bdds_solution_kernel(const BundleDesc* __restrict__ bundles, const HopRec* __restrict__ hops, const uint32_t* __restrict__ topo,
                                     const uint32_t* __restrict__ bdd_bundle, const uint32_t* __restrict__ bdd_ext_begin,
                                     const REAL* __restrict__ cfr, const REAL* __restrict__ cft, const REAL* __restrict__ lo_c, const REAL* __restrict__ hi_c,
                                     char* __restrict__ sol, uint32_t n_bdds)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if(b >= n_bdds) return;
    const uint32_t g = bdd_bundle[2 * b], q = bdd_bundle[2 * b + 1];
    const BundleDesc bd = bundles[g];
    const uint32_t bpw = 32u >> bd.logP;
    const uint32_t e0 = bdd_ext_begin[b], n_lay = bdd_ext_begin[b + 1] - e0 - 1;
    uint32_t ts = q << bd.logP;   // root's slot inside the hop-0 tile
    const REAL INF = real_inf<REAL>();
    for(uint32_t k = 0; k < n_lay; ++k)
    {
        const HopRec h = hops[bd.hop_base + k], hn = hops[bd.hop_base + k + 1];
        const uint32_t s = h.node_off + ts;
        const uint32_t t = topo[s];
        const uint32_t lay = bd.layer_base + k * bpw + q;
        const uint32_t lo = t & 0xFFFFu, hi = t >> 16;
        const REAL c = cfr[s];
        const REAL lo_path = c + ((lo == CHILD_BOT ? INF : cft[hn.node_off + lo]) + lo_c[lay]);
        const REAL hi_path = c + ((hi == CHILD_BOT ? INF : cft[hn.node_off + hi]) + hi_c[lay]);
        const bool take_lo = (hi_path - lo_path > 0);
        sol[e0 + k] = take_lo ? 0 : 1;
        ts = take_lo ? lo : hi;
    }
    sol[e0 + n_lay] = 0;   // terminal layer
}

} // namespace bddb200
