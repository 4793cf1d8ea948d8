// bdd_b200/csrc/resident.cuh -- the on-chip ("resident") form of the deferred min-marginal-averaging sweep.
//
// For collections of at most one wave of lane-class bundles (layout.hpp, CLS_LANE: one lane per BDD) the whole per-BDD state of
// a bundle -- topology words, {variable, nr_bdds}, {lo, hi} arc costs, cost_from_root and cost_from_terminal -- fits the shared
// memory of the SM whose warp owns the bundle.  ONE cooperative launch then runs any number of
//   iteration() = forward_mm, normalize_delta, backward_mm, normalize_delta      (bdd_cuda_parallel_mma.cu:142-153)
// with the state staying on chip between the passes: it is read from HBM once when the launch starts (bulk-async copies, one
// mbarrier per warp) and written back once when it ends (bulk-async stores).
//
// Cross-BDD averaging without atomics and without grid barriers.  What couples the BDDs is, per variable v and pass p, the sum of
// the min-marginal differences of the layers of v (compute_delta_atomic :358-376, normalize_delta :410-430).  Measured on B200
// (tools/microbench/scatter_cost2.cu, barrier_cost.cu) a scattered red.global.add costs 2.5 cycles per lane per SM, a scattered
// load 1.2-1.3, a grid barrier 3 000-4 000 cycles -- together more than the min-plus arithmetic of a pass.  So the exchange is
// a data-flow protocol of plain stores and loads, every datum travelling with the number of the pass that produced it:
//   * the hop loop of pass p writes, per layer entry, the record {mm_diff, p} (one coalesced 8-byte store per lane);
//   * every variable has an owner warp; when that warp starts pass p + 1 it loads the records of the variable's layers (waiting
//     until each carries p), adds them in BDD order, divides by nr_bdds(v) and publishes {lo, p, hi, p};
//   * every warp then loads the published sums of the variables of its own layers (waiting for p) and runs its hop loop.
// A record is written and read as naturally aligned 8-byte {value, pass} pairs, so value and pass number arrive together (the
// scheme of NCCL's LL protocol); no fence, no atomic, no barrier is involved, warps run ahead of each other by at most one pass,
// and the sums are added in a fixed order with an exact division: the result is bit-reproducible and equals the single-threaded
// CPU solver's.  Buffers are single: a record is only overwritten after everyone who needed the old value has produced the data
// that the overwrite itself waits for (contributions of pass p + 1 need the sums of pass p, which need the contributions of p).
// Cooperative launch is used only for its guarantee that all CTAs are co-resident (the warps wait for each other's data).
//
// The hop arithmetic is the streaming lane kernel's (sweep_lane_bundle, deterministic form) operation for operation.
#pragma once

#include "kernels.cuh"

namespace bddb200 {

constexpr int RES_TRACE_EVENTS = 16;
constexpr uint32_t RES_MAX_SPINS = 1u << 22;       // x (40 ns sleep + one L2 round trip): seconds
constexpr uint32_t VN_N_NONE = 0u, VN_N_TOP = 0xFFFFFFFFu;      // shared-memory copy of lay_vn.y for entries without a variable

// ---- exchange records --------------------------------------------------------------------------------------------------
// float : contribution {diff, pass} 8 B,                               sum {lo, pass, hi, pass} 16 B
// double: contribution {diff.lo32, pass, diff.hi32, pass} 16 B,        sum {lo.lo32, pass, lo.hi32, pass, hi.lo32, pass, hi.hi32, pass} 32 B
template<typename REAL> struct ExchangeRec;
template<> struct ExchangeRec<float> { static constexpr uint32_t CONTRIB = 8, SUM = 16; };
template<> struct ExchangeRec<double> { static constexpr uint32_t CONTRIB = 16, SUM = 32; };

__device__ __forceinline__ void st_contrib(unsigned char* p, float d, uint32_t pass)
{
    asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" :: "l"(p), "r"(__float_as_uint(d)), "r"(pass) : "memory");
}
__device__ __forceinline__ void st_contrib(unsigned char* p, double d, uint32_t pass)
{
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"((uint32_t)u), "r"(pass), "r"((uint32_t)(u >> 32)), "r"(pass) : "memory");
}
// loads bypass L1 (the producer is another SM); false = the record is not of pass `pass` yet
__device__ __forceinline__ bool ld_contrib(const unsigned char* p, uint32_t pass, float& d)
{
    uint32_t x, f;
    asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(f) : "l"(p) : "memory");
    d = __uint_as_float(x);
    return f == pass;
}
__device__ __forceinline__ bool ld_contrib(const unsigned char* p, uint32_t pass, double& d)
{
    uint32_t x0, f0, x1, f1;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(f0), "=r"(x1), "=r"(f1) : "l"(p) : "memory");
    d = __longlong_as_double((long long)(((unsigned long long)x1 << 32) | x0));
    return f0 == pass && f1 == pass;
}
__device__ __forceinline__ void st_sum(unsigned char* p, float lo, float hi, uint32_t pass)
{
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(__float_as_uint(lo)), "r"(pass), "r"(__float_as_uint(hi)), "r"(pass) : "memory");
}
__device__ __forceinline__ void st_sum(unsigned char* p, double lo, double hi, uint32_t pass)
{
    const unsigned long long a = (unsigned long long)__double_as_longlong(lo), b = (unsigned long long)__double_as_longlong(hi);
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"((uint32_t)a), "r"(pass), "r"((uint32_t)(a >> 32)), "r"(pass) : "memory");
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" :: "l"(p + 16), "r"((uint32_t)b), "r"(pass), "r"((uint32_t)(b >> 32)), "r"(pass) : "memory");
}
__device__ __forceinline__ bool ld_sum(const unsigned char* p, uint32_t pass, float& lo, float& hi)
{
    uint32_t x0, f0, x1, f1;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(f0), "=r"(x1), "=r"(f1) : "l"(p) : "memory");
    lo = __uint_as_float(x0); hi = __uint_as_float(x1);
    return f0 == pass && f1 == pass;
}
__device__ __forceinline__ bool ld_sum(const unsigned char* p, uint32_t pass, double& lo, double& hi)
{
    uint32_t x0, f0, x1, f1, y0, g0, y1, g1;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(f0), "=r"(x1), "=r"(f1) : "l"(p) : "memory");
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(y0), "=r"(g0), "=r"(y1), "=r"(g1) : "l"(p + 16) : "memory");
    lo = __longlong_as_double((long long)(((unsigned long long)x1 << 32) | x0));
    hi = __longlong_as_double((long long)(((unsigned long long)y1 << 32) | y0));
    return f0 == pass && f1 == pass && g0 == pass && g1 == pass;
}

template<typename REAL>
struct ResidentArgs {
    const LaneDesc* desc;          // one per bundle (used when n_classes == 0)
    const uint32_t* topo;
    const int2* lay_vn;
    const int32_t* bundle_bdd;
    REAL* cfr;
    REAL* cft;
    typename real2<REAL>::type* lohi;      // current {lo, hi} buffer, updated in place
    REAL* mmd;
    unsigned char* contrib;        // one contribution record per layer entry
    unsigned char* sums;           // one sum record per variable
    const uint32_t* var_lay_begin; // variable -> its layer entries in BDD order (layout.hpp)
    const uint32_t* var_lay;
    const int32_t* nr_bdds;        // per variable (global counts in shard mode)
    uint32_t n_vars;
    uint32_t vars_per_bundle;      // bundle g owns the variables [g * vars_per_bundle, (g + 1) * vars_per_bundle)
    uint32_t own_list_cap;         // > 0: the layer lists of the owned variables (at most this many entries) are staged in shared memory
    uint32_t pass0;                // number of the last pass before this launch (its records / sums carry it)
    uint32_t sums_published;       // the sums of pass0 are already in `sums` (published by the host from the rotating sum buffers)
    REAL* bdd_lb;
    double* lb_part;               // one partial lower bound per bundle (written by the last backward pass)
    REAL omega;
    uint32_t n_iterations;
    uint32_t init_backward;        // cost_from_terminal is stale (costs were changed): recompute it on chip first (backward_run, bdd_cuda_base.cu:670-713)
    uint32_t n_bundles, bundles_per_cta, bundles_rem;
    uint32_t warp_smem_bytes;
    uint32_t own_smem_off;         // byte offset of the owner-duty lists inside a warp's shared memory
    LaneDesc cls_first[LANE_MAX_CLASSES];
    uint32_t cls_begin[LANE_MAX_CLASSES];
    uint32_t n_classes;
    unsigned long long* trace;     // diagnostics: RES_TRACE_EVENTS clock stamps per bundle (null = off)
};

// bytes of shared memory one bundle of H hops and J rows occupies (plus one spare tile of DP values and one spare hop of
// everything else at the end, so that the one-hop-ahead operand loads of the hop loop never leave the region)
inline size_t resident_bundle_bytes(uint32_t H, uint32_t J, size_t R)
{
    return (size_t)H * (128 + 256 + 64 * R + 2 * (size_t)J * 32 * R + 64 * R) + (size_t)J * 32 * R + (128 + 256 + 128 * R + 2 * (size_t)J * 32 * R);
}

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(__cvta_generic_to_global(dst)), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}

// One bundle, all iterations of the launch.  Shared-memory layout of the bundle (bytes from `wsm`; every array is hop-major
// with 32 lanes per row, i.e. the global layout, so that it moves with plain bulk copies):
//   topo  H x 128            one-hot arc-target word per (hop, lane)
//   vn    H x 256            {byte offset of the variable's sum record, nr_bdds | VN_N_NONE | VN_N_TOP}
//   lohi  H x 64 R           {lo, hi} arc costs, updated in place
//   cfr   H x J x 32 R       cost_from_root
//   cft   (H + 1) x J x 32 R cost_from_terminal; tile H is never selected (the last hop has no arcs)
//   dl    H x 64 R           normalised {delta_lo, delta_hi} of the layer's variable for the current pass
template<typename REAL, int J>
__device__ __forceinline__ void resident_bundle(const ResidentArgs<REAL>& a, const LaneDesc d, const uint32_t g, unsigned char* wsm, uint64_t* bar_load, const int lane)
{
    using R2 = typename real2<REAL>::type;
    using XR = ExchangeRec<REAL>;
    constexpr uint32_t R = sizeof(REAL);
    constexpr uint32_t S_TOPO = 128, S_VN = 256, S_LOHI = 64 * R, S_DP = J * 32 * R, S_DL = 64 * R;
    const REAL INF = real_inf<REAL>();
    const uint32_t H = d.n_hops;
    unsigned char* const m_topo = wsm;
    unsigned char* const m_vn = m_topo + H * S_TOPO;
    unsigned char* const m_lohi = m_vn + H * S_VN;
    unsigned char* const m_cfr = m_lohi + H * S_LOHI;
    unsigned char* const m_cft = m_cfr + H * S_DP;
    unsigned char* const m_dl = m_cft + (H + 1) * S_DP;
    // owner duty (only when own_list_cap > 0): begin offsets of the owned variables' layer lists and the lists themselves, as byte
    // offsets of contribution records; placed behind the largest bundle's state
    uint32_t* const m_own_begin = reinterpret_cast<uint32_t*>(wsm + a.own_smem_off);
    uint32_t* const m_own_list = m_own_begin + ((a.vars_per_bundle + 1 + 3) & ~3u);
    const REAL omega = a.omega;

    unsigned long long* trace = a.trace ? a.trace + (size_t)g * RES_TRACE_EVENTS : nullptr;
    uint32_t trace_k = 0;
    auto stamp = [&]() { if(trace && lane == 0 && trace_k < RES_TRACE_EVENTS) trace[trace_k] = clock64(); ++trace_k; };
    stamp();   // 0: start

    // ---- load: four bulk copies (four lanes, one issue) onto the warp's mbarrier
    if(lane < 4)
    {
        const void* src = a.topo + d.topo_off; unsigned char* dst = m_topo; uint32_t bytes = H * S_TOPO;
        if(lane == 1) { src = a.lay_vn + d.lay_off; dst = m_vn; bytes = H * S_VN; }
        if(lane == 2) { src = a.lohi + d.lay_off; dst = m_lohi; bytes = H * S_LOHI; }
        if(lane == 3) { src = a.cft + d.slot_off; dst = m_cft; bytes = a.init_backward ? 0u : H * S_DP; }
        if(lane == 0) mbar_arrive_expect_tx(bar_load, H * (S_TOPO + S_VN + S_LOHI) + (a.init_backward ? 0u : H * S_DP));
        if(bytes > 0) bulk_g2s(dst, src, bytes, bar_load);
    }
    const int32_t bdd_index = a.bundle_bdd[d.bdd_base + lane];
    mbar_wait(bar_load, 0);
    stamp();   // 1: state on chip

    const uint32_t lane_topo = smem_u32(m_topo) + lane * 4u, lane_vn = smem_u32(m_vn) + lane * 8u, lane_lohi = smem_u32(m_lohi) + lane * 2u * R;
    const uint32_t lane_cfr = smem_u32(m_cfr) + lane * R, lane_cft = smem_u32(m_cft) + lane * R, lane_dl = smem_u32(m_dl) + lane * 2u * R;

    // every shared-memory access is a volatile asm with a memory clobber: the state changes between the passes of one launch, and
    // a plain asm load is a pure function of its address to the compiler (which then hoists it out of the iteration loop)
    auto lds_u32 = [](uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; };
    auto lds_i2 = [](uint32_t addr) { int2 v; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory"); return v; };
    auto lds_r = [](uint32_t addr) {
        REAL v;
        if constexpr (sizeof(REAL) == 4) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
        else asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
        return v;
    };
    auto lds_r2 = [](uint32_t addr) {
        R2 v;
        if constexpr (sizeof(REAL) == 4) asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
        else asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
        return v;
    };
    auto sts_r = [](uint32_t addr, REAL v) {
        if constexpr (sizeof(REAL) == 4) asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory");
        else asm volatile("st.shared.f64 [%0], %1;" :: "r"(addr), "d"(v) : "memory");
    };
    auto sts_r2 = [](uint32_t addr, REAL x, REAL y) {
        if constexpr (sizeof(REAL) == 4) asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(addr), "f"(x), "f"(y) : "memory");
        else asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(addr), "d"(x), "d"(y) : "memory");
    };
    auto bit = [](uint32_t t, int j, int arc, int r) -> bool { return (t & (1u << (j * 2 * J + arc * J + r))) != 0; };

    // {variable, nr_bdds} -> {byte offset of the variable's sum record, nr_bdds | VN_N_NONE | VN_N_TOP}
    for(uint32_t h = 0; h < H; ++h)
    {
        const int2 vn = lds_i2(lane_vn + h * S_VN);
        const uint32_t o = (uint32_t)max(vn.x, 0) * XR::SUM;
        const uint32_t c = vn.x >= 0 ? (uint32_t)max(vn.y, 1) : (vn.x == LAY_TOP ? VN_N_TOP : VN_N_NONE);
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(lane_vn + h * S_VN), "r"(o), "r"(c) : "memory");
    }

    if(a.init_backward)
    {   // backward_run(false): plain shortest paths to the top sink with the current arc costs
        REAL fr[J];
#pragma unroll
        for(int j = 0; j < J; ++j) fr[j] = INF;
        for(uint32_t hh = 0; hh < H; ++hh)
        {
            const uint32_t h = H - 1 - hh;
            const uint32_t t = lds_u32(lane_topo + h * S_TOPO);
            const uint32_t cnt = (uint32_t)lds_i2(lane_vn + h * S_VN).y;
            const R2 c2 = lds_r2(lane_lohi + h * S_LOHI);
            REAL val[J];
#pragma unroll
            for(int j = 0; j < J; ++j)
            {
                REAL ta = INF, tb = INF;
#pragma unroll
                for(int r = 0; r < J; ++r)
                {
                    ta = bit(t, j, 0, r) ? fr[r] : ta;
                    tb = bit(t, j, 1, r) ? fr[r] : tb;
                }
                val[j] = rmin(c2.y + tb, c2.x + ta);
                if(j == 0) val[j] = cnt == VN_N_TOP ? (REAL)0 : val[j];
            }
#pragma unroll
            for(int j = 0; j < J; ++j) { sts_r(lane_cft + h * S_DP + j * 32 * R, val[j]); fr[j] = val[j]; }
        }
    }

    struct In { uint32_t t, cnt; REAL lo, hi, d0, d1; REAL c[J]; };
    // operands of hop h: forward reads cost_from_terminal of tile h + 1, backward cost_from_root of tile h
    auto load_fwd = [&](uint32_t h) {
        In x;
        x.t = lds_u32(lane_topo + h * S_TOPO); x.cnt = 0;
        const R2 c2 = lds_r2(lane_lohi + h * S_LOHI); x.lo = c2.x; x.hi = c2.y;
        const R2 dl = lds_r2(lane_dl + h * S_DL); x.d0 = dl.x; x.d1 = dl.y;
#pragma unroll
        for(int r = 0; r < J; ++r) x.c[r] = lds_r(lane_cft + (h + 1) * S_DP + r * 32 * R);
        return x;
    };
    auto load_bwd = [&](uint32_t h) {
        In x;
        x.t = lds_u32(lane_topo + h * S_TOPO); x.cnt = lds_u32(lane_vn + h * S_VN + 4);
        const R2 c2 = lds_r2(lane_lohi + h * S_LOHI); x.lo = c2.x; x.hi = c2.y;
        const R2 dl = lds_r2(lane_dl + h * S_DL); x.d0 = dl.x; x.d1 = dl.y;
#pragma unroll
        for(int r = 0; r < J; ++r) x.c[r] = lds_r(lane_cfr + h * S_DP + r * 32 * R);
        return x;
    };

    // ---- owner duty: sums of pass `pass` for the variables this bundle owns (compute_delta :379-393 + normalize_delta :410-430,
    // in BDD order with an exact division like the single-threaded CPU solver)
    const uint32_t own_begin = min(a.n_vars, g * a.vars_per_bundle), own_end = min(a.n_vars, own_begin + a.vars_per_bundle);
    const bool own_staged = a.own_list_cap > 0;
    uint32_t dbg_reduce_retries = 0, dbg_gather_retries = 0; long long dbg_reduce_first = 0, dbg_gather_first = 0, dbg_stage = 0;
    const long long dbg_t_stage0 = trace ? clock64() : 0;
    const uint32_t list0 = own_begin < own_end ? a.var_lay_begin[own_begin] : 0u;       // first list entry of the owned range
    if(own_staged)
    {
        const uint32_t n_own = own_end - own_begin, n_list = (own_begin < own_end ? a.var_lay_begin[own_end] : 0u) - list0;
        for(uint32_t i = lane; i <= n_own; i += 32) m_own_begin[i] = a.var_lay_begin[own_begin + i] - list0;
        for(uint32_t i = lane; i < n_list; i += 32) m_own_list[i] = a.var_lay[list0 + i] * XR::CONTRIB;
        __syncwarp();
    }
    if(trace) dbg_stage = clock64() - dbg_t_stage0;
    // list entry i (relative to the owned range) -> byte offset of the contribution record
    auto own_entry = [&](uint32_t i) -> uint32_t { return own_staged ? m_own_list[i] : a.var_lay[list0 + i] * XR::CONTRIB; };
    auto reduce_owned = [&](const uint32_t pass) {
        constexpr uint32_t B = 12;
        // two variables per lane and round (A: v0 + lane, B: v0 + 32 + lane): their loads are in flight together
        for(uint32_t v0 = own_begin; v0 < own_end; v0 += 64)
        {
            uint32_t vv[2], ib[2], ie[2]; REAL lo[2], hi[2];
#pragma unroll
            for(int q = 0; q < 2; ++q)
            {
                vv[q] = v0 + q * 32 + lane;
                const bool mine = vv[q] < own_end;
                const uint32_t r = vv[q] - own_begin;
                ib[q] = !mine ? 0u : (own_staged ? m_own_begin[r] : a.var_lay_begin[vv[q]] - list0);
                ie[q] = !mine ? 0u : (own_staged ? m_own_begin[r + 1] : a.var_lay_begin[vv[q] + 1] - list0);
                lo[q] = 0; hi[q] = 0;
            }
            while(__any_sync(0xffffffffu, ib[0] < ie[0] || ib[1] < ie[1]))
            {
                uint32_t src[2][B]; REAL dv[2][B]; uint32_t want = 0, need = 0;       // bit q * B + k
#pragma unroll
                for(int q = 0; q < 2; ++q)
#pragma unroll
                    for(uint32_t k = 0; k < B; ++k)
                    {
                        const bool w = ib[q] + k < ie[q];
                        src[q][k] = w ? own_entry(ib[q] + k) : 0u;
                        dv[q][k] = 0;
                        want |= w ? (1u << (q * B + k)) : 0u;
                    }
                // first attempt: all loads in flight together (no load depends on the outcome of another)
                const long long t_first = trace ? clock64() : 0;
#pragma unroll
                for(int q = 0; q < 2; ++q)
#pragma unroll
                    for(uint32_t k = 0; k < B; ++k)
                        if(want & (1u << (q * B + k)))
                            need |= ld_contrib(a.contrib + src[q][k], pass, dv[q][k]) ? 0u : (1u << (q * B + k));
                if(trace) { dbg_reduce_first += clock64() - t_first + (need & 0); }
                for(uint32_t spins = 0; need != 0; ++spins)
                {   // records of producers that have not got there yet
                    ++dbg_reduce_retries;
                    if(spins > RES_MAX_SPINS) __trap();          // a record that never arrives (a protocol bug) must not hang the GPU
                    __nanosleep(20);
                    uint32_t still = 0;
#pragma unroll
                    for(int q = 0; q < 2; ++q)
#pragma unroll
                        for(uint32_t k = 0; k < B; ++k)
                            if(need & (1u << (q * B + k)))
                                still |= ld_contrib(a.contrib + src[q][k], pass, dv[q][k]) ? 0u : (1u << (q * B + k));
                    need = still;
                }
#pragma unroll
                for(int q = 0; q < 2; ++q)
                {
#pragma unroll
                    for(uint32_t k = 0; k < B; ++k)
                        if(ib[q] + k < ie[q]) { if(dv[q][k] > 0) hi[q] += dv[q][k]; else if(dv[q][k] < 0) lo[q] += -dv[q][k]; }      // compute_delta_atomic, :358-376
                    ib[q] = min(ib[q] + B, ie[q]);
                }
            }
#pragma unroll
            for(int q = 0; q < 2; ++q)
                if(vv[q] < own_end)
                {
                    const REAL nn = (REAL)max(a.nr_bdds[vv[q]], 1);
                    st_sum(a.sums + (size_t)vv[q] * XR::SUM, lo[q] / nn, hi[q] / nn, pass);
                }
        }
    };
    // the normalised sums of pass `pass` for the variables of this bundle's layers -> dl
    auto gather = [&](const uint32_t pass) {
        constexpr uint32_t B = 11;
        for(uint32_t h0 = 0; h0 < H; h0 += B)
        {
            uint32_t src[B]; REAL d0[B], d1[B]; uint32_t want = 0, need = 0;
#pragma unroll
            for(uint32_t k = 0; k < B; ++k)
            {
                const int2 vn = lds_i2(lane_vn + min(h0 + k, H - 1) * S_VN);
                src[k] = (uint32_t)vn.x;
                d0[k] = 0; d1[k] = 0;
                want |= (h0 + k < H && vn.y > 0) ? (1u << k) : 0u;          // VN_N_TOP is negative as int: entries without a variable read nothing
            }
            const long long t_first = trace ? clock64() : 0;
#pragma unroll
            for(uint32_t k = 0; k < B; ++k)
                if(want & (1u << k)) need |= ld_sum(a.sums + src[k], pass, d0[k], d1[k]) ? 0u : (1u << k);
            if(trace) { dbg_gather_first += clock64() - t_first + (need & 0); }
            for(uint32_t spins = 0; need != 0; ++spins)
            {
                ++dbg_gather_retries;
                if(spins > RES_MAX_SPINS) __trap();
                __nanosleep(20);
                uint32_t still = 0;
#pragma unroll
                for(uint32_t k = 0; k < B; ++k)
                    if(need & (1u << k)) still |= ld_sum(a.sums + src[k], pass, d0[k], d1[k]) ? 0u : (1u << k);
                need = still;
            }
#pragma unroll
            for(uint32_t k = 0; k < B; ++k)
                if(h0 + k < H) sts_r2(lane_dl + (h0 + k) * S_DL, d0[k], d1[k]);
        }
    };

    REAL* const g_mmd = a.mmd + d.lay_off + lane;
    unsigned char* const g_contrib = a.contrib + (size_t)(d.lay_off + lane) * XR::CONTRIB;
    uint32_t pass = a.pass0;       // number of the last completed pass
    for(uint32_t it = 0; it < a.n_iterations; ++it)
    {
        // =========================================================== forward_mm (bdd_cuda_parallel_mma.cu:207-257)
        {
            if(!(it == 0 && a.sums_published)) reduce_owned(pass);
            if(it == 0) stamp();   // 2: owned sums published
            gather(pass);
            ++pass;
            if(it == 0) stamp();   // 3: first gather done
            REAL fr[J];
#pragma unroll
            for(int j = 0; j < J; ++j) fr[j] = INF;
            if(lds_i2(lane_vn).y > 0) fr[0] = 0;          // flush_costs_from_root, bdd_cuda_base.cu:1438-1445: a lane with a BDD has a layer at hop 0
            In nx_in = load_fwd(0);
#pragma unroll 2
            for(uint32_t h = 0; h < H; ++h)
            {
                const In x = nx_in;
                nx_in = load_fwd(h + 1);                    // one hop ahead of the arithmetic (hop H reads the spare rows)
                REAL mm0 = INF, mm1 = INF;
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    REAL ta = INF, tb = INF;
#pragma unroll
                    for(int r = 0; r < J; ++r)
                    {
                        ta = bit(x.t, j, 0, r) ? x.c[r] : ta;
                        tb = bit(x.t, j, 1, r) ? x.c[r] : tb;
                    }
                    const REAL m0 = fr[j] + x.lo + ta;      // same association as bdd_cuda_parallel_mma.cu:83-84
                    const REAL m1 = fr[j] + x.hi + tb;
                    mm0 = j == 0 ? m0 : rmin(mm0, m0);
                    mm1 = j == 0 ? m1 : rmin(mm1, m1);
                }
                const REAL diff = mm_difference(omega, mm0, mm1);
                const REAL lo_n = x.lo + rmin(diff, (REAL)0) + x.d0;       // :185-193
                const REAL hi_n = x.hi + rmin(-diff, (REAL)0) + x.d1;
                REAL nx[J];
#pragma unroll
                for(int r = 0; r < J; ++r)
                {
#pragma unroll
                    for(int j = 0; j < J; ++j)
                    {
                        const REAL c0 = bit(x.t, j, 0, r) ? fr[j] + lo_n : INF;
                        const REAL c1 = bit(x.t, j, 1, r) ? fr[j] + hi_n : INF;
                        const REAL cm = rmin(c0, c1);
                        nx[r] = j == 0 ? cm : rmin(nx[r], cm);
                    }
                }
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    sts_r(lane_cfr + h * S_DP + j * 32 * R, fr[j]);
                    fr[j] = nx[j];
                }
                sts_r2(lane_lohi + h * S_LOHI, lo_n, hi_n);
                g_mmd[h * 32] = diff;
                st_contrib(g_contrib + (size_t)h * 32 * XR::CONTRIB, diff, pass);
            }
            if(it == 0) stamp();   // 4: forward hops done
        }

        // =========================================================== backward_mm (:301-346)
        {
            reduce_owned(pass);
            if(it == 0) stamp();   // 5: owned sums published
            gather(pass);
            ++pass;
            if(it == 0) stamp();   // 6: second gather done
            REAL fr[J];            // cost_from_terminal of the next hop's rows
#pragma unroll
            for(int j = 0; j < J; ++j) fr[j] = INF;
            In nx_in = load_bwd(H - 1);
#pragma unroll 2
            for(uint32_t hh = 0; hh < H; ++hh)
            {
                const uint32_t h = H - 1 - hh;
                const In x = nx_in;
                nx_in = load_bwd(h > 0 ? h - 1 : 0);
                REAL ta[J], tb[J];
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    ta[j] = INF; tb[j] = INF;
#pragma unroll
                    for(int r = 0; r < J; ++r)
                    {
                        ta[j] = bit(x.t, j, 0, r) ? fr[r] : ta[j];
                        tb[j] = bit(x.t, j, 1, r) ? fr[r] : tb[j];
                    }
                }
                REAL mm0 = INF, mm1 = INF;
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    const REAL m0 = x.c[j] + x.lo + ta[j], m1 = x.c[j] + x.hi + tb[j];
                    mm0 = j == 0 ? m0 : rmin(mm0, m0);
                    mm1 = j == 0 ? m1 : rmin(mm1, m1);
                }
                const REAL diff = mm_difference(omega, mm0, mm1);
                const REAL lo_n = x.lo + rmin(diff, (REAL)0) + x.d0;       // :280-281
                const REAL hi_n = x.hi + rmin(-diff, (REAL)0) + x.d1;
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    REAL val = rmin(hi_n + tb[j], lo_n + ta[j]);           // :286; +inf where there is no node
                    if(j == 0) val = x.cnt == VN_N_TOP ? (REAL)0 : val;   // set_special_nodes_costs, bdd_cuda_base.cu:217-227
                    sts_r(lane_cft + h * S_DP + j * 32 * R, val);
                    fr[j] = val;
                }
                sts_r2(lane_lohi + h * S_LOHI, lo_n, hi_n);
                g_mmd[h * 32] = diff;
                st_contrib(g_contrib + (size_t)h * 32 * XR::CONTRIB, diff, pass);
            }
            if(it == 0) stamp();   // 7: backward hops done
            if(it + 1 == a.n_iterations)
            {   // lower_bound, bdd_cuda_base.cu:1243-1251: sum of the roots' cost_from_terminal in double (fixed order)
                const REAL root = fr[0];
                const bool mine_valid = bdd_index >= 0;
                if(mine_valid) a.bdd_lb[bdd_index] = root;
                double v = mine_valid ? (double)root : 0.0;
#pragma unroll
                for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if(lane == 0) a.lb_part[g] = v;
            }
        }
    }

    // ---- write the state back: three bulk stores from shared memory
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes -> visible to the bulk-copy engine
    __syncwarp();
    if(lane < 3)
    {
        void* dst = a.lohi + d.lay_off; const unsigned char* src = m_lohi; uint32_t bytes = H * S_LOHI;
        if(lane == 1) { dst = a.cfr + d.slot_off; src = m_cfr; bytes = H * S_DP; }
        if(lane == 2) { dst = a.cft + d.slot_off; src = m_cft; bytes = H * S_DP; }
        bulk_s2g(dst, src, bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    stamp();   // 8: written back
    if(trace && lane == 0)
    {   // diagnostics of the exchange: retry rounds and cycles spent in the first-attempt loads (summed over the launch), list staging
        trace[9] = dbg_reduce_retries; trace[10] = dbg_gather_retries; trace[11] = (unsigned long long)dbg_reduce_first;
        trace[12] = (unsigned long long)dbg_gather_first; trace[13] = (unsigned long long)dbg_stage;
    }
}

// One CTA per SM, one warp per bundle (dealt evenly: CTA b owns bundles_per_cta (+1 if b < bundles_rem) consecutive bundles).
// Launched with cudaLaunchCooperativeKernel: all CTAs are co-resident (the warps wait for each other's records).
// MAXT = 256: up to 8 warps per CTA with up to 255 registers per thread (no spills); MAXT = 512: up to 16 warps.
template<typename REAL, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) resident_kernel(const ResidentArgs<REAL> a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars_all[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g_lo = blockIdx.x * a.bundles_per_cta + min(blockIdx.x, a.bundles_rem);
    const uint32_t g_hi = g_lo + a.bundles_per_cta + (blockIdx.x < a.bundles_rem ? 1u : 0u);
    const uint32_t g = g_lo + warp;
    if(threadIdx.x < (blockDim.x >> 5)) mbar_init(bars_all + threadIdx.x, 1);
    if(threadIdx.x == 0) mbar_fence_init();
    __syncthreads();
    if(g >= g_hi) return;
    LaneDesc d;
    if(a.n_classes == 0) d = a.desc[g];
    else
    {
        uint32_t c = 0;
#pragma unroll
        for(int k = 1; k < LANE_MAX_CLASSES; ++k) if((uint32_t)k < a.n_classes && g >= a.cls_begin[k]) c = k;
        d = a.cls_first[c];
        const uint32_t q = g - a.cls_begin[c];
        d.slot_off += q * d.n_hops * d.J * 32u; d.lay_off += q * d.n_hops * 32u; d.topo_off += q * d.n_hops * 32u; d.bdd_base += q * 32u;
    }
    unsigned char* wsm = smem_raw + (size_t)warp * a.warp_smem_bytes;
    switch(d.J)
    {
        case 1: resident_bundle<REAL, 1>(a, d, g, wsm, bars_all + warp, lane); break;
        case 2: resident_bundle<REAL, 2>(a, d, g, wsm, bars_all + warp, lane); break;
        case 3: resident_bundle<REAL, 3>(a, d, g, wsm, bars_all + warp, lane); break;
        default: resident_bundle<REAL, 4>(a, d, g, wsm, bars_all + warp, lane); break;
    }
}

// Hand-over from the rotating sum buffers of the streaming kernels: sums[v] = {delta[2v] / n, delta[2v + 1] / n} tagged with `pass`
// (normalize_delta, bdd_cuda_parallel_mma.cu:410-430; `normalize` = 0 when the buffer already holds normalised values)
template<typename REAL>
__global__ void publish_sums_kernel(const REAL* __restrict__ delta, const int32_t* __restrict__ nr_bdds, unsigned char* __restrict__ sums,
                                    uint32_t pass, int normalize, uint32_t n_vars)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if(v >= n_vars) return;
    const REAL nn = normalize ? (REAL)max(nr_bdds[v], 1) : (REAL)1;
    st_sum(sums + (size_t)v * ExchangeRec<REAL>::SUM, delta[2 * (size_t)v] / nn, delta[2 * (size_t)v + 1] / nn, pass);
}

} // namespace bddb200
