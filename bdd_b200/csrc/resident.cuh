// bdd_b200/csrc/resident.cuh -- the on-chip ("resident") form of the deferred min-marginal-averaging sweep.
//
// For collections of at most one wave of lane-class bundles (layout.hpp, CLS_LANE: one lane per BDD) the whole per-BDD state of
// a bundle -- topology words, {variable, nr_bdds}, {lo, hi} arc costs, cost_from_root and cost_from_terminal -- fits the shared
// memory of the SM whose warp owns the bundle.  ONE cooperative launch then runs any number of
//   iteration() = forward_mm, normalize_delta, backward_mm, normalize_delta      (bdd_cuda_parallel_mma.cu:142-153)
// with the state staying on chip between the passes: it is read from HBM once when the launch starts (bulk-async copies, one
// mbarrier per warp) and written back once when it ends (bulk-async stores).  Between two passes the only data that leaves the
// SM are the per-variable sums of min-marginal differences (compute_delta_atomic, :358-376: one predicated red.global.add per
// layer) and the only synchronisation is one grid-wide barrier -- the sums of pass p must be complete before pass p + 1 reads
// them.  The three sum buffers rotate as in the streaming kernels (kernels.cuh): read / accumulate / clear for the pass after next.
//
// The hop arithmetic is the streaming lane kernel's (sweep_lane_bundle) operation for operation, so both produce the same
// numbers; what differs is where operands live: every hop operand is a shared-memory word at a compile-time offset from one
// running address, the per-variable values of a pass are gathered and normalised in one batch right after the barrier
// (all loads in flight together, off the hop-to-hop dependency chain), and the hop loop itself contains no global load.
#pragma once

#include "kernels.cuh"

namespace bddb200 {

constexpr int RES_TRACE_EVENTS = 16;
constexpr uint32_t VN_N_NONE = 0u, VN_N_TOP = 0xFFFFFFFFu;      // shared-memory copy of lay_vn.y for entries without a variable
constexpr uint32_t RES_SCRATCH = 64;               // REALs behind the 2V sums of every sum buffer: targets of entries without a variable
#ifndef BDDB200_RES_RED
#define BDDB200_RES_RED 1       // 1: one unconditional reduction per layer entry (zero differences add +0), 0: skipped (branch) where the difference is 0
#endif

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
    REAL* delta[3];                // rotating per-variable sums; pass p of the launch reads delta[(cur + p) % 3]
    uint32_t cur;
    uint32_t n_delta;              // 2V
    REAL* bdd_lb;
    double* lb_sum;                // LB_SLOTS partial sums (zeroed by this kernel, filled by its last backward pass)
    REAL omega;
    uint32_t n_iterations;
    uint32_t init_backward;        // cost_from_terminal is stale (costs were changed): recompute it on chip first (backward_run, bdd_cuda_base.cu:670-713)
    uint32_t n_bundles, bundles_per_cta, bundles_rem;
    uint32_t zero_pairs_per_bundle;
    uint32_t warp_smem_bytes;
    uint32_t inv_count;
    uint32_t* barrier;             // {arrival count, generation}; both return to a consistent state after every barrier
    LaneDesc cls_first[LANE_MAX_CLASSES];
    uint32_t cls_begin[LANE_MAX_CLASSES];
    uint32_t n_classes;
    unsigned long long* trace;     // diagnostics: RES_TRACE_EVENTS clock stamps per bundle (null = off)
    uint32_t debug;                // diagnostics (BDDB200_RES_DEBUG): 1 = two barriers + fences, 2 = write back and reload the state between iterations
};

// bytes of shared memory one bundle of H hops and J rows occupies (plus one spare tile of DP values and one spare hop of
// everything else at the end, so that the one-hop-ahead operand loads of the hop loop never leave the region)
inline size_t resident_bundle_bytes(uint32_t H, uint32_t J, size_t R)
{
    return (size_t)H * (128 + 256 + 64 * R + 2 * (size_t)J * 32 * R + 64 * R) + (size_t)J * 32 * R + (128 + 256 + 128 * R + 2 * (size_t)J * 32 * R);
}

// ---- grid-wide barrier -------------------------------------------------------------------------------------------------
// All CTAs of the (cooperative, hence co-resident) launch arrive; the last one resets the count and opens the next
// generation.  Thread 0's release/acquire operations at gpu scope, bracketed by CTA barriers, order every thread's earlier
// global writes and reductions before every thread's later reads (the cooperative-groups grid.sync() pattern).
__device__ __forceinline__ void grid_barrier(uint32_t* bar, const uint32_t n_ctas)
{
    __syncthreads();
    if(threadIdx.x == 0)
    {
        uint32_t gen, old;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
        if(old == n_ctas - 1)
        {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(bar), "r"(0u) : "memory");
            asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(bar + 1), "r"(gen + 1u) : "memory");
        }
        else
        {
            uint32_t seen;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar + 1) : "memory"); } while(seen == gen);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(__cvta_generic_to_global(dst)), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}

// compute_delta_atomic, bdd_cuda_parallel_mma.cu:358-376: |diff| is added to the hi sum of the variable if diff > 0, to its lo sum if
// diff < 0.  `voff` is the byte offset of the variable's {lo, hi} pair.
template<typename REAL>
__device__ __forceinline__ void red_delta(REAL* base, uint32_t voff, REAL diff)
{
    REAL* addr = reinterpret_cast<REAL*>(reinterpret_cast<unsigned char*>(base) + (voff + (diff > 0 ? (uint32_t)sizeof(REAL) : 0u)));
#if BDDB200_RES_RED == 1
    red_add_if(true, addr, fabs(diff));
#else
    red_add_if(diff != 0, addr, fabs(diff));
#endif
}

template<typename REAL> __device__ __forceinline__ typename real2<REAL>::type ldcg2(const REAL* p);
template<> __device__ __forceinline__ float2 ldcg2<float>(const float* p) { return __ldcg(reinterpret_cast<const float2*>(p)); }
template<> __device__ __forceinline__ double2 ldcg2<double>(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

// One bundle, all iterations of the launch.  Shared-memory layout of the bundle (bytes from `wsm`; every array is hop-major
// with 32 lanes per row, i.e. the global layout, so that it moves with plain bulk copies):
//   topo  H x 128            one-hot arc-target word per (hop, lane)
//   vn    H x 256            {variable | LAY_NONE | LAY_TOP, nr_bdds(variable)}
//   lohi  H x 64 R           {lo, hi} arc costs, updated in place
//   cfr   H x J x 32 R       cost_from_root
//   cft   (H + 1) x J x 32 R cost_from_terminal; tile H is never selected (the last hop has no arcs)
//   dl    H x 64 R           normalised {delta_lo, delta_hi} of the layer's variable for the current pass
template<typename REAL, int J>
__device__ __forceinline__ void resident_bundle(const ResidentArgs<REAL>& a, const LaneDesc d, const uint32_t g, unsigned char* wsm, uint64_t* bar_load,
                                                const REAL* inv_tab, const int lane, const bool active)
{
    using R2 = typename real2<REAL>::type;
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
    const uint32_t n_ctas = gridDim.x;
    const REAL omega = a.omega;

    unsigned long long* trace = (a.trace && active) ? a.trace + (size_t)g * RES_TRACE_EVENTS : nullptr;
    uint32_t trace_k = 0;
    auto stamp = [&]() { if(trace && lane == 0 && trace_k < RES_TRACE_EVENTS) trace[trace_k] = clock64(); ++trace_k; };
    stamp();   // 0: start

    int32_t bdd_index = -1;
    if(active)
    {
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
        bdd_index = a.bundle_bdd[d.bdd_base + lane];
        mbar_wait(bar_load, 0);
    }
    stamp();   // 1: state on chip
    // {variable, nr_bdds} -> {byte offset of the variable's {lo, hi} pair in a sum buffer, nr_bdds}: the gather and the reductions
    // address the sum buffers without an index computation.  Entries without a variable point at this lane's scratch pair behind
    // the sums (they only ever add +0 there and read 0 from there) and carry VN_N_NONE / VN_N_TOP in place of the count.
    if(active)
        for(uint32_t h = 0; h < H; ++h)
        {
            const uint32_t addr = smem_u32(m_vn) + h * S_VN + lane * 8u;
            int v, n; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v), "=r"(n) : "r"(addr));
            const uint32_t o = v >= 0 ? (uint32_t)v * (2u * R) : (a.n_delta + 2u * lane) * R;
            const uint32_t c = v >= 0 ? (uint32_t)n : (v == LAY_TOP ? VN_N_TOP : VN_N_NONE);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(addr), "r"(o), "r"(c) : "memory");
        }

    const uint32_t lane_topo = smem_u32(m_topo) + lane * 4u, lane_vn = smem_u32(m_vn) + lane * 8u, lane_lohi = smem_u32(m_lohi) + lane * 2u * R;
    const uint32_t lane_cfr = smem_u32(m_cfr) + lane * R, lane_cft = smem_u32(m_cft) + lane * R, lane_dl = smem_u32(m_dl) + lane * 2u * R;
    const uint32_t inv_tab_s = smem_u32(inv_tab);

    auto lds_u32 = [](uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; };
    auto lds_i2 = [](uint32_t addr) { int2 v; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory"); return v; };
    // volatile: the DP rows change between the passes of one launch (a plain asm load is a pure function of its address to the
    // compiler, which hoists it out of the iteration loop)
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

    struct In { uint32_t t; uint32_t voff, cnt; REAL lo, hi, d0, d1; REAL c[J]; };
    // operands of hop h: forward reads cost_from_terminal of tile h + 1, backward cost_from_root of tile h
    auto load_fwd = [&](uint32_t h) {
        In x;
        x.t = lds_u32(lane_topo + h * S_TOPO);
        x.voff = lds_u32(lane_vn + h * S_VN); x.cnt = 0;
        const R2 c2 = lds_r2(lane_lohi + h * S_LOHI); x.lo = c2.x; x.hi = c2.y;
        const R2 dl = lds_r2(lane_dl + h * S_DL); x.d0 = dl.x; x.d1 = dl.y;
#pragma unroll
        for(int r = 0; r < J; ++r) x.c[r] = lds_r(lane_cft + (h + 1) * S_DP + r * 32 * R);
        return x;
    };
    auto load_bwd = [&](uint32_t h) {
        In x;
        x.t = lds_u32(lane_topo + h * S_TOPO);
        const int2 vn = lds_i2(lane_vn + h * S_VN); x.voff = (uint32_t)vn.x; x.cnt = (uint32_t)vn.y;
        const R2 c2 = lds_r2(lane_lohi + h * S_LOHI); x.lo = c2.x; x.hi = c2.y;
        const R2 dl = lds_r2(lane_dl + h * S_DL); x.d0 = dl.x; x.d1 = dl.y;
#pragma unroll
        for(int r = 0; r < J; ++r) x.c[r] = lds_r(lane_cfr + h * S_DP + r * 32 * R);
        return x;
    };

    // gather + normalise the per-variable sums of the previous pass for every hop of the bundle (normalize_delta,
    // bdd_cuda_parallel_mma.cu:410-430, folded into the read); L2 loads (.cg): the sums were accumulated by other SMs
    auto gather = [&](const REAL* delta_in) {
        constexpr uint32_t B = 8;
        for(uint32_t h0 = 0; h0 < H; h0 += B)
        {
            int2 vn[B]; R2 dl[B];
#pragma unroll
            for(uint32_t k = 0; k < B; ++k) vn[k] = lds_i2(lane_vn + min(h0 + k, H - 1) * S_VN);
#pragma unroll
            for(uint32_t k = 0; k < B; ++k)
                dl[k] = ldcg2<REAL>(reinterpret_cast<const REAL*>(reinterpret_cast<const unsigned char*>(delta_in) + (uint32_t)vn[k].x));
#pragma unroll
            for(uint32_t k = 0; k < B; ++k)
            {   // entries without a variable read the 0 of their scratch pair; the table holds 1 at 0
                const REAL r = lds_real<REAL>(inv_tab_s + (uint32_t)max(min(vn[k].y, INV_TAB - 1), 0) * R);      // VN_N_TOP is negative as int
                const REAL d0 = dl[k].x * r, d1 = dl[k].y * r;
                if(h0 + k < H) sts_r2(lane_dl + (h0 + k) * S_DL, d0, d1);
            }
        }
    };
    auto clear_share = [&](REAL* buf) {
        R2* z = reinterpret_cast<R2*>(buf);
        const uint32_t pairs = a.n_delta >> 1;
        const uint32_t z0 = min(pairs, g * a.zero_pairs_per_bundle), z1 = min(pairs, z0 + a.zero_pairs_per_bundle);
        R2 zero; zero.x = 0; zero.y = 0;
        for(uint32_t i = z0 + lane; i < z1; i += 32) z[i] = zero;
    };

    if(active && a.init_backward)
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

    REAL* const g_mmd = a.mmd + d.lay_off + lane;
    uint32_t cur = a.cur;
    for(uint32_t it = 0; it < a.n_iterations; ++it)
    {
        // =========================================================== forward_mm (bdd_cuda_parallel_mma.cu:207-257)
        if(active)
        {
            REAL* const delta_out = a.delta[(cur + 1) % 3];
            gather(a.delta[cur]);
            clear_share(a.delta[(cur + 2) % 3]);
            __syncwarp();
            if(it == 0) stamp();   // 2: first gather done
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
                // compute_delta_atomic, :358-376: |diff| goes to the hi slot if diff > 0, else to lo
                red_delta(delta_out, x.voff, diff);
            }
            if(it == 0) stamp();   // 3: forward hops done
        }
        grid_barrier(a.barrier, n_ctas);
        if(it == 0) stamp();       // 4: barrier passed
        cur = (cur + 1) % 3;

        // =========================================================== backward_mm (:301-346)
        if(active)
        {
            REAL* const delta_out = a.delta[(cur + 1) % 3];
            gather(a.delta[cur]);
            clear_share(a.delta[(cur + 2) % 3]);
            __syncwarp();
            if(it == 0) stamp();   // 5: second gather done
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
                red_delta(delta_out, x.voff, diff);
            }
            if(it == 0) stamp();   // 6: backward hops done
            if(it + 1 == a.n_iterations)
            {   // lower_bound, bdd_cuda_base.cu:1243-1251: sum of the roots' cost_from_terminal in double
                const REAL root = fr[0];
                const bool mine_valid = bdd_index >= 0;
                if(mine_valid) a.bdd_lb[bdd_index] = root;
                double v = mine_valid ? (double)root : 0.0;
#pragma unroll
                for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if(lane == 0) atomicAdd(a.lb_sum + (blockIdx.x & (LB_SLOTS - 1)), v);
            }
        }
        cur = (cur + 1) % 3;
        if(it + 1 < a.n_iterations)
        {
            if(a.debug & 1u) { __threadfence(); grid_barrier(a.barrier, n_ctas); __threadfence(); }
            if((a.debug & 2u) && active)
            {   // the state takes the round trip through global memory a relaunch would give it
                for(uint32_t h = 0; h < H; ++h)
                {
                    const R2 c2 = lds_r2(lane_lohi + h * S_LOHI);
                    a.lohi[d.lay_off + h * 32 + lane] = c2;
                    for(int j = 0; j < J; ++j) a.cft[d.slot_off + (h * J + j) * 32 + lane] = lds_r(lane_cft + h * S_DP + j * 32 * R);
                }
                __threadfence();
                for(uint32_t h = 0; h < H; ++h)
                {
                    const R2 c2 = __ldcg(&a.lohi[d.lay_off + h * 32 + lane]);
                    sts_r2(lane_lohi + h * S_LOHI, c2.x, c2.y);
                    for(int j = 0; j < J; ++j) sts_r(lane_cft + h * S_DP + j * 32 * R, __ldcg(&a.cft[d.slot_off + (h * J + j) * 32 + lane]));
                }
            }
            grid_barrier(a.barrier, n_ctas);       // the end of the launch orders the last pass
        }
    }

    // ---- write the state back: three bulk stores from shared memory
    if(active)
    {
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
    }
    stamp();   // 7: written back
}

// One CTA per SM, one warp per bundle (dealt evenly: CTA b owns bundles_per_cta (+1 if b < bundles_rem) consecutive bundles).
// Launched with cudaLaunchCooperativeKernel: all CTAs are co-resident, which the grid barrier needs.
template<typename REAL>
__global__ void __launch_bounds__(512, 1) resident_kernel(const ResidentArgs<REAL> a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ REAL inv_tab[INV_TAB];
    __shared__ uint64_t bars_all[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g_lo = blockIdx.x * a.bundles_per_cta + min(blockIdx.x, a.bundles_rem);
    const uint32_t g_hi = g_lo + a.bundles_per_cta + (blockIdx.x < a.bundles_rem ? 1u : 0u);
    const uint32_t g = g_lo + warp;
    const bool active = g < g_hi;
    LaneDesc d{};
    d.J = 1; d.n_hops = 1;
    if(active)
    {
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
    }
    if(threadIdx.x < (blockDim.x >> 5)) mbar_init(bars_all + threadIdx.x, 1);
    if(threadIdx.x == 0) mbar_fence_init();
    for(uint32_t i = threadIdx.x; i < a.inv_count; i += blockDim.x) inv_tab[i] = (REAL)1 / (REAL)(i > 0 ? i : 1);
    if(blockIdx.x == 0)
        for(uint32_t i = threadIdx.x; i < (uint32_t)LB_SLOTS; i += blockDim.x) a.lb_sum[i] = 0.0;       // filled after at least one grid barrier
    __syncthreads();
    unsigned char* wsm = smem_raw + (size_t)warp * a.warp_smem_bytes;
    switch(d.J)
    {
        case 1: resident_bundle<REAL, 1>(a, d, g, wsm, bars_all + warp, inv_tab, lane, active); break;
        case 2: resident_bundle<REAL, 2>(a, d, g, wsm, bars_all + warp, inv_tab, lane, active); break;
        case 3: resident_bundle<REAL, 3>(a, d, g, wsm, bars_all + warp, inv_tab, lane, active); break;
        default: resident_bundle<REAL, 4>(a, d, g, wsm, bars_all + warp, inv_tab, lane, active); break;
    }
}

} // namespace bddb200
