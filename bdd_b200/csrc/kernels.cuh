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
// (see layout.hpp).  No global load sits on the hop-to-hop dependency chain: the bundle's
// hops are cut into chunks, and every input of a chunk (topology, the opposite direction's
// DP values, {variable, nr_bdds}, {lo, hi} arc costs) is one contiguous range that lane 0
// brings into a shared-memory pipeline stage with cp.async.bulk (TMA 1-D bulk copy, SASS
// UBLKCP) completing on a per-warp mbarrier; the per-variable delta values are gathered into
// the same stage with cp.async (LDGSTS) one chunk ahead.  The chain-dependent frontier
// (cost_from_root going forward, cost_from_terminal going backward) lives in shared memory;
// with P == 1 a BDD never leaves its lane, so the walk needs no synchronisation at all,
// with P > 1 one __syncwarp per hop.  Per-layer min-marginals are lane-local minima followed
// by log2(P) xor-shuffles; the cross-BDD sum of min-marginal differences is a
// red.global.add per layer (or, in deterministic mode, a fixed-order segmented sum in
// delta_segsum_kernel); the division by the number of BDDs per variable is folded into the
// read of delta_in; the buffer of the pass after next is zeroed by the same launch.
// Outputs are coalesced streaming stores.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstdint>

#include "layout.hpp"

namespace bddb200 {

enum SweepMode { MODE_MMA = 0, MODE_PLAIN = 1, MODE_MM = 2 };
constexpr int LB_SLOTS = 64;

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

// ---- async-copy plumbing (PTX) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
#ifndef BDDB200_FENCE
#define BDDB200_FENCE 2
#endif
__device__ __forceinline__ void mbar_fence_init()
{
    // make the generic-proxy mbarrier initialisation visible to the async proxy (TMA) of this CTA
#if BDDB200_FENCE == 2
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#elif BDDB200_FENCE == 1
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// try_wait suspends the thread in hardware for a bounded time per call; a copy that never
// completes (a size/alignment bug) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    for(uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if(spins > (1u << 22)) __trap();
}
// 1-D bulk copy global -> shared (TMA engine), completion counted in bytes on `bar`.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template<int BYTES>
__device__ __forceinline__ void cp_async_gather(void* dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" :: "r"(smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

constexpr uint32_t INV_TAB_BYTES = 1024 * 8;   // INV_TAB REALs, sized for double
constexpr int TRACE_EVENTS = 16;

constexpr int LANE_MAX_CLASSES = 8;

template<typename REAL>
struct SweepArgs {
    const uint32_t* desc;      // DESC_WORDS words per bundle, in this pass's direction
    const ChunkRec* chunks;
    const uint32_t* topo;
    const int2* lay_vn;        // per layer entry {variable or -1, nr_bdds(variable)}
    const int32_t* bundle_bdd;
    REAL* cfr;                 // cost from root, per slot
    REAL* cft;                 // cost from terminal, per slot
    const typename real2<REAL>::type* lohi_in;   // per layer entry {lo, hi} arc cost
    typename real2<REAL>::type* lohi_out;
    REAL* mmd;                 // deferred min-marginal difference per layer entry
    const REAL* delta_in;      // 2V
    const REAL* delta_in_shared;   // sums of the variables [0, n_shared_vars) after the multi-GPU exchange (== delta_in on one GPU)
    uint32_t n_shared_vars;
    REAL* delta_out;           // 2V (zeroed before the launch or by the previous pass)
    REAL* zero_buf;            // 2V buffer to clear for the pass after next (may be null)
    REAL* mm_lo_out;           // MODE_MM
    REAL* mm_hi_out;
    REAL* bdd_lb;              // backward: cost_from_terminal of every BDD's root
    double* lb_sum;            // backward: LB_SLOTS partial sums of the roots' cost_from_terminal, slot = CTA index mod LB_SLOTS (one address
                               // would serialise tens of thousands of atomics in L2); null = off; forward kernels zero them
    REAL omega;
    uint32_t n_zero;
    uint32_t bundle_first, bundle_count;
    uint32_t tile_slots;       // capacity of one shared-memory frontier buffer, in slots
    uint32_t stage_bytes;      // capacity of one pipeline stage
    uint32_t n_stages;         // pipeline depth (>= 2)
    uint32_t chunk_hops;       // lane-class kernel: hops per pipeline stage
    const REAL* inv_tab_g;     // lane-class kernel: 1 / n for n < inv_count (global memory, copied into shared memory per CTA)
    uint32_t inv_count;
    uint32_t bundles_per_cta, bundles_rem;   // lane-class kernel: CTA b owns bundles_per_cta (+1 if b < bundles_rem) consecutive bundles
    uint32_t zero_pairs_per_bundle;          // lane-class kernel: every warp clears this many {lo, hi} pairs of zero_buf
    // lane-class kernel: runs of bundles with equal (J, n_hops) are arithmetic progressions in every array; with at most
    // LANE_MAX_CLASSES runs the descriptor of a bundle is computed from these kernel parameters instead of being loaded, so
    // that the CTA prologue contains no global load at all (one memory round trip less before the first copy is issued);
    // n_classes == 0: load LaneDesc from `desc`
    LaneDesc cls_first[LANE_MAX_CLASSES];    // descriptor of the first bundle of each run
    uint32_t cls_begin[LANE_MAX_CLASSES];    // index of that bundle
    uint32_t n_classes;
    uint32_t warp_smem_bytes;  // n_stages * stage_bytes + 2 frontier buffers + mbarriers, rounded to 128
                               // (the CTA's dynamic shared memory starts with the INV_TAB_BYTES reciprocal table)
    unsigned long long* trace; // diagnostics: TRACE_EVENTS clock64() stamps per bundle (null = off)
    int normalize_in;          // divide delta_in by nr_bdds(var) while reading
    int accumulate;            // add |mm_diff| to delta_out with atomics
    // multi-GPU push exchange (lane-class kernel, PUSH build): the |mm_diff| of the variables [0, n_push_vars) -- those that occur in
    // more than one shard -- is added to EVERY rank's sum buffer by one multimem.red through the multicast mapping of the symmetric
    // buffers (the NVSwitch replicates the reduction) or, where there is no multicast mapping, to this rank's buffer and with
    // peer-memory reductions to the buffers of the other ranks whose shards contain the variable (push_mask): each rank's own buffer
    // holds the global sums of its variables, there is no exchange kernel and no second buffer.
    // Only the bundles that contain such a variable (marked in their descriptor) take part in the flag barrier: they wait for the
    // peers' flags before they read sums, and when the last of them (and of the bundles that clear the shared prefix of the next
    // buffer) has finished, the peers are told -- all other bundles of pass p + 1 overlap the tail of pass p as on one GPU.  Flags
    // carry the barrier's ordinal mod 3 (a rank is never more than one barrier ahead of a peer), known to the host.
    REAL* delta_out_mc;            // multicast address of delta_out, or null: then the peers are addressed one by one
    REAL* const* push_peers;       // device array: entry r = rank r's block of the three sum buffers as mapped here
    size_t push_offset;            // delta_out relative to the start of the block, in REALs
    const uint16_t* push_mask;     // per shared variable: the OTHER ranks whose shards contain it
    uint32_t n_push_vars;
    uint32_t* push_counters;       // [1] slots completed, [2] error, [3] slots in use, [8 + s] counted bundles of slot s finished, [8 + PUSH_SLOTS + s] how many there are
    uint32_t* const* push_flags;   // device array: entry r = rank r's flag array as mapped here
    uint32_t* push_my_flags;       // this rank's flag array (slot r is written by rank r)
    int push_world, push_rank;
    uint32_t push_stale_phase;     // a peer whose flag still shows this value has not finished the pass before this one
    uint32_t push_send_phase;      // what this pass stores into the peers' flag arrays when its counted bundles are done
    uint32_t push_n_shared_bundles;    // bundles [0, n) contain a variable shared between shards
    uint32_t push_debug;           // diagnostics (BDDB200_PUSH_DEBUG): bit 0 skip the per-bundle fence, bit 1 skip the system fence, bit 2 no reductions to peers
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

enum NormalizeMode { NORM_NONE = 0, NORM_DIVIDE = 1, NORM_RECIPROCAL = 2 };
constexpr int INV_TAB = 1024;
constexpr uint32_t CHILD_LIMIT = 0xFFE0u;   // child slot indices are below this; CHILD_BOT and the halves of TOPO_TOP / TOPO_PAD are not

// omega * (mm_hi - mm_lo), 0 if either min-marginal is infinite (compute_mm_diff_flush_mm_lo,
// bdd_cuda_parallel_mma.cu:29-42): inf - finite, finite - inf and inf - inf all fail the test.
template<typename REAL>
__device__ __forceinline__ REAL mm_difference(REAL omega, REAL mm0, REAL mm1)
{
    const REAL d = mm1 - mm0;
    return fabs(d) < real_inf<REAL>() ? omega * d : (REAL)0;
}

// Where the pieces of one chunk live inside a pipeline stage.
template<typename REAL, int MODE, bool FORWARD>
struct ChunkGeom {
    static constexpr bool NEED_DP = FORWARD ? (MODE == MODE_MMA) : (MODE != MODE_PLAIN);
    static constexpr bool NEED_DELTA = (MODE == MODE_MMA);
    uint32_t slot_off, lay_off, n, J, Jn, ne;
    uint32_t topo_bytes, dp_bytes, o_dp, o_vn, o_lohi, o_delta;
    __device__ __forceinline__ ChunkGeom(const ChunkRec& c, uint32_t bpw)
    {
        slot_off = c.slot_off; lay_off = c.lay_off; n = c.n_hops; J = c.J; Jn = c.J_next;
        ne = (n * bpw + 1u) & ~1u;
        topo_bytes = n * J * 128u;
        const uint32_t dp_rows = FORWARD ? ((n - 1u) * J + Jn) : n * J;
        dp_bytes = NEED_DP ? dp_rows * 32u * (uint32_t)sizeof(REAL) : 0u;
        o_dp = topo_bytes;
        o_vn = o_dp + dp_bytes;
        o_lohi = o_vn + ne * 8u;
        o_delta = o_lohi + ne * 2u * (uint32_t)sizeof(REAL);
    }
};

// lane l keeps the chunk record of pipeline position (base + l)
struct ChunkLite { uint32_t slot_off, lay_off, n_hops, J, J_next; };
__device__ __forceinline__ ChunkRec to_rec(const ChunkLite& c)
{
    ChunkRec r;
    r.slot_off = c.slot_off; r.lay_off = c.lay_off; r.hop_first = 0; r.n_hops = c.n_hops; r.J = c.J; r.J_next = c.J_next;
    return r;
}
__device__ __forceinline__ ChunkLite shfl_chunk(const ChunkLite& r, int src)
{
    ChunkLite o;
    o.slot_off = __shfl_sync(0xffffffffu, r.slot_off, src);
    o.lay_off = __shfl_sync(0xffffffffu, r.lay_off, src);
    o.n_hops = __shfl_sync(0xffffffffu, r.n_hops, src);
    o.J = __shfl_sync(0xffffffffu, r.J, src);
    o.J_next = __shfl_sync(0xffffffffu, r.J_next, src);
    return o;
}

// One pass over one bundle.
//   forward : forward_mm (bdd_cuda_parallel_mma.cu:207-257) for MODE_MMA, forward_run
//             (bdd_cuda_base.cu:588-612) for MODE_PLAIN.
//   backward: backward_mm (:301-346) for MODE_MMA, backward_run(false) (bdd_cuda_base.cu:670-713)
//             for MODE_PLAIN, backward_run(true) + the per-layer min reduction of
//             min_marginals_cuda (:716-736) for MODE_MM.
// Layer update shared by both directions:
//   mm_diff = omega * (mm_hi - mm_lo), 0 if either is infinite  (bdd_cuda_parallel_mma.cu:29-42)
//   lo' = lo + min(mm_diff, 0) + delta[2v];  hi' = hi + min(-mm_diff, 0) + delta[2v+1]   (:185-193, :280-281)
//
// `w` is this lane's word of the bundle's 32-word descriptor block (layout.hpp, DescWord).
//
// JMAX > 0 selects the register variant for narrow bundles (at most JMAX rows per tile): the
// frontier entry of slot (j, lane) is register fr[j] of that lane, the hop body is branch
// free, relaxation along the arcs is a pull over the P lanes of the BDD with xor-shuffles
// (forward), children are selected from registers (backward, P == 1) or read from a
// shared-memory frontier (backward, P > 1).  JMAX == 0 is the generic variant with the
// frontier in shared memory and shared-memory atomic minima for the forward relaxation.
template<typename REAL, int LOGP, int MODE, bool FORWARD, int JMAX>
__device__ __forceinline__ void sweep_bundle(const SweepArgs<REAL>& a, const uint32_t w, unsigned char* wsm, const REAL* inv_tab, const int lane)
{
    constexpr bool RP = JMAX > 0;
    constexpr int JM = RP ? JMAX : 1;
    constexpr bool SMEM_FRONTIER = !RP || (!FORWARD && LOGP > 0);
    using R2 = typename real2<REAL>::type;
    using Geom = ChunkGeom<REAL, MODE, FORWARD>;
    constexpr int P = 1 << LOGP;
    constexpr uint32_t BPW = 32u >> LOGP;
    const int bl = lane >> LOGP;
    const int p = lane & (P - 1);
    const REAL INF = real_inf<REAL>();
    const uint32_t NS = a.n_stages;
    const uint32_t nc = __shfl_sync(0xffffffffu, w, DESC_N_CHUNKS);
    const uint32_t bdd_base = __shfl_sync(0xffffffffu, w, DESC_BDD_BASE);
    const uint32_t chunk_base = __shfl_sync(0xffffffffu, w, DESC_CHUNK_BASE);
    int32_t bdd_index = -1;
    if(!FORWARD && p == 0) bdd_index = a.bundle_bdd[bdd_base + bl];     // consumed after the last hop
    unsigned long long* trace = a.trace ? a.trace + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * TRACE_EVENTS : nullptr;
    uint32_t trace_k = 1;
    auto stamp = [&]() { if(trace && lane == 0 && trace_k < TRACE_EVENTS) trace[trace_k] = clock64(); ++trace_k; };
    stamp();   // 1: descriptor block arrived

    unsigned char* stages = wsm;
    REAL* cur = reinterpret_cast<REAL*>(wsm + (size_t)NS * a.stage_bytes);
    REAL* nxt = cur + a.tile_slots;
    uint64_t* bars = reinterpret_cast<uint64_t*>(nxt + a.tile_slots);

    if(lane == 0)
    {
        for(uint32_t s = 0; s < NS; ++s) mbar_init(bars + s, 1);
        mbar_fence_init();
    }
    REAL fr[JM];          // register frontier: cost_from_root of this hop / cost_from_terminal of the next
#pragma unroll
    for(int j = 0; j < JM; ++j) fr[j] = INF;
    if(FORWARD && SMEM_FRONTIER)
    {   // invariant: a frontier buffer is all +inf outside the tile it currently holds
        for(uint32_t i = lane; i < 2 * a.tile_slots; i += 32) cur[i] = INF;
    }
    __syncwarp();

    // ---- chunk records: the first DESC_CHUNKS pipeline positions come with the descriptor
    // block, later ones from the chunk table (off the critical path).  Lookups at loop
    // position i only ask for positions in [i + 1, i + NS]; on a miss the window restarts at
    // `low`, the lowest position that can still be asked for.
    auto chunk_of = [&](uint32_t i) { return chunk_base + (FORWARD ? i : nc - 1 - i); };
    ChunkLite mine;
    {
        const int k = DESC_FIRST_CHUNK + DESC_CHUNK_WORDS * min(lane, DESC_CHUNKS - 1);
        mine.slot_off = __shfl_sync(0xffffffffu, w, k);
        mine.lay_off = __shfl_sync(0xffffffffu, w, k + 1);
        mine.n_hops = __shfl_sync(0xffffffffu, w, k + 2);
        mine.J = __shfl_sync(0xffffffffu, w, k + 3);
        mine.J_next = __shfl_sync(0xffffffffu, w, k + 4);
    }
    uint32_t cache_base = 0, cache_count = DESC_CHUNKS;
    auto get_chunk = [&](uint32_t i, uint32_t low) -> ChunkRec {
        if(i >= cache_base + cache_count)
        {
            cache_base = low; cache_count = 32;
            const ChunkRec r = a.chunks[chunk_of(min(low + lane, nc - 1))];
            mine.slot_off = r.slot_off; mine.lay_off = r.lay_off; mine.n_hops = r.n_hops; mine.J = r.J; mine.J_next = r.J_next;
        }
        return to_rec(shfl_chunk(mine, (int)(i - cache_base)));
    };

    // lane 0: start the bulk copies of pipeline position i into stage i % NS
    auto issue = [&](uint32_t i, const ChunkRec& cr) {
        const Geom g(cr, BPW);
        if(lane == 0)
        {
            unsigned char* st = stages + (size_t)(i % NS) * a.stage_bytes;
            uint64_t* bar = bars + (i % NS);
            const uint32_t lay_bytes8 = g.ne * 8u, lohi_bytes = g.ne * 2u * (uint32_t)sizeof(REAL);
            mbar_arrive_expect_tx(bar, g.topo_bytes + g.dp_bytes + lay_bytes8 + lohi_bytes);
            bulk_g2s(st, a.topo + g.slot_off, g.topo_bytes, bar);
            if(Geom::NEED_DP && g.dp_bytes > 0)
            {
                const REAL* src = FORWARD ? a.cft + g.slot_off + 32u * g.J : a.cfr + g.slot_off;
                bulk_g2s(st + g.o_dp, src, g.dp_bytes, bar);
            }
            bulk_g2s(st + g.o_vn, a.lay_vn + g.lay_off, lay_bytes8, bar);
            bulk_g2s(st + g.o_lohi, a.lohi_in + g.lay_off, lohi_bytes, bar);
        }
    };
    // all lanes: wait for position i's bulk data, then gather its delta values
    auto land = [&](uint32_t i, const ChunkRec& cr) {
        mbar_wait(bars + (i % NS), (i / NS) & 1u);
        if(Geom::NEED_DELTA)
        {
            const Geom g(cr, BPW);
            unsigned char* st = stages + (size_t)(i % NS) * a.stage_bytes;
            const int2* s_vn = reinterpret_cast<const int2*>(st + g.o_vn);
            R2* s_delta = reinterpret_cast<R2*>(st + g.o_delta);
            for(uint32_t e = lane; e < g.n * BPW; e += 32)
            {
                const int var = s_vn[e].x;
                if(var >= 0) cp_async_gather<(int)sizeof(R2)>(s_delta + e, ((uint32_t)var < a.n_shared_vars ? a.delta_in_shared : a.delta_in) + 2 * (size_t)var);
            }
            cp_async_commit();
        }
    };

    {
        const uint32_t pre = min(NS, nc);
        for(uint32_t i = 0; i < pre; ++i) issue(i, get_chunk(i, 0));
    }
    stamp();       // 2: first bulk copies issued
    ChunkRec cr = get_chunk(0, 0);
    land(0, cr);
    stamp();       // 3: first chunk landed
    if(FORWARD)
    {   // flush_costs_from_root (bdd_cuda_base.cu:1438-1445): the root is slot `lane` of the first tile
        const bool has_bdd = p == 0 && reinterpret_cast<const uint32_t*>(stages)[lane] != TOPO_PAD;
        if(SMEM_FRONTIER) { if(has_bdd) cur[lane] = 0; }
        else if(has_bdd) fr[0] = 0;
    }

    for(uint32_t i = 0; i < nc; ++i)
    {
        ChunkRec cr_next = cr;
        if(i + 1 < nc)
        {
            cr_next = get_chunk(i + 1, i + 1);
            land(i + 1, cr_next);
            if(Geom::NEED_DELTA) cp_async_wait<1>();
        }
        else if(Geom::NEED_DELTA) cp_async_wait<0>();
        __syncwarp();
        stamp();   // 4 + 2i: chunk i ready (next chunk landed, own gathers complete)

        const Geom g(cr, BPW);
        const unsigned char* st = stages + (size_t)(i % NS) * a.stage_bytes;
        const uint32_t* s_topo = reinterpret_cast<const uint32_t*>(st);
        const REAL* s_dp = reinterpret_cast<const REAL*>(st + g.o_dp);
        const int2* s_vn = reinterpret_cast<const int2*>(st + g.o_vn);
        const R2* s_lohi = reinterpret_cast<const R2*>(st + g.o_lohi);
        const R2* s_delta = reinterpret_cast<const R2*>(st + g.o_delta);
        const uint32_t J = g.J;

        for(uint32_t hh = 0; hh < g.n; ++hh)
        {
            const uint32_t h = FORWARD ? hh : g.n - 1 - hh;
            const uint32_t e = h * BPW + bl;
            const uint32_t lay = g.lay_off + e;
            const int2 vn = s_vn[e];
            const int var = vn.x;
            const R2 c2 = s_lohi[e];                       // entries without a layer hold {0, 0}
            const REAL lo_c = c2.x, hi_c = c2.y;
            REAL d0 = 0, d1 = 0;
            if(MODE == MODE_MMA)
            {
                const R2 d = s_delta[e];                   // not gathered (garbage) where var < 0
                d0 = var >= 0 ? d.x : (REAL)0; d1 = var >= 0 ? d.y : (REAL)0;
                if(a.normalize_in == NORM_DIVIDE)
                {
                    const REAL n = (REAL)(var >= 0 ? vn.y : 1);
                    d0 /= n; d1 /= n;
                }
                else if(a.normalize_in == NORM_RECIPROCAL)
                {
                    REAL r = inv_tab[min(max(vn.y, 0), INV_TAB - 1)];
                    if(vn.y >= INV_TAB) r = (REAL)1 / (REAL)vn.y;
                    d0 *= r; d1 *= r;
                }
            }
            const uint32_t* trow = s_topo + h * J * 32u + lane;
            const uint32_t gslot = g.slot_off + h * J * 32u + lane;
            REAL lo_n = lo_c, hi_n = hi_c, diff = 0;

            if constexpr (RP)
            {
                uint32_t t[JM];
#pragma unroll
                for(int j = 0; j < JM; ++j) t[j] = (uint32_t)j < J ? trow[j * 32] : TOPO_PAD;
                // TOPO_TOP / TOPO_PAD words decode to child indices >= CHILD_LIMIT: their candidates are +inf
                if(FORWARD)
                {
                    if(MODE == MODE_MMA)
                    {
                        const REAL* child = s_dp + h * J * 32u;             // cost_from_terminal of the next hop's tile
                        REAL mm0 = INF, mm1 = INF;
#pragma unroll
                        for(int j = 0; j < JM; ++j)
                        {
                            const uint32_t lo = t[j] & 0xFFFFu, hi = t[j] >> 16;
                            const REAL ta = lo < CHILD_LIMIT ? child[lo] : INF;
                            const REAL tb = hi < CHILD_LIMIT ? child[hi] : INF;
                            const REAL m0 = fr[j] + lo_c + ta;     // same association as bdd_cuda_parallel_mma.cu:83-84
                            const REAL m1 = fr[j] + hi_c + tb;
                            mm0 = m0 < mm0 ? m0 : mm0;
                            mm1 = m1 < mm1 ? m1 : mm1;
                        }
                        mm0 = group_min<P>(mm0);
                        mm1 = group_min<P>(mm1);
                        diff = mm_difference(a.omega, mm0, mm1);
                        lo_n = lo_c + (diff < 0 ? diff : (REAL)0) + d0;
                        hi_n = hi_c + (-diff < 0 ? -diff : (REAL)0) + d1;
                    }
                    // relaxation as a pull: every slot collects the arcs of the P lanes of its BDD
                    REAL nx[JM];
#pragma unroll
                    for(int r = 0; r < JM; ++r) nx[r] = INF;
#pragma unroll
                    for(int d = 0; d < P; ++d)
                    {
#pragma unroll
                        for(int j = 0; j < JM; ++j)
                        {
                            uint32_t tp = t[j];
                            REAL x = fr[j];
                            if(d > 0)
                            {
                                tp = (uint32_t)j < J ? trow[j * 32 + ((lane ^ d) - lane)] : TOPO_PAD;
                                x = __shfl_xor_sync(0xffffffffu, fr[j], d);
                            }
                            const REAL v0 = x + lo_n, v1 = x + hi_n;
                            const uint32_t lo = tp & 0xFFFFu, hi = tp >> 16;
#pragma unroll
                            for(int r = 0; r < JM; ++r)
                            {
                                const uint32_t sr = (uint32_t)(r * 32 + lane);
                                nx[r] = (lo == sr && v0 < nx[r]) ? v0 : nx[r];
                                nx[r] = (hi == sr && v1 < nx[r]) ? v1 : nx[r];
                            }
                        }
                    }
#pragma unroll
                    for(int j = 0; j < JM; ++j)
                    {
                        if((uint32_t)j < J) a.cfr[gslot + j * 32] = fr[j];
                        fr[j] = nx[j];
                    }
                }
                else
                {
                    REAL ta[JM], tb[JM];
#pragma unroll
                    for(int j = 0; j < JM; ++j)
                    {
                        const uint32_t lo = t[j] & 0xFFFFu, hi = t[j] >> 16;
                        if(LOGP == 0)
                        {   // children live in this lane: select by row (CHILD_BOT / TOP / PAD -> no match)
                            ta[j] = INF; tb[j] = INF;
#pragma unroll
                            for(int r = 0; r < JM; ++r)
                            {
                                ta[j] = (lo >> 5) == (uint32_t)r ? fr[r] : ta[j];
                                tb[j] = (hi >> 5) == (uint32_t)r ? fr[r] : tb[j];
                            }
                        }
                        else
                        {
                            ta[j] = lo < CHILD_LIMIT ? nxt[lo] : INF;
                            tb[j] = hi < CHILD_LIMIT ? nxt[hi] : INF;
                        }
                    }
                    if(MODE != MODE_PLAIN)
                    {
                        const REAL* mine_dp = s_dp + h * J * 32u + lane;    // cost_from_root of this hop's tile
                        REAL mm0 = INF, mm1 = INF;
#pragma unroll
                        for(int j = 0; j < JM; ++j)
                        {
                            const REAL c = (uint32_t)j < J ? mine_dp[j * 32] : INF;
                            REAL m0, m1;
                            if(MODE == MODE_MMA) { m0 = c + lo_c + ta[j]; m1 = c + hi_c + tb[j]; }
                            else { m0 = c + (ta[j] + lo_c); m1 = c + (tb[j] + hi_c); }   // path costs, bdd_cuda_base.cu:636-641
                            mm0 = m0 < mm0 ? m0 : mm0;
                            mm1 = m1 < mm1 ? m1 : mm1;
                        }
                        mm0 = group_min<P>(mm0);
                        mm1 = group_min<P>(mm1);
                        if(MODE == MODE_MMA)
                        {
                            diff = mm_difference(a.omega, mm0, mm1);
                            lo_n = lo_c + (diff < 0 ? diff : (REAL)0) + d0;
                            hi_n = hi_c + (-diff < 0 ? -diff : (REAL)0) + d1;
                        }
                        else if(p == 0 && var >= 0)
                        {
                            a.mm_lo_out[lay] = mm0;
                            a.mm_hi_out[lay] = mm1;
                        }
                    }
#pragma unroll
                    for(int j = 0; j < JM; ++j)
                    {
                        const REAL vh = hi_n + tb[j], vl = lo_n + ta[j];      // bdd_cuda_parallel_mma.cu:286
                        REAL val = vh < vl ? vh : vl;                          // +inf for TOPO_PAD
                        val = t[j] == TOPO_TOP ? (REAL)0 : val;               // set_special_nodes_costs, bdd_cuda_base.cu:217-227
                        if((uint32_t)j < J)
                        {
                            a.cft[gslot + j * 32] = val;
                            if(LOGP > 0) cur[j * 32 + lane] = val;
                        }
                        if(LOGP == 0) fr[j] = (uint32_t)j < J ? val : INF;
                    }
                }
            }
            else if(FORWARD)
            {
                if(MODE == MODE_MMA)
                {
                    const REAL* child = s_dp + h * J * 32u;     // cost_from_terminal of the next hop's tile
                    REAL mm0 = INF, mm1 = INF;
                    for(uint32_t j = 0; j < J; ++j)
                    {
                        const uint32_t t = trow[j * 32];
                        if(t < TOPO_TOP)
                        {
                            const REAL c = cur[j * 32 + lane];
                            const uint32_t lo = t & 0xFFFFu, hi = t >> 16;
                            const REAL ta = lo == CHILD_BOT ? INF : child[lo];
                            const REAL tb = hi == CHILD_BOT ? INF : child[hi];
                            const REAL m0 = c + lo_c + ta;     // same association as bdd_cuda_parallel_mma.cu:83-84
                            const REAL m1 = c + hi_c + tb;
                            mm0 = m0 < mm0 ? m0 : mm0;
                            mm1 = m1 < mm1 ? m1 : mm1;
                        }
                    }
                    mm0 = group_min<P>(mm0);
                    mm1 = group_min<P>(mm1);
                    diff = mm_difference(a.omega, mm0, mm1);
                    lo_n = lo_c + (diff < 0 ? diff : (REAL)0) + d0;
                    hi_n = hi_c + (-diff < 0 ? -diff : (REAL)0) + d1;
                }
                for(uint32_t j = 0; j < J; ++j)
                {
                    const uint32_t t = trow[j * 32];
                    const REAL c = cur[j * 32 + lane];
                    cur[j * 32 + lane] = INF;                    // restore the all-inf invariant
                    a.cfr[gslot + j * 32] = c;
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
            }
            else
            {
                if(MODE != MODE_PLAIN)
                {
                    const REAL* mine_dp = s_dp + h * J * 32u + lane;   // cost_from_root of this hop's tile
                    REAL mm0 = INF, mm1 = INF;
                    for(uint32_t j = 0; j < J; ++j)
                    {
                        const uint32_t t = trow[j * 32];
                        if(t < TOPO_TOP)
                        {
                            const REAL c = mine_dp[j * 32];
                            const uint32_t lo = t & 0xFFFFu, hi = t >> 16;
                            const REAL ta = lo == CHILD_BOT ? INF : nxt[lo];
                            const REAL tb = hi == CHILD_BOT ? INF : nxt[hi];
                            REAL m0, m1;
                            if(MODE == MODE_MMA) { m0 = c + lo_c + ta; m1 = c + hi_c + tb; }
                            else { m0 = c + (ta + lo_c); m1 = c + (tb + hi_c); }   // path costs, bdd_cuda_base.cu:636-641
                            mm0 = m0 < mm0 ? m0 : mm0;
                            mm1 = m1 < mm1 ? m1 : mm1;
                        }
                    }
                    mm0 = group_min<P>(mm0);
                    mm1 = group_min<P>(mm1);
                    if(MODE == MODE_MMA)
                    {
                        diff = mm_difference(a.omega, mm0, mm1);
                        lo_n = lo_c + (diff < 0 ? diff : (REAL)0) + d0;
                        hi_n = hi_c + (-diff < 0 ? -diff : (REAL)0) + d1;
                    }
                    else if(p == 0 && var >= 0)
                    {
                        a.mm_lo_out[lay] = mm0;
                        a.mm_hi_out[lay] = mm1;
                    }
                }
                for(uint32_t j = 0; j < J; ++j)
                {
                    const uint32_t t = trow[j * 32];
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
                    a.cft[gslot + j * 32] = val;
                }
            }

            if(MODE == MODE_MMA && p == 0 && var >= 0)
            {
                R2 o; o.x = lo_n; o.y = hi_n;
                a.lohi_out[lay] = o;
                a.mmd[lay] = diff;
                // compute_delta_atomic, bdd_cuda_parallel_mma.cu:358-376: |diff| goes to the hi slot if diff > 0, else to lo
                if(a.accumulate && diff != 0) atomicAdd(a.delta_out + 2 * (size_t)var + (diff > 0 ? 1 : 0), fabs(diff));
            }
            if(SMEM_FRONTIER)
            {
                if(P > 1) __syncwarp();
                REAL* tmp = cur; cur = nxt; nxt = tmp;
            }
        }

        __syncwarp();                                   // every lane is done with stage i % NS
        stamp();   // 5 + 2i: chunk i computed
        if(i + NS < nc) issue(i + NS, get_chunk(i + NS, i + 1));
        cr = cr_next;
    }

    if(!FORWARD)
    {
        const REAL root = SMEM_FRONTIER ? nxt[lane] : fr[0];       // root = node 0 of hop 0
        const bool mine_valid = p == 0 && bdd_index >= 0;
        if(mine_valid) a.bdd_lb[bdd_index] = root;
        if(a.lb_sum != nullptr)
        {   // lower_bound, bdd_cuda_base.cu:1243-1251: sum over BDDs in double
            double v = mine_valid ? (double)root : 0.0;
#pragma unroll
            for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if(lane == 0) atomicAdd(a.lb_sum + (blockIdx.x & (LB_SLOTS - 1)), v);
        }
    }
}

template<typename REAL, int MODE, bool FORWARD>
__global__ void __launch_bounds__(512, 1) sweep_kernel(const SweepArgs<REAL> a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    REAL* inv_tab = reinterpret_cast<REAL*>(smem_raw);       // 1 / n for n < INV_TAB, shared by the CTA
    if(FORWARD && a.lb_sum != nullptr && blockIdx.x == 0)
        for(int i = threadIdx.x; i < LB_SLOTS; i += blockDim.x) a.lb_sum[i] = 0.0;
    if(MODE == MODE_MMA)
    {
        if(a.zero_buf != nullptr)
            for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_zero; i += gridDim.x * blockDim.x)
                a.zero_buf[i] = 0;
        if(a.normalize_in == NORM_RECIPROCAL)
        {
            for(int i = threadIdx.x; i < INV_TAB; i += blockDim.x) inv_tab[i] = (REAL)1 / (REAL)(i > 0 ? i : 1);
            __syncthreads();
        }
    }
    // bundles are dealt evenly over the CTAs: CTA b owns [b*n/G, (b+1)*n/G), at most blockDim.x/32 of them
    const uint32_t g_lo = (uint32_t)(((uint64_t)blockIdx.x * a.bundle_count) / gridDim.x);
    const uint32_t g_hi = (uint32_t)(((uint64_t)(blockIdx.x + 1) * a.bundle_count) / gridDim.x);
    uint32_t g = g_lo + warp;
    if(g >= g_hi) return;
    g += a.bundle_first;
    unsigned char* wsm = smem_raw + INV_TAB_BYTES + (size_t)warp * a.warp_smem_bytes;
    if(a.trace && lane == 0)
    {
        unsigned long long* tr = a.trace + (size_t)(blockIdx.x * wpc + warp) * TRACE_EVENTS;
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr[0] = clock64(); tr[TRACE_EVENTS - 1] = smid;
    }
    const uint32_t w = a.desc[(size_t)g * DESC_WORDS + lane];
    const uint32_t logP = __shfl_sync(0xffffffffu, w, DESC_LOGP);
    const uint32_t max_J = __shfl_sync(0xffffffffu, w, DESC_MAX_J);
    constexpr int M = (FORWARD && MODE == MODE_MM) ? MODE_PLAIN : MODE;
#define BDDB200_RUN(LP, JX) sweep_bundle<REAL, LP, M, FORWARD, JX>(a, w, wsm, inv_tab, lane)
    if(logP == 0)
    {
        if(max_J <= 1) BDDB200_RUN(0, 1);
        else if(max_J == 2) BDDB200_RUN(0, 2);
        else if(max_J <= 4) BDDB200_RUN(0, 4);
        else BDDB200_RUN(0, 0);
    }
    else if(logP <= 3 && max_J <= 2)
    {
        if(logP == 1) { if(max_J <= 1) BDDB200_RUN(1, 1); else BDDB200_RUN(1, 2); }
        else if(logP == 2) { if(max_J <= 1) BDDB200_RUN(2, 1); else BDDB200_RUN(2, 2); }
        else { if(max_J <= 1) BDDB200_RUN(3, 1); else BDDB200_RUN(3, 2); }
    }
    else
    {
        switch(logP)
        {
            case 1: BDDB200_RUN(1, 0); break;
            case 2: BDDB200_RUN(2, 0); break;
            case 3: BDDB200_RUN(3, 0); break;
            case 4: BDDB200_RUN(4, 0); break;
            case 5: BDDB200_RUN(5, 0); break;
            default: break;
        }
    }
#undef BDDB200_RUN
}

// ======================================================================================
// Lane-local class (layout.hpp, CLS_LANE): one lane per BDD, J <= 4 nodes per layer.
//
// A BDD never leaves its lane, so the whole walk is register arithmetic: the frontier
// (cost_from_root going forward, cost_from_terminal going backward) is J registers per lane,
// the topology of a (hop, lane) is one word of one-hot arc targets, child look-ups are
// predicated selects on its bits.  No shuffle, no shared-memory frontier, no __syncwarp inside
// a chunk, and the hop body is straight-line code (no branch): with about one warp per
// scheduler the pass time is the per-warp instruction chain, so everything that is not
// min-plus arithmetic is kept off it.
// Tiles are uniform (J rows per hop), so every per-hop array of a bundle is contiguous over any
// range of hops and the chunk size is a launch parameter (chunk_hops).
// Staging: four lanes issue the four bulk-async copies of a chunk (topology, opposite DP rows,
// {var, nr_bdds}, {lo, hi}) in one instruction onto the chunk's mbarrier; all lanes gather the
// chunk's delta values with cp.async one chunk ahead (the first chunk's variables are read
// straight from global memory so that its gathers overlap its bulk copies); the hop loop reads
// shared memory one hop ahead of the arithmetic (register double buffer).
// DET = deterministic mode: exact division by nr_bdds(var), no atomics (delta_segsum_kernel sums).
// ======================================================================================
template<typename REAL> __device__ __forceinline__ REAL rmin(REAL a, REAL b);
template<> __device__ __forceinline__ float rmin<float>(float a, float b) { return fminf(a, b); }
template<> __device__ __forceinline__ double rmin<double>(double a, double b) { return fmin(a, b); }

// shared-memory load through a 32-bit shared address (a generic pointer makes the compiler rebuild the window base per use)
template<typename REAL> __device__ __forceinline__ REAL lds_real(uint32_t addr);
template<> __device__ __forceinline__ float lds_real<float>(uint32_t addr) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
template<> __device__ __forceinline__ double lds_real<double>(uint32_t addr) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }

// predicated cp.async / red: no branch in the instruction stream
// (a scattered 8-byte cp.async.ca blocks the warp ~3x longer than two 4-byte ones or one 16-byte .cg:
// tools/microbench/issue_cost.cu)
__device__ __forceinline__ void cp_async_gather_if(bool p, float2* dst, const float* src)
{
#ifndef BDDB200_GATHER4X2
    // one 8-byte copy: the exchange of the per-variable sums is bound by the number of L2 requests, not by the issue cost
    // (profiles/r02_exchange_bound.md; two 4-byte copies block the issuing warp for less time but are two requests)
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q cp.async.ca.shared.global [%1], [%2], 8; }"
                 :: "r"((uint32_t)p), "r"(smem_u32(dst)), "l"(__cvta_generic_to_global(src)) : "memory");
    return;
#endif
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0;\n"
                 "@q cp.async.ca.shared.global [%1], [%2], 4;\n"
                 "@q cp.async.ca.shared.global [%1+4], [%2+4], 4; }"
                 :: "r"((uint32_t)p), "r"(smem_u32(dst)), "l"(__cvta_generic_to_global(src)) : "memory");
}
__device__ __forceinline__ void cp_async_gather_if(bool p, double2* dst, const double* src)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q cp.async.cg.shared.global [%1], [%2], 16; }"
                 :: "r"((uint32_t)p), "r"(smem_u32(dst)), "l"(__cvta_generic_to_global(src)) : "memory");
}
__device__ __forceinline__ void red_add_if(bool p, float* addr, float v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q red.global.add.f32 [%1], %2; }"
                 :: "r"((uint32_t)p), "l"(__cvta_generic_to_global(addr)), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_if(bool p, double* addr, double v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q red.global.add.f64 [%1], %2; }"
                 :: "r"((uint32_t)p), "l"(__cvta_generic_to_global(addr)), "d"(v) : "memory");
}

// Waiting for a peer is bounded by wall-clock time, not by a spin count: a peer that is merely late (a host stall, a debugger, an extra
// synchronisation on one rank) must not kill this rank's context.  After EXCHANGE_TIMEOUT_NS the waiter records the failure in
// counters[2] (the host turns it into BDDB200_ERR_EXCHANGE at its next synchronisation) and gives up waiting.
constexpr unsigned long long EXCHANGE_TIMEOUT_NS = 30ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void exchange_wait_flag(const uint32_t* flag, uint32_t epoch, uint32_t* counters)
{
    uint32_t seen;
    unsigned long long t0 = 0;
    for(uint32_t spins = 0;; ++spins)
    {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if((int32_t)(seen - epoch) >= 0) return;
        if((spins & 1023u) == 1023u)
        {
            const unsigned long long now = global_timer_ns();
            if(t0 == 0) t0 = now;
            else if(now - t0 > EXCHANGE_TIMEOUT_NS)
            {
                if(counters != nullptr) atomicExch(counters + 2, 1u); else __trap();
                return;
            }
        }
    }
}

constexpr int PUSH_SLOTS = 64;
// push exchange: wait while the peer's flag still shows `stale` (the phase before the one waited for)
__device__ __forceinline__ void push_wait_flag(const uint32_t* flag, uint32_t stale, uint32_t* counters)
{
    uint32_t seen;
    unsigned long long t0 = 0;
    for(uint32_t spins = 0;; ++spins)
    {   // relaxed polls, one acquire fence when the flag has moved (ld.acquire.sys invalidates L1 on every poll)
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if(seen != stale) { asm volatile("fence.acq_rel.sys;" ::: "memory"); return; }
        if((spins & 1023u) == 1023u)
        {
            const unsigned long long now = global_timer_ns();
            if(t0 == 0) t0 = now;
            else if(now - t0 > EXCHANGE_TIMEOUT_NS) { atomicExch(counters + 2, 1u); return; }
        }
    }
}

// the same reduction performed on every rank's copy of a symmetric buffer (addr = multicast address), in the switch.
// Never predicate it: ptxas 12.9 drops the guard of a predicated multimem.red (the SASS REDG is unconditional); branch around it.
__device__ __forceinline__ void mc_red_add(float* addr, float v)
{
    asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" :: "l"(__cvta_generic_to_global(addr)), "f"(v) : "memory");
}
__device__ __forceinline__ void mc_red_add(double* addr, double v)
{
    asm volatile("multimem.red.relaxed.sys.global.add.f64 [%0], %1;" :: "l"(__cvta_generic_to_global(addr)), "d"(v) : "memory");
}
// reduction on a peer GPU's memory (its mapping here), over NVLink
__device__ __forceinline__ void peer_red_add(float* addr, float v)
{
    asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" :: "l"(__cvta_generic_to_global(addr)), "f"(v) : "memory");
}
__device__ __forceinline__ void peer_red_add(double* addr, double v)
{
    asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" :: "l"(__cvta_generic_to_global(addr)), "d"(v) : "memory");
}

// programmatic dependent launch: everything before pdl_wait() may overlap the tail of the previous kernel in
// the stream and must not touch anything that kernel reads or writes
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template<typename REAL, int J>
struct LaneHopIn {
    uint32_t t;        // one-hot arc targets
    int var;           // variable, LAY_NONE or LAY_TOP
    REAL lo, hi;       // arc costs
    REAL d0, d1;       // normalised delta_in of the layer's variable
    REAL c[J];         // opposite direction's DP values: forward = cost_from_terminal of the NEXT hop's rows,
                       // backward = cost_from_root of THIS hop's rows
};

constexpr int LANE_MAX_STAGES = 4;
#ifndef BDDB200_LANE_UNROLL
#define BDDB200_LANE_UNROLL 2
#endif

template<typename REAL, int J, int MODE, bool FORWARD, bool DET, bool PUSH = false>
__device__ __forceinline__ void sweep_lane_bundle(const SweepArgs<REAL>& a, const LaneDesc d, const uint32_t bundle_in_launch, unsigned char* wsm, uint64_t* bars, const REAL* inv_tab, const int lane,
                                                  const bool waits_for_peers = false)
{
    using R2 = typename real2<REAL>::type;
    using In = LaneHopIn<REAL, J>;
    constexpr bool NEED_DP = FORWARD ? (MODE == MODE_MMA) : (MODE != MODE_PLAIN);
    constexpr bool NEED_DELTA = (MODE == MODE_MMA);
    constexpr uint32_t R = sizeof(REAL);
    constexpr uint32_t HOP_TOPO = 128, HOP_DP = J * 32 * R, HOP_VN = 256, HOP_LOHI = 64 * R;
    const REAL INF = real_inf<REAL>();
    const uint32_t NS = a.n_stages, n = a.chunk_hops, H = d.n_hops;
    const uint32_t nc = (H + n - 1) / n;
    const uint32_t o_dp = n * HOP_TOPO, o_vn = o_dp + n * HOP_DP, o_lohi = o_vn + n * HOP_VN, o_delta = o_lohi + n * HOP_LOHI;
    unsigned long long* trace = a.trace ? a.trace + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * TRACE_EVENTS : nullptr;
    uint32_t trace_k = 1;
    auto stamp = [&]() { if(trace && lane == 0 && trace_k < TRACE_EVENTS) trace[trace_k] = clock64(); ++trace_k; };
    stamp();   // 1: descriptor arrived

    // pipeline position i <-> chunk ci(i) = hops [ci*n, min(H, ci*n + n))
    auto chunk_first = [&](uint32_t i) { return (FORWARD ? i : nc - 1 - i) * n; };
    // positions [i0, i0 + count): lane l starts copy (l & 3) of position i0 + (l >> 2); all in one UBLKCP issue
    // copies 0 (topology) and 2 ({var, nr_bdds}) are static data, 1 (DP rows) and 3 ({lo, hi}) are written by the previous pass
    constexpr uint32_t COPY_ALL = 0xFu;
    auto issue = [&](uint32_t i0, uint32_t count, uint32_t mask) {
        const uint32_t k = lane >> 2, which = lane & 3;
        if(k < count && ((mask >> which) & 1u))
        {
            const uint32_t i = i0 + k;
            const uint32_t h0 = chunk_first(i), cnt = min(n, H - h0);
            unsigned char* st = wsm + (size_t)(i % NS) * a.stage_bytes;
            uint64_t* bar = bars + (i % NS);
            // forward reads cost_from_terminal of the tiles h0+1 .. h0+cnt (the tile after the last hop does not exist)
            const uint32_t dp_tiles = !NEED_DP ? 0u : (FORWARD ? min(cnt, H - 1 - h0) : cnt);
            const REAL* dp_src = FORWARD ? a.cft + d.slot_off + (size_t)(h0 + 1) * (J * 32) : a.cfr + d.slot_off + (size_t)h0 * (J * 32);
            const void* src = a.topo + d.topo_off + h0 * 32u;
            uint32_t off = 0, bytes = cnt * HOP_TOPO;
            if(which == 1) { src = dp_src; off = o_dp; bytes = dp_tiles * HOP_DP; }
            if(which == 2) { src = a.lay_vn + d.lay_off + h0 * 32u; off = o_vn; bytes = cnt * HOP_VN; }
            if(which == 3) { src = a.lohi_in + d.lay_off + h0 * 32u; off = o_lohi; bytes = cnt * HOP_LOHI; }
            // the transaction count may run negative until the expect_tx arrives: the phase cannot complete before the arrival
            if(which == 0) mbar_arrive_expect_tx(bar, cnt * (HOP_TOPO + HOP_VN + HOP_LOHI) + dp_tiles * HOP_DP);
            if(bytes > 0) bulk_g2s(st + off, src, bytes, bar);
        }
    };
    auto delta_src = [&](int var) -> const REAL* {
        const uint32_t v = (uint32_t)max(var, 0);
        return (v < a.n_shared_vars ? a.delta_in_shared : a.delta_in) + 2 * (size_t)v;
    };
    // gather the delta values of position i; from_global reads the variables from global memory
    // (does not need the chunk to have landed)
    auto gather = [&](uint32_t i, bool from_global, uint32_t h_begin) {
        const uint32_t h0 = chunk_first(i), cnt = min(n, H - h0);
        unsigned char* st = wsm + (size_t)(i % NS) * a.stage_bytes;
        const int2* s_vn = reinterpret_cast<const int2*>(st + o_vn) + lane;
        const int2* g_vn = a.lay_vn + d.lay_off + h0 * 32u + lane;
        R2* s_delta = reinterpret_cast<R2*>(st + o_delta) + lane;
        for(uint32_t h = h_begin; h < cnt; h += 4)
        {
            int var[4];
#pragma unroll
            for(int k = 0; k < 4; ++k)
            {
                const uint32_t hk = min(h + k, cnt - 1);
                var[k] = from_global ? __ldg(&g_vn[hk * 32].x) : s_vn[hk * 32].x;
            }
#pragma unroll
            for(int k = 0; k < 4; ++k)
                cp_async_gather_if(h + k < cnt && var[k] >= 0, s_delta + (h + k) * 32, delta_src(var[k]));
        }
        cp_async_commit();
    };

    // ---- start-up: static data first (may overlap the previous kernel's tail), then everything the previous pass wrote
    stamp();       // 2: (unused)
    constexpr int G0 = 16;                 // variables of the first chunk straight from global memory, in one batch of loads
    int var0[G0];
    const uint32_t cnt0 = min(n, H - chunk_first(0));
    if(NEED_DELTA)
    {
        const int2* g_vn = a.lay_vn + d.lay_off + chunk_first(0) * 32u + lane;
#pragma unroll
        for(int k = 0; k < G0; ++k) var0[k] = (uint32_t)k < cnt0 ? __ldg(&g_vn[k * 32].x) : -1;
    }
    int32_t bdd_index = -1;
    if(!FORWARD) bdd_index = a.bundle_bdd[d.bdd_base + lane];          // consumed after the last hop
#ifndef BDDB200_NO_PREFETCH_SUMS
    if(MODE == MODE_MMA && a.n_zero * (uint32_t)sizeof(REAL) <= (4u << 20))
    {   // the per-variable sum buffers are small and accessed at random (8-byte gathers, 4-byte reductions): when they are not in
        // L2 (first pass after other work evicted them) every such access is a DRAM sector read.  Every warp asks L2 for its
        // share of both buffers with coalesced prefetches before anything else (a hint: needs no ordering with the previous pass).
        // +4 % on the flushed 1 M-node pass; buffers beyond 4 MB are left alone (no gain measured on the 5 M / 20 M-node instances).
        const uint32_t lines = (a.n_zero * (uint32_t)sizeof(REAL) + 127u) / 128u;
        const uint32_t per = (lines + a.bundle_count - 1) / a.bundle_count;
        const uint32_t l0 = min(lines, bundle_in_launch * per), l1 = min(lines, l0 + per);
        for(uint32_t l = l0 + lane; l < l1; l += 32)
        {
            asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(a.delta_in) + (size_t)l * 128));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(a.delta_out) + (size_t)l * 128));
        }
    }
#endif
    stamp();       // 3: variable loads issued
    pdl_wait();
    pdl_launch_dependents();
    stamp();       // 4: previous kernel complete
    if(PUSH && waits_for_peers)
    {   // this bundle reads sums the peers add to: complete once every peer's flag has left the phase before the previous pass's
        if(lane < a.push_world && lane != a.push_rank) push_wait_flag(a.push_my_flags + lane, a.push_stale_phase, a.push_counters);
        __syncwarp();
    }
    issue(0, 1, COPY_ALL);
    stamp();       // 5: bulk copies of the first chunk issued
    R2 dl0[G0];
    if(NEED_DELTA)
    {   // scattered LDG is ~4x cheaper to issue than scattered LDGSTS (tools/microbench/issue_cost.cu)
#pragma unroll
        for(int k = 0; k < G0; ++k)
            dl0[k] = *reinterpret_cast<const R2*>(delta_src(var0[k]));
        if(cnt0 > (uint32_t)G0) gather(0, true, G0); else cp_async_commit();
    }
    if(nc > 1) issue(1, min(NS, nc) - 1, COPY_ALL);
    stamp();       // 6: first gathers (need the variable loads) and the bulk copies of the next chunks issued
    if(MODE == MODE_MMA && a.zero_buf != nullptr)
    {   // this warp's share of the delta buffer of the pass after next, cleared while the first chunk is in flight
        R2* z = reinterpret_cast<R2*>(a.zero_buf);
        const uint32_t pairs = a.n_zero >> 1;
        const uint32_t z0 = min(pairs, bundle_in_launch * a.zero_pairs_per_bundle), z1 = min(pairs, z0 + a.zero_pairs_per_bundle);
        R2 zero; zero.x = 0; zero.y = 0;
        for(uint32_t i = z0 + lane; i < z1; i += 32) z[i] = zero;
    }
    if(NEED_DELTA)
    {
        R2* s_delta = reinterpret_cast<R2*>(wsm + o_delta) + lane;
#pragma unroll
        for(int k = 0; k < G0; ++k)
            if((uint32_t)k < cnt0) s_delta[k * 32] = dl0[k];
    }
    stamp();       // 7: zeroing issued, first gathers stored
    mbar_wait(bars + 0, 0);
    stamp();       // 8: first chunk landed

    REAL fr[J];    // frontier: forward = cost_from_root of this hop's rows, backward = cost_from_terminal of the next hop's rows
#pragma unroll
    for(int j = 0; j < J; ++j) fr[j] = INF;
    if(FORWARD)
    {   // flush_costs_from_root (bdd_cuda_base.cu:1438-1445): a lane with a BDD has a layer at hop 0
        if(reinterpret_cast<const int2*>(wsm + o_vn)[lane].x >= 0) fr[0] = 0;
    }
    const REAL omega = a.omega;
    const bool normalize = a.normalize_in != NORM_NONE;
    const uint32_t inv_tab_s = smem_u32(inv_tab);

    for(uint32_t i = 0; i < nc; ++i)
    {
        if(i + 1 < nc)
        {
            mbar_wait(bars + ((i + 1) % NS), ((i + 1) / NS) & 1u);
            if(NEED_DELTA) { gather(i + 1, false, 0); cp_async_wait<1>(); }
        }
        else if(NEED_DELTA) cp_async_wait<0>();
        stamp();   // 9 + 2i: chunk i ready (next chunk landed, own gathers complete; every lane reads only what it gathered itself)

        const uint32_t h0 = chunk_first(i), cnt = min(n, H - h0);
        const unsigned char* st = wsm + (size_t)(i % NS) * a.stage_bytes;
        const uint32_t* s_topo = reinterpret_cast<const uint32_t*>(st) + lane;
        const REAL* s_dp = reinterpret_cast<const REAL*>(st + o_dp) + lane;
        const int2* s_vn = reinterpret_cast<const int2*>(st + o_vn) + lane;
        const R2* s_lohi = reinterpret_cast<const R2*>(st + o_lohi) + lane;
        const R2* s_delta = reinterpret_cast<const R2*>(st + o_delta) + lane;
        REAL* g_dp = (FORWARD ? a.cfr : a.cft) + d.slot_off + (size_t)h0 * (J * 32) + lane;
        R2* g_lohi = a.lohi_out + d.lay_off + h0 * 32u + lane;
        REAL* g_mmd = a.mmd + d.lay_off + h0 * 32u + lane;
        REAL* g_mm_lo = a.mm_lo_out + d.lay_off + h0 * 32u + lane;
        REAL* g_mm_hi = a.mm_hi_out + d.lay_off + h0 * 32u + lane;

        auto load = [&](uint32_t h) -> In {
            In x;
            x.t = s_topo[h * 32];
            const int2 vn = s_vn[h * 32];
            x.var = vn.x;
            const R2 c2 = s_lohi[h * 32];                      // entries without a layer hold {0, 0}
            x.lo = c2.x; x.hi = c2.y;
            x.d0 = 0; x.d1 = 0;
            if(NEED_DELTA)
            {
                const R2 dl = s_delta[h * 32];                 // not gathered (garbage) where var < 0
                REAL d0, d1;
                if(DET)
                {
                    const REAL nn = normalize ? (REAL)max(vn.y, 1) : (REAL)1;
                    d0 = dl.x / nn; d1 = dl.y / nn;
                }
                else
                {   // the table holds 1/n (the host selects the DET kernels when some n >= INV_TAB), or ones
                    const REAL r = normalize ? lds_real<REAL>(inv_tab_s + min((uint32_t)vn.y, (uint32_t)(INV_TAB - 1)) * R) : (REAL)1;
                    d0 = dl.x * r; d1 = dl.y * r;
                }
                x.d0 = vn.x >= 0 ? d0 : (REAL)0; x.d1 = vn.x >= 0 ? d1 : (REAL)0;
            }
#pragma unroll
            for(int r = 0; r < J; ++r) x.c[r] = NEED_DP ? s_dp[(h * J + r) * 32] : INF;
            return x;
        };
        auto bit = [](uint32_t t, int j, int arc, int r) -> bool { return (t & (1u << (j * 2 * J + arc * J + r))) != 0; };

        In cur = load(FORWARD ? 0u : cnt - 1u);
        constexpr int HOP_UNROLL = BDDB200_LANE_UNROLL;
#pragma unroll HOP_UNROLL
        for(uint32_t hh = 0; hh < cnt; ++hh)
        {
            const uint32_t h = FORWARD ? hh : cnt - 1 - hh;
            const In x = cur;
            cur = load(FORWARD ? min(h + 1, cnt - 1) : (uint32_t)max((int)h - 1, 0));     // one hop ahead of the arithmetic
            REAL lo_n = x.lo, hi_n = x.hi, diff = 0;

            if(FORWARD)
            {
                if(MODE == MODE_MMA)
                {
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
                        const REAL m0 = fr[j] + x.lo + ta;     // same association as bdd_cuda_parallel_mma.cu:83-84
                        const REAL m1 = fr[j] + x.hi + tb;
                        mm0 = j == 0 ? m0 : rmin(mm0, m0);
                        mm1 = j == 0 ? m1 : rmin(mm1, m1);
                    }
                    diff = mm_difference(omega, mm0, mm1);
                    lo_n = x.lo + rmin(diff, (REAL)0) + x.d0;
                    hi_n = x.hi + rmin(-diff, (REAL)0) + x.d1;
                }
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
                    g_dp[(size_t)h * (J * 32) + j * 32] = fr[j];
                    fr[j] = nx[j];
                }
            }
            else
            {
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
                if(MODE != MODE_PLAIN)
                {
                    REAL mm0 = INF, mm1 = INF;
#pragma unroll
                    for(int j = 0; j < J; ++j)
                    {
                        REAL m0, m1;
                        if(MODE == MODE_MMA) { m0 = x.c[j] + x.lo + ta[j]; m1 = x.c[j] + x.hi + tb[j]; }
                        else { m0 = x.c[j] + (ta[j] + x.lo); m1 = x.c[j] + (tb[j] + x.hi); }   // path costs, bdd_cuda_base.cu:636-641
                        mm0 = j == 0 ? m0 : rmin(mm0, m0);
                        mm1 = j == 0 ? m1 : rmin(mm1, m1);
                    }
                    if(MODE == MODE_MMA)
                    {
                        diff = mm_difference(omega, mm0, mm1);
                        lo_n = x.lo + rmin(diff, (REAL)0) + x.d0;
                        hi_n = x.hi + rmin(-diff, (REAL)0) + x.d1;
                    }
                    else if(x.var >= 0)
                    {
                        g_mm_lo[h * 32] = mm0;
                        g_mm_hi[h * 32] = mm1;
                    }
                }
#pragma unroll
                for(int j = 0; j < J; ++j)
                {
                    REAL val = rmin(hi_n + tb[j], lo_n + ta[j]);           // bdd_cuda_parallel_mma.cu:286; +inf where there is no node
                    if(j == 0) val = x.var == LAY_TOP ? (REAL)0 : val;    // set_special_nodes_costs, bdd_cuda_base.cu:217-227
                    g_dp[(size_t)h * (J * 32) + j * 32] = val;
                    fr[j] = val;
                }
            }

            if(MODE == MODE_MMA)
            {   // entries without a layer hold lo = hi = 0, collect diff = 0 and read delta 0: writing them back is harmless
                R2 o; o.x = lo_n; o.y = hi_n;
                g_lohi[h * 32] = o;
                g_mmd[h * 32] = diff;
                // compute_delta_atomic, bdd_cuda_parallel_mma.cu:358-376: |diff| goes to the hi slot if diff > 0, else to lo
                if(!DET)
                {
                    const size_t slot = 2 * (size_t)max(x.var, 0) + (diff > 0 ? 1 : 0);
                    if(PUSH)
                    {   // shared between shards: one reduction on every rank's buffer; everything else stays on this GPU
                        const bool shared = (uint32_t)x.var < a.n_push_vars, mc = a.delta_out_mc != nullptr;
                        red_add_if(diff != 0 && !(shared && mc), a.delta_out + slot, fabs(diff));
                        if(diff != 0 && shared && !(a.push_debug & 4u))
                        {   // shared between shards (rare): the same reduction on every rank's buffer, in the switch ...
                            if(mc) mc_red_add(a.delta_out_mc + slot, fabs(diff));
                            else   // ... or on the buffers of the other ranks that hold the variable, one by one
                                for(uint32_t m = a.push_mask[x.var]; m != 0; m &= m - 1)
                                    peer_red_add(a.push_peers[__ffs(m) - 1] + a.push_offset + slot, fabs(diff));
                        }
                    }
                    else red_add_if(diff != 0, a.delta_out + slot, fabs(diff));
                }
            }
        }

        __syncwarp();                                   // every lane is done with stage i % NS
        stamp();   // 10 + 2i: chunk i computed
        if(i + NS < nc) issue(i + NS, 1, COPY_ALL);
    }

    if(!FORWARD)
    {
        const REAL root = fr[0];                                    // root = row 0 of hop 0
        const bool mine_valid = bdd_index >= 0;
        if(mine_valid) a.bdd_lb[bdd_index] = root;
        if(a.lb_sum != nullptr)
        {   // lower_bound, bdd_cuda_base.cu:1243-1251: sum over BDDs in double
            double v = mine_valid ? (double)root : 0.0;
#pragma unroll
            for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if(lane == 0) atomicAdd(a.lb_sum + (blockIdx.x & (LB_SLOTS - 1)), v);
        }
    }
}

#ifndef BDDB200_LANE_MAX_THREADS
#define BDDB200_LANE_MAX_THREADS 512
#endif
// MAXT = BDDB200_LANE_MAX_THREADS: up to 16 warps per SM at up to 128 registers per thread.
// MAXT = 768 ("dense", MMA passes in float only -- the double kernels would spill): the same pass compiled for 24 resident warps per SM
// (<= 80 registers per thread); HBM-bound instances of many waves gain ~6 % from the extra warps (profiles/r01_v4_ncu_sweep_summary.md).
template<typename REAL, int MODE, bool FORWARD, bool DET, int MAXT = BDDB200_LANE_MAX_THREADS, bool PUSH = false>
__global__ void __launch_bounds__(MAXT, 1) sweep_lane_kernel(const SweepArgs<REAL> a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ REAL inv_tab[INV_TAB];                        // 1 / n for n < INV_TAB (ones when delta_in is already normalised)
    __shared__ uint64_t bars_all[16 * LANE_MAX_STAGES];      // one mbarrier per warp and pipeline stage
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g_lo = blockIdx.x * a.bundles_per_cta + min(blockIdx.x, a.bundles_rem);
    const uint32_t g_hi = g_lo + a.bundles_per_cta + (blockIdx.x < a.bundles_rem ? 1u : 0u);
    const uint32_t g = g_lo + warp;
    const bool active = g < g_hi;
    if(a.trace && lane == 0 && active)
    {
        unsigned long long* tr = a.trace + (size_t)(blockIdx.x * (blockDim.x >> 5) + warp) * TRACE_EVENTS;
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr[0] = clock64(); tr[TRACE_EVENTS - 1] = smid;
    }
    LaneDesc d{};
    if(active)
    {
        if(a.n_classes == 0) d = reinterpret_cast<const LaneDesc*>(a.desc)[a.bundle_first + g];     // in flight during the CTA prologue
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
    if(threadIdx.x < (blockDim.x >> 5) * a.n_stages) mbar_init(bars_all + threadIdx.x, 1);
    if(threadIdx.x == 0) mbar_fence_init();
    if(MODE == MODE_MMA && !DET)       // 1 / n for the counts that occur (computed, not loaded: no global load in the prologue)
        for(uint32_t i = threadIdx.x; i < a.inv_count; i += blockDim.x) inv_tab[i] = (REAL)1 / (REAL)(i > 0 ? i : 1);
    __syncthreads();
    if(FORWARD && a.lb_sum != nullptr && blockIdx.x == 0 && threadIdx.x < 32)
    {   // the previous backward pass accumulated into them
        pdl_wait();
        for(int i = threadIdx.x; i < LB_SLOTS; i += 32) a.lb_sum[i] = 0.0;
    }
    if(!active) return;
    unsigned char* wsm = smem_raw + (size_t)warp * a.warp_smem_bytes;
    uint64_t* bars = bars_all + warp * a.n_stages;
    constexpr int M = (FORWARD && MODE == MODE_MM) ? MODE_PLAIN : MODE;
    // push exchange: does this bundle take part in the flag barrier?  In shard mode the layout puts the bundles that contain a shared
    // variable first (their flag goes out early in the pass and has long arrived when the peers' next pass starts)
    const bool has_shared = PUSH && g < a.push_n_shared_bundles;
    const bool counted = PUSH && (has_shared || g == 0 || g * a.zero_pairs_per_bundle < a.n_push_vars);
    switch(d.J)
    {
        case 1: sweep_lane_bundle<REAL, 1, M, FORWARD, DET, PUSH>(a, d, g, wsm, bars, inv_tab, lane, has_shared); break;
        case 2: sweep_lane_bundle<REAL, 2, M, FORWARD, DET, PUSH>(a, d, g, wsm, bars, inv_tab, lane, has_shared); break;
        case 3: sweep_lane_bundle<REAL, 3, M, FORWARD, DET, PUSH>(a, d, g, wsm, bars, inv_tab, lane, has_shared); break;
        case 4: sweep_lane_bundle<REAL, 4, M, FORWARD, DET, PUSH>(a, d, g, wsm, bars, inv_tab, lane, has_shared); break;
        default: break;
    }
    if(PUSH && counted)
    {   // pass-end barrier, sending half.  Every counted bundle fences at GPU scope (its multimem reductions and its zeros in the shared
        // prefix precede its count); the last one, having seen every count, fences at system scope and tells every peer "everything I
        // push into your buffers in this pass has arrived, and the shared prefix of the buffer you push into next is clear".  One
        // system-scope fence per pass: one per warp serialises chip-wide (~90 ns each).  Counts are spread over PUSH_SLOTS addresses.
        __syncwarp();
        if(lane == 0)
        {
            if(!(a.push_debug & 1u)) asm volatile("fence.acq_rel.gpu;" ::: "memory");      // (not __threadfence(): that is the sequentially consistent MEMBAR.SC)
            const uint32_t slot = g % PUSH_SLOTS;
            if(atomicAdd(a.push_counters + 8 + slot, 1u) == a.push_counters[8 + PUSH_SLOTS + slot] - 1)
            {
                a.push_counters[8 + slot] = 0;
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                if(atomicAdd(a.push_counters + 1, 1u) == a.push_counters[3] - 1)
                {
                    a.push_counters[1] = 0;
                    // ONE system-scope fence, then relaxed flag stores: a st.release.sys per peer is a system-scope MEMBAR per peer
                    // (seven in a row on eight GPUs, on the critical path of every rank's next pass)
                    if(!(a.push_debug & 2u)) asm volatile("fence.acq_rel.sys;" ::: "memory");
                    for(int r = 0; r < a.push_world; ++r)
                        if(r != a.push_rank) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(a.push_flags[r] + a.push_rank), "r"(a.push_send_phase) : "memory");
                }
            }
        }
    }
}

// Push-exchange barrier outside a pass (before the host reads or clears the sum buffers): what & 1 waits until no peer's flag shows
// `stale_phase` any more (every peer has completed the last pass); what & 2 announces `send_phase` (this rank's buffers are ready to be
// pushed into again).  One warp.
__global__ void push_barrier_kernel(uint32_t* counters, uint32_t* const* flags, uint32_t* my_flags, int world, int rank, int what, uint32_t stale_phase, uint32_t send_phase)
{
    const int lane = threadIdx.x;
    if((what & 1) && lane < world && lane != rank) push_wait_flag(my_flags + lane, stale_phase, counters);
    __syncwarp();
    if(what & 2)
    {
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        if(lane < world && lane != rank)
            asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(flags[lane] + rank), "r"(send_phase) : "memory");
    }
}

// ------------------------------------------------------------------ small kernels ------

// construction: per layer entry {variable or marker, nr_bdds(variable)}
__global__ void pair_lay_vn_kernel(const int32_t* __restrict__ lay_var, const int32_t* __restrict__ nr_bdds, int2* __restrict__ out, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const int32_t v = lay_var[i];
    out[i] = make_int2(v, v >= 0 ? nr_bdds[v] : 0);
}

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
__global__ void update_costs_kernel(const int2* __restrict__ lay_vn, REAL* __restrict__ cost /* lohi + component */, const REAL* __restrict__ c, uint32_t n_c, uint32_t n_lay)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_lay) return;
    const int2 vn = lay_vn[i];
    if(vn.x < 0) return;
    if((uint32_t)vn.x >= n_c) { cost[2 * (size_t)i] = 0; return; }
    cost[2 * (size_t)i] += c[vn.x] / (REAL)vn.y;
}

// both cost vectors in one launch (host path): c = [lo costs (n_lo) | hi costs (n_hi)]
template<typename REAL>
__global__ void update_costs_lohi_kernel(const int2* __restrict__ lay_vn, REAL* __restrict__ lohi, const REAL* __restrict__ c, uint32_t n_lo, uint32_t n_hi, uint32_t n_lay)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_lay) return;
    const int2 vn = lay_vn[i];
    if(vn.x < 0) return;
    const REAL n = (REAL)vn.y;
    if(n_lo > 0) { if((uint32_t)vn.x >= n_lo) lohi[2 * (size_t)i] = 0; else lohi[2 * (size_t)i] += c[vn.x] / n; }
    if(n_hi > 0) { if((uint32_t)vn.x >= n_hi) lohi[2 * (size_t)i + 1] = 0; else lohi[2 * (size_t)i + 1] += c[n_lo + vn.x] / n; }
}

// set_var_cost_func, bdd_cuda_base.cu:425-452
template<typename REAL>
__global__ void set_cost_kernel(const int2* __restrict__ lay_vn, REAL* __restrict__ hi, int var, REAL add, uint32_t n_lay)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n_lay && lay_vn[i].x == var) hi[2 * (size_t)i] += add;   // hi = lohi + 1
}

// distribute_deffered_mm_diff_func, bdd_cuda_base.cu:1396-1414
template<typename REAL>
__global__ void distribute_kernel(const int2* __restrict__ lay_vn, REAL* __restrict__ lohi, REAL* __restrict__ mmd, uint32_t n_lay)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_lay || lay_vn[i].x < 0) return;
    const REAL d = mmd[i];
    if(d > 0) lohi[2 * (size_t)i + 1] += d; else lohi[2 * (size_t)i] -= d;
    mmd[i] = 0;
}

// out[e] = src[stride * ext2lay[e]] for inner layers, `fill` for terminal layers
template<typename T>
__global__ void gather_ext_kernel(const uint32_t* __restrict__ ext2lay, const int32_t* __restrict__ ext_var, const T* __restrict__ src, uint32_t stride, T* __restrict__ out, T fill, uint32_t n_ext)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n_ext) return;
    out[e] = ext_var[e] == INT_MAX ? fill : src[(size_t)stride * ext2lay[e]];
}

template<typename T>
__global__ void scatter_ext_kernel(const uint32_t* __restrict__ ext2lay, const int32_t* __restrict__ ext_var, const T* __restrict__ in, T* __restrict__ dst, uint32_t stride, uint32_t n_ext)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n_ext || ext_var[e] == INT_MAX) return;
    dst[(size_t)stride * ext2lay[e]] = in[e];
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
__global__ void net_costs_kernel(const uint32_t* __restrict__ ext2lay, const int32_t* __restrict__ ext_var, const REAL* __restrict__ lohi,
                                 const REAL* __restrict__ mmd, REAL* __restrict__ out, uint32_t n_ext)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n_ext) return;
    if(ext_var[e] == INT_MAX) { out[e] = 0; return; }
    const uint32_t l = ext2lay[e];
    out[e] = lohi[2 * (size_t)l + 1] - lohi[2 * (size_t)l] + mmd[l];
}

// add_scaled_product_func, bdd_cuda_parallel_mma.h:53-60
template<typename REAL>
__global__ void gradient_step_kernel(const uint32_t* __restrict__ ext2lay, const int32_t* __restrict__ ext_var, REAL* __restrict__ lohi, const REAL* __restrict__ g, REAL step, uint32_t n_ext)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n_ext || ext_var[e] == INT_MAX) return;
    const size_t l = 2 * (size_t)ext2lay[e] + 1;
    lohi[l] = lohi[l] + step * g[e];
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
__global__ void primal_objective_kernel(const uint32_t* __restrict__ var_lay_begin, const uint32_t* __restrict__ var_lay, const REAL* __restrict__ lohi,
                                        double* __restrict__ out, uint32_t n_vars)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if(v >= n_vars) return;
    REAL s = 0;
    for(uint32_t i = var_lay_begin[v]; i < var_lay_begin[v + 1]; ++i) s += lohi[2 * (size_t)var_lay[i] + 1] - lohi[2 * (size_t)var_lay[i]];
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

// ---- multi-GPU exchange over peer memory (NVLink / NVSwitch), SURVEY 8e ------------------------------------------
// One-shot all-reduce of the per-variable min-marginal sums of one pass: every rank's rotating sum buffers live in
// symmetric memory (mapped into every peer), peers[r] is rank r's mapping.  Step 1: tell every peer "my pass `epoch`
// is complete" (a release store of the epoch into slot `rank` of the peer's flag array) and wait until every peer has
// said so here.  Step 2: out[i] = sum over ranks (fixed order 0..world-1, so all ranks compute bit-identical sums) of
// peers[r][offset + i] over the exchanged prefix (the variables shared between shards are numbered first; the passes
// read those from `out` and all other variables from their own sum buffer).
// No trailing barrier: the buffers rotate with period three passes, so a buffer read here is next written (zeroed) two
// passes later, after its owner has waited for this rank's NEXT epoch, which is sent after this kernel.
constexpr int EXCHANGE_MAX_WORLD = 16;

__device__ __forceinline__ void ld_volatile2(const float* p, float& x, float& y) { asm volatile("ld.volatile.global.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "l"(p)); }
__device__ __forceinline__ void ld_volatile2(const double* p, double& x, double& y) { asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "l"(p)); }

// Device-side epoch (graph replay): with `counters` non-null the epoch of this exchange is counters[0] + 1, and the last CTA of the
// launch to finish advances counters[0] (counters[1] counts finished CTAs).  All ranks run the same sequence of exchanges, so their
// counters agree without any host involvement, and a captured launch stays valid however often it is replayed.
__device__ __forceinline__ uint32_t exchange_epoch(const uint32_t* counters, uint32_t epoch_arg)
{
    if(counters == nullptr) return epoch_arg;
    uint32_t e;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(e) : "l"(counters) : "memory");
    return e + 1u;
}
__device__ __forceinline__ void exchange_epoch_done(uint32_t* counters)
{   // called by every thread at the end of the kernel
    if(counters == nullptr) return;
    __syncthreads();
    if(threadIdx.x == 0)
    {
        __threadfence();
        const uint32_t done = atomicAdd(counters + 1, 1u);
        if(done == gridDim.x - 1) { counters[1] = 0; __threadfence(); atomicAdd(counters, 1u); }
    }
}

template<typename REAL>
__global__ void __launch_bounds__(256) delta_exchange_kernel(const REAL* const* __restrict__ peers, uint32_t* const* __restrict__ flags, int world, int rank,
                                                             uint32_t epoch_arg, size_t offset, REAL* __restrict__ out, size_t pairs, uint32_t* counters)
{
    using R2 = typename real2<REAL>::type;
    pdl_wait();                     // launched with programmatic stream serialisation: the pass before must be complete ...
    pdl_launch_dependents();        // ... and the next pass may set itself up while this exchange runs (it waits for its completion)
    const uint32_t epoch = exchange_epoch(counters, epoch_arg);
    __shared__ const REAL* peer_s[EXCHANGE_MAX_WORLD];
    if((int)threadIdx.x < world)
    {
        peer_s[threadIdx.x] = peers[threadIdx.x] + offset;
        if(blockIdx.x == 0)
        {
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flags[threadIdx.x] + rank), "r"(epoch) : "memory");
        }
        exchange_wait_flag(flags[rank] + threadIdx.x, epoch, counters);
    }
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride)
    {
        REAL sx = 0, sy = 0;
        for(int r = 0; r < world; ++r)
        {
            REAL x, y;
            ld_volatile2(peer_s[r] + 2 * i, x, y);
            sx += x; sy += y;
        }
        R2 o; o.x = sx; o.y = sy;
        reinterpret_cast<R2*>(out)[i] = o;
    }
    exchange_epoch_done(counters);
}

// Two-shot variant for many ranks and long prefixes (a one-shot exchange reads (world - 1) x the prefix per rank; this
// one 2 x (world - 1) / world of it): after the same arrival barrier every rank sums ONE slice of the prefix over all ranks
// into its own `out` buffer (symmetric memory as well), tells its peers "slice done" (second flag array, sent by the last
// CTA to finish), waits for theirs and copies the other slices from the peers' `out` buffers.  All CTAs of the launch must
// be co-resident (they wait for each other): the host launches at most 4 x 256 threads per SM.
template<typename REAL>
__global__ void __launch_bounds__(256) delta_exchange2_kernel(const REAL* const* __restrict__ peers, REAL* const* __restrict__ outs, uint32_t* const* __restrict__ flags,
                                                              int world, int rank, uint32_t epoch_arg, size_t offset, size_t pairs, uint32_t* counters)
{
    using R2 = typename real2<REAL>::type;
    pdl_wait();                     // launched with programmatic stream serialisation: the pass before must be complete ...
    pdl_launch_dependents();        // ... and the next pass may set itself up while this exchange runs (it waits for its completion)
    const uint32_t epoch = exchange_epoch(counters, epoch_arg);
    __shared__ const REAL* peer_s[EXCHANGE_MAX_WORLD];
    __shared__ REAL* out_s[EXCHANGE_MAX_WORLD];
    __shared__ bool last_cta;
    uint32_t* my_flags = flags[rank];
    if((int)threadIdx.x < world)
    {
        peer_s[threadIdx.x] = peers[threadIdx.x] + offset;
        out_s[threadIdx.x] = outs[threadIdx.x];
        if(blockIdx.x == 0)
        {
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flags[threadIdx.x] + rank), "r"(epoch) : "memory");
        }
        exchange_wait_flag(my_flags + threadIdx.x, epoch, counters);
    }
    __syncthreads();
    const size_t per = (pairs + world - 1) / world;
    const size_t stride = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    {   // shot 1: my slice, summed over all ranks in rank order
        const size_t lo = min(pairs, per * rank), hi = min(pairs, lo + per);
        REAL* out = out_s[rank];
        for(size_t i = lo + tid; i < hi; i += stride)
        {
            REAL sx = 0, sy = 0;
            for(int r = 0; r < world; ++r)
            {
                REAL x, y;
                ld_volatile2(peer_s[r] + 2 * i, x, y);
                sx += x; sy += y;
            }
            R2 o; o.x = sx; o.y = sy;
            reinterpret_cast<R2*>(out)[i] = o;
        }
    }
    // "slice done": the last CTA of this launch to get here tells every peer
    __threadfence_system();
    __syncthreads();
    if(threadIdx.x == 0)
    {
        const uint32_t done = atomicAdd(my_flags + 48, 1u);
        last_cta = done == gridDim.x - 1;
        if(last_cta) my_flags[48] = 0;
    }
    __syncthreads();
    if(last_cta && (int)threadIdx.x < world)
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flags[threadIdx.x] + 16 + rank), "r"(epoch) : "memory");
    if((int)threadIdx.x < world) exchange_wait_flag(my_flags + 16 + threadIdx.x, epoch, counters);
    __syncthreads();
    // shot 2: the other ranks' slices, copied from their `out` buffers
    REAL* out = out_s[rank];
    for(int q = 1; q < world; ++q)
    {
        const int r = (rank + q) % world;
        const size_t lo = min(pairs, per * r), hi = min(pairs, lo + per);
        for(size_t i = lo + tid; i < hi; i += stride)
        {
            R2 o;
            ld_volatile2(out_s[r] + 2 * i, o.x, o.y);
            reinterpret_cast<R2*>(out)[i] = o;
        }
    }
    exchange_epoch_done(counters);
}

// In-switch variant (NVLink SHARP / NVLS): the rotating sum buffers and the result buffer are symmetric memory with a multicast
// mapping.  After the arrival barrier every rank reduces ONE slice of the prefix with multimem.ld_reduce (the switch adds the
// ranks' values and returns the sum: one read of the slice instead of world of them) and broadcasts it into every rank's result
// buffer with multimem.st; a second barrier tells everyone that all slices are in place.  One rank computes each slice, so all
// ranks see bit-identical sums.  All CTAs of the launch must be co-resident (they wait for each other).
__device__ __forceinline__ void mc_reduce_store(const float* mc_in, float* mc_out)
{   // 16 bytes = two {lo, hi} pairs
    float a, b, c, d;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(mc_in) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(mc_out), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mc_reduce_store(const double* mc_in, double* mc_out)
{   // 16 bytes = one {lo, hi} pair
    double a, b;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(a) : "l"(mc_in) : "memory");
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(b) : "l"(mc_in + 1) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" :: "l"(mc_out), "d"(a) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" :: "l"(mc_out + 1), "d"(b) : "memory");
}

template<typename REAL>
__global__ void __launch_bounds__(256) delta_exchange_mc_kernel(const REAL* __restrict__ mc_in, REAL* __restrict__ mc_out, uint32_t* const* __restrict__ flags,
                                                                int world, int rank, size_t offset, size_t n_exchange, uint32_t* counters)
{
    pdl_wait();
    pdl_launch_dependents();
    const uint32_t epoch = exchange_epoch(counters, 0u);
    __shared__ bool last_cta;
    uint32_t* my_flags = flags[rank];
    auto wait_all = [&](uint32_t slot0) {
        if((int)threadIdx.x < world) exchange_wait_flag(my_flags + slot0 + threadIdx.x, epoch, counters);
        __syncthreads();
    };
    if(blockIdx.x == 0 && (int)threadIdx.x < world)
    {   // "my pass is complete" to every peer
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flags[threadIdx.x] + rank), "r"(epoch) : "memory");
    }
    wait_all(0);
    // my slice of the prefix, in 16-byte units
    constexpr size_t PER16 = 16 / sizeof(REAL);
    const size_t units = (n_exchange + PER16 - 1) / PER16, per = (units + world - 1) / world;
    const size_t lo = min(units, per * (size_t)rank), hi = min(units, lo + per);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for(size_t u = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; u < hi; u += stride)
        mc_reduce_store(mc_in + offset + u * PER16, mc_out + u * PER16);
    // "my slice is everywhere": the last CTA of this launch to get here tells every peer
    __threadfence_system();
    __syncthreads();
    if(threadIdx.x == 0)
    {
        const uint32_t done = atomicAdd(my_flags + 48, 1u);
        last_cta = done == gridDim.x - 1;
        if(last_cta) my_flags[48] = 0;
    }
    __syncthreads();
    if(last_cta && (int)threadIdx.x < world)
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flags[threadIdx.x] + 16 + rank), "r"(epoch) : "memory");
    wait_all(16);
    exchange_epoch_done(counters);
}

// ---- primal rounding: one perturbation round of incremental_mm_agreement_rounding_cuda ---------------------------
// (src/bdd_solver/incremental_mm_agreement_rounding_cuda.cu: mm_diff_direction_func :28-42, compute_mm_types :79-110,
// compute_mm_sums :124-146, mm_types_transform :148-212).  One thread per variable walks the variable's layers (BDD
// order): direction of every min-marginal pair (-1: mm_0 + 1e-6 <= mm_1, +1: mm_1 + 1e-6 <= mm_0, else 0), type of the
// variable from the min / max direction (one / zero / equal / inconsistent), sums of mm_0 and mm_1, and the cost
// perturbation: one -> (delta, 0), zero -> (0, delta), otherwise |r| * delta on the side the sums (or the sign of r)
// point away from, r uniform in (-delta, delta).  The reference draws r from thrust::default_random_engine
// (minstd_rand, x -> 48271 x mod 2^31 - 1, seed 1) discarded by (thread id + round); here the discard count is
// (variable + round), computed by modular exponentiation, so the perturbation is a pure function of (variable, round).
enum RoundingType { MM_ZERO = 0, MM_ONE = 1, MM_EQUAL = 2, MM_INCONSISTENT = 3 };

__device__ __forceinline__ uint32_t minstd_after(uint32_t n)
{   // state after n + 1 steps from seed 1: 48271^(n+1) mod (2^31 - 1)
    const uint64_t M = 2147483647ull;
    uint64_t result = 1, base = 48271ull, e = (uint64_t)n + 1ull;
    while(e) { if(e & 1ull) result = (result * base) % M; base = (base * base) % M; e >>= 1; }
    return (uint32_t)result;
}

template<typename REAL>
__global__ void rounding_kernel(const uint32_t* __restrict__ var_lay_begin, const uint32_t* __restrict__ var_lay, const REAL* __restrict__ mm_lo, const REAL* __restrict__ mm_hi,
                                double delta, uint32_t round_index, REAL* __restrict__ cost_delta_0, REAL* __restrict__ cost_delta_1, char* __restrict__ types,
                                unsigned long long* __restrict__ counts, uint32_t n_vars)
{
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if(v >= n_vars) return;
    int mn = 2, mx = -2;
    REAL s0 = 0, s1 = 0;
    for(uint32_t e = var_lay_begin[v]; e < var_lay_begin[v + 1]; ++e)
    {
        const uint32_t l = var_lay[e];
        const REAL m0 = mm_lo[l], m1 = mm_hi[l];
        const int dir = ((double)m0 + 1e-6 <= (double)m1) ? -1 : (((double)m1 + 1e-6 <= (double)m0) ? 1 : 0);
        mn = min(mn, dir); mx = max(mx, dir);
        s0 += m0; s1 += m1;
    }
    int type;
    if(mn == 2) type = MM_ZERO;               // variable in no BDD (cannot happen in the reference: bdd_cuda_base.cu:471): free, keep 0
    else if(mn > 0) type = MM_ONE;
    else if(mx < 0) type = MM_ZERO;
    else if(mx == 0 && mn == 0) type = MM_EQUAL;
    else type = MM_INCONSISTENT;
    types[v] = (char)type;
    atomicAdd(counts + type, 1ull);
    REAL d0 = 0, d1 = 0;
    if(type == MM_ONE) d0 = (REAL)delta;
    else if(type == MM_ZERO) { if(mn != 2) d1 = (REAL)delta; }
    else
    {
        const uint32_t x = minstd_after(v + round_index);
        const float u = (float)(x - 1u) / 2147483646.0f;                 // [0, 1)
        const float r = __fadd_rn((float)(-delta), __fmul_rn(u, (float)(2.0 * delta)));      // uniform_real_distribution<float>(-delta, delta), no fma contraction
        const REAL amount = (REAL)((double)fabsf(r) * delta);
        if(type == MM_EQUAL) { if(r < 0.0f) d0 = amount; else d1 = amount; }
        else { if(s0 < s1) d1 = amount; else d0 = amount; }
    }
    cost_delta_0[v] = d0; cost_delta_1[v] = d1;
}

// compute_bdd_sol_func, bdd_cuda_base.cu:1103-1135: per BDD follow the cheaper arc from the
// root (tie -> hi), comparing hi_path - lo_path > 0 with path = cfr + (cft[child] + cost).
template<typename REAL>
__global__ void bdds_solution_kernel(const BundleDesc* __restrict__ bundles, const HopRec* __restrict__ hops, const uint32_t* __restrict__ topo,
                                     const uint32_t* __restrict__ bdd_bundle, const uint32_t* __restrict__ bdd_ext_begin,
                                     const REAL* __restrict__ cfr, const REAL* __restrict__ cft, const REAL* __restrict__ lohi,
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
    if(bd.cls == CLS_LANE)
    {   // lane-local class: one-hot arc targets per (hop, lane), the BDD stays in lane q
        const uint32_t J = bd.max_J, mask = (1u << J) - 1u;
        uint32_t r = 0;
        for(uint32_t k = 0; k < n_lay; ++k)
        {
            const uint32_t word = topo[bd.topo_base + k * 32u + q];
            const uint32_t s = hops[bd.hop_base + k].node_off + r * 32u + q, sn = hops[bd.hop_base + k + 1].node_off + q;
            const uint32_t lay = bd.layer_base + k * 32u + q;
            const uint32_t lo_bits = (word >> (r * 2u * J)) & mask, hi_bits = (word >> (r * 2u * J + J)) & mask;
            const uint32_t lo_r = lo_bits ? (uint32_t)__ffs((int)lo_bits) - 1u : 0u, hi_r = hi_bits ? (uint32_t)__ffs((int)hi_bits) - 1u : 0u;
            const REAL c = cfr[s];
            const REAL lo_path = c + ((lo_bits ? cft[sn + lo_r * 32u] : INF) + lohi[2 * (size_t)lay]);
            const REAL hi_path = c + ((hi_bits ? cft[sn + hi_r * 32u] : INF) + lohi[2 * (size_t)lay + 1]);
            const bool take_lo = (hi_path - lo_path > 0);
            sol[e0 + k] = take_lo ? 0 : 1;
            r = take_lo ? lo_r : hi_r;
        }
        sol[e0 + n_lay] = 0;   // terminal layer
        return;
    }
    for(uint32_t k = 0; k < n_lay; ++k)
    {
        const HopRec h = hops[bd.hop_base + k], hn = hops[bd.hop_base + k + 1];
        const uint32_t s = h.node_off + ts;
        const uint32_t t = topo[s];
        const uint32_t lay = bd.layer_base + k * bpw + q;
        const uint32_t lo = t & 0xFFFFu, hi = t >> 16;
        const REAL c = cfr[s];
        const REAL lo_path = c + ((lo == CHILD_BOT ? INF : cft[hn.node_off + lo]) + lohi[2 * (size_t)lay]);
        const REAL hi_path = c + ((hi == CHILD_BOT ? INF : cft[hn.node_off + hi]) + lohi[2 * (size_t)lay + 1]);
        const bool take_lo = (hi_path - lo_path > 0);
        sol[e0 + k] = take_lo ? 0 : 1;
        ts = take_lo ? lo : hi;
    }
    sol[e0 + n_lay] = 0;   // terminal layer
}

} // namespace bddb200
