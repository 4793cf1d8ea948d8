// bdd_b200/csrc/layout.hpp -- host-side layout builder: BDD::bdd_collection (flat
// instruction array) -> the bundle/chunk/hop-major SoA the sm_100a sweep kernels stream.
//
// Replaces the reference's constructor chain bdd_cuda_base.cu:31-46 (initialize,
// populate_bdd_nodes, reorder_bdd_nodes, compress_bdd_nodes_to_layer,
// reorder_within_bdd_layers, set_special_nodes_indices, find_primal_variable_ordering;
// SURVEY 3.2), which sorts all nodes of all BDDs by hop with thrust and launches one kernel
// per hop.  Here the unit of work is a *bundle*: 32/P BDDs processed by ONE warp, P lanes
// per BDD, swept hop by hop without any grid- or block-level synchronisation (within a
// pass BDDs are independent).
//
// Memory layout (all arrays indexed by "slot" or "layer entry"):
//   A bundle's hops are cut into *chunks* of consecutive hops; inside a chunk every hop owns a
//   tile of 32 * J node slots (J uniform per chunk), tiles and chunks of a bundle are
//   contiguous.  One chunk is what the kernel stages into one shared-memory pipeline stage
//   with a handful of bulk-async copies (every per-slot / per-layer array of a chunk is one
//   contiguous, 16-byte aligned range).
//   slot (j, lane) of a tile = tile_off + j*32 + lane holds, for the BDD owning that lane group
//   (bdd_local = lane >> logP), the node with index  c = j*P + (lane & (P-1))  of that BDD's
//   layer.  A warp therefore touches every per-node array in fully coalesced 128-byte
//   rows, and a BDD's frontier stays in the same shared-memory banks from hop to hop.
//   topo[slot] = lo_child | hi_child << 16, children given as slot index inside the NEXT
//   hop's tile (0xFFFF = arc into the bot sink: value +inf, no memory access);
//   TOPO_TOP marks the BDD's top sink (cost_from_terminal 0), TOPO_PAD an unused slot.
//   Layer entry (g, k, bdd_local) = layer_base + k*(32/P) + bdd_local holds the layer's
//   variable, the (global) number of BDDs of that variable, lo/hi arc cost (interleaved) and
//   the deferred min-marginal difference.  layer_base is even so that 8-byte entries of a
//   chunk start 16-byte aligned.
#pragma once

#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstddef>
#include <cstdint>
#include <numeric>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/bdd_b200.h"

namespace bddb200 {

constexpr uint32_t TOPO_PAD = 0xFFFFFFFFu;
constexpr uint32_t TOPO_TOP = 0xFFFFFFFEu;
constexpr uint32_t CHILD_BOT = 0xFFFFu;
constexpr uint32_t MAX_TILE_SLOTS = 0xFFFFu;      // child slot indices are 16 bit
constexpr uint32_t SMALL_CLASS_MAX_J = 8;         // bundles with J <= 8 share multi-warp CTAs
constexpr uint32_t LANE_MAX_J = 4;                // lane-local class: one lane per BDD, at most 4 nodes per layer
constexpr int32_t LAY_NONE = -1;                  // layer entry without a layer (BDD shorter than its bundle)
constexpr int32_t LAY_TOP = -2;                   // layer entry of a BDD's terminal hop (its top sink sits in row 0)

struct HopRec {
    uint32_t node_off;   // first slot of the tile
    uint32_t J;          // rows of 32 slots
};

// One pipeline stage worth of consecutive hops of a bundle.
struct ChunkRec {
    uint32_t slot_off;   // first slot of the chunk's first tile
    uint32_t lay_off;    // first layer entry (even)
    uint32_t hop_first;  // first hop of the chunk inside its bundle
    uint32_t n_hops;
    uint32_t J;          // rows of 32 slots per tile, uniform inside the chunk
    uint32_t J_next;     // J of the following chunk (0: last chunk)
    uint32_t pad_[2];
};

struct BundleDesc {
    uint32_t hop_base;   // index of the bundle's first HopRec
    uint32_t n_hops;     // hops incl. the terminal hop of the longest BDD
    uint32_t layer_base; // first layer entry (even)
    uint32_t logP;       // log2(lanes per BDD)
    uint32_t bdd_base;   // first entry in bundle_bdd (32 >> logP entries)
    uint32_t max_J;
    uint32_t chunk_base; // index of the bundle's first ChunkRec
    uint32_t n_chunks;
    uint32_t cls;        // CLS_GENERIC or CLS_LANE
    uint32_t topo_base;  // CLS_LANE: first topology word (32 per hop)
};
enum BundleClass { CLS_GENERIC = 0, CLS_LANE = 1 };

// Lane-local class (CLS_LANE): one lane per BDD and at most LANE_MAX_J nodes per layer, so a
// BDD never leaves its lane.  Every hop of the bundle owns a tile of J rows (J uniform per
// bundle): any range of hops is contiguous in every array and chunking is a launch parameter.
// The topology of one (hop, lane) is ONE word of one-hot arc targets:
//   bit (j*2J + a*J + r) set  <=>  arc a (0 = lo, 1 = hi) of the node in row j enters row r of the
//   next hop's tile; no bit = arc into the bot sink (or no node in row j).
// The terminal hop of a BDD is marked in its layer entry (LAY_TOP).
struct LaneDesc {
    uint32_t slot_off;   // first slot of the bundle (hop h, row j, lane l -> slot_off + (h*J + j)*32 + l)
    uint32_t lay_off;    // first layer entry (hop h, lane l -> lay_off + h*32 + l)
    uint32_t topo_off;   // first topology word (hop h, lane l -> topo_off + h*32 + l)
    uint32_t n_hops;     // hops incl. the terminal hop of the longest BDD
    uint32_t J;
    uint32_t bdd_base;   // first entry in bundle_bdd (32 entries)
    uint32_t pad_[2];
};
inline uint32_t lane_arc_bit(uint32_t J, uint32_t j, uint32_t arc, uint32_t r) { return 1u << (j * 2 * J + arc * J + r); }

// Shared-memory bytes one hop of a lane-class bundle occupies in a pipeline stage:
// topology word + the opposite direction's DP rows + {var, nr_bdds} + {lo, hi} + gathered {delta_lo, delta_hi}
inline size_t lane_hop_bytes(uint32_t J, size_t R) { return 128 + (size_t)J * 32 * R + 256 + 64 * R + 64 * R; }

// Descriptor block of a bundle: DESC_WORDS 32-bit words that a warp fetches with ONE coalesced
// 128-byte load (lane l reads word l).  There is one block per bundle and pass direction; the
// chunk records are listed in the order the pass visits them (forward: first chunk first,
// backward: last chunk first), so the first DESC_CHUNKS pipeline stages can be started
// without a second, dependent global load.
constexpr int DESC_WORDS = 32;
enum DescWord { DESC_N_CHUNKS = 0, DESC_LOGP = 1, DESC_BDD_BASE = 2, DESC_MAX_J = 3, DESC_CHUNK_BASE = 4,
                DESC_N_HOPS = 5, DESC_LAYER_BASE = 6, DESC_FIRST_CHUNK = 7 };
constexpr int DESC_CHUNK_WORDS = 5;   // {slot_off, lay_off, n_hops, J, J_next}
constexpr int DESC_CHUNKS = 5;

// Shared-memory bytes one chunk occupies in a pipeline stage (worst case over the forward
// and the backward kernel): topo + the opposite direction's DP values + {var, nr_bdds} +
// {lo, hi} + gathered {delta_lo, delta_hi}.
inline size_t chunk_stage_bytes(uint32_t n_hops, uint32_t J, uint32_t J_next, uint32_t bpw, size_t R)
{
    const size_t ne = ((size_t)n_hops * bpw + 1) & ~(size_t)1;
    const size_t dp_rows = std::max<size_t>((size_t)n_hops * J, (size_t)(n_hops - 1) * J + J_next);
    return (size_t)n_hops * J * 128 + dp_rows * 32 * R + ne * (8 + 4 * R);
}

struct layout_error : std::runtime_error {
    int code;
    layout_error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// Large host arrays: std::vector value-initialises its elements on one thread, and for arrays of tens of megabytes that first touch
// (page faults) costs more than the work that fills them (20 M nodes: 130 ms of 350).  uvec<T> leaves new elements uninitialised
// (trivial T only); par_fill touches and fills them from all threads.
template<typename T>
struct default_init_allocator : std::allocator<T> {
    template<typename U> struct rebind { using other = default_init_allocator<U>; };
    using std::allocator<T>::allocator;
    template<typename U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new(static_cast<void*>(p)) U; }
    template<typename U, typename... Args> void construct(U* p, Args&&... args) { ::new(static_cast<void*>(p)) U(std::forward<Args>(args)...); }
};
template<typename T> using uvec = std::vector<T, default_init_allocator<T>>;
template<typename T>
inline void par_fill(uvec<T>& v, size_t n, T value)
{
    v.clear();
    v.resize(n);              // no touch
    T* p = v.data();
#pragma omp parallel for schedule(static)
    for(long long i = 0; i < (long long)n; ++i) p[i] = value;
}

struct HostLayout {
    size_t n_vars = 0, n_bdds = 0, n_instr = 0;
    size_t n_layers_ext = 0;   // sum over BDDs of (nr variables + 1)
    size_t n_real_nodes = 0;   // non-terminal nodes
    size_t n_slots = 0, n_lay = 0, max_hops = 0;
    size_t n_lane_bundles = 0;           // bundles [0, n_lane): lane-local class
    size_t n_lane_shared_bundles = 0;    // shard mode: lane-class bundles [0, n_lane_shared) contain a BDD with a variable shared between shards
    uint32_t lane_max_J = 0, lane_max_hops = 0;
    size_t n_generic_slots = 0;          // generic bundles own slots [0, n_generic_slots): topo[slot] is their topology word
    size_t n_topo = 0;                   // n_generic_slots + 32 words per hop of every lane-class bundle
    std::vector<LaneDesc> desc_lane;     // one per lane-class bundle
    size_t n_small_bundles = 0;          // generic bundles [n_lane, n_lane + n_small): every chunk fits the stage budget
    uint32_t max_tile_small = 0, max_tile_large = 0;  // slots
    size_t stage_small = 0, stage_large = 0;          // largest chunk_stage_bytes per class
    std::vector<BundleDesc> bundles;
    std::vector<ChunkRec> chunks;
    std::vector<uint32_t> desc_fwd, desc_bwd;   // DESC_WORDS per bundle
    std::vector<HopRec> hops;
    std::vector<int32_t> bundle_bdd;     // per bundle lane group: external BDD index or -1
    uvec<uint32_t> topo;          // per slot
    uvec<int32_t> lay_var;        // per layer entry: variable or -1
    uvec<uint32_t> ext2lay;       // external layer -> layer entry
    uvec<int32_t> ext_var;        // external layer -> variable (INT_MAX terminal)
    uvec<int32_t> ext_bdd;        // external layer -> BDD
    std::vector<uint32_t> bdd_ext_begin; // per BDD: first external layer (n_bdds+1)
    std::vector<uint32_t> root_slot, top_slot;  // per BDD
    std::vector<int32_t> nr_bdds_per_var;       // counted from this collection
    // variable -> layer entries in (variable, BDD index) order (deterministic delta sums,
    // make_dual_feasible, sorted min-marginals)
    std::vector<uint32_t> var_lay_begin; // n_vars+1
    uvec<uint32_t> var_lay;       // layer entries
    uvec<uint32_t> sorted_ext;    // external layers sorted by (var, bdd), terminals last
};

// BDDB200_LAYOUT_TIMING=1: phase durations of build_layout on stderr
struct LayoutTimer {
    bool on; std::chrono::steady_clock::time_point t;
    LayoutTimer() : on(std::getenv("BDDB200_LAYOUT_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char* what)
    {
        if(!on) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[bdd_b200 layout] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

inline uint32_t pow2ceil(uint32_t x) { uint32_t p = 1; while(p < x) p <<= 1; return p; }
inline uint32_t ilog2(uint32_t x) { uint32_t l = 0; while((1u << l) < x) ++l; return l; }

constexpr size_t DEFAULT_STAGE_BUDGET = 12 * 1024;   // bytes of one pipeline stage of the small class

// lanes_per_bdd: 0 = heuristic (about <= 4 nodes of a layer per lane), else forced.
// real_bytes: sizeof(REAL) of the solver (chunk sizes depend on it); stage_budget: shared-memory
// bytes one chunk of a small-class bundle may occupy.
inline HostLayout build_layout(const bddb200_instruction* instrs, size_t n_instr,
                               const size_t* delims, size_t n_bdds, int lanes_per_bdd,
                               size_t nr_variables_override = 0, size_t real_bytes = 4,
                               size_t stage_budget = DEFAULT_STAGE_BUDGET, bool lane_class = true, size_t n_sms = 0, size_t n_shared_vars = 0)
{
    constexpr size_t TOPSINK = (size_t)-1, BOTSINK = (size_t)-1 - 1;
    if(n_bdds == 0) throw layout_error(BDDB200_ERR_INVALID_ARGUMENT, "empty BDD collection");
    if(lanes_per_bdd != 0 && (lanes_per_bdd < 1 || lanes_per_bdd > 32 || (lanes_per_bdd & (lanes_per_bdd - 1))))
        throw layout_error(BDDB200_ERR_INVALID_ARGUMENT, "lanes_per_bdd must be 0 or a power of two <= 32");
    if(delims[n_bdds] > n_instr) throw layout_error(BDDB200_ERR_INVALID_ARGUMENT, "delimiters exceed instruction array");

    HostLayout L;
    L.n_bdds = n_bdds; L.n_instr = n_instr;
    LayoutTimer timer;

    // ---- pass 1: layers of every BDD, validation, widths ------------------------------
    L.bdd_ext_begin.assign(n_bdds + 1, 0);
    uvec<uint32_t> ext_first_instr;   // per external layer: first instruction; layer e spans [efi[e], efi[e+1])
    std::vector<uint32_t> bdd_maxw(n_bdds, 1);
    size_t max_var = 0;
    // errors found inside the parallel loops: the BDD with the smallest index reports (what a sequential scan would have found)
    size_t err_bdd = (size_t)-1; int err_code = 0; std::string err_msg;
    auto report = [&](size_t b, int code, const std::string& msg) {
#pragma omp critical(bddb200_layout_error)
        if(b < err_bdd) { err_bdd = b; err_code = code; err_msg = msg; }
    };
    // (a) layers per BDD
    std::vector<uint32_t> n_ext_of(n_bdds, 0);
    // shard mode: does the BDD contain a variable that other shards contain too (variables [0, n_shared_vars))?  Those BDDs are kept
    // together when bundles are formed, so that few bundles take part in the multi-GPU flag barrier of a pass
    std::vector<uint8_t> bdd_shared(n_shared_vars > 0 ? n_bdds : 0, 0);
    size_t n_real = 0;
#pragma omp parallel for schedule(static) reduction(max : max_var) reduction(+ : n_real)
    for(long long bb = 0; bb < (long long)n_bdds; ++bb)
    {
        const size_t b = (size_t)bb;
        const size_t first = delims[b], last = delims[b+1];
        if(last < first + 3) { report(b, BDDB200_ERR_INVALID_ARGUMENT, "BDD " + std::to_string(b) + " has no inner node"); continue; }
        // the two sinks close every BDD, in either order (bdd_collection.cpp:403-428 vs :1581-1586)
        if(!((instrs[last-2].index == BOTSINK && instrs[last-1].index == TOPSINK) || (instrs[last-2].index == TOPSINK && instrs[last-1].index == BOTSINK)))
        { report(b, BDDB200_ERR_INVALID_ARGUMENT, "BDD " + std::to_string(b) + ": last two instructions must be the bot and top sink"); continue; }
        size_t prev = TOPSINK;
        uint32_t cnt = 0;
        for(size_t i = first; i + 2 < last; ++i)
        {
            const size_t var = instrs[i].index;
            if(var >= BOTSINK) { report(b, BDDB200_ERR_INVALID_ARGUMENT, "terminal instruction inside BDD " + std::to_string(b)); break; }
            if(var != prev) { ++cnt; prev = var; if(var > max_var) max_var = var; if(var < n_shared_vars) bdd_shared[b] = 1; }
        }
        n_ext_of[b] = cnt + 1;                               // + the terminal layer entry
        n_real += last - first - 2;
    }
    if(err_code != 0) throw layout_error(err_code, err_msg);
    L.n_real_nodes = n_real;
    for(size_t b = 0; b < n_bdds; ++b) L.bdd_ext_begin[b + 1] = L.bdd_ext_begin[b] + n_ext_of[b];
    L.n_layers_ext = L.bdd_ext_begin[n_bdds];
    ext_first_instr.resize(L.n_layers_ext); L.ext_var.resize(L.n_layers_ext); L.ext_bdd.resize(L.n_layers_ext);
    // (b) fill
#pragma omp parallel for schedule(static)
    for(long long bb = 0; bb < (long long)n_bdds; ++bb)
    {
        const size_t b = (size_t)bb;
        const size_t first = delims[b], last = delims[b+1];
        uint32_t e = L.bdd_ext_begin[b];
        size_t prev = TOPSINK;
        for(size_t i = first; i + 2 < last; ++i)
        {
            const size_t var = instrs[i].index;
            if(var != prev)
            {
                ext_first_instr[e] = (uint32_t)i; L.ext_var[e] = (int32_t)var; L.ext_bdd[e] = (int32_t)b;
                ++e; prev = var;
            }
        }
        ext_first_instr[e] = (uint32_t)(last - 2);            // terminal layer entry = end of the inner nodes
        L.ext_var[e] = INT_MAX; L.ext_bdd[e] = (int32_t)b;
    }
    L.n_vars = std::max(max_var + 1, nr_variables_override);
    if(L.n_vars > (size_t)INT_MAX / 2) throw layout_error(BDDB200_ERR_INVALID_ARGUMENT, "too many variables");

    timer.lap("layers of every BDD");
    // QBDD check (reference: assert(is_qbdd && is_reordered), bdd_cuda_base.cu:100-101):
    // every arc of layer k enters layer k+1 or the bot sink; the top sink only from the last layer.
#pragma omp parallel for schedule(static)
    for(long long bb = 0; bb < (long long)n_bdds; ++bb)
    {
        const size_t b = (size_t)bb;
        const size_t last = delims[b+1];
        const size_t bot = instrs[last-2].index == BOTSINK ? last - 2 : last - 1, top = instrs[last-2].index == BOTSINK ? last - 1 : last - 2;
        const uint32_t eb = L.bdd_ext_begin[b], ee = L.bdd_ext_begin[b+1] - 1; // ee = terminal layer
        uint32_t maxw = 1;
        bool bad = false;
        for(uint32_t e = eb; e < ee && !bad; ++e)
        {
            const size_t lb = ext_first_instr[e], le = ext_first_instr[e+1];
            const bool last_layer = (e + 1 == ee);
            const size_t ne = last_layer ? le : ext_first_instr[e+2];   // next layer = [le, ne)
            maxw = std::max<uint32_t>(maxw, (uint32_t)(le - lb));
            for(size_t i = lb; i < le; ++i)
            {
                for(const size_t c : {instrs[i].lo, instrs[i].hi})
                {
                    if(c == bot) continue;
                    if(last_layer ? (c != top) : !(c >= le && c < ne)) bad = true;
                }
            }
        }
        if(bad) report(b, BDDB200_ERR_NOT_QBDD, "BDD " + std::to_string(b) + " is not a reordered QBDD (arc skips a layer)");
        bdd_maxw[b] = maxw;
    }
    if(err_code != 0) throw layout_error(err_code, err_msg);

    timer.lap("QBDD check, widths");
    // ---- lanes per BDD, sort, bundles --------------------------------------------------
    std::vector<uint8_t> bdd_logP(n_bdds);
    for(size_t b = 0; b < n_bdds; ++b)
    {
        uint32_t P = lanes_per_bdd ? (uint32_t)lanes_per_bdd : std::min<uint32_t>(32, pow2ceil((bdd_maxw[b] + 3) / 4));
        bdd_logP[b] = (uint8_t)ilog2(P);
    }
    auto nlay = [&](uint32_t b) { return L.bdd_ext_begin[b+1] - L.bdd_ext_begin[b]; };
    // lane-local class: one lane per BDD, narrow; everything else is generic
    std::vector<uint32_t> order, lane_order;
    for(size_t b = 0; b < n_bdds; ++b)
    {
        if(lane_class && bdd_logP[b] == 0 && bdd_maxw[b] <= LANE_MAX_J) lane_order.push_back((uint32_t)b);
        else order.push_back((uint32_t)b);
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        if(bdd_logP[x] != bdd_logP[y]) return bdd_logP[x] < bdd_logP[y];
        return nlay(x) > nlay(y);
    });
    {   // by (width ascending, number of layers descending), stable: a counting sort when the key range is small (it always is for
        // lane-class BDDs: width <= LANE_MAX_J), else a comparison sort
        uint32_t max_lay = 0;
        for(const uint32_t b : lane_order) max_lay = std::max(max_lay, nlay(b));
        const size_t n_keys = (size_t)(LANE_MAX_J + 1) * (max_lay + 1) * 2;
        if(n_keys <= ((size_t)1 << 24))
        {   // by (width, BDDs with a variable shared between shards first, length descending)
            auto key = [&](uint32_t b) { return ((size_t)bdd_maxw[b] * 2 + (!bdd_shared.empty() && bdd_shared[b] ? 0 : 1)) * (max_lay + 1) + (max_lay - nlay(b)); };
            std::vector<uint32_t> count(n_keys + 1, 0);
            for(const uint32_t b : lane_order) count[key(b) + 1]++;
            for(size_t k = 0; k < n_keys; ++k) count[k + 1] += count[k];
            std::vector<uint32_t> sorted(lane_order.size());
            for(const uint32_t b : lane_order) sorted[count[key(b)]++] = b;
            lane_order.swap(sorted);
        }
        else
            std::stable_sort(lane_order.begin(), lane_order.end(), [&](uint32_t x, uint32_t y) {
                if(bdd_maxw[x] != bdd_maxw[y]) return bdd_maxw[x] < bdd_maxw[y];
                if(!bdd_shared.empty() && bdd_shared[x] != bdd_shared[y]) return bdd_shared[x] > bdd_shared[y];
                return nlay(x) > nlay(y);
            });
    }
    const size_t n_generic_bdds = order.size();

    struct ProtoChunk { uint32_t first, n, J; };
    struct ProtoBundle { uint32_t first, count, logP, n_hops, max_J, work; bool small; std::vector<uint32_t> J; std::vector<ProtoChunk> chunks; };
    std::vector<ProtoBundle> protos;
    for(size_t pos = 0; pos < n_generic_bdds;)
    {
        const uint32_t logP = bdd_logP[order[pos]];
        const uint32_t bpw = 32u >> logP, P = 1u << logP;
        size_t end = pos;
        while(end < n_generic_bdds && end - pos < bpw && bdd_logP[order[end]] == logP) ++end;
        ProtoBundle pb;
        pb.first = (uint32_t)pos; pb.count = (uint32_t)(end - pos); pb.logP = logP;
        pb.n_hops = 0;
        for(size_t q = pos; q < end; ++q) pb.n_hops = std::max(pb.n_hops, nlay(order[q]));
        pb.J.assign(pb.n_hops, 0);
        for(size_t q = pos; q < end; ++q)
        {
            const uint32_t b = order[q];
            const uint32_t eb = L.bdd_ext_begin[b], ee = L.bdd_ext_begin[b+1];
            for(uint32_t e = eb; e < ee; ++e)
            {
                const uint32_t w = (e + 1 < ee) ? (ext_first_instr[e+1] - ext_first_instr[e]) : 1u;   // terminal layer: the top sink
                pb.J[e - eb] = std::max(pb.J[e - eb], (w + P - 1) / P);
            }
        }
        pb.max_J = *std::max_element(pb.J.begin(), pb.J.end());
        if(pb.max_J * 32u > MAX_TILE_SLOTS)
            throw layout_error(BDDB200_ERR_TOO_WIDE, "a BDD layer is wider than " + std::to_string(MAX_TILE_SLOTS) + " nodes; split the BDD (split_qbdd)");

        // Cut the hops into chunks, back to front (so that the J of the following chunk is
        // known): a chunk grows towards the root while it fits the stage budget and padding
        // its tiles to a common J wastes at most a quarter of its rows (+2).  With one BDD
        // per warp (bpw == 1) chunks start at even hops to keep layer ranges 16-byte aligned.
        {
            uint32_t e = pb.n_hops;          // exclusive end of the chunk being formed
            uint32_t J_next = 0;
            std::vector<ProtoChunk> rev;
            while(e > 0)
            {
                uint32_t a = e - 1, Jc = pb.J[a], real = pb.J[a];
                uint32_t best_a = a, best_J = Jc;
                while(a > 0)
                {
                    const uint32_t na = a - 1;
                    const uint32_t nJ = std::max(Jc, pb.J[na]);
                    const uint32_t nreal = real + pb.J[na];
                    const uint32_t n = e - na;
                    if(chunk_stage_bytes(n, nJ, J_next, bpw, real_bytes) > stage_budget) break;
                    if((size_t)n * nJ > (size_t)nreal + nreal / 4 + 2) break;
                    a = na; Jc = nJ; real = nreal;
                    if(bpw > 1 || (a % 2) == 0) { best_a = a; best_J = Jc; }
                }
                if(bpw == 1 && (best_a % 2) != 0)
                {   // a one-hop chunk at an odd hop: take the previous hop along regardless of the budget
                    best_a -= 1; best_J = std::max(best_J, pb.J[best_a]);
                }
                rev.push_back(ProtoChunk{best_a, e - best_a, best_J});
                J_next = best_J;
                e = best_a;
            }
            pb.chunks.assign(rev.rbegin(), rev.rend());
        }
        pb.work = 0; pb.small = true;
        for(size_t c = 0; c < pb.chunks.size(); ++c)
        {
            const uint32_t Jn = c + 1 < pb.chunks.size() ? pb.chunks[c+1].J : 0u;
            pb.work += pb.chunks[c].n * pb.chunks[c].J;
            if(chunk_stage_bytes(pb.chunks[c].n, pb.chunks[c].J, Jn, bpw, real_bytes) > stage_budget) pb.small = false;
        }
        protos.push_back(std::move(pb));
        pos = end;
    }
    // small class first; inside a class heavy bundles first (longest-processing-time order)
    std::vector<uint32_t> border(protos.size());
    std::iota(border.begin(), border.end(), 0u);
    std::stable_sort(border.begin(), border.end(), [&](uint32_t x, uint32_t y) {
        if(protos[x].small != protos[y].small) return protos[x].small;
        return protos[x].work > protos[y].work;
    });

    // ---- lane-class bundles: up to 32 BDDs of the same width ----------------------------
    // A pass over a collection of at most one wave of warps is bound by the scattered per-variable accesses, whose cost is per
    // lane and per SM (tools/microbench/scatter_cost2.cu): the slowest SM is the one with the most BDDs.  With the SM count known
    // (n_sms > 0) and at most 16 bundles per SM the bundles are made a multiple of the SM count and the BDDs dealt evenly over
    // them (bundles of fewer than 32 lanes), so that every SM gets the same number of warps AND of BDDs.
    struct LaneProto { uint32_t first, count, J, n_hops; bool shared; };
    std::vector<LaneProto> lane_protos;
    {
        struct WidthClass { size_t first, count, bundles; };
        std::vector<WidthClass> wc;
        for(size_t pos = 0; pos < lane_order.size();)
        {
            size_t end = pos;
            while(end < lane_order.size() && bdd_maxw[lane_order[end]] == bdd_maxw[lane_order[pos]]) ++end;
            wc.push_back(WidthClass{pos, end - pos, (end - pos + 31) / 32});
            pos = end;
        }
        size_t total = 0;
        for(const WidthClass& c : wc) total += c.bundles;
        if(n_sms > 0 && order.empty() && total > n_sms && total <= 16 * n_sms && total % n_sms != 0)
        {
            size_t extra = (total + n_sms - 1) / n_sms * n_sms - total;
            const size_t n_all = lane_order.size();
            size_t given = 0;
            for(WidthClass& c : wc) { const size_t add = extra * c.count / n_all; c.bundles += add; given += add; }
            for(size_t k = 0; given < extra; ++k, ++given) wc[k % wc.size()].bundles += 1;      // rounding leftovers
            for(WidthClass& c : wc) c.bundles = std::min(c.bundles, c.count);
        }
        for(const WidthClass& c : wc)
        {
            size_t pos = c.first;
            for(size_t i = 0; i < c.bundles; ++i)
            {
                const size_t cnt = c.count / c.bundles + (i < c.count % c.bundles ? 1 : 0);
                uint32_t nh = 0;
                bool sh = false;
                for(size_t q = pos; q < pos + cnt; ++q) { nh = std::max(nh, nlay(lane_order[q])); sh = sh || (!bdd_shared.empty() && bdd_shared[lane_order[q]]); }
                lane_protos.push_back(LaneProto{(uint32_t)pos, (uint32_t)cnt, bdd_maxw[lane_order[pos]], nh, sh});
                pos += cnt;
            }
        }
    }
    // heavy bundles first; in shard mode the bundles with shared variables before all others: they carry the multi-GPU flag barrier of a
    // pass (kernels.cuh, PUSH builds), whose flag should go out early in the pass
    std::stable_sort(lane_protos.begin(), lane_protos.end(), [](const LaneProto& x, const LaneProto& y) {
        if(x.shared != y.shared) return x.shared;
        return x.n_hops * x.J > y.n_hops * y.J;
    });
    for(const LaneProto& lp : lane_protos) L.n_lane_shared_bundles += lp.shared ? 1 : 0;

    timer.lap("sort, bundles");
    // ---- emit: generic bundles own the first slots (topo is indexed by slot there), lane-class
    // bundles follow; in bundle order the lane class comes first -------------------------------
    std::vector<BundleDesc> generic_bundles;
    generic_bundles.reserve(protos.size());
    size_t slot = 0, lay = 0;
    for(const uint32_t pi : border)
    {
        const ProtoBundle& pb = protos[pi];
        BundleDesc bd{};
        bd.cls = CLS_GENERIC;
        bd.hop_base = (uint32_t)L.hops.size();
        bd.n_hops = pb.n_hops; bd.logP = pb.logP; bd.max_J = 0;
        lay = (lay + 1) & ~(size_t)1;
        bd.layer_base = (uint32_t)lay;
        bd.bdd_base = (uint32_t)L.bundle_bdd.size();
        bd.chunk_base = (uint32_t)L.chunks.size();
        bd.n_chunks = (uint32_t)pb.chunks.size();
        const uint32_t bpw = 32u >> pb.logP;
        for(size_t c = 0; c < pb.chunks.size(); ++c)
        {
            const ProtoChunk& pc = pb.chunks[c];
            ChunkRec cr{};
            cr.slot_off = (uint32_t)slot; cr.lay_off = (uint32_t)(lay + (size_t)pc.first * bpw);
            cr.hop_first = pc.first; cr.n_hops = pc.n; cr.J = pc.J;
            cr.J_next = c + 1 < pb.chunks.size() ? pb.chunks[c+1].J : 0u;
            L.chunks.push_back(cr);
            bd.max_J = std::max(bd.max_J, pc.J);
            const size_t sb = chunk_stage_bytes(pc.n, pc.J, cr.J_next, bpw, real_bytes);
            if(pb.small) L.stage_small = std::max(L.stage_small, sb); else L.stage_large = std::max(L.stage_large, sb);
            for(uint32_t k = 0; k < pc.n; ++k)
            {
                if(slot > 0xFFFFFFFFull - 32ull * pc.J) throw layout_error(BDDB200_ERR_INVALID_ARGUMENT, "collection too large (slot index overflow)");
                L.hops.push_back(HopRec{(uint32_t)slot, pc.J});
                slot += 32ull * pc.J;
            }
        }
        for(uint32_t q = 0; q < bpw; ++q)
            L.bundle_bdd.push_back(q < pb.count ? (int32_t)order[pb.first + q] : -1);
        lay += (size_t)pb.n_hops * bpw;
        if(pb.small) { L.n_small_bundles++; L.max_tile_small = std::max(L.max_tile_small, bd.max_J * 32u); }
        else L.max_tile_large = std::max(L.max_tile_large, bd.max_J * 32u);
        L.max_hops = std::max<size_t>(L.max_hops, pb.n_hops);
        generic_bundles.push_back(bd);
    }
    L.n_generic_slots = slot;
    size_t topo_words = slot;
    lay = (lay + 1) & ~(size_t)1;
    for(const LaneProto& lp : lane_protos)
    {
        BundleDesc bd{};
        bd.cls = CLS_LANE;
        bd.hop_base = (uint32_t)L.hops.size();
        bd.n_hops = lp.n_hops; bd.logP = 0; bd.max_J = lp.J;
        bd.layer_base = (uint32_t)lay;
        bd.bdd_base = (uint32_t)L.bundle_bdd.size();
        bd.chunk_base = 0; bd.n_chunks = 0;
        bd.topo_base = (uint32_t)topo_words;
        if(slot + 32ull * lp.J * lp.n_hops > 0xFFFFFFFFull || topo_words + 32ull * lp.n_hops > 0xFFFFFFFFull)
            throw layout_error(BDDB200_ERR_INVALID_ARGUMENT, "collection too large (slot index overflow)");
        LaneDesc ld{};
        ld.slot_off = (uint32_t)slot; ld.lay_off = (uint32_t)lay; ld.topo_off = (uint32_t)topo_words;
        ld.n_hops = lp.n_hops; ld.J = lp.J; ld.bdd_base = bd.bdd_base;
        L.desc_lane.push_back(ld);
        for(uint32_t k = 0; k < lp.n_hops; ++k) { L.hops.push_back(HopRec{(uint32_t)slot, lp.J}); slot += 32ull * lp.J; }
        topo_words += 32ull * lp.n_hops;
        lay += 32ull * lp.n_hops;
        for(uint32_t q = 0; q < 32; ++q)
            L.bundle_bdd.push_back(q < lp.count ? (int32_t)lane_order[lp.first + q] : -1);
        L.lane_max_J = std::max(L.lane_max_J, lp.J);
        L.lane_max_hops = std::max(L.lane_max_hops, lp.n_hops);
        L.max_hops = std::max<size_t>(L.max_hops, lp.n_hops);
        L.bundles.push_back(bd);
    }
    L.n_lane_bundles = L.bundles.size();
    L.bundles.insert(L.bundles.end(), generic_bundles.begin(), generic_bundles.end());
    L.n_topo = topo_words;

    L.desc_fwd.assign(L.bundles.size() * DESC_WORDS, 0u);
    L.desc_bwd.assign(L.bundles.size() * DESC_WORDS, 0u);
    for(size_t g = L.n_lane_bundles; g < L.bundles.size(); ++g)
    {
        const BundleDesc& bd = L.bundles[g];
        for(int dir = 0; dir < 2; ++dir)
        {
            uint32_t* d = (dir == 0 ? L.desc_fwd.data() : L.desc_bwd.data()) + g * DESC_WORDS;
            d[DESC_N_CHUNKS] = bd.n_chunks; d[DESC_LOGP] = bd.logP; d[DESC_BDD_BASE] = bd.bdd_base; d[DESC_MAX_J] = bd.max_J;
            d[DESC_CHUNK_BASE] = bd.chunk_base; d[DESC_N_HOPS] = bd.n_hops; d[DESC_LAYER_BASE] = bd.layer_base;
            for(uint32_t k = 0; k < (uint32_t)DESC_CHUNKS && k < bd.n_chunks; ++k)
            {
                const ChunkRec& c = L.chunks[bd.chunk_base + (dir == 0 ? k : bd.n_chunks - 1 - k)];
                uint32_t* q = d + DESC_FIRST_CHUNK + DESC_CHUNK_WORDS * k;
                q[0] = c.slot_off; q[1] = c.lay_off; q[2] = c.n_hops; q[3] = c.J; q[4] = c.J_next;
            }
        }
    }
    lay = ((lay + 1) & ~(size_t)1) + 2;     // bulk copies of a chunk's layer range may read one entry past it
    L.n_slots = slot; L.n_lay = lay;
    if(lay > 0xFFFFFFF0ull) throw layout_error(BDDB200_ERR_INVALID_ARGUMENT, "collection too large (layer index overflow)");

    timer.lap("emit bundles");
    par_fill(L.topo, L.n_topo, (uint32_t)TOPO_PAD);
#pragma omp parallel for schedule(static)
    for(long long g = 0; g < (long long)L.n_lane_bundles; ++g)
        std::fill(L.topo.begin() + L.bundles[g].topo_base, L.topo.begin() + L.bundles[g].topo_base + 32ull * L.bundles[g].n_hops, 0u);
    par_fill(L.lay_var, L.n_lay, (int32_t)LAY_NONE);
    L.ext2lay.clear(); L.ext2lay.resize(L.n_layers_ext);         // every entry is written below
    L.root_slot.assign(n_bdds, 0);
    L.top_slot.assign(n_bdds, 0);
    L.nr_bdds_per_var.assign(L.n_vars, 0);

#pragma omp parallel for schedule(dynamic, 64)
    for(long long gg = 0; gg < (long long)L.bundles.size(); ++gg)
    {
        const size_t g = (size_t)gg;
        const BundleDesc& bd = L.bundles[g];
        const bool lane_cls = bd.cls == CLS_LANE;
        if(lane_cls)
        {   // hop-major: the 32 words of a hop are written together (the tile rows of this class are 32 consecutive entries; walking
            // BDD by BDD touches a different cache line of every array per layer), the 32 BDDs' instructions stay in L1 meanwhile
            struct Lane { size_t b, bot, top; uint32_t eb, ee; };
            Lane ln[32]; bool has[32];
            for(uint32_t q = 0; q < 32; ++q)
            {
                const int32_t bi = L.bundle_bdd[bd.bdd_base + q];
                has[q] = bi >= 0;
                if(!has[q]) continue;
                const size_t b = (size_t)bi, last = delims[b+1];
                ln[q] = Lane{b, instrs[last-2].index == BOTSINK ? last - 2 : last - 1, instrs[last-2].index == BOTSINK ? last - 1 : last - 2,
                             L.bdd_ext_begin[b], L.bdd_ext_begin[b+1] - 1};
            }
            for(uint32_t k = 0; k < bd.n_hops; ++k)
            {
                const HopRec& hr = L.hops[bd.hop_base + k];
                for(uint32_t q = 0; q < 32; ++q)
                {
                    if(!has[q] || ln[q].eb + k > ln[q].ee) continue;
                    const uint32_t e = ln[q].eb + k, layer_entry = bd.layer_base + k * 32u + q;
                    L.ext2lay[e] = layer_entry;
                    if(e == ln[q].ee)
                    {
                        L.lay_var[layer_entry] = LAY_TOP;
                        L.top_slot[ln[q].b] = hr.node_off + q;
                        continue;
                    }
                    const size_t lb = ext_first_instr[e], le = ext_first_instr[e+1];
                    L.lay_var[layer_entry] = L.ext_var[e];
                    if(k == 0) L.root_slot[ln[q].b] = hr.node_off + q;
                    uint32_t word = 0;
                    for(size_t i = lb; i < le; ++i)
                    {
                        const uint32_t j = (uint32_t)(i - lb);
                        const size_t arcs[2] = {instrs[i].lo, instrs[i].hi};
                        for(uint32_t arc = 0; arc < 2; ++arc)
                        {
                            const size_t c = arcs[arc];
                            if(c == ln[q].bot) continue;
                            const uint32_t r = (c == ln[q].top) ? 0u : (uint32_t)(c - le);
                            word |= lane_arc_bit(bd.max_J, j, arc, r);
                        }
                    }
                    L.topo[bd.topo_base + k * 32u + q] = word;
                }
            }
            continue;
        }
        const uint32_t logP = bd.logP, P = 1u << logP, bpw = 32u >> logP;
        for(uint32_t q = 0; q < bpw; ++q)
        {
            const int32_t bi = L.bundle_bdd[bd.bdd_base + q];
            if(bi < 0) continue;
            const size_t b = (size_t)bi;
            const size_t last = delims[b+1];
            const size_t bot = instrs[last-2].index == BOTSINK ? last - 2 : last - 1, top = instrs[last-2].index == BOTSINK ? last - 1 : last - 2;
            const uint32_t eb = L.bdd_ext_begin[b], ee = L.bdd_ext_begin[b+1] - 1;
            auto tile_slot = [&](uint32_t c) { return (c >> logP) * 32u + (q << logP) + (c & (P - 1)); };
            for(uint32_t e = eb; e <= ee; ++e)
            {
                const uint32_t k = e - eb;
                const uint32_t layer_entry = bd.layer_base + k * bpw + q;
                L.ext2lay[e] = layer_entry;
                const HopRec& hr = L.hops[bd.hop_base + k];
                if(e == ee)
                {
                    if(!lane_cls) L.topo[hr.node_off + tile_slot(0)] = TOPO_TOP;
                    L.lay_var[layer_entry] = LAY_TOP;
                    L.top_slot[b] = hr.node_off + tile_slot(0);
                    continue;
                }
                const size_t lb = ext_first_instr[e], le = ext_first_instr[e+1];
                const size_t next_first = le;   // first instruction of the next layer (or the bot sink)
                L.lay_var[layer_entry] = L.ext_var[e];
                if(k == 0) L.root_slot[b] = hr.node_off + tile_slot(0);
                if(lane_cls)
                {
                    uint32_t word = 0;
                    for(size_t i = lb; i < le; ++i)
                    {
                        const uint32_t j = (uint32_t)(i - lb);
                        const size_t arcs[2] = {instrs[i].lo, instrs[i].hi};
                        for(uint32_t arc = 0; arc < 2; ++arc)
                        {
                            const size_t c = arcs[arc];
                            if(c == bot) continue;
                            const uint32_t r = (c == top) ? 0u : (uint32_t)(c - next_first);
                            word |= lane_arc_bit(bd.max_J, j, arc, r);
                        }
                    }
                    L.topo[bd.topo_base + k * 32u + q] = word;
                    continue;
                }
                for(size_t i = lb; i < le; ++i)
                {
                    auto child = [&](size_t c) -> uint32_t {
                        if(c == bot) return CHILD_BOT;
                        if(c == top) return tile_slot(0);
                        return tile_slot((uint32_t)(c - next_first));
                    };
                    L.topo[hr.node_off + tile_slot((uint32_t)(i - lb))] = child(instrs[i].lo) | (child(instrs[i].hi) << 16);
                }
            }
        }
    }

    timer.lap("topology, layer entries");
    // ---- variable -> layers (sorted by variable, then BDD index) ------------------------
    // External order is BDD-major, so a STABLE sort of the inner layer entries by variable is what is needed.  Two levels: the entries
    // are first dealt into buckets of consecutive variables (every thread owns a range of entries; per-thread bucket counts make the
    // deal stable and race-free, all reads and writes sequential), then every bucket is counting-sorted on its own (its variables'
    // counters and its slice of the output fit the cache).
    L.var_lay_begin.assign(L.n_vars + 1, 0);
    L.sorted_ext.clear(); L.sorted_ext.resize(L.n_layers_ext);    // every entry is written below
    {
        int n_threads = 1;
#ifdef _OPENMP
        n_threads = std::max(1, std::min(omp_get_max_threads(), 64));
#endif
        uint32_t shift = 0;
        while(((L.n_vars + ((size_t)1 << shift) - 1) >> shift) > 2048) ++shift;
        const size_t n_buckets = std::max<size_t>(1, (L.n_vars + ((size_t)1 << shift) - 1) >> shift);
        const size_t n_e = L.n_layers_ext;
        std::vector<size_t> cnt((size_t)n_threads * n_buckets, 0);
#pragma omp parallel for schedule(static, 1) num_threads(n_threads)
        for(int t = 0; t < n_threads; ++t)
        {
            size_t* c = cnt.data() + (size_t)t * n_buckets;
            const size_t e0 = n_e * (size_t)t / n_threads, e1 = n_e * (size_t)(t + 1) / n_threads;
            for(size_t e = e0; e < e1; ++e) if(L.ext_var[e] != INT_MAX) ++c[(size_t)L.ext_var[e] >> shift];
        }
        // start of (bucket, thread) in the dealt arrays: bucket-major, threads in entry order inside a bucket
        std::vector<size_t> bucket_begin(n_buckets + 1, 0);
        size_t run = 0;
        for(size_t k = 0; k < n_buckets; ++k)
        {
            bucket_begin[k] = run;
            for(int t = 0; t < n_threads; ++t) { const size_t c = cnt[(size_t)t * n_buckets + k]; cnt[(size_t)t * n_buckets + k] = run; run += c; }
        }
        bucket_begin[n_buckets] = run;
        const size_t n_inner = run;
        uvec<uint32_t> d_e(n_inner), d_v(n_inner), d_l(n_inner);
#pragma omp parallel for schedule(static, 1) num_threads(n_threads)
        for(int t = 0; t < n_threads; ++t)
        {
            size_t* c = cnt.data() + (size_t)t * n_buckets;
            const size_t e0 = n_e * (size_t)t / n_threads, e1 = n_e * (size_t)(t + 1) / n_threads;
            for(size_t e = e0; e < e1; ++e)
            {
                const int32_t var = L.ext_var[e];
                if(var == INT_MAX) continue;
                const size_t p = c[(size_t)var >> shift]++;
                d_e[p] = (uint32_t)e; d_v[p] = (uint32_t)var; d_l[p] = L.ext2lay[e];
            }
        }
        // in how many BDDs every variable occurs (each bucket owns its variables' counters), then the global offsets
#pragma omp parallel for schedule(dynamic, 8)
        for(long long kk = 0; kk < (long long)n_buckets; ++kk)
            for(size_t p = bucket_begin[kk]; p < bucket_begin[kk + 1]; ++p) L.var_lay_begin[d_v[p] + 1]++;
        for(size_t v = 0; v < L.n_vars; ++v) L.nr_bdds_per_var[v] = (int32_t)L.var_lay_begin[v + 1];
        for(size_t v = 0; v < L.n_vars; ++v) L.var_lay_begin[v+1] += L.var_lay_begin[v];
        L.var_lay.clear(); L.var_lay.resize(n_inner);               // every entry is written below
#pragma omp parallel for schedule(dynamic, 8)
        for(long long kk = 0; kk < (long long)n_buckets; ++kk)
        {
            const size_t v0 = (size_t)kk << shift, v1 = std::min(L.n_vars, ((size_t)kk + 1) << shift);
            std::vector<uint32_t> fill(L.var_lay_begin.begin() + v0, L.var_lay_begin.begin() + v1);
            for(size_t p = bucket_begin[kk]; p < bucket_begin[kk + 1]; ++p)
            {
                const uint32_t q = fill[d_v[p] - v0]++;
                L.var_lay[q] = d_l[p];
                L.sorted_ext[q] = d_e[p];
            }
        }
        // terminal layers last, in BDD order
        size_t term = n_inner;
        for(size_t b = 0; b < n_bdds; ++b) L.sorted_ext[term++] = L.bdd_ext_begin[b + 1] - 1;
    }
    timer.lap("variable -> layers");
    return L;
}

} // namespace bddb200
