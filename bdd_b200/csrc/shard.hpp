// bdd_b200/csrc/shard.hpp -- host-side planning of a constraint-sharded solve: which BDDs a rank owns, which variables
// occur in more than one shard, and the relabelling that puts those first so that only a prefix of the per-variable sum vector
// crosses NVLink.  The reference's only multi-device code is the hybrid CPU + GPU solver, which splits the collection by BDD with
// bdd_collection::remove and keeps global per-variable counts (src/bdd_solver/bdd_multi_parallel_mma_base.cu:121-124, 191-215); this is
// that split for `world` GPU shards.  Same rules as bdd_b200/dist.py (partition_bdds, shared_first_relabeling), which the CPU tests
// compare it with.
#pragma once

#include <algorithm>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <vector>

#include "../../include/bdd_b200.h"

namespace bddb200 {

struct ShardPlan {
    size_t n_vars = 0;
    size_t n_shared = 0;                  // variables that occur in more than one shard: new indices [0, n_shared)
    size_t shared_entries = 0;            // layer entries of THIS shard that belong to shared variables
    size_t first_bdd = 0, n_bdds = 0;     // this rank's contiguous block of BDDs
    std::vector<int32_t> new_of_old;      // relabelling (shared first; both groups keep their relative order)
    std::vector<int32_t> counts_new;      // global nr_bdds_per_var, indexed by the NEW variable index
    std::vector<uint16_t> share_mask;     // per shared variable (new index): the shards that contain it, bit r = rank r (world <= 16)
};

// contiguous blocks of BDDs with balanced node counts (constraint order is kept: a grid-tile ordered instance gets compact shards)
inline std::vector<size_t> shard_bounds(const size_t* delims, size_t n_bdds, int world)
{
    std::vector<size_t> bounds((size_t)world + 1, 0);
    const double total = (double)(delims[n_bdds] - delims[0]);
    for(int r = 1; r < world; ++r)
    {
        const double want = total * r / world;
        // first BDD whose cumulative node count reaches `want`, then one past it
        size_t lo = 0, hi = n_bdds;
        while(lo < hi)
        {
            const size_t mid = (lo + hi) / 2;
            if((double)(delims[mid + 1] - delims[0]) < want) lo = mid + 1; else hi = mid;
        }
        bounds[(size_t)r] = std::min(n_bdds, lo + 1);
    }
    bounds[(size_t)world] = n_bdds;
    for(int r = 1; r <= world; ++r) bounds[(size_t)r] = std::max(bounds[(size_t)r], bounds[(size_t)r - 1]);
    return bounds;
}

inline ShardPlan plan_shard(const bddb200_instruction* instrs, const size_t* delims, size_t n_bdds, size_t n_vars_min, int world, int rank)
{
    if(world < 1 || world > 16 || rank < 0 || rank >= world) throw std::invalid_argument("plan_shard: rank / world (at most 16 ranks)");
    constexpr size_t BOT = (size_t)-2;
    ShardPlan p;
    const std::vector<size_t> bounds = shard_bounds(delims, n_bdds, world);
    p.first_bdd = bounds[(size_t)rank]; p.n_bdds = bounds[(size_t)rank + 1] - bounds[(size_t)rank];
    size_t max_var = 0; bool any = false;
    for(size_t i = delims[0]; i < delims[n_bdds]; ++i)
        if(instrs[i].index < BOT) { max_var = std::max(max_var, instrs[i].index); any = true; }
    p.n_vars = std::max(n_vars_min, any ? max_var + 1 : (size_t)0);
    std::vector<int32_t> lo(p.n_vars, std::numeric_limits<int32_t>::max()), hi(p.n_vars, -1), counts(p.n_vars, 0);
    std::vector<uint16_t> mask(p.n_vars, 0);
    int shard = 0;
    for(size_t b = 0; b < n_bdds; ++b)
    {
        while(b >= bounds[(size_t)shard + 1]) ++shard;
        size_t prev = BOT;
        for(size_t i = delims[b]; i < delims[b + 1]; ++i)
        {
            const size_t v = instrs[i].index;
            if(v >= BOT) continue;
            if(v != prev)
            {   // one layer entry per (variable, BDD): the nodes of a layer are adjacent (quasi-reduced, levelled BDDs)
                lo[v] = std::min(lo[v], (int32_t)shard); hi[v] = std::max(hi[v], (int32_t)shard); ++counts[v];
                mask[v] |= (uint16_t)(1u << shard);
                prev = v;
            }
        }
    }
    p.new_of_old.assign(p.n_vars, 0);
    size_t k = 0;
    for(size_t v = 0; v < p.n_vars; ++v) if(hi[v] > lo[v]) p.new_of_old[v] = (int32_t)k++;
    p.n_shared = k;
    for(size_t v = 0; v < p.n_vars; ++v) if(!(hi[v] > lo[v])) p.new_of_old[v] = (int32_t)k++;
    p.counts_new.assign(p.n_vars, 0);
    for(size_t v = 0; v < p.n_vars; ++v) p.counts_new[(size_t)p.new_of_old[v]] = counts[v];
    p.share_mask.assign(p.n_shared, 0);
    for(size_t v = 0; v < p.n_vars; ++v) if(hi[v] > lo[v]) p.share_mask[(size_t)p.new_of_old[v]] = mask[v];
    for(size_t b = p.first_bdd; b < p.first_bdd + p.n_bdds; ++b)
    {
        size_t prev = BOT;
        for(size_t i = delims[b]; i < delims[b + 1]; ++i)
        {
            const size_t v = instrs[i].index;
            if(v >= BOT || v == prev) continue;
            prev = v;
            if(hi[v] > lo[v]) ++p.shared_entries;
        }
    }
    return p;
}

} // namespace bddb200
