// bdd_b200/csrc/host/split.hpp -- splitting long BDDs into chunks linked by auxiliary variables, in C++ for the host driver.
// The chunk construction of bdd_collection::split_qbdd (src/bdd_collection/bdd_collection.cpp:507-790); same construction as
// bdd_b200/split.py, whose instruction arrays are bit-identical to the reference's (tests/test_split.py against oracle/_ref) and which
// tests/test_cpp_driver.py compares this with.  The collection class around it -- split_qbdd with the optional implication BDD, the
// preprocessor's loop over all BDDs -- is host/bdd_collection.hpp.
//
// Cutting a quasi-reduced BDD in front of a layer of width w introduces w auxiliary 0/1 variables that one-hot encode which node of
// that layer the path goes through (aux variable w-1-k is 1 iff node k is used): the chunk before the cut ends in a *tail* gadget that
// accepts exactly the one-hot pattern of the node each path reached, the chunk after it starts with a *head* gadget that routes the
// pattern to that node.  Auxiliary variables carry no cost.
#pragma once

#include <algorithm>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/bdd_b200.h"

namespace bddb200_host {

namespace split_detail {

constexpr size_t TOPSINK = (size_t)-1, BOTSINK = (size_t)-2;

struct Layers { std::vector<size_t> vars, offs, widths; };
inline Layers layers_of(const bddb200_instruction* ins, size_t first, size_t last)
{
    Layers l;
    for(size_t i = first; i + 2 < last; ++i)
        if(i == first || ins[i].index != ins[i - 1].index) { l.vars.push_back(ins[i].index); l.offs.push_back(i); }
    for(size_t k = 0; k < l.offs.size(); ++k) l.widths.push_back((k + 1 < l.offs.size() ? l.offs[k + 1] : last - 2) - l.offs[k]);
    return l;
}

// chunks of the QBDD ins[first, last) appended to `out` at absolute position `base` (= out.size() of the target array);
// returns the number of chunks (0: the BDD has at most chunk_size variables and stays as it is) and advances aux
inline size_t split_qbdd(const bddb200_instruction* ins, size_t first, size_t last, size_t chunk_size, size_t& aux, size_t base,
                         std::vector<bddb200_instruction>& out, std::vector<size_t>& out_delims)
{
    const Layers L = layers_of(ins, first, last);
    const size_t n_layers = L.vars.size();
    if(n_layers <= chunk_size) return 0;
    const size_t nr_chunks = (n_layers + chunk_size - 1) / chunk_size;
    for(size_t c = 0; c + 1 < nr_chunks; ++c)
        if(L.widths[(c + 1) * chunk_size] <= 1)
            throw std::invalid_argument("split_qbdd: cannot cut in front of a layer of width 1 (bdd_collection.cpp:598)");
    std::vector<size_t> aux_vars{aux};
    for(size_t c = 1; c + 1 < nr_chunks; ++c) aux_vars.push_back(aux_vars.back() + L.widths[c * chunk_size]);
    auto layer_offset = [&](size_t layer) { return layer < n_layers ? L.offs[layer] : last - 2; };

    for(size_t c = 0; c < nr_chunks; ++c)
    {
        const size_t first_layer = c * chunk_size, last_layer = std::min((c + 1) * chunk_size - 1, n_layers - 1);
        const size_t w_head = c > 0 ? L.widths[first_layer] : 0, n_head = w_head * (w_head + 1) / 2;
        const size_t b0 = layer_offset(first_layer), b1 = layer_offset(last_layer + 1), n_body = b1 - b0;
        const bool has_tail = c + 1 < nr_chunks;
        const size_t w_tail = has_tail ? L.widths[last_layer + 1] : 0;
        const size_t n_tail = has_tail ? w_tail * (w_tail + 1) / 2 + w_tail - 1 : 0;
        const size_t bottom = base + n_head + n_body + n_tail, top = bottom + 1;
        auto H = [&](size_t i, size_t j) { return base + i * (i + 1) / 2 + j; };
        auto T = [&](size_t i, size_t j) {
            if(i == 0) return base + n_head + n_body + j;
            return base + n_head + n_body + w_tail + w_tail * (i - 1) + j - (i - 1) * (i - 2) / 2;      // (i-1)(i-2)/2 is 0 for i = 1, 2
        };
        const size_t before = out.size();
        // 1) head gadget
        if(c > 0)
        {
            const size_t av = aux_vars[c - 1];
            for(size_t i = 0; i + 1 < w_head; ++i)
                for(size_t j = 0; j <= i; ++j)
                    out.push_back(j == 0 ? bddb200_instruction{H(i + 1, 0), H(i + 1, 1), av + i} : bddb200_instruction{H(i + 1, j + 1), bottom, av + i});
            const size_t i = w_head - 1;
            for(size_t j = 0; j <= i; ++j)
                out.push_back(j == 0 ? bddb200_instruction{bottom, base + n_head + j, av + i} : bddb200_instruction{base + n_head + j, bottom, av + i});
        }
        // 2) the chunk's own nodes keep their relative positions; arcs into the layer after the chunk land in the tail gadget
        for(size_t k = b0; k < b1; ++k)
        {
            const size_t pos = base + n_head + (k - b0);
            auto child = [&](size_t ch) -> size_t {
                if(ins[ch].index == BOTSINK) return bottom;
                if(ins[ch].index == TOPSINK) return top;
                return pos + ch - k;
            };
            out.push_back(bddb200_instruction{child(ins[k].lo), child(ins[k].hi), ins[k].index});
        }
        // 3) tail gadget
        if(has_tail)
        {
            const size_t av = aux_vars[c];
            for(size_t j = 0; j < w_tail; ++j)
                out.push_back(j + 1 == w_tail ? bddb200_instruction{bottom, T(1, w_tail - 1), av} : bddb200_instruction{T(1, j), bottom, av});
            for(size_t i = 1; i + 1 < w_tail; ++i)
                for(size_t j = 0; j < w_tail - i + 1; ++j)
                {
                    if(j + 1 == w_tail - i + 1) out.push_back(bddb200_instruction{T(i + 1, w_tail - i - 1), bottom, av + i});
                    else if(j + 1 == w_tail - i) out.push_back(bddb200_instruction{bottom, T(i + 1, j), av + i});
                    else out.push_back(bddb200_instruction{T(i + 1, j), bottom, av + i});
                }
            out.push_back(bddb200_instruction{bottom, top, av + w_tail - 1});
            out.push_back(bddb200_instruction{top, bottom, av + w_tail - 1});
        }
        if(out.size() - before != n_head + n_body + n_tail) throw std::logic_error("split_qbdd: gadget size mismatch");
        out.push_back(bddb200_instruction{BOTSINK, BOTSINK, BOTSINK});       // bot sink first, then top sink (:768-771)
        out.push_back(bddb200_instruction{TOPSINK, TOPSINK, TOPSINK});
        base = out.size();
        out_delims.push_back(out.size());
    }
    aux = aux_vars.back() + L.widths[(nr_chunks - 1) * chunk_size];
    return nr_chunks;
}

} // namespace split_detail

// Split length when the configuration gives none: the largest length that yields at least n_sms * warps_per_sm bundles of 32 BDDs, never
// below min_length; SIZE_MAX when nothing should be split (bdd_b200/split.py: compute_split_length; the reference's rule,
// bdd_preprocessor.cpp:32-121, targets the occupancy of its hop-synchronous kernels and does not transfer).
template<typename COLLECTION>
inline size_t compute_split_length(const COLLECTION& col, size_t n_sms = 148, size_t warps_per_sm = 16, size_t min_length = 16)
{
    using namespace split_detail;
    const size_t nb = col.delims.size() - 1;
    std::vector<size_t> layers(nb, 0);
    size_t max_layers = 0;
    for(size_t b = 0; b < nb; ++b)
    {
        for(size_t i = col.delims[b]; i + 2 < col.delims[b + 1]; ++i)
            if(i == col.delims[b] || col.instrs[i].index != col.instrs[i - 1].index) ++layers[b];
        max_layers = std::max(max_layers, layers[b]);
    }
    const size_t target = 32 * n_sms * warps_per_sm;
    if(nb >= target || max_layers <= min_length) return std::numeric_limits<size_t>::max();
    auto chunks_for = [&](size_t len) { size_t n = 0; for(const size_t l : layers) n += (l + len - 1) / len; return n; };
    size_t lo = min_length, hi = max_layers;
    if(chunks_for(lo) < target) return lo;
    while(lo < hi)
    {
        const size_t mid = (lo + hi + 1) / 2;
        if(chunks_for(mid) >= target) lo = mid; else hi = mid - 1;
    }
    return lo;
}

} // namespace bddb200_host
