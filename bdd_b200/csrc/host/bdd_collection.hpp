// bdd_b200/csrc/host/bdd_collection.hpp -- host-side BDD collection: direct generators for the common constraint shapes, the structural
// operations the solver's front end needs, and the long-BDD splitter with its optional implication BDD.
//
// What it stands for in the reference: BDD::bdd_collection (include/bdd_collection/bdd_collection.h, src/bdd_collection/bdd_collection.cpp)
//   generators     simplex_constraint :2039, not_all_false_constraint :2105, all_equal_constraint :2136, cardinality_constraint :2187
//   relabelling    rebase (header :311-352), negate :2023, invert :2029
//   structure      variables :1201, layer_widths :1344, layer_offsets :1367, is_reordered :1510, reorder :1429, make_qbdd :1670,
//                  remove (header :371-419), is_qbdd :500 / contiguous_vars :1614
//   conjunction    bdd_and :31-315 (two BDDs, N BDDs, iterator range)
//   splitting      split_qbdd :507-949 incl. the implication BDD over the auxiliary variables (:805-940)
// Every method returns instruction arrays identical to the reference's (tests/test_collection.py against oracle/_ref), but none of them
// is built the reference's way:
//   * the counting constraints are one state machine (layer, count so far) -> position, not four hand-unrolled emitters;
//   * bdd_and folds its operands pairwise in a small hash-consed ROBDD store and then writes the canonical result in the reference's
//     node order.  That order is the reverse of a lo-first post-order walk of the RESULT graph -- which is what the reference's
//     recursion over tuples of operand nodes produces whatever the number and grouping of operands -- so one emitter serves all arities
//     (the reference instantiates a 2-ary and 47 N-ary templates and batches above 49 operands);
//   * reorder is a stable counting sort by variable rank; make_qbdd lays the pass-through chains out directly;
//   * the implication BDD needs reachability only between the layers that were cut: bit sets pushed down the layers from each cut
//     instead of two transitive closures of the whole DAG.
// Variables of a BDD are ordered as the reference orders them (ascending when every arc goes to a larger variable, else Kahn's
// topological order with a FIFO queue over the sorted variable arcs, :1232-1304).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <deque>
#include <limits>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../../include/bdd_b200.h"
#include "split.hpp"

namespace bddb200_host {

namespace collection_detail {

struct pair_hash {
    size_t operator()(const std::array<size_t, 2>& k) const { return std::hash<size_t>()(k[0] * 0x9e3779b97f4a7c15ull ^ (k[1] + 0x7f4a7c15ull + (k[0] << 6))); }
};
struct triple_hash {
    size_t operator()(const std::array<size_t, 3>& k) const { return pair_hash()({pair_hash()({k[0], k[1]}), k[2]}); }
};

// Reduced ordered BDDs over ascending variable indices with shared nodes; 0 is false, 1 is true.
class robdd_store {
public:
    struct node { size_t var, lo, hi; };
    robdd_store() : nodes_{{(size_t)-2, 0, 0}, {(size_t)-1, 1, 1}} {}
    const node& operator[](size_t f) const { return nodes_[f]; }
    size_t make(size_t var, size_t lo, size_t hi)
    {
        if(lo == hi) return lo;
        const auto [it, fresh] = unique_.try_emplace({var, lo, hi}, nodes_.size());
        if(fresh) nodes_.push_back({var, lo, hi});
        return it->second;
    }
    size_t conjunction(size_t f, size_t g)
    {
        if(f == 0 || g == 0) return 0;
        if(f == 1 || f == g) return g;
        if(g == 1) return f;
        if(f > g) std::swap(f, g);
        const auto it = and_cache_.find({f, g});
        if(it != and_cache_.end()) return it->second;
        const node a = nodes_[f], b = nodes_[g];
        const size_t v = std::min(a.var, b.var);
        const size_t lo = conjunction(a.var == v ? a.lo : f, b.var == v ? b.lo : g);
        const size_t hi = conjunction(a.var == v ? a.hi : f, b.var == v ? b.hi : g);
        const size_t r = make(v, lo, hi);
        and_cache_.emplace(std::array<size_t, 2>{f, g}, r);
        return r;
    }
private:
    std::vector<node> nodes_;
    std::unordered_map<std::array<size_t, 3>, size_t, triple_hash> unique_;
    std::unordered_map<std::array<size_t, 2>, size_t, pair_hash> and_cache_;
};

} // namespace collection_detail

// ---------------------------------------------------------------------------------------------------------------- QBDD builder
// The quasi-reduced BDD of one linear constraint  sum_k a_k x_k  {<=, >=, =}  rhs  over 0/1 variables, built directly: states are the
// reachable partial sums, equal sub-functions are merged bottom-up, layers the function does not depend on are spliced out.  It is the
// canonical form the reference reaches through lineq_bdd -> bdd_mgr -> add_bdd -> reorder -> make_qbdd (bdd_preprocessor.cpp:175-228)
// up to the order of the nodes inside a layer (tests/test_host.py, tests/test_collection.py against the reference's converter).
enum Ineq { LE = 0, GE = 1, EQ = 2 };
constexpr size_t MAX_PARTIAL_SUMS = size_t(1) << 20;     // the builder enumerates the reachable partial sums of a layer; a solver layer holds at most 65 504 nodes anyway

// One BDD with local numbering: node i branches on position layer[i] of the constraint's variable list; children are local node ids,
// -1 = bot sink, -2 = top sink; nodes are ordered layer by layer.
struct QbddTemplate {
    bool trivial = false;              // the constraint is always satisfied: no BDD
    std::vector<long long> layer, lo, hi;
};

inline QbddTemplate qbdd_template(const std::vector<long long>& a, int ineq, long long rhs)
{
    const size_t n = a.size();
    if(n == 0) throw std::runtime_error("empty constraint");
    constexpr long long BOT = -1, TOP = -2;
    std::vector<std::vector<long long>> sums(n + 1);
    sums[0] = {0};
    for(size_t k = 0; k < n; ++k)
    {
        std::set<long long> nxt;
        for(long long s : sums[k]) { nxt.insert(s); nxt.insert(s + a[k]); }
        if(nxt.size() > MAX_PARTIAL_SUMS)
            throw std::runtime_error("constraint with more than " + std::to_string(MAX_PARTIAL_SUMS) + " distinct partial sums in a layer: outside the direct BDD "
                                     "builder's scope (the reference converts such knapsack rows through its BDD manager, test/hard_ineqs.h)");
        sums[k + 1].assign(nxt.begin(), nxt.end());
    }
    auto accept = [&](long long s) { return ineq == LE ? s <= rhs : (ineq == GE ? s >= rhs : s == rhs); };
    std::vector<std::map<long long, long long>> ident(n + 1);         // sub-function id of every state, bottom-up
    for(long long s : sums[n]) ident[n][s] = accept(s) ? TOP : BOT;
    std::vector<std::vector<std::pair<long long, long long>>> nodes(n);
    for(size_t kk = n; kk-- > 0;)
    {
        std::map<std::pair<long long, long long>, long long> table;
        for(long long s : sums[kk])
        {
            const std::pair<long long, long long> key(ident[kk + 1][s], ident[kk + 1][s + a[kk]]);
            if(key.first == BOT && key.second == BOT) { ident[kk][s] = BOT; continue; }
            auto it = table.find(key);
            if(it == table.end()) { it = table.emplace(key, (long long)nodes[kk].size()).first; nodes[kk].push_back(key); }
            ident[kk][s] = it->second;
        }
    }
    if(ident[0][0] == BOT) throw std::runtime_error("problem is infeasible");
    QbddTemplate t;
    bool any_bot = false;
    for(const auto& nl : nodes) for(const auto& key : nl) any_bot = any_bot || key.first == BOT || key.second == BOT;
    if(!any_bot) { t.trivial = true; return t; }
    // a variable the function does not depend on (every node of its layer has lo == hi) gets no layer: splice such layers out
    std::vector<char> keep(n);
    for(size_t k = 0; k < n; ++k) { keep[k] = 0; for(const auto& key : nodes[k]) if(key.first != key.second) keep[k] = 1; }
    std::vector<long long> offset(n, 0);
    long long off = 0;
    for(size_t k = 0; k < n; ++k) if(keep[k]) { offset[k] = off; off += (long long)nodes[k].size(); }
    auto stands_for = [&](size_t k, long long i) -> long long {      // global id of the first kept node reached, or a terminal code
        while(true)
        {
            if(i < 0) return i;
            if(keep[k]) return offset[k] + i;
            i = nodes[k][(size_t)i].first;
            ++k;
        }
    };
    for(size_t k = 0; k < n; ++k)
    {
        if(!keep[k]) continue;
        for(const auto& key : nodes[k])
        {
            t.layer.push_back((long long)k);
            t.lo.push_back(key.first >= 0 ? stands_for(k + 1, key.first) : key.first);
            t.hi.push_back(key.second >= 0 ? stands_for(k + 1, key.second) : key.second);
        }
    }
    return t;
}

class bdd_collection {
public:
    static constexpr size_t TOPSINK = (size_t)-1, BOTSINK = (size_t)-2;        // bdd_instruction::topsink_index / botsink_index
    std::vector<bddb200_instruction> instrs;
    std::vector<size_t> delims{0};

    bdd_collection() = default;
    bdd_collection(const bddb200_instruction* ins, size_t n_instr, const size_t* d, size_t n_bdds)
        : instrs(ins, ins + n_instr), delims(d, d + n_bdds + 1)
    {
        if(delims.front() != 0 || delims.back() != n_instr || !std::is_sorted(delims.begin(), delims.end())) throw std::invalid_argument("bdd_collection: delimiters do not partition the instruction array");
        for(size_t b = 0; b < n_bdds; ++b) check_structure(b);
    }

    size_t nr_bdds() const { return delims.size() - 1; }
    size_t nr_bdd_nodes(size_t b) const { need(b); return delims[b + 1] - delims[b]; }               // the two sinks count, as in the reference (:430)
    size_t nr_bdd_nodes() const { return instrs.size(); }
    size_t offset(size_t b) const { need(b); return delims[b]; }
    bool is_terminal(size_t i) const { return instrs[i].index >= BOTSINK; }

    // ---------------------------------------------------------------------------------------- generators
    // All three counting constraints walk the same state space: node (i, s) = "s of x_0 .. x_{i-1} are set".
    // exactly one of n variables (:2039-2103): layer 0 has one node, every other layer the states 0 and 1
    size_t simplex_constraint(size_t n)
    {
        if(n == 0) throw std::invalid_argument("simplex_constraint: n must be positive");
        const size_t off = instrs.size(), bot = off + 2 * n - 1, top = bot + 1;
        auto pos = [&](size_t i, size_t s) { return i == 0 ? off : off + 2 * i - 1 + s; };
        auto after = [&](size_t i, size_t s) { return s > 1 ? bot : (i + 1 < n ? pos(i + 1, s) : (s == 1 ? top : bot)); };
        for(size_t i = 0; i < n; ++i)
            for(size_t s = 0; s < (i == 0 ? 1u : 2u); ++s) instrs.push_back({after(i, s), after(i, s + 1), i});
        return close(bot, top);
    }
    // at least one of n variables (:2105-2134): a reduced chain, NOT quasi-reduced (hi arcs jump to the top sink)
    size_t not_all_false_constraint(size_t n)
    {
        if(n == 0) throw std::invalid_argument("not_all_false_constraint: n must be positive");
        const size_t off = instrs.size(), bot = off + n, top = bot + 1;
        for(size_t i = 0; i < n; ++i) instrs.push_back({i + 1 < n ? off + i + 1 : bot, top, i});
        return close(bot, top);
    }
    // x_0 = x_1 = ... = x_{n-1} (:2136-2185): state = the common value; reduced, not quasi-reduced
    size_t all_equal_constraint(size_t n)
    {
        if(n < 2) throw std::invalid_argument("all_equal_constraint: n must be at least 2");
        const size_t off = instrs.size(), bot = off + 2 * n - 1, top = bot + 1;
        auto next = [&](size_t i, size_t s) { return i + 1 < n ? off + 2 * (i + 1) - 1 + s : top; };
        instrs.push_back({next(0, 0), next(0, 1), 0});
        for(size_t i = 1; i < n; ++i)
        {
            instrs.push_back({next(i, 0), bot, i});
            instrs.push_back({bot, next(i, 1), i});
        }
        return close(bot, top);
    }
    // exactly k of n variables (:2187-2263): layer i holds the counts that can still reach k; top sink before bot sink
    size_t cardinality_constraint(size_t n, size_t k)
    {
        if(n < 2 || k > n) throw std::invalid_argument("cardinality_constraint: need n > 1 and k <= n");
        if(k == 0) { const size_t b = not_all_false_constraint(n); negate(b); return b; }
        if(k == 1) return simplex_constraint(n);
        auto first_count = [&](size_t i) { return k - std::min(k, n - i); };
        auto last_count = [&](size_t i) { return std::min(k, i); };
        std::vector<size_t> layer_start(n + 1, instrs.size());
        for(size_t i = 0; i < n; ++i) layer_start[i + 1] = layer_start[i] + last_count(i) - first_count(i) + 1;
        const size_t top = layer_start[n], bot = top + 1;
        auto at = [&](size_t i, size_t s) {
            if(s > k || s + (n - i) < k) return bot;
            if(i == n) return s == k ? top : bot;
            return layer_start[i] + s - first_count(i);
        };
        for(size_t i = 0; i < n; ++i)
            for(size_t s = first_count(i); s <= last_count(i); ++s) instrs.push_back({at(i + 1, s), at(i + 1, s + 1), i});
        instrs.push_back(topsink());
        instrs.push_back(botsink());
        delims.push_back(instrs.size());
        return nr_bdds() - 1;
    }

    // the BDD of a template over the given variables (position k of the constraint -> variables[k]); SIZE_MAX when the constraint is
    // always satisfied and no BDD is added (bdd_preprocessor.cpp:183-184)
    size_t add_bdd(const QbddTemplate& t, const std::vector<size_t>& variables)
    {
        if(t.trivial) return std::numeric_limits<size_t>::max();
        const size_t first = instrs.size(), nn = t.layer.size();
        auto child = [&](long long x) -> size_t { return x == -1 ? first + nn : (x == -2 ? first + nn + 1 : first + (size_t)x); };
        for(size_t i = 0; i < nn; ++i) instrs.push_back({child(t.lo[i]), child(t.hi[i]), variables.at((size_t)t.layer[i])});
        return close(first + nn, first + nn + 1);
    }
    // one linear constraint over 0/1 variables (ascending variable indices, as the reference's converter expects them)
    size_t add_linear_constraint(const std::vector<long long>& coefficients, const std::vector<size_t>& variables, int ineq, long long rhs)
    {
        if(coefficients.size() != variables.size()) throw std::invalid_argument("add_linear_constraint: one coefficient per variable");
        if(ineq != LE && ineq != GE && ineq != EQ) throw std::invalid_argument("add_linear_constraint: relation must be 0 (<=), 1 (>=) or 2 (=)");
        for(size_t k = 0; k + 1 < variables.size(); ++k)
            if(!(variables[k] < variables[k + 1])) throw std::invalid_argument("add_linear_constraint: variables must ascend");
        return add_bdd(qbdd_template(coefficients, ineq, rhs), variables);
    }

    // ---------------------------------------------------------------------------------------- relabelling
    // variable i of the BDD becomes vars[i] (header :311-333)
    template<typename ITERATOR>
    void rebase(size_t b, ITERATOR vars_begin, ITERATOR vars_end)
    {
        need(b);
        const size_t n = std::distance(vars_begin, vars_end);
        for(size_t i = delims[b]; i + 2 < delims[b + 1]; ++i)
        {
            if(instrs[i].index >= n) throw std::invalid_argument("rebase: variable " + std::to_string(instrs[i].index) + " has no image");
            instrs[i].index = *(vars_begin + instrs[i].index);
        }
    }
    void rebase(size_t b, const std::unordered_map<size_t, size_t>& var_map)
    {
        need(b);
        for(size_t i = delims[b]; i + 2 < delims[b + 1]; ++i)
        {
            const auto it = var_map.find(instrs[i].index);
            if(it == var_map.end()) throw std::invalid_argument("rebase: variable " + std::to_string(instrs[i].index) + " has no image");
            instrs[i].index = it->second;
        }
    }
    void negate(size_t b) { need(b); std::swap(instrs[delims[b + 1] - 1], instrs[delims[b + 1] - 2]); }
    void invert(size_t b, size_t var)
    {
        need(b);
        for(size_t i = delims[b]; i + 2 < delims[b + 1]; ++i)
            if(instrs[i].index == var) std::swap(instrs[i].lo, instrs[i].hi);
    }

    // ---------------------------------------------------------------------------------------- structure
    // the variables of a BDD in the order its layers follow one another (:1201-1306)
    std::vector<size_t> variables(size_t b) const
    {
        need(b);
        bool ascending = true;
        std::vector<size_t> vars;
        for(size_t i = delims[b]; i + 2 < delims[b + 1]; ++i)
        {
            const bddb200_instruction& x = instrs[i];
            for(const size_t child : {x.lo, x.hi})
                if(!is_terminal(child) && x.index > instrs[child].index) ascending = false;
            if(vars.empty() || vars.back() != x.index) vars.push_back(x.index);
        }
        std::sort(vars.begin(), vars.end());
        vars.erase(std::unique(vars.begin(), vars.end()), vars.end());
        if(ascending) return vars;
        std::vector<std::array<size_t, 2>> arcs;
        for(size_t i = delims[b]; i + 2 < delims[b + 1]; ++i)
            for(const size_t child : {instrs[i].lo, instrs[i].hi})
                if(!is_terminal(child)) arcs.push_back({instrs[i].index, instrs[child].index});
        std::sort(arcs.begin(), arcs.end());
        arcs.erase(std::unique(arcs.begin(), arcs.end()), arcs.end());
        std::unordered_map<size_t, size_t> first_arc, pending;
        for(size_t a = arcs.size(); a-- > 0;) { first_arc[arcs[a][0]] = a; ++pending[arcs[a][1]]; }
        std::deque<size_t> ready{instrs[delims[b]].index};
        std::vector<size_t> order;
        while(!ready.empty())
        {
            const size_t v = ready.front();
            ready.pop_front();
            order.push_back(v);
            const auto it = first_arc.find(v);
            if(it == first_arc.end()) continue;
            for(size_t a = it->second; a < arcs.size() && arcs[a][0] == v; ++a)
                if(--pending[arcs[a][1]] == 0) ready.push_back(arcs[a][1]);
        }
        if(order.size() != vars.size()) throw std::invalid_argument("variables: the BDD has no consistent variable order");
        return order;
    }
    size_t nr_variables(size_t b) const { return variables(b).size(); }
    std::array<size_t, 2> min_max_variables(size_t b) const
    {
        need(b);
        std::array<size_t, 2> mm{std::numeric_limits<size_t>::max(), 0};
        for(size_t i = delims[b]; i + 2 < delims[b + 1]; ++i) { mm[0] = std::min(mm[0], instrs[i].index); mm[1] = std::max(mm[1], instrs[i].index); }
        return mm;
    }
    std::vector<size_t> layer_offsets(size_t b) const                        // absolute index of the first node of every layer (:1367)
    {
        need(b);
        std::vector<size_t> offs;
        for(size_t i = delims[b]; i + 2 < delims[b + 1]; ++i)
            if(i == delims[b] || instrs[i].index != instrs[i - 1].index) offs.push_back(i);
        return offs;
    }
    std::vector<size_t> layer_widths(size_t b) const                         // (:1344)
    {
        const std::vector<size_t> offs = layer_offsets(b);
        std::vector<size_t> w(offs.size());
        for(size_t l = 0; l < offs.size(); ++l) w[l] = (l + 1 < offs.size() ? offs[l + 1] : delims[b + 1] - 2) - offs[l];
        return w;
    }
    // nodes grouped by variable, groups in variable order (:1510-1530)
    bool is_reordered(size_t b) const
    {
        const std::vector<size_t> r = node_ranks(b, variables(b));
        for(size_t i = 0; i + 1 < r.size(); ++i)
            if(r[i + 1] != r[i] && r[i + 1] != r[i] + 1) return false;
        return true;
    }
    // every arc ends in the next layer or in the bot sink (:500-505, :1614-1640)
    bool is_qbdd(size_t b) const
    {
        need(b);
        const std::vector<size_t> vars = variables(b), r = node_ranks(b, vars);
        const size_t first = delims[b], n = r.size();
        for(size_t i = 0; i < n; ++i)
            for(const size_t child : {instrs[first + i].lo, instrs[first + i].hi})
            {
                if(child <= first + i || child >= delims[b + 1]) return false;
                if(instrs[child].index == BOTSINK) continue;
                if(r[i] + 1 == vars.size() ? instrs[child].index != TOPSINK : (child - first >= n || r[child - first] != r[i] + 1)) return false;
            }
        return true;
    }
    bool evaluate(size_t b, const std::vector<char>& labeling) const
    {
        need(b);
        size_t i = delims[b];
        while(!is_terminal(i)) i = labeling.at(instrs[i].index) ? instrs[i].hi : instrs[i].lo;
        return instrs[i].index == TOPSINK;
    }

    // stable sort of the nodes by the rank of their variable (:1429-1508); nothing moves when the BDD is layered already
    void reorder(size_t b)
    {
        const std::vector<size_t> vars = variables(b), r = node_ranks(b, vars);
        bool layered = true;
        for(size_t i = 0; i + 1 < r.size() && layered; ++i) layered = r[i + 1] == r[i] || r[i + 1] == r[i] + 1;
        if(layered) return;
        const size_t first = delims[b], n = r.size();
        std::vector<size_t> start(vars.size() + 1, 0), where(n + 2);
        for(size_t i = 0; i < n; ++i) ++start[r[i] + 1];
        for(size_t l = 0; l < vars.size(); ++l) start[l + 1] += start[l];
        for(size_t i = 0; i < n; ++i) where[i] = start[r[i]]++;
        where[n] = n; where[n + 1] = n + 1;
        std::vector<bddb200_instruction> sorted(n);
        for(size_t i = 0; i < n; ++i)
        {
            const bddb200_instruction& x = instrs[first + i];
            sorted[where[i]] = {first + where[x.lo - first], first + where[x.hi - first], x.index};
        }
        std::copy(sorted.begin(), sorted.end(), instrs.begin() + first);
    }

    // Appends the quasi-reduced form of BDD b (:1670-1810): an arc that skips layers goes through one pass-through node (lo = hi) per
    // skipped layer, shared between all arcs into the same target; arcs into the bot sink stay direct.  Layer by layer the nodes are
    // the original ones in their order followed by the pass-through nodes in the order the arcs first ask for them.
    size_t make_qbdd(size_t b)
    {
        need(b);
        const std::vector<size_t> vars = variables(b), rank = node_ranks(b, vars);
        const size_t first = delims[b], n = delims[b + 1] - 2 - first, depth = vars.size();
        constexpr size_t TOP = (size_t)-1, BOT = (size_t)-2, NONE = (size_t)-3;
        struct raw { size_t lo, hi, layer; };
        std::vector<raw> nodes(n);
        auto local = [&](size_t child) { return instrs[child].index == TOPSINK ? TOP : (instrs[child].index == BOTSINK ? BOT : child - first); };
        for(size_t i = 0; i < n; ++i) nodes[i] = {local(instrs[first + i].lo), local(instrs[first + i].hi), rank[i]};
        std::unordered_map<std::array<size_t, 2>, size_t, collection_detail::pair_hash> chain;          // (layer, target) -> pass-through node
        auto target_layer = [&](size_t t) { return t == TOP ? depth : nodes[t].layer; };
        auto route = [&](size_t from_layer, size_t t) -> size_t {            // where an arc from from_layer towards t has to land
            if(t == BOT) return BOT;
            const size_t end = target_layer(t);
            size_t landing = NONE, prev = NONE;
            for(size_t l = from_layer + 1; l < end; ++l)
            {
                const auto [it, fresh] = chain.try_emplace({l, t}, nodes.size());
                if(prev != NONE) nodes[prev].lo = nodes[prev].hi = it->second;
                if(landing == NONE) landing = it->second;
                if(!fresh) return landing;
                nodes.push_back({t, t, l});
                prev = it->second;
            }
            return landing == NONE ? t : landing;
        };
        for(size_t i = 0; i < n; ++i)
        {
            const size_t lo = route(nodes[i].layer, nodes[i].lo);
            nodes[i].lo = lo;
            const size_t hi = route(nodes[i].layer, nodes[i].hi);
            nodes[i].hi = hi;
        }
        // stable counting sort by layer, then the sinks: top first (:1774-1776)
        std::vector<size_t> start(depth + 1, 0), where(nodes.size());
        for(const raw& x : nodes) ++start[x.layer + 1];
        for(size_t l = 0; l < depth; ++l) start[l + 1] += start[l];
        for(size_t i = 0; i < nodes.size(); ++i) where[i] = start[nodes[i].layer]++;
        const size_t off = instrs.size(), top = off + nodes.size(), bot = top + 1;
        auto global = [&](size_t t) { return t == TOP ? top : (t == BOT ? bot : off + where[t]); };
        instrs.resize(off + nodes.size());
        for(size_t i = 0; i < nodes.size(); ++i) instrs[off + where[i]] = {global(nodes[i].lo), global(nodes[i].hi), vars[nodes[i].layer]};
        instrs.push_back(topsink());
        instrs.push_back(botsink());
        delims.push_back(instrs.size());
        return nr_bdds() - 1;
    }

    // removes the BDDs with the given (ascending) numbers and closes the gaps (header :371-419)
    template<typename ITERATOR>
    void remove(ITERATOR nrs_begin, ITERATOR nrs_end)
    {
        if(nrs_begin == nrs_end) return;
        for(ITERATOR it = nrs_begin; it != nrs_end; ++it)
            if(*it >= nr_bdds() || (it != nrs_begin && !(*(it - 1) < *it))) throw std::invalid_argument("remove: BDD numbers must be ascending and valid");
        size_t to = delims[*nrs_begin], kept = *nrs_begin;
        ITERATOR gone = nrs_begin;
        for(size_t b = *nrs_begin; b < nr_bdds(); ++b)
        {
            if(gone != nrs_end && *gone == b) { ++gone; continue; }
            const size_t shift = delims[b] - to;
            for(size_t i = delims[b]; i < delims[b + 1]; ++i, ++to)
            {
                instrs[to] = instrs[i];
                if(instrs[to].index < BOTSINK) { instrs[to].lo -= shift; instrs[to].hi -= shift; }
            }
            delims[++kept] = to;
        }
        delims.resize(kept + 1);
        instrs.resize(to);
    }
    void remove(size_t b) { const std::array<size_t, 1> one{b}; remove(one.begin(), one.end()); }

    // ---------------------------------------------------------------------------------------- conjunction
    // Appends the reduced BDD of the conjunction of the given BDDs (all over ascending variables) and returns its number (:31-315 and
    // header :494-600).  Node order: reverse lo-first post-order of the result, then top sink, bot sink.
    template<typename ITERATOR>
    size_t bdd_and(ITERATOR nrs_begin, ITERATOR nrs_end)
    {
        if(std::distance(nrs_begin, nrs_end) < 2) throw std::invalid_argument("bdd_and: needs at least two BDDs");
        collection_detail::robdd_store store;
        size_t f = 1;
        for(ITERATOR it = nrs_begin; it != nrs_end; ++it) f = store.conjunction(f, load(store, *it));
        if(f < 2) throw std::invalid_argument(f == 0 ? "bdd_and: the conjunction is infeasible" : "bdd_and: the conjunction is always true");
        // post-order numbers ("stack positions" 2, 3, ...; the sinks hold 0 and 1), children before parents, lo before hi
        std::unordered_map<size_t, size_t> number{{0, 0}, {1, 1}};
        std::vector<size_t> order;
        std::vector<std::array<size_t, 2>> walk{{f, 0}};
        while(!walk.empty())
        {
            auto& [u, stage] = walk.back();
            if(number.count(u)) { walk.pop_back(); continue; }
            if(stage == 0) { stage = 1; walk.push_back({store[u].lo, 0}); }
            else if(stage == 1) { stage = 2; walk.push_back({store[u].hi, 0}); }
            else { number.emplace(u, order.size() + 2); order.push_back(u); walk.pop_back(); }
        }
        const size_t last = instrs.size() + order.size() + 1;              // position of stack entry 0
        for(size_t s = order.size(); s-- > 0;)
        {
            const auto& x = store[order[s]];
            instrs.push_back({last - number.at(x.lo), last - number.at(x.hi), x.var});
        }
        instrs.push_back(topsink());
        instrs.push_back(botsink());
        delims.push_back(instrs.size());
        return nr_bdds() - 1;
    }
    size_t bdd_and(size_t i, size_t j) { const std::array<size_t, 2> two{i, j}; return bdd_and(two.begin(), two.end()); }
    size_t bdd_and(int i, int j)                                               // integer literals, as in the reference (:73-82)
    {
        if(i < 0 || j < 0) throw std::invalid_argument("bdd_and: negative BDD number");
        return bdd_and(size_t(i), size_t(j));
    }

    // ---------------------------------------------------------------------------------------- splitting
    // Cuts the quasi-reduced BDD b into chunks of at most chunk_size variables linked by one-hot auxiliary variables (numbered from
    // aux_var_start on) and appends the chunks; with_implication_bdd adds, when there are at least three chunks and the BDD's paths rule
    // out some combination of cut nodes, one more BDD over the auxiliary variables alone: per cut exactly one node is used, and a node
    // of one cut implies the nodes of every other cut it is connected to by a path (:805-940).  Returns the numbers of the new BDDs
    // ({b} when nothing was cut; the caller removes b otherwise, bdd_preprocessor.cpp:393-410) and the next free auxiliary variable.
    // Throws std::invalid_argument when a cut would land in front of a layer of width 1 (the reference asserts, :598).
    std::tuple<std::vector<size_t>, size_t> split_qbdd(size_t b, size_t chunk_size, size_t aux_var_start, bool with_implication_bdd = false)
    {
        need(b);
        if(chunk_size == 0) throw std::invalid_argument("split_qbdd: chunk size must be positive");
        if(!is_layered_qbdd(b)) throw std::invalid_argument("split_qbdd: BDD " + std::to_string(b) + " is not a layered quasi-reduced BDD");
        // a private copy with indices relative to the BDD: the chunks are appended to the array the BDD lives in
        const size_t first = delims[b], n = delims[b + 1] - first;
        std::vector<bddb200_instruction> src(instrs.begin() + first, instrs.begin() + first + n);
        for(bddb200_instruction& x : src) if(x.index < BOTSINK) { x.lo -= first; x.hi -= first; }
        size_t aux = aux_var_start;
        const size_t nr_before = nr_bdds();
        const size_t nr_chunks = split_detail::split_qbdd(src.data(), 0, n, chunk_size, aux, instrs.size(), instrs, delims);
        if(nr_chunks == 0) return {{b}, aux_var_start};
        std::vector<size_t> new_nrs(nr_chunks);
        for(size_t c = 0; c < nr_chunks; ++c) new_nrs[c] = nr_before + c;
        if(with_implication_bdd && nr_chunks > 2 && add_implication_bdd(src, chunk_size, nr_chunks, aux_var_start)) new_nrs.push_back(nr_bdds() - 1);
        return {new_nrs, aux};
    }

    // ---------------------------------------------------------------------------------------- export
    // The linear programme whose optimum is the bound the dual solvers approach, in .lp syntax (write_bdd_lp, header :731-830; "export
    // bdd lp" of the driver): one 0/1 variable arc_<bdd>_<node>_<value> per arc that does not end in the bot sink, one unit of flow
    // from every root (R_<bdd>), conserved at every other node (FC_<bdd>_<node>), and per BDD and layer the flow over the hi arcs
    // equal to the shared variable x_<var>.  Text identical to the reference's for BDDs over ascending variables
    // (tests/test_collection.py); for the chunks of a split BDD, whose auxiliary variables come first, the reference ties the first
    // layer to the wrong variable (:800) -- here every layer is tied to its own.
    template<typename STREAM>
    void write_bdd_lp(STREAM& s, const std::vector<double>& costs) const
    {
        auto arc = [](size_t b, size_t node, int value) { return "arc_" + std::to_string(b) + "_" + std::to_string(node) + "_" + std::to_string(value); };
        s << "Minimize\n";
        for(size_t v = 0; v < costs.size(); ++v) s << (costs[v] < 0 ? "-" : "+") << std::abs(costs[v]) << " x_" << v << "\n";
        s << "Subject To\n";
        std::vector<std::string> inflow;                                       // per node: " - arc" of every arc that enters it
        for(size_t b = 0; b < nr_bdds(); ++b)
        {
            const size_t first = delims[b], n = delims[b + 1] - 2 - first;
            inflow.assign(n, std::string());
            for(size_t k = 0; k < n; ++k)
            {
                const bddb200_instruction& x = instrs[first + k];
                if(k == 0) s << "R_" << b << ": "; else s << "FC_" << b << "_" << k << ": ";
                if(instrs[x.lo].index != BOTSINK) s << arc(b, k, 0);
                if(instrs[x.hi].index != BOTSINK) s << " + " << arc(b, k, 1);
                s << inflow[k] << (k == 0 ? " = 1\n" : " = 0\n");
                if(!is_terminal(x.lo)) inflow[x.lo - first] += " - " + arc(b, k, 0);
                if(!is_terminal(x.hi)) inflow[x.hi - first] += " - " + arc(b, k, 1);
            }
        }
        for(size_t b = 0; b < nr_bdds(); ++b)
        {
            const size_t first = delims[b], n = delims[b + 1] - 2 - first;
            size_t layer_var = instrs[first].index;      // (the reference starts from the BDD's smallest variable, which is the root's only when the variables ascend)
            for(size_t k = 0; k < n; ++k)
            {
                const bddb200_instruction& x = instrs[first + k];
                if(x.index != layer_var) { s << " - x_" << layer_var << " = 0\n"; layer_var = x.index; }
                if(instrs[x.hi].index != BOTSINK) s << " + " << arc(b, k, 1);
            }
            s << " - x_" << layer_var << " = 0\n";
        }
        s << "Bounds\nBinaries\n";
        for(size_t b = 0; b < nr_bdds(); ++b)
            for(size_t k = 0; k + 2 < delims[b + 1] - delims[b]; ++k)
                for(int value = 0; value < 2; ++value)
                    if(instrs[value ? instrs[delims[b] + k].hi : instrs[delims[b] + k].lo].index != BOTSINK) s << arc(b, k, value) << "\n";
        s << "End\n";
    }

    static bddb200_instruction botsink() { return {BOTSINK, BOTSINK, BOTSINK}; }
    static bddb200_instruction topsink() { return {TOPSINK, TOPSINK, TOPSINK}; }

private:
    void need(size_t b) const { if(b >= nr_bdds()) throw std::out_of_range("bdd_collection: no BDD " + std::to_string(b)); }
    // what every method relies on (bdd_basic_check, :951-988): the two sinks close the BDD, every other instruction is an inner node
    // whose children lie behind it inside the same BDD
    void check_structure(size_t b) const
    {
        const size_t first = delims[b], e = delims[b + 1];
        if(e - first < 3 || !((instrs[e - 1].index == TOPSINK && instrs[e - 2].index == BOTSINK) || (instrs[e - 1].index == BOTSINK && instrs[e - 2].index == TOPSINK)))
            throw std::invalid_argument("bdd_collection: BDD " + std::to_string(b) + " does not end in its two sinks");
        for(size_t i = first; i + 2 < e; ++i)
            if(instrs[i].index >= BOTSINK || instrs[i].lo <= i || instrs[i].lo >= e || instrs[i].hi <= i || instrs[i].hi >= e)
                throw std::invalid_argument("bdd_collection: instruction " + std::to_string(i - first) + " of BDD " + std::to_string(b) + " is not an inner node with its children behind it");
    }
    size_t close(size_t bot, size_t top)
    {
        if(instrs.size() != bot || top != bot + 1) throw std::logic_error("bdd_collection: generator size mismatch");
        instrs.push_back(botsink());
        instrs.push_back(topsink());
        delims.push_back(instrs.size());
        return nr_bdds() - 1;
    }
    static std::unordered_map<size_t, size_t> ranks(const std::vector<size_t>& vars)
    {
        std::unordered_map<size_t, size_t> r;
        r.reserve(vars.size());
        for(size_t i = 0; i < vars.size(); ++i) r.emplace(vars[i], i);
        return r;
    }
    // is_qbdd and is_reordered in one pass over the BDD (what split_qbdd needs of its input)
    bool is_layered_qbdd(size_t b) const
    {
        const std::vector<size_t> vars = variables(b), r = node_ranks(b, vars);
        const size_t first = delims[b], n = r.size();
        for(size_t i = 0; i < n; ++i)
        {
            if(i + 1 < n && r[i + 1] != r[i] && r[i + 1] != r[i] + 1) return false;
            for(const size_t child : {instrs[first + i].lo, instrs[first + i].hi})
            {
                if(child <= first + i || child >= delims[b + 1]) return false;
                if(instrs[child].index == BOTSINK) continue;
                if(r[i] + 1 == vars.size() ? instrs[child].index != TOPSINK : (child - first >= n || r[child - first] != r[i] + 1)) return false;
            }
        }
        return true;
    }
    // per inner node the position of its variable in `vars`: when `vars` ascends (the usual case) the next layer's variable is tried
    // first and a binary search decides otherwise; a topological order goes through a hash map, one lookup per run of equal variables
    std::vector<size_t> node_ranks(size_t b, const std::vector<size_t>& vars) const
    {
        const size_t first = delims[b], n = delims[b + 1] - 2 - first;
        std::vector<size_t> r(n);
        if(std::is_sorted(vars.begin(), vars.end()))
        {
            for(size_t i = 0; i < n; ++i)
            {
                const size_t var = instrs[first + i].index, prev = i > 0 ? r[i - 1] : 0;
                if(vars[prev] == var) r[i] = prev;
                else if(prev + 1 < vars.size() && vars[prev + 1] == var) r[i] = prev + 1;
                else r[i] = size_t(std::lower_bound(vars.begin(), vars.end(), var) - vars.begin());
            }
            return r;
        }
        const std::unordered_map<size_t, size_t> rank = ranks(vars);
        for(size_t i = 0; i < n; ++i)
            r[i] = i > 0 && instrs[first + i].index == instrs[first + i - 1].index ? r[i - 1] : rank.at(instrs[first + i].index);
        return r;
    }
    // BDD b as a function in the store; children sit behind their parents in the array (bdd_basic_check, :951-988)
    size_t load(collection_detail::robdd_store& store, size_t b) const
    {
        need(b);
        const size_t first = delims[b], n = delims[b + 1] - first;
        std::vector<size_t> f(n);
        for(size_t i = n; i-- > 0;)
        {
            const bddb200_instruction& x = instrs[first + i];
            if(x.index == TOPSINK) { f[i] = 1; continue; }
            if(x.index == BOTSINK) { f[i] = 0; continue; }
            for(const size_t child : {x.lo, x.hi})
                if(child <= first + i || child >= first + n || (!is_terminal(child) && instrs[child].index <= x.index))
                    throw std::invalid_argument("bdd_and: BDD " + std::to_string(b) + " is not ordered by ascending variables");
            f[i] = store.make(x.index, f[x.lo - first], f[x.hi - first]);
        }
        return f[0];
    }

    // src: the BDD that was cut, indices relative to its root.  Returns true when an implication BDD was appended.
    bool add_implication_bdd(const std::vector<bddb200_instruction>& src, size_t chunk_size, size_t nr_chunks, size_t aux_var_start)
    {
        const split_detail::Layers L = split_detail::layers_of(src.data(), 0, src.size());
        const size_t nr_cuts = nr_chunks - 1;                                  // cut c (1-based) sits in front of layer c * chunk_size
        auto width = [&](size_t c) { return L.widths[c * chunk_size]; };
        auto offset = [&](size_t c) { return L.offs[c * chunk_size]; };
        std::vector<size_t> aux_first(nr_cuts + 1, aux_var_start);           // auxiliary variable aux_first[c] + w - 1 - k stands for node k of cut c
        for(size_t c = 2; c <= nr_cuts; ++c) aux_first[c] = aux_first[c - 1] + width(c - 1);
        auto aux_var = [&](size_t c, size_t k) { return aux_first[c] + width(c) - 1 - k; };

        // reach[a][c - a - 1]: row k2 of cut c = bit set of the nodes of cut a with a path to node k2
        std::vector<std::vector<std::vector<uint64_t>>> reach(nr_cuts + 1);
        for(size_t a = 1; a < nr_cuts; ++a)
        {
            const size_t words = (width(a) + 63) / 64, begin = offset(a), end = offset(nr_cuts) + width(nr_cuts);
            std::vector<uint64_t> from((end - begin) * words, 0);
            for(size_t k = 0; k < width(a); ++k) from[k * words + k / 64] |= uint64_t(1) << (k % 64);
            for(size_t i = begin; i < end; ++i)
                for(const size_t child : {src[i].lo, src[i].hi})
                    if(child < end && src[child].index < BOTSINK)
                        for(size_t w = 0; w < words; ++w) from[(child - begin) * words + w] |= from[(i - begin) * words + w];
            for(size_t c = a + 1; c <= nr_cuts; ++c)
                reach[a].emplace_back(from.begin() + (offset(c) - begin) * words, from.begin() + (offset(c) - begin + width(c)) * words);
        }
        auto connected = [&](size_t a, size_t k1, size_t c, size_t k2) {
            const size_t words = (width(a) + 63) / 64;
            return (reach[a][c - a - 1][k2 * words + k1 / 64] >> (k1 % 64)) & 1;
        };

        const size_t nr_before = nr_bdds();
        for(size_t c = 1; c <= nr_cuts; ++c)
        {
            std::vector<size_t> vars(width(c));
            for(size_t k = 0; k < vars.size(); ++k) vars[k] = aux_first[c] + k;
            const size_t s = simplex_constraint(vars.size());
            rebase(s, vars.begin(), vars.end());
        }
        // "node k1 of cut a is used" implies "one of the nodes of cut c it is connected to is used", in both directions;
        // a node connected to the whole other cut adds nothing beyond the simplex constraints
        auto implication = [&](size_t premise, std::vector<size_t> conclusion, size_t full) {
            if(conclusion.size() == full) return;
            conclusion.push_back(premise);
            std::sort(conclusion.begin(), conclusion.end());
            const size_t nr = not_all_false_constraint(conclusion.size());
            rebase(nr, conclusion.begin(), conclusion.end());
            invert(nr, premise);
        };
        for(size_t a = 1; a < nr_cuts; ++a)
            for(size_t c = a + 1; c <= nr_cuts; ++c)
            {
                for(size_t k1 = 0; k1 < width(a); ++k1)
                {
                    std::vector<size_t> ends;
                    for(size_t k2 = 0; k2 < width(c); ++k2) if(connected(a, k1, c, k2)) ends.push_back(aux_var(c, k2));
                    implication(aux_var(a, k1), ends, width(c));
                }
                for(size_t k2 = 0; k2 < width(c); ++k2)
                {
                    std::vector<size_t> starts;
                    for(size_t k1 = 0; k1 < width(a); ++k1) if(connected(a, k1, c, k2)) starts.push_back(aux_var(a, k1));
                    implication(aux_var(c, k2), starts, width(a));
                }
            }
        std::vector<size_t> parts(nr_bdds() - nr_before);
        for(size_t i = 0; i < parts.size(); ++i) parts[i] = nr_before + i;
        if(parts.size() == nr_cuts) { remove(parts.begin(), parts.end()); return false; }          // nothing beyond the simplex constraints (:907-911)
        size_t result = bdd_and(parts.begin(), parts.end());
        reorder(result);
        if(!is_qbdd(result)) { parts.push_back(result); result = make_qbdd(result); }
        remove(parts.begin(), parts.end());
        return true;
    }
};

// The driver loop of the preprocessor with a forced split length (bdd_preprocessor.cpp:372-415), in place: every BDD over more than
// split_length variables is replaced by its chunks (appended behind the kept BDDs in BDD order, each followed by its implication BDD
// when asked for and non-trivial); auxiliary variables are numbered from nr_variables on.  Returns the total number of variables.
// A BDD whose cut would land in front of a layer of width 1 stays whole.
inline size_t split_long_bdds(bdd_collection& col, size_t split_length, size_t nr_variables, bool with_implication_bdd, size_t* n_split_out = nullptr)
{
    if(split_length == 0) throw std::invalid_argument("split length must be positive");
    size_t aux = nr_variables;
    for(const bddb200_instruction& x : col.instrs) if(x.index < bdd_collection::BOTSINK) aux = std::max(aux, x.index + 1);
    std::vector<size_t> replaced;
    const size_t nr_orig = col.nr_bdds();
    size_t long_nodes = 0;                             // the chunks repeat the nodes of the BDDs they replace and add the gadgets
    std::vector<char> is_long(nr_orig, 0);
    for(size_t b = 0; b < nr_orig; ++b)
        if(col.layer_offsets(b).size() > split_length) { is_long[b] = 1; long_nodes += col.nr_bdd_nodes(b); }
    col.instrs.reserve(col.instrs.size() + long_nodes + long_nodes / 2);
    for(size_t b = 0; b < nr_orig; ++b)
    {
        if(!is_long[b]) continue;
        try
        {
            const auto [new_nrs, next_aux] = col.split_qbdd(b, split_length, aux, with_implication_bdd);
            if(new_nrs.size() > 1) { replaced.push_back(b); aux = next_aux; }
        }
        catch(const std::invalid_argument&) {}
    }
    col.remove(replaced.begin(), replaced.end());
    if(n_split_out) *n_split_out = replaced.size();
    return aux;
}

} // namespace bddb200_host
