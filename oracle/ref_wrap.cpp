// oracle/ref_wrap.cpp -- C wrapper around the UNMODIFIED reference CPU sources.
//
// TEST INFRASTRUCTURE ONLY.  This file is compiled together with the reference's own
// source files (where they lie under /root/reference; see oracle/Makefile) into
// oracle/_ref/libbdd_ref.so.  It exposes, through a plain C interface that ctypes can
// bind,
//   * the reference's constraint -> BDD conversion
//       (src/bdd_conversion/bdd_preprocessor.cpp:175-228 is the call sequence followed
//        here: simplex shortcut, else convert_to_bdd -> add_bdd -> reorder -> make_qbdd
//        -> rebase),
//   * the reference's CPU `parallel mma` solver
//       bdd_parallel_mma_base<bdd_branch_instruction<REAL,uint16_t>>
//       (src/bdd_solver/bdd_parallel_mma_base.cpp), the type behind the config string
//       "parallel mma" (src/bdd_solver/bdd_solver.cpp:153-163).
// Nothing in the product path (bdd_b200/) links or loads this library; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
// No reference source text is copied: this file only CALLS the reference's public API.

#include <vector>
#include <array>
#include <memory>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <iostream>
#include <sstream>
#include <iterator>
#include <ostream>
#include <limits>
#include <cassert>
// refw_collection_from_arrays must fill a bdd_collection with QBDDs verbatim.  The public
// builder (new_bdd / add_bdd_node / close_bdd) runs reduce() on close (bdd_collection.cpp:1610),
// which would strip the lo == hi nodes a QBDD needs, so this TEST HARNESS reaches the two
// private vectors directly.  Access control does not change the class layout.
#include "bdd_manager/bdd_mgr.h"
#define private public
#include "bdd_collection/bdd_collection.h"
#undef private
#include "bdd_conversion/convert_pb_to_bdd.h"
#include "bdd_solver/bdd_parallel_mma_base.h"
#include "bdd_solver/bdd_branch_instruction.h"
#include "mm_primal_decoder.h"
#include <vector>
#include <array>
#include <memory>
#include <cstring>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace LPMP;

namespace {

struct ref_collection {
    BDD::bdd_collection col;
    BDD::bdd_mgr mgr;
    std::unique_ptr<bdd_converter> conv;
    std::string last_error;
    ref_collection() : conv(new bdd_converter(mgr)) {}
};

template<typename REAL>
using ref_solver_t = bdd_parallel_mma_base<bdd_branch_instruction<REAL, uint16_t>>;

struct ref_solver {
    int is_double;
    std::unique_ptr<ref_solver_t<double>> d;
    std::unique_ptr<ref_solver_t<float>> f;
    std::vector<std::array<double,2>> delta_d;
    std::vector<std::array<float,2>> delta_f;
};

template<typename F>
auto visit(ref_solver* s, F&& f) { if(s->is_double) return f(*s->d); else return f(*s->f); }

} // namespace

extern "C" {

// ---------------------------------------------------------------- collection ---------
void* refw_collection_new() { return new ref_collection(); }
void refw_collection_free(void* c) { delete static_cast<ref_collection*>(c); }
const char* refw_collection_error(void* c) { return static_cast<ref_collection*>(c)->last_error.c_str(); }

// ineq: 0 '<=', 1 '>=', 2 '='.  Returns the new bdd number, -1 if the constraint is
// trivially satisfied (no BDD emitted), -2 on error (message in refw_collection_error).
long refw_add_constraint(void* c, const int* coeffs, const size_t* vars, size_t n, int ineq, int rhs)
{
    ref_collection* rc = static_cast<ref_collection*>(c);
    try {
        const std::vector<int> coefficients(coeffs, coeffs + n);
        const std::vector<size_t> variables(vars, vars + n);
        const ILP_input::inequality_type it = ineq == 0 ? ILP_input::inequality_type::smaller_equal
            : (ineq == 1 ? ILP_input::inequality_type::greater_equal : ILP_input::inequality_type::equal);
        bool simplex = (ineq == 2) && rhs != 0;
        for(size_t i = 0; i < n && simplex; ++i)
            if(coefficients[i] != rhs)
                simplex = false;
        if(simplex)
        {
            const size_t bdd_nr = rc->col.simplex_constraint(n);
            rc->col.rebase(bdd_nr, variables.begin(), variables.end());
            return long(bdd_nr);
        }
        BDD::node_ref bdd = rc->conv->convert_to_bdd(coefficients, it, rhs);
        if(bdd.is_topsink())
            return -1;
        if(bdd.is_botsink())
            throw std::runtime_error("problem is infeasible");
        size_t bdd_nr = rc->col.add_bdd(bdd);
        rc->col.reorder(bdd_nr);
        if(!rc->col.is_qbdd(bdd_nr))
        {
            rc->col.make_qbdd(bdd_nr);
            rc->col.remove(bdd_nr);
        }
        rc->col.rebase(bdd_nr, variables.begin(), variables.end());
        return long(bdd_nr);
    } catch(const std::exception& e) {
        rc->last_error = e.what();
        return -2;
    }
}

size_t refw_nr_bdds(void* c) { return static_cast<ref_collection*>(c)->col.nr_bdds(); }

// bdd_collection::split_qbdd(bdd_nr, chunk_size, aux_var_start, false) (bdd_collection.cpp:507-790): appends the chunk BDDs,
// returns the next free auxiliary variable (or (size_t)-1 on error); *nr_new = number of BDDs appended (1 = not split).
size_t refw_split_qbdd(void* c, size_t bdd_nr, size_t chunk_size, size_t aux_var_start, size_t* nr_new)
{
    auto* rc = static_cast<ref_collection*>(c);
    try {
        const auto [new_nrs, next_aux] = rc->col.split_qbdd(bdd_nr, chunk_size, aux_var_start, false);
        *nr_new = new_nrs.size();
        return next_aux;
    } catch(const std::exception& e) { rc->last_error = e.what(); return (size_t)-1; }
}

// The reference's direct generators and structural operations (bdd_collection.cpp:2039-2263, :1429, :1670, :31-315, :2023-2037),
// called one to one so that tests/test_collection.py can compare bddb200_host::bdd_collection with them bit for bit.
size_t refw_simplex_constraint(void* c, size_t n) { return static_cast<ref_collection*>(c)->col.simplex_constraint(n); }
size_t refw_not_all_false_constraint(void* c, size_t n) { return static_cast<ref_collection*>(c)->col.not_all_false_constraint(n); }
size_t refw_all_equal_constraint(void* c, size_t n) { return static_cast<ref_collection*>(c)->col.all_equal_constraint(n); }
size_t refw_cardinality_constraint(void* c, size_t n, size_t k) { return static_cast<ref_collection*>(c)->col.cardinality_constraint(n, k); }
void refw_rebase(void* c, size_t bdd_nr, const size_t* vars, size_t n) { static_cast<ref_collection*>(c)->col.rebase(bdd_nr, vars, vars + n); }
void refw_negate(void* c, size_t bdd_nr) { static_cast<ref_collection*>(c)->col.negate(bdd_nr); }
void refw_invert(void* c, size_t bdd_nr, size_t var) { static_cast<ref_collection*>(c)->col.invert(bdd_nr, var); }
void refw_reorder(void* c, size_t bdd_nr) { static_cast<ref_collection*>(c)->col.reorder(bdd_nr); }
size_t refw_make_qbdd(void* c, size_t bdd_nr) { return static_cast<ref_collection*>(c)->col.make_qbdd(bdd_nr); }
size_t refw_bdd_and(void* c, const size_t* nrs, size_t n) { return static_cast<ref_collection*>(c)->col.bdd_and(nrs, nrs + n); }
void refw_remove(void* c, const size_t* nrs, size_t n) { static_cast<ref_collection*>(c)->col.remove(nrs, nrs + n); }
int refw_is_qbdd(void* c, size_t bdd_nr) { return static_cast<ref_collection*>(c)->col.is_qbdd(bdd_nr) ? 1 : 0; }
int refw_is_reordered(void* c, size_t bdd_nr) { return static_cast<ref_collection*>(c)->col.is_reordered(bdd_nr) ? 1 : 0; }
size_t refw_variables(void* c, size_t bdd_nr, size_t* out)
{
    const auto vars = static_cast<ref_collection*>(c)->col.variables(bdd_nr);
    if(out != nullptr) std::copy(vars.begin(), vars.end(), out);
    return vars.size();
}
// split_qbdd with the implication BDD switched on (bdd_collection.cpp:805-940); same return convention as refw_split_qbdd
size_t refw_split_qbdd_implication(void* c, size_t bdd_nr, size_t chunk_size, size_t aux_var_start, size_t* nr_new)
{
    auto* rc = static_cast<ref_collection*>(c);
    try {
        const auto [new_nrs, next_aux] = rc->col.split_qbdd(bdd_nr, chunk_size, aux_var_start, true);
        *nr_new = new_nrs.size();
        return next_aux;
    } catch(const std::exception& e) { rc->last_error = e.what(); return (size_t)-1; }
}

// bdd_collection::write_bdd_lp (include/bdd_collection/bdd_collection.h:731-830) into a caller buffer; returns the length of the text
size_t refw_write_bdd_lp(void* c, const double* costs, size_t n_costs, char* out, size_t capacity)
{
    std::ostringstream s;
    static_cast<ref_collection*>(c)->col.write_bdd_lp(s, costs, costs + n_costs);
    const std::string text = s.str();
    if(out != nullptr) std::memcpy(out, text.data(), std::min(capacity, text.size()));
    return text.size();
}

size_t refw_nr_instructions(void* c)
{
    const BDD::bdd_collection& col = static_cast<ref_collection*>(c)->col;
    return col.nr_bdds() == 0 ? 0 : col.offset(col.nr_bdds()-1) + col.nr_bdd_nodes(col.nr_bdds()-1);
}

// triples: 3*nr_instructions size_t values {lo, hi, index} in the layout of
// BDD::bdd_instruction (include/bdd_collection/bdd_collection.h:14-17).
void refw_export(void* c, size_t* triples, size_t* delimiters)
{
    const BDD::bdd_collection& col = static_cast<ref_collection*>(c)->col;
    const size_t n = refw_nr_instructions(c);
    for(size_t i = 0; i < n; ++i)
    {
        const BDD::bdd_instruction& instr = col.get_bdd_instruction(i);
        triples[3*i+0] = instr.lo;
        triples[3*i+1] = instr.hi;
        triples[3*i+2] = instr.index;
    }
    for(size_t b = 0; b < col.nr_bdds(); ++b)
        delimiters[b] = col.offset(b);
    delimiters[col.nr_bdds()] = n;
}

// Rebuild a reference bdd_collection from flat arrays (instruction triples are copied
// verbatim into bdd_instructions / bdd_delimiters, include/bdd_collection/bdd_collection.h:282-283).
void* refw_collection_from_arrays(const size_t* triples, size_t n_instr, const size_t* delimiters, size_t n_bdds)
{
    ref_collection* rc = new ref_collection();
    rc->col.bdd_instructions.resize(n_instr);
    for(size_t i = 0; i < n_instr; ++i)
    {
        rc->col.bdd_instructions[i].lo = triples[3*i];
        rc->col.bdd_instructions[i].hi = triples[3*i+1];
        rc->col.bdd_instructions[i].index = triples[3*i+2];
    }
    rc->col.bdd_delimiters.assign(delimiters, delimiters + n_bdds + 1);
    for(size_t b = 0; b < n_bdds; ++b)
        if(!rc->col.is_qbdd(b))
        {
            rc->last_error = "BDD " + std::to_string(b) + " is not a QBDD";
            break;
        }
    return rc;
}

// ---------------------------------------------------------------- solver -------------
void refw_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int refw_max_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// costs may be NULL (n_costs = 0): solver is built with zero costs (one-argument ctor).
void* refw_solver_new(void* c, const double* costs, size_t n_costs, int is_double)
{
    ref_collection* rc = static_cast<ref_collection*>(c);
    ref_solver* s = new ref_solver();
    s->is_double = is_double;
    if(costs != nullptr)
    {
        const std::vector<double> cv(costs, costs + n_costs);
        if(is_double) s->d.reset(new ref_solver_t<double>(rc->col, cv));
        else s->f.reset(new ref_solver_t<float>(rc->col, cv));
    }
    else
    {
        if(is_double) s->d.reset(new ref_solver_t<double>(rc->col));
        else s->f.reset(new ref_solver_t<float>(rc->col));
    }
    return s;
}
void refw_solver_free(void* s) { delete static_cast<ref_solver*>(s); }

size_t refw_solver_nr_variables(void* s) { return visit(static_cast<ref_solver*>(s), [](auto& x) { return x.nr_variables(); }); }
size_t refw_solver_nr_bdds(void* s) { return visit(static_cast<ref_solver*>(s), [](auto& x) { return x.nr_bdds(); }); }
size_t refw_solver_nr_bdds_of_var(void* s, size_t v) { return visit(static_cast<ref_solver*>(s), [&](auto& x) { return x.nr_bdds(v); }); }
size_t refw_solver_nr_layers(void* s) { return visit(static_cast<ref_solver*>(s), [](auto& x) { return x.nr_layers(); }); }

double refw_solver_lower_bound(void* s) { return visit(static_cast<ref_solver*>(s), [](auto& x) { return x.lower_bound(); }); }
void refw_solver_iteration(void* s) { visit(static_cast<ref_solver*>(s), [](auto& x) { x.iteration(); return 0; }); }
void refw_solver_distribute_delta(void* s) { visit(static_cast<ref_solver*>(s), [](auto& x) { x.distribute_delta(); return 0; }); }

// Costs are only settable at construction (refw_solver_new): the reference's public
// two-vector update_costs overload passes its iterators in the wrong order at this commit
// (bdd_parallel_mma_base.cpp:623), while the constructor path (:30) is correct.

// delta: 2*nr_variables values, interleaved lo/hi, in and out (double precision on the
// C side; converted to the solver's REAL).
static void pass_impl(ref_solver* s, double omega, double* delta, bool forward)
{
    if(s->is_double)
    {
        auto& x = *s->d;
        s->delta_d.resize(x.nr_variables());
        for(size_t v = 0; v < x.nr_variables(); ++v) s->delta_d[v] = {delta[2*v], delta[2*v+1]};
        if(forward) x.forward_mm(omega, s->delta_d); else x.backward_mm(omega, s->delta_d);
        for(size_t v = 0; v < x.nr_variables(); ++v) { delta[2*v] = s->delta_d[v][0]; delta[2*v+1] = s->delta_d[v][1]; }
    }
    else
    {
        auto& x = *s->f;
        s->delta_f.resize(x.nr_variables());
        for(size_t v = 0; v < x.nr_variables(); ++v) s->delta_f[v] = {float(delta[2*v]), float(delta[2*v+1])};
        if(forward) x.forward_mm(float(omega), s->delta_f); else x.backward_mm(float(omega), s->delta_f);
        for(size_t v = 0; v < x.nr_variables(); ++v) { delta[2*v] = s->delta_f[v][0]; delta[2*v+1] = s->delta_f[v][1]; }
    }
}
void refw_solver_forward_mm(void* s, double omega, double* delta) { pass_impl(static_cast<ref_solver*>(s), omega, delta, true); }
void refw_solver_backward_mm(void* s, double omega, double* delta) { pass_impl(static_cast<ref_solver*>(s), omega, delta, false); }

// Min-marginals in the reference's (variable, k-th BDD containing it) order
// (bdd_parallel_mma_base.cpp:373-414).  out has 2 * sum_v nr_bdds(v) doubles.
size_t refw_solver_min_marginals(void* s, double* out)
{
    return visit(static_cast<ref_solver*>(s), [&](auto& x) {
        const auto mm = x.min_marginals();
        size_t c = 0;
        for(size_t v = 0; v < mm.size(); ++v)
            for(size_t j = 0; j < mm.size(v); ++j, ++c)
                if(out != nullptr) { out[2*c] = mm(v,j)[0]; out[2*c+1] = mm(v,j)[1]; }
        return c;
    });
}

// L-BFGS support surface of the CPU solver (bdd_parallel_mma_base.cpp:1197-1400).
void refw_solver_bdds_solution(void* s, char* out)
{
    visit(static_cast<ref_solver*>(s), [&](auto& x) { const auto v = x.bdds_solution_vec(); std::memcpy(out, v.data(), v.size()); return 0; });
}
void refw_solver_net_solver_costs(void* s, double* out)
{
    visit(static_cast<ref_solver*>(s), [&](auto& x) { const auto v = x.net_solver_costs(); for(size_t i = 0; i < v.size(); ++i) out[i] = v[i]; return 0; });
}

// The reference's min-marginal decoder of the CPU primal rounding (include/mm_primal_decoder.h, src/bdd_solver/mm_primal_decoder.cpp;
// used by include/bdd_solver/incremental_mm_agreement_rounding.hxx:78-140): per variable the agreement type (0 zero, 1 one, 2 equal,
// 3 inconsistent), the sums of the min-marginals, the type statistics {#one, #zero, #equal, #inconsistent} and, when every variable is
// zero / one, the solution.  mms: 2 doubles per (variable, BDD) entry, variable-major; counts[v] entries for variable v.
int refw_mm_decode(const double* mms, const size_t* counts, size_t n_vars, char* types_out, double* sums_out, size_t* stats_out, char* solution_out)
{
    std::vector<std::vector<std::array<double,2>>> rows(n_vars);
    size_t c = 0;
    for(size_t v = 0; v < n_vars; ++v)
        for(size_t j = 0; j < counts[v]; ++j, ++c) rows[v].push_back({mms[2*c], mms[2*c+1]});
    two_dim_variable_array<std::array<double,2>> arr(rows);
    const mm_primal_decoder dec(std::move(arr));
    for(size_t v = 0; v < n_vars; ++v)
    {
        const mm_type t = dec.compute_mm_type(v);
        types_out[v] = t == mm_type::zero ? 0 : (t == mm_type::one ? 1 : (t == mm_type::equal ? 2 : 3));
        const auto sum = dec.mm_sum(v);
        sums_out[2*v] = sum[0]; sums_out[2*v+1] = sum[1];
    }
    const auto st = dec.mm_type_statistics();
    for(int k = 0; k < 4; ++k) stats_out[k] = st[k];
    if(dec.can_reconstruct_solution() && solution_out != nullptr)
    {
        const auto sol = dec.solution_from_mms();
        std::memcpy(solution_out, sol.data(), sol.size());
        return 1;
    }
    return 0;
}

} // extern "C"
