// bdd_b200/csrc/host/collection_abi.cpp -- the C ABI of include/bdd_b200_collection.h over bddb200_host::bdd_collection.
// Host code only (compiled by g++ and linked into libbdd_b200.so next to the CUDA translation unit); failures go to the same
// per-thread string bddb200_last_error() returns.
#include <algorithm>
#include <fstream>
#include <memory>
#include <new>
#include <string>

#include "../../../include/bdd_b200_collection.h"
#include "../last_error.hpp"
#include "bdd_collection.hpp"
#include "lp_reader.hpp"

struct bddb200_collection { bddb200_host::bdd_collection col; };
struct bddb200_ilp { bddb200_host::ILP ilp; };

namespace {
template<typename F>
int guarded(F&& f)
{
    try { f(); return BDDB200_OK; }
    catch(const std::bad_alloc&) { bddb200::detail::set_last_error("out of host memory"); return BDDB200_ERR_INVALID_ARGUMENT; }
    catch(const std::exception& e) { bddb200::detail::set_last_error(e.what()); return BDDB200_ERR_INVALID_ARGUMENT; }
}
int fail(const char* message) { bddb200::detail::set_last_error(message); return BDDB200_ERR_INVALID_ARGUMENT; }
}

extern "C" {

#define REQUIRE_COLLECTION(c) if((c) == nullptr) return fail("null collection handle");
#define REQUIRE_OUT(p) if((p) == nullptr) return fail("null argument");

int bddb200_collection_create(const bddb200_instruction* instrs, size_t n_instr, const size_t* delims, size_t n_bdds, bddb200_collection** out)
{
    REQUIRE_OUT(out);
    *out = nullptr;
    if(n_bdds > 0 && (instrs == nullptr || delims == nullptr)) return fail("null argument");
    return guarded([&] {
        std::unique_ptr<bddb200_collection> c(new bddb200_collection());
        if(n_bdds > 0) c->col = bddb200_host::bdd_collection(instrs, n_instr, delims, n_bdds);
        *out = c.release();
    });
}
int bddb200_collection_destroy(bddb200_collection* c) { delete c; return BDDB200_OK; }
int bddb200_collection_nr_bdds(const bddb200_collection* c, size_t* out) { REQUIRE_COLLECTION(c); REQUIRE_OUT(out); *out = c->col.nr_bdds(); return BDDB200_OK; }
int bddb200_collection_nr_instructions(const bddb200_collection* c, size_t* out) { REQUIRE_COLLECTION(c); REQUIRE_OUT(out); *out = c->col.instrs.size(); return BDDB200_OK; }
int bddb200_collection_export(const bddb200_collection* c, bddb200_instruction* instrs_out, size_t* delims_out)
{
    REQUIRE_COLLECTION(c); REQUIRE_OUT(delims_out);
    if(instrs_out != nullptr) std::copy(c->col.instrs.begin(), c->col.instrs.end(), instrs_out);
    std::copy(c->col.delims.begin(), c->col.delims.end(), delims_out);
    return BDDB200_OK;
}

int bddb200_collection_simplex_constraint(bddb200_collection* c, size_t n, size_t* bdd_nr_out)
{ REQUIRE_COLLECTION(c); REQUIRE_OUT(bdd_nr_out); return guarded([&] { *bdd_nr_out = c->col.simplex_constraint(n); }); }
int bddb200_collection_not_all_false_constraint(bddb200_collection* c, size_t n, size_t* bdd_nr_out)
{ REQUIRE_COLLECTION(c); REQUIRE_OUT(bdd_nr_out); return guarded([&] { *bdd_nr_out = c->col.not_all_false_constraint(n); }); }
int bddb200_collection_all_equal_constraint(bddb200_collection* c, size_t n, size_t* bdd_nr_out)
{ REQUIRE_COLLECTION(c); REQUIRE_OUT(bdd_nr_out); return guarded([&] { *bdd_nr_out = c->col.all_equal_constraint(n); }); }
int bddb200_collection_cardinality_constraint(bddb200_collection* c, size_t n, size_t k, size_t* bdd_nr_out)
{ REQUIRE_COLLECTION(c); REQUIRE_OUT(bdd_nr_out); return guarded([&] { *bdd_nr_out = c->col.cardinality_constraint(n, k); }); }

int bddb200_collection_add_linear_constraint(bddb200_collection* c, const long long* coefficients, const size_t* variables, size_t n,
                                             int relation, long long rhs, size_t* bdd_nr_out)
{
    REQUIRE_COLLECTION(c); REQUIRE_OUT(coefficients); REQUIRE_OUT(variables); REQUIRE_OUT(bdd_nr_out);
    return guarded([&] {
        *bdd_nr_out = c->col.add_linear_constraint(std::vector<long long>(coefficients, coefficients + n), std::vector<size_t>(variables, variables + n), relation, rhs);
    });
}

int bddb200_collection_rebase(bddb200_collection* c, size_t bdd_nr, const size_t* vars, size_t n_vars)
{ REQUIRE_COLLECTION(c); REQUIRE_OUT(vars); return guarded([&] { c->col.rebase(bdd_nr, vars, vars + n_vars); }); }
int bddb200_collection_negate(bddb200_collection* c, size_t bdd_nr) { REQUIRE_COLLECTION(c); return guarded([&] { c->col.negate(bdd_nr); }); }
int bddb200_collection_invert(bddb200_collection* c, size_t bdd_nr, size_t var) { REQUIRE_COLLECTION(c); return guarded([&] { c->col.invert(bdd_nr, var); }); }

int bddb200_collection_variables(const bddb200_collection* c, size_t bdd_nr, size_t* vars_out, size_t capacity, size_t* n_out)
{
    REQUIRE_COLLECTION(c); REQUIRE_OUT(n_out);
    return guarded([&] {
        const std::vector<size_t> vars = c->col.variables(bdd_nr);
        *n_out = vars.size();
        if(vars_out != nullptr) std::copy(vars.begin(), vars.begin() + std::min(capacity, vars.size()), vars_out);
    });
}
int bddb200_collection_is_qbdd(const bddb200_collection* c, size_t bdd_nr, int* out)
{ REQUIRE_COLLECTION(c); REQUIRE_OUT(out); return guarded([&] { *out = c->col.is_qbdd(bdd_nr) ? 1 : 0; }); }
int bddb200_collection_is_reordered(const bddb200_collection* c, size_t bdd_nr, int* out)
{ REQUIRE_COLLECTION(c); REQUIRE_OUT(out); return guarded([&] { *out = c->col.is_reordered(bdd_nr) ? 1 : 0; }); }
int bddb200_collection_evaluate(const bddb200_collection* c, size_t bdd_nr, const char* labeling, size_t n, int* out)
{
    REQUIRE_COLLECTION(c); REQUIRE_OUT(out); REQUIRE_OUT(labeling);
    return guarded([&] { *out = c->col.evaluate(bdd_nr, std::vector<char>(labeling, labeling + n)) ? 1 : 0; });
}
int bddb200_collection_reorder(bddb200_collection* c, size_t bdd_nr) { REQUIRE_COLLECTION(c); return guarded([&] { c->col.reorder(bdd_nr); }); }
int bddb200_collection_make_qbdd(bddb200_collection* c, size_t bdd_nr, size_t* bdd_nr_out)
{ REQUIRE_COLLECTION(c); REQUIRE_OUT(bdd_nr_out); return guarded([&] { *bdd_nr_out = c->col.make_qbdd(bdd_nr); }); }
int bddb200_collection_bdd_and(bddb200_collection* c, const size_t* bdd_nrs, size_t n, size_t* bdd_nr_out)
{ REQUIRE_COLLECTION(c); REQUIRE_OUT(bdd_nrs); REQUIRE_OUT(bdd_nr_out); return guarded([&] { *bdd_nr_out = c->col.bdd_and(bdd_nrs, bdd_nrs + n); }); }
int bddb200_collection_remove(bddb200_collection* c, const size_t* bdd_nrs, size_t n)
{ REQUIRE_COLLECTION(c); if(n == 0) return BDDB200_OK; REQUIRE_OUT(bdd_nrs); return guarded([&] { c->col.remove(bdd_nrs, bdd_nrs + n); }); }

int bddb200_create_from_collection(const bddb200_collection* c, const double* costs_hi, size_t n_costs, int precision,
                                   const bddb200_options* opts, bddb200_solver** out)
{
    REQUIRE_COLLECTION(c); REQUIRE_OUT(out);
    if(c->col.nr_bdds() == 0) { *out = nullptr; return fail("empty BDD collection"); }
    return bddb200_create(c->col.instrs.data(), c->col.instrs.size(), c->col.delims.data(), c->col.nr_bdds(), costs_hi, n_costs, precision, opts, out);
}

int bddb200_collection_write_bdd_lp(const bddb200_collection* c, const double* costs, size_t n_costs, const char* path)
{
    REQUIRE_COLLECTION(c); REQUIRE_OUT(path);
    if(n_costs > 0 && costs == nullptr) return fail("null argument");
    return guarded([&] {
        std::ofstream f(path);
        if(!f.good()) throw std::runtime_error(std::string("cannot write ") + path);
        c->col.write_bdd_lp(f, std::vector<double>(costs, costs + n_costs));
    });
}

int bddb200_collection_split_qbdd(bddb200_collection* c, size_t bdd_nr, size_t chunk_size, size_t aux_var_start, int with_implication_bdd,
                                  size_t* nr_new_out, size_t* next_aux_out)
{
    REQUIRE_COLLECTION(c); REQUIRE_OUT(nr_new_out); REQUIRE_OUT(next_aux_out);
    return guarded([&] {
        const auto [new_nrs, next_aux] = c->col.split_qbdd(bdd_nr, chunk_size, aux_var_start, with_implication_bdd != 0);
        *nr_new_out = new_nrs.size() == 1 && new_nrs[0] == bdd_nr ? 0 : new_nrs.size();
        *next_aux_out = next_aux;
    });
}
int bddb200_collection_split_long_bdds(bddb200_collection* c, size_t split_length, size_t nr_variables, int with_implication_bdd,
                                       size_t* n_split_out, size_t* nr_variables_out)
{
    REQUIRE_COLLECTION(c); REQUIRE_OUT(nr_variables_out);
    return guarded([&] { *nr_variables_out = bddb200_host::split_long_bdds(c->col, split_length, nr_variables, with_implication_bdd != 0, n_split_out); });
}

// ---- ILP input ----
#define REQUIRE_ILP(p) if((p) == nullptr) return fail("null ILP handle");

int bddb200_ilp_read(const char* file_or_text, bddb200_ilp** out)
{
    REQUIRE_OUT(out); *out = nullptr; REQUIRE_OUT(file_or_text);
    return guarded([&] {
        std::unique_ptr<bddb200_ilp> p(new bddb200_ilp());
        p->ilp = bddb200_host::read_ilp(file_or_text);
        *out = p.release();
    });
}
int bddb200_ilp_destroy(bddb200_ilp* ilp) { delete ilp; return BDDB200_OK; }
int bddb200_ilp_nr_variables(const bddb200_ilp* ilp, size_t* out) { REQUIRE_ILP(ilp); REQUIRE_OUT(out); *out = ilp->ilp.nr_variables(); return BDDB200_OK; }
int bddb200_ilp_nr_constraints(const bddb200_ilp* ilp, size_t* out) { REQUIRE_ILP(ilp); REQUIRE_OUT(out); *out = ilp->ilp.constraints.size(); return BDDB200_OK; }
int bddb200_ilp_objective(const bddb200_ilp* ilp, double* coefficients_out, double* constant_out)
{
    REQUIRE_ILP(ilp);
    if(coefficients_out != nullptr) std::copy(ilp->ilp.objective.begin(), ilp->ilp.objective.end(), coefficients_out);
    if(constant_out != nullptr) *constant_out = ilp->ilp.constant;
    return BDDB200_OK;
}
int bddb200_ilp_variable_name(const bddb200_ilp* ilp, size_t var, const char** name_out)
{
    REQUIRE_ILP(ilp); REQUIRE_OUT(name_out);
    if(var >= ilp->ilp.nr_variables()) return fail("no such variable");
    *name_out = ilp->ilp.var_names[var].c_str();
    return BDDB200_OK;
}
int bddb200_ilp_constraint(const bddb200_ilp* ilp, size_t c, size_t* n_out, size_t* variables_out, long long* coefficients_out, size_t capacity,
                           int* relation_out, long long* rhs_out)
{
    REQUIRE_ILP(ilp); REQUIRE_OUT(n_out);
    if(c >= ilp->ilp.constraints.size()) return fail("no such constraint");
    const bddb200_host::Constraint& k = ilp->ilp.constraints[c];
    *n_out = k.variables.size();
    const size_t n = std::min(capacity, k.variables.size());
    if(variables_out != nullptr) std::copy(k.variables.begin(), k.variables.begin() + n, variables_out);
    if(coefficients_out != nullptr) std::copy(k.coefficients.begin(), k.coefficients.begin() + n, coefficients_out);
    if(relation_out != nullptr) *relation_out = k.ineq;
    if(rhs_out != nullptr) *rhs_out = k.rhs;
    return BDDB200_OK;
}
int bddb200_ilp_to_bdds(const bddb200_ilp* ilp, bddb200_collection** out)
{
    REQUIRE_ILP(ilp); REQUIRE_OUT(out); *out = nullptr;
    return guarded([&] {
        std::unique_ptr<bddb200_collection> c(new bddb200_collection());
        c->col = bddb200_host::bdds_from_ilp(ilp->ilp);
        *out = c.release();
    });
}

} // extern "C"
