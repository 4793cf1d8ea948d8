/* include/bdd_b200_collection.h -- C ABI of the host-side BDD collection in libbdd_b200.so: direct generators for the common
 * constraint shapes, the structural operations of the solver's front end, and long-BDD splitting.
 *
 * What each entry point replaces is a method of BDD::bdd_collection (include/bdd_collection/bdd_collection.h,
 * src/bdd_collection/bdd_collection.cpp of the reference); the instruction arrays it produces are identical to the reference's
 * (tests/test_collection.py against the reference's own object code).  Host code only -- none of these calls needs a GPU.
 * The C++ class behind it is bddb200_host::bdd_collection (bdd_b200/csrc/host/bdd_collection.hpp, header-only, usable without the
 * library); bdd_b200/collection.py is the ctypes mirror.
 *
 * Conventions as in bdd_b200.h: every function returns a bddb200_status, the message of the last failure on this thread is
 * bddb200_last_error(); instruction arrays are bddb200_instruction {lo, hi, index} with absolute child indices; a BDD number is
 * the position of a BDD in the collection and shifts down when BDDs in front of it are removed.
 */
#ifndef BDD_B200_COLLECTION_H
#define BDD_B200_COLLECTION_H

#include "bdd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bddb200_collection bddb200_collection;

/* an empty collection (instrs == NULL, n_instr == 0, n_bdds == 0) or a copy of the given arrays (delims: n_bdds + 1 entries) */
int bddb200_collection_create(const bddb200_instruction* instrs, size_t n_instr, const size_t* delims, size_t n_bdds, bddb200_collection** out);
int bddb200_collection_destroy(bddb200_collection* c);
int bddb200_collection_nr_bdds(const bddb200_collection* c, size_t* out);                         /* bdd_collection::nr_bdds                    */
int bddb200_collection_nr_instructions(const bddb200_collection* c, size_t* out);                 /* all BDDs, sinks included                   */
int bddb200_collection_export(const bddb200_collection* c, bddb200_instruction* instrs_out, size_t* delims_out);   /* nr_instructions, nr_bdds + 1 */

/* generators: the new BDD is over variables 0 .. n-1 and is appended; *bdd_nr_out is its number */
int bddb200_collection_simplex_constraint(bddb200_collection* c, size_t n, size_t* bdd_nr_out);              /* bdd_collection.cpp:2039 */
int bddb200_collection_not_all_false_constraint(bddb200_collection* c, size_t n, size_t* bdd_nr_out);        /* :2105 (reduced, not quasi-reduced) */
int bddb200_collection_all_equal_constraint(bddb200_collection* c, size_t n, size_t* bdd_nr_out);            /* :2136 (reduced, not quasi-reduced) */
int bddb200_collection_cardinality_constraint(bddb200_collection* c, size_t n, size_t k, size_t* bdd_nr_out);/* :2187 */

/* one linear constraint sum_k coefficients[k] * x[variables[k]] {0 '<=', 1 '>=', 2 '='} rhs over 0/1 variables (ascending), converted
 * directly to its quasi-reduced BDD: what bdd_preprocessor.cpp:175-228 reaches through lineq_bdd -> bdd_mgr -> add_bdd -> reorder ->
 * make_qbdd -> rebase, up to the node order inside a layer.  *bdd_nr_out = SIZE_MAX when the constraint is always satisfied (no BDD). */
int bddb200_collection_add_linear_constraint(bddb200_collection* c, const long long* coefficients, const size_t* variables, size_t n,
                                             int relation, long long rhs, size_t* bdd_nr_out);

int bddb200_collection_rebase(bddb200_collection* c, size_t bdd_nr, const size_t* vars, size_t n_vars);     /* variable i -> vars[i], header :311 */
int bddb200_collection_negate(bddb200_collection* c, size_t bdd_nr);                                         /* :2023 */
int bddb200_collection_invert(bddb200_collection* c, size_t bdd_nr, size_t var);                             /* :2029 */

int bddb200_collection_variables(const bddb200_collection* c, size_t bdd_nr, size_t* vars_out, size_t capacity, size_t* n_out);   /* :1201; vars_out may be NULL to ask for n */
int bddb200_collection_is_qbdd(const bddb200_collection* c, size_t bdd_nr, int* out);                        /* :500 */
int bddb200_collection_is_reordered(const bddb200_collection* c, size_t bdd_nr, int* out);                   /* :1510 */
int bddb200_collection_evaluate(const bddb200_collection* c, size_t bdd_nr, const char* labeling, size_t n, int* out);
int bddb200_collection_reorder(bddb200_collection* c, size_t bdd_nr);                                        /* :1429 */
int bddb200_collection_make_qbdd(bddb200_collection* c, size_t bdd_nr, size_t* bdd_nr_out);                  /* :1670, appended */
int bddb200_collection_bdd_and(bddb200_collection* c, const size_t* bdd_nrs, size_t n, size_t* bdd_nr_out);  /* :31-315, header :494; appended */
int bddb200_collection_remove(bddb200_collection* c, const size_t* bdd_nrs, size_t n);                       /* ascending numbers, header :371 */

/* bddb200_create (bdd_b200.h) on the collection's own arrays: the constructor bdd_cuda_parallel_mma(bdd_col, costs)
 * (src/bdd_solver/bdd_cuda_parallel_mma.cu:7-27) without an export in between; costs beyond n_costs are zero (auxiliary variables) */
int bddb200_create_from_collection(const bddb200_collection* c, const double* costs_hi, size_t n_costs, int precision,
                                   const bddb200_options* opts, bddb200_solver** out);

/* bdd_collection::write_bdd_lp (header :731-830; "export bdd lp" of the driver, bdd_solver.cpp:400-410): the linear programme of the
 * relaxation the dual solvers work on -- arc-flow variables per BDD, linked through the shared variables x_<var> -- as an .lp file */
int bddb200_collection_write_bdd_lp(const bddb200_collection* c, const double* costs, size_t n_costs, const char* path);

/* bdd_collection::split_qbdd (:507-949): chunks (and, with_implication_bdd != 0, the implication BDD when it is not trivial) are
 * appended; *nr_new_out is how many BDDs were appended (0: the BDD is short enough and stays), *next_aux_out the next free
 * auxiliary variable.  BDDB200_ERR_INVALID_ARGUMENT when a cut would land in front of a layer of width 1 (reference: assert, :598). */
int bddb200_collection_split_qbdd(bddb200_collection* c, size_t bdd_nr, size_t chunk_size, size_t aux_var_start, int with_implication_bdd,
                                  size_t* nr_new_out, size_t* next_aux_out);
/* the preprocessor's loop over all BDDs with a forced split length (bdd_preprocessor.cpp:372-415): long BDDs are replaced by their
 * chunks; *n_split_out = number of BDDs that were cut, *nr_variables_out = variables including the auxiliary ones */
int bddb200_collection_split_long_bdds(bddb200_collection* c, size_t split_length, size_t nr_variables, int with_implication_bdd,
                                       size_t* n_split_out, size_t* nr_variables_out);

/* ---- ILP input: the .lp reader in front of the collection (src/ILP/ILP_parser.cpp:25-160, ILP_input; bdd_solver::read_ILP,
 * src/bdd_solver/bdd_solver.cpp:44-66).  Minimisation over 0/1 variables, linear constraints with integer coefficients. ---- */
typedef struct bddb200_ilp bddb200_ilp;

/* file_or_text: the name of a readable .lp file, else the LP text itself (as "input" of the JSON configuration) */
int bddb200_ilp_read(const char* file_or_text, bddb200_ilp** out);
int bddb200_ilp_destroy(bddb200_ilp* ilp);
int bddb200_ilp_nr_variables(const bddb200_ilp* ilp, size_t* out);
int bddb200_ilp_nr_constraints(const bddb200_ilp* ilp, size_t* out);
int bddb200_ilp_objective(const bddb200_ilp* ilp, double* coefficients_out, double* constant_out);   /* nr_variables doubles; either may be NULL */
int bddb200_ilp_variable_name(const bddb200_ilp* ilp, size_t var, const char** name_out);            /* valid until the ILP is destroyed */
/* constraint c: *n_out terms; with non-NULL outputs (capacity >= *n_out of a first call) its variables, coefficients, relation
 * (0 '<=', 1 '>=', 2 '=') and right-hand side */
int bddb200_ilp_constraint(const bddb200_ilp* ilp, size_t c, size_t* n_out, size_t* variables_out, long long* coefficients_out, size_t capacity,
                           int* relation_out, long long* rhs_out);
/* one quasi-reduced BDD per constraint that is not always satisfied, in constraint order (bdd_preprocessor::add_ilp,
 * bdd_preprocessor.cpp:123-228); an infeasible constraint is an error */
int bddb200_ilp_to_bdds(const bddb200_ilp* ilp, bddb200_collection** out);

#ifdef __cplusplus
}
#endif
#endif
