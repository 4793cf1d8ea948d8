/* include/bdd_b200.h -- C ABI of libbdd_b200.so, the B200-native (sm_100a) deferred
 * min-marginal-averaging sweep.
 *
 * Drop-in boundary.  The reference has no FFI for this path: callers use the C++ class
 * LPMP::bdd_cuda_parallel_mma<REAL> (include/bdd_solver/bdd_cuda_parallel_mma.h:7-52) and
 * its base LPMP::bdd_cuda_base<REAL> (include/bdd_solver/bdd_cuda_base.h:57-226) by duck
 * typing.  Every entry point below is what a method of those two classes binds to; the
 * header-only shim bdd_b200/csrc/host/bdd_solver/bdd_cuda_parallel_mma.h re-creates the class on top
 * of this ABI, and bdd_b200/solver.py is the ctypes mirror.
 *
 * Conventions
 *  - Plain pointers and sizes only.  `*_host` arguments are host memory, `*_dev` arguments
 *    are device memory on the solver's device; `void*` REAL buffers hold float when the
 *    solver was created with BDDB200_FLOAT and double with BDDB200_DOUBLE.
 *  - Every function returns a bddb200_status; BDDB200_OK == 0.  The reference signals
 *    construction / configuration errors with std::runtime_error
 *    (SURVEY 8b "Error convention"); the C++ shim maps non-zero codes to that exception
 *    with bddb200_last_error() as message.
 *  - Not thread safe per solver (same as the reference, which runs on one host thread and
 *    the default stream, include/cuda_utils.h:111-114).  All work of one solver is issued
 *    on one CUDA stream (bddb200_options.stream, default: a private non-blocking stream).
 *  - "Layer order" of every per-layer vector (length bddb200_nr_layers): BDD-major -- the
 *    layers of BDD 0 in BDD order followed by its terminal layer, then BDD 1, ...  The
 *    terminal layer entries carry primal index INT_MAX, solution 0 and are left untouched
 *    by cost functions, as in the reference (bdd_cuda_base.cu:124, 1198, 1270-1271).
 *    (The reference's own order is its hop-sorted one, bdd_cuda_base.cu:146-188; callers
 *    only rely on all per-layer vectors of one solver sharing one order.)
 *  - delta vectors: 2 * nr_variables REALs, delta[2v] = lo, delta[2v+1] = hi
 *    (delta_lo_hi_, include/bdd_solver/bdd_cuda_base.h:201).
 */
#ifndef BDD_B200_H
#define BDD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bddb200_solver bddb200_solver;

/* Memory layout of BDD::bdd_instruction (include/bdd_collection/bdd_collection.h:14-17):
 * lo / hi are absolute indices into the instruction array; index is the variable, or
 * SIZE_MAX for the top sink and SIZE_MAX-1 for the bot sink (:19-24).  The last two
 * instructions of every BDD are its two sinks, in either order (bdd_collection.cpp:403-428,
 * :1581-1586); a sink is recognised by the index field of the instruction an arc points to. */
typedef struct bddb200_instruction {
    size_t lo;
    size_t hi;
    size_t index;
} bddb200_instruction;

typedef enum bddb200_status {
    BDDB200_OK = 0,
    BDDB200_ERR_INVALID_ARGUMENT = 1,
    BDDB200_ERR_CUDA = 2,
    BDDB200_ERR_NOT_QBDD = 3,       /* reference: assert(is_qbdd), bdd_cuda_base.cu:100-101 */
    BDDB200_ERR_TOO_WIDE = 4,       /* a BDD layer does not fit the shared-memory frontier  */
    BDDB200_ERR_STATE = 5,          /* e.g. backward_mm without a valid forward state (reference: assert, bdd_cuda_parallel_mma.cu:304) */
    BDDB200_ERR_NO_DEVICE = 6,
    BDDB200_ERR_EXCHANGE = 7        /* multi-GPU: a peer did not reach the exchange of a pass within the time limit */
} bddb200_status;

typedef enum bddb200_precision { BDDB200_FLOAT = 0, BDDB200_DOUBLE = 1 } bddb200_precision;

typedef struct bddb200_options {
    int device;                     /* CUDA device ordinal (reference: always 0)                      */
    void* stream;                   /* cudaStream_t to run on; NULL = private non-blocking stream     */
    int deterministic;              /* 1: per-variable sums in fixed BDD order (bit-reproducible, and  *
                                     *    bit-identical to the single-threaded CPU solver in double);  *
                                     * 0: atomic accumulation like bdd_cuda_parallel_mma.cu:358-393     */
    int lanes_per_bdd;              /* 0 = choose from the BDD widths; else force 1,2,4,8,16 or 32     */
    size_t nr_variables;            /* 0 = max variable + 1 (bdd_cuda_base.cu:59-64); larger when this  *
                                     * solver holds one shard of a bigger problem                       */
    const int32_t* nr_bdds_per_var_host; /* NULL = counted from this collection; else global counts of *
                                     * length nr_variables (shard mode, cf. the hybrid solver's         *
                                     * bdd_multi_parallel_mma_base.cu:191-215)                          */
    /* kernel tuning; 0 = automatic */
    int stage_bytes;                /* shared-memory budget of one pipeline stage (chunk of hops)      */
    int n_stages;                   /* pipeline depth per warp, 2..8 (default 3)                       */
    int warps_per_cta;              /* bundles (warps) per CTA for bundles that fit the stage budget   */
    size_t n_shared_vars;           /* shard mode: variables [0, n_shared_vars) occur in other shards too: the BDDs that contain one *
                                     * are bundled together (few bundles then take part in the multi-GPU flag barrier); 0 = none   */
} bddb200_options;

void bddb200_default_options(bddb200_options* opts);
const char* bddb200_last_error(void);
const char* bddb200_version(void);

/* ---- construction: bdd_cuda_parallel_mma(const bdd_collection&[, costs]) --------------
 * (bdd_cuda_parallel_mma.cu:7-27 on top of bdd_cuda_base.cu:31-53).  costs_hi_host may be
 * NULL (n_costs = 0).  opts may be NULL. */
int bddb200_create(const bddb200_instruction* instrs_host, size_t n_instr,
                   const size_t* delimiters_host, size_t n_bdds,
                   const double* costs_hi_host, size_t n_costs,
                   int precision, const bddb200_options* opts, bddb200_solver** out);
void bddb200_destroy(bddb200_solver* s);

/* ---- constraint-sharded construction (multi-GPU) ----------------------------------------
 * One process per GPU calls bddb200_create_shard with the WHOLE collection and its (rank, world): the library cuts the collection
 * into `world` contiguous blocks of BDDs with balanced node counts, renumbers the variables so that those occurring in more than
 * one block come first ([0, n_shared): only that prefix of the per-variable sums crosses NVLink), and builds this rank's solver in
 * shard mode (global nr_bdds_per_var, costs split by the global count -- the hybrid solver's split,
 * bdd_multi_parallel_mma_base.cu:121-124, 191-215).  new_of_old_out (nr_variables entries, may be NULL) receives the renumbering;
 * solver vectors indexed by variable (delta sums, min-marginals) use the NEW numbering.  bddb200_plan_shard is the planning step
 * alone (no GPU needed); counts_new_out = global BDD count per NEW variable index.  Afterwards register the symmetric sum buffers
 * and the exchange (bddb200_set_delta_buffers, bddb200_set_exchange with n_exchange = 2 * n_shared); shared_entries (this shard's
 * layer entries of shared variables) is what decides between the push exchange (mode 4, few) and the pull forms (many). */
typedef struct bddb200_shard_info {
    size_t nr_variables;   /* of the whole problem */
    size_t n_shared;       /* variables that occur in more than one shard */
    size_t shared_entries; /* this shard's layer entries of shared variables */
    size_t first_bdd;      /* this rank's block of BDDs: [first_bdd, first_bdd + n_bdds) */
    size_t n_bdds;
} bddb200_shard_info;
int bddb200_plan_shard(const bddb200_instruction* instrs_host, size_t n_instr, const size_t* delimiters_host, size_t n_bdds,
                       size_t nr_variables_min, int world, int rank, bddb200_shard_info* info,
                       int32_t* new_of_old_out, int32_t* counts_new_out, uint16_t* share_mask_out);
/* Which ranks' shards contain each shared variable (n = n_shared entries, NEW numbering, bit r = rank r; share_mask_out of
 * bddb200_plan_shard): the push exchange sends a variable's differences only there.  bddb200_create_shard sets it itself; a caller
 * that planned the shard elsewhere passes it before bddb200_set_exchange(mode 4) (default: every shared variable goes to all ranks). */
int bddb200_set_push_masks(bddb200_solver* s, const uint16_t* masks_host, size_t n);
int bddb200_create_shard(const bddb200_instruction* instrs_host, size_t n_instr, const size_t* delimiters_host, size_t n_bdds,
                         const double* costs_hi_host, size_t n_costs, int precision, const bddb200_options* opts,
                         int world, int rank, bddb200_shard_info* info, int32_t* new_of_old_out, bddb200_solver** out);
/* copy constructor of the reference class (all its members are thrust::device_vectors): a deep
 * copy with its own stream on the same device */
int bddb200_clone(const bddb200_solver* s, bddb200_solver** out);
/* cereal save / load of the reference class (bdd_cuda_base.cu:1486-1544, bdd_cuda_parallel_mma.cu:475-490; what its pybind module
 * pickles, bdd_cuda_parallel_mma_py.cu:29-38): the whole solver state -- layout, costs, pending sums, validity flags -- as one
 * byte blob.  bddb200_save writes at most `bytes` bytes and reports the blob size (ask bddb200_save_size first); bddb200_load builds
 * a new solver from a blob on `device` (same GPU model as the one it was saved on: the launch plan is part of the blob).  Solvers
 * whose sum buffers live in caller-owned memory (multi-GPU exchange) cannot be saved (BDDB200_ERR_STATE). */
int bddb200_save_size(const bddb200_solver* s, size_t* bytes_out);
int bddb200_save(const bddb200_solver* s, void* buf, size_t bytes, size_t* written_out);
int bddb200_load(const void* buf, size_t bytes, int device, bddb200_solver** out);

/* ---- sizes (bdd_cuda_base.h:101-117) -------------------------------------------------- */
size_t bddb200_nr_variables(const bddb200_solver* s);
size_t bddb200_nr_bdds(const bddb200_solver* s);
size_t bddb200_nr_layers(const bddb200_solver* s);        /* incl. one terminal layer per BDD */
size_t bddb200_nr_bdd_nodes(const bddb200_solver* s);     /* incl. both terminals per BDD      */
size_t bddb200_nr_hops(const bddb200_solver* s);          /* length of the longest BDD         */
int bddb200_precision_of(const bddb200_solver* s);
int bddb200_device_of(const bddb200_solver* s);
int bddb200_nr_bdds_per_var(const bddb200_solver* s, int32_t* out_host);   /* nr_bdds(var), get_num_bdds_per_var */
int bddb200_layer_primal_indices(const bddb200_solver* s, int32_t* out_host); /* get_primal_variable_index: INT_MAX on terminal layers */
int bddb200_layer_bdd_indices(const bddb200_solver* s, int32_t* out_host);    /* get_bdd_index */

/* ---- the hot path ---------------------------------------------------------------------
 * iteration(omega): forward_mm, normalize, backward_mm, normalize
 * (bdd_cuda_parallel_mma.cu:142-153) on the solver's own delta vector. */
int bddb200_iteration(bddb200_solver* s, double omega);
/* n back-to-back iterations captured once in a CUDA graph and replayed. */
int bddb200_iterations(bddb200_solver* s, double omega, size_t n);
/* The two halves of iteration() on the solver's own (rotating, un-normalised) delta sums;
 * a multi-GPU driver all-reduces bddb200_delta_sum_buffer() between them. */
int bddb200_forward_pass(bddb200_solver* s, double omega);
int bddb200_backward_pass(bddb200_solver* s, double omega);
/* forward_mm / backward_mm(omega, delta) (bdd_cuda_parallel_mma.cu:207-257, 301-346):
 * delta_dev is read as the values to add to the arc costs and overwritten with the newly
 * collected, NOT yet normalised, per-variable min-marginal differences. */
int bddb200_forward_mm(bddb200_solver* s, double omega, void* delta_dev);
int bddb200_backward_mm(bddb200_solver* s, double omega, void* delta_dev);
/* normalize_delta (bdd_cuda_parallel_mma.cu:421-430): delta[i] /= nr_bdds(i/2). */
int bddb200_normalize_delta(const bddb200_solver* s, void* delta_dev);
/* the solver's own delta vector after iteration() (normalised), copied out */
int bddb200_get_delta(bddb200_solver* s, void* out_dev_or_host, int out_is_host);
/* lower_bound() (bdd_cuda_base.cu:1243-1251): runs a plain backward pass if needed. */
int bddb200_lower_bound(bddb200_solver* s, double* out_host);
int bddb200_lower_bound_per_bdd(bddb200_solver* s, void* out_dev);   /* :1253-1259 */

/* ---- plain shortest-path runs (bdd_cuda_base.cu:588-612, 670-713) -------------------- */
int bddb200_forward_run(bddb200_solver* s);
int bddb200_backward_run(bddb200_solver* s);
void bddb200_flush_forward_states(bddb200_solver* s);   /* bdd_cuda_base.cu:393-403 */
void bddb200_flush_backward_states(bddb200_solver* s);

/* ---- costs ----------------------------------------------------------------------------
 * update_costs (bdd_cuda_base.cu:476-558): cost[layer] += c[var] / nr_bdds(var). */
int bddb200_update_costs_host(bddb200_solver* s, const double* lo_host, size_t n_lo, const double* hi_host, size_t n_hi);
/* update_costs(const std::vector<REAL>&, const std::vector<REAL>&) (bdd_cuda_base.cu:519-523): host
 * arrays of the solver's REAL type.  Both host variants stage through pinned memory and return
 * without waiting for the device; the caller's arrays may be reused immediately. */
int bddb200_update_costs_host_real(bddb200_solver* s, const void* lo_host, size_t n_lo, const void* hi_host, size_t n_hi);
int bddb200_update_costs_dev(bddb200_solver* s, const void* lo_dev, size_t n_lo, const void* hi_dev, size_t n_hi);
/* One step of a host-driven loop in one call: update_costs(lo, hi) from host vectors (of REAL when src_is_real != 0, else of double),
 * iteration(omega), lower_bound() -- what a perturbation round of bdd_solver.cpp:318-380 / one turn of run_solver
 * (run_solver_util.h:37-49) does with three calls.  REAL vectors are copied straight from the caller's memory (DMA when it is pinned);
 * everything after the upload is one CUDA graph launch.  Returns after the bound has arrived: the caller's arrays are free again. */
int bddb200_step_host(bddb200_solver* s, const void* lo_host, size_t n_lo, const void* hi_host, size_t n_hi, int src_is_real, double omega, double* lb_out);
int bddb200_set_cost(bddb200_solver* s, double c, size_t var);      /* :439-452 */
/* distribute_delta (bdd_cuda_base.cu:1396-1436) */
int bddb200_distribute_delta(bddb200_solver* s);
/* get/set_solver_costs (bdd_cuda_base.cu:1308-1344): three per-layer vectors, layer order */
int bddb200_get_solver_costs(const bddb200_solver* s, void* lo_dev, void* hi_dev, void* mm_diff_dev);
int bddb200_set_solver_costs(bddb200_solver* s, const void* lo_dev, const void* hi_dev, const void* mm_diff_dev);
/* get_primal_objective_vector_host (bdd_cuda_base.cu:1352-1373): sum over layers of hi - lo */
int bddb200_primal_objective_host(bddb200_solver* s, double* out_host);

/* ---- min-marginals (bdd_cuda_base.cu:716-751) -----------------------------------------
 * sorted = 1: entries ordered by (variable, BDD index), the nr_bdds terminal entries last,
 * as min_marginals_cuda(true); sorted = 0: layer order.  All outputs have nr_layers entries
 * and may be NULL. */
int bddb200_min_marginals(bddb200_solver* s, int sorted, int32_t* primal_index_dev, void* mm_lo_dev, void* mm_hi_dev);
/* the same into HOST memory as doubles (what min_marginals() returns to host callers, bdd_cuda_base.cu:753-786, and what
 * bdd_solver::min_marginals hands to Python, bdd_solver.cpp:497-514) */
int bddb200_min_marginals_host(bddb200_solver* s, int sorted, int32_t* primal_index_host, double* mm_lo_host, double* mm_hi_host);

/* ---- L-BFGS support surface (include/bdd_solver/lbfgs.h:22-27) -------------------------- */
int bddb200_bdds_solution(bddb200_solver* s, char* sol_dev);                 /* bdds_solution_vec, bdd_cuda_base.cu:1137-1202 */
int bddb200_net_solver_costs(const bddb200_solver* s, void* out_dev);         /* bdd_cuda_parallel_mma.cu:449-463 */
int bddb200_make_dual_feasible(const bddb200_solver* s, void* inout_dev);     /* bdd_cuda_base.cu:1277-1303 */
int bddb200_gradient_step(bddb200_solver* s, const void* dir_dev, double step); /* bdd_cuda_parallel_mma.h:62-77 */

/* ---- L-BFGS wrapper: lbfgs<bdd_cuda_parallel_mma<REAL>, device_vector<REAL>, REAL, device_vector<char>, true> --------
 * (include/bdd_solver/lbfgs.h:35-110, src/bdd_solver/lbfgs_impl.h; config strings "lbfgs cuda mma" / "cuda lbfgs parallel mma",
 * src/bdd_solver/bdd_solver.cpp:222-265).  bddb200_lbfgs_iteration is lbfgs<>::iteration(): store the iterate, then either an
 * L-BFGS step (two-loop recursion, make_dual_feasible, step-size search by lower bound) followed by one MMA iteration, or a plain
 * MMA iteration while the history fills.  All vectors stay on the device.  Parameters <= 0 select the reference defaults
 * (history 5, step 1e-6, required relative increase 1e-6, decrease 0.8, increase 1.1; lbfgs.h:29-33).  The solver must outlive the
 * wrapper; call bddb200_lbfgs_flush after changing costs from outside (lbfgs<>::update_costs, lbfgs_impl.h:343-348).
 * The reference template is broken at this commit on both back ends, so there is no reference trajectory to pin. */
typedef struct bddb200_lbfgs bddb200_lbfgs;
int bddb200_lbfgs_create(bddb200_solver* s, int history_size, double init_step_size, double req_rel_lb_increase,
                         double step_size_decrease_factor, double step_size_increase_factor, bddb200_lbfgs** out);
void bddb200_lbfgs_destroy(bddb200_lbfgs* l);
int bddb200_lbfgs_iteration(bddb200_lbfgs* l);
int bddb200_lbfgs_flush(bddb200_lbfgs* l);
int bddb200_lbfgs_stats(const bddb200_lbfgs* l, size_t* lbfgs_iterations, size_t* mma_iterations, double* step_size);

/* ---- termination loop and primal rounding (the steps after construction in bdd_solver::solve, bdd_solver.cpp:277-380) --
 * bddb200_run_solver = run_solver (include/run_solver_util.h:10-77): iterate until the iteration limit, the time limit, a relative
 * improvement below `tolerance`, or an improvement below `improvement_slope` x the first iteration's.  `lbfgs` may be NULL.
 * bddb200_rounding_perturb = perturb_primal_costs (src/bdd_solver/incremental_mm_agreement_rounding_cuda.cu:264-335): distribute_delta,
 * min-marginals, per-variable agreement type {0 zero, 1 one, 2 equal, 3 inconsistent}; *solved = 1 and sol_host (nr_variables chars) filled
 * when every variable is decided, else the costs are perturbed (update_costs) and *solved = 0.  counts_out = number of variables per type;
 * types_dev (nr_variables chars, device) may be NULL.
 * bddb200_incremental_mm_agreement_rounding = the whole loop (:338-375): delta grows by delta_growth_rate per round (capped at 1e6),
 * run_solver(num_itr_lb, 1e-7, 1e-4) between rounds. */
int bddb200_run_solver(bddb200_solver* s, bddb200_lbfgs* lbfgs, size_t max_iter, double tolerance, double improvement_slope,
                       double time_limit_s, double* lb_out);
int bddb200_rounding_perturb(bddb200_solver* s, double delta, int round_index, unsigned long long counts_out[4], char* types_dev,
                             char* sol_host, int* solved);
int bddb200_incremental_mm_agreement_rounding(bddb200_solver* s, bddb200_lbfgs* lbfgs, double init_delta, double delta_growth_rate,
                                              int num_itr_lb, int num_rounds, char* sol_host, int* solved, int* rounds_used);

/* ---- stream plumbing / diagnostics ------------------------------------------------------ */
int bddb200_synchronize(bddb200_solver* s);
void* bddb200_stream(bddb200_solver* s);
/* number of kernels of this library launched by this solver so far (bench.py's gpu_launches) */
size_t bddb200_kernel_launches(const bddb200_solver* s);
/* device buffers of the solver's own delta sums, for multi-GPU exchange: *sum_dev receives the
 * buffer holding the UN-normalised sums written by the last pass (2 * nr_variables REALs). */
int bddb200_delta_sum_buffer(bddb200_solver* s, void** sum_dev);

/* The library issues the exchange itself after every forward_pass / backward_pass (and inside iteration / iterations, whose CUDA
 * graph then contains it): register the peer mappings once.  mode 1 = one-shot reads of every peer's sum buffer, 2 = two-shot
 * (slice-wise reduce, then copy), 3 = in-switch reduction through multicast mappings (multimem.ld_reduce / multimem.st; mc_in = multicast
 * address of the symmetric block holding the three sum buffers, mc_out = multicast address of the symmetric result buffer), 0 = off.
 * mode 4 = push: no exchange kernel -- the pass itself adds the min-marginal differences of the first n_exchange / 2 variables to its own
 * sum buffer and, with peer-memory reductions over NVLink through peer_bufs_dev, to the buffers of the other ranks that hold the variable
 * (bddb200_set_push_masks); the bundles that contain such variables run first, signal the peers when done, and wait for the peers' flags
 * in the next pass; all sums are then read from the rank's own buffer (no bddb200_set_delta_input).  Needs bddb200_push_exchange_supported.
 * The epoch of the flag barriers lives on the device, so all ranks must make the same sequence of passes.  Replaces the host-staged
 * exchange of the hybrid solver, bdd_multi_parallel_mma_base.cu:266-318. */
int bddb200_set_exchange(bddb200_solver* s, int world, int rank, const void* const* peer_bufs_dev, uint32_t* const* flags_dev, void* out_dev,
                         void* const* peer_outs_dev, const void* mc_in, void* mc_out, size_t n_exchange, int mode);

/* ---- multi-GPU exchange over peer memory (NVLink / NVSwitch), SURVEY 8e -----------------------------
 * The reference has no multi-GPU code; its only multi-device exchange is the hybrid solver's host-staged
 * copy of the delta vector (bdd_multi_parallel_mma_base.cu:266-318).  Here each rank keeps its three rotating
 * sum buffers in symmetric memory that every peer maps (bddb200_set_delta_buffers; zero-filled, 2 * nr_variables
 * REALs each), numbers the variables that occur in more than one shard first, lets the passes read the exchanged
 * sums of those n_shared variables from a separate buffer (bddb200_set_delta_input) and, after every pass,
 * launches bddb200_delta_exchange on the solver's stream: a one-shot all-reduce that reads all peers' buffer
 * `index` (bddb200_delta_sum_index) directly over NVLink.
 *   peer_bufs_dev : device array of `world` pointers, entry r = rank r's sum-buffer block as mapped here
 *   flags_dev     : device array of `world` pointers to each rank's flag array (>= world uint32, zero-filled)
 *   epoch         : 1, 2, 3, ... one per exchange, the same on every rank
 *   offset_elems  : start of the exchanged buffer inside the block, in REALs
 *   n_exchange    : leading REALs (2 * n_shared) summed over all ranks into out_dev */
int bddb200_delta_sum_index(const bddb200_solver* s, int* index_out);
/* 1 when this solver can run the push exchange (mode 4 of bddb200_set_exchange): atomic (non-deterministic) sums and every BDD in the
 * one-lane-per-BDD class */
int bddb200_push_exchange_supported(const bddb200_solver* s, int* yes_out);
int bddb200_set_delta_buffers(bddb200_solver* s, void* buf0_dev, void* buf1_dev, void* buf2_dev);
int bddb200_set_delta_input(bddb200_solver* s, void* shared_in_dev, size_t n_shared_vars);
int bddb200_delta_exchange(void* stream, int precision, int world, int rank, const void* const* peer_bufs_dev,
                           uint32_t* const* flags_dev, uint32_t epoch, size_t offset_elems, void* out_dev,
                           size_t n_exchange);
/* Two-shot variant for many ranks and long prefixes: every rank sums one slice over all ranks into its own out buffer
 * (symmetric memory too: peer_outs_dev[r] = rank r's out buffer as mapped here; this rank's result is peer_outs_dev[rank]),
 * then copies the other slices from the peers.  Reads 2 (world-1)/world of the prefix per rank instead of (world-1) times it.
 * flags: >= 64 uint32 per rank, zero-filled. */
int bddb200_delta_exchange_two_shot(void* stream, int precision, int world, int rank, const void* const* peer_bufs_dev,
                                    void* const* peer_outs_dev, uint32_t* const* flags_dev, uint32_t epoch,
                                    size_t offset_elems, size_t n_exchange);

/* Diagnostics: run ONE forward (forward != 0) or backward MMA pass and return, for the first
 * max_bundles bundles, 16 clock64() stamps each: [0] warp start, [1] descriptor loaded,
 * [2] first bulk copies issued, [3] first chunk landed, [4+2i] chunk i ready, [5+2i] chunk i
 * computed, [15] SM id.  The pass is a real one (solver state advances). */
int bddb200_trace_pass(bddb200_solver* s, int forward, double omega, unsigned long long* out_host, size_t max_bundles, size_t* n_out);

/* Layout statistics computed on the host only (no GPU needed): fills up to n of
 * {slots, layer entries, bundles, real nodes, max hops, max tile slots, small-class bundles,
 *  layers, chunks, largest small-class stage bytes, largest large-class stage bytes}. */
int bddb200_layout_stats(const bddb200_instruction* instrs_host, size_t n_instr,
                         const size_t* delimiters_host, size_t n_bdds, int lanes_per_bdd,
                         uint64_t* out, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* BDD_B200_H */
