/*
 * onephase_b200.h -- C ABI of libonephase_b200.so, the B200-native (sm_100a)
 * replacement for the per-iteration KKT solve of ohinder/OnePhase.jl.
 *
 * Every entry point replaces a method of the reference's two plugin interfaces
 * (paths relative to the reference's src/):
 *   abstract_linear_system_solver   linear_system_solvers/linear_system_solvers.jl:11,40-46
 *   abstract_KKT_system_solver      kkt_system_solver/kkt_system_solver.jl:10-19
 * The Julia shim that binds them with ccall is in julia/ (see INTEGRATION.md).
 *
 * Conventions
 *   - all pointers are HOST pointers owned by the caller; the library copies in /
 *     out during the call and never retains them;
 *   - sparse matrices are CSC with int64 indices, `index_base` 1 (Julia) or 0;
 *   - every function returns 0 on success or a negative OPB_ERR_* code; the text
 *     is available from opb_last_error(h).  "Not positive definite" is a RESULT
 *     (inertia_ok == 0), never an error (julia.jl:39-45);
 *   - a handle is bound to one CUDA device and one stream; calls on a handle are
 *     synchronous with respect to the host and must not be issued concurrently;
 *   - there is no CPU fallback: numeric calls on a handle created with
 *     device_id < 0 return OPB_ERR_NO_DEVICE.
 */
#ifndef ONEPHASE_B200_H
#define ONEPHASE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct opb_handle opb_handle;

#define OPB_OK 0
#define OPB_ERR_INVALID (-1)    /* bad argument / malformed pattern            */
#define OPB_ERR_STATE (-2)      /* call order violates the `ready` state machine (kkt_system_solver.jl:181-199) */
#define OPB_ERR_CUDA (-3)       /* CUDA runtime error (incl. out of memory)    */
#define OPB_ERR_NO_DEVICE (-4)  /* numeric call on a host-only handle          */
#define OPB_ERR_INTERNAL (-5)

#define OPB_MODE_CHOLESKY 0     /* sym == :definite  (julia.jl:28-46)  */
#define OPB_MODE_LDLT 1         /* sym == :symmetric (julia.jl:47-90)  */

/* status_out of opb_factor_delta_loop (delta_strategy.jl:37-114) */
#define OPB_DELTA_SUCCESS 1     /* :success                                  */
#define OPB_DELTA_FAILURE 0     /* :failure, delta > delta.max -> :MAX_DELTA */
#define OPB_DELTA_MAX_IT (-1)   /* error("max it")                           */

/* initialize! / finalize!  (linear_system_solvers.jl:40-46, kkt_system_solver.jl:21-25).
 * device_id >= 0 selects the CUDA device; device_id < 0 creates a host-only
 * handle (symbolic analysis + introspection only). */
int opb_create(opb_handle** h, int device_id, unsigned flags);
int opb_destroy(opb_handle* h);
const char* opb_last_error(const opb_handle* h);

/* Run all kernels of this handle on an existing cudaStream_t (e.g. the host
 * framework's current stream) instead of the handle's own stream. */
int opb_set_stream(opb_handle* h, void* cuda_stream);

/* Options.  Symbolic (before opb_set_structure): "ordering" (4 auto = fewer flops of 0 and 3
 * [default], 0 level-structure nested dissection + minimum-degree leaves, 1 natural,
 * 3 METIS_NodeND), "shard_split_flops" (sharded instance: update blocks of top fronts with at least this many
 * flops are split over the ranks of their range, default 2e10), "nd_leaf", "nd_balance" (a separator level must leave at least this fraction
 * of the part on either side, default 0.40), "metis_max_n", "relax" (0/1), "relax_small".
 * Numeric (any time): "attempts_per_sync" (delta-loop attempts enqueued per host
 * synchronisation, default 2), "graphs" (0/1: replay the launch sequences from CUDA graphs),
 * "outer_block" (columns of the outer block of the panel updates, 128 * 2^k, default 4096),
 * "lookahead" (blocked panel factorisation of the big fronts: 0 = one stream, 1 = the bulk of each
 * update on a side stream, 2 [default] = deep look-ahead: every update cut into pieces by the step at which
 * its columns are next touched, one prioritised stream per piece class), "chain_priority" (0/1: the
 * latency chain of that factorisation runs on the highest-priority stream, default 1),
 * "barrier_timeout_s" (sharded instance: seconds a rank waits for its peers, default 20), "cb_small_k" (tuning: levels
 * whose fronts have at most this many pivot columns form their update blocks in 64-row tiles, default all; an odd
 * value runs the bulk panel updates in 128-row tiles), "ldlt_scalar" (0/1: LDL' of the big fronts on the first,
 * scalar path instead of the tensor-core tile engine; for A/B runs), "loop_graph" (0/1: the delta loop as one
 * CUDA-graph WHILE node, default 1), "solve_overlap" (0/1). */
int opb_set_option(opb_handle* h, const char* key, double value);
/* Optional fill-reducing permutation supplied by the caller (0-based, perm[new] = old). */
int opb_set_permutation(opb_handle* h, int64_t n, const int64_t* perm);

/* Sparsity of the iterate's cached matrices: J (m x n CSC, Class_iterate.jl:334-355)
 * and H (n x n CSC, lower triangular incl. diagonal, eval.jl:132-134).  Runs the
 * symbolic analysis once per pattern (cached by pattern hash per process): pattern
 * of tril(J'DJ + H) with full diagonal, gather map, ordering, supernodes, level
 * sets.  CHOLMOD redoes this on every ls_factor! (julia.jl:34, recycle=false). */
int opb_set_structure(opb_handle* h, int64_t n, int64_t m,
                      const int64_t* J_colptr, const int64_t* J_rowval,
                      const int64_t* H_colptr, const int64_t* H_rowval, int index_base);

/* form_system!  (schur.jl:47-62): Q = J' diag(y./s) J + H, schur_diag = diag(Q);
 * also returns diag_min(kkt_solver) (kkt_system_solver.jl:291-294).
 * schur_diag_out (n) and diag_min_out may be NULL. */
int opb_form(opb_handle* h, const double* J_nzval, const double* H_nzval,
             const double* y, const double* s, double* schur_diag_out, double* diag_min_out);

/* Lower triangle of Q as assembled on the device: pattern (0-based) and values,
 * for is_diag_dom (delta_strategy.jl:95) and for tests. */
int opb_get_M_pattern(opb_handle* h, int64_t* colptr_out, int64_t* rowval_out);
int opb_get_M_values(opb_handle* h, double* nzval_out);

/* ipopt_strategy!  (delta_strategy.jl:37-114) with update_delta_vecs! + ls_factor!
 * (schur.jl:64-87, julia.jl:28-46) inside: the whole delta loop runs on the
 * device, the PD check never returns to the host between attempts. */
int opb_factor_delta_loop(opb_handle* h, double delta_prev, double delta_zero, double delta_min,
                          double delta_max, double delta_start, double inc, double dec, int max_it,
                          double* delta_out, int* num_fac_out, int* status_out);

/* factor!(kkt_solver, delta, timer)  (kkt_system_solver.jl:98-107,190-204): one
 * attempt with a given shift (one_phase.jl:241, test/kkt_system_solvers.jl:78). */
int opb_factor(opb_handle* h, double delta, int* inertia_ok);

/* compute_direction_implementation!  (schur.jl:89-128) with solver_schur_rhs's
 * iterative refinement (schur.jl:131-182) and update_kkt_error!
 * (kkt_system_solver.jl:67-96).  kkt_err_out = [error_D, error_P, error_mu,
 * overall, rhs_norm, ratio] (Class_kkt_error; ratio is the log's "N err"). */
int opb_direction(opb_handle* h, const double* dual_r, const double* primal_r, const double* comp_r,
                  int n_refine, double* dx_out, double* dy_out, double* ds_out, double* kkt_err_out);

/* ls_factor!(solver, Q, n, m, timer)  (julia.jl:21-97) for an arbitrary
 * SparseMatrixCSC: only entries with row >= col are read.  mode CHOLESKY:
 * inertia_ok = 1 iff PD; mode LDLT: inertia_ok = inertia_status(pos, neg, zero,
 * n_pos_expected, m_neg_expected) (linear_system_solvers.jl:48-91). */
int opb_ls_factor_csc(opb_handle* h, int64_t dim, const int64_t* colptr, const int64_t* rowval,
                      const double* nzval, int index_base, int mode,
                      int64_t n_pos_expected, int64_t m_neg_expected, int* inertia_ok);
/* ls_solve! / ls_solve  (julia.jl:99-113): sol = F \ rhs with the last factor. */
int opb_ls_solve(opb_handle* h, const double* rhs, double* sol);

/* eval_diag_J_T_J(iter, diag_vals)  (utils/eval.jl:89-100): out[i] = sum_j J[j,i]^2 * diag_vals[j]
 * for an m x n CSC matrix, in the reference's operation order.  compute_schur_diag
 * (kkt_system_solver.jl:296-300) = diag(H) + this with diag_vals = y ./ s; used by the symmetric
 * KKT solver (symmetric.jl:49).  Needs no opb_set_structure. */
int opb_eval_diag_JtDJ(opb_handle* h, int64_t n, int64_t m, const int64_t* J_colptr, const int64_t* J_rowval,
                       const double* J_nzval, int index_base, const double* diag_vals, double* out);

/* --- device-resident variants used by bench.py's `value` leg: inputs are uploaded
 *     once, the timed region launches kernels only. --- */
int opb_upload_values(opb_handle* h, const double* J_nzval, const double* H_nzval,
                      const double* y, const double* s);
int opb_upload_rhs(opb_handle* h, const double* dual_r, const double* primal_r, const double* comp_r);
int opb_form_resident(opb_handle* h);
int opb_delta_loop_resident(opb_handle* h, double delta_prev, double delta_zero, double delta_min,
                            double delta_max, double delta_start, double inc, double dec, int max_it);
int opb_direction_resident(opb_handle* h, int n_refine);
int opb_solve_resident(opb_handle* h, int nsolves);   /* triangular solves only, on the residual vector */
/* --- the iterate's (J, H, y, s) stay resident between the calls of one outer iteration; the
 *     right-hand side is built and the step bounds are reduced on the device, so per direction only the NLP's
 *     gradient / constraint values go up and a few scalars come back. ---
 * System_rhs(iter, reduct_factors)  (kkt_system_solver/system_rhs.jl:57-73 with eval_grad_lag / eval_grad_r,
 * utils/eval.jl:59-63,136-142) from the resident (J, y, s) and the caller's grad (n) and cons (m):
 *   dual_r = -(grad - J'y + (mu*eta_mu) * (a_norm_penalty * J'1)) * (1 - eta_D)
 *   primal_r = -(cons - s) * (1 - eta_P),   comp_r = mu*eta_mu - s .* y
 * The three vectors become the resident rhs of the next opb_direction_resident; the *_out pointers may be
 * NULL (nothing is copied back then and the call does not synchronise). */
int opb_system_rhs(opb_handle* h, const double* grad, const double* cons, double mu, double a_norm_penalty,
                   double eta_P, double eta_D, double eta_mu, double* dual_r_out, double* primal_r_out,
                   double* comp_r_out);
/* Fraction-to-the-boundary scalars of the resident direction (line_search/frac_boundary.jl:3-35):
 * out4 = [norm(dx,Inf), norm(dy,Inf), norm(ds,Inf), simple_max_step(s, ds, lb_s)] with
 * lb_s = frac_bd * min.(s, norm(dx,Inf) * norm(dx,Inf)^predict_exp). */
int opb_step_bounds(opb_handle* h, double frac_bd, double predict_exp, double* out4);
/* The resident direction of the last opb_direction_resident and its Class_kkt_error (any pointer may be NULL). */
int opb_get_direction(opb_handle* h, double* dx_out, double* dy_out, double* ds_out, double* kkt_err_out);

/* One factorisation attempt at shift `delta` (like opb_factor) with CUDA events around every
 * launch of the two FP64 tensor-pipe kernels: time of the whole attempt, summed time of the
 * update-block kernel and of the panel-update kernel, and the algorithmic flops those kernels
 * serve (for bench.py's per-kernel roofline).  Plain launches, no look-ahead stream. */
int opb_profile_factor(opb_handle* h, double delta, double* total_ms, double* cb_ms, double* update_ms,
                       double* cb_flops, double* update_flops, int* inertia_ok);
/* One factorisation attempt at shift `delta` with the look-ahead streams as configured (plain launches, no
 * graph) and event marks on the main stream at the phase boundaries of every elimination-tree level:
 * out[3*l + 0] = ms before the big panels of level l (small fronts, medium panels, extend-add), out[3*l + 1] =
 * blocked panel factorisation of the big fronts, out[3*l + 2] = update blocks; out[3*nlevels] = front fill
 * before the first level, out[3*nlevels + 1] = pivot-block inverses after the last. */
int opb_profile_levels(opb_handle* h, double delta, double* out, int cap, int* nlevels_out, double* total_ms);
/* blocks until the stream is idle and reads the controller state */
int opb_sync_state(opb_handle* h, double* delta_out, int* num_fac_out, int* status_out, double* kkt_err_out);

/* --- one instance sharded over the GPUs of a box (no counterpart in the
 * reference, which is single-threaded CPU code: julia.jl:21-113 factorises and solves on one
 * core).  One process (or at least one handle) per GPU, all of them making the SAME sequence
 * of calls with the SAME data (SPMD).  Independent subtrees of the elimination tree are mapped
 * to ranks by factorisation flops; the top separators are owned by one rank each and pull
 * their children's update blocks from the owners' HBM over NVLink (peer-mapped memory, no
 * collective library); every rank ends up with the full direction.  Cholesky mode only.
 *   opb_shard_init    before opb_set_structure: this handle is rank `rank` of `world` (<= 8)
 *   opb_shard_export  after every opb_set_structure: OPB_SHARD_BLOB_BYTES bytes describing this
 *                     rank's peer-visible buffers (CUDA IPC handles); exchange them between the
 *                     ranks by any means (the Python host uses torch.distributed.all_gather)
 *   opb_shard_attach  map the buffers of rank `peer` from its blob
 * The update blocks of the top separators are formed by all ranks of the separator's range: the owner
 * factorises the panel, the other ranks pull it over NVLink and store their share of the tiles into the owner's
 * update-block arena.
 * Extra info keys: shard_rank, shard_world, shard_load (flops owned by this rank), shard_split (split fronts),
 * shard_top_flops, shard_barriers; symbolic arrays: owner, top. */
#define OPB_SHARD_BLOB_BYTES 384
int opb_shard_init(opb_handle* h, int rank, int world);
int opb_shard_export(opb_handle* h, unsigned char* blob);
int opb_shard_attach(opb_handle* h, int peer, const unsigned char* blob);

/* --- introspection --- */
/* keys: n, m, nnzJ, nnzH, nnzM, npairs, nnzL, nnzL_true, flops, nsuper, nlevels,
 *       max_front, cb_total, n_tiny, n_small, n_big, symbolic_cached,
 *       device_bytes (device memory held by the library in this process, all handles and cached structures);
 *       t_<phase>: host seconds of the one-off analysis of this structure -- pattern, analyze (= order_own |
 *       order_candidates + order_compare, etree_counts, supernodes, row_structures, storage, rel_gather,
 *       tile_cuts, amap), shard_map, plan, upload (0 for a phase that did not run).
 * The analysis uses up to 16 host threads (environment OPB_HOST_THREADS overrides the core count); its result
 * does not depend on the number of threads. */
int opb_get_info(opb_handle* h, const char* key, double* out);
/* symbolic arrays for tests: perm, sfirst, sparent, rowptr, rowidx, rel, Loff, CBoff,
 * level, amap, dpos, Mp, Mi, pair_ptr, pairA, pairB, hmap.  Values are widened to
 * int64.  Returns the element count (or a negative error); copies min(count, cap). */
int64_t opb_get_symbolic(opb_handle* h, const char* name, int64_t* out, int64_t cap);
/* factor values of supernode panels as stored on the device (tests) */
int opb_get_L_values(opb_handle* h, double* out, int64_t cap);
/* kernels launched by the library since process start */
long long opb_launch_count(void);
/* Symbolic analyses are cached per process (up to 8 patterns, with their device-side maps) so that the solver
 * objects of one solve share one analysis.  opb_cache_clear drops the cache; structures still bound to a live
 * handle stay valid until that handle is destroyed or given another structure.  Returns the entries dropped. */
int opb_cache_clear(void);
const char* opb_version(void);

#ifdef __cplusplus
}
#endif
#endif
