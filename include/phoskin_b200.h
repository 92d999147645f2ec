/*
 * phoskin_b200.h — C ABI of the B200-native ensemble ODE engine for PhosKinTime's hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch / numpy types),
 * and replaces a *loop of single calls* of one reference function (reference = Python, so the
 * binding a maintainer adds is a ctypes stub — see INTEGRATION.md):
 *
 *   pk_local_solve_batch   <- models/distmod.py:93-134, models/succmod.py:114-152,
 *                             models/randmod.py:249-305   solve_ode(params, init_cond, num_psites, t)
 *                             selected by models/__init__.py:6-12 (ODE_MODEL plugin point)
 *   loss fields of the job <- config/config.py:176-226 score_fit; paramest/normest.py:403-423
 *                             (regularised curve_fit residual, sigma-weighted)
 *   Y fields of the job    <- sensitivity/analysis.py:89-176 _compute_Y (5 metrics)
 *   pk_morris_ee           <- SALib.analyze.morris.analyze as called at sensitivity/analysis.py:264
 *   pk_allgather_f64       <- the per-sample result gather the reference does by pickling futures
 *                             (sensitivity/analysis.py:241-259), here one NCCL all-gather.
 *
 * Conventions
 *   - return value: 0 = ok, <0 = API misuse or CUDA error (message via pk_last_error()).
 *     A failing *system* never fails the call: its int32 status is 1 (max steps), 2 (step size
 *     underflow) or 3 (non-finite), and its outputs are NaN (the reference only warns and returns
 *     garbage: models/distmod.py:112; consumers treat non-finite as a bad sample).
 *   - all arrays are C-order doubles unless noted; `memspace` says where *all* data pointers of a
 *     job live (host: the library stages through its own device workspace; device: used in place).
 *   - the library never frees caller memory; a handle owns one stream + workspaces; one handle
 *     per GPU; not fork-safe.
 */
#ifndef PHOSKIN_B200_H
#define PHOSKIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_ABI_VERSION 1

typedef struct pk_handle_s* pk_handle_t;

enum pk_model { PK_DISTMOD = 0, PK_SUCCMOD = 1, PK_RANDMOD = 2 };
enum pk_memspace { PK_HOST = 0, PK_DEVICE = 1 };
/* sensitivity/analysis.py:114-176 */
enum pk_y_metric { PK_Y_NONE = -1, PK_Y_TOTAL_SIGNAL = 0, PK_Y_MEAN_ACTIVITY = 1, PK_Y_VARIANCE = 2,
                   PK_Y_DYNAMICS = 3, PK_Y_L2_NORM = 4 };
/* integrator coefficient set: ROS6L = 7-solve order-6(5) Rosenbrock for linear systems (default of the
 * thread-per-system kernels: dist/succ up to 8 sites; selectable on the dense kernel), ROS5L = 6-solve order-5(4)
 * (default of the dense kernel), RODAS4 = Hairer-Wanner order-4(3); see DESIGN.md section 2 */
enum pk_method { PK_METHOD_DEFAULT = 0, PK_METHOD_RODAS4 = 1, PK_METHOD_ROS5L = 2, PK_METHOD_ROS6L = 3 };
enum pk_status { PK_OK = 0, PK_MAX_STEPS = 1, PK_STEP_UNDERFLOW = 2, PK_NON_FINITE = 3 };

/* One batched call of solve_ode(params[b], init_cond, num_psites, t) for b in [0,B). */
typedef struct pk_local_job {
    int32_t model;          /* pk_model                                                          */
    int32_t n_sites;        /* num_psites                                                        */
    int64_t B;              /* number of independent systems                                     */
    int32_t T;              /* number of output times (t[0] is the initial time)                 */
    int32_t memspace;       /* pk_memspace of every pointer below                                */
    const double* params;   /* [B,P]  P = 4+2ns (dist/succ) or 4+ns+2^ns-1 (rand), reference order */
    const double* y0;       /* [n] if y0_stride==0 else [B,n] with row stride y0_stride doubles  */
    int64_t y0_stride;
    const double* t;        /* [T] strictly increasing                                           */
    double rtol, atol;      /* <=0 -> defaults of the method: ROS6L 2e-5 / 2e-11, ROS5L and RODAS4 2e-6 / 2e-9 */
    int32_t max_steps;      /* per system, <=0 -> 100000                                         */
    int32_t normalize;      /* NORMALIZE_MODEL_OUTPUT (models/distmod.py:115-122)                */
    int32_t log_params;     /* 1: params hold log-values, model uses exp(params) (normest.py:54) */
    int32_t y_metric;       /* pk_y_metric, PK_Y_NONE to skip                                    */
    int32_t method;         /* pk_method                                                         */
    int32_t reserved0;
    /* outputs, each may be NULL */
    double* out_sol;        /* [B,T,n]  clipped at 0                                             */
    double* out_flat;       /* [B,L]    L = (T-5)+T+ns*T  (needs T>5)                            */
    double* out_Y;          /* [B]      Morris scalar                                            */
    double* out_ssr;        /* [B]      sum(((model-target)/sigma)^2) incl. lam/P*params^2 terms */
    double* out_score;      /* [B]      score_fit(params, target, flat)                          */
    int32_t* out_status;    /* [B]                                                               */
    int32_t* out_nsteps;    /* [B]      accepted steps                                           */
    int32_t* out_nrej;      /* [B]      rejected steps                                           */
    /* fused loss inputs (needed iff out_ssr or out_score) */
    const double* target;   /* [G,L]                                                             */
    const double* sigma;    /* NULL (ones) or [G,sigma_len], sigma_len in {L, L+P}               */
    const int32_t* group;   /* NULL (all group 0) or [B] values in [0,G)                         */
    int32_t n_groups;       /* G >= 1                                                            */
    int32_t sigma_len;
    double lam;             /* regularisation lambda (normest.py:56, 421)                        */
    double score_w[5];      /* alpha(rmse) beta(mae) gamma(var) delta(mse) mu(l2); config.toml:212-217 */
    const double* lam_group;/* NULL (lam for every system) or [G]: lambda of every group — the lambda scan of
                               normest.py:22-165 (10 lambdas x weight options) becomes groups of ONE call */
} pk_local_job;

int pk_abi_version(void);
const char* pk_last_error(void);

int pk_create(int device, pk_handle_t* out);
int pk_destroy(pk_handle_t h);
int pk_device_count(int* out);
/* multiprocessor count / SM clock (kHz) / name of the handle's device */
int pk_device_info(pk_handle_t h, int* sm_count, int* clock_khz, char* name, int name_len);

/* sizes implied by (model, n_sites, T): number of states, parameters, flat length */
int pk_local_dims(int model, int n_sites, int T, int* n_states, int* n_params, int* flat_len);

void pk_local_job_init(pk_local_job* job);        /* zero + defaults (score_w = 1, y_metric none)   */
int pk_sizeof_local_job(void);                    /* sizeof(pk_local_job), for binding self-checks  */
int pk_local_solve_batch(pk_handle_t h, const pk_local_job* job);
/* number of kernel launches the last pk_* call on this handle issued, and its device time (ms,
 * CUDA events on the handle's stream around the kernels only, copies excluded) */
int pk_last_launch_info(pk_handle_t h, int* n_launches, float* kernel_ms);

/* Region timing on the handle's stream (the stream every pk_* kernel and copy is issued on):
 * pk_region_begin records a CUDA event, pk_region_end records a second one, synchronises and
 * returns the device time between them in ms. */
int pk_region_begin(pk_handle_t h);
int pk_region_end(pk_handle_t h, float* ms);

/* Morris elementary effects on device data already gathered: X[N*(D+1),D], Y[N*(D+1)].
 * scaled!=0 -> sigma-scaled EE (analysis.py:264 `scaled=True`). Outputs [D] each (may be NULL). */
int pk_morris_ee(pk_handle_t h, int memspace, const double* X, const double* Y, int64_t N, int32_t D,
                 int32_t num_levels, int32_t scaled, double* out_mu, double* out_mu_star,
                 double* out_sigma, double* out_ee /* [N,D] or NULL */);

/* pinned host buffers for fast staging */
int pk_host_alloc(void** ptr, int64_t bytes);
int pk_host_free(void* ptr);

/* FP64 FMA peak of the handle's device, measured with a register-resident DFMA kernel (TFLOP/s). */
int pk_measure_fp64_peak(pk_handle_t h, double* tflops, float* ms);

/* NCCL all-gather of per-sample doubles (device pointers). `comm` is an ncclComm_t created by the
 * caller (or NULL with world==1 -> plain copy). */
/* Coefficients MU[6], EPS[6] of the ROS5L(gamma) member: the dense kernel derives them at run time
 * for steps that reuse an inverse computed for a larger step; exported for verification (host arithmetic). */
int pk_ros5l_coeffs(double gamma, double* mu6, double* eps6);
int pk_ros6l_coeffs(double gamma, double* mu7, double* eps7);   /* the seven-solve family (ROS6L) */

/* The solve fused with the path's one collective: the batch is integrated in `chunks` pieces (<=0 -> 4) and the
 * all-gather of piece c's per-sample output (which: 0 out_score, 1 out_ssr, 2 out_Y; must be requested in the job)
 * runs on a second, high-priority stream while piece c+1 integrates — only the last piece's gather is exposed.
 * recv_dev [world*B] is filled rank-major (the layout pk_allgather_f64 produces).  Device buffers only. */
int pk_local_solve_allgather(pk_handle_t h, const pk_local_job* job, int32_t which, int32_t chunks, double* recv_dev);

/* Precondition of pk_local_solve_allgather: EVERY rank passes the same job->B (the piece boundaries and the send
 * counts of the collectives are derived from it) and recv_dev holds world*B doubles; ranks with fewer systems pad. */

/* The same collective WITHOUT a collective call — the replacement for the reference's pickled futures
 * (sensitivity/analysis.py:241-259) when all ranks sit on one NVSwitch: every rank owns a symmetric buffer of
 * [world][slots_per_rank] doubles, maps every peer's buffer (CUDA IPC; the 64-byte handles travel through the caller's
 * process group), and the solve kernel stores each finished system's scalar straight into all `world` buffers over
 * NVLink as the system completes.  No staging, no exposed tail; the call ends with an 8-byte rendezvous.  Ranks may
 * pass different job->B (<= slots_per_rank): rank r's values land at [r*slots_per_rank, r*slots_per_rank + B). */
int pk_sym_alloc(pk_handle_t h, int64_t slots_per_rank, char* handle_out64);
int pk_sym_open(pk_handle_t h, const char* handles_world_x64);            /* rank-major, 64 bytes each */
int pk_sym_buffer(pk_handle_t h, double** local_dev, int64_t* slots_per_rank);
int pk_sym_free(pk_handle_t h);
int pk_local_solve_gather_p2p(pk_handle_t h, const pk_local_job* job, int32_t which);

int pk_nccl_unique_id(char* out128);
int pk_nccl_init(pk_handle_t h, const char* id128, int world, int rank);
int pk_allgather_f64(pk_handle_t h, const double* send_dev, int64_t count, double* recv_dev);


/* ------------------------------------------------------------------------------------------------
 * Batched bounded non-linear least squares over the local models.
 *
 *   pk_local_nlls_batch  <- a loop of scipy.optimize.curve_fit(model_func, time_points, target_fit, p0=p0_try,
 *                           bounds=free_bounds, sigma=sigma, x_scale='jac') calls: the multistart loop
 *                           paramest/normest.py:274-316 (one call per start), the lambda scan :79-89 and the
 *                           sequential / bootstrap fits :494-509 — with the residual model of normest.py:403-423
 *                              r(theta) = ([flat(solve_ode(theta or exp theta)) | lam/P*theta^2] - [target | 0]) / sigma
 *                           Every problem (one start of one protein) is a row of theta[B,P]; group[b] selects its
 *                           target / sigma row.  On return theta holds the minimiser, out_cost = 0.5*sum(r^2) there and
 *                           out_score = score_fit(theta, target, flat) (config/config.py:176-226), the quantity
 *                           normest.py:293-306 ranks the starts by.
 * The optimiser is a projected Levenberg-Marquardt with MINPACK column scaling (SciPy's TRF is policy, outside the
 * parity contract); Jacobians are forward differences, B*(P+1) solves in ONE launch per iteration.
 * out_status: 1 gtol, 2 ftol, 3 xtol, 4 max_iter reached, -1 numerical failure (integrator failed / non-finite).
 * -----------------------------------------------------------------------------------------------*/
typedef struct pk_nlls_job {
    int32_t model;          /* pk_model                                                          */
    int32_t n_sites;
    int64_t B;              /* problems                                                          */
    int32_t T;
    int32_t memspace;       /* of theta, y0, target, sigma, group, out_*; t, lb, ub are HOST     */
    double* theta;          /* [B,P] in: start points (clipped into the box), out: minimisers    */
    const double* y0;       /* [n] (y0_stride 0) or [B, y0_stride]                               */
    int64_t y0_stride;
    const double* t;        /* [T] HOST                                                          */
    const double* lb;       /* [P] HOST, finite                                                  */
    const double* ub;       /* [P] HOST, finite                                                  */
    const double* target;   /* [G,L]                                                             */
    const double* sigma;    /* NULL or [G,sigma_len], sigma_len in {L, L+P}                      */
    const int32_t* group;   /* NULL or [B]                                                       */
    int32_t n_groups;
    int32_t sigma_len;
    double lam;             /* regularisation lambda (normest.py:56, 421)                        */
    int32_t log_params;     /* 1: model uses exp(theta) (randmod, normest.py:54)                 */
    int32_t max_iter;       /* Jacobian evaluations per problem                                  */
    double ftol, xtol, gtol;/* curve_fit defaults 1e-8                                           */
    double fd_rel;          /* forward-difference step relative to max(1,|theta_j|), default 1e-4 */
    double mu0;             /* initial damping (<=0 -> 1e-3)                                     */
    double rtol, atol;      /* integrator tolerances, defaults 2e-7 / 2e-10                      */
    int32_t max_steps;
    int32_t method;         /* pk_method                                                         */
    double score_w[5];
    double* out_cost;       /* [B]                                                               */
    double* out_score;      /* [B]                                                               */
    int32_t* out_status;    /* [B]                                                               */
    int32_t* out_iters;     /* [B]                                                               */
    int32_t* out_nfev;      /* [B] ODE solves spent on the problem                               */
    const double* lam_group;/* NULL or [G] (memspace of target): per-group lambda, see pk_local_job      */
} pk_nlls_job;
void pk_nlls_job_init(pk_nlls_job* job);
int pk_sizeof_nlls_job(void);
int pk_local_nlls_batch(pk_handle_t h, const pk_nlls_job* job);

/* ------------------------------------------------------------------------------------------------
 * Global coupled kinase-TF-protein network (reference: global_model/)
 *
 *   pk_global_upload        <- the static array part of System.odeint_args()  global_model/network.py:508-526
 *                              (kin_grid, kin_Kmat, W/TF CSR, offsets, n_sites, tf_deg, driver_map), copied once
 *   pk_global_set_loss_data <- the argument list of LOSS_FN minus Y  global_model/lossfn.py:113-121
 *                              (tables built by global_model/cache.py:19-155)
 *   pk_global_set_prior     <- defaults of GlobalODE_MOO (global_model/optproblem.py:105-114)
 *   pk_global_solve_batch   <- a loop of simulate_odeint(sys, t_eval, rtol, atol, mxstep)
 *                              (global_model/simulate.py:34-80) after System.update(**params)
 *                              (network.py:293-302), optionally followed by LOSS_FN, the objectives of
 *                              GlobalODE_MOO._evaluate (optproblem.py:87-160) and the Morris scalar of
 *                              global_model/sensitivity.py:106-140 over simulate_and_measure's fold changes
 *                              (simulate.py:105-182).
 * Kinetic models (global_model/models.py): 0 distributive, 1 sequential, 4 saturating with the block
 * [mRNA, P0, site_1..site_ns]; 2 combinatorial with the block [mRNA, pattern_0 .. pattern_{2^ns-1}]
 * (network.py:131-149; at most 8 sites per protein - blocks up to 16 patterns are inverted in registers, larger
 * ones by one warp in an L2-resident scratch; driver_map is ignored as in jacspeedup.py:318-325; the
 * reference's per-bucket rate table S_cache, jacspeedup.py:114-145, is recomputed on the device).
 * -----------------------------------------------------------------------------------------------*/
typedef struct pk_global_topology {
    int32_t model;               /* 0, 1, 2 or 4                                                    */
    int32_t N;                   /* proteins                                                        */
    int32_t K;                   /* kinases                                                         */
    int32_t n_bins;              /* kinase-grid points                                              */
    const int32_t* n_sites;      /* [N]; block of protein i = [mRNA, P0, site_1..site_ns] (model 2: patterns) */
    const int32_t* W_indptr;     /* CSR [total_sites, K]: site <- kinase weights                    */
    const int32_t* W_indices;
    const double* W_data;
    const int32_t* TF_indptr;    /* CSR [N, N]: gene <- TF weights                                  */
    const int32_t* TF_indices;
    const double* TF_data;
    const double* kin_grid;      /* [n_bins]                                                        */
    const double* kin_Kmat;      /* [K, n_bins] row-major                                           */
    const double* tf_deg;        /* [N]                                                             */
    const int32_t* driver_map;   /* [N] kinase index driving protein i's TF activity, or -1         */
    int32_t force_generic_schur; /* testing: 1 = shared-memory LU path even when the register path fits;
                                    2 = additionally all large per-system arrays in the per-CTA global scratch
                                    (the capacity fallback chosen automatically for networks beyond one CTA's
                                    shared memory) */
    int32_t reserved0;
} pk_global_topology;

typedef struct pk_global_loss_data {
    int32_t n_prot, n_rna, n_pho;
    const int32_t *p_prot, *t_prot;           /* protein index, time index into t_eval              */
    const double *obs_prot, *w_prot;
    const int32_t *p_rna, *t_rna;
    const double *obs_rna, *w_rna;
    const int32_t *p_pho, *s_pho, *t_pho;     /* protein, site-within-protein, time index           */
    const double *obs_pho, *w_pho;
    int32_t prot_base_idx, rna_base_idx, pho_base_idx;   /* runner.py:545-547                       */
} pk_global_loss_data;

enum pk_global_metric { PK_GM_NONE = -1, PK_GM_TOTAL_SIGNAL = 0, PK_GM_MEAN = 1, PK_GM_VARIANCE = 2, PK_GM_L2_NORM = 3 };

typedef struct pk_global_job {
    int32_t topo;                /* id returned by pk_global_upload                                 */
    int32_t memspace;            /* of params, y0, out_*; t_eval and mt_* are always HOST pointers  */
    int64_t B;
    int32_t T;
    int32_t theta_mode;          /* 1: params hold raw theta, physical = softplus(theta) (params.py:106-132) */
    const double* params;        /* [B,P], P = K+5N+total_sites+1: c_k|A_i|B_i|C_i|D_i|Dp_i|E_i|tf_scale */
    const double* y0;            /* [state_dim] (y0_stride 0) or [B, y0_stride]                     */
    int64_t y0_stride;
    const double* t_eval;        /* [T] strictly increasing, HOST                                   */
    double rtol, atol;           /* <=0 -> 2e-6 / 2e-9                                              */
    int32_t max_steps;           /* per system, <=0 -> 200000                                       */
    int32_t loss_mode;           /* LOSS_MODE 0..7 (lossfn.py:150-246)                              */
    int32_t metric;              /* pk_global_metric                                                */
    int32_t n_mt_prot, n_mt_rna, n_mt_pho;        /* metric time indices per modality               */
    const int32_t *mt_prot, *mt_rna, *mt_pho;     /* HOST                                           */
    int32_t mb_prot, mb_rna, mb_pho;              /* base time indices (t=0 / t=4 / t=0)            */
    int32_t reserved0;
    double lambdas[3];           /* protein, rna, phospho (optproblem.py:146-148)                   */
    double lambda_prior;
    double* out_Y;               /* [B,T,state_dim] or NULL                                         */
    double* out_loss;            /* [B,3] raw weighted sums (needs pk_global_set_loss_data) or NULL */
    double* out_F;               /* [B,3] objectives incl. prior penalty or NULL                    */
    double* out_metric;          /* [B] or NULL                                                     */
    int32_t* out_status;         /* [B] pk_status                                                   */
    int32_t* out_nsteps;
    int32_t* out_nrej;
    double* out_fc;              /* [B, n_fc] or NULL: the fold changes simulate_and_measure tabulates
                                  * (global_model/simulate.py:105-182, floors 1e-12, bases mb_*), n_fc =
                                  * N*n_mt_prot + N*n_mt_rna + total_sites*n_mt_pho, protein-major then (site,) time */
} pk_global_job;

int pk_global_upload(pk_handle_t h, const pk_global_topology* topo, int32_t* topo_id);
int pk_global_set_loss_data(pk_handle_t h, int32_t topo_id, const pk_global_loss_data* ld);
int pk_global_set_prior(pk_handle_t h, int32_t topo_id, const double* defaults /* [P] physical, HOST */);
int pk_global_release(pk_handle_t h, int32_t topo_id);
/* state_dim, number of parameters P, size of the regulator set (dense Schur block), shared memory bytes per CTA */
int pk_global_dims(pk_handle_t h, int32_t topo_id, int32_t* state_dim, int32_t* n_params, int32_t* n_reg,
                   int32_t* smem_bytes);
/* proteins, kinases and total phosphorylation sites of an uploaded network */
int pk_global_counts(pk_handle_t h, int32_t topo_id, int32_t* n_proteins, int32_t* n_kinases, int32_t* total_sites);
void pk_global_job_init(pk_global_job* job);
int pk_sizeof_global_job(void);
/* The linear systems of a step are solved through the Schur complement on the regulator set: by fixed-point sweeps to an
 * error bound of 1e-12 when the infinity norm of its iteration matrix is below 0.5 (most steps), by an exact dense
 * inverse otherwise.  Environment PHOSKIN_SCHUR_ITER=<norm bound> overrides the 0.5 (0 = exact inverse in every step;
 * the parity tests compare the two). */
int pk_global_solve_batch(pk_handle_t h, const pk_global_job* job);
/* LOSS_FN(Y, tables...) (global_model/lossfn.py:113-121, dispatch :386) on B trajectories that already
 * exist: Y [B,T,state_dim] -> out_loss [B,3] = (loss_p, loss_r, loss_ph) raw weighted sums. */
int pk_global_loss_batch(pk_handle_t h, int32_t topo_id, int32_t memspace, const double* Y, int64_t B, int32_t T,
                         int32_t loss_mode, double* out_loss);

/* The reference's CUSTOM solver — `solve_custom(sys, y0, t_eval, rtol, atol)` (global_model/jacspeedup.py:31-67) ->
 * `adaptive_rk45_model01` / `adaptive_rk45_model2` (global_model/solvers.py:292-758) — reproduced step for step on the
 * device for B parameter vectors: explicit Dormand-Prince 5(4), PI control, dt <= 1, landing on kinase-bucket
 * boundaries, cubic-Hermite outputs (rtol/atol <= 0 -> the reference's 1e-5 / 1e-7, max_steps <= 0 -> 2 000 000).
 * out_Y [B,T,state_dim]; status 0 ok, 1 max steps, 3 non-finite.  (pk_global_solve_batch is the accurate default.) */
int pk_global_solve_custom(pk_handle_t h, int32_t topo_id, int32_t memspace, int64_t B, const double* params, int32_t theta_mode,
                           const double* y0, int64_t y0_stride, const double* t_eval_host, int32_t T, double rtol, double atol,
                           int32_t max_steps, double* out_Y, int32_t* out_status, int32_t* out_nsteps, int32_t* out_nrej);

/* f(t, y) and the analytic Jacobian df/dy for B (parameter vector, state, time) triples of one uploaded network — the
 * device form of the reference's `fun(t, y)` closures (global_model/model_ivp.py:49-277: make_solve_ivp_fun_*), of
 * `rhs_odeint` (jacspeedup.py:175-375) and of `fd_jacobian_odeint` (jacspeedup.py:397-588; here analytic, not a finite
 * difference).  params [B,P] (raw thetas with theta_mode = 1), Y [B,state_dim], t_host [B] HOST (selects the kinase
 * bucket, utils.py:210-225), out_f [B,state_dim], out_J [B,state_dim,state_dim] row-major (J[i][j] = df_i/dy_j) or NULL.
 * Direct mode = the exact argument list of make_solve_ivp_fun_* (model_ivp.py:49-61): tf_direct [B,N] (the TF_inputs the
 * closure receives) and S_direct [B,total_sites] (its S_all / the S_cache column) replace the topology's kinase and TF
 * matrices (t_host may be NULL); the TF input is then squashed once, as the block kernels do (models.py:52). */
int pk_global_rhs_batch(pk_handle_t h, int32_t topo_id, int32_t memspace, int64_t B, const double* params, int32_t theta_mode,
                        const double* Y, const double* t_host, double* out_f, double* out_J, const double* tf_direct,
                        const double* S_direct);

#ifdef __cplusplus
}
#endif
#endif
