/*
 * fg_abi.h -- C ABI of the B200-native factor-graph least-squares backend (libfg_b200.so).
 *
 * This is the drop-in boundary behind rising-turtle/graph_slam's wrapper classes.  The reference
 * has no FFI layer of its own: it calls GTSAM 4.0 / g2o C++ objects directly from CGraphGT /
 * CGraphG2O / CImuBase.  Every entry point below replaces one of those call sites; the
 * reference file:line each one stands in for is cited next to it (paths relative to the
 * reference root).  INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - plain C types only; all matrices row-major fp64; a pose is 12 doubles: R (3x3 row-major) then t.
 *   - keys are gtsam::Symbol-compatible: (uint64(chr) << 56) | index  (gtsam_graph.cpp:50-54).
 *   - tangent order of Pose3 is [rot(3), trans(3)]; bias is [acc(3), gyro(3)]  (SURVEY.md A.1).
 *   - the caller owns every input buffer (copied on add); the ctx owns all device memory.
 *   - every function returns 0 (FG_OK) or a negative fg_status; fg_last_error(ctx) gives the text.
 *   - no exceptions cross the ABI.  A ctx is NOT thread-safe: one host thread per ctx, one ctx per GPU.
 *   - there is NO CPU fallback: every numeric entry point runs hand-written sm_100a CUDA and fails
 *     with FG_ERR_CUDA when no device is usable.
 */
#ifndef FG_ABI_H
#define FG_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fg_ctx fg_ctx;
typedef uint64_t fg_key;

typedef enum {
  FG_OK = 0,
  FG_ERR_INVALID = -1,       /* bad argument                                                     */
  FG_ERR_DUPLICATE_KEY = -2, /* Values::insert on an existing key (gtsam ValuesKeyAlreadyExists)  */
  FG_ERR_UNKNOWN_KEY = -3,   /* Values::at / factor on a missing key (ValuesKeyDoesNotExist)      */
  FG_ERR_CUDA = -4,          /* CUDA runtime failure / no device                                  */
  FG_ERR_INDETERMINATE = -5, /* reduced system not positive definite at lambda_upper              */
  FG_ERR_NCCL = -6,
  FG_ERR_STATE = -7          /* call made in the wrong state (e.g. optimise an empty graph)       */
} fg_status;

/* ------------------------------------------------------------------ lifetime */
/* One context = one NonlinearFactorGraph + Values pair on one GPU
 * (CGraphGT::CGraphGT, gtsam/gtsam_graph.cpp:75-91; rank/nranks for landmark sharding). */
fg_ctx* fg_create(int device, int rank, int nranks);
void fg_destroy(fg_ctx* ctx);
const char* fg_last_error(fg_ctx* ctx);
int fg_abi_version(void);

/* ------------------------------------------------------------------ values (gtsam::Values) */
/* mp_node_values->insert(X(id), Pose3)            gtsam_graph.cpp:333, 619, 653, 659 */
int fg_add_pose(fg_ctx* ctx, fg_key key, const double T[12]);
/* mp_node_values->insert(V(id), Vector3)          gtsam_graph.cpp:350, 622 */
int fg_add_vec3(fg_ctx* ctx, fg_key key, const double v[3]);
/* mp_node_values->insert(B(id), ConstantBias)     gtsam_graph.cpp:354, 623 */
int fg_add_bias(fg_ctx* ctx, fg_key key, const double b[6]);
/* mp_node_values->insert<Point3>(Q(id), q)        gtsam_graph.cpp:391 */
int fg_add_point(fg_ctx* ctx, fg_key key, const double p[3]);
int fg_add_points(fg_ctx* ctx, int64_t n, const fg_key* keys, const double* p3);
/* mp_node_values->insert(L(id), OrientedPlane3)   gtsam_graph.cpp:1198 ; pl = (nx,ny,nz,d), n normalised */
int fg_add_plane(fg_ctx* ctx, fg_key key, const double pl[4]);
/* mp_node_values->update(key, value)              gtsam_graph.cpp:666 */
int fg_update_value(fg_ctx* ctx, fg_key key, const double* value);
/* mp_node_values->exists(key)                     gtsam_graph.cpp:632-633 ; returns 1/0 */
int fg_exists(fg_ctx* ctx, fg_key key);
/* mp_node_values->at<T>(key)                      gtsam_graph.cpp:651, test_vro_imu_graph.cpp:348-350
 * writes 12/3/6/3/4 doubles depending on the key's type; *n_out receives the count (may be NULL). */
int fg_get_value(fg_ctx* ctx, fg_key key, double* out, int* n_out);
/* batch accessors: all values of one type in insertion order (type codes below). */
enum { FG_T_POSE = 0, FG_T_VEC3 = 1, FG_T_BIAS = 2, FG_T_POINT = 3, FG_T_PLANE = 4 };
int64_t fg_num_values(fg_ctx* ctx, int type);
int fg_get_values(fg_ctx* ctx, int type, double* out);        /* device -> host, whole array   */
int fg_set_values(fg_ctx* ctx, int type, const double* in);   /* host -> device, whole array   */

/* ------------------------------------------------------------------ factors (NonlinearFactorGraph::add) */
/* PriorFactor<Pose3>(key, T, Gaussian information)       gtsam_graph.cpp:341 */
int fg_add_prior_pose(fg_ctx* ctx, fg_key key, const double T[12], const double info[36]);
/* PriorFactor<Vector3>                                   gtsam_graph.cpp:362 */
int fg_add_prior_vec3(fg_ctx* ctx, fg_key key, const double mean[3], const double info[9]);
/* PriorFactor<imuBias::ConstantBias>                     gtsam_graph.cpp:363 */
int fg_add_prior_bias(fg_ctx* ctx, fg_key key, const double mean[6], const double info[36]);
/* PriorFactor<Point3>(Q(id), q, Isotropic::Sigma(3,s))   gtsam_graph.cpp:379, 394 */
int fg_add_prior_point(fg_ctx* ctx, fg_key key, const double mean[3], double sigma);
int fg_add_prior_points(fg_ctx* ctx, int64_t n, const fg_key* keys, const double* mean3, double sigma);
/* BetweenFactor<Pose3>(k1, k2, T, Gaussian::Information(info))   gtsam_graph.cpp:689-692 */
int fg_add_between(fg_ctx* ctx, fg_key k1, fg_key k2, const double T[12], const double info[36]);

/* Cal3DS2(fx,fy,s,u0,v0,k1,k2[,p1,p2]) and body_P_sensor (mp_u2c)   gtsam_graph.cpp:373, 405-406 */
int fg_set_calibration(fg_ctx* ctx, int calib_id, const double K[9]);
int fg_set_sensor(fg_ctx* ctx, int sensor_id, const double T[12]);
/* GenericProjectionFactor<Pose3,Point3,Cal3DS2>(uv, Isotropic(2,sigma), X, Q, K, false, false, body_P_sensor)
 *                                                                    gtsam_graph.cpp:405-406, 423, 433 */
int fg_add_projection(fg_ctx* ctx, fg_key kpose, fg_key kpoint, const double uv[2], double sigma,
                      int calib_id, int sensor_id);
int fg_add_projections(fg_ctx* ctx, int64_t n, const fg_key* kpose, const fg_key* kpoint, const double* uv2,
                       double sigma, int calib_id, int sensor_id);
/* OrientedPlane3Factor(z, Gaussian::Covariance(S), X(node), L(landmark))   gtsam_graph.cpp:1265 */
int fg_add_plane_factor(fg_ctx* ctx, fg_key kpose, fg_key kplane, const double z[4], const double cov[9]);

/* ------------------------------------------------------------------ IMU preintegration */
/* PreintegratedCombinedMeasurements::Params (imu_vn100.cpp:24-67, imu_base.cpp:258-263) */
typedef struct {
  double acc_cov[9];            /* accelerometerCovariance  */
  double gyro_cov[9];           /* gyroscopeCovariance      */
  double int_cov[9];            /* integrationCovariance    */
  double bias_acc_cov[9];       /* biasAccCovariance        */
  double bias_gyro_cov[9];      /* biasOmegaCovariance      */
  double bias_acc_omega_int[36];/* biasAccOmegaInt          */
  double gravity[3];            /* n_gravity (MakeSharedD(g) -> (0,0,+g)) */
} fg_imu_params;

/* State of a PreintegratedCombinedMeasurements (TangentPreintegration). */
typedef struct {
  double dt;           /* deltaTij                                   */
  double preint[9];    /* [theta, position, velocity]                */
  double H_ba[27];     /* preintegrated_H_biasAcc   (9x3 row-major)  */
  double H_bg[27];     /* preintegrated_H_biasOmega (9x3 row-major)  */
  double bias_hat[6];  /* biasHat [acc, gyro]                        */
  double cov[225];     /* preintMeasCov (15x15 row-major)            */
  double gravity[3];
} fg_pim;

/* The CImuBase::predictNext sample loop (imu_base.cpp:76-85): for every sample [gx gy gz ax ay az]
 * integrateMeasurement(acc, gyro, dt).  n_intervals independent intervals run in one launch; interval k
 * uses samples [offsets[k], offsets[k+1]).  `ctx` may be NULL (device 0 is used).  */
int fg_preintegrate(fg_ctx* ctx, int n_intervals, const int* offsets, const double* imu6, double dt,
                    const fg_imu_params* params, const double* bias_hat6, fg_pim* out);
/* PreintegrationBase::predict(state_i, bias_i)  (imu_base.cpp:86) -- O(1) host arithmetic on a finished pim. */
int fg_pim_predict(const fg_pim* pim, const double pose_i[12], const double vel_i[3], const double bias_i[6],
                   double pose_j[12], double vel_j[3]);
/* CombinedImuFactor(X_i, V_i, X_j, V_j, B_i, B_j, pim)   test_vro_imu_graph.cpp:191-196; keys in that order */
int fg_add_imu(fg_ctx* ctx, const fg_key keys[6], const fg_pim* pim);

/* ------------------------------------------------------------------ optimise */
/* gtsam::LevenbergMarquardtParams, defaults = GTSAM defaults */
typedef struct {
  double lambda_initial;     /* 1e-5  */
  double lambda_factor;      /* 10    */
  double lambda_upper;       /* 1e5   */
  double lambda_lower;       /* 0     */
  double min_model_fidelity; /* 1e-3  */
  int max_iterations;        /* 100   */
  double relative_error_tol; /* 1e-5  */
  double absolute_error_tol; /* 1e-5  */
  double error_tol;          /* 0     */
  int force_iterations;      /* benchmark mode: run exactly max_iterations outer iterations */
  int verbosity;
} fg_lm_params;
void fg_lm_params_default(fg_lm_params* p);

#define FG_TRACE_MAX 512
typedef struct {
  int iterations;            /* outer LM iterations performed                */
  int trials;                /* inner lambda trials (damped solves)          */
  double initial_error;      /* graph.error(values) before                   */
  double final_error;        /* graph.error(values) after                    */
  double lambda;             /* lambda on exit                               */
  int status;                /* fg_status of the run                         */
  int trace_len;             /* entries valid in the trace arrays            */
  double trace_lambda[FG_TRACE_MAX];
  double trace_error[FG_TRACE_MAX];      /* error before the trial            */
  double trace_new_error[FG_TRACE_MAX];  /* error after the trial (inf if the solve failed) */
  int trace_accepted[FG_TRACE_MAX];
  /* device time per phase, ms, summed over the run (CUDA events) */
  double ms_linearize, ms_schur, ms_factor, ms_solve, ms_retract_error, ms_total;
  /* workload description used for the roofline */
  int64_t n_reduced_dims, n_supernodes, nnz_L, n_projections, n_landmarks;
  /* single-kernel device times (ms, summed): the observation pass of the linearisation and the Schur tile kernel */
  double ms_proj_obs, ms_schur_blocks;
  int64_t n_schur_pairs, n_levels;
  /* multi-GPU: bytes of the per-trial allreduce of the reduced pose Hessian (packed structural non-zeros + rhs + chi2) */
  int64_t allreduce_bytes;
  int64_t nnz_S;             /* structural non-zeros of the reduced system before the factorisation (nnz_L includes the fill) */
} fg_lm_report;

/* Build the symbolic structure and upload the graph (idempotent; called lazily by the functions below). */
int fg_finalize(fg_ctx* ctx);
/* LevenbergMarquardtOptimizer(graph, values).optimize(); values <- result
 *   CGraphGT::optimizeGraphBatch  gtsam_graph.cpp:1784-1788 (and optimizeGraph :1779-1782) */
int fg_optimize_lm(fg_ctx* ctx, const fg_lm_params* params, fg_lm_report* report);
/* graph.error(values) = 1/2 sum |r|^2_Sigma     CGraphGT::error  gtsam_graph.cpp:173-176 */
int fg_error(fg_ctx* ctx, double* error);

/* ------------------------------------------------------------------ manifold charts */
/* GTSAM chooses the Pose3 / Rot3 retraction at COMPILE time (GTSAM_POSE3_EXPMAP, GTSAM_ROT3_EXPMAP); the reference links an
 * author-local GTSAM 4.0 whose flags are unknown (gtsam/CMakeLists.txt:12-18).  The chart decides Values::retract of a
 * Pose3, the residual Local(measured, h) of BetweenFactor / PriorFactor<Pose3> at non-zero error, and the 6-vector of the
 * VRO log (gtsam_graph.cpp:60,1532), not the minimiser.  Default: full EXPMAP. */
enum {
  FG_CHART_EXPMAP = 0,              /* Pose3 EXPMAP (Rot3 chart irrelevant): GTSAM 4.1+ default           */
  FG_CHART_FIRST_ORDER_EXPMAP = 2,  /* Pose3 FIRST_ORDER over Rot3 EXPMAP                                  */
  FG_CHART_FIRST_ORDER_CAYLEY = 3   /* Pose3 FIRST_ORDER over Rot3 CAYLEY: GTSAM 4.0's default build       */
};
int fg_set_pose_chart(fg_ctx* ctx, int chart);

/* ------------------------------------------------------------------ g2o back-end (CGraphG2O, BASELINE config 1) */
/* g2o::EdgeSE3 between two VertexSE3 (poses added with fg_add_pose):  setMeasurement(mr.edge.transform),
 * setInformation(mr.edge.informationMatrix)   CGraphG2O::addToGraph  g2o/g2o_graph.cpp:125-132.
 * Error = toVectorMQT(Z^-1 X1^-1 X2) = [t, q_xyz]; `info` is 6x6 row-major in that [trans, rot] order.  A graph holds
 * either GTSAM pose factors or g2o edges, not both (fg_finalize rejects the mix). */
int fg_add_g2o_edge(fg_ctx* ctx, fg_key k1, fg_key k2, const double T[12], const double info[36]);
/* VertexSE3::setFixed(true)   CGraphG2O::firstNode  g2o/g2o_graph.cpp:90.  Only g2o edges may touch a fixed vertex. */
int fg_set_fixed(fg_ctx* ctx, fg_key key, int fixed);
typedef struct {
  int iterations;           /* 20: `int iter = 20`                                  g2o_graph.cpp:244 */
  int iterations_per_call;  /* 2 : mp_optimizer->optimize(ceil(iter/10))            g2o_graph.cpp:249 */
  double tau;               /* 1e-5: OptimizationAlgorithmLevenberg lambda init = tau * max diag(H)    */
  int max_trials;           /* 10 : maxTrialsAfterFailure                                              */
} fg_g2o_params;
void fg_g2o_params_default(fg_g2o_params* p);
#define FG_G2O_TRACE_MAX 64
typedef struct {
  int iterations;            /* LM iterations performed (the reference asks for 20)                       */
  int calls;                 /* optimize() calls made (lambda is re-initialised at the start of each)     */
  double initial_chi2, final_chi2;   /* sum e^T Omega e, no 1/2 (g2o chi2())                               */
  double lambda;
  int status;
  int trace_len;
  double trace_chi2[FG_G2O_TRACE_MAX];     /* chi2 after the iteration  */
  double trace_lambda[FG_G2O_TRACE_MAX];   /* lambda after the iteration */
  int trace_trials[FG_G2O_TRACE_MAX];      /* damped solves it took      */
  double ms_total;
} fg_g2o_report;
/* mp_optimizer->initializeOptimization(); for (i = 0; i < iter; i += currIt) currIt = mp_optimizer->optimize(2);
 *   CGraphG2O::optimizeGraph  g2o/g2o_graph.cpp:241-252  (OptimizationAlgorithmLevenberg + BlockSolver<6,3> + CSparse
 *   Cholesky, :65-77): gain ratio rho = (chi2 - chi2_new) / (dx.(lambda dx + b) + 1e-3); on success
 *   lambda *= max(1/3, min(1 - (2 rho - 1)^3, 2/3)), on failure lambda *= nu, nu *= 2, at most max_trials solves; an
 *   iteration that ends without progress terminates the optimize() call.  (The reference's loop never ends if optimize()
 *   returns 0; this entry point returns instead.) */
int fg_optimize_g2o(fg_ctx* ctx, const fg_g2o_params* params, fg_g2o_report* report);
/* mp_optimizer->computeActiveErrors(); return mp_optimizer->chi2();   CGraphG2O::error  g2o_graph.cpp:254-258 */
int fg_g2o_chi2(fg_ctx* ctx, double* chi2);

/* ISAM2Params as CGraphGT sets them (initISAM2Params, gtsam_graph.cpp:93-99). */
typedef struct {
  double relinearize_threshold;   /* 0.1 */
  int relinearize_skip;           /* 1   */
} fg_isam2_params;
void fg_isam2_params_default(fg_isam2_params* p);
typedef struct {
  double error_before;       /* graph.error at the linearisation point of this update        */
  double error_after;        /* graph.error at the new estimate                              */
  int64_t n_variables;       /* variables in the graph (all types)                           */
  int64_t n_relinearized;    /* variables whose linearisation point moved in this update     */
  int64_t n_new_variables;   /* variables that entered since the previous update             */
  double ms_update;          /* device time of the update (gating, linearise, solve, retract) */
  double ms_rebuild;         /* host time spent re-analysing / uploading the grown graph      */
  int status;
} fg_inc_report;
/* isam2->update(new factors, new values); values <- isam2->calculateEstimate()
 *   CGraphGT::optimizeGraphIncremental  gtsam_graph.cpp:1768-1776 (called once per frame by
 *   test_vro_imu_graph.cpp:344 and test_ba_imu_graph.cpp:427).
 * The factors and values added through fg_add_* since the previous call are the "new" ones.  The context keeps ISAM2's
 * two states per variable -- the linearisation point theta and the estimate theta (+) delta; fg_get_value(s) return the
 * estimate.  One call = one update(): new variables enter at their initial value, every variable whose delta has a
 * component >= relinearize_threshold moves its linearisation point (every relinearize_skip calls), the graph is
 * linearised at theta and ONE undamped Gauss-Newton system is solved on the device (full re-factorisation instead of
 * ISAM2's partial Bayes-tree re-elimination: the same delta when ISAM2's wildfire threshold is 0).  An indefinite
 * system returns FG_ERR_INDETERMINATE (IndeterminantLinearSystemException) and leaves the estimate unchanged.
 * fg_optimize_lm / fg_error / fg_marginal_cov afterwards start from the estimate (the reference copies it into
 * mp_node_values) and end the incremental session. */
int fg_update_incremental(fg_ctx* ctx, const fg_isam2_params* params, fg_inc_report* report);

/* Marginals(graph, values, Marginals::CHOLESKY).marginalCovariance(key)
 *   CGraphGT::bundleAdjust gtsam_graph.cpp:598-601 (edge information = inverse of pose 1's marginal covariance),
 *   planeNodeAssociation :1357, gtsam/test/convert_vo2ba.cpp:413-416.
 * Linearises at the current values, factors the UNDAMPED reduced system on the device and returns the d x d
 * (row-major) covariance of one pose (d = 6), velocity (3), bias (6) or plane (3) variable.  Point3 keys are
 * eliminated by the Schur complement and are not supported (FG_ERR_INVALID).  */
int fg_marginal_cov(fg_ctx* ctx, fg_key key, double* cov, int* dim);

/* Diagnostic: measured fp64 throughput of `device` in TFLOP/s, out[0] = DFMA (CUDA cores), out[1] = DMMA
 * (mma.sync.m8n8k4.f64).  bench.py divides its fp64 roofline by these (MEASURED_PEAKS.json holds bf16 only). */
int fg_debug_fp64_peak(int device, double out[2]);

/* ------------------------------------------------------------------ multi-GPU */
/* Rank 0 fills a 128-byte NCCL unique id; the host launcher ships it to the other ranks (any transport);
 * every rank then calls fg_comm_init.  Landmarks added on a rank are that rank's shard; pose-side
 * variables and factors must be added identically on every rank. */
int fg_comm_unique_id(char id[128]);
int fg_comm_init(fg_ctx* ctx, const char id[128]);
/* Declares pose-pose couplings that exist only through landmarks held by OTHER ranks, so that every rank builds
 * the same reduced-system structure (no numeric effect; a superset of the true co-visibility is allowed). */
int fg_add_structure_edges(fg_ctx* ctx, int64_t n, const fg_key* kpose_a, const fg_key* kpose_b);

/* ------------------------------------------------------------------ introspection for tests/bench */
/* Linearise at the current values and copy out chi2 (=2*error), the reduced gradient norm and sizes. */
int fg_debug_sizes(fg_ctx* ctx, int64_t out[8]);

#ifdef __cplusplus
}
#endif
#endif /* FG_ABI_H */
