/*
 * gpgrid.h -- C ABI of libgpgrid.so: the B200 (sm_100a) exact-GP-on-grids engine that sits under
 * the gpim.reconstructor / gpim.boptimizer Python API.
 *
 * The reference (ziatdinovmax/GPim) has no FFI: its boundary to the arithmetic is the Python
 * object protocol of pyro.contrib.gp.models.GPRegression.  Each entry point below names the
 * reference call site whose arithmetic it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer argument is a DEVICE pointer owned by the caller (torch tensor data_ptr())
 *     unless its name ends in _host; the library never frees or retains caller memory;
 *   - matrices are row-major with an explicit leading dimension (in elements);
 *   - work is enqueued on the cudaStream_t passed as `stream` (void* here so the header needs
 *     no CUDA include); calls do not synchronise unless documented;
 *   - dtype selects the arithmetic type of ALL floating-point buffers of the call
 *     (GPG_F32 = float, GPG_F64 = double), mirroring precision="single"/"double"
 *     (gpim/gpreg/gpr.py:92-99);
 *   - return value: GPG_OK, or an error code with text in gpg_last_error();
 *   - theta (constrained hyper-parameters) is a device array of dtype, length 3 + d:
 *       theta[0] variance, theta[1] noise, theta[2] scale_mixture (RationalQuadratic only),
 *       theta[3 + k] lengthscale of input dimension k (an isotropic kernel repeats its value);
 *   - thread-compatible: a handle must not be used from two host threads at once.
 */
#ifndef GPGRID_H
#define GPGRID_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPG_VERSION 120

enum { GPG_OK = 0, GPG_EINVAL = 1, GPG_ENOTPD = 2, GPG_ECUDA = 3 };
enum { GPG_F32 = 0, GPG_F64 = 1 };
/* kernel book of gpim/kernels/pyro_kernels.py:58-68 */
enum { GPG_RBF = 0, GPG_MATERN52 = 1, GPG_RATQUAD = 2 };
/* gpim/gpbayes/acqfunc.py:11-92 */
enum { GPG_ACQ_CB = 0, GPG_ACQ_EI = 1, GPG_ACQ_POI = 2 };
/* gpg_set_option keys.  Every default is the fastest correct setting; the alternatives exist for A/B measurements. */
enum {
    GPG_OPT_GEMM_PATH = 1,        /* 0 auto (f32: tcgen05 when N >= 1024; f64: DMMA tiles when N >= 256), 1 SIMT only, 2 force tcgen05 */
    GPG_OPT_PREDICT_CHUNK = 2,    /* test points per internal tile of gpg_predict (0 = auto: 16384) */
    GPG_OPT_STAGE_TIMING = 3,     /* != 0: bracket every stage with CUDA events, read by gpg_stage_times */
    GPG_OPT_PANEL_REFINE = 4,     /* recursive factorisation: refine every panel solve against L11 (default 1) */
    GPG_OPT_SYRK_CHUNK = 5,       /* recursive factorisation: longest K accumulated in TMEM before an fp32 round-to-nearest
                                     add (multiple of 64; 0 = unlimited, the default) */
    GPG_OPT_FACTOR_ALGO = 6,      /* f32 tensor-core factorisation: 0 (default) two-level blocked Cholesky followed by the
                                     batched triangular inverse; 1 recursive Cholesky + inverse */
    GPG_OPT_FIT_GRAPH = 7,        /* gpg_fit_adam on small problems (SIMT path): replay one captured iteration as a CUDA
                                     graph (default 1) */
    GPG_OPT_PANEL_MODE = 8,       /* blocked Cholesky panel: 0 forward substitution against the diagonal factor, 1 tcgen05
                                     GEMM through the block inverse, 2 SIMT GEMM through the block inverse, 3 (default) the
                                     whole 512-column panel as one cooperative kernel with a device-side dependency
                                     chain: diagonal blocks on one CTA, panel products and updates of every row block on
                                     tcgen05 with the update sums held in tensor memory (chol_panel.cuh) */
    GPG_OPT_OUTER_PANEL = 9,      /* blocked Cholesky: width of the outer panel (multiple of 128; default 512 -- wider is a
                                     few per cent faster at N > 15 000 but doubles the TMEM accumulation bias of the update) */
    GPG_OPT_COMPACT_SUPPORT = 10, /* gpg_predict, tcgen05 path (default 1; 0 = always dense): per 128-row tile of test points,
                                     restrict the variance GEMM to the contiguous range of training rows whose covariance
                                     with the tile exceeds 1e-14 x variance (what lies outside contributes below fp32
                                     resolution, whatever theta is: |L^-1_ij| <= (noise + jitter)^-1/2).  The range is found
                                     on the device while K* is assembled, so the policy is automatic: a tile whose support
                                     is everything (long lengthscales, unordered training rows) runs the dense product */
    GPG_OPT_INNER_LEFT = 11,      /* blocked Cholesky, update inside the outer panel: 1 (default) left-looking (next block
                                     column only, all inner panels so far), 0 right-looking (all remaining columns) */
    GPG_OPT_LOOKAHEAD = 12,       /* blocked Cholesky with the cooperative panel (panel mode 3): 1 (default) the trailing
                                     update of an outer panel is split -- the next panel's columns first, on the caller's
                                     stream; the rest on a side stream with a capped grid, UNDER the next panel's chain */
    GPG_OPT_PANEL_WORKERS = 13    /* cooperative panel: cap on the number of worker CTAs (0 = auto) */
};
/* stages reported by gpg_stage_times */
enum {
    GPG_ST_KMAT = 0,        /* K(X,X) assembly */
    GPG_ST_CHOLESKY = 1,
    GPG_ST_TRTRI = 2,
    GPG_ST_SOLVE = 3,       /* vector solves + logdet */
    GPG_ST_KCROSS = 4,      /* K(X*,X) tile assembly + predictive mean */
    GPG_ST_PGEMM = 5,       /* Linv K* with the fused column-sum-of-squares epilogue */
    GPG_ST_PFINAL = 6,      /* sd epilogue */
    GPG_ST_GRAD = 7,        /* Kinv + gradient reduction + Adam step */
    GPG_ST_ACQ = 8,
    GPG_ST_COUNT = 9
};

typedef struct gpg_handle_s *gpg_handle_t;

int gpg_version(void);
const char *gpg_last_error(void);
int gpg_create(int device, gpg_handle_t *out);
int gpg_destroy(gpg_handle_t h);
int gpg_set_option(gpg_handle_t h, int key, long long value);
/* number of kernel launches this handle has enqueued since creation (bench.py gpu_launches) */
long long gpg_launch_count(gpg_handle_t h);
/* Multiply-accumulates (at tile granularity: 128 x 256 x 32 per executed k-block, algorithmic MACs -- the tensor
 * pipe executes three fp16 MMAs for each) the variance GEMM of gpg_predict has executed while GPG_OPT_STAGE_TIMING
 * was on, since the last call; synchronises the device and clears the count.  With GPG_OPT_COMPACT_SUPPORT this is
 * what the kernel really did, as opposed to the dense N^2 / 2 per test point. */
int gpg_variance_gemm_macs(gpg_handle_t h, double *macs_host);
/* bytes of device workspace currently owned by the handle */
size_t gpg_workspace_bytes(gpg_handle_t h);
/* Synchronises the device, then adds up the CUDA-event spans recorded since the last call:
 * ms_host[GPG_ST_COUNT] device milliseconds per stage, spans_host[GPG_ST_COUNT] number of
 * brackets per stage (host arrays).  Clears the record. */
int gpg_stage_times(gpg_handle_t h, double *ms_host, long long *spans_host);

/* The engine's tensor-core GEMM on caller matrices (f32, row-major): C = alpha * A B^T + beta * C with
 * A [M x K], B [N x K].  Operands are split into fp16 hi/lo planes after multiplication by the
 * power-of-two scale_a / scale_b (choose them so that scale * max|x| <= 2^15) and multiplied as
 * Ahi Bhi + Ahi Blo + Alo Bhi on tcgen05 with fp32 accumulation in TMEM. */
int gpg_gemm_nt_f32(gpg_handle_t h, const float *A, int64_t lda, const float *B, int64_t ldb, float *C,
                    int64_t ldc, int64_t M, int64_t N, int64_t K, double alpha, double beta,
                    double scale_a, double scale_b, void *stream);

/* K1/K2 -- kernel-matrix assembly.  Replaces Pyro Isotropy.forward reached from
 * gpim/gpreg/gpr.py:192,248 (kernel(X), kernel(X, Xnew)) with the kernels configured at
 * gpim/kernels/pyro_kernels.py:58-68.
 * out[i*ld + j] = k_theta(X_i, Z_j) for i < N, j < P;  Z == NULL means Z = X, P = N and
 * (theta.noise + jitter) is added on the diagonal (GPRegression.model: Kff.view(-1)[::N+1] += ...).
 * lower_only != 0 (Z == NULL only) skips tiles strictly above the diagonal. */
int gpg_kmat(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta,
             const void *X, int64_t N, const void *Z, int64_t P, double jitter, int lower_only,
             void *out, int64_t ld, void *stream);

/* K3 -- in-place blocked Cholesky of the lower triangle of A (torch.linalg.cholesky inside
 * GPRegression.model / .forward, reached from gpr.py:192,248).  The strict upper triangle is
 * not referenced and is left unspecified.  *info (device int32) receives 0, or 1 + the index of
 * the first non-positive pivot; no host synchronisation. */
int gpg_cholesky(gpg_handle_t h, int dtype, void *A, int64_t N, int64_t ld, int32_t *info, void *stream);

/* Linv = L^-1 (lower; strict upper triangle of Linv is written as zero).  L and Linv must not alias. */
int gpg_trtri(gpg_handle_t h, int dtype, const void *L, int64_t N, int64_t ld,
              void *Linv, int64_t ldinv, void *stream);

/* K7a -- vhat = L^-1 y, alpha = L^-T vhat, logdet = sum_i log L_ii  (MultivariateNormal.log_prob
 * in GPRegression.model, gpr.py:192; the y column of util.conditional's pack, gpr.py:248).
 * Uses Linv with residual correction against L.  scalars_out (dtype[2]): {0.5*|vhat|^2, logdet}. */
int gpg_solve_vec(gpg_handle_t h, int dtype, const void *L, const void *Linv, int64_t N, int64_t ld,
                  const void *y, void *vhat_out, void *alpha_out, void *scalars_out, void *stream);

/* K1+K3+trtri+K7a in one call: the factor cache {L, Linv, alpha, vhat} for fixed theta.
 * Replaces the kernel(X)+cholesky that GPRegression.forward redoes on every predict (gpr.py:248).
 * wsplit_out / scales_out (both NULL or both set; f32 only): the tensor-core form of Linv --
 * 2*N*ld fp16 values (hi plane then lo plane, power-of-two scaled) and float[16] operand scales --
 * which gpg_predict needs to run on the tcgen05 path (without them it runs the SIMT kernels). */
int gpg_factorize(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta,
                  const void *X, const void *y, int64_t N, double jitter,
                  void *L, void *Linv, int64_t ld, void *vhat_out, void *alpha_out,
                  void *scalars_out, int32_t *info, void *wsplit_out, float *scales_out, void *stream);

/* K2+K4+K5 -- predictive mean and standard deviation at M test points, tiled over M internally
 * (never materialises the N x M cross-kernel).  Replaces util.conditional + the noise add and
 * sqrt at gpr.py:248-250:  mean = K*^T alpha,  sd = sqrt(max(v - colsum((Linv K*)^2), 0) + noise).
 * Rows of Xs that contain NaN give NaN outputs (acqfunc.py:57-59 relies on it). */
int gpg_predict(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta,
                const void *X, int64_t N, const void *Linv, int64_t ld, const void *alpha,
                const void *wsplit, const float *scales,
                const void *Xs, int64_t M, void *mean_out, void *sd_out, void *stream);

/* Analytic grid variant: Xs is not read; test point j has coordinates unravel(j0 + j, dims)*step
 * (np.mgrid layout of gprutils.get_full_grid, gprutils.py:136).  dims_host/step_host: host arrays [d]. */
int gpg_predict_grid(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta,
                     const void *X, int64_t N, const void *Linv, int64_t ld, const void *alpha,
                     const void *wsplit, const float *scales,
                     const int64_t *dims_host, const double *step_host, int64_t j0, int64_t M,
                     void *mean_out, void *sd_out, void *stream);

/* K7 -- negative log marginal likelihood and its gradient w.r.t. the constrained theta
 * (Trace_ELBO.differentiable_loss + backward at gpr.py:192-193, minus the constant Uniform
 * log-priors).  nll_out: dtype[1]; grad_out: dtype[3 + d] in theta layout. */
int gpg_nll_grad(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta,
                 const void *X, const void *y, int64_t N, double jitter,
                 void *nll_out, void *grad_out, int32_t *info, void *stream);

/* Whole training loop on the device (reconstructor.train, gpr.py:170-217): `iters` Adam steps
 * (torch.optim.Adam defaults) on the unconstrained parameters, constrained values recorded after
 * each step, no host synchronisation.
 *   u        dtype[3 + n_ls] in/out: unconstrained {variance, noise, scale_mixture, lengthscale[n_ls]}
 *            variance, lengthscale: sigmoid onto [lo, hi] (interval constraint); noise,
 *            scale_mixture: exp (positive constraint)
 *   bounds_host  double[2 + 2*n_ls]: {var_lo, var_hi, ls_lo[n_ls], ls_hi[n_ls]}
 *   n_ls     1 (isotropic) or d (ARD)
 *   traj_out dtype[iters * (4 + d)]: per iteration {theta[0..3+d), loss}
 *   theta_out dtype[3 + d]: constrained theta after the last step */
int gpg_fit_adam(gpg_handle_t h, int dtype, int kernel_id, int d, int n_ls,
                 const void *X, const void *y, int64_t N, double jitter,
                 void *u, const double *bounds_host, int iters, double lr,
                 void *traj_out, void *theta_out, int32_t *info, void *stream);

/* gpg_fit_adam with the GPyTorch parametrisation of skreconstructor(ski=False) -- the exact-GP branch of
 * gpim/gpreg/skgpr.py:143-150,189-203 (ExactGP: ConstantMean + ScaleKernel(RBF | Matern52) + GaussianLikelihood,
 * loss = -ExactMarginalLogLikelihood = nll / N, torch.optim.Adam on the raw parameters):
 *   u  dtype[3 + n_ls] in/out: raw {outputscale, noise, mean constant, lengthscale[n_ls]}
 *      outputscale = softplus(u0) (Positive), noise = softplus(u1) + 1e-4 (GreaterThan(1e-4)), constant = u2,
 *      lengthscale = lo + (hi - lo) sigmoid(u) (gpytorch.constraints.Interval, gpytorch_kernels.py:55-57)
 *   bounds_host  as gpg_fit_adam; only the lengthscale bounds are used
 *   traj_out / theta_out  as gpg_fit_adam, with theta[0] = outputscale and theta[2] = mean constant; the recorded
 *      loss is nll / N.  kernel_id: GPG_RBF or GPG_MATERN52 (gpytorch_kernels.py:60-69); jitter is normally 0. */
int gpg_fit_adam_sk(gpg_handle_t h, int dtype, int kernel_id, int d, int n_ls,
                    const void *X, const void *y, int64_t N, double jitter,
                    void *u, const double *bounds_host, int iters, double lr,
                    void *traj_out, void *theta_out, int32_t *info, void *stream);

/* gpg_fit_adam for vreconstructor(independent=True) -- gpim/gpreg/vgpr.py:320-354 (ivgprmodel) trained as at
 * vgpr.py:157-179: `ntasks` exact GPs on one shared X with ONE shared lengthscale (the base kernel's parameter exists
 * before its batch_shape is overwritten, vgpr.py:346), per-task ScaleKernel outputscales and ConstantMean constants,
 * MultitaskGaussianLikelihood noise_t = task_noise_t + global noise (both softplus + 1e-4);
 * loss = -sum_t log N(y_t; c_t, s_t K_l + noise_t I) / (N ntasks), torch.optim.Adam on the raw parameters.
 *   Y        dtype[ntasks * N], task-major (row t = observations of output t at the N training rows)
 *   u        dtype[3 ntasks + 1 + n_ls] in/out raw {outputscale[ntasks] | task noise[ntasks] | global noise |
 *            constant[ntasks] | lengthscale[n_ls]} (GPyTorch initialises all of them to 0)
 *   bounds_host  double[2 n_ls] {ls_lo[n_ls], ls_hi[n_ls]} (gpytorch.constraints.Interval), or NULL for GPyTorch's
 *            default Positive constraint (softplus) -- the reference's lengthscale=None
 *   traj_out dtype[iters * (d + 1)]: per iteration {lengthscale[d], loss}
 *   theta_out dtype[ntasks * (3 + d)]: per task {outputscale, total noise, constant, lengthscale[d]} after the last
 *            step -- the theta gpg_factorize / gpg_predict take for that task (on y_t - constant). */
int gpg_fit_adam_mt(gpg_handle_t h, int dtype, int kernel_id, int d, int n_ls, int ntasks, const void *X,
                    const void *Y, int64_t N, double jitter, void *u, const double *bounds_host, int iters, double lr,
                    void *traj_out, void *theta_out, int32_t *info, void *stream);

/* K6 -- acquisition sweep + top-k (acqfunc.py:11-92, boptim.py:303-315).
 *   acq_id CB: alpha*mean + beta*sd;  EI: imp*Phi(z) + sd*phi(z), imp = mean - mu_best - xi,
 *   z = imp/sd;  POI: Phi(z).   mask (nullable, dtype[M]): multiplied in, NaN entries excluded.
 *   topk_val dtype[k], topk_idx int64[k]: descending value, ties by descending flat index
 *   (the reversed ascending argsort of boptim.py:304-306); unmasked NaNs rank first, as there.
 *   count_out int32[1]: number of valid entries written (< k only when masked).  acq_out nullable. */
int gpg_acq_sweep(gpg_handle_t h, int dtype, int acq_id, const void *mean, const void *sd,
                  const void *mask, int64_t M, double mu_best, double xi, double alpha, double beta,
                  int k, void *topk_val, int64_t *topk_idx, int32_t *count_out, void *acq_out, void *stream);

/* K6, the point filters on the ranked list (boptim.py:378-429 checkvalues, boptim.py:326-376 update_points; Python lists
 * and a scipy cKDTree in the reference).  topk_val / topk_idx / count: the output of gpg_acq_sweep, still on the device.
 *   dims_host int64[ndim]: shape of the dense grid (flat indices are row-major over it);
 *   visited int64[n_visited] (device): flat indices of every point measured so far, oldest first;
 *   a candidate is admissible when it is not in `visited` and farther than dscale * gamma^q from the q-th most recent
 *   visited point, q < memory (dscale 0 = the reference's dscale=None);
 *   do_batch: greedy ball suppression of radius batch_dscale (closed ball) from the first candidate whose value equals
 *   the first admissible one's, at most batch_out_max picks.
 *   sel_out int32[4 + batch_out_max] (device): {first admissible position or -1, start position of the cut list or -1,
 *   number of picks, 1 if a NaN value was seen, pick positions...}.  The exit strategies and the random padding of the
 *   reference draw from numpy's generator and stay with the caller. */
int gpg_acq_select(gpg_handle_t h, int dtype, const void *topk_val, const int64_t *topk_idx, const int32_t *count,
                   int k, int ndim, const int64_t *dims_host, const int64_t *visited, int n_visited, int memory,
                   double dscale, double gamma, int do_batch, double batch_dscale, int batch_out_max, int32_t *sel_out,
                   void *stream);

/* ---------------------------------------------------------------------------------------------
 * Inducing-point GP: reconstructor(sparse=True), gpim/gpreg/gpr.py:145-155,198-199 over pyro's
 * SparseGPRegression with its default VFE approximation.  Xu: m x d inducing inputs
 * (gpr.py:151: X[::len(X) // indpoints]), 0 < m <= N, m < 65536.  theta as above; jitter goes on the
 * diagonal of k(Xu, Xu) only.  *info: 0, or 1 + the first non-positive pivot of either m x m
 * factorisation (k(Xu, Xu) + jitter I, then I + W^T W / noise).
 * --------------------------------------------------------------------------------------------- */

/* One evaluation of SparseGPRegression.model's objective (Trace_ELBO.differentiable_loss at gpr.py:192
 * minus the constant Uniform log-priors):
 *   loss = -log N(y; 0, Qff + noise I) + clamp(tr(Kff - Qff) / noise, 0) / 2,  Qff = Kfu Kuu^-1 Kuf,
 * and what loss.backward() leaves behind: grad_theta_out dtype[3 + d] (theta layout, constrained
 * values), grad_xu_out dtype[m * d]. */
int gpg_sparse_loss_grad(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta,
                         const void *X, const void *y, int64_t N, const void *Xu, int64_t m, double jitter,
                         void *loss_out, void *grad_theta_out, void *grad_xu_out, int32_t *info, void *stream);

/* reconstructor.train with sparse=True (gpr.py:186-199) on the device: `iters` Adam steps on the
 * unconstrained hyper-parameters (u, bounds_host, n_ls, traj_out, theta_out exactly as gpg_fit_adam)
 * AND on the inducing inputs Xu (in/out; a plain Parameter in pyro: no constraint), one optimiser.
 * xu_traj_out (nullable) dtype[iters * m * d]: Xu after every step (hyperparams["inducing_points"],
 * gpr.py:198-199). */
int gpg_sparse_fit_adam(gpg_handle_t h, int dtype, int kernel_id, int d, int n_ls,
                        const void *X, const void *y, int64_t N, void *Xu, int64_t m, double jitter,
                        void *u, const double *bounds_host, int iters, double lr,
                        void *traj_out, void *xu_traj_out, void *theta_out, int32_t *info, void *stream);

/* Factor cache of the inducing-point posterior for fixed (theta, Xu) -- what SparseGPRegression.forward
 * recomputes on every call (gpr.py:248): Ui = Luu^-1 and Pm = LA^-1 Luu^-1 (m x m lower triangular,
 * row-major, leading dimension ld, strict upper triangle zero) and w (dtype[m]) with
 * mean(x*) = k(x*, Xu) . w.
 * split_out / scales_out (both NULL or both set; f32 with ld % 8 == 0 only): the tensor-core form of the two factors
 * -- 4 * m * ld fp16 values (Ui hi plane, Ui lo plane, Pm hi, Pm lo, power-of-two scaled) and float[24] operand scales
 * -- which gpg_sparse_predict needs to run on the tcgen05 path (without them it runs the SIMT kernels). */
int gpg_sparse_factorize(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta,
                         const void *X, const void *y, int64_t N, const void *Xu, int64_t m, double jitter,
                         void *Ui_out, void *Pm_out, int64_t ld, void *w_out, int32_t *info,
                         void *split_out, float *scales_out, void *stream);

/* SparseGPRegression.forward(Xnew, full_cov=False, noiseless=False) + sqrt (gpr.py:248-250), tiled over M:
 *   mean = K*^T w,  sd = sqrt(v + noise - colsum((Ui K*)^2) + colsum((Pm K*)^2)),  K* = k(Xu, X*).
 * Rows of Xs that contain NaN give NaN outputs. */
int gpg_sparse_predict(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta,
                       const void *Xu, int64_t m, const void *Ui, const void *Pm, int64_t ld, const void *w,
                       const void *split, const float *scales,
                       const void *Xs, int64_t M, void *mean_out, void *sd_out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU of one box, one NCCL communicator per handle.  The reference has no
 * multi-device path (gpr.py:136-140 moves the one model to the one GPU); what is sharded is reconstructor.predict
 * (gpr.py:219-255): test points are independent given the factor cache, so rank `root` alone factorises (training /
 * Cholesky stay replicas-only), the cache is broadcast over NVLink, every rank predicts its contiguous tile of
 * X_full rows and one all-gather returns (mean, sd).  NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy
 * the process already has loaded, else the system one); without it these entry points return GPG_ECUDA and
 * everything else keeps working.  Collective calls: every rank of the communicator must make the same call.
 * --------------------------------------------------------------------------------------------- */

/* Rank 0 draws the 128-byte NCCL unique id (host memory) and hands it to the other ranks by any out-of-band means
 * (torch.distributed object broadcast, MPI_Bcast, a file); every rank then calls gpg_comm_init with it. */
int gpg_comm_unique_id(void *id_host);
int gpg_comm_init(gpg_handle_t h, int nranks, int rank, const void *id_host);
int gpg_comm_destroy(gpg_handle_t h);
/* nranks / rank of the handle's communicator; 0 / -1 when there is none */
int gpg_comm_info(gpg_handle_t h, int *nranks_out, int *rank_out);

/* 1 when gpg_predict would read the fp16 planes (wsplit) of a cache for this dtype / N under the handle's options,
 * 0 when it would read Linv: what a factor cache has to carry to another GPU. */
int gpg_predict_uses_planes(gpg_handle_t h, int dtype, int64_t N, int have_planes);

/* Broadcast of the factor cache of gpg_factorize from rank `root`, in place on every rank, stream-ordered on
 * `stream`: {theta[3 + d], X[N x d], alpha[N], scales[16], info} as one fused NCCL launch, then the N x N part --
 * the fp16 planes when gpg_predict_uses_planes(), otherwise Linv (wsplit / scales may be NULL then). */
int gpg_bcast_factor(gpg_handle_t h, int dtype, int d, int64_t N, int64_t ld, void *theta, void *X, void *Linv,
                     void *alpha, void *wsplit, float *scales, int32_t *info, int root, void *stream);

/* One all-gather: pred_all[r * count + i] = rank r's pred_local[i]  (dtype elements; count equal on all ranks). */
int gpg_allgather_pred(gpg_handle_t h, int dtype, const void *pred_local, int64_t count, void *pred_all, void *stream);

/* The sharded predict in one call: gpg_bcast_factor from `root` + gpg_predict of this rank's M_local rows +
 * gpg_allgather_pred.  On the tcgen05 route the broadcast of the planes is pipelined by row blocks on the handle's
 * communication stream and the variance GEMM of the first tile of test points starts on the n-blocks of Linv that
 * have landed while the rest is in flight (the n-blocks are independent: block b needs rows [256 b, 256 b + 256)).
 *   pred_local  dtype[2 * M_pad]: mean at [0, M_local), sd at [M_pad, M_pad + M_local)
 *   pred_all    dtype[nranks * 2 * M_pad] or NULL (no gather): rank r's pred_local at offset r * 2 * M_pad
 * M_pad (>= M_local) must be the same on all ranks.  *info on every rank receives root's factorisation status. */
int gpg_predict_sharded(gpg_handle_t h, int dtype, int kernel_id, int d, void *theta, void *X, int64_t N,
                        void *Linv, int64_t ld, void *alpha, void *wsplit, float *scales, int32_t *info, int root,
                        const void *Xs_local, int64_t M_local, void *pred_local, int64_t M_pad, void *pred_all,
                        void *stream);

/* K6 over a sharded grid (boptim.py:303-315 is what it replaces): gpg_acq_sweep on this rank's tile with global flat
 * indices idx_offset + j, a k * nranks-element all-gather, and the merge -- every rank ends with the global top-k
 * in the reference's order.  mean_local / sd_local / mask_local / acq_out_local: this rank's M_local entries. */
int gpg_acq_sweep_sharded(gpg_handle_t h, int dtype, int acq_id, const void *mean_local, const void *sd_local,
                          const void *mask_local, int64_t M_local, int64_t idx_offset, double mu_best, double xi,
                          double alpha, double beta, int k, void *topk_val, int64_t *topk_idx, int32_t *count_out,
                          void *acq_out_local, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GPGRID_H */
