// train.cuh -- K7: marginal-likelihood gradient and the on-device Adam loop
// (reconstructor.train, gpr.py:170-217).  One iteration = kmat -> cholesky -> trtri -> solves ->
// Kinv = Linv^T Linv -> fused gradient reduction -> Adam step, all enqueued without host syncs.
#pragma once
#include "common.cuh"

#define GPG_MAX_P (3 + GPG_MAX_D)

struct FitState {               // lives in device workspace, double regardless of dtype
    double m[GPG_MAX_P], v[GPG_MAX_P];
    double dtheta_du[GPG_MAX_P];
    int step;
};

// mode 0: pyro parametrisation of reconstructor (Uniform priors -> interval constraints on variance and
// lengthscale, positive noise / scale mixture through exp; gpr.py + pyro_kernels.py:81-94).
// mode 1: GPyTorch parametrisation of skreconstructor(ski=False) (skgpr.py:143-150, gpytorch_kernels.py:55-73):
// slot 0 = ScaleKernel outputscale (Positive: softplus), slot 1 = GaussianLikelihood noise (GreaterThan(1e-4):
// softplus + 1e-4), slot 2 = ConstantMean constant (unconstrained), lengthscale: Interval (sigmoid).
struct FitCfg {
    int d, n_ls, is_rq, mode;
    double var_lo, var_hi, ls_lo[GPG_MAX_D], ls_hi[GPG_MAX_D];
    double lr, beta1, beta2, eps;
    double half_n_log2pi;
};

// torch.distributions SigmoidTransform uses a clipped sigmoid (finfo.tiny .. 1 - finfo.eps)
template <typename T> __device__ __forceinline__ void interval_fwd(T u, double lo, double hi, T &val, double &dval) {
    const T tiny = sizeof(T) == 4 ? T(1.17549435e-38f) : T(2.2250738585072014e-308);
    const T eps = sizeof(T) == 4 ? T(1.1920929e-07f) : T(2.220446049250313e-16);
    T s = T(1) / (T(1) + gpg_exp(-u));
    bool clipped = false;
    if (s < tiny) { s = tiny; clipped = true; }
    if (s > T(1) - eps) { s = T(1) - eps; clipped = true; }
    const T scale = (T)(hi - lo);
    val = (T)lo + scale * s;
    dval = clipped ? 0.0 : (double)(scale * s * (T(1) - s));
}

// u -> theta (+ d theta / d u).  u layout: {variance, noise, scale_mixture, lengthscale[n_ls]}.
// torch.nn.functional.softplus (beta = 1, threshold = 20) and its derivative
template <typename T> __device__ __forceinline__ void softplus_fwd(T u, T &val, double &dval) {
    if (u > T(20)) { val = u; dval = 1.0; return; }
    val = gpg_log(T(1) + gpg_exp(u));
    dval = 1.0 / (1.0 + exp(-(double)u));
}

template <typename T>
__device__ void constrain_params(const T *u, const FitCfg &c, T *theta, double *dtheta_du) {
    T val; double dv;
    if (c.mode == 1) {
        softplus_fwd<T>(u[0], val, dv);
        theta[0] = val; dtheta_du[0] = dv;
        softplus_fwd<T>(u[1], val, dv);
        theta[1] = val + T(1e-4); dtheta_du[1] = dv;
        theta[2] = u[2]; dtheta_du[2] = 1.0;
    } else {
        interval_fwd<T>(u[0], c.var_lo, c.var_hi, val, dv);
        theta[0] = val; dtheta_du[0] = dv;
        theta[1] = gpg_exp(u[1]); dtheta_du[1] = (double)theta[1];
        theta[2] = c.is_rq ? gpg_exp(u[2]) : T(1); dtheta_du[2] = c.is_rq ? (double)theta[2] : 0.0;
    }
    for (int k = 0; k < c.n_ls; ++k) {
        interval_fwd<T>(u[3 + k], c.ls_lo[k], c.ls_hi[k], val, dv);
        dtheta_du[3 + k] = dv;
        if (c.n_ls == 1) { for (int q = 0; q < c.d; ++q) theta[3 + q] = val; }
        else theta[3 + k] = val;
    }
}

// 0.5 * sum_ij G_ij dA_ij/dtheta_p over the lower triangle, G = Kinv - alpha alpha^T.
// One warp per (row i, column chunk blockIdx.y); partial[chunk][block][3+D] in double.  Reads Kinv's lower triangle once.
constexpr int GRAD_CHUNKS_MAX = 8;
template <typename T, int KID, int D>
__global__ void __launch_bounds__(256) grad_partial_kernel(const T *__restrict__ theta, const T *__restrict__ X,
                                                           const T *__restrict__ alpha, const T *__restrict__ Kinv,
                                                           int64_t ld, int64_t N, double *__restrict__ partial) {
    constexpr int P = 3 + D;
    __shared__ double red[8][P];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 8 + w;
    double g[P];
#pragma unroll
    for (int p = 0; p < P; ++p) g[p] = 0.0;
    if (i < N) {
        const Theta<T> th = load_theta<T, D>(theta);
        T x[D];
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = X[i * D + k];
        const T ai = alpha[i];
        const int64_t qc = (((N + gridDim.y - 1) / gridDim.y + 31) / 32) * 32;          // columns per chunk
        const int64_t j_end = min(i + 1, (int64_t)(blockIdx.y + 1) * qc);
        for (int64_t j = (int64_t)blockIdx.y * qc + lane; j < j_end; j += 32) {
            const double G = (double)Kinv[i * ld + j] - (double)ai * (double)alpha[j];
            const double wgt = (j < i) ? 1.0 : 0.5;
            T q[D];
            T r2 = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                T dlt = (x[k] - X[j * D + k]) * th.inv_ls[k];
                q[k] = dlt * dlt;
                r2 += q[k];
            }
            T dv, dl_common, da = T(0);     // dA/dv, factor such that dA/dl_k = dl_common * q_k / l_k
            if (KID == GPG_RBF) {
                const T e = gpg_exp(T(-0.5) * r2);
                dv = e;
                dl_common = th.variance * e;
            } else if (KID == GPG_MATERN52) {
                const T r = gpg_sqrt(r2 + T(1e-12));
                const T s = T(2.23606797749978969641) * r;
                const T e = gpg_exp(-s);
                dv = (T(1) + s + (T(5) / T(3)) * r * r) * e;
                dl_common = th.variance * e * (T(5) / T(3)) * (T(1) + s);
            } else {
                const T base = T(1) + (T(0.5) / th.alpha) * r2;
                const T kb = gpg_pow(base, -th.alpha);
                dv = kb;
                dl_common = th.variance * kb / base;
                da = th.variance * kb * (-gpg_log(base) + r2 / (T(2) * th.alpha * base));
            }
            const double gw = G * wgt;
            g[0] += gw * (double)dv;
            if (j == i) g[1] += gw;
            g[2] += gw * (double)da;
#pragma unroll
            for (int k = 0; k < D; ++k) g[3 + k] += gw * (double)(dl_common * q[k] * th.inv_ls[k]);
        }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const double s = warp_sum(g[p]);
        if (lane == 0) red[w][p] = s;
    }
    __syncthreads();
    if (threadIdx.x < P) {
        double s = 0.0;
        for (int r = 0; r < 8; ++r) s += red[r][threadIdx.x];
        partial[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * P + threadIdx.x] = s;
    }
}

// Sums the partials into grad_theta (constrained-theta layout, dtype T) and nll.
template <typename T>
__global__ void __launch_bounds__(256) grad_finish_kernel(const double *__restrict__ partial, int nblocks, int P,
                                                          const T *__restrict__ scalars, double half_n_log2pi,
                                                          T *__restrict__ grad_out, T *__restrict__ nll_out) {
    __shared__ double red[8];
    for (int p = 0; p < P; ++p) {
        double s = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += 256) s += partial[(int64_t)b * P + p];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0;
            for (int w = 0; w < 8; ++w) tot += red[w];
            grad_out[p] = (T)tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && nll_out) nll_out[0] = (T)((double)scalars[0] + (double)scalars[1] + half_n_log2pi);
}

// GPyTorch semantics (FitCfg.mode 1).  yc = y - constant (the ConstantMean, theta[2]):
template <typename T>
__global__ void center_y_kernel(const T *__restrict__ y, int64_t N, const T *__restrict__ theta, T *__restrict__ yc) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < N) yc[i] = y[i] - theta[2];
}

// ... and after the exact-GP gradient: d nll / d constant = -sum_i alpha_i, then everything divided by N
// (ExactMarginalLogLikelihood returns log p(y) / num_data, skgpr.py:189-196).  Single CTA.
template <typename T>
__global__ void __launch_bounds__(256) sk_grad_fix_kernel(const T *__restrict__ alpha, int64_t N, int P,
                                                          T *__restrict__ grad, T *__restrict__ nll) {
    __shared__ double red[8];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += 256) s += (double)alpha[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x != 0) return;
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    const double inv_n = 1.0 / (double)N;
    for (int p = 0; p < P; ++p) grad[p] = (T)((p == 2 ? -tot : (double)grad[p]) * inv_n);
    nll[0] = (T)((double)nll[0] * inv_n);
}

// mode 0: theta = constrain(u) (start of a train() call: fresh Adam state).
// mode 1: chain rule, one torch.optim.Adam step on u, re-constrain, record {theta, loss} in row step-1 of traj.
template <typename T>
__global__ void adam_step_kernel(int mode, FitCfg c, T *__restrict__ u, FitState *__restrict__ st,
                                 const T *__restrict__ grad_theta, const T *__restrict__ nll, T *__restrict__ theta,
                                 T *__restrict__ traj_row) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int P = 3 + c.n_ls;
    if (mode == 0) {
        for (int p = 0; p < GPG_MAX_P; ++p) { st->m[p] = 0.0; st->v[p] = 0.0; }
        st->step = 0;
        constrain_params<T>(u, c, theta, st->dtheta_du);
        return;
    }
    st->step += 1;
    const double bc1 = 1.0 - pow(c.beta1, (double)st->step);
    const double bc2 = 1.0 - pow(c.beta2, (double)st->step);
    const double step_size = c.lr / bc1;
    const double bc2_sqrt = sqrt(bc2);
    for (int p = 0; p < P; ++p) {
        if (p == 2 && !c.is_rq && c.mode == 0) continue;
        double gth;
        if (p >= 3 && c.n_ls == 1) {
            gth = 0.0;
            for (int q = 0; q < c.d; ++q) gth += (double)grad_theta[3 + q];
        } else gth = (double)grad_theta[p];
        const T g = (T)(gth * st->dtheta_du[p]);
        // state kept in double but rounded through T so fp32 runs follow torch's fp32 state
        T m = (T)st->m[p], v = (T)st->v[p];
        m = m + (T)(1.0 - c.beta1) * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1)
        v = v * (T)c.beta2 + ((T)(1.0 - c.beta2) * g) * g;          // mul_(beta2).addcmul_(g, g, 1 - beta2)
        const T denom = gpg_sqrt(v) / (T)bc2_sqrt + (T)c.eps;
        u[p] = u[p] + ((T)(-step_size) * m) / denom;                // addcdiv_(exp_avg, denom, -step_size)
        st->m[p] = (double)m; st->v[p] = (double)v;
    }
    constrain_params<T>(u, c, theta, st->dtheta_du);
    if (traj_row) {
        // traj_row is the BASE of the trajectory: the row is picked from the device-side step counter, so that
        // the same launch (a replayed CUDA graph node) serves every iteration
        traj_row += (size_t)(st->step - 1) * (4 + c.d);
        for (int p = 0; p < 3 + c.d; ++p) traj_row[p] = theta[p];
        traj_row[3 + c.d] = nll[0];
    }
}

// ---------------------------------------------------------------------------------------------
// Independent multi-output GP: vreconstructor(independent=True) (gpim/gpreg/vgpr.py:320-354 over GPyTorch):
// T exact GPs on one shared X with ONE shared lengthscale (the base kernel's parameter is created before its
// batch_shape is overwritten, vgpr.py:346), per-task ScaleKernel outputscales, per-task ConstantMean constants and a
// MultitaskGaussianLikelihood (rank 0): noise_t = task_noise_t + global noise, both GreaterThan(1e-4).
//   loss = -sum_t log N(y_t; c_t, s_t K_l + noise_t I) / (N T)        (ExactMarginalLogLikelihood, vgpr.py:171-179)
// Raw-parameter layout u: {outputscale[T] | task noise[T] | global noise | constant[T] | lengthscale[n_ls]};
// lengthscale constraint: Interval (sigmoid) when bounds are given, else GPyTorch's default Positive (softplus).
// ---------------------------------------------------------------------------------------------
#define GPG_MT_MAX_TASKS 16
#define GPG_MT_MAX_P (3 * GPG_MT_MAX_TASKS + 1 + GPG_MAX_D)

struct MtCfg {
    int d, n_ls, T, ls_softplus;
    double ls_lo[GPG_MAX_D], ls_hi[GPG_MAX_D];
    double lr, beta1, beta2, eps;
};

struct MtState {
    double m[GPG_MT_MAX_P], v[GPG_MT_MAX_P];
    double dls_du[GPG_MAX_D];
    int step;
};

// theta_all[t] = {outputscale_t, task_noise_t + noise, constant_t, lengthscale[d]}
template <typename T>
__device__ void mt_constrain(const T *u, const MtCfg &c, T *theta_all, double *dls_du) {
    const int nt = c.T;
    T gn; double dgn;
    softplus_fwd<T>(u[2 * nt], gn, dgn);
    T ls[GPG_MAX_D];
    for (int k = 0; k < c.n_ls; ++k) {
        T val; double dv;
        if (c.ls_softplus) softplus_fwd<T>(u[3 * nt + 1 + k], val, dv);
        else interval_fwd<T>(u[3 * nt + 1 + k], c.ls_lo[k], c.ls_hi[k], val, dv);
        ls[k] = val; dls_du[k] = dv;
    }
    for (int t = 0; t < nt; ++t) {
        T *th = theta_all + t * GPG_MAX_P;
        T val; double dv;
        softplus_fwd<T>(u[t], val, dv);
        th[0] = val;
        softplus_fwd<T>(u[nt + t], val, dv);
        th[1] = (val + T(1e-4)) + (gn + T(1e-4));
        th[2] = u[2 * nt + 1 + t];
        for (int q = 0; q < c.d; ++q) th[3 + q] = ls[c.n_ls == 1 ? 0 : q];
    }
}

// asum[t] = sum_i alpha_i  (d nll_t / d constant_t = -asum[t]); single CTA
template <typename T>
__global__ void __launch_bounds__(256) mt_alpha_sum_kernel(const T *__restrict__ alpha, int64_t N, double *__restrict__ out) {
    __shared__ double red[8];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += 256) s += (double)alpha[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { double tot = 0.0; for (int w = 0; w < 8; ++w) tot += red[w]; *out = tot; }
}

// mode 0: fresh Adam state + theta_all = constrain(u).  mode 1: chain rule from the T per-task gradients
// (constrained-theta layout, grad_all[t][3 + d]), one Adam step, re-constrain, record {lengthscale[d], loss}.
template <typename T>
__global__ void mt_adam_kernel(int mode, MtCfg c, T *__restrict__ u, MtState *__restrict__ st,
                               const T *__restrict__ grad_all, const T *__restrict__ nll_all,
                               const double *__restrict__ asum, double n_rows, T *__restrict__ theta_all,
                               T *__restrict__ traj) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nt = c.T, P = 3 * nt + 1 + c.n_ls;
    if (mode == 0) {
        for (int p = 0; p < GPG_MT_MAX_P; ++p) { st->m[p] = 0.0; st->v[p] = 0.0; }
        st->step = 0;
        mt_constrain<T>(u, c, theta_all, st->dls_du);
        return;
    }
    const double inv = 1.0 / (n_rows * (double)nt);
    double g[GPG_MT_MAX_P];
    double gnoise = 0.0, loss = 0.0;
    double gls[GPG_MAX_D] = {0.0, 0.0, 0.0, 0.0};
    T tmp; double dgn;
    softplus_fwd<T>(u[2 * nt], tmp, dgn);
    for (int t = 0; t < nt; ++t) {
        const T *gt = grad_all + t * GPG_MAX_P;
        double dv;
        softplus_fwd<T>(u[t], tmp, dv);
        g[t] = (double)gt[0] * dv * inv;
        softplus_fwd<T>(u[nt + t], tmp, dv);
        g[nt + t] = (double)gt[1] * dv * inv;
        gnoise += (double)gt[1];
        g[2 * nt + 1 + t] = -asum[t] * inv;
        for (int q = 0; q < c.d; ++q) gls[c.n_ls == 1 ? 0 : q] += (double)gt[3 + q];
        loss += (double)nll_all[t];
    }
    g[2 * nt] = gnoise * dgn * inv;
    for (int k = 0; k < c.n_ls; ++k) g[3 * nt + 1 + k] = gls[k] * st->dls_du[k] * inv;
    st->step += 1;
    const double bc1 = 1.0 - pow(c.beta1, (double)st->step);
    const double bc2 = 1.0 - pow(c.beta2, (double)st->step);
    const double step_size = c.lr / bc1, bc2_sqrt = sqrt(bc2);
    for (int p = 0; p < P; ++p) {
        const T gp = (T)g[p];
        T m = (T)st->m[p], v = (T)st->v[p];
        m = m + (T)(1.0 - c.beta1) * (gp - m);
        v = v * (T)c.beta2 + ((T)(1.0 - c.beta2) * gp) * gp;
        const T denom = gpg_sqrt(v) / (T)bc2_sqrt + (T)c.eps;
        u[p] = u[p] + ((T)(-step_size) * m) / denom;
        st->m[p] = (double)m; st->v[p] = (double)v;
    }
    mt_constrain<T>(u, c, theta_all, st->dls_du);
    if (traj) {
        T *row = traj + (size_t)(st->step - 1) * (c.d + 1);
        for (int q = 0; q < c.d; ++q) row[q] = theta_all[3 + q];
        row[c.d] = (T)(loss * inv);
    }
}
