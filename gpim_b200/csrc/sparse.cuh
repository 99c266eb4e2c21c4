// sparse.cuh -- inducing-point GP (VFE, Titsias 2009): the kernels behind reconstructor(sparse=True)
// (gpim/gpreg/gpr.py:145-155,198-199 over pyro's SparseGPRegression).
//
// With Xu (m x d) the inducing inputs, s2 = noise, Kuu = k(Xu) + jitter I = Luu Luu^T, Ui = Luu^-1,
// B = Ui k(Xu, X) (m x N), A' = I + B B^T / s2 = LA LA^T, beta = B y, a = A'^-1 beta / s2:
//
//   loss   = 1/2 [y.y/s2 - |LA^-1 beta|^2/s2^2 + N log s2 + 2 sum log LA_ii + N log 2pi] + 1/2 max(0, (N v - |B|_F^2)/s2)
//   dF/dKuu = -1/2 Ui^T (2 I - A'^-1 - A' - a a^T) Ui                          (symmetric, sum over all i, j)
//   dF/dKuf = (1/s2) [Ui^T (A'^-1 - I) B - (Ui^T a) rho^T],  rho = y - B^T a
//   dF/ds2  = -y.y/(2 s2^2) + a.beta/s2^2 - (a.beta/s2 - a.a)/(2 s2) + N/(2 s2) - (m - tr A'^-1)/(2 s2) - T/(2 s2^2)
//   dF/dv  += N/(2 s2)   (the k(x, x) = v diagonal of the trace term),   T = N v - |B|_F^2
//
// (derivation in DESIGN.md section 8; checked against autograd of the oracle in tests/test_sparse_oracle.py).
// The chain onto (variance, lengthscales, scale mixture, inducing inputs) is one fused pass per sensitivity
// matrix that re-evaluates the covariance derivative on the fly (sgp_kgrad_kernel), as grad_partial_kernel does
// for the exact GP.  Every dense product runs on the engine's GEMM; nothing here synchronises with the host.
#pragma once
#include "common.cuh"
#include "train.cuh"

enum { SGP_SC_YY = 0, SGP_SC_A0BETA = 1, SGP_SC_A0A0 = 2, SGP_SC_TRAINV = 3, SGP_SC_BFRO = 4, SGP_SC_LOGDET = 5,
       SGP_SC_C0C0 = 6, SGP_SC_COUNT = 8 };

// Operand scales of the tcgen05 products (fp32 path, large m and N).  The split-fp16 GEMM wants every operand
// multiplied by a power of two that brings its largest entry just under 2^14.
//   SGP_S_K  k(X, Xu) entries are <= variance                         (rigorous, from theta)
//   SGP_S_B  B = Luu^-1 k(Xu, X): column norms^2 = Qff_nn <= k(x, x) = variance, so |B_ij| <= sqrt(variance)
//   SGP_S_U, SGP_S_T  Luu^-1 and T2 = Ui^T (A'^-1 - I): measured (one pass over an m x m matrix)
//   SGP_S_LA LA^-1: A' = I + B B^T / s2 >= I, so |LA^-1_ij| <= 1    (rigorous)
//   SGP_S_PHI, SGP_S_H, SGP_S_XT  the m x m operands of the gradient chain: measured
enum { SGP_S_K = 0, SGP_S_U = 1, SGP_S_UK_INV = 2, SGP_S_B = 3, SGP_S_T = 4, SGP_S_TB_INV = 5, SGP_S_BB_INV = 6,
       SGP_S_LA = 7, SGP_S_LALA_INV = 8, SGP_S_PHI = 9, SGP_S_UPHI_INV = 10, SGP_S_H = 11, SGP_S_UH_INV = 12,
       SGP_S_XT = 13, SGP_S_XTU_INV = 14,
       SGP_S_P = 16, SGP_S_KP_INV = 17,      // predict: Pm = LA^-1 Luu^-1 operand (measured) and 1 / (s_K s_P)
       SGP_S_COUNT = 24 };

__device__ __forceinline__ float sgp_pow2_scale(float bound) {
    return (bound > 0.f && bound < 3.0e38f) ? exp2f(floorf(log2f(16384.f / bound))) : 1.f;
}

template <typename T> __global__ void sgp_scales_theta_kernel(const T *__restrict__ theta, float *__restrict__ scales) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float v = (float)theta[0];
    scales[SGP_S_K] = sgp_pow2_scale(v);
    scales[SGP_S_B] = sgp_pow2_scale(sqrtf(v));
    scales[SGP_S_BB_INV] = 1.0f / (scales[SGP_S_B] * scales[SGP_S_B]);
    scales[SGP_S_LA] = 16384.f;
    scales[SGP_S_LALA_INV] = 1.0f / (16384.f * 16384.f);
}

// max |M_ij| over rows x cols in two launches: per-block maxima (SGP_AMAX_BLOCKS blocks), then
// scales[slot] = power-of-two scale for the maximum and scales[slot_inv] = 1 / (scales[slot] * scales[other])
constexpr int SGP_AMAX_BLOCKS = 64;

__global__ void __launch_bounds__(256) sgp_absmax_partial_kernel(const float *__restrict__ M, int64_t ld, int64_t rows,
                                                                 int64_t cols, float *__restrict__ blockmax) {
    __shared__ float red[8];
    float mx = 0.f;
    const int lane = threadIdx.x & 31;
    for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < rows; i += (int64_t)gridDim.x * 8)
        for (int64_t j = lane; j < cols; j += 32) {
            const float a = fabsf(M[i * ld + j]);
            if (a < 3.0e38f) mx = fmaxf(mx, a);      // ignores inf / NaN (a failed factorisation is reported through info)
        }
    mx = warp_max(mx);
    if (lane == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t = fmaxf(t, red[w]);
        blockmax[blockIdx.x] = t;
    }
}

__global__ void sgp_absmax_finish_kernel(const float *__restrict__ blockmax, int nblocks, float *__restrict__ scales,
                                         int slot, int other, int slot_inv) {
    float t = 0.f;
    for (int b = threadIdx.x; b < nblocks; b += 32) t = fmaxf(t, blockmax[b]);
    t = warp_max(t);
    if (threadIdx.x == 0) {
        const float sc = sgp_pow2_scale(t);
        scales[slot] = sc;
        scales[slot_inv] = 1.0f / (sc * scales[other]);
    }
}

// out[j] = beta * yin[j] + alpha * sum_i A[i][j] x[i] in two deterministic passes: a 2-D grid of (128-column strip) x
// (row chunk) blocks leaves double partial sums per chunk, a second kernel adds the chunks up.  (One thread per
// column over all rows keeps too few loads in flight: 7688 x 769 took 110 us.)
constexpr int SGP_GEMVT_CHUNKS = 16;

template <typename T>
__global__ void __launch_bounds__(128) gemvT_rect_partial_kernel(const T *__restrict__ A, int64_t lda, int64_t rows,
                                                                 int64_t cols, const T *__restrict__ x,
                                                                 double *__restrict__ part) {
    const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (j >= cols) return;
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = min(rows, r0 + per);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int64_t i = r0;
    for (; i + 3 < r1; i += 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] += (double)A[(i + e) * lda + j] * (double)x[i + e];
    }
    for (; i < r1; ++i) acc[0] += (double)A[i * lda + j] * (double)x[i];
    part[(int64_t)blockIdx.y * cols + j] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

template <typename T>
__global__ void __launch_bounds__(256) gemvT_rect_finish_kernel(const double *__restrict__ part, int nchunks, int64_t cols,
                                                                const T *__restrict__ yin, T alpha, T beta,
                                                                T *__restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= cols) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += part[(int64_t)c * cols + j];
    out[j] = (T)((double)alpha * s + (yin ? (double)beta * (double)yin[j] : 0.0));
}

// theta0 = theta with the noise entry cleared (Kuu carries jitter only on its diagonal)
template <typename T> __global__ void sgp_theta0_kernel(const T *__restrict__ theta, int P, T *__restrict__ theta0) {
    const int p = threadIdx.x;
    if (blockIdx.x == 0 && p < P) theta0[p] = (p == 1) ? T(0) : theta[p];
}

// out[i] = sum_j A[i][j] x[j]   (rows x cols, one warp per row, double accumulation, four loads in flight per lane)
template <typename T>
__global__ void __launch_bounds__(256) gemv_rect_kernel(const T *__restrict__ A, int64_t lda, int64_t rows, int64_t cols,
                                                        const T *__restrict__ x, T *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= rows) return;
    const T *__restrict__ row = A + i * lda;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    int64_t j = lane;
    for (; j + 96 < cols; j += 128) {
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] += (double)row[j + 32 * e] * (double)x[j + 32 * e];
    }
    for (; j < cols; j += 32) acc[0] += (double)row[j] * (double)x[j];
    const double tot = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
    if (lane == 0) out[i] = (T)tot;
}

// out[i] = v[i] / theta[1]
template <typename T>
__global__ void sgp_div_noise_kernel(const T *__restrict__ v, int64_t n, const T *__restrict__ theta, T *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = (T)((double)v[i] / (double)theta[1]);
}

// S = sum_z Spart[z] (the split-K partial products of B B^T, lower tiles valid) and A' = I + S / s2: S, Ap and LA
// are written in full (symmetric)
template <typename T>
__global__ void __launch_bounds__(256) sgp_form_A_kernel(const T *__restrict__ Spart, int nz, int64_t zstride, int64_t ld,
                                                         int64_t m, const T *__restrict__ theta, T *__restrict__ S,
                                                         T *__restrict__ Ap, T *__restrict__ LA) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
    if (j >= m || i >= m) return;
    const int64_t src = (j <= i) ? i * ld + j : j * ld + i;
    double s = 0.0;
    for (int z = 0; z < nz; ++z) s += (double)Spart[(int64_t)z * zstride + src];
    const T v = (T)(s / (double)theta[1]) + (i == j ? T(1) : T(0));
    S[i * ld + j] = (T)s;
    Ap[i * ld + j] = v;
    LA[i * ld + j] = v;
}

// Phi = 2 I - A'^-1 - A' - a a^T and H = A'^-1 - I, both full, from the lower triangle of Ainv
template <typename T>
__global__ void __launch_bounds__(256) sgp_form_phi_kernel(const T *__restrict__ Ainv, const T *__restrict__ Ap,
                                                           const T *__restrict__ a, int64_t ld, int64_t m,
                                                           T *__restrict__ Phi, T *__restrict__ H) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
    if (j >= m || i >= m) return;
    const T ai = (j <= i) ? Ainv[i * ld + j] : Ainv[j * ld + i];
    const T dlt = (i == j) ? T(1) : T(0);
    Phi[i * ld + j] = T(2) * dlt - ai - Ap[i * ld + j] - a[i] * a[j];
    H[i * ld + j] = ai - dlt;
}

// sum_ij G_ij d k(A_i, Z_j) / d{variance, scale_mixture, lengthscale_k, A_i}, G_ij = gscale * (Gm[i][j] - w_i rho_j)
// with gscale = host_scale * (inv_noise ? 1 / theta[1] : 1).  One warp per (row i, column chunk blockIdx.y): the
// columns are split over gridDim.y chunks so that a 769 x 7 688 sensitivity matrix keeps 776 CTAs busy instead of 97
// (one chunk: 12.5 % of the warp slots, 2.3 % of the HBM roofline).  partial[chunk][block][3 + D] in double (entry 1,
// the noise, stays zero); row sums for the inducing inputs either straight into gA[i][k] (= xu_factor * sum_j ..., or +=
// when accumulate; gridDim.y == 1 only) or, chunked, as doubles into gAp[chunk][i][k] for sgp_gxu_reduce_kernel.
template <typename T, int KID, int D>
__global__ void __launch_bounds__(256) sgp_kgrad_kernel(const T *__restrict__ theta, const T *__restrict__ A, int64_t P,
                                                        const T *__restrict__ Z, int64_t Q, const T *__restrict__ Gm,
                                                        int64_t ld, const T *__restrict__ w, const T *__restrict__ rho,
                                                        double host_scale, int inv_noise, double xu_factor,
                                                        int accumulate, double *__restrict__ partial,
                                                        T *__restrict__ gA, double *__restrict__ gAp = nullptr) {
    constexpr int NP = 3 + D;
    __shared__ double red[8][NP];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 8 + wp;
    double g[NP], gx[D];
#pragma unroll
    for (int p = 0; p < NP; ++p) g[p] = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) gx[k] = 0.0;
    if (i < P) {
        const Theta<T> th = load_theta<T, D>(theta);
        const double gscale = host_scale * (inv_noise ? 1.0 / (double)theta[1] : 1.0);
        T x[D];
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = A[i * D + k];
        const double wi = w ? (double)w[i] : 0.0;
        const int64_t qc = (((Q + gridDim.y - 1) / gridDim.y + 31) / 32) * 32;     // columns per chunk
        const int64_t j_end = min(Q, (int64_t)(blockIdx.y + 1) * qc);
        for (int64_t j = (int64_t)blockIdx.y * qc + lane; j < j_end; j += 32) {
            double G = (double)Gm[i * ld + j];
            if (w) G -= wi * (double)rho[j];
            G *= gscale;
            T dl[D], q[D];
            T r2 = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                dl[k] = (x[k] - Z[j * D + k]) * th.inv_ls[k];
                q[k] = dl[k] * dl[k];
                r2 += q[k];
            }
            T dv, dl_common, da = T(0);     // dk/dv; dk/dl_k = dl_common q_k / l_k; dk/dA_ik = -dl_common dl_k / l_k
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
            g[0] += G * (double)dv;
            g[2] += G * (double)da;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                g[3 + k] += G * (double)(dl_common * q[k] * th.inv_ls[k]);
                gx[k] -= G * (double)(dl_common * dl[k] * th.inv_ls[k]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const double s = warp_sum(g[p]);
        if (lane == 0) red[wp][p] = s;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const double s = warp_sum(gx[k]) * xu_factor;
        if (lane == 0 && i < P) {
            if (gAp) gAp[((int64_t)blockIdx.y * P + i) * D + k] = s;
            else if (gA) gA[i * D + k] = (T)(s + (accumulate ? (double)gA[i * D + k] : 0.0));
        }
    }
    __syncthreads();
    if (threadIdx.x < NP) {
        double s = 0.0;
        for (int r = 0; r < 8; ++r) s += red[r][threadIdx.x];
        partial[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * NP + threadIdx.x] = s;
    }
}

// gA[i][k] = sum over the chunk rows of gAp (already scaled by their xu_factor): nrows chunks of P x D doubles
template <typename T>
__global__ void __launch_bounds__(256) sgp_gxu_reduce_kernel(const double *__restrict__ gAp, int nrows, int64_t PD,
                                                             T *__restrict__ gA) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= PD) return;
    double s = 0.0;
    for (int r = 0; r < nrows; ++r) s += gAp[(int64_t)r * PD + e];
    gA[e] = (T)s;
}

// the scalar reductions of one evaluation (single CTA, double): see SGP_SC_*
template <typename T>
__global__ void __launch_bounds__(1024) sgp_scalars_kernel(const T *__restrict__ y, int64_t N, const T *__restrict__ beta,
                                                           const T *__restrict__ a0, const T *__restrict__ c0,
                                                           const T *__restrict__ S, const T *__restrict__ Ainv,
                                                           const T *__restrict__ LA, int64_t ld, int64_t m,
                                                           double *__restrict__ sc) {
    __shared__ double red[32][SGP_SC_COUNT];
    double v[SGP_SC_COUNT];
#pragma unroll
    for (int p = 0; p < SGP_SC_COUNT; ++p) v[p] = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += 1024) { const double t = (double)y[i]; v[SGP_SC_YY] += t * t; }
    for (int64_t i = threadIdx.x; i < m; i += 1024) {
        const double b = (double)beta[i], a = (double)a0[i], c = (double)c0[i];
        v[SGP_SC_A0BETA] += a * b;
        v[SGP_SC_A0A0] += a * a;
        v[SGP_SC_C0C0] += c * c;
        v[SGP_SC_TRAINV] += (double)Ainv[i * ld + i];
        v[SGP_SC_BFRO] += (double)S[i * ld + i];
        v[SGP_SC_LOGDET] += log((double)LA[i * ld + i]);
    }
#pragma unroll
    for (int p = 0; p < SGP_SC_COUNT; ++p) {
        const double s = warp_sum(v[p]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][p] = s;
    }
    __syncthreads();
    if (threadIdx.x < SGP_SC_COUNT) {
        double tot = 0.0;
        for (int w = 0; w < 32; ++w) tot += red[w][threadIdx.x];
        sc[threadIdx.x] = tot;
    }
}

// grad_theta (theta layout: variance, noise, scale_mixture, lengthscale[d]) and the loss from the partial sums
// of the two sgp_kgrad passes and the scalars
template <typename T>
__global__ void __launch_bounds__(256) sgp_finish_kernel(const double *__restrict__ partA, int nbA,
                                                         const double *__restrict__ partB, int nbB, int P,
                                                         const double *__restrict__ sc, const T *__restrict__ theta,
                                                         int64_t N, int64_t m, T *__restrict__ grad_out,
                                                         T *__restrict__ loss_out) {
    __shared__ double red[8];
    __shared__ double tot[GPG_MAX_P];
    for (int p = 0; p < P; ++p) {
        double s = 0.0;
        for (int b = threadIdx.x; b < nbA; b += 256) s += partA[(int64_t)b * P + p];
        for (int b = threadIdx.x; b < nbB; b += 256) s += partB[(int64_t)b * P + p];
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[w];
            tot[p] = t;
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    const double v = (double)theta[0], s2 = (double)theta[1];
    const double dN = (double)N, dm = (double)m;
    const double yy = sc[SGP_SC_YY];
    const double ab = sc[SGP_SC_A0BETA] / s2;            // a . beta
    const double aa = sc[SGP_SC_A0A0] / (s2 * s2);       // a . a
    const double Tr = dN * v - sc[SGP_SC_BFRO];          // N v - |B|_F^2
    const double trace_term = Tr / s2 > 0.0 ? Tr / s2 : 0.0;
    const double loss = 0.5 * (yy / s2 - sc[SGP_SC_C0C0] / (s2 * s2) + dN * log(s2) + 2.0 * sc[SGP_SC_LOGDET] +
                               dN * 1.8378770664093454835606594728112) + 0.5 * trace_term;
    // The clamp(min=0) of the trace term only engages through rounding (T >= 0 mathematically, T -> 0 when every
    // point is an inducing point); the gradient is that of the unclamped objective.
    tot[0] += dN / (2.0 * s2);
    tot[1] = -yy / (2.0 * s2 * s2) + ab / (s2 * s2) - (ab / s2 - aa) / (2.0 * s2) + dN / (2.0 * s2) -
             (dm - sc[SGP_SC_TRAINV]) / (2.0 * s2) - Tr / (2.0 * s2 * s2);
    for (int p = 0; p < P; ++p) grad_out[p] = (T)tot[p];
    if (loss_out) loss_out[0] = (T)loss;
}

// torch.optim.Adam step on the inducing inputs (same hyper-parameters and step count as the theta group, which
// adam_step_kernel has already advanced: st->step is the CURRENT step).  traj (nullable): [iters][n] record.
template <typename T>
__global__ void __launch_bounds__(256) sgp_adam_xu_kernel(FitCfg c, const FitState *__restrict__ st, int64_t n,
                                                          T *__restrict__ xu, const T *__restrict__ g,
                                                          T *__restrict__ m1, T *__restrict__ m2, T *__restrict__ traj) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int step = st->step;
    const double bc1 = 1.0 - pow(c.beta1, (double)step);
    const double bc2 = 1.0 - pow(c.beta2, (double)step);
    const double step_size = c.lr / bc1;
    const double bc2_sqrt = sqrt(bc2);
    const T gi = g[i];
    T m = m1[i], v = m2[i];
    m = m + (T)(1.0 - c.beta1) * (gi - m);
    v = v * (T)c.beta2 + ((T)(1.0 - c.beta2) * gi) * gi;
    const T denom = gpg_sqrt(v) / (T)bc2_sqrt + (T)c.eps;
    const T x = xu[i] + ((T)(-step_size) * m) / denom;
    xu[i] = x;
    m1[i] = m;
    m2[i] = v;
    if (traj) traj[(int64_t)(step - 1) * n + i] = x;
}

// sd = sqrt(v + noise - sum_t part1[t][j] + sum_t part2[t][j])  (SparseGPRegression.forward, full_cov=False,
// noiseless=False: no clamp there either); NaN coordinates -> NaN
template <typename T, int D>
__global__ void sgp_predict_finalize_kernel(const T *__restrict__ theta, const T *__restrict__ part1,
                                            const T *__restrict__ part2, int ntiles, int64_t ldpart,
                                            TestPoints<T, D> tp, int64_t mc, T *__restrict__ sd) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= mc) return;
    double s1 = 0.0, s2 = 0.0;
    for (int t = 0; t < ntiles; ++t) {
        s1 += (double)part1[(int64_t)t * ldpart + j];
        s2 += (double)part2[(int64_t)t * ldpart + j];
    }
    T z[D];
    tp.load(j, z);
    bool bad = false;
#pragma unroll
    for (int k = 0; k < D; ++k) bad |= (z[k] != z[k]);
    const double var = (double)theta[0] + (double)theta[1] - s1 + s2;
    sd[j] = bad ? T(NAN) : (T)sqrt(var);
}
