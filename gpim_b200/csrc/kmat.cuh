// kmat.cuh -- K1 (symmetric K(X,X) + diag) and K2 (cross-kernel rows for a tile of test points,
// fused with the predictive mean row-reduction).  Both are HBM-write-bound: one FMA chain + one
// exp per 4-byte store, 128-bit stores, coordinates re-read through L1/L2.
#pragma once
#include "common.cuh"

// Where test-point coordinates come from: an (M, d) array, or the analytic np.mgrid layout.
template <typename T, int D> struct TestPoints {
    const T *Xs;            // nullptr -> analytic grid
    int64_t dims[GPG_MAX_D];
    T step[GPG_MAX_D];
    int64_t j0;
    __device__ __forceinline__ void load(int64_t j, T *z) const {
        if (Xs) {
#pragma unroll
            for (int k = 0; k < D; ++k) z[k] = Xs[j * D + k];
        } else {
            int64_t f = j0 + j;
#pragma unroll
            for (int k = D - 1; k >= 0; --k) {
                z[k] = T(f % dims[k]) * step[k];
                f /= dims[k];
            }
        }
    }
};

template <typename T> struct Vec4 { T v[4]; };

template <typename T> __device__ __forceinline__ void store4(T *dst, const T *v, bool vec_ok, int n_valid) {
    if (vec_ok && n_valid == 4) {
        if (sizeof(T) == 4) {
            *reinterpret_cast<float4 *>(dst) = *reinterpret_cast<const float4 *>(v);
        } else {
            reinterpret_cast<double2 *>(dst)[0] = reinterpret_cast<const double2 *>(v)[0];
            reinterpret_cast<double2 *>(dst)[1] = reinterpret_cast<const double2 *>(v)[1];
        }
    } else {
        for (int e = 0; e < n_valid; ++e) dst[e] = v[e];
    }
}

// Optional fp16 hi/lo planes of the matrix being assembled (operand of the tensor-core factorisation).
struct KmatSplit {
    __half *hi = nullptr, *lo = nullptr;
    int64_t ld = 0;
    const float *scale = nullptr;
};

// out[i*ld + j] = k(X_i, Z_j) (+ diag_add on i == j when sym).  Block: 64 x 4 threads, each
// thread 4 rows x 4 consecutive columns -> tile of 16 rows x 256 columns.
template <typename T, int KID, int D>
__global__ void __launch_bounds__(256) kmat_kernel(const T *__restrict__ theta, const T *__restrict__ X, int64_t N,
                                                   const T *__restrict__ Z, int64_t P, int sym, T jitter,
                                                   int lower_only, T *__restrict__ out, int64_t ld, KmatSplit sp) {
    const int64_t j = ((int64_t)blockIdx.x * 64 + threadIdx.x) * 4;
    const int64_t i0 = ((int64_t)blockIdx.y * 4 + threadIdx.y) * 4;
    if (lower_only && (int64_t)blockIdx.x * 256 > (int64_t)blockIdx.y * 16 + 15) return;
    if (j >= P) return;
    const Theta<T> th = load_theta<T, D>(theta);
    const T diag_add = sym ? th.noise + jitter : T(0);
    T z[4][D];
    const int nv = (int)min((int64_t)4, P - j);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < D; ++k) z[c][k] = (c < nv) ? Z[(j + c) * D + k] : T(0);
    const bool vec_ok = ((ld % 4) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int64_t i = i0 + r;
        if (i >= N) break;
        T x[D];
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = X[i * D + k];
        T v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            v[c] = cov_from_r2<T, KID>(scaled_r2<T, D>(x, z[c], th), th);
            if (sym && i == j + c) v[c] += diag_add;
        }
        store4(out + i * ld + j, v, vec_ok, nv);
        if (sizeof(T) == 4 && sp.hi) {
            const float sc = *sp.scale;
            __align__(8) __half h4[4], l4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float sv = (float)v[c] * sc;
                h4[c] = __float2half_rn(sv);
                l4[c] = __float2half_rn(sv - __half2float(h4[c]));
            }
            if (nv == 4 && (sp.ld & 3) == 0) {
                *reinterpret_cast<uint2 *>(sp.hi + i * sp.ld + j) = *reinterpret_cast<const uint2 *>(h4);
                *reinterpret_cast<uint2 *>(sp.lo + i * sp.ld + j) = *reinterpret_cast<const uint2 *>(l4);
            } else {
                for (int c = 0; c < nv; ++c) { sp.hi[i * sp.ld + j + c] = h4[c]; sp.lo[i * sp.ld + j + c] = l4[c]; }
            }
        }
    }
}

// K2: rows of the cross-kernel for test points j in [0, mc): Ks[j*ldk + i] = k(Xs_j, X_i), and
// mean[j] = sum_i Ks[j][i] * alpha[i].  One warp per test point, lanes sweep i four at a time.
// Test points with a NaN coordinate get a zero row (so the variance GEMM stays finite) and
// mean = NaN.  SPLIT: additionally/instead emit fp16 hi/lo operands scaled by `scale`
// (tensor-core path, see gemm_tc.cuh); then Ks is not written.  Ks == nullptr (non-SPLIT): rows are not
// stored at all.  y_resid != nullptr: instead of the mean, write the residual of the linear system
// (K + diag_add I) alpha = y at row j:  y_j - sum_i k(Xs_j, X_i) alpha_i - diag_add alpha_j  (test points = X).
template <typename T, int KID, int D, bool SPLIT>
__global__ void __launch_bounds__(256) kcross_mean_kernel(const T *__restrict__ theta, const T *__restrict__ X, int64_t N,
                                                          TestPoints<T, D> tp, int64_t mc, const T *__restrict__ alpha,
                                                          T *__restrict__ Ks, int64_t ldk,
                                                          __half *__restrict__ Khi, __half *__restrict__ Klo, int64_t ldh,
                                                          const float *__restrict__ scale_ptr, T *__restrict__ mean,
                                                          const T *__restrict__ y_resid = nullptr, T jitter = T(0)) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= mc) return;
    const Theta<T> th = load_theta<T, D>(theta);
    T z[D];
    tp.load(j, z);
    bool bad = false;
#pragma unroll
    for (int k = 0; k < D; ++k) bad |= (z[k] != z[k]);
    double acc = 0.0;
    const float scale = SPLIT ? *scale_ptr : 1.0f;
    const bool vec_ok = ((ldk % 4) == 0);
    // row padding [N, ldk) (and [N, ldh)) is zero-filled so K-tiles may over-read it
    const int64_t width = SPLIT ? ldh : (Ks ? ldk : N);
    for (int64_t i = (int64_t)lane * 4; i < width; i += 128) {
        T v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            T val = T(0);
            if (i + c < N && !bad) {
                T x[D];
#pragma unroll
                for (int k = 0; k < D; ++k) x[k] = X[(i + c) * D + k];
                val = cov_from_r2<T, KID>(scaled_r2<T, D>(z, x, th), th);
                acc += (double)val * (double)alpha[i + c];
            }
            v[c] = val;
        }
        const int nvalid = (int)min((int64_t)4, width - i);
        if (SPLIT) {
            __half hi[4], lo[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float s = (float)v[c] * scale;
                hi[c] = __float2half_rn(s);
                lo[c] = __float2half_rn(s - __half2float(hi[c]));
            }
            if (nvalid == 4 && (ldh % 4) == 0) {
                *reinterpret_cast<uint2 *>(Khi + j * ldh + i) = *reinterpret_cast<uint2 *>(hi);
                *reinterpret_cast<uint2 *>(Klo + j * ldh + i) = *reinterpret_cast<uint2 *>(lo);
            } else {
                for (int c = 0; c < nvalid; ++c) {
                    Khi[j * ldh + i + c] = hi[c];
                    Klo[j * ldh + i + c] = lo[c];
                }
            }
        } else if (Ks) {
            store4(Ks + j * ldk + i, v, vec_ok && ((reinterpret_cast<uintptr_t>(Ks) & 15) == 0), nvalid);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        if (y_resid) acc = (double)y_resid[j] - acc - ((double)th.noise + (double)jitter) * (double)alpha[j];
        mean[j] = bad ? T(NAN) : (T)acc;
    }
}

// K5: sd = sqrt(max(v - sum_tiles part[t][j], 0) + noise); NaN coordinates -> NaN.
template <typename T, int D>
__global__ void predict_finalize_kernel(const T *__restrict__ theta, const T *__restrict__ part, int ntiles, int64_t ldpart,
                                        TestPoints<T, D> tp, int64_t mc, T inv_scale2, T *__restrict__ sd) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= mc) return;
    double s = 0.0;
    for (int t = 0; t < ntiles; ++t) s += (double)part[(int64_t)t * ldpart + j];
    s *= (double)inv_scale2;
    T z[D];
    tp.load(j, z);
    bool bad = false;
#pragma unroll
    for (int k = 0; k < D; ++k) bad |= (z[k] != z[k]);
    const double v = (double)theta[0], noise = (double)theta[1];
    double var = v - s;
    var = (var > 0.0 ? var : 0.0) + noise;
    sd[j] = bad ? T(NAN) : (T)sqrt(var);
}
