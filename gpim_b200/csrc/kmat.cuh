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
#pragma unroll
        for (int e = 0; e < 4; ++e)                // predicated, fully unrolled: v must stay in registers
            if (e < n_valid) dst[e] = v[e];
    }
}

// Optional fp16 hi/lo planes of the matrix being assembled (operand of the tensor-core factorisation).
struct KmatSplit {
    __half *hi = nullptr, *lo = nullptr;
    int64_t ld = 0;
    const float *scale = nullptr;
    int64_t ncols = 0;           // > 0: only columns j < ncols are emitted
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
        if (sizeof(T) == 4 && sp.hi && (sp.ncols == 0 || j < sp.ncols)) {
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
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < nv) { sp.hi[i * sp.ld + j + c] = h4[c]; sp.lo[i * sp.ld + j + c] = l4[c]; }
            }
        }
    }
}

// Squared lengthscale-scaled distance beyond which k(r2) < rel * variance (compact support of K* at fp32 resolution).
template <typename T, int KID>
__device__ __forceinline__ float support_r2(const Theta<T> &th, float rel) {
    const float lt = -logf(rel);                         // > 0
    if (KID == GPG_RBF) return 2.0f * lt;
    if (KID == GPG_MATERN52) {                           // (1 + s + s^2 / 3) exp(-s) = rel, s = sqrt(5) r: Newton on the log
        float s = lt + 8.0f;
#pragma unroll
        for (int it = 0; it < 6; ++it) {
            const float q = 1.0f + s + s * s * (1.0f / 3.0f);
            const float f = logf(q) - s + lt, df = (1.0f + s * (2.0f / 3.0f)) / q - 1.0f;
            s -= f / df;
        }
        return s * s * 0.2f;
    }
    const float a = (float)th.alpha;                     // (1 + r2 / (2 a))^-a = rel
    const float e = lt / a;
    return e > 80.0f ? 3.0e38f : 2.0f * a * (expf(e) - 1.0f);
}

// Bounding boxes of the training rows in blocks of 32 consecutive rows: bbox[b][0..D) = min, [D..2D) = max.
template <typename T, int D>
__global__ void __launch_bounds__(256) block_bbox_kernel(const T *__restrict__ X, int64_t N, float *__restrict__ bbox) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b * 32 >= N) return;
    const int64_t i = b * 32 + lane;
    float lo[D], hi[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const float v = i < N ? (float)X[i * D + k] : 0.f;
        lo[k] = i < N ? v : 3.0e38f;
        hi[k] = i < N ? v : -3.0e38f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) { bbox[b * 2 * D + k] = lo[k]; bbox[b * 2 * D + D + k] = hi[k]; }
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
                                                          const T *__restrict__ y_resid = nullptr, T jitter = T(0),
                                                          int *__restrict__ krange = nullptr, float support_rel = 0.f,
                                                          const float *__restrict__ bbox = nullptr, int nblk32 = 0,
                                                          float var_target = 0.f) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= mc) return;
    const Theta<T> th = load_theta<T, D>(theta);
    T z[D];
    tp.load(j, z);
    bool bad = false;
#pragma unroll
    for (int k = 0; k < D; ++k) bad |= (z[k] != z[k]);
    // mean accumulator: double for T = double; for T = float a Kahan-compensated fp32 pair, which keeps the
    // long alternating sum accurate without per-element conversions (they run on the slow XU pipe)
    double accd = 0.0;
    float acc_s = 0.f, acc_c = 0.f;
    const float scale = SPLIT ? *scale_ptr : 1.0f;
    // row padding [N, ldk) (and [N, ldh)) is zero-filled so K-tiles may over-read it
    const int64_t width = SPLIT ? ldh : (Ks ? ldk : N);
    // Compact support (optional, SPLIT only): the contiguous range of training rows that can have a covariance above
    // support_rel * variance with ANY of the 128 test points of this row's tile -- found geometrically, from the
    // tile's bounding box against the boxes of 32-row blocks of X (conservative: box distance <= point distance).
    // Everything outside contributes below fp32 resolution; it is neither evaluated nor written nor read by the
    // variance GEMM (krange), and the warps of a tile all derive the same range.
    // Two ranges (SPLIT): [m_lo, m_hi) for the MEAN (linear in K*, weights alpha ~ y / noise: threshold support_rel) and
    // the narrower [c_lo, c_hi) for the planes the VARIANCE product reads.  Dropping entries below eps x variance changes
    // q = L^-1 k* by at most ||L^-1||_2 eps v sqrt(N) and sum q^2 <= v by at most 2 sqrt(v) ||dq||, i.e. the variance by
    // at most 2 eps sqrt(v N / noise) relative to v: eps = var_target / (2 sqrt(v N / noise)) keeps it below var_target
    // (2e-7, the rounding of the fp32 result itself) whatever the data.
    int64_t c_lo = 0, c_hi = width, m_lo = 0, m_hi = width;
    if (!SPLIT && bbox != nullptr && sizeof(T) == 4 && !bad) {
        // no tile to agree with (nothing is stored for a GEMM): the range of THIS point, box = the point itself
        const float r2max = 1.02f * support_r2<T, KID>(th, support_rel);
        int first = 0x7fffffff, last = 0;
        for (int b = lane; b < nblk32; b += 32) {
            float d2 = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float zk = (float)z[k];
                const float gap = fmaxf(0.f, fmaxf(bbox[b * 2 * D + k] - zk, zk - bbox[b * 2 * D + D + k])) * (float)th.inv_ls[k];
                d2 += gap * gap;
            }
            if (d2 < r2max) { first = min(first, b); last = max(last, b + 1); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
        }
        if (last <= first) { c_lo = 0; c_hi = 0; }
        else { c_lo = ((int64_t)first * 32 / 128) * 128; c_hi = min(width, (((int64_t)last * 32 + 127) / 128) * 128); }
        m_lo = c_lo; m_hi = c_hi;
    }
    if (krange && bbox) {
        const int64_t tile0 = (j / 128) * 128;
        float blo[D], bhi[D];
#pragma unroll
        for (int k = 0; k < D; ++k) { blo[k] = 3.0e38f; bhi[k] = -3.0e38f; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t jj = tile0 + lane + 32 * q;
            if (jj < mc) {
                T zz[D];
                tp.load(jj, zz);
                bool nanp = false;
#pragma unroll
                for (int k = 0; k < D; ++k) nanp |= (zz[k] != zz[k]);
                if (!nanp) {
#pragma unroll
                    for (int k = 0; k < D; ++k) { blo[k] = fminf(blo[k], (float)zz[k]); bhi[k] = fmaxf(bhi[k], (float)zz[k]); }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < D; ++k)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                blo[k] = fminf(blo[k], __shfl_xor_sync(0xffffffffu, blo[k], o));
                bhi[k] = fmaxf(bhi[k], __shfl_xor_sync(0xffffffffu, bhi[k], o));
            }
        const float r2max = 1.02f * support_r2<T, KID>(th, support_rel);
        float rel_var = support_rel;
        if (var_target > 0.f)
            rel_var = fminf(1e-6f, fmaxf(support_rel, var_target / (2.0f * sqrtf(fmaxf((float)th.variance, 1e-30f) * (float)N /
                                                                                    fmaxf((float)th.noise, 1e-30f)))));
        const float r2var = 1.02f * support_r2<T, KID>(th, rel_var);
        int first = 0x7fffffff, last = 0, vfirst = 0x7fffffff, vlast = 0;
        for (int b = lane; b < nblk32; b += 32) {
            float d2 = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float gap = fmaxf(0.f, fmaxf(bbox[b * 2 * D + k] - bhi[k], blo[k] - bbox[b * 2 * D + D + k])) * (float)th.inv_ls[k];
                d2 += gap * gap;
            }
            if (d2 < r2max) { first = min(first, b); last = max(last, b + 1); }
            if (d2 < r2var) { vfirst = min(vfirst, b); vlast = max(vlast, b + 1); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
            vfirst = min(vfirst, __shfl_xor_sync(0xffffffffu, vfirst, o));
            vlast = max(vlast, __shfl_xor_sync(0xffffffffu, vlast, o));
        }
        if (last <= first) { m_lo = 0; m_hi = 0; }
        else {
            m_lo = ((int64_t)first * 32 / 128) * 128;
            m_hi = min(width, (((int64_t)last * 32 + 127) / 128) * 128);
        }
        if (vlast <= vfirst) { c_lo = 0; c_hi = 0; }
        else {
            c_lo = ((int64_t)vfirst * 32 / 128) * 128;
            c_hi = min(width, (((int64_t)vlast * 32 + 127) / 128) * 128);
        }
        if (lane == 0) { krange[2 * (j / 128)] = (int)c_lo; krange[2 * (j / 128) + 1] = (int)min(c_hi, N); }
    }
    __half *__restrict__ hrow = SPLIT ? Khi + j * ldh : nullptr;
    __half *__restrict__ lrow = SPLIT ? Klo + j * ldh : nullptr;
    T *__restrict__ krow = (!SPLIT && Ks) ? Ks + j * ldk : nullptr;
    const bool fast_ok = (D == 2 && sizeof(T) == 4 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(alpha) & 7) == 0 && !bad);
    // lane owns the column pairs i0 + 2 lane + {0, 1} and i0 + 64 + 2 lane + {0, 1}: coordinates, alpha and the
    // fp16 planes are all touched with unit stride across the warp (128-bit / 64-bit / 32-bit per lane)
    for (int64_t i0 = m_lo; i0 < m_hi; i0 += 128) {
        T v[4];
        if (fast_ok && i0 + 128 <= N) {
#pragma unroll
            for (int hblk = 0; hblk < 2; ++hblk) {
                const int64_t i = i0 + 64 * hblk + 2 * lane;
                const float4 xy = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(X) + i * 2);
                const float2 a2 = *reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(alpha) + i);
                T xa[D], xb[D];
                xa[0] = (T)xy.x; xa[D - 1] = (T)xy.y; xb[0] = (T)xy.z; xb[D - 1] = (T)xy.w;
                v[2 * hblk] = cov_from_r2<T, KID>(scaled_r2<T, D>(z, xa, th), th);
                v[2 * hblk + 1] = cov_from_r2<T, KID>(scaled_r2<T, D>(z, xb, th), th);
                const float av[2] = {a2.x, a2.y};
#pragma unroll
                for (int e = 0; e < 2; ++e) {            // Kahan: no FMA contraction may touch these
                    const float yk = __fsub_rn(__fmul_rn((float)v[2 * hblk + e], av[e]), acc_c);
                    const float t = __fadd_rn(acc_s, yk);
                    acc_c = __fsub_rn(__fsub_rn(t, acc_s), yk);
                    acc_s = t;
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t i = i0 + 64 * (c >> 1) + 2 * lane + (c & 1);
                T val = T(0);
                if (i < N && !bad) {
                    T x[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) x[k] = X[i * D + k];
                    val = cov_from_r2<T, KID>(scaled_r2<T, D>(z, x, th), th);
                    if (sizeof(T) == 4) {
                        const float yk = __fsub_rn(__fmul_rn((float)val, (float)alpha[i]), acc_c);
                        const float t = __fadd_rn(acc_s, yk);
                        acc_c = __fsub_rn(__fsub_rn(t, acc_s), yk);
                        acc_s = t;
                    } else {
                        accd += (double)val * (double)alpha[i];
                    }
                }
                v[c] = val;
            }
        }
        if (SPLIT && i0 >= c_lo && i0 < c_hi) {
#pragma unroll
            for (int hblk = 0; hblk < 2; ++hblk) {
                const int64_t i = i0 + 64 * hblk + 2 * lane;
                const float s0 = (float)v[2 * hblk] * scale, s1 = (float)v[2 * hblk + 1] * scale;
                const __half2 h2 = __floats2half2_rn(s0, s1);
                const float2 f2 = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(s0 - f2.x, s1 - f2.y);
                if (i + 1 < width) {                      // width is even: pairs are in or out together
                    *reinterpret_cast<__half2 *>(hrow + i) = h2;
                    *reinterpret_cast<__half2 *>(lrow + i) = l2;
                }
            }
        } else if (krow) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t i = i0 + 64 * (c >> 1) + 2 * lane + (c & 1);
                if (i < width) krow[i] = v[c];
            }
        }
    }
    if (sizeof(T) == 4) acc_c = -acc_c;               // Kahan keeps the NEGATIVE of the running error
    double acc = (sizeof(T) == 4) ? (double)acc_s + (double)acc_c : accd;
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
