// acq.cuh -- K6: acquisition sweep over the dense grid + top-k
// (acqfunc.py:11-92 evaluated on the host in the reference; boptim.py:303-315 full argsort).
#pragma once
#include "common.cuh"

template <typename T> struct Cand { T val; int64_t idx; };

// a ranks before b: larger value first; NaN first of all (np.argsort puts NaN last ascending, the
// reference then reverses, boptim.py:304-306); ties -> larger flat index first; idx < 0 = excluded.
template <typename T> __device__ __forceinline__ bool ranks_before(const Cand<T> &a, const Cand<T> &b) {
    if (a.idx < 0) return false;
    if (b.idx < 0) return true;
    const bool an = a.val != a.val, bn = b.val != b.val;
    if (an || bn) return an && (!bn || a.idx > b.idx);
    if (a.val != b.val) return a.val > b.val;
    return a.idx > b.idx;
}

template <typename T>
__global__ void __launch_bounds__(256) acq_eval_kernel(int acq_id, const T *__restrict__ mean, const T *__restrict__ sd,
                                                       const T *__restrict__ mask, int64_t M, int64_t idx_offset,
                                                       double mu_best, double xi,
                                                       double alpha, double beta, T *__restrict__ acq_out,
                                                       Cand<T> *__restrict__ cand) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (j >= M) return;
    const double mu = (double)mean[j], s = (double)sd[j];
    double a;
    if (acq_id == GPG_ACQ_CB) {
        a = alpha * mu + beta * s;
    } else {
        const double imp = mu - mu_best - xi;
        const double z = imp / s;
        const double cdf = 0.5 * erfc(-z * 0.70710678118654752440);
        if (acq_id == GPG_ACQ_EI) {
            const double pdf = 0.39894228040143267794 * exp(-0.5 * z * z);
            a = imp * cdf + s * pdf;
        } else a = cdf;
    }
    Cand<T> c;
    c.idx = j + idx_offset;
    if (mask) {
        a = (double)mask[j] * a;
        if (a != a) c.idx = -1;                 // masked-out entries are stripped (boptim.py:311)
    }
    c.val = (T)a;
    if (acq_out) acq_out[j] = (T)a;
    cand[j] = c;
}

// One tournament round: each CTA bitonic-sorts CH = 2048 candidates in shared memory and keeps
// its best k (k <= 1024), so every round shrinks the field by >= 2x.
template <typename T>
__global__ void __launch_bounds__(1024) topk_round_kernel(const Cand<T> *__restrict__ in, int64_t n, int k,
                                                          Cand<T> *__restrict__ out) {
    constexpr int CH = 2048;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cand<T> *s = reinterpret_cast<Cand<T> *>(smem_raw);
    const int64_t base = (int64_t)blockIdx.x * CH;
    for (int e = threadIdx.x; e < CH; e += 1024) {
        Cand<T> c;
        c.val = T(0); c.idx = -1;
        if (base + e < n) c = in[base + e];
        s[e] = c;
    }
    __syncthreads();
    for (int size = 2; size <= CH; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int e = threadIdx.x; e < CH / 2; e += 1024) {
                const int lo = (e / stride) * (stride * 2) + (e % stride);
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);          // "descending" = best first
                const Cand<T> a = s[lo], b = s[hi];
                const bool swap = desc ? ranks_before(b, a) : ranks_before(a, b);
                if (swap) { s[lo] = b; s[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int e = threadIdx.x; e < k; e += 1024) out[(int64_t)blockIdx.x * k + e] = s[e];
}

template <typename T>
__global__ void topk_emit_kernel(const Cand<T> *__restrict__ in, int k, T *__restrict__ val, int64_t *__restrict__ idx,
                                 int32_t *__restrict__ count) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int local = 0;
    for (int e = threadIdx.x; e < k; e += blockDim.x) {
        val[e] = in[e].val;
        idx[e] = in[e].idx;
        if (in[e].idx >= 0) local++;
    }
    atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) *count = cnt;
}
