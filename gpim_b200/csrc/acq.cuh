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

// Order-preserving 32-bit key of a candidate's value as the ranking sees it: larger key = ranks earlier, NaN on top,
// -0 == +0.  Values are looked at in single precision (monotone, so a key range is a superset of a value range).
__device__ __forceinline__ unsigned rank_key(double v) {
    if (v != v) return 0xffffffffu;
    float f = (float)v;
    if (f == 0.f) f = 0.f;                                  // -0 -> +0
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
constexpr int TOPK_BINS = 2048;                             // 11 key bits per histogram level

// Top-k pre-filter (the arg-sort of boptim.py:303-315 only ever uses the first batch_size entries): a two-level
// histogram of the keys locates, without sorting anything, a key threshold that keeps the k best plus at most the
// population of ONE bin of 2^-22 of the key space; only those survivors enter the tournament below.
//   level 1 (fused into the sweep): hist1[key >> 21]
//   level 2: hist2[(key >> 10) & 2047] over the candidates of the level-1 bin in which the count from the top reaches k
// sel[0] = level-1 threshold bin, sel[1] = candidates above it, sel[2] = level-2 threshold bin.
__device__ __forceinline__ void topk_find_bin(const unsigned *__restrict__ hist, int need, int *bin_out, int *above_out,
                                              int *scratch /* >= 33 ints of shared memory */) {
    // 1024 threads: thread t owns bins 2 t, 2 t + 1; suffix sums from the top by warp shuffles + one shared pass
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int c0 = (int)hist[2 * t], c1 = (int)hist[2 * t + 1];
    int s = c0 + c1;                                        // inclusive suffix sum over threads t .. 1023
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_down_sync(0xffffffffu, s, o);
        if (lane + o < 32) s += v;
    }
    if (lane == 0) scratch[warp] = s;
    __syncthreads();
    if (warp == 0) {
        int w = scratch[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_down_sync(0xffffffffu, w, o);
            if (lane + o < 32) w += v;
        }
        scratch[lane] = w - scratch[lane];                  // exclusive: warps above this one
    }
    if (t == 0) scratch[32] = -1;
    __syncthreads();
    const int above_t = s - (c0 + c1) + scratch[warp];      // candidates in bins above 2 t + 1
    // the threshold bin: the highest bin b with (count in bins >= b) >= need
    if (above_t < need && above_t + c1 >= need) { *bin_out = 2 * t + 1; *above_out = above_t; scratch[32] = 1; }
    else if (above_t + c1 < need && above_t + c1 + c0 >= need) { *bin_out = 2 * t; *above_out = above_t + c1; scratch[32] = 1; }
    __syncthreads();
    if (t == 0 && scratch[32] < 0) { *bin_out = 0; *above_out = 0; }     // fewer than `need` candidates: keep all
    __syncthreads();
}

__global__ void __launch_bounds__(1024) topk_level1_kernel(const unsigned *__restrict__ hist1, int k, int *__restrict__ sel) {
    __shared__ int scratch[33];
    topk_find_bin(hist1, k, sel, sel + 1, scratch);
}

template <typename T>
__global__ void __launch_bounds__(256) topk_hist2_kernel(const Cand<T> *__restrict__ cand, int64_t M, const int *__restrict__ sel,
                                                         unsigned *__restrict__ hist2) {
    __shared__ unsigned h[TOPK_BINS];
    for (int i = threadIdx.x; i < TOPK_BINS; i += 256) h[i] = 0;
    __syncthreads();
    const unsigned b1 = (unsigned)sel[0];
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < M; j += (int64_t)gridDim.x * 256) {
        const Cand<T> c = cand[j];
        if (c.idx < 0) continue;
        const unsigned key = rank_key((double)c.val);
        if ((key >> 21) == b1) atomicAdd(&h[(key >> 10) & (TOPK_BINS - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TOPK_BINS; i += 256)
        if (h[i]) atomicAdd(&hist2[i], h[i]);
}

__global__ void __launch_bounds__(1024) topk_level2_kernel(const unsigned *__restrict__ hist2, int k, int *__restrict__ sel,
                                                           int *__restrict__ n_surv) {
    __shared__ int scratch[33];
    __shared__ int above2;
    topk_find_bin(hist2, k - sel[1], sel + 2, &above2, scratch);
    if (threadIdx.x == 0) *n_surv = 0;
}

// survivors: key above the level-1 bin, or inside it at or above the level-2 bin
template <typename T>
__global__ void __launch_bounds__(256) topk_compact_kernel(const Cand<T> *__restrict__ cand, int64_t M, const int *__restrict__ sel,
                                                           Cand<T> *__restrict__ surv, int *__restrict__ n_surv) {
    const unsigned b1 = (unsigned)sel[0], b2 = (unsigned)sel[2];
    for (int64_t j0 = (int64_t)blockIdx.x * 256; j0 < M; j0 += (int64_t)gridDim.x * 256) {
        const int64_t j = j0 + threadIdx.x;
        bool keep = false;
        Cand<T> c;
        if (j < M) {
            c = cand[j];
            if (c.idx >= 0) {
                const unsigned key = rank_key((double)c.val);
                const unsigned d1 = key >> 21;
                keep = d1 > b1 || (d1 == b1 && ((key >> 10) & (TOPK_BINS - 1)) >= b2);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m) {
            int base = 0;
            const int lane = threadIdx.x & 31;
            if (lane == 0) base = atomicAdd(n_surv, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) surv[base + __popc(m & ((1u << lane) - 1))] = c;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) acq_eval_kernel(int acq_id, const T *__restrict__ mean, const T *__restrict__ sd,
                                                       const T *__restrict__ mask, int64_t M, int64_t idx_offset,
                                                       double mu_best, double xi,
                                                       double alpha, double beta, T *__restrict__ acq_out,
                                                       Cand<T> *__restrict__ cand, unsigned *__restrict__ hist1) {
    __shared__ unsigned h[TOPK_BINS];
    if (hist1) {
        for (int i = threadIdx.x; i < TOPK_BINS; i += 256) h[i] = 0;
        __syncthreads();
    }
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < M; j += (int64_t)gridDim.x * 256) {
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
    if (hist1 && c.idx >= 0) atomicAdd(&h[rank_key((double)c.val) >> 21], 1u);
    }
    if (hist1) {
        __syncthreads();
        for (int i = threadIdx.x; i < TOPK_BINS; i += 256)
            if (h[i]) atomicAdd(&hist1[i], h[i]);
    }
}

// One tournament round: each CTA bitonic-sorts CH = 2048 candidates in shared memory and keeps
// its best k (k <= 1024), so every round shrinks the field by >= 2x.
// n_in (device, optional): the number of candidates actually present (<= n, the host's bound the grid was sized for);
// a value < 0 means "-n_in entries, already ranked": the round only copies them.  n_out (device): what the next round
// will find, in the same convention.
template <typename T>
__global__ void __launch_bounds__(1024) topk_round_kernel(const Cand<T> *__restrict__ in, int64_t n, int k,
                                                          Cand<T> *__restrict__ out, const int *__restrict__ n_in,
                                                          int *__restrict__ n_out) {
    constexpr int CH = 2048;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cand<T> *s = reinterpret_cast<Cand<T> *>(smem_raw);
    if (n_in) {
        const int nd = *n_in;
        if (nd < 0) {                                       // final already: pass the k ranked entries on
            if (blockIdx.x == 0) {
                for (int e = threadIdx.x; e < k; e += 1024) out[e] = in[e];
                if (threadIdx.x == 0 && n_out) *n_out = nd;
            }
            return;
        }
        n = nd;
        const int64_t nblk = n > 0 ? (n + CH - 1) / CH : 1;
        if (blockIdx.x >= nblk) return;
        if (threadIdx.x == 0 && blockIdx.x == 0 && n_out) *n_out = nblk == 1 ? -k : (int)(nblk * k);
    }
    const int64_t base = (int64_t)blockIdx.x * CH;
    // sort no more than this CTA holds: the next power of two above max(count, k)
    const int cnt = (int)max((int64_t)0, min((int64_t)CH, n - base));
    int sz = 2;
    while (sz < cnt || sz < k) sz <<= 1;
    for (int e = threadIdx.x; e < sz; e += 1024) {
        Cand<T> c;
        c.val = T(0); c.idx = -1;
        if (base + e < n) c = in[base + e];
        s[e] = c;
    }
    __syncthreads();
    for (int size = 2; size <= sz; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int e = threadIdx.x; e < sz / 2; e += 1024) {
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

// ---------------------------------------------------------------------------------------------
// The point filters of the BO loop on the ranked candidate list (boptim.py:326-429; the reference walks Python
// lists and a cKDTree).  One CTA, one thread per candidate (k <= 1024).
//   admissible(t): candidate t was not measured before and keeps the distance dscale * gamma^q from the q-th most
//                  recent measured point (q < memory)                                     -- checkvalues
//   first        = the first admissible candidate, -1 when the list is exhausted (the caller applies the
//                  reference's exit strategy, which draws from numpy's generator)
//   batch mode   : the list is cut at the first candidate whose VALUE equals that of `first`
//                  (np.where(vals == val)[0][0]); greedy ball suppression from there: take the best candidate
//                  left (= the earliest: the list is ranked), drop everything within batch_dscale of it (closed
//                  ball, as cKDTree.query_ball_point), until batch_out_max picks or nothing is left -- update_points
// Coordinates are the grid indices of the flat index (row-major over dims), distances exact in integers.
// sel_out int32[4 + batch_out_max]: {first, start, npicks, nan_seen, picks...} (positions in the candidate list).
// ---------------------------------------------------------------------------------------------
struct SelectArgs {
    int ndim;
    long long dims[4];
    int n_visited, memory, do_batch, batch_out_max;
    double dscale, gamma, batch_dscale;
};

template <typename T>
__global__ void __launch_bounds__(1024) acq_select_kernel(const T *__restrict__ val, const int64_t *__restrict__ idx,
                                                          const int32_t *__restrict__ count, int kmax,
                                                          const int64_t *__restrict__ visited, SelectArgs a,
                                                          int32_t *__restrict__ sel_out) {
    __shared__ int s_first, s_start, s_next, s_nan;
    __shared__ long long s_pick[4];
    const int t = threadIdx.x;
    const int n = min(*count, kmax);
    if (t == 0) { s_first = 0x7fffffff; s_start = 0x7fffffff; s_nan = 0; }
    __syncthreads();
    long long c[4] = {0, 0, 0, 0};
    bool ok = false;
    T v = T(0);
    if (t < n) {
        v = val[t];
        if (v != v) atomicOr(&s_nan, 1);
        long long f = idx[t];
        for (int q = a.ndim - 1; q >= 0; --q) { c[q] = f % a.dims[q]; f /= a.dims[q]; }
        ok = true;
        for (int q = 0; q < a.n_visited && ok; ++q) ok = visited[q] != idx[t];
        double lim = a.dscale;
        for (int q = 0; q < a.memory && q < a.n_visited && ok; ++q) {     // q-th most recent measured point
            long long f2 = visited[a.n_visited - 1 - q], d2 = 0;
            for (int e = a.ndim - 1; e >= 0; --e) { const long long ce = f2 % a.dims[e]; f2 /= a.dims[e]; d2 += (c[e] - ce) * (c[e] - ce); }
            ok = sqrt((double)d2) > lim;
            lim *= a.gamma;
        }
        if (ok) atomicMin(&s_first, t);
    }
    __syncthreads();
    const int first = s_first == 0x7fffffff ? -1 : s_first;
    int npicks = 0;
    if (a.do_batch && first >= 0) {
        const T v0 = val[first];
        if (t < n && v == v0) atomicMin(&s_start, t);
        __syncthreads();
        const int start = s_start;
        bool alive = t < n && t >= start;
        const double r2 = a.batch_dscale * a.batch_dscale;
        while (npicks < a.batch_out_max) {
            if (t == 0) s_next = 0x7fffffff;
            __syncthreads();
            if (alive) atomicMin(&s_next, t);
            __syncthreads();
            const int b = s_next;
            if (b == 0x7fffffff) break;
            if (t == b) {
                for (int e = 0; e < 4; ++e) s_pick[e] = c[e];
                sel_out[4 + npicks] = b;
            }
            __syncthreads();
            if (alive) {
                long long d2 = 0;
                for (int e = 0; e < a.ndim; ++e) d2 += (c[e] - s_pick[e]) * (c[e] - s_pick[e]);
                if ((double)d2 <= r2) alive = false;
            }
            ++npicks;
            __syncthreads();
        }
    }
    if (t == 0) { sel_out[0] = first; sel_out[1] = a.do_batch && first >= 0 ? s_start : -1; sel_out[2] = npicks; sel_out[3] = s_nan; }
}
