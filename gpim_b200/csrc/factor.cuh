// factor.cuh -- K3 blocked Cholesky, triangular inverse, and the vector solves (K7a).
//
// Cholesky: right-looking, block NB = GemmCfg<T>::BN.  Per block column: (1) one CTA factors the
// NB x NB diagonal block in shared memory and inverts it, (2) the panel below becomes
// A21 * inv(L11)^T (a GEMM against the small inverse), (3) the trailing matrix gets the SYRK
// update A22 -= A21 A21^T on lower tiles only.  Steps (2)/(3) go through gemm_dispatch(), i.e.
// tcgen05 when eligible, SIMT otherwise.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"

template <typename T> int gemm_dispatch(gpg_handle_s *h, const GemmArgs<T> &g, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// diagonal block: Cholesky + inverse in shared memory.  grid = 1 CTA (potf2) or one CTA per
// diagonal block (inverse-only mode, used by trtri).  S, W: NB x (NB+1).
// ---------------------------------------------------------------------------------------------
template <typename T, int NB>
__global__ void __launch_bounds__(256) diag_block_kernel(T *__restrict__ A, int64_t ld, int64_t N, int64_t j0_first,
                                                         int do_factor, T *__restrict__ inv_out, int64_t ld_inv,
                                                         int64_t inv_block_stride, int dense_out,
                                                         int32_t *__restrict__ info) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *S = reinterpret_cast<T *>(smem_raw);
    T *W = S + NB * (NB + 1);
    constexpr int LDS = NB + 1;
    const int t = threadIdx.x;
    const int64_t j0 = j0_first + (int64_t)blockIdx.x * NB;
    const int nb = (int)min((int64_t)NB, N - j0);
    T *Ab = A + j0 * ld + j0;
    for (int idx = t; idx < NB * NB; idx += 256) {
        const int i = idx / NB, k = idx % NB;
        T v = (i == k) ? T(1) : T(0);
        if (i < nb && k <= i) v = Ab[(int64_t)i * ld + k];
        S[i * LDS + k] = v;
    }
    __syncthreads();
    if (do_factor) {
        const int tx = t & 15, ty = t >> 4;
        for (int j = 0; j < nb; ++j) {
            __syncthreads();
            const T d = S[j * LDS + j];
            if (!(d > T(0)) && t == 0) atomicCAS(info, 0, (int32_t)(j0 + j + 1));
            const T ljj = gpg_sqrt(d);
            const T inv = T(1) / ljj;
            __syncthreads();
            for (int i = j + 1 + t; i < nb; i += 256) S[i * LDS + j] *= inv;
            if (t == 0) S[j * LDS + j] = ljj;
            __syncthreads();
            for (int i = j + 1 + ty; i < nb; i += 16) {
                const T lij = S[i * LDS + j];
                for (int k = j + 1 + tx; k <= i; k += 16) S[i * LDS + k] -= lij * S[k * LDS + j];
            }
        }
        __syncthreads();
        for (int idx = t; idx < nb * nb; idx += 256) {
            const int i = idx / nb, k = idx % nb;
            if (k <= i) Ab[(int64_t)i * ld + k] = S[i * LDS + k];
        }
    }
    // inverse of the lower-triangular block, one thread per column
    if (t < NB) {
        const int c = t;
        for (int i = 0; i < NB; ++i) {
            T s = T(0);
            if (i >= c) {
                s = (i == c) ? T(1) : T(0);
                for (int k = c; k < i; ++k) s -= S[i * LDS + k] * W[k * LDS + c];
                s /= S[i * LDS + i];
            }
            W[i * LDS + c] = s;
        }
    }
    __syncthreads();
    T *out = inv_out + (int64_t)blockIdx.x * inv_block_stride;
    for (int idx = t; idx < NB * NB; idx += 256) {
        const int i = idx / NB, k = idx % NB;
        // dense NB x NB scratch block, or in-place block of Linv (guard the ragged last block)
        if (dense_out || (i < nb && k < nb)) out[(int64_t)i * ld_inv + k] = W[i * LDS + k];
    }
}

template <typename T, int NB> static int diag_block_smem() { return 2 * NB * (NB + 1) * (int)sizeof(T); }

template <typename T>
static int cholesky_blocked(gpg_handle_s *h, T *A, int64_t N, int64_t ld, int32_t *info, int reset_info, T *dinv,
                            cudaStream_t stream) {   // dinv: NB*NB scratch
    constexpr int NB = GemmCfg<T>::BN;
    static bool attr_set = false;
    if (!attr_set) {
        GPG_CUDA_CHECK(cudaFuncSetAttribute(diag_block_kernel<T, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            diag_block_smem<T, NB>()));
        attr_set = true;
    }
    if (reset_info) GPG_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), stream));
    for (int64_t j0 = 0; j0 < N; j0 += NB) {
        const int64_t nb = std::min<int64_t>(NB, N - j0);
        diag_block_kernel<T, NB><<<1, 256, diag_block_smem<T, NB>(), stream>>>(A, ld, N, j0, 1, dinv, NB, 0, 1, info);
        GPG_LAUNCH_CHECK(h);
        const int64_t rows = N - j0 - nb;
        if (rows <= 0) break;
        T *A21 = A + (j0 + nb) * ld + j0;
        GemmArgs<T> p;                       // A21 <- A21 * inv(L11)^T   (in place: one n-tile)
        p.A = A21; p.lda = ld; p.a_kmajor = 1;
        p.B = dinv; p.ldb = NB; p.b_kmajor = 1;
        p.C = A21; p.ldc = ld;
        p.M = (int)rows; p.N = (int)nb; p.K = (int)nb;
        p.ke_mode = GEMM_KE_N;
        GPG_TRY(gemm_simt<T>(h, p, stream));
        GemmArgs<T> s;                       // A22 -= A21 A21^T (lower tiles)
        s.A = A21; s.lda = ld; s.B = A21; s.ldb = ld;
        s.C = A + (j0 + nb) * ld + (j0 + nb); s.ldc = ld;
        s.M = (int)rows; s.N = (int)rows; s.K = (int)nb;
        s.alpha = T(-1); s.beta = T(1);
        s.tile_mode = GEMM_TILES_LOWER;
        GPG_TRY(gemm_dispatch<T>(h, s, stream));
    }
    return GPG_OK;
}

// ---------------------------------------------------------------------------------------------
// Linv = L^-1, recursive doubling over block size: with L = [[L11,0],[L21,L22]],
// W21 = -W22 (L21 W11).  All pairs of a level are independent -> batched GEMMs.
// tmp: N x N scratch (same ld as L).
// ---------------------------------------------------------------------------------------------
template <typename T>
static int trtri_blocked(gpg_handle_s *h, const T *L, int64_t N, int64_t ld, T *Linv, int64_t ldi, T *tmp,
                         cudaStream_t stream) {
    constexpr int NB = GemmCfg<T>::BN;
    static bool attr_set = false;
    if (!attr_set) {
        GPG_CUDA_CHECK(cudaFuncSetAttribute(diag_block_kernel<T, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            diag_block_smem<T, NB>()));
        attr_set = true;
    }
    GPG_CUDA_CHECK(cudaMemset2DAsync(Linv, ldi * sizeof(T), 0, N * sizeof(T), N, stream));
    const int nblk = (int)((N + NB - 1) / NB);
    diag_block_kernel<T, NB><<<nblk, 256, diag_block_smem<T, NB>(), stream>>>(
        const_cast<T *>(L), ld, N, 0, 0, Linv, ldi, (int64_t)NB * (ldi + 1), 0, nullptr);
    GPG_LAUNCH_CHECK(h);
    for (int64_t b = NB; b < N; b *= 2) {
        const int64_t npairs_full = N / (2 * b);                   // pairs whose second block is complete
        const int64_t rem_start = npairs_full * 2 * b;
        const int64_t rem_rows = (N - rem_start > b) ? (N - rem_start - b) : 0;   // ragged last pair
        for (int pass = 0; pass < 2; ++pass) {
            const int64_t s0 = pass == 0 ? 0 : rem_start;
            const int64_t rows = pass == 0 ? b : rem_rows;
            const int64_t batch = pass == 0 ? npairs_full : (rem_rows > 0 ? 1 : 0);
            if (batch == 0 || rows == 0) continue;
            GemmArgs<T> g1;                  // tmp21 = L21 * W11
            g1.A = L + (s0 + b) * ld + s0; g1.lda = ld; g1.a_kmajor = 1;
            g1.B = Linv + s0 * ldi + s0; g1.ldb = ldi; g1.b_kmajor = 0;
            g1.C = tmp + (s0 + b) * ld + s0; g1.ldc = ld;
            g1.M = (int)rows; g1.N = (int)b; g1.K = (int)b;
            g1.kb_mode = GEMM_KB_N0;
            g1.batch = (int)batch;
            g1.strideA = 2 * b * (ld + 1); g1.strideB = 2 * b * (ldi + 1); g1.strideC = 2 * b * (ld + 1);
            GPG_TRY(gemm_dispatch<T>(h, g1, stream));
            GemmArgs<T> g2;                  // W21 = -W22 * tmp21
            g2.A = Linv + (s0 + b) * (ldi + 1); g2.lda = ldi; g2.a_kmajor = 1;
            g2.B = tmp + (s0 + b) * ld + s0; g2.ldb = ld; g2.b_kmajor = 0;
            g2.C = Linv + (s0 + b) * ldi + s0; g2.ldc = ldi;
            g2.M = (int)rows; g2.N = (int)b; g2.K = (int)rows;
            g2.ke_mode = GEMM_KE_M;
            g2.alpha = T(-1);
            g2.batch = (int)batch;
            g2.strideA = 2 * b * (ldi + 1); g2.strideB = 2 * b * (ld + 1); g2.strideC = 2 * b * (ldi + 1);
            GPG_TRY(gemm_dispatch<T>(h, g2, stream));
        }
    }
    return GPG_OK;
}

// ---------------------------------------------------------------------------------------------
// triangular GEMV: out = alpha * op(Mx) x + beta * yin, Mx lower-triangular row-major, double
// accumulation.  TRANS = false: one warp per row.  TRANS = true: 32-column strips.
// ---------------------------------------------------------------------------------------------
template <typename T, bool TRANS>
__global__ void __launch_bounds__(256) gemv_tri_kernel(const T *__restrict__ Mx, int64_t ld, int64_t N,
                                                       const T *__restrict__ x, const T *__restrict__ yin, T alpha,
                                                       T beta, T *__restrict__ out) {
    if (!TRANS) {
        const int lane = threadIdx.x & 31;
        const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
        if (i >= N) return;
        double acc = 0.0;
        for (int64_t k = lane; k <= i; k += 32) acc += (double)Mx[i * ld + k] * (double)x[k];
        acc = warp_sum(acc);
        if (lane == 0) out[i] = (T)((double)alpha * acc + (yin ? (double)beta * (double)yin[i] : 0.0));
    } else {
        __shared__ double red[8][33];
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        const int64_t k = (int64_t)blockIdx.x * 32 + tx;
        double acc = 0.0;
        if (k < N)
            for (int64_t i = (int64_t)blockIdx.x * 32 + ty; i < N; i += 8)
                if (i >= k) acc += (double)Mx[i * ld + k] * (double)x[i];
        red[ty][tx] = acc;
        __syncthreads();
        if (ty == 0 && k < N) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) s += red[r][tx];
            out[k] = (T)((double)alpha * s + (yin ? (double)beta * (double)yin[k] : 0.0));
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) solve_scalars_kernel(const T *__restrict__ L, int64_t ld, int64_t N,
                                                            const T *__restrict__ vhat, T *__restrict__ scalars) {
    __shared__ double r0[8], r1[8];
    double q = 0.0, ld_sum = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += 256) {
        const double v = (double)vhat[i];
        q += v * v;
        ld_sum += log((double)L[i * ld + i]);
    }
    q = warp_sum(q);
    ld_sum = warp_sum(ld_sum);
    if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = q; r1[threadIdx.x >> 5] = ld_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < 8; ++w) { a += r0[w]; b += r1[w]; }
        scalars[0] = (T)(0.5 * a);
        scalars[1] = (T)b;
    }
}

// vhat = L^-1 y and alpha = L^-T vhat through the explicit inverse plus two residual corrections
// against L (each correction restores the backward error of a substitution solve).
// scratch: 2*N elements.
template <typename T>
static int solve_vec_refined(gpg_handle_s *h, const T *L, const T *Linv, int64_t N, int64_t ld, const T *y, T *vhat,
                             T *alpha, T *scalars, T *scratch, cudaStream_t stream) {
    T *r = scratch, *dx = scratch + N;
    const int gN = (int)((N + 7) / 8), gT = (int)((N + 31) / 32);
    // forward: L v = y
    gemv_tri_kernel<T, false><<<gN, 256, 0, stream>>>(Linv, ld, N, y, nullptr, T(1), T(0), vhat);
    GPG_LAUNCH_CHECK(h);
    for (int it = 0; it < 2; ++it) {
        gemv_tri_kernel<T, false><<<gN, 256, 0, stream>>>(L, ld, N, vhat, y, T(-1), T(1), r);        // r = y - L v
        GPG_LAUNCH_CHECK(h);
        gemv_tri_kernel<T, false><<<gN, 256, 0, stream>>>(Linv, ld, N, r, vhat, T(1), T(1), dx);     // dx = v + Linv r
        GPG_LAUNCH_CHECK(h);
        GPG_CUDA_CHECK(cudaMemcpyAsync(vhat, dx, N * sizeof(T), cudaMemcpyDeviceToDevice, stream));
    }
    // backward: L^T a = v
    gemv_tri_kernel<T, true><<<gT, 256, 0, stream>>>(Linv, ld, N, vhat, nullptr, T(1), T(0), alpha);
    GPG_LAUNCH_CHECK(h);
    for (int it = 0; it < 2; ++it) {
        gemv_tri_kernel<T, true><<<gT, 256, 0, stream>>>(L, ld, N, alpha, vhat, T(-1), T(1), r);
        GPG_LAUNCH_CHECK(h);
        gemv_tri_kernel<T, true><<<gT, 256, 0, stream>>>(Linv, ld, N, r, alpha, T(1), T(1), dx);
        GPG_LAUNCH_CHECK(h);
        GPG_CUDA_CHECK(cudaMemcpyAsync(alpha, dx, N * sizeof(T), cudaMemcpyDeviceToDevice, stream));
    }
    if (scalars) {
        solve_scalars_kernel<T><<<1, 256, 0, stream>>>(L, ld, N, vhat, scalars);
        GPG_LAUNCH_CHECK(h);
    }
    return GPG_OK;
}
