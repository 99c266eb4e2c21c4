// factor.cuh -- K3 blocked Cholesky, triangular inverse, and the vector solves (K7a): the SIMT
// (fp64, small fp32) drivers and the kernels both they and the tensor-core drivers of factor_tc.cuh
// share (diagonal block, panel forward substitution, triangular GEMVs, refinement bookkeeping).
//
// Cholesky (this file): right-looking, block NB = GemmCfg<T>::BN.  Per block column: (1) one CTA
// factors the NB x NB diagonal block in shared memory and inverts it, (2) the panel below becomes
// A21 * inv(L11)^T (a GEMM against the small inverse), (3) the trailing matrix gets the SYRK
// update A22 -= A21 A21^T on lower tiles only, both on the SIMT GEMM.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"

template <typename T> int gemm_dispatch(gpg_handle_s *h, const GemmArgs<T> &g, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// diagonal block: Cholesky + inverse of one NB x NB block in shared memory, blocked by 32.
// grid = 1 CTA (potf2 inside the blocked Cholesky) or one CTA per diagonal block (inverse-only
// mode, level 0 of trtri).  S: the block / its factor, W: the inverse, both NB x (NB+4).
//   factor   per 32-column panel: (1) warp 0 factors the 32 x 32 diagonal sub-block in registers
//            (lane = row; pivot by shuffle, scaled column through a shared buffer), (2) one thread per
//            row solves the panel below against it by forward substitution, (3) all threads apply the
//            rank-32 update to the rest (128-bit shared-memory products);
//   inverse  (only when an output consumes it) the 32 x 32 diagonal sub-blocks by forward substitution,
//            one warp each, then recursive doubling W21 = -W22 (L21 W11).
// Optional fp16 hi/lo emission (f32 only) of the factor block (Lh/Ll) and of the inverse block,
// plain (Wh/Wl) and transposed (WTh/WTl), for the tensor-core GEMMs that follow.
// ---------------------------------------------------------------------------------------------
// development aid (tools/diag_bench.cu): phase timestamps of diag_block_kernel
#ifdef GPG_DIAG_PROFILE
__device__ long long g_diag_clk[64];
#define GPG_PHASE(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_diag_clk[i] = clock64(); } while (0)
#else
#define GPG_PHASE(i)
#endif

struct DiagEmit {
    __half *Lh = nullptr, *Ll = nullptr;        // factor block at [j0+i][j0+k]
    __half *Wh = nullptr, *Wl = nullptr;        // inverse block at [j0+i][j0+k]
    __half *WTh = nullptr, *WTl = nullptr;      // inverse block transposed at [j0+k][j0+i]
    int64_t lds = 0;
    const float *scale_L = nullptr, *scale_W = nullptr;
};

__device__ __forceinline__ void emit_split(__half *hi, __half *lo, int64_t off, float v) {
    const __half h = __float2half_rn(v);
    hi[off] = h;
    lo[off] = __float2half_rn(v - __half2float(h));
}

// 128-bit shared-memory vectors (rows of S / W are 16-byte aligned: lds is a multiple of 4)
template <typename T> struct V4 { T v[4]; };
__device__ __forceinline__ V4<float> ld4(const float *p) {
    const float4 a = *reinterpret_cast<const float4 *>(p);
    V4<float> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; return r;
}
__device__ __forceinline__ V4<double> ld4(const double *p) {
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    V4<double> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y; return r;
}
__device__ __forceinline__ void st4(float *p, const V4<float> &r) {
    *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}
__device__ __forceinline__ void st4(double *p, const V4<double> &r) {
    *reinterpret_cast<double2 *>(p) = make_double2(r.v[0], r.v[1]);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(r.v[2], r.v[3]);
}

// 256 threads as a 16 x 16 grid (ty, tx); both products read four k at a time with 128-bit loads.
//   mm_kk: out(ty + 16 a, tx + 16 b)  += sum_k A[(ty + 16 a) * lds + k] * B[(tx + 16 b) * lds + k]
//   mm_kn: out(ty + 16 a, TJ tx + b)  += sum_k A[(ty + 16 a) * lds + k] * B[k * lds + TJ tx + b]
// (rows of A are warp-uniform up to two values -> broadcasts; rows of B in mm_kk are 4 banks apart per lane,
// conflict-free within each 8-lane phase of a 128-bit access; mm_kn reads B rows contiguously.)
template <typename T, int TI, int TJ>
__device__ __forceinline__ void mm_kk(const T *__restrict__ A, const T *__restrict__ B, int lds, int K,
                                      T (&acc)[TI][TJ], int ty, int tx) {
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        V4<T> a[TI], b[TJ];
#pragma unroll
        for (int i = 0; i < TI; ++i) a[i] = ld4(A + (ty + 16 * i) * lds + k);
#pragma unroll
        for (int j = 0; j < TJ; ++j) b[j] = ld4(B + (tx + 16 * j) * lds + k);
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) acc[i][j] = fma(a[i].v[e], b[j].v[e], acc[i][j]);
    }
}

template <typename T, int TI, int TJ>
__device__ __forceinline__ void mm_kn(const T *__restrict__ A, const T *__restrict__ B, int lds, int K,
                                      T (&acc)[TI][TJ], int ty, int tx) {
    static_assert(TJ == 2 || TJ == 4, "mm_kn: two or four consecutive columns per thread");
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        V4<T> a[TI];
#pragma unroll
        for (int i = 0; i < TI; ++i) a[i] = ld4(A + (ty + 16 * i) * lds + k);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            T b[TJ];
            if (TJ == 4) {
                const V4<T> bv = ld4(B + (k + e) * lds + 4 * tx);
#pragma unroll
                for (int j = 0; j < TJ; ++j) b[j] = bv.v[j];
            } else {
                b[0] = B[(k + e) * lds + 2 * tx];
                b[TJ - 1] = B[(k + e) * lds + 2 * tx + 1];
            }
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) acc[i][j] = fma(a[i].v[e], b[j], acc[i][j]);
        }
    }
}

// Rows below a freshly factored 32 x 32 diagonal block: X <- X L11^-T by forward substitution, one thread
// per row (the row lives in registers, the entries of L11 arrive as warp-wide broadcasts, 128 bits at a time).
template <typename T>
__device__ __forceinline__ void panel_solve32(T *__restrict__ S, int lds, int c0, int row, const T *__restrict__ rdiag) {
    T x[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const V4<T> t4 = ld4(S + row * lds + c0 + 4 * q);
#pragma unroll
        for (int e = 0; e < 4; ++e) x[4 * q + e] = t4.v[e];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        T s[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
        for (int k4 = 0; k4 + 4 <= j; k4 += 4) {
            const V4<T> l4 = ld4(S + (c0 + j) * lds + c0 + k4);
#pragma unroll
            for (int e = 0; e < 4; ++e) s[e] = fma(x[k4 + e], l4.v[e], s[e]);
        }
#pragma unroll
        for (int k = (j / 4) * 4; k < j; ++k) s[k & 3] = fma(x[k], S[(c0 + j) * lds + c0 + k], s[k & 3]);
        x[j] = (x[j] - ((s[0] + s[1]) + (s[2] + s[3]))) * rdiag[j];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        V4<T> t4;
#pragma unroll
        for (int e = 0; e < 4; ++e) t4.v[e] = x[4 * q + e];
        st4(S + row * lds + c0 + 4 * q, t4);
    }
}

// inverse of the 32 x 32 lower-triangular block at S[b0.., b0..] into W[b0.., b0..]; one warp, lane = column.
// Must be called by all 32 lanes (converged): the reciprocal diagonal travels by warp shuffle.
template <typename T>
__device__ __forceinline__ void invert_tri32(const T *__restrict__ S, T *__restrict__ W, int lds, int b0, int lane) {
    const T rd = T(1) / S[(b0 + lane) * lds + b0 + lane];
    T x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        T s0 = (i == lane) ? T(1) : T(0), s1 = T(0);
#pragma unroll
        for (int k = 0; k + 1 < i; k += 2) {
            s0 -= S[(b0 + i) * lds + b0 + k] * x[k];
            s1 -= S[(b0 + i) * lds + b0 + k + 1] * x[k + 1];
        }
        if (i & 1) s0 -= S[(b0 + i) * lds + b0 + i - 1] * x[i - 1];
        x[i] = (s0 + s1) * __shfl_sync(0xffffffffu, rd, i);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) W[(b0 + i) * lds + b0 + lane] = x[i];
}

__device__ __forceinline__ float gpg_rsqrt(float x) {          // the bare MUFU.RSQ (2 ulp): it sits on the pivot chain
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double gpg_rsqrt(double x) { return 1.0 / sqrt(x); }

// Cholesky of the 32 x 32 block at S[c0.., c0..] in registers (lane = row).  The serial chain of a pivot is kept as
// short as the data flow allows: the NEXT pivot's diagonal entry only needs the owning lane's own l_ij, so it is
// updated, broadcast (one shuffle) and sent through rsqrt BEFORE the scaled column is exchanged; the exchange goes
// through a double-buffered 32-entry shared array (one __syncwarp per pivot) and comes back as broadcast 128-bit
// loads for the rank-1 update of the rest of the row.  Returns 0 or 1 + the local index of the first non-positive
// pivot among the first `valid` columns.  colbuf: 64 entries.
template <typename T>
__device__ __forceinline__ int factor_tri32(T *__restrict__ S, int lds, int c0, int lane, int valid,
                                            T *__restrict__ colbuf, T *__restrict__ rdiag) {
    constexpr unsigned FULL = 0xffffffffu;
    T row[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) row[k] = S[(c0 + lane) * lds + c0 + k];
    T djj = __shfl_sync(FULL, row[0], 0);
    int bad = (!(djj > T(0)) && 0 < valid) ? 1 : 0;
    T rs = gpg_rsqrt(djj);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const T lij = row[j] * rs;                    // lane j: djj / sqrt(djj) = l_jj
        row[j] = lij;
        if (lane == j) rdiag[j] = rs;                 // 1 / l_jj for the panel solve
        if (j < 31) {
            T *cb = colbuf + (j & 1) * 32;
            cb[lane] = lij;
            if (lane == j + 1) row[j + 1] = fma(-lij, lij, row[j + 1]);
            const T dn = __shfl_sync(FULL, row[j + 1], j + 1);
            bad = (bad == 0 && !(dn > T(0)) && j + 1 < valid) ? j + 2 : bad;
            const T rsn = gpg_rsqrt(dn);
            __syncwarp();
#pragma unroll
            for (int k4 = ((j + 1) / 4) * 4; k4 < 32; k4 += 4) {
                T c[4];
                if (sizeof(T) == 4) {
                    const float4 v = *reinterpret_cast<const float4 *>(cb + k4);
                    c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
                } else {
                    const double2 a = *reinterpret_cast<const double2 *>(cb + k4);
                    const double2 b = *reinterpret_cast<const double2 *>(cb + k4 + 2);
                    c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (k4 + e == j + 1) {
                        if (lane != j + 1) row[j + 1] = fma(-lij, c[e], row[j + 1]);   // the owner has done its own already
                    } else if (k4 + e > j + 1) {
                        row[k4 + e] = fma(-lij, c[e], row[k4 + e]);     // meaningful for lane >= k only
                    }
                }
            }
            rs = rsn;
        }
    }
#pragma unroll
    for (int k = 0; k < 32; ++k)
        if (k <= lane) S[(c0 + lane) * lds + c0 + k] = row[k];
    return bad;
}

// Trailing update of the blocked factorisation inside a diagonal block: the lower triangle of
// S[base.., base..] (R = 16 TI rows) -= P P^T with P = S[base.., c0 .. c0 + 32).  256 threads as a 16 x 16 grid, each
// TI x TI outputs (rows ty + 16 i, columns tx + 16 j); tiles strictly above the diagonal (j > i) are not computed.
template <typename T, int TI>
__device__ __forceinline__ void trailing_lower32(T *__restrict__ S, int lds, int base, int c0, int ty, int tx) {
    T acc[TI][TI];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TI; ++j) acc[i][j] = T(0);
    const T *P = S + base * lds + c0;
#pragma unroll 2
    for (int k = 0; k < 32; k += 4) {
        V4<T> a[TI], b[TI];
#pragma unroll
        for (int i = 0; i < TI; ++i) a[i] = ld4(P + (ty + 16 * i) * lds + k);
#pragma unroll
        for (int j = 0; j < TI; ++j) b[j] = ld4(P + (tx + 16 * j) * lds + k);
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TI; ++j)
                    if (j <= i) acc[i][j] = fma(a[i].v[e], b[j].v[e], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TI; ++j) {
            const int gi = base + ty + 16 * i, gk = base + tx + 16 * j;
            if (j <= i && gk <= gi) S[gi * lds + gk] -= acc[i][j];
        }
}

// Blocked Cholesky of the NB x NB block held in shared memory (S: NB x (NB + 4), lower triangle; the strict upper
// triangle must be zero).  256 threads.  Right-looking over 32-column panels: (1) warp 0 factors the 32 x 32 diagonal
// sub-block in registers, (2) one thread per row below solves X L11^T = A21 by forward substitution, (3) all threads
// apply the rank-32 update to the trailing lower triangle in 32 x 32 chunks.  nb: valid rows of a ragged last block.
template <typename T, int NB>
__device__ __forceinline__ void diag_factor_smem(T *__restrict__ S, T *__restrict__ colbuf, T *__restrict__ rdiag, int nb,
                                                 int64_t j0, int32_t *__restrict__ info) {
    constexpr int LDS = NB + 4;
    constexpr int NSUB = NB / 32;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int tx = t & 15, ty = t >> 4;
    for (int p = 0; p < NSUB; ++p) {
        const int c0 = p * 32;
        if (warp == 0) {
            GPG_PHASE(2 + 5 * p);
            const int bad = factor_tri32<T>(S, LDS, c0, lane, nb - c0, colbuf, rdiag);
            if (lane == 0 && bad && info) atomicCAS(info, 0, (int32_t)(j0 + c0 + bad));
            GPG_PHASE(3 + 5 * p);
        }
        __syncthreads();
        const int base = c0 + 32;
        if (base < NB) {
            const int R = NB - base;
            if (t < R) panel_solve32<T>(S, LDS, c0, base + t, rdiag);
            __syncthreads();
            GPG_PHASE(5 + 5 * p);
            // trailing lower triangle: S[i][k] -= sum_m S[i][c0+m] S[k][c0+m], register tiles of R / 16 squared
            if (R == 96) trailing_lower32<T, 6>(S, LDS, base, c0, ty, tx);
            else if (R == 64) trailing_lower32<T, 4>(S, LDS, base, c0, ty, tx);
            else trailing_lower32<T, 2>(S, LDS, base, c0, ty, tx);
            __syncthreads();
            GPG_PHASE(6 + 5 * p);
        }
    }
}

// W = S^-1 for the lower-triangular factor in S (W must be zero on entry; both NB x (NB + 4)).  256 threads.
// The 32 x 32 diagonal sub-blocks by forward substitution, one warp each, then recursive doubling
// W21 = -W22 (L21 W11); the upper triangles of S and W are zero, so the products are plain dense ones.
template <typename T, int NB>
__device__ __forceinline__ void diag_invert_smem(const T *__restrict__ S, T *__restrict__ W) {
    constexpr int LDS = NB + 4;
    constexpr int NSUB = NB / 32;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int tx = t & 15, ty = t >> 4;
    if (warp < NSUB) invert_tri32<T>(S, W, LDS, warp * 32, lane);
    __syncthreads();
    if (NB >= 64) {
        for (int s0 = 0; s0 < NB; s0 += 64) {        // hb = 32: out(ty + 16 i, 2 tx + j)
            T acc[2][2] = {{T(0), T(0)}, {T(0), T(0)}};
            mm_kn<T, 2, 2>(S + (s0 + 32) * LDS + s0, W + s0 * LDS + s0, LDS, 32, acc, ty, tx);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) W[(s0 + 32 + ty + 16 * i) * LDS + s0 + 2 * tx + j] = acc[i][j];
            __syncthreads();
            T acc2[2][2] = {{T(0), T(0)}, {T(0), T(0)}};
            mm_kn<T, 2, 2>(W + (s0 + 32) * LDS + s0 + 32, W + (s0 + 32) * LDS + s0, LDS, 32, acc2, ty, tx);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) W[(s0 + 32 + ty + 16 * i) * LDS + s0 + 2 * tx + j] = -acc2[i][j];
            __syncthreads();
        }
    }
    GPG_PHASE(23);
    if (NB >= 128) {                                  // hb = 64: out(ty + 16 i, 4 tx + j)
        T acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
        mm_kn<T, 4, 4>(S + 64 * LDS, W, LDS, 64, acc, ty, tx);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) W[(64 + ty + 16 * i) * LDS + 4 * tx + j] = acc[i][j];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
        mm_kn<T, 4, 4>(W + 64 * LDS + 64, W + 64 * LDS, LDS, 64, acc, ty, tx);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) W[(64 + ty + 16 * i) * LDS + 4 * tx + j] = -acc[i][j];
        __syncthreads();
    }
}

template <typename T, int NB>
__global__ void __launch_bounds__(256) diag_block_kernel(T *__restrict__ A, int64_t ld, int64_t N, int64_t j0_first,
                                                         int do_factor, T *__restrict__ inv_out, int64_t ld_inv,
                                                         int64_t inv_block_stride, int dense_out,
                                                         int32_t *__restrict__ info, DiagEmit em) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LDS = NB + 4;                       // 16-byte aligned rows, 4 banks of skew per row
    T *S = reinterpret_cast<T *>(smem_raw);
    T *W = S + NB * LDS;
    __shared__ __align__(16) T colbuf[64];
    __shared__ __align__(16) T rdiag[32];
    const int t = threadIdx.x;
    const int64_t j0 = j0_first + (int64_t)blockIdx.x * NB;
    const int nb = (int)min((int64_t)NB, N - j0);
    T *Ab = A + j0 * ld + j0;
    pdl_trigger();
    pdl_wait();
    for (int idx = t; idx < NB * NB; idx += 256) {
        const int i = idx / NB, k = idx % NB;
        T v = (i == k) ? T(1) : T(0);                 // identity padding of a ragged last block
        if (i < nb && k <= i) v = Ab[(int64_t)i * ld + k];
        S[i * LDS + k] = v;
        W[i * LDS + k] = T(0);
    }
    GPG_PHASE(0);
    __syncthreads();
    GPG_PHASE(1);
    if (do_factor) {
        diag_factor_smem<T, NB>(S, colbuf, rdiag, nb, j0, info);
        GPG_PHASE(22);
    }
    // the inverse is computed only when somebody consumes it (the blocked Cholesky solves its panels against
    // the factor itself and leaves all inverses to the batched trtri that follows)
    const bool need_inv = (inv_out != nullptr) || (em.Wh != nullptr) || (em.WTh != nullptr);
    if (need_inv) diag_invert_smem<T, NB>(S, W);
    else __syncthreads();
    GPG_PHASE(24);
    static_assert(NB == 64 || NB == 128, "diag_block_kernel handles NB = 64 or 128");
    // Outputs, four consecutive columns per thread: factor block (fp32 + fp16 hi/lo), inverse block (fp32 +
    // hi/lo), then the transposed inverse planes.
    T *out = inv_out ? inv_out + (int64_t)blockIdx.x * inv_block_stride : nullptr;
    const float sL = em.scale_L ? *em.scale_L : 1.0f, sW = em.scale_W ? *em.scale_W : 1.0f;
    const bool f32 = sizeof(T) == 4;
    const bool vecA = f32 && ((reinterpret_cast<uintptr_t>(Ab) & 15) == 0) && ((ld & 3) == 0);
    const bool vecO = f32 && out && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && ((ld_inv & 3) == 0);
    const bool vecS = ((em.lds & 3) == 0);
    for (int q = t; q < NB * NB / 4; q += 256) {
        const int i = q / (NB / 4), k4 = (q % (NB / 4)) * 4;
        const int nvalid = (i < nb) ? max(0, min(4, nb - k4)) : 0;       // columns of this group inside the block
        const V4<T> s4 = ld4(S + i * LDS + k4), w4 = ld4(W + i * LDS + k4);
        if (do_factor && nvalid > 0 && k4 <= i) {      // entries right of the diagonal inside the group are zero in S
            T *dst = Ab + (int64_t)i * ld + k4;
            if (nvalid == 4 && vecA) *reinterpret_cast<float4 *>(dst) = make_float4((float)s4.v[0], (float)s4.v[1], (float)s4.v[2], (float)s4.v[3]);
            else for (int e = 0; e < nvalid; ++e) if (k4 + e <= i) dst[e] = s4.v[e];
        }
        if (out) {
            const int nw = dense_out ? 4 : nvalid;
            T *dst = out + (int64_t)i * ld_inv + k4;
            if (nw == 4 && vecO) *reinterpret_cast<float4 *>(dst) = make_float4((float)w4.v[0], (float)w4.v[1], (float)w4.v[2], (float)w4.v[3]);
            else for (int e = 0; e < nw; ++e) dst[e] = w4.v[e];
        }
        if (f32 && nvalid > 0 && (em.Lh || em.Wh)) {
            const int64_t off = (j0 + i) * em.lds + j0 + k4;
            if (em.Lh) {
                const float a0 = (k4 <= i) ? (float)s4.v[0] * sL : 0.f, a1 = (k4 + 1 <= i) ? (float)s4.v[1] * sL : 0.f;
                const float a2 = (k4 + 2 <= i) ? (float)s4.v[2] * sL : 0.f, a3 = (k4 + 3 <= i) ? (float)s4.v[3] * sL : 0.f;
                const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(a0 - f01.x, a1 - f01.y), l23 = __floats2half2_rn(a2 - f23.x, a3 - f23.y);
                if (nvalid == 4 && vecS) {
                    *reinterpret_cast<uint2 *>(em.Lh + off) = make_uint2(*reinterpret_cast<const unsigned *>(&h01), *reinterpret_cast<const unsigned *>(&h23));
                    *reinterpret_cast<uint2 *>(em.Ll + off) = make_uint2(*reinterpret_cast<const unsigned *>(&l01), *reinterpret_cast<const unsigned *>(&l23));
                } else {
                    const __half hh[4] = {__low2half(h01), __high2half(h01), __low2half(h23), __high2half(h23)};
                    const __half ll[4] = {__low2half(l01), __high2half(l01), __low2half(l23), __high2half(l23)};
                    for (int e = 0; e < nvalid; ++e) { em.Lh[off + e] = hh[e]; em.Ll[off + e] = ll[e]; }
                }
            }
            if (em.Wh) {
                const float a0 = (float)w4.v[0] * sW, a1 = (float)w4.v[1] * sW, a2 = (float)w4.v[2] * sW, a3 = (float)w4.v[3] * sW;
                const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                const __half2 l01 = __floats2half2_rn(a0 - f01.x, a1 - f01.y), l23 = __floats2half2_rn(a2 - f23.x, a3 - f23.y);
                if (nvalid == 4 && vecS) {
                    *reinterpret_cast<uint2 *>(em.Wh + off) = make_uint2(*reinterpret_cast<const unsigned *>(&h01), *reinterpret_cast<const unsigned *>(&h23));
                    *reinterpret_cast<uint2 *>(em.Wl + off) = make_uint2(*reinterpret_cast<const unsigned *>(&l01), *reinterpret_cast<const unsigned *>(&l23));
                } else {
                    const __half hh[4] = {__low2half(h01), __high2half(h01), __low2half(h23), __high2half(h23)};
                    const __half ll[4] = {__low2half(l01), __high2half(l01), __low2half(l23), __high2half(l23)};
                    for (int e = 0; e < nvalid; ++e) { em.Wh[off + e] = hh[e]; em.Wl[off + e] = ll[e]; }
                }
            }
        }
    }
    if (f32 && em.WTh) {
        // transposed planes: lane = source row i (consecutive destination columns), four source columns k per
        // thread: 128-bit conflict-free reads of W, 64-byte contiguous stores per warp and plane
        for (int q = t; q < NB * NB / 4; q += 256) {
            const int i = q % NB, k4 = (q / NB) * 4;
            if (i >= nb || k4 >= nb) continue;
            const V4<T> w4 = ld4(W + i * LDS + k4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (k4 + e < nb) {
                    const float a = (float)w4.v[e] * sW;
                    const __half hh = __float2half_rn(a);
                    const int64_t off = (j0 + k4 + e) * em.lds + j0 + i;
                    em.WTh[off] = hh;
                    em.WTl[off] = __float2half_rn(a - __half2float(hh));
                }
            }
        }
    }
    GPG_PHASE(25);
}

// ---------------------------------------------------------------------------------------------
// Cholesky panel: X <- X L11^-T for the rows below a 128 x 128 diagonal block, by forward substitution
// against the factor itself (backward stable, and the diagonal block's inverse drops off the critical
// path).  One thread per row: the row lives in 128 registers, the entries of L11 arrive as warp-wide
// broadcast 128-bit shared loads.  CTA = 128 rows; global traffic goes through shared memory so that it
// is coalesced.  Writes the fp32 result in place and its fp16 hi/lo planes.
// ---------------------------------------------------------------------------------------------
constexpr int PANEL_NB = 128, PANEL_LDS = PANEL_NB + 4;
static int panel_trsm_smem() { return (2 * PANEL_NB * PANEL_LDS + PANEL_NB) * (int)sizeof(float); }

__global__ void __launch_bounds__(PANEL_NB) panel_trsm_kernel(float *__restrict__ A, int64_t ld, int64_t j0, int64_t rows,
                                                               __half *__restrict__ Lh, __half *__restrict__ Ll,
                                                               int64_t lds_planes, const float *__restrict__ scale_L) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NB = PANEL_NB, LDS = PANEL_LDS;
    float *Ls = reinterpret_cast<float *>(smem_raw);          // L11, lower triangle
    float *Xs = Ls + NB * LDS;                                 // this CTA's rows of the panel
    float *rd = Xs + NB * LDS;                                 // 1 / diag(L11)
    const int t = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * NB;               // first row of this CTA (relative to the panel)
    const int nrows = (int)min((int64_t)NB, rows - r0);
    pdl_trigger();
    pdl_wait();
    const float *L11 = A + j0 * ld + j0;
    float *A21 = A + (j0 + NB + r0) * ld + j0;
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    for (int q = t; q < NB * NB / 4; q += NB) {
        const int i = q / (NB / 4), k4 = (q % (NB / 4)) * 4;
        float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f), x4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec) {
            if (k4 <= i) l4 = *reinterpret_cast<const float4 *>(L11 + (int64_t)i * ld + k4);
            if (i < nrows) x4 = *reinterpret_cast<const float4 *>(A21 + (int64_t)i * ld + k4);
        } else {
            float lt[4] = {0.f, 0.f, 0.f, 0.f}, xt[4] = {0.f, 0.f, 0.f, 0.f};
            for (int e = 0; e < 4; ++e) {
                if (k4 + e <= i) lt[e] = L11[(int64_t)i * ld + k4 + e];
                if (i < nrows) xt[e] = A21[(int64_t)i * ld + k4 + e];
            }
            l4 = make_float4(lt[0], lt[1], lt[2], lt[3]);
            x4 = make_float4(xt[0], xt[1], xt[2], xt[3]);
        }
        *reinterpret_cast<float4 *>(Ls + i * LDS + k4) = l4;       // entries right of the diagonal are never read
        *reinterpret_cast<float4 *>(Xs + i * LDS + k4) = x4;
    }
    __syncthreads();
    rd[t] = 1.0f / Ls[t * LDS + t];
    __syncthreads();
    float x[NB];
#pragma unroll
    for (int q = 0; q < NB / 4; ++q) {
        const float4 v = *reinterpret_cast<const float4 *>(Xs + t * LDS + 4 * q);
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int k4 = 0; k4 + 4 <= j; k4 += 4) {
            const float4 l4 = *reinterpret_cast<const float4 *>(Ls + j * LDS + k4);
            s0 = fmaf(x[k4], l4.x, s0); s1 = fmaf(x[k4 + 1], l4.y, s1);
            s2 = fmaf(x[k4 + 2], l4.z, s2); s3 = fmaf(x[k4 + 3], l4.w, s3);
        }
        float tail = 0.f;
#pragma unroll
        for (int k = (j / 4) * 4; k < j; ++k) tail = fmaf(x[k], Ls[j * LDS + k], tail);
        x[j] = (x[j] - (((s0 + s1) + (s2 + s3)) + tail)) * rd[j];
    }
#pragma unroll
    for (int q = 0; q < NB / 4; ++q)
        *reinterpret_cast<float4 *>(Xs + t * LDS + 4 * q) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
    __syncthreads();
    const float sL = *scale_L;
    const bool vecS = ((lds_planes & 3) == 0);
    for (int q = t; q < NB * NB / 4; q += NB) {
        const int i = q / (NB / 4), k4 = (q % (NB / 4)) * 4;
        if (i >= nrows) continue;
        const float4 v = *reinterpret_cast<const float4 *>(Xs + i * LDS + k4);
        float *dst = A21 + (int64_t)i * ld + k4;
        if (vec) *reinterpret_cast<float4 *>(dst) = v;
        else { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }
        const float a0 = v.x * sL, a1 = v.y * sL, a2 = v.z * sL, a3 = v.w * sL;
        const __half2 h01 = __floats2half2_rn(a0, a1), h23 = __floats2half2_rn(a2, a3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(a0 - f01.x, a1 - f01.y), l23 = __floats2half2_rn(a2 - f23.x, a3 - f23.y);
        const int64_t off = (j0 + NB + r0 + i) * lds_planes + j0 + k4;
        if (vecS) {
            *reinterpret_cast<uint2 *>(Lh + off) = make_uint2(*reinterpret_cast<const unsigned *>(&h01), *reinterpret_cast<const unsigned *>(&h23));
            *reinterpret_cast<uint2 *>(Ll + off) = make_uint2(*reinterpret_cast<const unsigned *>(&l01), *reinterpret_cast<const unsigned *>(&l23));
        } else {
            const __half hh[4] = {__low2half(h01), __high2half(h01), __low2half(h23), __high2half(h23)};
            const __half ll[4] = {__low2half(l01), __high2half(l01), __low2half(l23), __high2half(l23)};
            for (int e = 0; e < 4; ++e) { Lh[off + e] = hh[e]; Ll[off + e] = ll[e]; }
        }
    }
}

template <typename T, int NB> static int diag_block_smem() { return 2 * NB * (NB + 4) * (int)sizeof(T); }

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
        diag_block_kernel<T, NB><<<1, 256, diag_block_smem<T, NB>(), stream>>>(A, ld, N, j0, 1, dinv, NB, 0, 1, info,
                                                                               DiagEmit());
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
        const_cast<T *>(L), ld, N, 0, 0, Linv, ldi, (int64_t)NB * (ldi + 1), 0, nullptr, DiagEmit());
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
        if (sizeof(T) == 4 && (ld & 3) == 0 && ((reinterpret_cast<uintptr_t>(Mx) | reinterpret_cast<uintptr_t>(x)) & 15) == 0) {
            // 128-bit loads, four rows of loads in flight per lane; the tail group of the row is masked
            const float *row = reinterpret_cast<const float *>(Mx) + i * ld;
            const float *xf = reinterpret_cast<const float *>(x);
            double a0 = 0.0, a1 = 0.0;
            int64_t k = 4 * lane;
            for (; k + 128 + 3 <= i; k += 256) {
                const float4 m0 = *reinterpret_cast<const float4 *>(row + k), m1 = *reinterpret_cast<const float4 *>(row + k + 128);
                const float4 x0 = *reinterpret_cast<const float4 *>(xf + k), x1 = *reinterpret_cast<const float4 *>(xf + k + 128);
                a0 += (double)m0.x * x0.x + (double)m0.y * x0.y + (double)m0.z * x0.z + (double)m0.w * x0.w;
                a1 += (double)m1.x * x1.x + (double)m1.y * x1.y + (double)m1.z * x1.z + (double)m1.w * x1.w;
            }
            for (; k <= i; k += 128) {
                const float4 m0 = *reinterpret_cast<const float4 *>(row + k);        // k + 3 < ld: ld is a multiple of 4
                const float mm[4] = {m0.x, m0.y, m0.z, m0.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (k + e <= i) a0 += (double)mm[e] * (double)xf[k + e];
            }
            acc = a0 + a1;
        } else {
            for (int64_t k = lane; k <= i; k += 32) acc += (double)Mx[i * ld + k] * (double)x[k];
        }
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

// scalars = {0.5 y^T K^-1 y, sum log L_ii}; the quadratic form is |vhat|^2, or y . alpha when both are given
// Transposed triangular GEMV, out = alpha * Mx^T x + beta * yin, in two deterministic passes: a 2-D grid of
// (128-column strip) x (512-row chunk) blocks streams the lower triangle with 128-bit loads and leaves double
// partial sums per chunk; a second kernel adds the chunks up.  (One block per strip cannot pull the matrix
// fast enough: a single SM's load bandwidth caps it.)
constexpr int GEMVT_ROWS = 512;

template <typename T>
__global__ void __launch_bounds__(256) gemvT_partial_kernel(const T *__restrict__ Mx, int64_t ld, int64_t N,
                                                            const T *__restrict__ x, double *__restrict__ part) {
    __shared__ double red[8][128];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * 128, r0 = (int64_t)blockIdx.y * GEMVT_ROWS;
    const int64_t r1 = min(N, r0 + GEMVT_ROWS);
    if (r1 <= c0) return;                              // chunk entirely above the diagonal: never summed
    const int64_t c = c0 + 4 * lane;
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(Mx) & 15) == 0) && (sizeof(T) == 4);
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t i = max(r0, c0) + w; i < r1; i += 8) {
        if (c > i) continue;
        const double xi = (double)x[i];
        T m[4];
        if (vec && c + 3 < ld) {
            const float4 v = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(Mx) + i * ld + c);
            m[0] = (T)v.x; m[1] = (T)v.y; m[2] = (T)v.z; m[3] = (T)v.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) m[e] = (c + e <= i) ? Mx[i * ld + c + e] : T(0);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (c + e <= i) a[e] += (double)m[e] * xi;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) red[w][4 * lane + e] = a[e];
    __syncthreads();
    if (threadIdx.x < 128 && c0 + threadIdx.x < N) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r][threadIdx.x];
        part[(int64_t)blockIdx.y * N + c0 + threadIdx.x] = s;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) gemvT_finish_kernel(const double *__restrict__ part, int64_t N,
                                                           const T *__restrict__ yin, T alpha, T beta,
                                                           T *__restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (k >= N) return;
    const int64_t first = ((k / 128) * 128) / GEMVT_ROWS;    // first row chunk that reaches this strip's diagonal
    const int64_t nch = (N + GEMVT_ROWS - 1) / GEMVT_ROWS;
    double s = 0.0;
    for (int64_t ch = first; ch < nch; ++ch) s += part[ch * N + k];
    out[k] = (T)((double)alpha * s + (yin ? (double)beta * (double)yin[k] : 0.0));
}

template <typename T>
static int gemv_tri_T(gpg_handle_s *h, const T *Mx, int64_t ld, int64_t N, const T *x, const T *yin, T alpha, T beta,
                      T *out, cudaStream_t stream) {
    const int64_t nch = (N + GEMVT_ROWS - 1) / GEMVT_ROWS;
    double *part;
    GPG_TRY(gpg_gemv_part_reserve(h, (size_t)nch * N, &part));
    dim3 grid((unsigned)((N + 127) / 128), (unsigned)nch);
    gemvT_partial_kernel<T><<<grid, 256, 0, stream>>>(Mx, ld, N, x, part);
    GPG_LAUNCH_CHECK(h);
    gemvT_finish_kernel<T><<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(part, N, yin, alpha, beta, out);
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

template <typename T>
__global__ void __launch_bounds__(1024) solve_scalars_kernel(const T *__restrict__ L, int64_t ld, int64_t N,
                                                             const T *__restrict__ vhat, T *__restrict__ scalars,
                                                             const T *__restrict__ y = nullptr,
                                                             const T *__restrict__ alpha = nullptr) {
    // one CTA; the diagonal gathers are latency-bound (one sector each), so as many threads as a CTA can hold
    __shared__ double r0[32], r1[32];
    double q = 0.0, ld_sum = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        if (y) q += (double)y[i] * (double)alpha[i];
        else { const double v = (double)vhat[i]; q += v * v; }
        ld_sum += log((double)L[i * ld + i]);
    }
    q = warp_sum(q);
    ld_sum = warp_sum(ld_sum);
    if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = q; r1[threadIdx.x >> 5] = ld_sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += r0[w]; b += r1[w]; }
        scalars[0] = (T)(0.5 * a);
        scalars[1] = (T)b;
    }
}

// One guarded step of iterative refinement bookkeeping (single CTA): n1 = |r_new|^2; if it beats the best
// residual so far (*best, < 0 = none yet) the candidate {alpha_new, r_new} replaces {alpha, r}.
template <typename T>
__global__ void __launch_bounds__(1024) refine_select_kernel(int64_t N, const T *__restrict__ alpha_new,
                                                             const T *__restrict__ r_new, T *__restrict__ alpha,
                                                             T *__restrict__ r, double *__restrict__ best) {
    __shared__ double red[32];
    __shared__ int accept;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += 1024) { const double v = (double)r_new[i]; s += v * v; }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 32; ++w) tot += red[w];
        accept = (*best < 0.0 || tot < *best) ? 1 : 0;        // NaN never wins
        if (accept) *best = tot;
    }
    __syncthreads();
    if (!accept) return;
    for (int64_t i = threadIdx.x; i < N; i += 1024) {
        if (alpha_new != alpha) alpha[i] = alpha_new[i];
        if (r_new != r) r[i] = r_new[i];
    }
}

// vhat = L^-1 y and alpha = L^-T vhat through the explicit inverse plus two residual corrections
// against L (each correction restores the backward error of a substitution solve).
// scratch: 2*N elements.
template <typename T>
static int solve_vec_refined(gpg_handle_s *h, const T *L, const T *Linv, int64_t N, int64_t ld, const T *y, T *vhat,
                             T *alpha, T *scalars, T *scratch, cudaStream_t stream, int corrections = 2) {
    T *r = scratch, *dx = scratch + N;
    const int gN = (int)((N + 7) / 8);
    // forward: L v = y
    gemv_tri_kernel<T, false><<<gN, 256, 0, stream>>>(Linv, ld, N, y, nullptr, T(1), T(0), vhat);
    GPG_LAUNCH_CHECK(h);
    for (int it = 0; it < corrections; ++it) {
        gemv_tri_kernel<T, false><<<gN, 256, 0, stream>>>(L, ld, N, vhat, y, T(-1), T(1), r);        // r = y - L v
        GPG_LAUNCH_CHECK(h);
        gemv_tri_kernel<T, false><<<gN, 256, 0, stream>>>(Linv, ld, N, r, vhat, T(1), T(1), dx);     // dx = v + Linv r
        GPG_LAUNCH_CHECK(h);
        GPG_CUDA_CHECK(cudaMemcpyAsync(vhat, dx, N * sizeof(T), cudaMemcpyDeviceToDevice, stream));
    }
    // backward: L^T a = v
    GPG_TRY(gemv_tri_T<T>(h, Linv, ld, N, vhat, nullptr, T(1), T(0), alpha, stream));
    for (int it = 0; it < corrections; ++it) {
        GPG_TRY(gemv_tri_T<T>(h, L, ld, N, alpha, vhat, T(-1), T(1), r, stream));
        GPG_TRY(gemv_tri_T<T>(h, Linv, ld, N, r, alpha, T(1), T(1), dx, stream));
        GPG_CUDA_CHECK(cudaMemcpyAsync(alpha, dx, N * sizeof(T), cudaMemcpyDeviceToDevice, stream));
    }
    if (scalars) {
        solve_scalars_kernel<T><<<1, (N >= 1024 ? 1024 : 256), 0, stream>>>(L, ld, N, vhat, scalars);
        GPG_LAUNCH_CHECK(h);
    }
    return GPG_OK;
}
