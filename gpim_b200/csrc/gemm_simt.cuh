// gemm_simt.cuh -- register-tiled SIMT GEMM used by every dense step when the tcgen05 path does
// not apply (fp64 everywhere; fp32 for small problems):
//     C[m][n] (+)= alpha * sum_k Aop(m,k) * Bop(n,k)
// with Aop(m,k) = a_kmajor ? A[m*lda+k] : A[k*lda+m] and the same for B, per-tile triangular
// k-ranges (so the structural zeros of L / L^-1 are skipped at tile granularity), an optional
// lower-tiles-only sweep (SYRK) and a fused column-sum-of-squares epilogue (predictive variance:
// diag(K*^T K^-1 K*) = colsum((L^-1 K*)^2), util.conditional's W.pow(2).sum(-1)).
#pragma once
#include "common.cuh"

enum { GEMM_KB_NONE = 0, GEMM_KB_N0 = 1, GEMM_KB_MAXMN = 2 };
enum { GEMM_KE_NONE = 0, GEMM_KE_M = 1, GEMM_KE_N = 2 };
enum { GEMM_TILES_ALL = 0, GEMM_TILES_LOWER = 1 };
enum { GEMM_EPI_STORE = 0, GEMM_EPI_COLSUMSQ = 1 };

template <typename T> struct GemmArgs {
    const T *A = nullptr;
    const T *B = nullptr;
    T *C = nullptr;
    int64_t lda = 0, ldb = 0, ldc = 0;
    int M = 0, N = 0, K = 0;
    int a_kmajor = 1, b_kmajor = 1;
    int kb_mode = GEMM_KB_NONE, ke_mode = GEMM_KE_NONE, tile_mode = GEMM_TILES_ALL, epi = GEMM_EPI_STORE;
    T alpha = T(1), beta = T(0);
    int64_t strideA = 0, strideB = 0, strideC = 0;
    int batch = 1;
    T *part = nullptr;       // COLSUMSQ: part[tile_m * ldpart + n]
    int64_t ldpart = 0;
    // STORE, f32 only: also emit the fp16 hi/lo split of the stored value times *split_scale at
    // [m * ld_split + n] (operand of a following tensor-core GEMM, see gemm_tc.cuh)
    __half *split_hi = nullptr, *split_lo = nullptr;
    int64_t ld_split = 0;
    const float *split_scale = nullptr;
    // optional compact support of operand B (K* rows of the predict product): krange[2 t], krange[2 t + 1] = the k range
    // [lo, hi) outside of which the B rows of the 128-row group t = n / 128 are negligible (and not even stored);
    // lo is a multiple of 128
    const int *krange = nullptr;
};

template <typename T> struct GemmCfg;
template <> struct GemmCfg<float> { static constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8; static constexpr bool COLSUM = true; };
template <> struct GemmCfg<double> { static constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4; static constexpr bool COLSUM = true; };
// skinny tiles for the Cholesky panel (M rows x 128 x 128): four times as many CTAs as the square config,
// which is what the short dependent steps of the factorisation need
struct GemmCfgPanelF32 { static constexpr int BM = 32, BN = 128, BK = 16, TM = 2, TN = 8; static constexpr bool COLSUM = false; };

template <typename T, typename C = GemmCfg<T>>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs<T> g) {
    constexpr int BM = C::BM, BN = C::BN, BK = C::BK, TM = C::TM, TN = C::TN;
    constexpr int TMH = TM / 2, TNH = TN / 2;
    constexpr int NT = 256;
    constexpr int ELA = BM * BK / NT, ELB = BN * BK / NT;
    constexpr int LDS_A = BM + 4, LDS_B = BN + 4;
    __shared__ __align__(16) T As[BK * LDS_A];
    __shared__ __align__(16) T Bs[BK * LDS_B];

    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (g.tile_mode == GEMM_TILES_LOWER && n0 > m0 + BM - 1) return;
    const T *A = g.A + (int64_t)blockIdx.z * g.strideA;
    const T *B = g.B + (int64_t)blockIdx.z * g.strideB;
    T *Cp = g.C + (int64_t)blockIdx.z * g.strideC;

    int kb = 0, ke = g.K;
    if (g.kb_mode == GEMM_KB_N0) kb = n0;
    else if (g.kb_mode == GEMM_KB_MAXMN) kb = max(m0, n0);
    if (g.ke_mode == GEMM_KE_M) ke = min(g.K, m0 + BM);
    else if (g.ke_mode == GEMM_KE_N) ke = min(g.K, n0 + BN);
    if (g.krange) {
        static_assert(128 % BN == 0, "a column tile must not straddle two support groups");
        kb = max(kb, __ldg(g.krange + 2 * (n0 / 128)));
        ke = min(ke, __ldg(g.krange + 2 * (n0 / 128) + 1));
    }
    kb = (kb / BK) * BK;

    const int t = threadIdx.x;
    const int tx = t % (BN / TN), ty = t / (BN / TN);

    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

    T ra[ELA], rb[ELB];

    auto load_a = [&](int k0) {
        if (g.a_kmajor) {
            constexpr int TPR = BK / ELA;
            const int r = t / TPR, ko = (t % TPR) * ELA;
            const int m = m0 + r;
            const T *src = A + (int64_t)m * g.lda + k0 + ko;
#pragma unroll
            for (int e = 0; e < ELA; ++e) ra[e] = (m < g.M && k0 + ko + e < ke) ? src[e] : T(0);
        } else {
            constexpr int TPK = BM / ELA;
            const int k = k0 + t / TPK, mo = (t % TPK) * ELA;
            const T *src = A + (int64_t)k * g.lda + m0 + mo;
#pragma unroll
            for (int e = 0; e < ELA; ++e) ra[e] = (k < ke && m0 + mo + e < g.M) ? src[e] : T(0);
        }
    };
    auto load_b = [&](int k0) {
        if (g.b_kmajor) {
            constexpr int TPR = BK / ELB;
            const int r = t / TPR, ko = (t % TPR) * ELB;
            const int n = n0 + r;
            const T *src = B + (int64_t)n * g.ldb + k0 + ko;
#pragma unroll
            for (int e = 0; e < ELB; ++e) rb[e] = (n < g.N && k0 + ko + e < ke) ? src[e] : T(0);
        } else {
            constexpr int TPK = BN / ELB;
            const int k = k0 + t / TPK, no = (t % TPK) * ELB;
            const T *src = B + (int64_t)k * g.ldb + n0 + no;
#pragma unroll
            for (int e = 0; e < ELB; ++e) rb[e] = (k < ke && n0 + no + e < g.N) ? src[e] : T(0);
        }
    };
    auto store_smem = [&]() {
        if (g.a_kmajor) {
            constexpr int TPR = BK / ELA;
            const int r = t / TPR, ko = (t % TPR) * ELA;
#pragma unroll
            for (int e = 0; e < ELA; ++e) As[(ko + e) * LDS_A + r] = ra[e];
        } else {
            constexpr int TPK = BM / ELA;
            const int k = t / TPK, mo = (t % TPK) * ELA;
#pragma unroll
            for (int e = 0; e < ELA; ++e) As[k * LDS_A + mo + e] = ra[e];
        }
        if (g.b_kmajor) {
            constexpr int TPR = BK / ELB;
            const int r = t / TPR, ko = (t % TPR) * ELB;
#pragma unroll
            for (int e = 0; e < ELB; ++e) Bs[(ko + e) * LDS_B + r] = rb[e];
        } else {
            constexpr int TPK = BN / ELB;
            const int k = t / TPK, no = (t % TPK) * ELB;
#pragma unroll
            for (int e = 0; e < ELB; ++e) Bs[k * LDS_B + no + e] = rb[e];
        }
    };

    if (kb < ke) {
        load_a(kb);
        load_b(kb);
    }
    for (int k0 = kb; k0 < ke; k0 += BK) {
        __syncthreads();
        store_smem();
        __syncthreads();
        if (k0 + BK < ke) {
            load_a(k0 + BK);
            load_b(k0 + BK);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            T a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TMH; ++i) {
                a[i] = As[kk * LDS_A + ty * TMH + i];
                a[TMH + i] = As[kk * LDS_A + BM / 2 + ty * TMH + i];
            }
#pragma unroll
            for (int j = 0; j < TNH; ++j) {
                b[j] = Bs[kk * LDS_B + tx * TNH + j];
                b[TNH + j] = Bs[kk * LDS_B + BN / 2 + tx * TNH + j];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }

    if (g.epi == GEMM_EPI_STORE) {
        const float sscale = (sizeof(T) == 4 && g.split_hi) ? *g.split_scale : 0.f;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int m = m0 + (i < TMH ? ty * TMH + i : BM / 2 + ty * TMH + (i - TMH));
            if (m >= g.M) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = n0 + (j < TNH ? tx * TNH + j : BN / 2 + tx * TNH + (j - TNH));
                if (n >= g.N) continue;
                T *dst = Cp + (int64_t)m * g.ldc + n;
                T v = g.alpha * acc[i][j];
                if (g.beta != T(0)) v += g.beta * *dst;
                *dst = v;
                if (sizeof(T) == 4 && g.split_hi) {
                    const float sv = (float)v * sscale;
                    const __half hh = __float2half_rn(sv);
                    g.split_hi[(int64_t)m * g.ld_split + n] = hh;
                    g.split_lo[(int64_t)m * g.ld_split + n] = __float2half_rn(sv - __half2float(hh));
                }
            }
        }
    } else if constexpr (C::COLSUM) {
        // column sums of squares over this tile's rows -> part[tile_m][n]
        __syncthreads();
        T *red = As;   // (BM/TM) x BN values fit: 16 x 128 <= BK * LDS_A
        static_assert((BM / TM) * BN <= BK * LDS_A, "reduction scratch too small");
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            T s = T(0);
#pragma unroll
            for (int i = 0; i < TM; ++i) s = fma(acc[i][j], acc[i][j], s);
            const int nl = (j < TNH ? tx * TNH + j : BN / 2 + tx * TNH + (j - TNH));
            red[ty * BN + nl] = s;
        }
        __syncthreads();
        for (int nl = t; nl < BN; nl += NT) {
            T s = T(0);
#pragma unroll
            for (int r = 0; r < BM / TM; ++r) s += red[r * BN + nl];
            if (n0 + nl < g.N)
                g.part[(int64_t)blockIdx.z * g.strideC + (int64_t)blockIdx.y * g.ldpart + n0 + nl] = s;
        }
    }
}

template <typename T, typename C = GemmCfg<T>>
static int gemm_simt(gpg_handle_s *h, const GemmArgs<T> &g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return GPG_OK;
    if (!C::COLSUM && g.epi != GEMM_EPI_STORE) { gpg_set_error("this GEMM tile config has no reduction epilogue"); return GPG_EINVAL; }
    dim3 grid((g.N + C::BN - 1) / C::BN, (g.M + C::BM - 1) / C::BM, g.batch);
    gemm_simt_kernel<T, C><<<grid, 256, 0, stream>>>(g);
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

// ---------------------------------------------------------------------------------------------
// fp64 on the tensor cores (DMMA, mma.sync.m8n8k4.f64): the reference's default precision (gpr.py:92).
// Same contract as gemm_simt_kernel<double> (operand layouts, per-tile triangular k-ranges, compact support, lower-tiles
// sweep, both epilogues); CTA tile 128 x 128 x 16, 8 warps as 2 x 4, warp tile 64 x 32 = 8 x 4 fragments of 8 x 8.
// B200's fp64 tensor peak equals its DFMA peak; the gain is operand delivery: a fragment pair feeds 256 FMAs per
// 64-bit shared-memory load per lane, where the scalar 4 x 4 register tile of the SIMT kernel gets 2 -- that kernel is
// bound by the load/store unit at ~35 % of peak.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) gemm_dmma_kernel(GemmArgs<double> g) {
    using T = double;
    constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
    constexpr int ELA = BM * BK / NT, ELB = BN * BK / NT;
    constexpr int LDS_A = BM + 4, LDS_B = BN + 4;          // (k * LDS + m) * 8 bytes: fragment loads hit 16 distinct bank pairs
    __shared__ __align__(16) T As[BK * LDS_A];
    __shared__ __align__(16) T Bs[BK * LDS_B];

    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (g.tile_mode == GEMM_TILES_LOWER && n0 > m0 + BM - 1) return;
    const T *A = g.A + (int64_t)blockIdx.z * g.strideA;
    const T *B = g.B + (int64_t)blockIdx.z * g.strideB;
    T *Cp = g.C + (int64_t)blockIdx.z * g.strideC;

    int kb = 0, ke = g.K;
    if (g.kb_mode == GEMM_KB_N0) kb = n0;
    else if (g.kb_mode == GEMM_KB_MAXMN) kb = max(m0, n0);
    if (g.ke_mode == GEMM_KE_M) ke = min(g.K, m0 + BM);
    else if (g.ke_mode == GEMM_KE_N) ke = min(g.K, n0 + BN);
    if (g.krange) {
        kb = max(kb, __ldg(g.krange + 2 * (n0 / 128)));
        ke = min(ke, __ldg(g.krange + 2 * (n0 / 128) + 1));
    }
    kb = (kb / BK) * BK;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;  // this warp's 64 x 32 corner inside the CTA tile
    const int fg = lane >> 2, ft = lane & 3;                // fragment row (A) / column (B); k index

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    T ra[ELA], rb[ELB];
    auto load_a = [&](int k0) {
        if (g.a_kmajor) {
            constexpr int TPR = BK / ELA;
            const int r = t / TPR, ko = (t % TPR) * ELA;
            const int m = m0 + r;
            const T *src = A + (int64_t)m * g.lda + k0 + ko;
#pragma unroll
            for (int e = 0; e < ELA; ++e) ra[e] = (m < g.M && k0 + ko + e < ke) ? src[e] : T(0);
        } else {
            constexpr int TPK = BM / ELA;
            const int k = k0 + t / TPK, mo = (t % TPK) * ELA;
            const T *src = A + (int64_t)k * g.lda + m0 + mo;
#pragma unroll
            for (int e = 0; e < ELA; ++e) ra[e] = (k < ke && m0 + mo + e < g.M) ? src[e] : T(0);
        }
    };
    auto load_b = [&](int k0) {
        if (g.b_kmajor) {
            constexpr int TPR = BK / ELB;
            const int r = t / TPR, ko = (t % TPR) * ELB;
            const int n = n0 + r;
            const T *src = B + (int64_t)n * g.ldb + k0 + ko;
#pragma unroll
            for (int e = 0; e < ELB; ++e) rb[e] = (n < g.N && k0 + ko + e < ke) ? src[e] : T(0);
        } else {
            constexpr int TPK = BN / ELB;
            const int k = k0 + t / TPK, no = (t % TPK) * ELB;
            const T *src = B + (int64_t)k * g.ldb + n0 + no;
#pragma unroll
            for (int e = 0; e < ELB; ++e) rb[e] = (k < ke && n0 + no + e < g.N) ? src[e] : T(0);
        }
    };
    auto store_smem = [&]() {
        if (g.a_kmajor) {
            constexpr int TPR = BK / ELA;
            const int r = t / TPR, ko = (t % TPR) * ELA;
#pragma unroll
            for (int e = 0; e < ELA; ++e) As[(ko + e) * LDS_A + r] = ra[e];
        } else {
            constexpr int TPK = BM / ELA;
            const int k = t / TPK, mo = (t % TPK) * ELA;
#pragma unroll
            for (int e = 0; e < ELA; ++e) As[k * LDS_A + mo + e] = ra[e];
        }
        if (g.b_kmajor) {
            constexpr int TPR = BK / ELB;
            const int r = t / TPR, ko = (t % TPR) * ELB;
#pragma unroll
            for (int e = 0; e < ELB; ++e) Bs[(ko + e) * LDS_B + r] = rb[e];
        } else {
            constexpr int TPK = BN / ELB;
            const int k = t / TPK, no = (t % TPK) * ELB;
#pragma unroll
            for (int e = 0; e < ELB; ++e) Bs[k * LDS_B + no + e] = rb[e];
        }
    };

    if (kb < ke) {
        load_a(kb);
        load_b(kb);
    }
    for (int k0 = kb; k0 < ke; k0 += BK) {
        __syncthreads();
        store_smem();
        __syncthreads();
        if (k0 + BK < ke) {
            load_a(k0 + BK);
            load_b(k0 + BK);
        }
#pragma unroll
        for (int k4 = 0; k4 < BK; k4 += 4) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = As[(k4 + ft) * LDS_A + wm + 8 * i + fg];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(k4 + ft) * LDS_B + wn + 8 * j + fg];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j], a[i], b[j]);
        }
    }

    // accumulator fragment (i, j): rows wm + 8 i + fg, columns wn + 8 j + 2 ft, + 1
    if (g.epi == GEMM_EPI_STORE) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + wm + 8 * i + fg;
            if (m >= g.M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int n = n0 + wn + 8 * j + 2 * ft + e;
                    if (n >= g.N) continue;
                    T *dst = Cp + (int64_t)m * g.ldc + n;
                    T v = g.alpha * acc[i][j][e];
                    if (g.beta != T(0)) v += g.beta * *dst;
                    *dst = v;
                }
        }
    } else {
        // column sums of squares over this tile's 128 rows -> part[tile_m][n]: over the 8 fragments of a lane, over the
        // 8 lanes that share ft (shuffles), over the two warps that share the columns (shared memory)
        __syncthreads();
        T *red = As;                                        // 2 x 128 values
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                T s = T(0);
#pragma unroll
                for (int i = 0; i < 8; ++i) s = fma(acc[i][j][e], acc[i][j][e], s);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                if (fg == 0) red[(warp >> 2) * BN + wn + 8 * j + 2 * ft + e] = s;
            }
        __syncthreads();
        for (int nl = t; nl < BN; nl += NT)
            if (n0 + nl < g.N)
                g.part[(int64_t)blockIdx.z * g.strideC + (int64_t)blockIdx.y * g.ldpart + n0 + nl] = red[nl] + red[BN + nl];
    }
}

// fp64 GEMM: DMMA tiles of 128 x 128 when the problem fills them, the 64 x 64 SIMT tiles otherwise.  `tile_m_rows`
// tells a COLSUMSQ caller how many rows one partial-sum row stands for.
// (the rank-64 updates and the short dependent GEMMs of the blocked fp64 Cholesky keep the 64 x 64 SIMT tiles: four
// times the CTAs, and with K = 64 the larger tile is all prologue and epilogue -- measured 23.3 against 18.6 ms at C2)
static inline bool gemm_f64_uses_dmma(const GemmArgs<double> &g) {
    return g.K >= 256 && (int64_t)((g.M + 127) / 128) * ((g.N + 127) / 128) * g.batch >= 148;
}
static int gemm_dmma(gpg_handle_s *h, const GemmArgs<double> &g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || g.batch <= 0) return GPG_OK;
    dim3 grid((g.N + 127) / 128, (g.M + 127) / 128, g.batch);
    gemm_dmma_kernel<<<grid, 256, 0, stream>>>(g);
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}
