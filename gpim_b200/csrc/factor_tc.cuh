// factor_tc.cuh -- f32 factorisation on the tensor cores: blocked Cholesky whose trailing SYRK
// updates run on the split-fp16 tcgen05 GEMM (gemm_tc.cuh), and the triangular inverse by
// recursive doubling with both GEMMs of every level on the same kernel.
//
// Operand bookkeeping.  Every tensor-core operand is an fp16 hi/lo pair of K-major planes with the
// geometry of the N x ld matrix, produced where the fp32 value is born (no separate conversion
// passes):
//   Ls  = split(s_L L)        diag blocks by diag_block_kernel, panels by the SIMT panel GEMM
//   Ws  = split(s_W L^-1)     diag blocks by diag_block_kernel, off-diagonal blocks by the epilogue
//   WTs = split(s_W L^-T)     of the level that computes them (plain and transposed emission)
//   TTs = split(s_T (L21 W11)^T)   intermediate of a doubling level, transposed emission
// The power-of-two scales come from rigorous bounds (scales_from_theta_kernel), so no data pass
// is needed to find them and nothing can overflow fp16.
#pragma once
#include "factor.cuh"
#include "gemm_tc.cuh"

// scales[] layout (device float[16])
enum {
    SC_K = 0,          // s_K   K* operand
    SC_W = 1,          // s_W   L^-1 operand
    SC_INV_KW = 2,     // 1 / (s_K s_W)
    SC_L = 3,          // s_L   L operand
    SC_INV_LL = 4,     // 1 / s_L^2
    SC_T = 5,          // s_T   L21 W11 intermediate
    SC_INV_LW = 6,     // 1 / (s_L s_W)
    SC_INV_WT = 7,     // 1 / (s_W s_T)
    SC_INV_WW = 8,     // 1 / s_W^2
    SC_COUNT = 16
};

__device__ __forceinline__ int pow2_exp_below(float x) {       // floor(log2(x)) clamped
    int e = (int)floorf(log2f(fmaxf(x, 1e-30f)));
    return max(-40, min(40, e));
}

// Bounds (K = K_f + (noise + jitter) I, K_f PSD with entries <= variance):
//   |K*_ij| <= v;  |L_ij| <= sqrt(v + nz);  |(L^-1)_ij| <= ||L^-1||_2 <= nz^-1/2;
//   |(L21 W11)_ij| <= ||L||_2 ||W11||_2 <= sqrt(N v + nz) nz^-1/2.
template <typename T>
__global__ void scales_from_theta_kernel(const T *__restrict__ theta, float jitter, float n_rows,
                                         float *__restrict__ scales) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float v = fmaxf((float)theta[0], 1e-30f);
    const float nz = fmaxf((float)theta[1] + jitter, 1e-30f);
    const int ek = pow2_exp_below(16384.0f / v);
    const int ew = pow2_exp_below(16384.0f * sqrtf(nz));
    const int el = pow2_exp_below(16384.0f / sqrtf(v + nz));
    const int et = pow2_exp_below(16384.0f * sqrtf(nz) / sqrtf(n_rows * v + nz));
    for (int i = 0; i < SC_COUNT; ++i) scales[i] = 0.f;
    scales[SC_K] = exp2f((float)ek);
    scales[SC_W] = exp2f((float)ew);
    scales[SC_INV_KW] = exp2f((float)(-ek - ew));
    scales[SC_L] = exp2f((float)el);
    scales[SC_INV_LL] = exp2f((float)(-2 * el));
    scales[SC_T] = exp2f((float)et);
    scales[SC_INV_LW] = exp2f((float)(-el - ew));
    scales[SC_INV_WT] = exp2f((float)(-ew - et));
    scales[SC_INV_WW] = exp2f((float)(-2 * ew));
}

// For a caller-supplied SPD matrix (gpg_cholesky): |L_ij| <= sqrt(max_i A_ii).
__global__ void __launch_bounds__(256) scales_from_diag_kernel(const float *__restrict__ A, int64_t ld, int64_t N,
                                                               float *__restrict__ scales) {
    __shared__ float red[8];
    float m = 0.f;
    for (int64_t i = threadIdx.x; i < N; i += 256) m = fmaxf(m, fabsf(A[i * ld + i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
        const int el = pow2_exp_below(16384.0f / sqrtf(fmaxf(m, 1e-30f)));
        for (int i = 0; i < SC_COUNT; ++i) scales[i] = 0.f;
        scales[SC_L] = exp2f((float)el);
        scales[SC_INV_LL] = exp2f((float)(-2 * el));
    }
}

struct TcPlanes {             // hi plane followed by lo plane, each N x ld halves
    __half *hi = nullptr, *lo = nullptr;
    TcPlanes() {}
    TcPlanes(void *base, int64_t N, int64_t ld) : hi((__half *)base), lo((__half *)base + (size_t)N * ld) {}
};

static inline void tc_params_clear(tc::Launch &g) { memset(&g.p, 0, sizeof(g.p)); }

// ---------------------------------------------------------------------------------------------
// Blocked right-looking Cholesky, NB = 128.  Per block column: diagonal block factor + inverse in
// one CTA; panel A21 <- A21 inv(L11)^T on the SIMT GEMM (K = 128, also emits the panel's split);
// trailing A22 -= A21 A21^T (lower tiles) on tcgen05.
// ---------------------------------------------------------------------------------------------
static int cholesky_blocked_tc(gpg_handle_s *h, float *A, int64_t N, int64_t ld, int32_t *info, int reset_info,
                               float *dinv, TcPlanes Ls, const float *scales, cudaStream_t stream) {
    constexpr int NB = 128;
    static bool attr_set = false;
    if (!attr_set) {
        GPG_CUDA_CHECK(cudaFuncSetAttribute(diag_block_kernel<float, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            diag_block_smem<float, NB>()));
        attr_set = true;
    }
    if (reset_info) GPG_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), stream));
    for (int64_t j0 = 0; j0 < N; j0 += NB) {
        const int64_t nb = std::min<int64_t>(NB, N - j0);
        DiagEmit em;
        em.Lh = Ls.hi; em.Ll = Ls.lo; em.lds = ld; em.scale_L = scales + SC_L;
        diag_block_kernel<float, NB><<<1, 256, diag_block_smem<float, NB>(), stream>>>(A, ld, N, j0, 1, dinv, NB, 0, 1,
                                                                                       info, em);
        GPG_LAUNCH_CHECK(h);
        const int64_t rows = N - j0 - nb;
        if (rows <= 0) break;
        float *A21 = A + (j0 + nb) * ld + j0;
        GemmArgs<float> p;                   // A21 <- A21 * inv(L11)^T (in place: one n-tile), + split
        p.A = A21; p.lda = ld; p.a_kmajor = 1;
        p.B = dinv; p.ldb = NB; p.b_kmajor = 1;
        p.C = A21; p.ldc = ld;
        p.M = (int)rows; p.N = (int)nb; p.K = (int)nb;
        p.ke_mode = GEMM_KE_N;
        p.split_hi = Ls.hi + (j0 + nb) * ld + j0;
        p.split_lo = Ls.lo + (j0 + nb) * ld + j0;
        p.ld_split = ld;
        p.split_scale = scales + SC_L;
        GPG_TRY(gemm_simt<float>(h, p, stream));
        tc::Launch g;                        // A22 -= A21 A21^T, lower tiles
        tc_params_clear(g);
        g.A.hi = Ls.hi; g.A.lo = Ls.lo; g.A.rows = N; g.A.cols = N; g.A.ld = ld;
        g.B = g.A;
        g.p.M = (int)rows; g.p.N = (int)rows; g.p.K = (int)nb; g.p.batch = 1;
        g.p.a_row0 = g.p.b_row0 = (int)(j0 + nb);
        g.p.a_col0 = g.p.b_col0 = (int)j0;
        g.p.tile_mode = GEMM_TILES_LOWER;
        g.p.epi = tc::EPI_STORE;
        g.p.scale_inv = scales + SC_INV_LL;
        g.p.C = A + (j0 + nb) * ld + (j0 + nb); g.p.ldc = ld;
        g.p.alpha = -1.f; g.p.beta = 1.f;
        GPG_TRY(tc::launch(h, g, stream));
    }
    return GPG_OK;
}

// ---------------------------------------------------------------------------------------------
// Linv = L^-1 by recursive doubling, W21 = -W22 (L21 W11), both products per level on tcgen05.
// Outputs: Linv (fp32, strict upper triangle zero), Ws = split(s_W Linv), WTs = split(s_W Linv^T).
// TTs is scratch.  Ls must hold split(s_L L) (cholesky_blocked_tc leaves it behind).
// ---------------------------------------------------------------------------------------------
static int trtri_tc(gpg_handle_s *h, const float *L, int64_t N, int64_t ld, float *Linv, TcPlanes Ls, TcPlanes Ws,
                    TcPlanes WTs, TcPlanes TTs, const float *scales, cudaStream_t stream) {
    constexpr int NB = 128;
    static bool attr_set = false;
    if (!attr_set) {
        GPG_CUDA_CHECK(cudaFuncSetAttribute(diag_block_kernel<float, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            diag_block_smem<float, NB>()));
        attr_set = true;
    }
    const size_t plane2 = 2 * (size_t)N * ld * sizeof(__half);
    GPG_CUDA_CHECK(cudaMemsetAsync(Linv, 0, (size_t)N * ld * sizeof(float), stream));
    GPG_CUDA_CHECK(cudaMemsetAsync(Ws.hi, 0, plane2, stream));
    GPG_CUDA_CHECK(cudaMemsetAsync(WTs.hi, 0, plane2, stream));
    const int nblk = (int)((N + NB - 1) / NB);
    DiagEmit em;
    em.Wh = Ws.hi; em.Wl = Ws.lo; em.WTh = WTs.hi; em.WTl = WTs.lo; em.lds = ld; em.scale_W = scales + SC_W;
    diag_block_kernel<float, NB><<<nblk, 256, diag_block_smem<float, NB>(), stream>>>(
        const_cast<float *>(L), ld, N, 0, 0, Linv, ld, (int64_t)NB * (ld + 1), 0, nullptr, em);
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
            const long long bs = 2 * b * (ld + 1);
            tc::Launch g1;                   // T = L21 W11, emitted transposed into TTs
            tc_params_clear(g1);
            g1.A.hi = Ls.hi; g1.A.lo = Ls.lo; g1.A.rows = N; g1.A.cols = N; g1.A.ld = ld;
            g1.B.hi = WTs.hi; g1.B.lo = WTs.lo; g1.B.rows = N; g1.B.cols = N; g1.B.ld = ld;
            g1.p.M = (int)rows; g1.p.N = (int)b; g1.p.K = (int)b; g1.p.batch = (int)batch;
            g1.p.a_row0 = (int)(s0 + b); g1.p.a_col0 = (int)s0; g1.p.a_bs = (int)(2 * b);
            g1.p.b_row0 = (int)s0; g1.p.b_col0 = (int)s0; g1.p.b_bs = (int)(2 * b);
            g1.p.kb_mode = GEMM_KB_N0;
            g1.p.epi = tc::EPI_STORE;
            g1.p.scale_inv = scales + SC_INV_LW;
            g1.p.alpha = 1.f;
            g1.p.T_hi = TTs.hi + s0 * ld + (s0 + b); g1.p.T_lo = TTs.lo + s0 * ld + (s0 + b);
            g1.p.ldt = ld; g1.p.t_bs = bs;
            g1.p.scale_out = scales + SC_T;
            GPG_TRY(tc::launch(h, g1, stream));
            tc::Launch g2;                   // W21 = -W22 T
            tc_params_clear(g2);
            g2.A.hi = Ws.hi; g2.A.lo = Ws.lo; g2.A.rows = N; g2.A.cols = N; g2.A.ld = ld;
            g2.B.hi = TTs.hi; g2.B.lo = TTs.lo; g2.B.rows = N; g2.B.cols = N; g2.B.ld = ld;
            g2.p.M = (int)rows; g2.p.N = (int)b; g2.p.K = (int)rows; g2.p.batch = (int)batch;
            g2.p.a_row0 = (int)(s0 + b); g2.p.a_col0 = (int)(s0 + b); g2.p.a_bs = (int)(2 * b);
            g2.p.b_row0 = (int)s0; g2.p.b_col0 = (int)(s0 + b); g2.p.b_bs = (int)(2 * b);
            g2.p.ke_mode = GEMM_KE_M;
            g2.p.epi = tc::EPI_STORE;
            g2.p.scale_inv = scales + SC_INV_WT;
            g2.p.alpha = -1.f;
            g2.p.C = Linv + (s0 + b) * ld + s0; g2.p.ldc = ld; g2.p.c_bs = bs;
            g2.p.S_hi = Ws.hi + (s0 + b) * ld + s0; g2.p.S_lo = Ws.lo + (s0 + b) * ld + s0;
            g2.p.lds = ld; g2.p.s_bs = bs;
            g2.p.T_hi = WTs.hi + s0 * ld + (s0 + b); g2.p.T_lo = WTs.lo + s0 * ld + (s0 + b);
            g2.p.ldt = ld; g2.p.t_bs = bs;
            g2.p.scale_out = scales + SC_W;
            GPG_TRY(tc::launch(h, g2, stream));
        }
    }
    return GPG_OK;
}

// Kinv = Linv^T Linv on the lower tiles (gradient of the marginal likelihood needs K^-1):
// Kinv[i][j] = sum_{k >= max(i,j)} WT[i][k] WT[j][k].
static int kinv_tc(gpg_handle_s *h, int64_t N, int64_t ld, TcPlanes WTs, float *Kinv, const float *scales,
                   cudaStream_t stream) {
    tc::Launch g;
    tc_params_clear(g);
    g.A.hi = WTs.hi; g.A.lo = WTs.lo; g.A.rows = N; g.A.cols = N; g.A.ld = ld;
    g.B = g.A;
    g.p.M = (int)N; g.p.N = (int)N; g.p.K = (int)N; g.p.batch = 1;
    g.p.kb_mode = GEMM_KB_MAXMN;
    g.p.tile_mode = GEMM_TILES_LOWER;
    g.p.epi = tc::EPI_STORE;
    g.p.scale_inv = scales + SC_INV_WW;
    g.p.alpha = 1.f;
    g.p.C = Kinv; g.p.ldc = ld;
    return tc::launch(h, g, stream);
}
