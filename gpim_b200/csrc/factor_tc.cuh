// factor_tc.cuh -- f32 factorisation on the tensor cores: a two-level blocked Cholesky whose panels and
// trailing SYRK updates run on the split-fp16 tcgen05 GEMM (gemm_tc.cuh), the triangular inverse by
// batched recursive doubling with both GEMMs of every level on the same kernel, and (optional
// alternative) a fully recursive Cholesky + inverse.
//
// Operand bookkeeping.  Every tensor-core operand is an fp16 hi/lo pair of K-major planes with the
// geometry of the N x ld matrix, produced where the fp32 value is born (no separate conversion
// passes):
//   Ls  = split(s_L L)        diag blocks by diag_block_kernel, panels by the panel GEMM's epilogue
//   Ws  = split(s_W L^-1)     diag blocks by diag_block_kernel, off-diagonal blocks by the epilogue
//   WTs = split(s_W L^-T)     of the level that computes them (plain and transposed emission)
//   TTs = split(s_T (L21 W11)^T)   intermediate of a doubling level, transposed emission
//   As  = split(s_A A)        not-yet-factored columns of K: by kmat_kernel, then by the SYRK epilogues
//                             (blocked algorithm: only the next panel's 128 columns are kept current)
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
    SC_A = 9,          // s_A   not-yet-factored part of K (Schur complements), |a_ij| <= v + nz
    SC_INV_AW = 10,    // 1 / (s_A s_W)
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
    const int ea = pow2_exp_below(16384.0f / (v + nz));
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
    scales[SC_A] = exp2f((float)ea);
    scales[SC_INV_AW] = exp2f((float)(-ea - ew));
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

// For an SPD matrix with a known lower bound lam_min on its spectrum (Kuu + jitter I: jitter; A' = I + B B^T / s2: 1):
//   |a_ij| <= max_i A_ii =: D,  |L_ij| <= sqrt(D),  |(L^-1)_ij| <= lam_min^-1/2,  |(L21 W11)_ij| <= sqrt(N D / lam_min).
// Fills every slot the cooperative panel Cholesky and trtri_tc read.
__global__ void __launch_bounds__(256) scales_from_diag_bound_kernel(const float *__restrict__ A, int64_t ld, int64_t N,
                                                                     float lam_min, float *__restrict__ scales) {
    __shared__ float red[8];
    float m = 0.f;
    for (int64_t i = threadIdx.x; i < N; i += 256) m = fmaxf(m, fabsf(A[i * ld + i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
        m = fmaxf(m, 1e-30f);
        const float lm = fmaxf(lam_min, 1e-30f);
        const int el = pow2_exp_below(16384.0f / sqrtf(m));
        const int ea = pow2_exp_below(16384.0f / m);
        const int ew = pow2_exp_below(16384.0f * sqrtf(lm));
        const int et = pow2_exp_below(16384.0f * sqrtf(lm) / sqrtf((float)N * m));
        for (int i = 0; i < SC_COUNT; ++i) scales[i] = 0.f;
        scales[SC_L] = exp2f((float)el);
        scales[SC_INV_LL] = exp2f((float)(-2 * el));
        scales[SC_A] = exp2f((float)ea);
        scales[SC_W] = exp2f((float)ew);
        scales[SC_INV_AW] = exp2f((float)(-ea - ew));
        scales[SC_T] = exp2f((float)et);
        scales[SC_INV_LW] = exp2f((float)(-el - ew));
        scales[SC_INV_WT] = exp2f((float)(-ew - et));
        scales[SC_INV_WW] = exp2f((float)(-2 * ew));
    }
}

template <typename E>
__global__ void __launch_bounds__(256) zero_band_kernel(E *__restrict__ p, int64_t ld, int64_t N, int64_t lo_off,
                                                        int64_t hi_off) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= N) return;
    const int64_t c0 = max((int64_t)0, i + lo_off), c1 = min(ld, i + hi_off);
    for (int64_t c = c0 + lane; c < c1; c += 32) p[i * ld + c] = E(0);
}

struct TcPlanes {             // hi plane followed by lo plane, each N x ld halves
    __half *hi = nullptr, *lo = nullptr;
    TcPlanes() {}
    TcPlanes(void *base, int64_t N, int64_t ld) : hi((__half *)base), lo((__half *)base + (size_t)N * ld) {}
};

static inline void tc_params_clear(tc::Launch &g) { memset(&g.p, 0, sizeof(g.p)); }

// ---------------------------------------------------------------------------------------------
// Two-level blocked right-looking Cholesky.  Inner block NB = 128: the diagonal block is factored and
// inverted in one CTA (diag_block_kernel), the panel below becomes A21 inv(L11)^T on the SIMT GEMM
// (K = 128, also emits the panel's fp16 split), and the trailing update A22 -= A21 A21^T runs on
// tcgen05 -- but only over the columns of the current OUTER panel (width NB2): the rest of the trailing
// matrix receives one K = NB2 update per outer panel, so the fp32 matrix is read-modified-written
// N / NB2 times instead of N / 128 times and the big updates are tensor-bound, not HBM-bound.
// Accumulation chains in TMEM stay <= NB2 long (the tensor core truncates when it accumulates).
// ---------------------------------------------------------------------------------------------
// panel_mode 1 (default; needs As / Ws, else falls back to 2): the panel runs on tcgen05 through the block's
// inverse -- As holds split(s_A A) of the NEXT panel's 128 columns (written by kmat_kernel for the first panel,
// then by the update that last touched them), Ws receives the inverse of each diagonal block.
// panel_mode 0: forward substitution against the diagonal factor (panel_trsm_kernel; diag_block_kernel then
// computes no inverse) -- backward stable but slower (one thread per row, 8128 dependent FMAs).
// panel_mode 2: SIMT GEMM through the inverse.
static int cholesky_blocked_tc(gpg_handle_s *h, float *A, int64_t N, int64_t ld, int32_t *info, int reset_info,
                               float *dinv, TcPlanes Ls, const float *scales, cudaStream_t stream,
                               TcPlanes As = TcPlanes(), TcPlanes Ws = TcPlanes()) {
    constexpr int NB = 128;
    static bool attr_set = false;
    if (!attr_set) {
        GPG_CUDA_CHECK(cudaFuncSetAttribute(diag_block_kernel<float, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            diag_block_smem<float, NB>()));
        GPG_CUDA_CHECK(cudaFuncSetAttribute(panel_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, panel_trsm_smem()));
        attr_set = true;
    }
    int panel_mode = h->opt_panel_mode;
    // mode 3 needs the W planes, 16-byte aligned fp32 rows and a device on which the cooperative grid fits
    if (panel_mode == 3 && !(Ws.hi && (ld % 8) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && h->sm_count >= 4))
        panel_mode = 1;
    if (panel_mode == 1 && !(As.hi && Ws.hi)) panel_mode = 2;
    // the cooperative panel keeps the update sums of one row block in the 512 TMEM columns: at most 4 block columns
    const int64_t NB2 = panel_mode == 3 ? std::min<int64_t>(h->opt_outer_panel, 512) : h->opt_outer_panel;
    if (reset_info) GPG_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), stream));
    auto syrk = [&](int64_t row0, int64_t col0, int64_t k0, int64_t rows, int64_t cols, int64_t kk,
                    cudaStream_t on = nullptr, int max_ctas = 0) -> int {
        // A[row0.., col0..] -= L[row0.., k0..k0+kk) L[col0.., k0..k0+kk)^T on the tiles that touch row >= col
        if (rows <= 0 || cols <= 0) return GPG_OK;
        tc::Launch g;
        tc_params_clear(g);
        g.max_ctas = max_ctas;
        g.A.hi = Ls.hi; g.A.lo = Ls.lo; g.A.rows = N; g.A.cols = N; g.A.ld = ld;
        g.B = g.A;
        g.p.M = (int)rows; g.p.N = (int)cols; g.p.K = (int)kk; g.p.batch = 1;
        g.p.a_row0 = (int)row0; g.p.b_row0 = (int)col0;
        g.p.a_col0 = g.p.b_col0 = (int)k0;
        g.p.tile_mode = GEMM_TILES_LOWER;            // row0 == col0 in every call: C's origin is on the diagonal
        g.p.epi = tc::EPI_STORE;
        g.p.scale_inv = scales + SC_INV_LL;
        g.p.C = A + row0 * ld + col0; g.p.ldc = ld;
        g.p.alpha = -1.f; g.p.beta = 1.f;
        if (panel_mode == 1) {                       // the first 128 columns are the next panel: keep their split current
            g.p.S_hi = As.hi + row0 * ld + col0; g.p.S_lo = As.lo + row0 * ld + col0; g.p.lds = ld;
            g.p.s_ncols = NB;
            g.p.scale_out = scales + SC_A;
        }
        return tc::launch(h, g, on ? on : stream);
    };
    bool side_pending = false;
    for (int64_t J0 = 0; J0 < N; J0 += NB2) {
        const int64_t Jend = std::min<int64_t>(N, J0 + NB2);
        if (panel_mode == 3) {
            // the whole column panel (diagonal blocks, panel products, updates inside the panel) in ONE cooperative
            // kernel with a device-side dependency chain (chol_panel.cuh); then the K = panel-width update of the rest
            const int64_t nbp = (Jend - J0 + NB - 1) / NB;
            const int64_t rows_blk = (N - J0 + NB - 1) / NB;                  // row blocks including the diagonal one
            if (rows_blk <= 1) {                                              // a lone last block: nothing below it
                DiagEmit em;
                em.Lh = Ls.hi; em.Ll = Ls.lo; em.lds = ld; em.scale_L = scales + SC_L;
                diag_block_kernel<float, NB><<<1, 256, diag_block_smem<float, NB>(), stream>>>(
                    A, ld, N, J0, 1, (float *)nullptr, (int64_t)NB, (int64_t)0, 1, info, em);
                GPG_LAUNCH_CHECK(h);
                break;
            }
            cpanel::Args pa;
            pa.A = A; pa.ld = ld; pa.N = N; pa.J0 = J0; pa.nbp = (int)std::min<int64_t>(nbp, 4);
            pa.Ls_hi = Ls.hi; pa.Ls_lo = Ls.lo; pa.Ws_hi = Ws.hi; pa.Ws_lo = Ws.lo;
            pa.scales = scales; pa.sc_A = SC_A; pa.sc_W = SC_W; pa.sc_L = SC_L; pa.sc_inv_AW = SC_INV_AW; pa.sc_inv_LL = SC_INV_LL;
            pa.info = info;
            GPG_TRY(gpg_tc_counters(h, stream, cpanel::NFLAGS, &pa.flags));
            CUtensorMap mLhi, mLlo, mWhi, mWlo;
            GPG_TRY(tc::make_tensor_map(&mLhi, Ls.hi, N, N, ld, NB));
            GPG_TRY(tc::make_tensor_map(&mLlo, Ls.lo, N, N, ld, NB));
            GPG_TRY(tc::make_tensor_map(&mWhi, Ws.hi, N, N, ld, NB));
            GPG_TRY(tc::make_tensor_map(&mWlo, Ws.lo, N, N, ld, NB));
            // chain + workers.  With look-ahead the rest of the trailing update of the PREVIOUS panel runs next to this
            // kernel on the side stream, so the workers are capped at half the SMs (a worker then owns several row blocks)
            // Look-ahead pays where the chain dominates, i.e. once one worker per row block leaves at least half of the
            // SMs to the side stream; further up the matrix the trailing update is the bulk of the work, wants every
            // SM, and runs in stream order.
            const int64_t rest = N - Jend - NB2;                              // trailing columns beyond the next panel
            const bool small = rows_blk - 1 <= h->sm_count / 2;
            const bool ahead = h->opt_lookahead && small && rest > 0;
            int wcap = h->opt_panel_workers > 0 ? h->opt_panel_workers : h->sm_count - 1;
            wcap = std::max(3, std::min(wcap, h->sm_count - 1));
            const int grid = 1 + (int)std::min<int64_t>(rows_blk - 1, wcap);
            void *kargs[] = {&mLhi, &mLlo, &mWhi, &mWlo, &pa};
            GPG_CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)cpanel::chol_panel_kernel, dim3(grid), dim3(cpanel::NUM_THREADS),
                                                       kargs, (size_t)cpanel::SMEM_BYTES, stream));
            GPG_LAUNCH_CHECK(h);
            if (!ahead && !side_pending) {
                GPG_TRY(syrk(Jend, Jend, J0, N - Jend, N - Jend, Jend - J0));
                continue;
            }
            if (!h->side_stream) {
                GPG_CUDA_CHECK(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
                GPG_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
                GPG_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming));
            }
            if (ahead) {           // the rest, K = panel width, on the side stream as soon as this panel is done ...
                GPG_CUDA_CHECK(cudaEventRecord(h->ev_fork, stream));
                GPG_CUDA_CHECK(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
            }
            // ... the next panel's columns here -- after the previous panel's rest, which writes the same tiles
            if (side_pending) GPG_CUDA_CHECK(cudaStreamWaitEvent(stream, h->ev_side, 0));
            GPG_TRY(syrk(Jend, Jend, J0, N - Jend, std::min<int64_t>(NB2, N - Jend), Jend - J0));
            if (ahead) {
                // next to it runs the NEXT panel's kernel: chain + one worker per row block below Jend
                const int next_grid = 1 + (int)std::min<int64_t>((N - Jend + NB - 1) / NB - 1, wcap);
                GPG_TRY(syrk(Jend + NB2, Jend + NB2, J0, rest, rest, Jend - J0, h->side_stream, std::max(16, h->sm_count - next_grid)));
                GPG_CUDA_CHECK(cudaEventRecord(h->ev_side, h->side_stream));
                side_pending = true;
            }
            continue;
        }
        // inner update after the panel of block column j0.  Right-looking (opt_inner_left == 0): all remaining
        // columns of the outer panel, K = nb.  Left-looking (default): only the NEXT block column, but with the
        // contributions of every inner panel so far (K = j0 + nb - J0) -- half the epilogue per step and a third
        // of the fp32 read-modify-write traffic; the columns further right catch up when their turn comes.
        auto inner_update = [&](int64_t j0, int64_t nb, int64_t rows) -> int {
            if (h->opt_inner_left)
                return syrk(j0 + nb, j0 + nb, J0, rows, std::min<int64_t>(NB, Jend - (j0 + nb)), j0 + nb - J0);
            return syrk(j0 + nb, j0 + nb, j0, rows, Jend - (j0 + nb), nb);
        };
        for (int64_t j0 = J0; j0 < Jend; j0 += NB) {
            const int64_t nb = std::min<int64_t>(NB, N - j0);
            DiagEmit em;
            em.Lh = Ls.hi; em.Ll = Ls.lo; em.lds = ld; em.scale_L = scales + SC_L;
            if (panel_mode == 1) { em.Wh = Ws.hi; em.Wl = Ws.lo; em.scale_W = scales + SC_W; }
            GPG_CUDA_CHECK(launch_pdl(diag_block_kernel<float, NB>, dim3(1), dim3(256), (size_t)diag_block_smem<float, NB>(),
                                      stream, A, ld, N, j0, 1, panel_mode == 2 ? dinv : (float *)nullptr, (int64_t)NB,
                                      (int64_t)0, 1, info, em));
            GPG_LAUNCH_CHECK(h);
            const int64_t rows = N - j0 - nb;
            if (rows <= 0) break;
            float *A21 = A + (j0 + nb) * ld + j0;
            if (panel_mode == 0) {                   // A21 <- A21 L11^-T by forward substitution, fp32 in place + Ls planes
                GPG_CUDA_CHECK(launch_pdl(panel_trsm_kernel, dim3((unsigned)((rows + NB - 1) / NB)), dim3(PANEL_NB),
                                          (size_t)panel_trsm_smem(), stream, A, ld, j0, rows, Ls.hi, Ls.lo, ld,
                                          scales + SC_L));
                GPG_LAUNCH_CHECK(h);
                GPG_TRY(inner_update(j0, nb, rows));
                continue;
            }
            if (panel_mode == 1) {                   // A21 <- A21 inv(L11)^T on tcgen05, fp32 in place + Ls planes
                tc::Launch g;
                tc_params_clear(g);
                g.A.hi = As.hi; g.A.lo = As.lo; g.A.rows = N; g.A.cols = N; g.A.ld = ld;
                g.B.hi = Ws.hi; g.B.lo = Ws.lo; g.B.rows = N; g.B.cols = N; g.B.ld = ld;
                g.p.M = (int)rows; g.p.N = (int)nb; g.p.K = (int)nb; g.p.batch = 1;
                g.p.a_row0 = (int)(j0 + nb); g.p.a_col0 = (int)j0;
                g.p.b_row0 = (int)j0; g.p.b_col0 = (int)j0;
                g.p.epi = tc::EPI_STORE;
                g.p.scale_inv = scales + SC_INV_AW;
                g.p.alpha = 1.f;
                g.p.C = A21; g.p.ldc = ld;
                g.p.S_hi = Ls.hi + (j0 + nb) * ld + j0; g.p.S_lo = Ls.lo + (j0 + nb) * ld + j0; g.p.lds = ld;
                g.p.scale_out = scales + SC_L;
                GPG_TRY(tc::launch(h, g, stream));
                GPG_TRY(inner_update(j0, nb, rows));
                continue;
            }
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
            GPG_TRY((gemm_simt<float, GemmCfgPanelF32>(h, p, stream)));
            GPG_TRY(inner_update(j0, nb, rows));
        }
        // outer update: everything right of the panel, K = panel width
        GPG_TRY(syrk(Jend, Jend, J0, N - Jend, N - Jend, Jend - J0));
    }
    if (side_pending) GPG_CUDA_CHECK(cudaStreamWaitEvent(stream, h->ev_side, 0));
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
    cudaStream_t stream_main = stream;
    static bool attr_set = false;
    if (!attr_set) {
        GPG_CUDA_CHECK(cudaFuncSetAttribute(diag_block_kernel<float, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            diag_block_smem<float, NB>()));
        attr_set = true;
    }
    // zero what is structurally zero AND read: Linv's strict upper triangle, and the bands of Ws / WTs next to
    // the diagonal that triangular k-ranges over-read at tile granularity
    {
        const unsigned gz = (unsigned)((N + 7) / 8);
        constexpr int64_t BAND = 384;      // >= tc::BN + tc::BK
        zero_band_kernel<float><<<gz, 256, 0, stream>>>(Linv, ld, N, 1, ld);
        GPG_LAUNCH_CHECK(h);
        for (__half *pl : {Ws.hi, Ws.lo}) {
            zero_band_kernel<__half><<<gz, 256, 0, stream>>>(pl, ld, N, 1, BAND);
            GPG_LAUNCH_CHECK(h);
        }
        for (__half *pl : {WTs.hi, WTs.lo}) {
            zero_band_kernel<__half><<<gz, 256, 0, stream>>>(pl, ld, N, -BAND, 0);
            GPG_LAUNCH_CHECK(h);
        }
    }
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
        // the ragged last pair is independent of the full pairs of its level: its two (small, latency-bound) GEMMs run
        // on the side stream next to them
        const bool fork = npairs_full > 0 && rem_rows > 0;
        if (fork) {
            if (!h->side_stream) {
                GPG_CUDA_CHECK(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
                GPG_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
                GPG_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming));
            }
            GPG_CUDA_CHECK(cudaEventRecord(h->ev_fork, stream_main));
            GPG_CUDA_CHECK(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        }
        for (int pass = 0; pass < 2; ++pass) {
            const int64_t s0 = pass == 0 ? 0 : rem_start;
            const int64_t rows = pass == 0 ? b : rem_rows;
            const int64_t batch = pass == 0 ? npairs_full : (rem_rows > 0 ? 1 : 0);
            if (batch == 0 || rows == 0) continue;
            cudaStream_t stream = (fork && pass == 1) ? h->side_stream : stream_main;
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
        if (fork) {
            GPG_CUDA_CHECK(cudaEventRecord(h->ev_side, h->side_stream));
            GPG_CUDA_CHECK(cudaStreamWaitEvent(stream_main, h->ev_side, 0));
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

// ---------------------------------------------------------------------------------------------
// Recursive Cholesky + inverse, every contraction on tcgen05 with K = half the node size, so the
// fp32 matrix is read-modified-written O(log N) times instead of N / 128 times:
//   node [s, s + n), split at h:   (L11, W11) = node(s, h)
//                                  L21 = A21 W11^T                    (panel through the inverse)
//                                  A22 -= L21 L21^T                   (SYRK, lower tiles)
//                                  (L22, W22) = node(s + h, n - h)
//                                  W21 = -W22 (L21 W11)               (two products)
//   leaf (n <= 128): diag_block_kernel (Cholesky + inverse of the block in one CTA).
// In: A fp32 lower triangle and As = split(s_A A) (lower tiles).  Out: L in A (+ Ls), W = L^-1 in Linv
// (+ Ws, WTs).  Rf is an fp32 N x ld scratch.  Linv's strict upper triangle and the bands of Ws / WTs next to the diagonal that the
// triangular k-ranges over-read at tile granularity are zeroed first.
// ---------------------------------------------------------------------------------------------
struct PotrfCtx {
    gpg_handle_s *h;
    float *A, *Linv, *Rf;          // Rf: fp32 N x ld scratch (first panel estimate)
    int64_t N, ld;
    TcPlanes As, Ls, Ws, WTs, TTs;
    const float *scales;
    int32_t *info;
    cudaStream_t stream;
};

static int potrf_inv_node(const PotrfCtx &c, int64_t s, int64_t n) {
    constexpr int NB = 128;
    const int64_t ld = c.ld;
    if (n <= NB) {
        DiagEmit em;
        em.Lh = c.Ls.hi; em.Ll = c.Ls.lo; em.Wh = c.Ws.hi; em.Wl = c.Ws.lo; em.WTh = c.WTs.hi; em.WTl = c.WTs.lo;
        em.lds = ld; em.scale_L = c.scales + SC_L; em.scale_W = c.scales + SC_W;
        diag_block_kernel<float, NB><<<1, 256, diag_block_smem<float, NB>(), c.stream>>>(
            c.A, ld, s + n, s, 1, c.Linv + s * (ld + 1), ld, 0, 0, c.info, em);
        GPG_LAUNCH_CHECK(c.h);
        return GPG_OK;
    }
    const int64_t nblk = (n + NB - 1) / NB;
    const int64_t h = ((nblk + 1) / 2) * NB, r2 = n - h;
    GPG_TRY(potrf_inv_node(c, s, h));
    auto whole = [&](tc::SplitMat &m, const TcPlanes &pl) { m.hi = pl.hi; m.lo = pl.lo; m.rows = c.N; m.cols = c.N; m.ld = ld; };
    // Panel L21 = A21 L11^-T through the explicit inverse W11, plus one step of iterative refinement
    // against L11 itself, which restores the backward stability of a triangular solve:
    //   L0 = A21 W11^T;   R = A21 - L0 L11^T;   L21 = L0 + R W11^T.
    const bool refine = c.h->opt_panel_refine != 0;
    {   // L0 = A21 W11^T  -> Rf (fp32; straight into the factor when not refining), Ls
        tc::Launch g;
        tc_params_clear(g);
        whole(g.A, c.As); whole(g.B, c.Ws);
        g.p.M = (int)r2; g.p.N = (int)h; g.p.K = (int)h; g.p.batch = 1;
        g.p.a_row0 = (int)(s + h); g.p.a_col0 = (int)s;
        g.p.b_row0 = (int)s; g.p.b_col0 = (int)s;
        g.p.ke_mode = GEMM_KE_N;
        g.p.epi = tc::EPI_STORE;
        g.p.scale_inv = c.scales + SC_INV_AW;
        g.p.alpha = 1.f;
        g.p.C = (refine ? c.Rf : c.A) + (s + h) * ld + s; g.p.ldc = ld;
        g.p.S_hi = c.Ls.hi + (s + h) * ld + s; g.p.S_lo = c.Ls.lo + (s + h) * ld + s; g.p.lds = ld;
        g.p.scale_out = c.scales + SC_L;
        GPG_TRY(tc::launch(c.h, g, c.stream));
    }
    if (refine) {   // R = A21 - L0 L11^T  -> fp32 in place of A21, As
        tc::Launch g;
        tc_params_clear(g);
        whole(g.A, c.Ls); whole(g.B, c.Ls);
        g.p.M = (int)r2; g.p.N = (int)h; g.p.K = (int)h; g.p.batch = 1;
        g.p.a_row0 = (int)(s + h); g.p.a_col0 = (int)s;
        g.p.b_row0 = (int)s; g.p.b_col0 = (int)s;
        g.p.ke_mode = GEMM_KE_N;
        g.p.epi = tc::EPI_STORE;
        g.p.scale_inv = c.scales + SC_INV_LL;
        g.p.alpha = -1.f; g.p.beta = 1.f;
        g.p.C = c.A + (s + h) * ld + s; g.p.ldc = ld;
        g.p.S_hi = c.As.hi + (s + h) * ld + s; g.p.S_lo = c.As.lo + (s + h) * ld + s; g.p.lds = ld;
        g.p.scale_out = c.scales + SC_A;
        GPG_TRY(tc::launch(c.h, g, c.stream));
    }
    if (refine) {   // L21 = L0 + R W11^T  -> Rf, fp32 into the factor, Ls
        tc::Launch g;
        tc_params_clear(g);
        whole(g.A, c.As); whole(g.B, c.Ws);
        g.p.M = (int)r2; g.p.N = (int)h; g.p.K = (int)h; g.p.batch = 1;
        g.p.a_row0 = (int)(s + h); g.p.a_col0 = (int)s;
        g.p.b_row0 = (int)s; g.p.b_col0 = (int)s;
        g.p.ke_mode = GEMM_KE_N;
        g.p.epi = tc::EPI_STORE;
        g.p.scale_inv = c.scales + SC_INV_AW;
        g.p.alpha = 1.f; g.p.beta = 1.f;
        g.p.C = c.Rf + (s + h) * ld + s; g.p.ldc = ld;
        g.p.C2 = c.A + (s + h) * ld + s; g.p.ldc2 = ld;
        g.p.S_hi = c.Ls.hi + (s + h) * ld + s; g.p.S_lo = c.Ls.lo + (s + h) * ld + s; g.p.lds = ld;
        g.p.scale_out = c.scales + SC_L;
        GPG_TRY(tc::launch(c.h, g, c.stream));
    }
    // A22 -= L21 L21^T (lower tiles) -> fp32 in place, As.  The sums on and near the diagonal are all-positive
    // and the tensor core truncates when it adds into the TMEM accumulator, a bias that grows with the length
    // of the chain (about -6e-9 K relative): K is therefore cut into chunks whose partial results meet in an
    // fp32 round-to-nearest add (beta = 1).
    {
        const int64_t kc = c.h->opt_syrk_chunk > 0 ? c.h->opt_syrk_chunk : h;
        for (int64_t k0 = 0; k0 < h; k0 += kc) {
            const int64_t kk = std::min<int64_t>(kc, h - k0);
            const bool last = k0 + kk >= h;
            tc::Launch g;
            tc_params_clear(g);
            whole(g.A, c.Ls); g.B = g.A;
            g.p.M = (int)r2; g.p.N = (int)r2; g.p.K = (int)kk; g.p.batch = 1;
            g.p.a_row0 = g.p.b_row0 = (int)(s + h);
            g.p.a_col0 = g.p.b_col0 = (int)(s + k0);
            g.p.tile_mode = GEMM_TILES_LOWER;
            g.p.epi = tc::EPI_STORE;
            g.p.scale_inv = c.scales + SC_INV_LL;
            g.p.alpha = -1.f; g.p.beta = 1.f;
            g.p.C = c.A + (s + h) * (ld + 1); g.p.ldc = ld;
            if (last && r2 > NB) {       // a leaf reads the fp32 block; only larger nodes need the split
                g.p.S_hi = c.As.hi + (s + h) * (ld + 1); g.p.S_lo = c.As.lo + (s + h) * (ld + 1); g.p.lds = ld;
                g.p.scale_out = c.scales + SC_A;
            }
            GPG_TRY(tc::launch(c.h, g, c.stream));
        }
    }
    GPG_TRY(potrf_inv_node(c, s + h, r2));
    {   // T = L21 W11, emitted transposed into TTs
        tc::Launch g;
        tc_params_clear(g);
        whole(g.A, c.Ls); whole(g.B, c.WTs);
        g.p.M = (int)r2; g.p.N = (int)h; g.p.K = (int)h; g.p.batch = 1;
        g.p.a_row0 = (int)(s + h); g.p.a_col0 = (int)s;
        g.p.b_row0 = (int)s; g.p.b_col0 = (int)s;
        g.p.kb_mode = GEMM_KB_N0;
        g.p.epi = tc::EPI_STORE;
        g.p.scale_inv = c.scales + SC_INV_LW;
        g.p.alpha = 1.f;
        g.p.T_hi = c.TTs.hi + s * ld + (s + h); g.p.T_lo = c.TTs.lo + s * ld + (s + h); g.p.ldt = ld;
        g.p.scale_out = c.scales + SC_T;
        GPG_TRY(tc::launch(c.h, g, c.stream));
    }
    {   // W21 = -W22 T -> Linv, Ws, WTs
        tc::Launch g;
        tc_params_clear(g);
        whole(g.A, c.Ws); whole(g.B, c.TTs);
        g.p.M = (int)r2; g.p.N = (int)h; g.p.K = (int)r2; g.p.batch = 1;
        g.p.a_row0 = (int)(s + h); g.p.a_col0 = (int)(s + h);
        g.p.b_row0 = (int)s; g.p.b_col0 = (int)(s + h);
        g.p.ke_mode = GEMM_KE_M;
        g.p.epi = tc::EPI_STORE;
        g.p.scale_inv = c.scales + SC_INV_WT;
        g.p.alpha = -1.f;
        g.p.C = c.Linv + (s + h) * ld + s; g.p.ldc = ld;
        g.p.S_hi = c.Ws.hi + (s + h) * ld + s; g.p.S_lo = c.Ws.lo + (s + h) * ld + s; g.p.lds = ld;
        g.p.T_hi = c.WTs.hi + s * ld + (s + h); g.p.T_lo = c.WTs.lo + s * ld + (s + h); g.p.ldt = ld;
        g.p.scale_out = c.scales + SC_W;
        GPG_TRY(tc::launch(c.h, g, c.stream));
    }
    return GPG_OK;
}

static int potrf_inv_tc(gpg_handle_s *h, float *A, int64_t N, int64_t ld, float *Linv, float *Rf, int32_t *info,
                        int reset_info, TcPlanes As, TcPlanes Ls, TcPlanes Ws, TcPlanes WTs, TcPlanes TTs,
                        const float *scales, cudaStream_t stream) {
    constexpr int NB = 128;
    static bool attr_set = false;
    if (!attr_set) {
        GPG_CUDA_CHECK(cudaFuncSetAttribute(diag_block_kernel<float, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            diag_block_smem<float, NB>()));
        attr_set = true;
    }
    if (reset_info) GPG_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), stream));
    const unsigned gz = (unsigned)((N + 7) / 8);
    constexpr int64_t BAND = 384;      // >= tc::BN + tc::BK: what a triangular k-range can over-read
    zero_band_kernel<float><<<gz, 256, 0, stream>>>(Linv, ld, N, 1, ld);
    GPG_LAUNCH_CHECK(h);
    for (__half *pl : {Ws.hi, Ws.lo, Ls.hi, Ls.lo}) {
        zero_band_kernel<__half><<<gz, 256, 0, stream>>>(pl, ld, N, 1, BAND);
        GPG_LAUNCH_CHECK(h);
    }
    for (__half *pl : {WTs.hi, WTs.lo}) {
        zero_band_kernel<__half><<<gz, 256, 0, stream>>>(pl, ld, N, -BAND, 0);
        GPG_LAUNCH_CHECK(h);
    }
    PotrfCtx c{h, A, Linv, Rf, N, ld, As, Ls, Ws, WTs, TTs, scales, info, stream};
    return potrf_inv_node(c, 0, N);
}
