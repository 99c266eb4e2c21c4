// sparse_driver.cuh -- host-side sequencing of the inducing-point GP (kernels: sparse.cuh) and the extern "C"
// gpg_sparse_* entry points of include/gpgrid.h.  Included at the end of gpgrid.cu (not a stand-alone header).
// Reference: reconstructor(sparse=True), gpr.py:145-155,198-199 over pyro SparseGPRegression (VFE); sparse.cuh has the
// algebra.  The m x m factorisations run on the blocked SIMT drivers of factor.cuh; the m x m x N and m^3 products on the
// SIMT GEMM (fp64, small problems) or the tcgen05 split-fp16 GEMM (fp32, m >= 512 and N >= 2048).
#pragma once
constexpr int SGP_KGRAD_CHUNKS = 16;     // most column chunks of the m x N kernel-derivative reduction

template <typename T> struct SgpBufs {
    double *gxup = nullptr;      // double partials of the inducing-input gradient: [1 + SGP_KGRAD_CHUNKS][m][d]
    // tcgen05 route of the two m x m factorisations (cooperative panel Cholesky + trtri_tc): four plane pairs and scales
    unsigned char *fplanes = nullptr;
    float *fscales = nullptr;
    int64_t m = 0, N = 0, ldm = 0, ldn = 0;
    int nz = 1;                  // SIMT route: split-K factor of S = B B^T; ldn = nz * kchunk (columns [N, ldn) of B are kept zero)
    int64_t kchunk = 0;
    T *Spart;
    // tcgen05 route (fp32, m >= 512 and N >= 2048, see sgp_uses_tc): fp16 hi/lo planes of the operands and their scales
    bool tc = false;
    int64_t ldk = 0;             // leading dimension of the N x m operands
    float *Kfu = nullptr, *scales = nullptr;
    __half *Kfus = nullptr, *Bts = nullptr, *Uis = nullptr, *T2s = nullptr, *Bs = nullptr;      // hi plane, then lo plane
    __half *P1 = nullptr, *P2 = nullptr, *P3 = nullptr;      // m x m operand planes of the gradient chain
    float *amax = nullptr;       // per-block maxima of the operand-scale measurement
    int nzt = 0;                 // S = B B^T on tcgen05: split-K batches of SGP_TC_KCHUNK columns, partial products in SpartTc
    float *SpartTc = nullptr;
    T *Luu, *Ui, *tmp, *Kuf, *B, *S, *Ap, *LA, *LAi, *Ainv, *Phi, *H, *T1, *Guu, *T2;   // Kuf doubles as dF/dKuf
    T *beta, *c0, *a0, *a, *w, *rho, *dinv, *theta0, *gxu, *grad, *loss, *theta, *m1, *m2;
    double *sc, *partA, *partB;
    int nb;
    FitState *st;
};

// S = B B^T has only (m / tile)^2 / 2 output tiles against a contraction length of N: it is split along K into nz
// batches of kchunk columns (nz * kchunk >= N; the tail columns of B are zero) whose partial products
// sgp_form_A_kernel adds up.
static void sgp_split(int64_t N, int &nz, int64_t &kchunk) {
    nz = (int)std::min<int64_t>(16, std::max<int64_t>(1, N / 1024));
    kchunk = (int64_t)gpg_align_up((size_t)((N + nz - 1) / nz), 64);
}

// The three m x m x N products -- B = Ui k(Xu, X), S = B B^T and dF/dKuf = T2 B -- go to the split-fp16 tcgen05 GEMM when
// they are large enough to matter (fp32 only).
// S = B B^T is a sum of N same-sign-dominated products per diagonal entry; the TMEM accumulation truncates (a bias of
// about -6e-9 per accumulated term), so the contraction is cut into batches of 512 columns whose
// partial products are added in double by sgp_form_A_kernel: the bias stays at 3e-6 relative, the level of the
// rounding error of an fp32 SIMT accumulation of this length.
constexpr int SGP_TC_KCHUNK = 512;

template <typename T> static bool sgp_uses_tc(const gpg_handle_s *h, int64_t m, int64_t N) {
    return std::is_same<T, float>::value && h->opt_gemm_path != 1 && (h->opt_gemm_path == 2 || (m >= 512 && N >= 2048));
}

template <typename T> static size_t sgp_ws_bytes(const gpg_handle_s *h, int64_t m, int64_t N, int d) {
    constexpr int NB = GemmCfg<T>::BN;
    int nz; int64_t kchunk;
    sgp_split(N, nz, kchunk);
    const bool tcp = sgp_uses_tc<T>(h, m, N);
    const size_t ldk_ = gpg_align_up((size_t)m, 64), ldm_ = ldk_;
    const size_t nzt_ = (size_t)((N + SGP_TC_KCHUNK - 1) / SGP_TC_KCHUNK);
    const size_t tc_bytes = tcp ? bump_size({(size_t)N * ldk_ * 4, (size_t)N * ldk_ * 4, (size_t)N * ldk_ * 4,
                                             (size_t)m * ldm_ * 4, (size_t)m * ldm_ * 4, SGP_S_COUNT * sizeof(float),
                                             (size_t)m * (size_t)(nz * kchunk) * 4, nzt_ * m * ldm_ * 4,
                                             (size_t)m * ldm_ * 4, (size_t)m * ldm_ * 4, (size_t)m * ldm_ * 4,
                                             SGP_AMAX_BLOCKS * sizeof(float), 4 * (size_t)m * ldm_ * 4,
                                             SC_COUNT * sizeof(float)}) : 0;
    const int64_t ldm = gpg_align_up((size_t)m, 64), ldn = nz * kchunk;
    const size_t mm = (size_t)m * ldm * sizeof(T), mn = (size_t)m * ldn * sizeof(T), mv = (size_t)m * sizeof(T);
    const size_t nb = (size_t)((m + 7) / 8);
    return bump_size({(size_t)nz * mm, mm, mm, mm, mn, mn, mm, mm, mm, mm, mm, mm, mm, mm, mm, mm,
                      mv, mv, mv, mv, mv, (size_t)N * sizeof(T), NB * NB * sizeof(T), GPG_MAX_P * sizeof(T),
                      mv * d, GPG_MAX_P * sizeof(T), sizeof(T), GPG_MAX_P * sizeof(T), mv * d, mv * d,
                      SGP_SC_COUNT * sizeof(double), nb * GPG_MAX_P * sizeof(double),
                      SGP_KGRAD_CHUNKS * nb * GPG_MAX_P * sizeof(double), (size_t)(1 + SGP_KGRAD_CHUNKS) * m * d * sizeof(double),
                      sizeof(FitState)}) + tc_bytes;
}

template <typename T> static SgpBufs<T> sgp_carve(const gpg_handle_s *h, void *ws, int64_t m, int64_t N, int d) {
    constexpr int NB = GemmCfg<T>::BN;
    Bump b(ws);
    SgpBufs<T> s;
    s.m = m; s.N = N;
    sgp_split(N, s.nz, s.kchunk);
    s.ldm = gpg_align_up((size_t)m, 64); s.ldn = s.nz * s.kchunk;
    const size_t mm = (size_t)m * s.ldm, mn = (size_t)m * s.ldn;
    s.Spart = b.take<T>((size_t)s.nz * mm);
    s.Luu = b.take<T>(mm); s.Ui = b.take<T>(mm); s.tmp = b.take<T>(mm);
    s.Kuf = b.take<T>(mn); s.B = b.take<T>(mn);
    s.S = b.take<T>(mm); s.Ap = b.take<T>(mm); s.LA = b.take<T>(mm); s.LAi = b.take<T>(mm); s.Ainv = b.take<T>(mm);
    s.Phi = b.take<T>(mm); s.H = b.take<T>(mm); s.T1 = b.take<T>(mm); s.Guu = b.take<T>(mm); s.T2 = b.take<T>(mm);
    s.beta = b.take<T>(m); s.c0 = b.take<T>(m); s.a0 = b.take<T>(m); s.a = b.take<T>(m); s.w = b.take<T>(m);
    s.rho = b.take<T>(N);
    s.dinv = b.take<T>(NB * NB);
    s.theta0 = b.take<T>(GPG_MAX_P);
    s.gxu = b.take<T>((size_t)m * d);
    s.grad = b.take<T>(GPG_MAX_P); s.loss = b.take<T>(1); s.theta = b.take<T>(GPG_MAX_P);
    s.m1 = b.take<T>((size_t)m * d); s.m2 = b.take<T>((size_t)m * d);
    s.sc = b.take<double>(SGP_SC_COUNT);
    s.nb = (int)((m + 7) / 8);
    s.partA = b.take<double>((size_t)s.nb * GPG_MAX_P);
    s.partB = b.take<double>((size_t)SGP_KGRAD_CHUNKS * s.nb * GPG_MAX_P);
    s.gxup = b.take<double>((size_t)(1 + SGP_KGRAD_CHUNKS) * m * d);
    s.st = b.take<FitState>(1);
    s.tc = sgp_uses_tc<T>(h, m, N);
    s.ldk = gpg_align_up((size_t)m, 64);
    if (s.tc) {
        s.Kfu = b.take<float>((size_t)N * s.ldk);
        s.Kfus = b.take<__half>(2 * (size_t)N * s.ldk);
        s.Bts = b.take<__half>(2 * (size_t)N * s.ldk);
        s.Uis = b.take<__half>(2 * (size_t)m * s.ldm);
        s.T2s = b.take<__half>(2 * (size_t)m * s.ldm);
        s.scales = b.take<float>(SGP_S_COUNT);
        s.Bs = b.take<__half>(2 * (size_t)m * s.ldn);
        s.nzt = (int)((N + SGP_TC_KCHUNK - 1) / SGP_TC_KCHUNK);
        s.SpartTc = b.take<float>((size_t)s.nzt * m * s.ldm);
        s.P1 = b.take<__half>(2 * (size_t)m * s.ldm);
        s.P2 = b.take<__half>(2 * (size_t)m * s.ldm);
        s.P3 = b.take<__half>(2 * (size_t)m * s.ldm);
        s.amax = b.take<float>(SGP_AMAX_BLOCKS);
        s.fplanes = b.take<unsigned char>(4 * (size_t)m * s.ldm * 4);
        s.fscales = b.take<float>(SC_COUNT);
    }
    return s;
}

// Cholesky (in place) + inverse of an m x m SPD matrix whose spectrum is bounded below by lam_min, on the tensor-core
// drivers of the exact path: the cooperative panel kernel (chol_panel.cuh) and trtri_tc.  For m of a few hundred the SIMT
// drivers are a chain of ~80 us steps per 128 columns; here a 512-column panel is one ~125 us launch.
static bool sgp_factor_uses_tc(const gpg_handle_s *h, int64_t m, int64_t ldm) {
    return m >= 384 && (ldm % 8) == 0 && h->opt_panel_mode == 3 && h->sm_count >= 4 && h->opt_gemm_path != 1;
}
static int sgp_factor_inverse_tc(gpg_handle_s *h, float *A, int64_t m, int64_t ldm, float *Ainv, float lam_min,
                                 int32_t *info, float *dinv, unsigned char *planes, float *scales, cudaStream_t s) {
    const size_t pl = (size_t)m * ldm * 4;
    TcPlanes Ls(planes, m, ldm), Ws(planes + pl, m, ldm), WTs(planes + 2 * pl, m, ldm), TTs(planes + 3 * pl, m, ldm);
    scales_from_diag_bound_kernel<<<1, 256, 0, s>>>(A, ldm, m, lam_min, scales);
    GPG_LAUNCH_CHECK(h);
    { StageTimer st(h, GPG_ST_CHOLESKY, s); GPG_TRY(cholesky_blocked_tc(h, A, m, ldm, info, 0, dinv, Ls, scales, s, TcPlanes(), Ws)); }
    { StageTimer st(h, GPG_ST_TRTRI, s); GPG_TRY(trtri_tc(h, A, m, ldm, Ainv, Ls, Ws, WTs, TTs, scales, s)); }
    return GPG_OK;
}

// C (m x N, fp32, leading dimension ldc) = A (m x m, fp16 planes As) * Bt^T with Bt (N x m, fp16 planes) on tcgen05;
// optionally emits the transposed split of the result (N x m planes Ts, scaled by *scale_out).
static int sgp_tc_product(gpg_handle_s *h, int64_t m, int64_t N, const __half *As, int64_t lda, const __half *Bts,
                          int64_t ldb, float *C, int64_t ldc, const float *scale_inv, int ke_mode, __half *Ts,
                          const float *scale_out, cudaStream_t s, __half *Ss = nullptr) {
    tc::Launch g;
    memset(&g.p, 0, sizeof(g.p));
    g.A.hi = As; g.A.lo = As + (size_t)m * lda; g.A.rows = m; g.A.cols = m; g.A.ld = lda;
    g.B.hi = Bts; g.B.lo = Bts + (size_t)N * ldb; g.B.rows = N; g.B.cols = m; g.B.ld = ldb;
    g.p.M = (int)m; g.p.N = (int)N; g.p.K = (int)m; g.p.batch = 1;
    g.p.ke_mode = ke_mode;
    g.p.epi = tc::EPI_STORE;
    g.p.scale_inv = scale_inv;
    g.p.C = C; g.p.ldc = ldc;
    g.p.alpha = 1.0f; g.p.beta = 0.0f;
    if (Ts) { g.p.T_hi = Ts; g.p.T_lo = Ts + (size_t)N * ldb; g.p.ldt = ldb; g.p.scale_out = scale_out; }
    if (Ss) { g.p.S_hi = Ss; g.p.S_lo = Ss + (size_t)m * ldc; g.p.lds = ldc; g.p.scale_out = scale_out; }   // same geometry as C
    return tc::launch(h, g, s);
}

// measured power-of-two operand scale of an fp32 matrix (see sgp_absmax_partial_kernel)
static int sgp_absmax_scale(gpg_handle_s *h, const float *M, int64_t ld, int64_t rows, int64_t cols, float *amax,
                            float *scales, int slot, int other, int slot_inv, cudaStream_t s) {
    sgp_absmax_partial_kernel<<<SGP_AMAX_BLOCKS, 256, 0, s>>>(M, ld, rows, cols, amax);
    GPG_LAUNCH_CHECK(h);
    sgp_absmax_finish_kernel<<<1, 32, 0, s>>>(amax, SGP_AMAX_BLOCKS, scales, slot, other, slot_inv);
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

// rho = beta * yin + alpha * A^T x for a rows x cols matrix, deterministic two-pass reduction (scratch from the handle)
template <typename T>
static int sgp_gemvT_rect(gpg_handle_s *h, const T *A, int64_t lda, int64_t rows, int64_t cols, const T *x, const T *yin,
                          T alpha, T beta, T *out, cudaStream_t s) {
    double *part;
    GPG_TRY(gpg_gemv_part_reserve(h, (size_t)SGP_GEMVT_CHUNKS * cols, &part));
    const dim3 grid((unsigned)((cols + 127) / 128), SGP_GEMVT_CHUNKS);
    gemvT_rect_partial_kernel<T><<<grid, 128, 0, s>>>(A, lda, rows, cols, x, part);
    GPG_LAUNCH_CHECK(h);
    gemvT_rect_finish_kernel<T><<<(unsigned)((cols + 255) / 256), 256, 0, s>>>(part, SGP_GEMVT_CHUNKS, cols, yin, alpha, beta, out);
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

// C (m x m, fp32) = alpha * A Bm^T on tcgen05, both operands m x m fp16 plane pairs with leading dimension ld
static int sgp_tc_mm(gpg_handle_s *h, int64_t m, int64_t ld, const __half *As, const __half *Bms, float *C,
                     const float *scale_inv, float alpha, int kb_mode, int tile_mode, cudaStream_t s) {
    tc::Launch g;
    memset(&g.p, 0, sizeof(g.p));
    g.A.hi = As; g.A.lo = As + (size_t)m * ld; g.A.rows = m; g.A.cols = m; g.A.ld = ld;
    g.B.hi = Bms; g.B.lo = Bms + (size_t)m * ld; g.B.rows = m; g.B.cols = m; g.B.ld = ld;
    g.p.M = (int)m; g.p.N = (int)m; g.p.K = (int)m; g.p.batch = 1;
    g.p.kb_mode = kb_mode; g.p.tile_mode = tile_mode;
    g.p.epi = tc::EPI_STORE;
    g.p.scale_inv = scale_inv;
    g.p.C = C; g.p.ldc = ld;
    g.p.alpha = alpha; g.p.beta = 0.0f;
    return tc::launch(h, g, s);
}

// fp16 planes of M^T (m x m, leading dimension ld) scaled by *scale
static int sgp_split_T(gpg_handle_s *h, const float *M, int64_t m, int64_t ld, const float *scale, __half *planes,
                       cudaStream_t s) {
    const dim3 grid((unsigned)((m + 31) / 32), (unsigned)((m + 31) / 32));
    tc::split_transpose_kernel<<<grid, 256, 0, s>>>(M, ld, m, m, scale, planes, planes + (size_t)m * ld, ld);
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

// Luu, Ui, B, S, A', LA, LAi, beta, c0, a0, a, w for the theta stored on the device.  info keeps the first failing pivot of
// either factorisation (the caller resets it).
template <typename T>
static int sgp_lowrank_core(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *X, const T *y, int64_t N,
                            const T *Xu, int64_t m, double jitter, const SgpBufs<T> &b, int32_t *info, cudaStream_t s) {
    const int64_t ldm = b.ldm, ldn = b.ldn;
    sgp_theta0_kernel<T><<<1, 32, 0, s>>>(theta, 3 + d, b.theta0);
    GPG_LAUNCH_CHECK(h);
    { StageTimer st(h, GPG_ST_KMAT, s);
      GPG_TRY(kmat_launch<T>(h, kernel_id, d, b.theta0, Xu, m, nullptr, m, jitter, 0, b.Luu, ldm, s));
      if constexpr (std::is_same<T, float>::value) {
          if (b.tc) {                // k(X, Xu) (N x m) straight into the fp16 planes the tcgen05 product reads
              sgp_scales_theta_kernel<T><<<1, 32, 0, s>>>(theta, b.scales);
              GPG_LAUNCH_CHECK(h);
              KmatSplit sp;
              sp.hi = b.Kfus; sp.lo = b.Kfus + (size_t)N * b.ldk; sp.ld = b.ldk; sp.scale = b.scales + SGP_S_K;
              GPG_TRY(kmat_launch<T>(h, kernel_id, d, theta, X, N, Xu, m, 0.0, 0, b.Kfu, b.ldk, s, sp));
          }
      }
      if (!b.tc) GPG_TRY(kmat_launch<T>(h, kernel_id, d, theta, Xu, m, X, N, 0.0, 0, b.Kuf, ldn, s)); }
    bool fdone = false;
    if constexpr (std::is_same<T, float>::value) {
        if (b.tc && sgp_factor_uses_tc(h, m, ldm)) {     // Kuu + jitter I >= jitter
            GPG_TRY(sgp_factor_inverse_tc(h, b.Luu, m, ldm, b.Ui, (float)jitter, info, b.dinv, b.fplanes, b.fscales, s));
            fdone = true;
        }
    }
    if (!fdone) {
        { StageTimer st(h, GPG_ST_CHOLESKY, s); GPG_TRY(cholesky_blocked<T>(h, b.Luu, m, ldm, info, 0, b.dinv, s)); }
        { StageTimer st(h, GPG_ST_TRTRI, s); GPG_TRY(trtri_blocked<T>(h, b.Luu, m, ldm, b.Ui, ldm, b.tmp, s)); }
    }
    {
        StageTimer stg(h, GPG_ST_PGEMM, s);      // B = Ui Kuf (booked under the predict GEMM's stage; S = B B^T under PFINAL)
        if (ldn > N)                 // the split-K batches of S = B B^T read B up to column ldn
            GPG_CUDA_CHECK(cudaMemset2DAsync(b.B + N, ldn * sizeof(T), 0, (ldn - N) * sizeof(T), m, s));
        bool done = false;
        if constexpr (std::is_same<T, float>::value) {
            if (b.tc) {              // B = Ui k(X, Xu)^T on tcgen05; the epilogue also leaves B^T as fp16 planes for dF/dKuf
                GPG_TRY(sgp_absmax_scale(h, b.Ui, ldm, m, m, b.amax, b.scales, SGP_S_U, SGP_S_K, SGP_S_UK_INV, s));
                GPG_TRY(tc::split_matrix(h, b.Ui, ldm, m, m, b.scales + SGP_S_U, b.Uis, b.Uis + (size_t)m * ldm, ldm, 1, s));
                GPG_TRY(sgp_tc_product(h, m, N, b.Uis, ldm, b.Kfus, b.ldk, b.B, ldn, b.scales + SGP_S_UK_INV, GEMM_KE_M,
                                       b.Bts, b.scales + SGP_S_B, s, b.Bs));
                done = true;
            }
        }
        if (!done) {
            GemmArgs<T> g;           // B = Ui Kuf  (Ui lower triangular: k <= i)
            g.A = b.Ui; g.lda = ldm; g.a_kmajor = 1;
            g.B = b.Kuf; g.ldb = ldn; g.b_kmajor = 0;
            g.C = b.B; g.ldc = ldn;
            g.M = (int)m; g.N = (int)N; g.K = (int)m;
            g.ke_mode = GEMM_KE_M;
            GPG_TRY(gemm_simt<T>(h, g, s));
        }
    }
    {
        StageTimer stg(h, GPG_ST_PFINAL, s);
        const dim3 gmm((unsigned)((m + 255) / 256), (unsigned)m);
        bool done = false;
        if constexpr (std::is_same<T, float>::value) {
            if (b.tc) {              // S = B B^T on tcgen05 from the fp16 planes of B, lower tiles, split-K batches
                tc::Launch g;
                memset(&g.p, 0, sizeof(g.p));
                g.A.hi = b.Bs; g.A.lo = b.Bs + (size_t)m * ldn; g.A.rows = m; g.A.cols = N; g.A.ld = ldn;
                g.B = g.A;
                g.p.M = (int)m; g.p.N = (int)m; g.p.K = SGP_TC_KCHUNK; g.p.batch = b.nzt;
                g.p.a_kbs = SGP_TC_KCHUNK; g.p.b_kbs = SGP_TC_KCHUNK;
                g.p.tile_mode = GEMM_TILES_LOWER;
                g.p.epi = tc::EPI_STORE;
                g.p.scale_inv = b.scales + SGP_S_BB_INV;
                g.p.C = b.SpartTc; g.p.ldc = ldm; g.p.c_bs = m * ldm;
                g.p.alpha = 1.0f; g.p.beta = 0.0f;
                GPG_TRY(tc::launch(h, g, s));
                sgp_form_A_kernel<T><<<gmm, 256, 0, s>>>(b.SpartTc, b.nzt, m * ldm, ldm, m, theta, b.S, b.Ap, b.LA);
                GPG_LAUNCH_CHECK(h);
                done = true;
            }
        }
        if (!done) {
            GemmArgs<T> g;           // S = B B^T, lower tiles, split along K into nz batches (see sgp_split)
            g.A = b.B; g.lda = ldn; g.a_kmajor = 1;
            g.B = b.B; g.ldb = ldn; g.b_kmajor = 1;
            g.C = b.Spart; g.ldc = ldm;
            g.M = (int)m; g.N = (int)m; g.K = (int)b.kchunk;
            g.batch = b.nz; g.strideA = b.kchunk; g.strideB = b.kchunk; g.strideC = m * ldm;
            g.tile_mode = GEMM_TILES_LOWER;
            GPG_TRY(gemm_simt<T>(h, g, s));
            sgp_form_A_kernel<T><<<gmm, 256, 0, s>>>(b.Spart, b.nz, m * ldm, ldm, m, theta, b.S, b.Ap, b.LA);
            GPG_LAUNCH_CHECK(h);
        }
    }
    fdone = false;
    if constexpr (std::is_same<T, float>::value) {
        if (b.tc && sgp_factor_uses_tc(h, m, ldm)) {     // A' = I + B B^T / s2 >= I
            GPG_TRY(sgp_factor_inverse_tc(h, b.LA, m, ldm, b.LAi, 1.0f, info, b.dinv, b.fplanes, b.fscales, s));
            fdone = true;
        }
    }
    if (!fdone) {
        { StageTimer st(h, GPG_ST_CHOLESKY, s); GPG_TRY(cholesky_blocked<T>(h, b.LA, m, ldm, info, 0, b.dinv, s)); }
        { StageTimer st(h, GPG_ST_TRTRI, s); GPG_TRY(trtri_blocked<T>(h, b.LA, m, ldm, b.LAi, ldm, b.tmp, s)); }
    }
    StageTimer st(h, GPG_ST_SOLVE, s);
    const unsigned gm8 = (unsigned)((m + 7) / 8), gm256 = (unsigned)((m + 255) / 256);
    gemv_rect_kernel<T><<<gm8, 256, 0, s>>>(b.B, ldn, m, N, y, b.beta);                              // beta = B y
    GPG_LAUNCH_CHECK(h);
    gemv_tri_kernel<T, false><<<gm8, 256, 0, s>>>(b.LAi, ldm, m, b.beta, nullptr, T(1), T(0), b.c0);  // c0 = LA^-1 beta
    GPG_LAUNCH_CHECK(h);
    GPG_TRY(gemv_tri_T<T>(h, b.LAi, ldm, m, b.c0, nullptr, T(1), T(0), b.a0, s));                     // a0 = A'^-1 beta
    sgp_div_noise_kernel<T><<<gm256, 256, 0, s>>>(b.a0, m, theta, b.a);                              // a = a0 / s2
    GPG_LAUNCH_CHECK(h);
    GPG_TRY(gemv_tri_T<T>(h, b.Ui, ldm, m, b.a, nullptr, T(1), T(0), b.w, s));                        // w = Ui^T a
    return GPG_OK;
}

// loss and gradient w.r.t. the constrained theta (b.grad layout of theta) and the inducing inputs (gxu, m x d)
template <typename T>
static int sgp_loss_grad_core(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *X, const T *y, int64_t N,
                              const T *Xu, int64_t m, double jitter, const SgpBufs<T> &b, T *loss_out, T *grad_out,
                              T *gxu_out, int32_t *info, cudaStream_t s) {
    GPG_TRY(sgp_lowrank_core<T>(h, kernel_id, d, theta, X, y, N, Xu, m, jitter, b, info, s));
    StageTimer st(h, GPG_ST_GRAD, s);
    const int64_t ldm = b.ldm, ldn = b.ldn;
    auto mm_gemm = [&](const T *A, int a_km, const T *Bm, int b_km, T *C, T alpha, int kb_mode, int tile_mode) -> int {
        GemmArgs<T> g;
        g.A = A; g.lda = ldm; g.a_kmajor = a_km;
        g.B = Bm; g.ldb = ldm; g.b_kmajor = b_km;
        g.C = C; g.ldc = ldm;
        g.M = (int)m; g.N = (int)m; g.K = (int)m;
        g.alpha = alpha; g.kb_mode = kb_mode; g.tile_mode = tile_mode;
        // m ~ 10^3: with 128 x 128 tiles an fp32 product has (m / 128)^2 ~ 50 CTAs for 148 SMs; the 32-row panel
        // tiles give four times as many (fp64 already runs 64 x 64 tiles)
        if constexpr (std::is_same<T, float>::value) return gemm_simt<T, GemmCfgPanelF32>(h, g, s);
        else return gemm_simt<T>(h, g, s);
    };
    bool tcmm = false;               // fp32, large m: the m x m x m products on tcgen05 as well
    if constexpr (std::is_same<T, float>::value) tcmm = b.tc;
    if constexpr (std::is_same<T, float>::value) {
        if (tcmm) {                  // A'^-1 = LAi^T LAi (lower) = (LAi^T) (LAi^T)^T: operand = transposed planes of LAi
            GPG_TRY(sgp_split_T(h, b.LAi, m, ldm, b.scales + SGP_S_LA, b.P1, s));
            GPG_TRY(sgp_tc_mm(h, m, ldm, b.P1, b.P1, b.Ainv, b.scales + SGP_S_LALA_INV, 1.0f, GEMM_KB_MAXMN,
                              GEMM_TILES_LOWER, s));
        }
    }
    if (!tcmm) GPG_TRY(mm_gemm(b.LAi, 0, b.LAi, 0, b.Ainv, T(1), GEMM_KB_MAXMN, GEMM_TILES_LOWER));     // A'^-1 = LAi^T LAi (lower)
    sgp_scalars_kernel<T><<<1, 1024, 0, s>>>(y, N, b.beta, b.a0, b.c0, b.S, b.Ainv, b.LA, ldm, m, b.sc);
    GPG_LAUNCH_CHECK(h);
    GPG_TRY(sgp_gemvT_rect<T>(h, b.B, ldn, m, N, b.a, y, T(-1), T(1), b.rho, s));                  // rho = y - B^T a
    const dim3 gmm((unsigned)((m + 255) / 256), (unsigned)m);
    sgp_form_phi_kernel<T><<<gmm, 256, 0, s>>>(b.Ainv, b.Ap, b.a, ldm, m, b.Phi, b.H);
    GPG_LAUNCH_CHECK(h);
    if constexpr (std::is_same<T, float>::value) {
        if (tcmm) {
            // with UiT = planes of Ui^T (P2):  T1^T = Ui^T Phi = UiT Phi^T (Phi symmetric),  T2 = Ui^T H = UiT H^T,
            // dF/dKuu = -1/2 (Ui^T Phi) Ui = -1/2 T1^T UiT^T;  P3 carries Phi, H and T1^T in turn
            GPG_TRY(sgp_split_T(h, b.Ui, m, ldm, b.scales + SGP_S_U, b.P2, s));
            GPG_TRY(sgp_absmax_scale(h, b.Phi, ldm, m, m, b.amax, b.scales, SGP_S_PHI, SGP_S_U, SGP_S_UPHI_INV, s));
            GPG_TRY(tc::split_matrix(h, b.Phi, ldm, m, m, b.scales + SGP_S_PHI, b.P3, b.P3 + (size_t)m * ldm, ldm, 0, s));
            GPG_TRY(sgp_tc_mm(h, m, ldm, b.P2, b.P3, b.T1, b.scales + SGP_S_UPHI_INV, 1.0f, GEMM_KB_NONE, GEMM_TILES_ALL, s));
            GPG_TRY(sgp_absmax_scale(h, b.H, ldm, m, m, b.amax, b.scales, SGP_S_H, SGP_S_U, SGP_S_UH_INV, s));
            GPG_TRY(tc::split_matrix(h, b.H, ldm, m, m, b.scales + SGP_S_H, b.P3, b.P3 + (size_t)m * ldm, ldm, 0, s));
            GPG_TRY(sgp_tc_mm(h, m, ldm, b.P2, b.P3, b.T2, b.scales + SGP_S_UH_INV, 1.0f, GEMM_KB_NONE, GEMM_TILES_ALL, s));
            GPG_TRY(sgp_absmax_scale(h, b.T1, ldm, m, m, b.amax, b.scales, SGP_S_XT, SGP_S_U, SGP_S_XTU_INV, s));
            GPG_TRY(tc::split_matrix(h, b.T1, ldm, m, m, b.scales + SGP_S_XT, b.P3, b.P3 + (size_t)m * ldm, ldm, 0, s));
            GPG_TRY(sgp_tc_mm(h, m, ldm, b.P3, b.P2, b.Guu, b.scales + SGP_S_XTU_INV, -0.5f, GEMM_KB_N0, GEMM_TILES_ALL, s));
        }
    }
    if (!tcmm) {
        GPG_TRY(mm_gemm(b.Phi, 1, b.Ui, 0, b.T1, T(1), GEMM_KB_NONE, GEMM_TILES_ALL));           // T1 = Phi Ui
        GPG_TRY(mm_gemm(b.Ui, 0, b.T1, 0, b.Guu, T(-0.5), GEMM_KB_NONE, GEMM_TILES_ALL));        // dF/dKuu = -1/2 Ui^T T1
        GPG_TRY(mm_gemm(b.Ui, 0, b.H, 0, b.T2, T(1), GEMM_KB_NONE, GEMM_TILES_ALL));             // T2 = Ui^T H
    }
    bool guf_done = false;           // s2 dF/dKuf + w rho^T = T2 B   (into the Kuf buffer, which is dead by now)
    if constexpr (std::is_same<T, float>::value) {
        if (b.tc) {
            GPG_TRY(sgp_absmax_scale(h, b.T2, ldm, m, m, b.amax, b.scales, SGP_S_T, SGP_S_B, SGP_S_TB_INV, s));
            GPG_TRY(tc::split_matrix(h, b.T2, ldm, m, m, b.scales + SGP_S_T, b.T2s, b.T2s + (size_t)m * ldm, ldm, 0, s));
            GPG_TRY(sgp_tc_product(h, m, N, b.T2s, ldm, b.Bts, b.ldk, b.Kuf, ldn, b.scales + SGP_S_TB_INV, GEMM_KE_NONE,
                                   nullptr, nullptr, s));
            guf_done = true;
        }
    }
    if (!guf_done) {
        GemmArgs<T> g;
        g.A = b.T2; g.lda = ldm; g.a_kmajor = 1;
        g.B = b.B; g.ldb = ldn; g.b_kmajor = 0;
        g.C = b.Kuf; g.ldc = ldn;
        g.M = (int)m; g.N = (int)N; g.K = (int)m;
        GPG_TRY(gemm_simt<T>(h, g, s));
    }
    T *gxu = gxu_out ? gxu_out : b.gxu;
    const int kchunks = (int)std::max<int64_t>(1, std::min<int64_t>(SGP_KGRAD_CHUNKS, (N + 1023) / 1024));
    GPG_DISPATCH_KID(kernel_id, GPG_DISPATCH_D(d, {
        // row sums for the inducing inputs go through double partials: [0] the m x m term (factor 2), [1..] the chunks of
        // the m x N term; one small kernel adds them up (deterministic, no atomics)
        sgp_kgrad_kernel<T, KID, D><<<b.nb, 256, 0, s>>>(theta, Xu, m, Xu, m, b.Guu, ldm, nullptr, nullptr, 1.0, 0, 2.0, 0,
                                                        b.partA, nullptr, b.gxup);
        sgp_kgrad_kernel<T, KID, D><<<dim3(b.nb, kchunks), 256, 0, s>>>(theta, Xu, m, X, N, b.Kuf, ldn, b.w, b.rho, 1.0, 1, 1.0,
                                                                       0, b.partB, nullptr, b.gxup + (size_t)m * D);
    }));
    GPG_LAUNCH_CHECK(h);
    h->launches++;
    if (sizeof(T) == 4) sgp_gxu_reduce_kernel<float><<<(unsigned)((m * d + 255) / 256), 256, 0, s>>>(b.gxup, 1 + kchunks, m * d, (float *)gxu);
    else sgp_gxu_reduce_kernel<double><<<(unsigned)((m * d + 255) / 256), 256, 0, s>>>(b.gxup, 1 + kchunks, m * d, (double *)gxu);
    GPG_LAUNCH_CHECK(h);
    sgp_finish_kernel<T><<<1, 256, 0, s>>>(b.partA, b.nb, b.partB, b.nb * kchunks, 3 + d, b.sc, theta, N, m, grad_out, loss_out);
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

#define SGP_COMMON_REQUIRE()                                                                         \
    GPG_REQUIRE(N > 0 && m > 0 && m <= N, "need 0 < m <= N");                                        \
    GPG_REQUIRE(m < 65536, "at most 65535 inducing points");                                         \
    GPG_REQUIRE(d >= 1 && d <= GPG_MAX_D, "d not in 1..4");                                          \
    GPG_REQUIRE(kernel_id >= 0 && kernel_id <= 2, "unknown kernel id")

template <typename T>
static int sgp_loss_grad_entry(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *X, const T *y, int64_t N,
                               const T *Xu, int64_t m, double jitter, T *loss_out, T *grad_out, T *gxu_out, int32_t *info,
                               cudaStream_t s) {
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, sgp_ws_bytes<T>(h, m, N, d), &ws));
    SgpBufs<T> b = sgp_carve<T>(h, ws, m, N, d);
    GPG_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), s));
    return sgp_loss_grad_core<T>(h, kernel_id, d, theta, X, y, N, Xu, m, jitter, b, loss_out, grad_out, gxu_out, info, s);
}

extern "C" int gpg_sparse_loss_grad(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta, const void *X,
                                    const void *y, int64_t N, const void *Xu, int64_t m, double jitter, void *loss_out,
                                    void *grad_theta_out, void *grad_xu_out, int32_t *info, void *stream) {
    GPG_REQUIRE(h && theta && X && y && Xu && loss_out && grad_theta_out && grad_xu_out && info, "NULL argument");
    DeviceGuard device_guard(h->device);
    SGP_COMMON_REQUIRE();
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return sgp_loss_grad_entry<float>(h, kernel_id, d, (const float *)theta, (const float *)X, (const float *)y, N,
                                          (const float *)Xu, m, jitter, (float *)loss_out, (float *)grad_theta_out,
                                          (float *)grad_xu_out, info, s);
    if (dtype == GPG_F64)
        return sgp_loss_grad_entry<double>(h, kernel_id, d, (const double *)theta, (const double *)X, (const double *)y, N,
                                           (const double *)Xu, m, jitter, (double *)loss_out, (double *)grad_theta_out,
                                           (double *)grad_xu_out, info, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

template <typename T>
static int sgp_fit_entry(gpg_handle_s *h, int kernel_id, int d, int n_ls, const T *X, const T *y, int64_t N, T *Xu,
                         int64_t m, double jitter, T *u, const double *bounds, int iters, double lr, T *traj, T *xu_traj,
                         T *theta_out, int32_t *info, cudaStream_t s) {
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, sgp_ws_bytes<T>(h, m, N, d), &ws));
    SgpBufs<T> b = sgp_carve<T>(h, ws, m, N, d);
    FitCfg c;
    memset(&c, 0, sizeof(c));
    c.d = d; c.n_ls = n_ls; c.is_rq = (kernel_id == GPG_RATQUAD);
    c.var_lo = bounds[0]; c.var_hi = bounds[1];
    for (int k = 0; k < n_ls; ++k) { c.ls_lo[k] = bounds[2 + k]; c.ls_hi[k] = bounds[2 + n_ls + k]; }
    c.lr = lr; c.beta1 = 0.9; c.beta2 = 0.999; c.eps = 1e-8;
    const int64_t nxu = m * d;
    adam_step_kernel<T><<<1, 32, 0, s>>>(0, c, u, b.st, nullptr, nullptr, b.theta, nullptr);      // fresh Adam state
    GPG_LAUNCH_CHECK(h);
    GPG_CUDA_CHECK(cudaMemsetAsync(b.m1, 0, nxu * sizeof(T), s));
    GPG_CUDA_CHECK(cudaMemsetAsync(b.m2, 0, nxu * sizeof(T), s));
    GPG_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), s));
    for (int it = 0; it < iters; ++it) {
        GPG_TRY(sgp_loss_grad_core<T>(h, kernel_id, d, b.theta, X, y, N, Xu, m, jitter, b, b.loss, b.grad, nullptr, info, s));
        adam_step_kernel<T><<<1, 32, 0, s>>>(1, c, u, b.st, (const T *)b.grad, (const T *)b.loss, b.theta, traj);
        GPG_LAUNCH_CHECK(h);
        sgp_adam_xu_kernel<T><<<(unsigned)((nxu + 255) / 256), 256, 0, s>>>(c, b.st, nxu, Xu, b.gxu, b.m1, b.m2, xu_traj);
        GPG_LAUNCH_CHECK(h);
    }
    if (theta_out) GPG_CUDA_CHECK(cudaMemcpyAsync(theta_out, b.theta, (3 + d) * sizeof(T), cudaMemcpyDeviceToDevice, s));
    return GPG_OK;
}

extern "C" int gpg_sparse_fit_adam(gpg_handle_t h, int dtype, int kernel_id, int d, int n_ls, const void *X,
                                   const void *y, int64_t N, void *Xu, int64_t m, double jitter, void *u,
                                   const double *bounds_host, int iters, double lr, void *traj_out, void *xu_traj_out,
                                   void *theta_out, int32_t *info, void *stream) {
    GPG_REQUIRE(h && X && y && Xu && u && bounds_host && info, "NULL argument");
    DeviceGuard device_guard(h->device);
    SGP_COMMON_REQUIRE();
    GPG_REQUIRE(iters >= 0, "iters must not be negative");
    GPG_REQUIRE(iters == 0 || traj_out != nullptr, "traj_out is NULL");
    GPG_REQUIRE(n_ls == 1 || n_ls == d, "n_ls must be 1 or d");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return sgp_fit_entry<float>(h, kernel_id, d, n_ls, (const float *)X, (const float *)y, N, (float *)Xu, m, jitter,
                                    (float *)u, bounds_host, iters, lr, (float *)traj_out, (float *)xu_traj_out,
                                    (float *)theta_out, info, s);
    if (dtype == GPG_F64)
        return sgp_fit_entry<double>(h, kernel_id, d, n_ls, (const double *)X, (const double *)y, N, (double *)Xu, m,
                                     jitter, (double *)u, bounds_host, iters, lr, (double *)traj_out,
                                     (double *)xu_traj_out, (double *)theta_out, info, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

// Factor cache of the inducing-point posterior: Ui = Luu^-1, Pm = LA^-1 Luu^-1 (both m x m lower, leading dimension ld),
// w = Luu^-T A'^-1 B y / s2, so that mean = k(x*, Xu) w and var = v + noise - |Ui k*|^2 + |Pm k*|^2.
template <typename T>
static int sgp_factorize_entry(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *X, const T *y, int64_t N,
                               const T *Xu, int64_t m, double jitter, T *Ui_out, T *P_out, int64_t ld, T *w_out,
                               int32_t *info, void *split_out, float *scales_out, cudaStream_t s) {
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, sgp_ws_bytes<T>(h, m, N, d), &ws));
    SgpBufs<T> b = sgp_carve<T>(h, ws, m, N, d);
    GPG_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), s));
    GPG_TRY(sgp_lowrank_core<T>(h, kernel_id, d, theta, X, y, N, Xu, m, jitter, b, info, s));
    GemmArgs<T> g;                   // Pm = LAi Ui (lower x lower)
    g.A = b.LAi; g.lda = b.ldm; g.a_kmajor = 1;
    g.B = b.Ui; g.ldb = b.ldm; g.b_kmajor = 0;
    g.C = P_out; g.ldc = ld;
    g.M = (int)m; g.N = (int)m; g.K = (int)m;
    g.ke_mode = GEMM_KE_M;
    GPG_TRY(gemm_simt<T>(h, g, s));
    GPG_CUDA_CHECK(cudaMemcpy2DAsync(Ui_out, ld * sizeof(T), b.Ui, b.ldm * sizeof(T), m * sizeof(T), m,
                                     cudaMemcpyDeviceToDevice, s));
    GPG_CUDA_CHECK(cudaMemcpyAsync(w_out, b.w, m * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if constexpr (std::is_same<T, float>::value) {
        if (split_out) {             // tensor-core form of the two factors: fp16 planes {Ui hi, Ui lo, Pm hi, Pm lo} + scales
            __half *pl = (__half *)split_out;
            const size_t plane = (size_t)m * ld;
            sgp_scales_theta_kernel<T><<<1, 32, 0, s>>>(theta, scales_out);
            GPG_LAUNCH_CHECK(h);
            float *amax = b.Spart;   // dead by now (>= 64 x m floats): scratch for the per-block maxima
            GPG_TRY(sgp_absmax_scale(h, Ui_out, ld, m, m, amax, scales_out, SGP_S_U, SGP_S_K, SGP_S_UK_INV, s));
            GPG_TRY(sgp_absmax_scale(h, P_out, ld, m, m, amax, scales_out, SGP_S_P, SGP_S_K, SGP_S_KP_INV, s));
            GPG_TRY(tc::split_matrix(h, Ui_out, ld, m, m, scales_out + SGP_S_U, pl, pl + plane, ld, 1, s));
            GPG_TRY(tc::split_matrix(h, P_out, ld, m, m, scales_out + SGP_S_P, pl + 2 * plane, pl + 3 * plane, ld, 1, s));
        }
    }
    return GPG_OK;
}

extern "C" int gpg_sparse_factorize(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta, const void *X,
                                    const void *y, int64_t N, const void *Xu, int64_t m, double jitter, void *Ui_out,
                                    void *P_out, int64_t ld, void *w_out, int32_t *info, void *split_out,
                                    float *scales_out, void *stream) {
    GPG_REQUIRE(h && theta && X && y && Xu && Ui_out && P_out && w_out && info, "NULL argument");
    DeviceGuard device_guard(h->device);
    SGP_COMMON_REQUIRE();
    GPG_REQUIRE(ld >= m, "ld smaller than m");
    GPG_REQUIRE((split_out == nullptr) == (scales_out == nullptr), "split_out and scales_out go together");
    GPG_REQUIRE(split_out == nullptr || (dtype == GPG_F32 && ld % 8 == 0), "the split factors need f32 and ld % 8 == 0");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return sgp_factorize_entry<float>(h, kernel_id, d, (const float *)theta, (const float *)X, (const float *)y, N,
                                          (const float *)Xu, m, jitter, (float *)Ui_out, (float *)P_out, ld,
                                          (float *)w_out, info, split_out, scales_out, s);
    if (dtype == GPG_F64)
        return sgp_factorize_entry<double>(h, kernel_id, d, (const double *)theta, (const double *)X, (const double *)y, N,
                                           (const double *)Xu, m, jitter, (double *)Ui_out, (double *)P_out, ld,
                                           (double *)w_out, info, nullptr, nullptr, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

template <typename T, int D>
static int sgp_predict_core(gpg_handle_s *h, int kernel_id, const T *theta, const T *Xu, int64_t m, const T *Ui,
                            const T *Pm, int64_t ld, const T *w, const T *Xs, int64_t M, T *mean, T *sd, cudaStream_t s) {
    using C = GemmCfg<T>;
    const int64_t ldk = gpg_align_up((size_t)m, 64);
    int64_t chunk = h->opt_predict_chunk > 0 ? h->opt_predict_chunk : 16384;
    while (chunk > 512 && chunk * ldk * (int64_t)sizeof(T) > (int64_t)768 << 20) chunk /= 2;
    chunk = std::min<int64_t>(chunk, gpg_align_up((size_t)M, 128));
    const int tiles_m = (int)((m + C::BM - 1) / C::BM);
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, bump_size({(size_t)chunk * ldk * sizeof(T), (size_t)tiles_m * chunk * sizeof(T),
                                         (size_t)tiles_m * chunk * sizeof(T)}), &ws));
    Bump b(ws);
    T *Ks = b.take<T>((size_t)chunk * ldk);
    T *part1 = b.take<T>((size_t)tiles_m * chunk);
    T *part2 = b.take<T>((size_t)tiles_m * chunk);
    TestPoints<T, D> tp;
    tp.j0 = 0;
    for (int k = 0; k < GPG_MAX_D; ++k) { tp.dims[k] = 1; tp.step[k] = T(1); }
    for (int64_t c0 = 0; c0 < M; c0 += chunk) {
        const int64_t mc = std::min<int64_t>(chunk, M - c0);
        tp.Xs = Xs + c0 * D;
        {
            StageTimer st(h, GPG_ST_KCROSS, s);
            GPG_DISPATCH_KID(kernel_id, kcross_mean_kernel<T, KID, D, false><<<(unsigned)((mc + 7) / 8), 256, 0, s>>>(
                                            theta, Xu, m, tp, mc, w, Ks, ldk, nullptr, nullptr, 0, nullptr, mean + c0));
            GPG_LAUNCH_CHECK(h);
        }
        {
            StageTimer st(h, GPG_ST_PGEMM, s);
            for (int pass = 0; pass < 2; ++pass) {      // colsum((Ui Ks^T)^2), colsum((Pm Ks^T)^2)
                GemmArgs<T> g;
                g.A = pass == 0 ? Ui : Pm; g.lda = ld; g.a_kmajor = 1;
                g.B = Ks; g.ldb = ldk; g.b_kmajor = 1;
                g.M = (int)m; g.N = (int)mc; g.K = (int)m;
                g.ke_mode = GEMM_KE_M;
                g.epi = GEMM_EPI_COLSUMSQ;
                g.part = pass == 0 ? part1 : part2; g.ldpart = chunk;
                GPG_TRY(gemm_simt<T>(h, g, s));
            }
        }
        StageTimer st(h, GPG_ST_PFINAL, s);
        sgp_predict_finalize_kernel<T, D><<<(unsigned)((mc + 255) / 256), 256, 0, s>>>(theta, part1, part2, tiles_m, chunk,
                                                                                       tp, mc, sd + c0);
        GPG_LAUNCH_CHECK(h);
    }
    return GPG_OK;
}

// fp32 tensor-core route: K* rows are generated directly as fp16 planes, both column-sum-of-squares reductions run in
// the epilogue of the tcgen05 GEMM (the machinery of predict_core_tc on the two m x m factors).
template <int D>
static int sgp_predict_core_tc(gpg_handle_s *h, int kernel_id, const float *theta, const float *Xu, int64_t m,
                               const __half *planes, int64_t ld, const float *scales, const float *w, const float *Xs,
                               int64_t M, float *mean, float *sd, cudaStream_t s) {
    const int64_t ldh = gpg_align_up((size_t)m, 64);
    int64_t chunk = h->opt_predict_chunk > 0 ? h->opt_predict_chunk : 16384;
    chunk = gpg_align_up((size_t)std::min<int64_t>(chunk, gpg_align_up((size_t)M, 128)), 128);
    const int tiles_n = (int)((m + tc::BN - 1) / tc::BN);
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, bump_size({(size_t)chunk * ldh * 2, (size_t)chunk * ldh * 2,
                                         (size_t)tiles_n * chunk * sizeof(float), (size_t)tiles_n * chunk * sizeof(float)}), &ws));
    Bump b(ws);
    __half *Khi = b.take<__half>((size_t)chunk * ldh);
    __half *Klo = b.take<__half>((size_t)chunk * ldh);
    float *part1 = b.take<float>((size_t)tiles_n * chunk);
    float *part2 = b.take<float>((size_t)tiles_n * chunk);
    const size_t plane = (size_t)m * ld;
    const int m_group = (int)std::max<int64_t>(1, ((int64_t)64 << 20) / (tc::BM * ldh * 4));
    TestPoints<float, D> tp;
    tp.j0 = 0;
    for (int k = 0; k < GPG_MAX_D; ++k) { tp.dims[k] = 1; tp.step[k] = 1.0f; }
    for (int64_t c0 = 0; c0 < M; c0 += chunk) {
        const int64_t mc = std::min<int64_t>(chunk, M - c0);
        tp.Xs = Xs + c0 * D;
        {
            StageTimer st(h, GPG_ST_KCROSS, s);
            GPG_DISPATCH_KID(kernel_id, kcross_mean_kernel<float, KID, D, true><<<(unsigned)((mc + 7) / 8), 256, 0, s>>>(
                                            theta, Xu, m, tp, mc, w, nullptr, 0, Khi, Klo, ldh, scales + SGP_S_K, mean + c0));
            GPG_LAUNCH_CHECK(h);
        }
        {
            StageTimer st(h, GPG_ST_PGEMM, s);
            for (int pass = 0; pass < 2; ++pass) {      // rowsum((K* Ui^T)^2), rowsum((K* Pm^T)^2)
                tc::Launch g;
                memset(&g.p, 0, sizeof(g.p));
                g.A.hi = Khi; g.A.lo = Klo; g.A.rows = mc; g.A.cols = m; g.A.ld = ldh;
                g.B.hi = planes + 2 * pass * plane; g.B.lo = planes + (2 * pass + 1) * plane;
                g.B.rows = m; g.B.cols = m; g.B.ld = ld;
                g.p.M = (int)mc; g.p.N = (int)m; g.p.K = (int)m; g.p.batch = 1;
                g.p.m_group = m_group;
                g.p.ke_mode = GEMM_KE_N;
                g.p.epi = tc::EPI_ROWSUMSQ;
                g.p.scale_inv = scales + (pass == 0 ? SGP_S_UK_INV : SGP_S_KP_INV);
                g.p.part = pass == 0 ? part1 : part2; g.p.ldpart = chunk;
                GPG_TRY(tc::launch(h, g, s));
            }
        }
        StageTimer st(h, GPG_ST_PFINAL, s);
        sgp_predict_finalize_kernel<float, D><<<(unsigned)((mc + 255) / 256), 256, 0, s>>>(theta, part1, part2, tiles_n, chunk,
                                                                                           tp, mc, sd + c0);
        GPG_LAUNCH_CHECK(h);
    }
    return GPG_OK;
}

template <typename T>
static int sgp_predict_entry(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *Xu, int64_t m, const T *Ui,
                             const T *Pm, int64_t ld, const T *w, const void *split, const float *scales, const T *Xs,
                             int64_t M, T *mean, T *sd, cudaStream_t s) {
    if (M == 0) return GPG_OK;
    if constexpr (std::is_same<T, float>::value) {
        if (split != nullptr && h->opt_gemm_path != 1 && (h->opt_gemm_path == 2 || m >= 512)) {
            GPG_DISPATCH_D(d, { return sgp_predict_core_tc<D>(h, kernel_id, theta, Xu, m, (const __half *)split, ld, scales, w,
                                                              Xs, M, mean, sd, s); });
        }
    }
    GPG_DISPATCH_D(d, { return sgp_predict_core<T, D>(h, kernel_id, theta, Xu, m, Ui, Pm, ld, w, Xs, M, mean, sd, s); });
    return GPG_OK;
}

extern "C" int gpg_sparse_predict(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta, const void *Xu,
                                  int64_t m, const void *Ui, const void *Pm, int64_t ld, const void *w, const void *split,
                                  const float *scales, const void *Xs, int64_t M, void *mean_out, void *sd_out,
                                  void *stream) {
    GPG_REQUIRE(h && theta && Xu && Ui && Pm && w && mean_out && sd_out, "NULL argument");
    GPG_REQUIRE((split == nullptr) == (scales == nullptr), "split and scales go together");
    GPG_REQUIRE(split == nullptr || (dtype == GPG_F32 && ld % 8 == 0), "the split factors need f32 and ld % 8 == 0");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(M == 0 || Xs != nullptr, "Xs is NULL");
    GPG_REQUIRE(m > 0 && M >= 0 && ld >= m, "bad size");
    GPG_REQUIRE(kernel_id >= 0 && kernel_id <= 2, "unknown kernel id");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return sgp_predict_entry<float>(h, kernel_id, d, (const float *)theta, (const float *)Xu, m, (const float *)Ui,
                                        (const float *)Pm, ld, (const float *)w, split, scales, (const float *)Xs, M,
                                        (float *)mean_out, (float *)sd_out, s);
    if (dtype == GPG_F64)
        return sgp_predict_entry<double>(h, kernel_id, d, (const double *)theta, (const double *)Xu, m, (const double *)Ui,
                                         (const double *)Pm, ld, (const double *)w, nullptr, nullptr, (const double *)Xs, M,
                                         (double *)mean_out, (double *)sd_out, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}
