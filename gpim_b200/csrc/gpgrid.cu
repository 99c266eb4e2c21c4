// gpgrid.cu -- extern "C" entry points of libgpgrid.so (see include/gpgrid.h) and the drivers
// that sequence the kernels of kmat.cuh / factor.cuh / train.cuh / acq.cuh / gemm_*.cuh.
#include <cstdarg>
#include <cstring>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "kmat.cuh"
#include "factor.cuh"
#include "chol_panel.cuh"
#include "factor_tc.cuh"
#include "train.cuh"
#include "acq.cuh"
#include "sparse.cuh"
#include "comm.cuh"

// ---------------------------------------------------------------------------------------------
// errors, handle, workspace
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void gpg_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int gpg_ws_reserve(gpg_handle_s *h, size_t bytes, void **out) {
    if (bytes > h->ws_bytes) {
        GPG_CUDA_CHECK(cudaDeviceSynchronize());
        if (h->ws) GPG_CUDA_CHECK(cudaFree(h->ws));
        h->ws = nullptr;
        h->ws_bytes = 0;
        const size_t want = gpg_align_up(bytes + bytes / 8, 1 << 20);
        GPG_CUDA_CHECK(cudaMalloc(&h->ws, want));
        h->ws_bytes = want;
    }
    *out = h->ws;
    return GPG_OK;
}

int gpg_gemv_part_reserve(gpg_handle_s *h, size_t elems, double **out) {
    if (elems > h->gemv_part_elems) {
        GPG_CUDA_CHECK(cudaDeviceSynchronize());
        if (h->gemv_part) GPG_CUDA_CHECK(cudaFree(h->gemv_part));
        h->gemv_part = nullptr;
        h->gemv_part_elems = 0;
        const size_t want = elems + elems / 4 + 1024;
        GPG_CUDA_CHECK(cudaMalloc(&h->gemv_part, want * sizeof(double)));
        h->gemv_part_elems = want;
    }
    *out = h->gemv_part;
    return GPG_OK;
}

int gpg_tc_counters(gpg_handle_s *h, cudaStream_t stream, int n, int **out) {
    constexpr int POOL = 8192;
    // one pool per stream the library launches on: the pool is re-zeroed stream-ordered when it wraps, which is only
    // ordered against launches of the SAME stream
    const int which = (h->side_stream != nullptr && stream == h->side_stream) ? 1 : 0;
    int *&pool = which ? h->tc_counters_side : h->tc_counters;
    int &pos = which ? h->tc_counter_pos_side : h->tc_counter_pos;
    if (!pool) {
        GPG_CUDA_CHECK(cudaMalloc(&pool, POOL * sizeof(int)));
        pos = POOL;
    }
    if (pos + n > POOL) {
        GPG_CUDA_CHECK(cudaMemsetAsync(pool, 0, POOL * sizeof(int), stream));
        pos = 0;
    }
    *out = pool + pos;
    pos += n;
    return GPG_OK;
}
int gpg_tc_counter(gpg_handle_s *h, cudaStream_t stream, int **out) { return gpg_tc_counters(h, stream, 1, out); }

struct Bump {            // carve a reserved workspace into aligned pieces
    unsigned char *base;
    size_t off = 0;
    explicit Bump(void *b) : base(reinterpret_cast<unsigned char *>(b)) {}
    template <typename U> U *take(size_t count) {
        off = gpg_align_up(off, 1024);
        U *p = reinterpret_cast<U *>(base + off);
        off += count * sizeof(U);
        return p;
    }
};
static size_t bump_size(std::initializer_list<size_t> parts) {
    size_t off = 0;
    for (size_t p : parts) off = gpg_align_up(off, 1024) + p;
    return off + 1024;
}

extern "C" int gpg_version(void) { return GPG_VERSION; }
extern "C" const char *gpg_last_error(void) { return g_err; }

// Opt-in shared-memory sizes are per device: set them for the handle's device when it is created (the lazy
// process-wide guards next to the launches only cover the first device a process touches).
static int set_function_attributes() {
    GPG_CUDA_CHECK(cudaFuncSetAttribute(diag_block_kernel<float, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        diag_block_smem<float, 128>()));
    GPG_CUDA_CHECK(cudaFuncSetAttribute(diag_block_kernel<double, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        diag_block_smem<double, 64>()));
    GPG_CUDA_CHECK(cudaFuncSetAttribute(panel_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, panel_trsm_smem()));
    GPG_CUDA_CHECK(cudaFuncSetAttribute(tc::gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    GPG_CUDA_CHECK(cudaFuncSetAttribute(cpanel::chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cpanel::SMEM_BYTES));
    GPG_CUDA_CHECK(cudaFuncSetAttribute(topk_round_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        2048 * (int)sizeof(Cand<float>)));
    GPG_CUDA_CHECK(cudaFuncSetAttribute(topk_round_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        2048 * (int)sizeof(Cand<double>)));
    return GPG_OK;
}

extern "C" int gpg_create(int device, gpg_handle_t *out) {
    GPG_REQUIRE(out != nullptr, "out is NULL");
    int count = 0;
    GPG_CUDA_CHECK(cudaGetDeviceCount(&count));
    GPG_REQUIRE(device >= 0 && device < count, "device index out of range");
    DeviceGuard device_guard(device);              // the caller's current device is restored on return
    cudaDeviceProp prop;
    GPG_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        gpg_set_error("libgpgrid is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
        return GPG_ECUDA;
    }
    GPG_TRY(set_function_attributes());
    gpg_handle_s *h = new gpg_handle_s();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    *out = h;
    return GPG_OK;
}

extern "C" int gpg_destroy(gpg_handle_t h) {
    if (!h) return GPG_OK;
    gpg_comm_destroy(h);
    if (h->ws) cudaFree(h->ws);
    if (h->tc_counters) cudaFree(h->tc_counters);
    if (h->tc_counters_side) cudaFree(h->tc_counters_side);
    if (h->work_counter) cudaFree(h->work_counter);
    if (h->gemv_part) cudaFree(h->gemv_part);
    if (h->fit_stream) cudaStreamDestroy(h->fit_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_side) cudaEventDestroy(h->ev_side);
    for (auto &sp : h->spans) { cudaEventDestroy(sp.beg); cudaEventDestroy(sp.end); }
    for (auto &e : h->event_pool) cudaEventDestroy(e);
    delete h;
    return GPG_OK;
}

extern "C" int gpg_set_option(gpg_handle_t h, int key, long long value) {
    GPG_REQUIRE(h != nullptr, "handle is NULL");
    switch (key) {
        case GPG_OPT_GEMM_PATH: GPG_REQUIRE(value >= 0 && value <= 2, "gemm path 0..2"); h->opt_gemm_path = (int)value; break;
        case GPG_OPT_PREDICT_CHUNK: GPG_REQUIRE(value >= 0, "chunk >= 0"); h->opt_predict_chunk = value; break;
        case GPG_OPT_STAGE_TIMING: h->opt_stage_timing = value != 0; break;
        case GPG_OPT_FIT_GRAPH: h->opt_fit_graph = value != 0; break;
        case GPG_OPT_INNER_LEFT: h->opt_inner_left = value != 0; break;
        case GPG_OPT_COMPACT_SUPPORT: h->opt_compact_support = value != 0; break;
        case GPG_OPT_OUTER_PANEL: GPG_REQUIRE(value >= 128 && value % 128 == 0, "outer panel: a multiple of 128"); h->opt_outer_panel = (int)value; break;
        case GPG_OPT_PANEL_MODE: GPG_REQUIRE(value >= 0 && value <= 3, "panel mode 0..3"); h->opt_panel_mode = (int)value; break;
        case GPG_OPT_FACTOR_ALGO: GPG_REQUIRE(value == 0 || value == 1, "factor algorithm 0..1"); h->opt_factor_algo = (int)value; break;
        case GPG_OPT_PANEL_REFINE: h->opt_panel_refine = value != 0; break;
        case GPG_OPT_LOOKAHEAD: h->opt_lookahead = value != 0; break;
        case GPG_OPT_PANEL_WORKERS: GPG_REQUIRE(value >= 0, "panel workers >= 0"); h->opt_panel_workers = (int)value; break;
        case GPG_OPT_SYRK_CHUNK: GPG_REQUIRE(value >= 0 && value % 64 == 0, "chunk must be a multiple of 64"); h->opt_syrk_chunk = (int)value; break;
        default: gpg_set_error("unknown option %d", key); return GPG_EINVAL;
    }
    return GPG_OK;
}

extern "C" int gpg_stage_times(gpg_handle_t h, double *ms_host, long long *spans_host) {
    GPG_REQUIRE(h && ms_host && spans_host, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_CUDA_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < GPG_ST_COUNT; ++i) { ms_host[i] = 0.0; spans_host[i] = 0; }
    for (auto &sp : h->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.beg, sp.end) == cudaSuccess && sp.stage >= 0 && sp.stage < GPG_ST_COUNT) {
            ms_host[sp.stage] += ms;
            spans_host[sp.stage] += 1;
        }
        h->event_pool.push_back(sp.beg);
        h->event_pool.push_back(sp.end);
    }
    h->spans.clear();
    return GPG_OK;
}

extern "C" long long gpg_launch_count(gpg_handle_t h) { return h ? h->launches : -1; }

extern "C" int gpg_variance_gemm_macs(gpg_handle_t h, double *macs_host) {
    GPG_REQUIRE(h && macs_host, "NULL argument");
    DeviceGuard device_guard(h->device);
    *macs_host = 0.0;
    if (!h->work_counter) return GPG_OK;
    GPG_CUDA_CHECK(cudaDeviceSynchronize());
    unsigned long long kb = 0;
    GPG_CUDA_CHECK(cudaMemcpy(&kb, h->work_counter, sizeof(kb), cudaMemcpyDeviceToHost));
    GPG_CUDA_CHECK(cudaMemset(h->work_counter, 0, sizeof(kb)));
    *macs_host = (double)kb * (double)tc::BM * (double)tc::BN * (double)tc::BK;
    return GPG_OK;
}
extern "C" size_t gpg_workspace_bytes(gpg_handle_t h) { return h ? h->ws_bytes : 0; }

// ---------------------------------------------------------------------------------------------
// GEMM dispatch: tcgen05 split-fp16 path for large fp32 problems, SIMT otherwise
// ---------------------------------------------------------------------------------------------
template <> int gemm_dispatch<double>(gpg_handle_s *h, const GemmArgs<double> &g, cudaStream_t stream) {
    if (g.epi == GEMM_EPI_STORE && gemm_f64_uses_dmma(g) && h->opt_gemm_path != 1) return gemm_dmma(h, g, stream);
    return gemm_simt<double>(h, g, stream);
}
template <> int gemm_dispatch<float>(gpg_handle_s *h, const GemmArgs<float> &g, cudaStream_t stream) {
    return gemm_simt<float>(h, g, stream);
}

// ---------------------------------------------------------------------------------------------
// fp32-faithful tensor-core GEMM as a stand-alone entry (unit tests, and callers that want the
// split-fp16 tcgen05 kernel on their own matrices):  C = alpha * A B^T + beta * C
// ---------------------------------------------------------------------------------------------
__global__ void set_scales_kernel(float sa, float sb, float *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = sa; out[1] = sb; out[2] = 1.0f / (sa * sb); out[3] = 1.0f; }
}

extern "C" int gpg_gemm_nt_f32(gpg_handle_t h, const float *A, int64_t lda, const float *B, int64_t ldb, float *C,
                               int64_t ldc, int64_t M, int64_t N, int64_t K, double alpha, double beta, double scale_a,
                               double scale_b, void *stream) {
    GPG_REQUIRE(h && A && B && C, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldb >= K && ldc >= N, "bad size");
    GPG_REQUIRE(scale_a > 0 && scale_b > 0, "scales must be positive");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int64_t ldk = gpg_align_up((size_t)K, 64);
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, bump_size({(size_t)M * ldk * 2, (size_t)M * ldk * 2, (size_t)N * ldk * 2, (size_t)N * ldk * 2,
                                         64, 64}), &ws));
    Bump b(ws);
    __half *Ahi = b.take<__half>((size_t)M * ldk), *Alo = b.take<__half>((size_t)M * ldk);
    __half *Bhi = b.take<__half>((size_t)N * ldk), *Blo = b.take<__half>((size_t)N * ldk);
    float *scales = b.take<float>(16);
    set_scales_kernel<<<1, 32, 0, s>>>((float)scale_a, (float)scale_b, scales);
    GPG_LAUNCH_CHECK(h);
    GPG_TRY(tc::split_matrix(h, A, lda, M, K, scales, Ahi, Alo, ldk, 0, s));
    GPG_TRY(tc::split_matrix(h, B, ldb, N, K, scales + 1, Bhi, Blo, ldk, 0, s));
    tc::Launch g;
    memset(&g.p, 0, sizeof(g.p));
    g.A.hi = Ahi; g.A.lo = Alo; g.A.rows = M; g.A.cols = K; g.A.ld = ldk;
    g.B.hi = Bhi; g.B.lo = Blo; g.B.rows = N; g.B.cols = K; g.B.ld = ldk;
    g.p.M = (int)M; g.p.N = (int)N; g.p.K = (int)K; g.p.batch = 1;
    g.p.epi = tc::EPI_STORE;
    g.p.scale_inv = scales + 2;
    g.p.C = C; g.p.ldc = ldc;
    g.p.alpha = (float)alpha; g.p.beta = (float)beta;
    return tc::launch(h, g, s);
}

// ---------------------------------------------------------------------------------------------
// K1 / K2
// ---------------------------------------------------------------------------------------------
template <typename T>
static int kmat_launch(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *X, int64_t N, const T *Z,
                       int64_t P, double jitter, int lower_only, T *out, int64_t ld, cudaStream_t stream,
                       KmatSplit sp = KmatSplit()) {
    const int sym = (Z == nullptr);
    if (sym) { Z = X; P = N; }
    if (N <= 0 || P <= 0) return GPG_OK;
    dim3 block(64, 4);
    dim3 grid((unsigned)((P + 255) / 256), (unsigned)((N + 15) / 16));
    GPG_DISPATCH_KID(kernel_id, GPG_DISPATCH_D(d, kmat_kernel<T, KID, D><<<grid, block, 0, stream>>>(
                                                       theta, X, N, Z, P, sym, (T)jitter, sym ? lower_only : 0, out, ld, sp)));
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

extern "C" int gpg_kmat(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta, const void *X, int64_t N,
                        const void *Z, int64_t P, double jitter, int lower_only, void *out, int64_t ld, void *stream) {
    GPG_REQUIRE(h && theta && X && out, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(N >= 0 && (Z == nullptr || P >= 0), "negative size");
    GPG_REQUIRE(ld >= (Z ? P : N), "ld smaller than row length");
    GPG_REQUIRE((int64_t)((N + 15) / 16) < 65536, "N too large for one launch");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return kmat_launch<float>(h, kernel_id, d, (const float *)theta, (const float *)X, N, (const float *)Z, P, jitter,
                                  lower_only, (float *)out, ld, s);
    if (dtype == GPG_F64)
        return kmat_launch<double>(h, kernel_id, d, (const double *)theta, (const double *)X, N, (const double *)Z, P,
                                   jitter, lower_only, (double *)out, ld, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

// ---------------------------------------------------------------------------------------------
// K3, trtri, K7a
// ---------------------------------------------------------------------------------------------
template <typename T> static int cholesky_entry(gpg_handle_s *h, T *A, int64_t N, int64_t ld, int32_t *info, cudaStream_t s) {
    constexpr int NB = GemmCfg<T>::BN;
    StageTimer st(h, GPG_ST_CHOLESKY, s);
    if constexpr (std::is_same<T, float>::value) {
        if ((ld % 8) == 0 && h->opt_gemm_path != 1 && (h->opt_gemm_path == 2 || N >= 1024)) {
            void *ws;
            GPG_TRY(gpg_ws_reserve(h, bump_size({(size_t)N * ld * 4, NB * NB * sizeof(T), SC_COUNT * sizeof(float)}), &ws));
            Bump b(ws);
            TcPlanes Ls(b.take<unsigned char>((size_t)N * ld * 4), N, ld);
            float *dinv = b.take<float>(NB * NB);
            float *scales = b.take<float>(SC_COUNT);
            scales_from_diag_kernel<<<1, 256, 0, s>>>(A, ld, N, scales);
            GPG_LAUNCH_CHECK(h);
            return cholesky_blocked_tc(h, A, N, ld, info, 1, dinv, Ls, scales, s);
        }
    }
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, bump_size({NB * NB * sizeof(T)}), &ws));
    Bump b(ws);
    return cholesky_blocked<T>(h, A, N, ld, info, 1, b.take<T>(NB * NB), s);
}

extern "C" int gpg_cholesky(gpg_handle_t h, int dtype, void *A, int64_t N, int64_t ld, int32_t *info, void *stream) {
    GPG_REQUIRE(h && A && info, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(N >= 0 && ld >= N, "bad size");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (N == 0) return GPG_OK;
    if (dtype == GPG_F32) return cholesky_entry<float>(h, (float *)A, N, ld, info, s);
    if (dtype == GPG_F64) return cholesky_entry<double>(h, (double *)A, N, ld, info, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

template <typename T>
static int trtri_entry(gpg_handle_s *h, const T *L, int64_t N, int64_t ld, T *Linv, int64_t ldi, cudaStream_t s) {
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, bump_size({(size_t)N * ld * sizeof(T)}), &ws));
    Bump b(ws);
    return trtri_blocked<T>(h, L, N, ld, Linv, ldi, b.take<T>((size_t)N * ld), s);
}

extern "C" int gpg_trtri(gpg_handle_t h, int dtype, const void *L, int64_t N, int64_t ld, void *Linv, int64_t ldinv,
                         void *stream) {
    GPG_REQUIRE(h && L && Linv && L != Linv, "NULL or aliased argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(N >= 0 && ld >= N && ldinv >= N, "bad size");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (N == 0) return GPG_OK;
    if (dtype == GPG_F32) return trtri_entry<float>(h, (const float *)L, N, ld, (float *)Linv, ldinv, s);
    if (dtype == GPG_F64) return trtri_entry<double>(h, (const double *)L, N, ld, (double *)Linv, ldinv, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

template <typename T>
static int solve_entry(gpg_handle_s *h, const T *L, const T *Linv, int64_t N, int64_t ld, const T *y, T *vhat, T *alpha,
                       T *scalars, cudaStream_t s) {
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, bump_size({2 * (size_t)N * sizeof(T)}), &ws));
    Bump b(ws);
    return solve_vec_refined<T>(h, L, Linv, N, ld, y, vhat, alpha, scalars, b.take<T>(2 * N), s);
}

extern "C" int gpg_solve_vec(gpg_handle_t h, int dtype, const void *L, const void *Linv, int64_t N, int64_t ld,
                             const void *y, void *vhat_out, void *alpha_out, void *scalars_out, void *stream) {
    GPG_REQUIRE(h && L && Linv && y && vhat_out && alpha_out, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(N > 0 && ld >= N, "bad size");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return solve_entry<float>(h, (const float *)L, (const float *)Linv, N, ld, (const float *)y, (float *)vhat_out,
                                  (float *)alpha_out, (float *)scalars_out, s);
    if (dtype == GPG_F64)
        return solve_entry<double>(h, (const double *)L, (const double *)Linv, N, ld, (const double *)y,
                                   (double *)vhat_out, (double *)alpha_out, (double *)scalars_out, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

// Scratch of one factorisation.  SIMT path: tmp (N x ld of T).  Tensor-core path (f32): the three
// fp16 plane pairs Ls / WTs / TTs of factor_tc.cuh (the fourth, Ws, is caller memory: it is part
// of the factor cache gpg_predict consumes).
template <typename T> struct FactorWs {
    T *tmp = nullptr, *dinv = nullptr, *vs = nullptr;       // vs: 4 N vector scratch
    double *best = nullptr;
    float *bbox = nullptr;            // boxes of 32-row blocks of X (support ranges of the residual kernel), f32 only
    void *planes = nullptr;           // Ls | WTs | TTs | As (2 N ld halves each) | Rf (N ld floats): tensor-core path
};

static bool tc_factor_wanted(const gpg_handle_s *h, int64_t N, int64_t ld, const void *wsplit) {
    return wsplit != nullptr && (ld % 8) == 0 && h->opt_gemm_path != 1 && (h->opt_gemm_path == 2 || N >= 1024);
}

template <typename T> static size_t factor_ws_bytes(const gpg_handle_s *h, int64_t N, int64_t ld, const void *wsplit) {
    constexpr int NB = GemmCfg<T>::BN;
    const bool tcp = std::is_same<T, float>::value && tc_factor_wanted(h, N, ld, wsplit);
    const size_t big = tcp ? 5 * (size_t)N * ld * 4 : (size_t)N * ld * sizeof(T);
    return bump_size({big, NB * NB * sizeof(T), 4 * (size_t)N * sizeof(T), sizeof(double),
                      (size_t)((N + 31) / 32) * 2 * GPG_MAX_D * sizeof(float)});
}

template <typename T> static FactorWs<T> factor_ws_carve(const gpg_handle_s *h, Bump &b, int64_t N, int64_t ld, const void *wsplit) {
    constexpr int NB = GemmCfg<T>::BN;
    const bool tcp = std::is_same<T, float>::value && tc_factor_wanted(h, N, ld, wsplit);
    FactorWs<T> w;
    if (tcp) w.planes = b.take<unsigned char>(5 * (size_t)N * ld * 4);
    else w.tmp = b.take<T>((size_t)N * ld);
    w.dinv = b.take<T>(NB * NB);
    w.vs = b.take<T>(4 * N);
    w.best = b.take<double>(1);
    w.bbox = b.take<float>((size_t)((N + 31) / 32) * 2 * GPG_MAX_D);
    return w;
}

// Iterative refinement of alpha = K^-1 y against K ITSELF: the residual y - K alpha is formed by
// re-evaluating the covariance function (no N x N memory traffic), the correction goes through
// W^T W ~= K^-1.  This removes the factorisation's backward error (split-fp16 contractions, panels
// through explicit inverses) from the predictive mean.  Each step is guarded on the device: a
// candidate is kept only if it lowers |y - K alpha|.  scratch: 4 N elements of T + one double.
__global__ void set_double_kernel(double *p, double v) { if (threadIdx.x == 0 && blockIdx.x == 0) *p = v; }

template <typename T>
static int refine_alpha_against_K(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *X, const T *y,
                                  int64_t N, double jitter, const T *Linv, int64_t ld, T *alpha, T *scratch,
                                  double *best, int rounds, cudaStream_t s, float *bbox = nullptr) {
    T *r = scratch, *t = scratch + N, *an = scratch + 2 * N, *rn = scratch + 3 * N;
    const unsigned gk = (unsigned)((N + 7) / 8), gN = gk;
    const int nblk32 = (int)((N + 31) / 32);
    if (bbox && !h->opt_compact_support) bbox = nullptr;
    if (bbox) {                      // the residual only visits the training rows inside each point's support
        GPG_DISPATCH_D(d, block_bbox_kernel<T, D><<<(unsigned)((nblk32 + 7) / 8), 256, 0, s>>>(X, N, bbox));
        GPG_LAUNCH_CHECK(h);
    }
    auto resid = [&](const T *a, T *out) -> int {
        GPG_DISPATCH_KID(kernel_id, GPG_DISPATCH_D(d, {
            TestPoints<T, D> tp;
            tp.Xs = X; tp.j0 = 0;
            for (int k = 0; k < GPG_MAX_D; ++k) { tp.dims[k] = 1; tp.step[k] = T(1); }
            kcross_mean_kernel<T, KID, D, false><<<gk, 256, 0, s>>>(theta, X, N, tp, N, a, nullptr, 0, nullptr, nullptr, 0,
                                                                     nullptr, out, y, (T)jitter, nullptr, 1e-14f, bbox, nblk32);
        }));
        GPG_LAUNCH_CHECK(h);
        return GPG_OK;
    };
    set_double_kernel<<<1, 32, 0, s>>>(best, -1.0);
    GPG_LAUNCH_CHECK(h);
    GPG_TRY(resid(alpha, r));
    refine_select_kernel<T><<<1, 1024, 0, s>>>(N, alpha, r, alpha, r, best);
    GPG_LAUNCH_CHECK(h);
    for (int it = 0; it < rounds; ++it) {
        gemv_tri_kernel<T, false><<<gN, 256, 0, s>>>(Linv, ld, N, r, nullptr, T(1), T(0), t);          // t = W r
        GPG_LAUNCH_CHECK(h);
        GPG_TRY(gemv_tri_T<T>(h, Linv, ld, N, t, alpha, T(1), T(1), an, s));                            // an = alpha + W^T t
        GPG_TRY(resid(an, rn));
        refine_select_kernel<T><<<1, 1024, 0, s>>>(N, an, rn, alpha, r, best);
        GPG_LAUNCH_CHECK(h);
    }
    return GPG_OK;
}

// K1 + K3 + trtri + K7a: the factor cache for the theta stored on the device.
// wsplit / scales (f32, nullable together): tensor-core form of Linv for gpg_predict.
template <typename T>
static int factorize_core(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *X, const T *y, int64_t N,
                          double jitter, T *L, T *Linv, int64_t ld, T *vhat, T *alpha, T *scalars, int32_t *info,
                          int reset_info, const FactorWs<T> &w, void *wsplit, float *scales, cudaStream_t s) {
    bool done = false;
    if constexpr (std::is_same<T, float>::value) {
        if (wsplit) {
            scales_from_theta_kernel<float><<<1, 32, 0, s>>>(theta, (float)jitter, (float)N, scales);
            GPG_LAUNCH_CHECK(h);
        }
        if (w.planes) {
            const size_t pl = (size_t)N * ld * 4;
            unsigned char *base = (unsigned char *)w.planes;
            TcPlanes Ls(base, N, ld), WTs(base + pl, N, ld), TTs(base + 2 * pl, N, ld), As(base + 3 * pl, N, ld);
            TcPlanes Ws(wsplit, N, ld);
            if (h->opt_factor_algo == 1) {
                {
                    StageTimer st(h, GPG_ST_KMAT, s);
                    KmatSplit sp;
                    sp.hi = As.hi; sp.lo = As.lo; sp.ld = ld; sp.scale = scales + SC_A;
                    GPG_TRY(kmat_launch<T>(h, kernel_id, d, theta, X, N, nullptr, N, jitter, 1, L, ld, s, sp));
                }
                // Cholesky and inverse are interleaved by the recursion; the stage clock books both under CHOLESKY
                float *Rf = reinterpret_cast<float *>(base + 4 * pl);
                StageTimer st(h, GPG_ST_CHOLESKY, s);
                GPG_TRY(potrf_inv_tc(h, L, N, ld, Linv, Rf, info, reset_info, As, Ls, Ws, WTs, TTs, scales, s));
            } else {
                {
                    StageTimer st(h, GPG_ST_KMAT, s);
                    KmatSplit sp;                // fp16 planes of the first panel (the updates keep the next one current)
                    sp.hi = As.hi; sp.lo = As.lo; sp.ld = ld; sp.scale = scales + SC_A; sp.ncols = 128;
                    GPG_TRY(kmat_launch<T>(h, kernel_id, d, theta, X, N, nullptr, N, jitter, 1, L, ld, s, sp));
                }
                { StageTimer st(h, GPG_ST_CHOLESKY, s); GPG_TRY(cholesky_blocked_tc(h, L, N, ld, info, reset_info, w.dinv, Ls, scales, s, As, Ws)); }
                { StageTimer st(h, GPG_ST_TRTRI, s); GPG_TRY(trtri_tc(h, L, N, ld, Linv, Ls, Ws, WTs, TTs, scales, s)); }
            }
            done = true;
        }
    }
    if (!done) { StageTimer st(h, GPG_ST_KMAT, s); GPG_TRY(kmat_launch<T>(h, kernel_id, d, theta, X, N, nullptr, N, jitter, 1, L, ld, s)); }
    if (!done) {
        { StageTimer st(h, GPG_ST_CHOLESKY, s); GPG_TRY(cholesky_blocked<T>(h, L, N, ld, info, reset_info, w.dinv, s)); }
        { StageTimer st(h, GPG_ST_TRTRI, s); GPG_TRY(trtri_blocked<T>(h, L, N, ld, Linv, ld, w.tmp, s)); }
        if constexpr (std::is_same<T, float>::value) {
            if (wsplit) {
                StageTimer st(h, GPG_ST_TRTRI, s);
                TcPlanes Ws(wsplit, N, ld);
                GPG_TRY(tc::split_matrix(h, Linv, ld, N, N, scales + SC_W, Ws.hi, Ws.lo, ld, 1, s));
            }
        }
    }
    {
        StageTimer st(h, GPG_ST_SOLVE, s);
        GPG_TRY(solve_vec_refined<T>(h, L, Linv, N, ld, y, vhat, alpha, nullptr, w.vs, s, 0));
        // The refinement against K removes the backward error of the fp32 factorisation (split-fp16 contractions,
        // panels through explicit inverses).  In fp64 the factor cache is accurate to ~cond(L) eps already and the
        // small-N regime of the Bayesian-optimisation loop is bound by the NUMBER of dependent launches per Adam
        // iteration: there the 13 launches of the two refinement rounds are left out.
        if (sizeof(T) == 4)
            GPG_TRY(refine_alpha_against_K<T>(h, kernel_id, d, theta, X, y, N, jitter, Linv, ld, alpha, w.vs, w.best, 2, s, w.bbox));
        if (scalars) {               // 0.5 y^T K^-1 y from the refined alpha, log-determinant from diag L
            solve_scalars_kernel<T><<<1, (N >= 1024 ? 1024 : 256), 0, s>>>(L, ld, N, vhat, scalars, y, alpha);
            GPG_LAUNCH_CHECK(h);
        }
    }
    return GPG_OK;
}

template <typename T>
static int factorize_entry(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *X, const T *y, int64_t N,
                           double jitter, T *L, T *Linv, int64_t ld, T *vhat, T *alpha, T *scalars, int32_t *info,
                           void *wsplit, float *scales, cudaStream_t s) {
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, factor_ws_bytes<T>(h, N, ld, wsplit), &ws));
    Bump b(ws);
    FactorWs<T> w = factor_ws_carve<T>(h, b, N, ld, wsplit);
    return factorize_core<T>(h, kernel_id, d, theta, X, y, N, jitter, L, Linv, ld, vhat, alpha, scalars, info, 1, w,
                             wsplit, scales, s);
}

extern "C" int gpg_factorize(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta, const void *X,
                             const void *y, int64_t N, double jitter, void *L, void *Linv, int64_t ld, void *vhat_out,
                             void *alpha_out, void *scalars_out, int32_t *info, void *wsplit_out, float *scales_out,
                             void *stream) {
    GPG_REQUIRE(h && theta && X && y && L && Linv && vhat_out && alpha_out && info, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(N > 0 && ld >= N, "bad size");
    GPG_REQUIRE(d >= 1 && d <= GPG_MAX_D, "d not in 1..4");
    GPG_REQUIRE((wsplit_out == nullptr) == (scales_out == nullptr), "wsplit_out and scales_out go together");
    GPG_REQUIRE(wsplit_out == nullptr || (dtype == GPG_F32 && ld % 8 == 0), "the split factor needs f32 and ld % 8 == 0");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return factorize_entry<float>(h, kernel_id, d, (const float *)theta, (const float *)X, (const float *)y, N, jitter,
                                      (float *)L, (float *)Linv, ld, (float *)vhat_out, (float *)alpha_out,
                                      (float *)scalars_out, info, wsplit_out, scales_out, s);
    if (dtype == GPG_F64)
        return factorize_entry<double>(h, kernel_id, d, (const double *)theta, (const double *)X, (const double *)y, N,
                                       jitter, (double *)L, (double *)Linv, ld, (double *)vhat_out, (double *)alpha_out,
                                       (double *)scalars_out, info, nullptr, nullptr, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

// ---------------------------------------------------------------------------------------------
// K2 + K4 + K5: predict
// ---------------------------------------------------------------------------------------------
template <typename T, int D>
static int predict_core(gpg_handle_s *h, int kernel_id, const T *theta, const T *X, int64_t N, const T *Linv, int64_t ld,
                        const T *alpha, TestPoints<T, D> tp, int64_t M, T *mean, T *sd, cudaStream_t s) {
    using C = GemmCfg<T>;
    const int64_t ldk = gpg_align_up((size_t)N, 64);
    int64_t chunk = h->opt_predict_chunk;
    if (chunk <= 0) {
        chunk = 8192;
        while (chunk > 512 && chunk * ldk * (int64_t)sizeof(T) > (int64_t)768 << 20) chunk /= 2;
    }
    chunk = std::min<int64_t>(chunk, gpg_align_up((size_t)M, 128));
    // fp64: the variance product runs on DMMA tiles of 128 rows when the problem fills them (GPG_OPT_GEMM_PATH = 1 keeps SIMT)
    bool dmma = false;
    if constexpr (std::is_same<T, double>::value) dmma = N >= 256 && h->opt_gemm_path != 1;
    const int bm = dmma ? 128 : C::BM;
    const int tiles_m = (int)((N + bm - 1) / bm);
    const int nblk32 = (int)((N + 31) / 32);
    const int groups = (int)((chunk + 127) / 128);
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, bump_size({(size_t)chunk * ldk * sizeof(T), (size_t)tiles_m * chunk * sizeof(T),
                                         2 * (size_t)groups * sizeof(int), (size_t)nblk32 * 2 * D * sizeof(float)}), &ws));
    Bump b(ws);
    T *Ks = b.take<T>((size_t)chunk * ldk);
    T *part = b.take<T>((size_t)tiles_m * chunk);
    // compact support of K* (GPG_OPT_COMPACT_SUPPORT): per group of 128 test points the range of training rows that
    // can matter at the resolution of T -- found geometrically by the K* kernel, honoured by the product
    int *krange = h->opt_compact_support ? b.take<int>(2 * (size_t)groups) : nullptr;
    float *bbox = h->opt_compact_support ? b.take<float>((size_t)nblk32 * 2 * D) : nullptr;
    const float support_rel = sizeof(T) == 4 ? 1e-14f : 1e-30f;
    if (bbox) {
        StageTimer st(h, GPG_ST_KCROSS, s);
        block_bbox_kernel<T, D><<<(unsigned)((nblk32 + 7) / 8), 256, 0, s>>>(X, N, bbox);
        GPG_LAUNCH_CHECK(h);
    }
    for (int64_t c0 = 0; c0 < M; c0 += chunk) {
        const int64_t mc = std::min<int64_t>(chunk, M - c0);
        TestPoints<T, D> tpc = tp;
        if (tpc.Xs) tpc.Xs += c0 * D; else tpc.j0 += c0;
        const unsigned gk = (unsigned)((mc + 7) / 8);
        {
            StageTimer st(h, GPG_ST_KCROSS, s);
            GPG_DISPATCH_KID(kernel_id, kcross_mean_kernel<T, KID, D, false><<<gk, 256, 0, s>>>(
                                            theta, X, N, tpc, mc, alpha, Ks, ldk, nullptr, nullptr, 0, nullptr, mean + c0,
                                            nullptr, T(0), krange, support_rel, bbox, nblk32));
            GPG_LAUNCH_CHECK(h);
        }
        GemmArgs<T> g;           // colsum((Linv Ks^T)^2): C[i][j] = sum_k Linv[i][k] Ks[j][k], k <= i
        g.A = Linv; g.lda = ld; g.a_kmajor = 1;
        g.B = Ks; g.ldb = ldk; g.b_kmajor = 1;
        g.M = (int)N; g.N = (int)mc; g.K = (int)N;
        g.ke_mode = GEMM_KE_M;
        g.epi = GEMM_EPI_COLSUMSQ;
        g.part = part; g.ldpart = chunk;
        g.krange = krange;
        {
            StageTimer st(h, GPG_ST_PGEMM, s);
            if constexpr (std::is_same<T, double>::value) {
                if (dmma) GPG_TRY(gemm_dmma(h, g, s));
                else GPG_TRY(gemm_simt<T>(h, g, s));
            } else {
                GPG_TRY(gemm_simt<T>(h, g, s));
            }
        }
        StageTimer st(h, GPG_ST_PFINAL, s);
        predict_finalize_kernel<T, D><<<(unsigned)((mc + 255) / 256), 256, 0, s>>>(theta, part, tiles_m, chunk, tpc, mc,
                                                                                   T(1), sd + c0);
        GPG_LAUNCH_CHECK(h);
    }
    return GPG_OK;
}

// fp32 tensor-core path: K* tiles are generated directly as fp16 hi/lo operands, the variance
// reduction runs in the epilogue of the tcgen05 GEMM (test points on the TMEM lanes).
// Row blocks of the Linv planes that are still arriving (pipelined broadcast, comm.cuh): block c covers rows
// [row_end[c-1], row_end[c]) -- boundaries are multiples of tc::BN -- and is complete once ev[c] has fired.
struct PlaneArrival {
    int nchunks = 0;
    const int64_t *row_end = nullptr;
    const cudaEvent_t *ev = nullptr;
};

template <int D>
static int predict_core_tc(gpg_handle_s *h, int kernel_id, const float *theta, const float *X, int64_t N,
                           const __half *Whi, const __half *Wlo, int64_t ldh, const float *scales, const float *alpha,
                           TestPoints<float, D> tp, int64_t M, float *mean, float *sd, cudaStream_t s,
                           const PlaneArrival *arrival = nullptr) {
    int64_t chunk = h->opt_predict_chunk;
    if (chunk <= 0) {
        chunk = 16384;               // K* planes of one chunk: at most 2 GiB
        while (chunk > 1024 && chunk * ldh * 4 > (int64_t)2 << 30) chunk /= 2;
    }
    chunk = gpg_align_up((size_t)std::min<int64_t>(chunk, gpg_align_up((size_t)M, 128)), 128);
    const int tiles_n = (int)((N + tc::BN - 1) / tc::BN);
    void *ws;
    const int mtiles = (int)(chunk / tc::BM);
    const int nblk32 = (int)((N + 31) / 32);
    GPG_TRY(gpg_ws_reserve(h, bump_size({(size_t)chunk * ldh * 2, (size_t)chunk * ldh * 2,
                                         (size_t)tiles_n * chunk * sizeof(float), 2 * (size_t)mtiles * sizeof(int),
                                         (size_t)nblk32 * 2 * D * sizeof(float)}), &ws));
    Bump b(ws);
    __half *Khi = b.take<__half>((size_t)chunk * ldh);
    __half *Klo = b.take<__half>((size_t)chunk * ldh);
    float *part = b.take<float>((size_t)tiles_n * chunk);
    int *krange = h->opt_compact_support ? b.take<int>(2 * (size_t)mtiles) : nullptr;
    float *bbox = h->opt_compact_support ? b.take<float>((size_t)nblk32 * 2 * D) : nullptr;
    if (bbox) {                      // boxes of 32-row blocks of X: what the K* kernel derives the support ranges from
        StageTimer st(h, GPG_ST_KCROSS, s);
        block_bbox_kernel<float, D><<<(unsigned)((nblk32 + 7) / 8), 256, 0, s>>>(X, N, bbox);
        GPG_LAUNCH_CHECK(h);
    }
    // A-operand groups sized to stay L2-resident while the n-blocks sweep over them
    const int m_group = (int)std::max<int64_t>(1, ((int64_t)64 << 20) / (tc::BM * ldh * 4));
    for (int64_t c0 = 0; c0 < M; c0 += chunk) {
        const int64_t mc = std::min<int64_t>(chunk, M - c0);
        TestPoints<float, D> tpc = tp;
        if (tpc.Xs) tpc.Xs += c0 * D; else tpc.j0 += c0;
        if (krange)                  // n-blocks outside a tile's support range are skipped: their partial sums must read zero
            GPG_CUDA_CHECK(cudaMemsetAsync(part, 0, (size_t)tiles_n * chunk * sizeof(float), s));
        {
            StageTimer st(h, GPG_ST_KCROSS, s);
            GPG_DISPATCH_KID(kernel_id, kcross_mean_kernel<float, KID, D, true><<<(unsigned)((mc + 7) / 8), 256, 0, s>>>(
                                            theta, X, N, tpc, mc, alpha, nullptr, 0, Khi, Klo, ldh, scales, mean + c0,
                                            nullptr, 0.f, krange, 1e-14f, bbox, nblk32, 2e-7f));
            GPG_LAUNCH_CHECK(h);
        }
        {
            StageTimer st(h, GPG_ST_PGEMM, s);
            // one launch over all n-blocks -- or, for the first tile of test points behind a pipelined broadcast,
            // one launch per row block of Linv, each gated on that block's arrival
            const int nlaunch = (arrival && c0 == 0) ? arrival->nchunks : 1;
            int64_t r0 = 0;
            for (int li = 0; li < nlaunch; ++li) {
                const int64_t r1 = nlaunch == 1 ? N : std::min<int64_t>(N, arrival->row_end[li]);
                if (nlaunch > 1) GPG_CUDA_CHECK(cudaStreamWaitEvent(s, arrival->ev[li], 0));
                if (r1 > r0) {
                    tc::Launch g;
                    memset(&g.p, 0, sizeof(g.p));
                    g.A.hi = Khi; g.A.lo = Klo; g.A.rows = mc; g.A.cols = N; g.A.ld = ldh;
                    g.B.hi = Whi; g.B.lo = Wlo; g.B.rows = N; g.B.cols = N; g.B.ld = ldh;
                    g.p.M = (int)mc; g.p.N = (int)(r1 - r0); g.p.K = (int)N; g.p.batch = 1;
                    g.p.b_row0 = (int)r0; g.p.n_off = (int)r0;
                    g.p.m_group = m_group;
                    g.p.ke_mode = GEMM_KE_N;
                    g.p.epi = tc::EPI_ROWSUMSQ;
                    g.p.scale_inv = scales + 2;
                    g.p.part = part + (r0 / tc::BN) * chunk; g.p.ldpart = chunk;
                    g.p.krange = krange;
                    if (h->opt_stage_timing) {
                        if (!h->work_counter) {
                            GPG_CUDA_CHECK(cudaMalloc(&h->work_counter, sizeof(unsigned long long)));
                            GPG_CUDA_CHECK(cudaMemsetAsync(h->work_counter, 0, sizeof(unsigned long long), s));
                        }
                        g.p.work_counter = h->work_counter;
                    }
                    GPG_TRY(tc::launch(h, g, s));
                }
                r0 = r1;
            }
        }
        StageTimer st(h, GPG_ST_PFINAL, s);
        predict_finalize_kernel<float, D><<<(unsigned)((mc + 255) / 256), 256, 0, s>>>(theta, part, tiles_n, chunk, tpc, mc,
                                                                                       1.0f, sd + c0);
        GPG_LAUNCH_CHECK(h);
    }
    return GPG_OK;
}

// THE routing rule of gpg_predict: the tcgen05 kernels consume the fp16 planes, everything else the fp32 / fp64 Linv.
// (gpg_predict_uses_planes exports it, so that whoever ships a factor cache between GPUs sends what will be read.)
static bool predict_wants_tc(const gpg_handle_s *h, int dtype, int64_t N, bool have_planes) {
    return dtype == GPG_F32 && have_planes && h->opt_gemm_path != 1 && (h->opt_gemm_path == 2 || N >= 1024);
}

template <typename T, int D>
static int predict_route(gpg_handle_s *h, int kernel_id, const T *theta, const T *X, int64_t N, const T *Linv, int64_t ld,
                         const T *alpha, const void *wsplit, const float *scales, TestPoints<T, D> tp, int64_t M, T *mean,
                         T *sd, cudaStream_t s, const PlaneArrival *arrival = nullptr) {
    if constexpr (std::is_same<T, float>::value) {
        if (predict_wants_tc(h, GPG_F32, N, wsplit != nullptr)) {
            const __half *whi = (const __half *)wsplit;
            return predict_core_tc<D>(h, kernel_id, theta, X, N, whi, whi + (size_t)N * ld, ld, scales, alpha, tp, M, mean,
                                      sd, s, arrival);
        }
    }
    return predict_core<T, D>(h, kernel_id, theta, X, N, Linv, ld, alpha, tp, M, mean, sd, s);
}

template <typename T>
static int predict_entry(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *X, int64_t N, const T *Linv,
                         int64_t ld, const T *alpha, const void *wsplit, const float *scales, const T *Xs,
                         const int64_t *dims, const double *step, int64_t j0,
                         int64_t M, T *mean, T *sd, cudaStream_t s, const PlaneArrival *arrival = nullptr) {
    if (M == 0) return GPG_OK;
    GPG_DISPATCH_D(d, {
        TestPoints<T, D> tp;
        tp.Xs = Xs;
        tp.j0 = j0;
        for (int k = 0; k < GPG_MAX_D; ++k) {
            tp.dims[k] = (dims && k < D) ? dims[k] : 1;
            tp.step[k] = (step && k < D) ? (T)step[k] : T(1);
        }
        return predict_route<T, D>(h, kernel_id, theta, X, N, Linv, ld, alpha, wsplit, scales, tp, M, mean, sd, s, arrival);
    });
    return GPG_OK;
}

extern "C" int gpg_predict(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta, const void *X, int64_t N,
                           const void *Linv, int64_t ld, const void *alpha, const void *wsplit, const float *scales,
                           const void *Xs, int64_t M, void *mean_out, void *sd_out, void *stream) {
    GPG_REQUIRE(h && theta && X && Linv && alpha && mean_out && sd_out, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(M == 0 || Xs != nullptr, "Xs is NULL");
    GPG_REQUIRE(N > 0 && M >= 0 && ld >= N, "bad size");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return predict_entry<float>(h, kernel_id, d, (const float *)theta, (const float *)X, N, (const float *)Linv, ld,
                                    (const float *)alpha, wsplit, scales, (const float *)Xs, nullptr, nullptr, 0, M,
                                    (float *)mean_out, (float *)sd_out, s);
    if (dtype == GPG_F64)
        return predict_entry<double>(h, kernel_id, d, (const double *)theta, (const double *)X, N, (const double *)Linv,
                                     ld, (const double *)alpha, nullptr, nullptr, (const double *)Xs, nullptr, nullptr, 0, M,
                                     (double *)mean_out, (double *)sd_out, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

extern "C" int gpg_predict_grid(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta, const void *X,
                                int64_t N, const void *Linv, int64_t ld, const void *alpha, const void *wsplit,
                                const float *scales, const int64_t *dims_host,
                                const double *step_host, int64_t j0, int64_t M, void *mean_out, void *sd_out,
                                void *stream) {
    GPG_REQUIRE(h && theta && X && Linv && alpha && mean_out && sd_out && dims_host && step_host, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(N > 0 && M >= 0 && ld >= N && j0 >= 0, "bad size");
    GPG_REQUIRE(d >= 1 && d <= GPG_MAX_D, "d not in 1..4");
    int64_t total = 1;
    for (int k = 0; k < d; ++k) { GPG_REQUIRE(dims_host[k] > 0, "grid dims must be positive"); total *= dims_host[k]; }
    GPG_REQUIRE(j0 + M <= total, "grid tile exceeds the grid");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return predict_entry<float>(h, kernel_id, d, (const float *)theta, (const float *)X, N, (const float *)Linv, ld,
                                    (const float *)alpha, wsplit, scales, nullptr, dims_host, step_host, j0, M,
                                    (float *)mean_out, (float *)sd_out, s);
    if (dtype == GPG_F64)
        return predict_entry<double>(h, kernel_id, d, (const double *)theta, (const double *)X, N, (const double *)Linv,
                                     ld, (const double *)alpha, nullptr, nullptr, nullptr, dims_host, step_host, j0, M,
                                     (double *)mean_out, (double *)sd_out, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

// ---------------------------------------------------------------------------------------------
// K7: nll + gradient, Adam loop
// ---------------------------------------------------------------------------------------------
struct TrainBufs {
    void *L, *Linv, *Kinv, *dinv, *vs, *vhat, *alpha, *scalars, *grad, *nll, *theta, *yc;
    float *bbox;
    void *planes;            // tensor-core path: Ls | WTs | TTs | As (Kinv aliases TTs, which is dead by then)
    void *wsplit;            // tensor-core path: Ws
    float *scales;
    double *partial;
    FitState *st;
    int64_t ld;
    int nblocks;
};

template <typename T> static bool train_uses_tc(const gpg_handle_s *h, int64_t N, int64_t ld) {
    return std::is_same<T, float>::value && tc_factor_wanted(h, N, ld, (const void *)1);
}

template <typename T> static size_t train_ws_bytes(const gpg_handle_s *h, int64_t N, int64_t ld) {
    constexpr int NB = GemmCfg<T>::BN;
    const size_t nn = (size_t)N * ld * sizeof(T);
    const size_t nb = (size_t)((N + 7) / 8);
    const bool tcp = train_uses_tc<T>(h, N, ld);
    // L, Linv + (SIMT: Kinv | TC: 4 plane pairs + Rf + Ws)
    return bump_size({nn, nn, tcp ? 6 * nn : nn, NB * NB * sizeof(T), 4 * (size_t)N * sizeof(T) + 64, (size_t)N * sizeof(T),
                      (size_t)N * sizeof(T), 2 * sizeof(T), GPG_MAX_P * sizeof(T), sizeof(T), GPG_MAX_P * sizeof(T),
                      SC_COUNT * sizeof(float), GRAD_CHUNKS_MAX * nb * GPG_MAX_P * sizeof(double), sizeof(FitState), (size_t)N * sizeof(T),
                      (size_t)((N + 31) / 32) * 2 * GPG_MAX_D * sizeof(float)});
}

template <typename T> static TrainBufs train_carve(const gpg_handle_s *h, void *ws, int64_t N, int64_t ld) {
    constexpr int NB = GemmCfg<T>::BN;
    Bump b(ws);
    TrainBufs t;
    t.ld = ld;
    t.L = b.take<T>((size_t)N * ld);
    t.Linv = b.take<T>((size_t)N * ld);
    if (train_uses_tc<T>(h, N, ld)) {
        const size_t pl = (size_t)N * ld * 4;
        unsigned char *big = b.take<unsigned char>(6 * pl);
        t.planes = big;                              // Ls | WTs | TTs | As | Rf
        t.wsplit = big + 5 * pl;
        t.Kinv = big + 2 * pl;                       // aliases TTs, dead once the factorisation is done
    } else {
        t.planes = nullptr;
        t.wsplit = nullptr;
        t.Kinv = b.take<T>((size_t)N * ld);
    }
    t.dinv = b.take<T>(NB * NB);
    t.vs = b.take<T>(4 * N + 64 / sizeof(T));        // vector scratch + the refinement's best-residual double
    t.vhat = b.take<T>(N);
    t.alpha = b.take<T>(N);
    t.scalars = b.take<T>(2);
    t.grad = b.take<T>(GPG_MAX_P);
    t.nll = b.take<T>(1);
    t.theta = b.take<T>(GPG_MAX_P);
    t.scales = b.take<float>(SC_COUNT);
    t.nblocks = (int)((N + 7) / 8);
    t.partial = b.take<double>((size_t)GRAD_CHUNKS_MAX * t.nblocks * GPG_MAX_P);
    t.st = b.take<FitState>(1);
    t.yc = b.take<T>(N);
    t.bbox = b.take<float>((size_t)((N + 31) / 32) * 2 * GPG_MAX_D);
    return t;
}

// one evaluation of nll and d nll / d theta at the theta stored in `theta`
template <typename T>
static int nll_grad_core(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *X, const T *y, int64_t N,
                         double jitter, const TrainBufs &tb, T *nll_out, T *grad_out, int32_t *info, int reset_info,
                         cudaStream_t s) {
    T *L = (T *)tb.L, *Linv = (T *)tb.Linv, *Kinv = (T *)tb.Kinv;
    FactorWs<T> w;
    w.planes = tb.planes;
    w.tmp = tb.planes ? nullptr : Kinv;          // SIMT trtri scratch; Kinv is overwritten afterwards
    w.dinv = (T *)tb.dinv;
    w.vs = (T *)tb.vs;
    w.best = reinterpret_cast<double *>((T *)tb.vs + 4 * N);       // 16 N or 32 N bytes past a 1 KB boundary: aligned
    w.bbox = tb.bbox;
    GPG_TRY(factorize_core<T>(h, kernel_id, d, theta, X, y, N, jitter, L, Linv, tb.ld, (T *)tb.vhat, (T *)tb.alpha,
                              (T *)tb.scalars, info, reset_info, w, tb.wsplit, tb.scales, s));
    StageTimer st(h, GPG_ST_GRAD, s);
    bool done = false;
    if constexpr (std::is_same<T, float>::value) {
        if (tb.planes) {             // Kinv = Linv^T Linv on tcgen05 from the transposed split of Linv
            TcPlanes WTs((unsigned char *)tb.planes + (size_t)N * tb.ld * 4, N, tb.ld);
            GPG_TRY(kinv_tc(h, N, tb.ld, WTs, Kinv, tb.scales, s));
            done = true;
        }
    }
    if (!done) {
        GemmArgs<T> g;           // Kinv = Linv^T Linv on lower tiles: sum_k Linv[k][i] Linv[k][j], k >= max(i,j)
        g.A = Linv; g.lda = tb.ld; g.a_kmajor = 0;
        g.B = Linv; g.ldb = tb.ld; g.b_kmajor = 0;
        g.C = Kinv; g.ldc = tb.ld;
        g.M = (int)N; g.N = (int)N; g.K = (int)N;
        g.kb_mode = GEMM_KB_MAXMN;
        g.tile_mode = GEMM_TILES_LOWER;
        GPG_TRY(gemm_dispatch<T>(h, g, s));
    }
    // rows split over column chunks for large N (one warp per row keeps too few loads in flight: 11 % of the HBM roofline)
    const int gchunks = (int)std::max<int64_t>(1, std::min<int64_t>(GRAD_CHUNKS_MAX, N / 1024));
    GPG_DISPATCH_KID(kernel_id, GPG_DISPATCH_D(d, grad_partial_kernel<T, KID, D><<<dim3(tb.nblocks, gchunks), 256, 0, s>>>(
                                                       theta, X, (const T *)tb.alpha, Kinv, tb.ld, N, tb.partial)));
    GPG_LAUNCH_CHECK(h);
    const double hl = 0.5 * (double)N * 1.8378770664093454835606594728112;   // N/2 log(2 pi)
    grad_finish_kernel<T><<<1, 256, 0, s>>>(tb.partial, tb.nblocks * gchunks, 3 + d, (const T *)tb.scalars, hl, grad_out, nll_out);
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

template <typename T>
static int nll_grad_entry(gpg_handle_s *h, int kernel_id, int d, const T *theta, const T *X, const T *y, int64_t N,
                          double jitter, T *nll_out, T *grad_out, int32_t *info, cudaStream_t s) {
    const int64_t ld = gpg_align_up((size_t)N, 64);
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, train_ws_bytes<T>(h, N, ld), &ws));
    TrainBufs tb = train_carve<T>(h, ws, N, ld);
    return nll_grad_core<T>(h, kernel_id, d, theta, X, y, N, jitter, tb, nll_out, grad_out, info, 1, s);
}

extern "C" int gpg_nll_grad(gpg_handle_t h, int dtype, int kernel_id, int d, const void *theta, const void *X,
                            const void *y, int64_t N, double jitter, void *nll_out, void *grad_out, int32_t *info,
                            void *stream) {
    GPG_REQUIRE(h && theta && X && y && nll_out && grad_out && info, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(N > 0, "bad size");
    GPG_REQUIRE(d >= 1 && d <= GPG_MAX_D, "d not in 1..4");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return nll_grad_entry<float>(h, kernel_id, d, (const float *)theta, (const float *)X, (const float *)y, N, jitter,
                                     (float *)nll_out, (float *)grad_out, info, s);
    if (dtype == GPG_F64)
        return nll_grad_entry<double>(h, kernel_id, d, (const double *)theta, (const double *)X, (const double *)y, N,
                                      jitter, (double *)nll_out, (double *)grad_out, info, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

template <typename T>
static int fit_entry(gpg_handle_s *h, int kernel_id, int d, int n_ls, const T *X, const T *y, int64_t N, double jitter,
                     T *u, const double *bounds, int iters, double lr, T *traj, T *theta_out, int32_t *info,
                     cudaStream_t s, int mode = 0) {
    const int64_t ld = gpg_align_up((size_t)N, 64);
    if (s == nullptr && h->opt_fit_graph && iters >= 8 && !train_uses_tc<T>(h, N, ld)) {
        // The legacy default stream cannot be captured.  A blocking stream is implicitly ordered against it in
        // both directions, so the loop can run (and be captured) there with unchanged semantics for the caller.
        if (!h->fit_stream) GPG_CUDA_CHECK(cudaStreamCreate(&h->fit_stream));
        s = h->fit_stream;
    }
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, train_ws_bytes<T>(h, N, ld), &ws));
    TrainBufs tb = train_carve<T>(h, ws, N, ld);
    FitCfg c;
    memset(&c, 0, sizeof(c));
    c.d = d; c.n_ls = n_ls; c.is_rq = (kernel_id == GPG_RATQUAD); c.mode = mode;
    c.var_lo = bounds[0]; c.var_hi = bounds[1];
    for (int k = 0; k < n_ls; ++k) { c.ls_lo[k] = bounds[2 + k]; c.ls_hi[k] = bounds[2 + n_ls + k]; }
    c.lr = lr; c.beta1 = 0.9; c.beta2 = 0.999; c.eps = 1e-8;
    T *theta = (T *)tb.theta;
    adam_step_kernel<T><<<1, 32, 0, s>>>(0, c, u, tb.st, nullptr, nullptr, theta, nullptr);
    GPG_LAUNCH_CHECK(h);
    // info keeps the FIRST failing pivot over all iterations (reset once, atomicCAS afterwards)
    GPG_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), s));
    auto one_iteration = [&]() -> int {
        const T *yt = y;
        if (mode == 1) {             // GPyTorch semantics: constant mean subtracted, loss and gradient per datum
            center_y_kernel<T><<<(unsigned)((N + 255) / 256), 256, 0, s>>>(y, N, theta, (T *)tb.yc);
            GPG_LAUNCH_CHECK(h);
            yt = (const T *)tb.yc;
        }
        GPG_TRY(nll_grad_core<T>(h, kernel_id, d, theta, X, yt, N, jitter, tb, (T *)tb.nll, (T *)tb.grad, info, 0, s));
        if (mode == 1) {
            sk_grad_fix_kernel<T><<<1, 256, 0, s>>>((const T *)tb.alpha, N, 3 + d, (T *)tb.grad, (T *)tb.nll);
            GPG_LAUNCH_CHECK(h);
        }
        adam_step_kernel<T><<<1, 32, 0, s>>>(1, c, u, tb.st, (const T *)tb.grad, (const T *)tb.nll, theta, traj);
        GPG_LAUNCH_CHECK(h);
        return GPG_OK;
    };
    // Small problems are launch-bound (~35 dependent kernels of a few microseconds per iteration: the Bayesian
    // optimisation regime, N ~ 10..500, 1000 iterations x 50 trainings): after one eager iteration the second is
    // captured into a CUDA graph and replayed.  Every launch of an iteration is shape-static; the only
    // per-iteration state (Adam moments, step count, trajectory row) lives on the device.
    const bool use_graph = iters >= 8 && !h->opt_stage_timing && !train_uses_tc<T>(h, N, ld) && h->opt_fit_graph;
    int it = 0;
    if (use_graph) {
        GPG_TRY(one_iteration());
        it = 1;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const long long launches_before = h->launches;
        bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        int rc = GPG_OK;
        if (ok) {
            rc = one_iteration();
            ok = (cudaStreamEndCapture(s, &graph) == cudaSuccess) && rc == GPG_OK && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        if (ok) {
            const long long per_iter = h->launches - launches_before;
            for (; it < iters; ++it) {
                if (cudaGraphLaunch(exec, s) != cudaSuccess) { ok = false; break; }
                if (it > 1) h->launches += per_iter;       // the captured pass counted itself once
            }
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {
            cudaGetLastError();
            if (rc != GPG_OK) return rc;
            // capture refused (e.g. the caller's stream is already being captured): fall through to eager launches
        }
    }
    for (; it < iters; ++it) GPG_TRY(one_iteration());
    if (theta_out) GPG_CUDA_CHECK(cudaMemcpyAsync(theta_out, theta, (3 + d) * sizeof(T), cudaMemcpyDeviceToDevice, s));
    return GPG_OK;
}

extern "C" int gpg_fit_adam(gpg_handle_t h, int dtype, int kernel_id, int d, int n_ls, const void *X, const void *y,
                            int64_t N, double jitter, void *u, const double *bounds_host, int iters, double lr,
                            void *traj_out, void *theta_out, int32_t *info, void *stream) {
    GPG_REQUIRE(h && X && y && u && bounds_host && info, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(N > 0 && iters >= 0, "bad size");
    GPG_REQUIRE(d >= 1 && d <= GPG_MAX_D, "d not in 1..4");
    GPG_REQUIRE(n_ls == 1 || n_ls == d, "n_ls must be 1 or d");
    GPG_REQUIRE(kernel_id >= 0 && kernel_id <= 2, "unknown kernel id");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return fit_entry<float>(h, kernel_id, d, n_ls, (const float *)X, (const float *)y, N, jitter, (float *)u,
                                bounds_host, iters, lr, (float *)traj_out, (float *)theta_out, info, s);
    if (dtype == GPG_F64)
        return fit_entry<double>(h, kernel_id, d, n_ls, (const double *)X, (const double *)y, N, jitter, (double *)u,
                                 bounds_host, iters, lr, (double *)traj_out, (double *)theta_out, info, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

extern "C" int gpg_fit_adam_sk(gpg_handle_t h, int dtype, int kernel_id, int d, int n_ls, const void *X, const void *y,
                               int64_t N, double jitter, void *u, const double *bounds_host, int iters, double lr,
                               void *traj_out, void *theta_out, int32_t *info, void *stream) {
    GPG_REQUIRE(h && X && y && u && bounds_host && info, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(N > 0 && iters >= 0, "bad size");
    GPG_REQUIRE(d >= 1 && d <= GPG_MAX_D, "d not in 1..4");
    GPG_REQUIRE(n_ls == 1 || n_ls == d, "n_ls must be 1 or d");
    GPG_REQUIRE(kernel_id == GPG_RBF || kernel_id == GPG_MATERN52, "gpytorch kernel book: RBF or Matern52");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return fit_entry<float>(h, kernel_id, d, n_ls, (const float *)X, (const float *)y, N, jitter, (float *)u,
                                bounds_host, iters, lr, (float *)traj_out, (float *)theta_out, info, s, 1);
    if (dtype == GPG_F64)
        return fit_entry<double>(h, kernel_id, d, n_ls, (const double *)X, (const double *)y, N, jitter, (double *)u,
                                 bounds_host, iters, lr, (double *)traj_out, (double *)theta_out, info, s, 1);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

// vreconstructor(independent=True): T exact GPs on one X, shared lengthscale (train.cuh, MtCfg)
template <typename T>
static int fit_mt_entry(gpg_handle_s *h, int kernel_id, int d, int n_ls, int ntasks, const T *X, const T *Y, int64_t N,
                        double jitter, T *u, const double *bounds, int iters, double lr, T *traj, T *theta_out,
                        int32_t *info, cudaStream_t s) {
    const int64_t ld = gpg_align_up((size_t)N, 64);
    const size_t base = gpg_align_up(train_ws_bytes<T>(h, N, ld), 1024);
    const size_t extra = bump_size({(size_t)ntasks * GPG_MAX_P * sizeof(T), (size_t)ntasks * GPG_MAX_P * sizeof(T),
                                    (size_t)ntasks * sizeof(T), (size_t)ntasks * sizeof(double), sizeof(MtState)});
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, base + extra, &ws));
    TrainBufs tb = train_carve<T>(h, ws, N, ld);
    Bump b((unsigned char *)ws + base);
    T *theta_all = b.take<T>((size_t)ntasks * GPG_MAX_P);
    T *grad_all = b.take<T>((size_t)ntasks * GPG_MAX_P);
    T *nll_all = b.take<T>(ntasks);
    double *asum = b.take<double>(ntasks);
    MtState *st = b.take<MtState>(1);
    MtCfg c;
    memset(&c, 0, sizeof(c));
    c.d = d; c.n_ls = n_ls; c.T = ntasks; c.ls_softplus = (bounds == nullptr);
    if (bounds) for (int k = 0; k < n_ls; ++k) { c.ls_lo[k] = bounds[k]; c.ls_hi[k] = bounds[n_ls + k]; }
    c.lr = lr; c.beta1 = 0.9; c.beta2 = 0.999; c.eps = 1e-8;
    mt_adam_kernel<T><<<1, 32, 0, s>>>(0, c, u, st, nullptr, nullptr, nullptr, (double)N, theta_all, nullptr);
    GPG_LAUNCH_CHECK(h);
    GPG_CUDA_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), s));
    for (int it = 0; it < iters; ++it) {
        for (int t = 0; t < ntasks; ++t) {
            const T *theta = theta_all + (size_t)t * GPG_MAX_P;
            center_y_kernel<T><<<(unsigned)((N + 255) / 256), 256, 0, s>>>(Y + (size_t)t * N, N, theta, (T *)tb.yc);
            GPG_LAUNCH_CHECK(h);
            GPG_TRY(nll_grad_core<T>(h, kernel_id, d, theta, X, (const T *)tb.yc, N, jitter, tb, nll_all + t,
                                     grad_all + (size_t)t * GPG_MAX_P, info, 0, s));
            mt_alpha_sum_kernel<T><<<1, 256, 0, s>>>((const T *)tb.alpha, N, asum + t);
            GPG_LAUNCH_CHECK(h);
        }
        mt_adam_kernel<T><<<1, 32, 0, s>>>(1, c, u, st, grad_all, nll_all, asum, (double)N, theta_all, traj);
        GPG_LAUNCH_CHECK(h);
    }
    if (theta_out)
        GPG_CUDA_CHECK(cudaMemcpy2DAsync(theta_out, (3 + d) * sizeof(T), theta_all, GPG_MAX_P * sizeof(T), (3 + d) * sizeof(T),
                                         ntasks, cudaMemcpyDeviceToDevice, s));
    return GPG_OK;
}

extern "C" int gpg_fit_adam_mt(gpg_handle_t h, int dtype, int kernel_id, int d, int n_ls, int ntasks, const void *X,
                               const void *Y, int64_t N, double jitter, void *u, const double *bounds_host, int iters,
                               double lr, void *traj_out, void *theta_out, int32_t *info, void *stream) {
    GPG_REQUIRE(h && X && Y && u && info, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(N > 0 && iters >= 0, "bad size");
    GPG_REQUIRE(d >= 1 && d <= GPG_MAX_D, "d not in 1..4");
    GPG_REQUIRE(n_ls == 1 || n_ls == d, "n_ls must be 1 or d");
    GPG_REQUIRE(ntasks >= 1 && ntasks <= GPG_MT_MAX_TASKS, "number of tasks not in 1..16");
    GPG_REQUIRE(kernel_id == GPG_RBF || kernel_id == GPG_MATERN52, "gpytorch kernel book: RBF or Matern52");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return fit_mt_entry<float>(h, kernel_id, d, n_ls, ntasks, (const float *)X, (const float *)Y, N, jitter, (float *)u,
                                   bounds_host, iters, lr, (float *)traj_out, (float *)theta_out, info, s);
    if (dtype == GPG_F64)
        return fit_mt_entry<double>(h, kernel_id, d, n_ls, ntasks, (const double *)X, (const double *)Y, N, jitter,
                                    (double *)u, bounds_host, iters, lr, (double *)traj_out, (double *)theta_out, info, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

// ---------------------------------------------------------------------------------------------
// K6: acquisition sweep + top-k
// ---------------------------------------------------------------------------------------------
// Sweep + tournament: leaves the k best candidates (ranked, idx < 0 = none) in workspace memory at *best_out.
// idx_offset is added to every flat index (a rank's tile of a sharded grid reports global indices);
// extra_cands: room reserved behind the tournament buffers for the caller (sharded merge).
template <typename T>
static int acq_local_topk(gpg_handle_s *h, int acq_id, const T *mean, const T *sd, const T *mask, int64_t M,
                          int64_t idx_offset, double mu_best, double xi, double alpha, double beta, int k, T *acq_out,
                          size_t extra_cands, Cand<T> **best_out, Cand<T> **extra_out, cudaStream_t s) {
    constexpr int CH = 2048;
    const int64_t nblk0 = std::max<int64_t>(1, (M + CH - 1) / CH);
    // large grids: a two-level key histogram keeps the k best plus one fine bin's worth of candidates out of M before
    // anything is sorted (acq.cuh); the tournament then runs on the survivors with its counts on the device
    const bool prefilter = M >= 65536 && M < ((int64_t)1 << 31);
    const size_t aux_bytes = (2 * TOPK_BINS + 8) * sizeof(unsigned);
    void *ws;
    GPG_TRY(gpg_ws_reserve(h, bump_size({(size_t)(M + 1) * sizeof(Cand<T>), (size_t)(nblk0 * k + CH) * sizeof(Cand<T>),
                                         (size_t)(nblk0 * k + CH) * sizeof(Cand<T>), 2 * (extra_cands + CH) * sizeof(Cand<T>),
                                         prefilter ? (size_t)(M + 1) * sizeof(Cand<T>) : 0, aux_bytes}),
                           &ws));
    Bump b(ws);
    Cand<T> *cand = b.take<Cand<T>>(M + 1);
    Cand<T> *bufA = b.take<Cand<T>>(nblk0 * k + CH);
    Cand<T> *bufB = b.take<Cand<T>>(nblk0 * k + CH);
    if (extra_out) *extra_out = b.take<Cand<T>>(2 * (extra_cands + CH));
    Cand<T> *surv = prefilter ? b.take<Cand<T>>(M + 1) : nullptr;
    unsigned *hist1 = b.take<unsigned>(2 * TOPK_BINS + 8), *hist2 = hist1 + TOPK_BINS;
    int *sel = reinterpret_cast<int *>(hist2 + TOPK_BINS), *ncnt = sel + 4;
    const unsigned gsweep = (unsigned)std::min<int64_t>((M + 255) / 256, (int64_t)h->sm_count * 8);
    if (prefilter) GPG_CUDA_CHECK(cudaMemsetAsync(hist1, 0, aux_bytes, s));
    if (M > 0) {                         // an empty tile (sharded sweep) contributes k excluded entries
        acq_eval_kernel<T><<<std::max(1u, gsweep), 256, 0, s>>>(acq_id, mean, sd, mask, M, idx_offset, mu_best, xi,
                                                                 alpha, beta, acq_out, cand, prefilter ? hist1 : nullptr);
        GPG_LAUNCH_CHECK(h);
    }
    const Cand<T> *in = cand;
    int64_t n = M;
    if (prefilter) {
        topk_level1_kernel<<<1, 1024, 0, s>>>(hist1, k, sel);
        GPG_LAUNCH_CHECK(h);
        topk_hist2_kernel<T><<<gsweep, 256, 0, s>>>(cand, M, sel, hist2);
        GPG_LAUNCH_CHECK(h);
        topk_level2_kernel<<<1, 1024, 0, s>>>(hist2, k, sel, ncnt);
        GPG_LAUNCH_CHECK(h);
        topk_compact_kernel<T><<<gsweep, 256, 0, s>>>(cand, M, sel, surv, ncnt);
        GPG_LAUNCH_CHECK(h);
        in = surv;
    }
    Cand<T> *out = bufA;
    int round = 0;
    while (true) {
        const int64_t nblk = std::max<int64_t>(1, (n + CH - 1) / CH);
        // with the pre-filter the grid is sized for the worst case (nothing filtered out: all values equal) and the
        // blocks past the survivors return at once
        topk_round_kernel<T><<<(unsigned)nblk, 1024, CH * sizeof(Cand<T>), s>>>(in, n, k, out, prefilter ? ncnt + (round & 1) : nullptr,
                                                                                prefilter ? ncnt + ((round + 1) & 1) : nullptr);
        GPG_LAUNCH_CHECK(h);
        if (nblk == 1) break;
        in = out;
        n = nblk * k;
        out = (out == bufA) ? bufB : bufA;
        ++round;
    }
    *best_out = out;
    return GPG_OK;
}

template <typename T>
static int acq_entry(gpg_handle_s *h, int acq_id, const T *mean, const T *sd, const T *mask, int64_t M, double mu_best,
                     double xi, double alpha, double beta, int k, T *topk_val, int64_t *topk_idx, int32_t *count,
                     T *acq_out, cudaStream_t s) {
    StageTimer st(h, GPG_ST_ACQ, s);
    Cand<T> *best;
    GPG_TRY(acq_local_topk<T>(h, acq_id, mean, sd, mask, M, 0, mu_best, xi, alpha, beta, k, acq_out, 0, &best, nullptr, s));
    topk_emit_kernel<T><<<1, 256, 0, s>>>(best, k, topk_val, topk_idx, count);
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

extern "C" int gpg_acq_sweep(gpg_handle_t h, int dtype, int acq_id, const void *mean, const void *sd, const void *mask,
                             int64_t M, double mu_best, double xi, double alpha, double beta, int k, void *topk_val,
                             int64_t *topk_idx, int32_t *count_out, void *acq_out, void *stream) {
    GPG_REQUIRE(h && mean && sd && topk_val && topk_idx && count_out, "NULL argument");
    DeviceGuard device_guard(h->device);
    GPG_REQUIRE(M > 0, "M must be positive");
    GPG_REQUIRE(k >= 1 && k <= 1024, "k must be in 1..1024");
    GPG_REQUIRE(acq_id >= 0 && acq_id <= 2, "unknown acquisition id");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == GPG_F32)
        return acq_entry<float>(h, acq_id, (const float *)mean, (const float *)sd, (const float *)mask, M, mu_best, xi,
                                alpha, beta, k, (float *)topk_val, topk_idx, count_out, (float *)acq_out, s);
    if (dtype == GPG_F64)
        return acq_entry<double>(h, acq_id, (const double *)mean, (const double *)sd, (const double *)mask, M, mu_best, xi,
                                 alpha, beta, k, (double *)topk_val, topk_idx, count_out, (double *)acq_out, s);
    gpg_set_error("unknown dtype %d", dtype);
    return GPG_EINVAL;
}

extern "C" int gpg_acq_select(gpg_handle_t h, int dtype, const void *topk_val, const int64_t *topk_idx,
                              const int32_t *count, int k, int ndim, const int64_t *dims_host, const int64_t *visited,
                              int n_visited, int memory, double dscale, double gamma, int do_batch, double batch_dscale,
                              int batch_out_max, int32_t *sel_out, void *stream) {
    GPG_REQUIRE(h && topk_val && topk_idx && count && dims_host && sel_out, "NULL argument");
    GPG_REQUIRE(k >= 1 && k <= 1024, "k must be in 1..1024");
    GPG_REQUIRE(ndim >= 1 && ndim <= 4, "ndim not in 1..4");
    GPG_REQUIRE(n_visited >= 0 && (n_visited == 0 || visited != nullptr), "visited list missing");
    GPG_REQUIRE(batch_out_max >= 0 && batch_out_max <= 1024 && memory >= 0, "bad batch_out_max / memory");
    DeviceGuard device_guard(h->device);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    SelectArgs a;
    a.ndim = ndim;
    for (int q = 0; q < 4; ++q) a.dims[q] = q < ndim ? dims_host[q] : 1;
    a.n_visited = n_visited; a.memory = memory; a.do_batch = do_batch; a.batch_out_max = batch_out_max;
    a.dscale = dscale; a.gamma = gamma; a.batch_dscale = batch_dscale;
    if (dtype == GPG_F32)
        acq_select_kernel<float><<<1, 1024, 0, s>>>((const float *)topk_val, topk_idx, count, k, visited, a, sel_out);
    else if (dtype == GPG_F64)
        acq_select_kernel<double><<<1, 1024, 0, s>>>((const double *)topk_val, topk_idx, count, k, visited, a, sel_out);
    else { gpg_set_error("unknown dtype %d", dtype); return GPG_EINVAL; }
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

// ---------------------------------------------------------------------------------------------
// Inducing-point GP (sparse=True): drivers and gpg_sparse_* entry points.  Same translation unit: they use the
// workspace carving, kmat_launch and the factorisation drivers defined above.
// ---------------------------------------------------------------------------------------------
#include "sparse_driver.cuh"
#include "comm_driver.cuh"
