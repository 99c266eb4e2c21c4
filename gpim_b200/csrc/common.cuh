// common.cuh -- handle, error plumbing and the covariance-function math shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <string>
#include "../../include/gpgrid.h"

#define GPG_MAX_D 4

#include <vector>

struct gpg_stage_span { int stage; cudaEvent_t beg, end; };

struct gpg_handle_s {
    int device = 0;
    int sm_count = 148;
    long long launches = 0;
    int opt_gemm_path = 0;
    long long opt_predict_chunk = 0;
    int opt_stage_timing = 0;
    int opt_factor_algo = 0;
    int opt_fit_graph = 1;
    int opt_panel_mode = 3;
    int opt_outer_panel = 512;
    int opt_compact_support = 1;
    int opt_inner_left = 1;
    int opt_panel_refine = 1;
    int opt_syrk_chunk = 0;
    int opt_lookahead = 1;
    int opt_panel_workers = 0;
    void *ws = nullptr;          // grow-only device workspace
    size_t ws_bytes = 0;
    double *gemv_part = nullptr;             // partial sums of the transposed triangular GEMV (grow-only)
    size_t gemv_part_elems = 0;
    int *tc_counters = nullptr;              // pool of zeroed tile counters for the persistent GEMM
    int tc_counter_pos = 0;
    int *tc_counters_side = nullptr;         // the same for launches on side_stream
    int tc_counter_pos_side = 0;
    cudaStream_t fit_stream = nullptr;       // blocking stream the small-N Adam loop is captured on
    cudaStream_t side_stream = nullptr;      // look-ahead of the blocked Cholesky: trailing updates beyond the next panel
    cudaEvent_t ev_fork = nullptr, ev_side = nullptr;
    unsigned long long *work_counter = nullptr;   // device: k-blocks executed by the variance GEMM while stage timing is on
    void *comm = nullptr;                    // comm::State (comm.cuh): NCCL communicator + communication stream
    std::vector<gpg_stage_span> spans;       // recorded while opt_stage_timing != 0
    std::vector<cudaEvent_t> event_pool;
};

// CUDA-event bracket around the launches of one stage (bench.py roofline: average device time of
// the dominant kernel measured on the launching stream).  No-ops unless GPG_OPT_STAGE_TIMING is set.
struct StageTimer {
    gpg_handle_s *h;
    cudaStream_t s;
    cudaEvent_t end = nullptr;
    StageTimer(gpg_handle_s *h_, int stage, cudaStream_t s_) : h(h_), s(s_) {
        if (!h->opt_stage_timing) return;
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; ++i) {
            if (!h->event_pool.empty()) { ev[i] = h->event_pool.back(); h->event_pool.pop_back(); }
            else if (cudaEventCreate(&ev[i]) != cudaSuccess) return;
        }
        cudaEventRecord(ev[0], s);
        end = ev[1];
        h->spans.push_back({stage, ev[0], ev[1]});
    }
    ~StageTimer() { if (end) cudaEventRecord(end, s); }
};

void gpg_set_error(const char *fmt, ...);
// next zeroed tile counter of the handle's pool (re-zeroed stream-ordered when it wraps)
int gpg_tc_counter(gpg_handle_s *h, cudaStream_t stream, int **out);
// n consecutive zeroed ints of the same pool (flags of the cooperative panel kernel)
int gpg_tc_counters(gpg_handle_s *h, cudaStream_t stream, int n, int **out);
// returns pointer into the handle workspace, growing it if needed (synchronises on growth)
int gpg_ws_reserve(gpg_handle_s *h, size_t bytes, void **out);
// scratch for gemv_tri_T: at least `elems` doubles (synchronises on growth)
int gpg_gemv_part_reserve(gpg_handle_s *h, size_t elems, double **out);

#define GPG_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            gpg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return GPG_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

#define GPG_LAUNCH_CHECK(h)                                                               \
    do {                                                                                  \
        (h)->launches++;                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            gpg_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return GPG_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

#define GPG_REQUIRE(cond, msg)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            gpg_set_error("%s:%d invalid argument: %s", __FILE__, __LINE__, msg);         \
            return GPG_EINVAL;                                                            \
        }                                                                                 \
    } while (0)

#define GPG_TRY(expr)                                                                     \
    do {                                                                                  \
        int _rc = (expr);                                                                 \
        if (_rc != GPG_OK) return _rc;                                                    \
    } while (0)

static inline size_t gpg_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Every entry point runs on the handle's device, whatever the caller's current device is (restored on exit).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = (cudaSetDevice(device) == cudaSuccess);
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// Programmatic dependent launch (the factorisation is a chain of short dependent kernels): a kernel calls
// pdl_trigger() on entry, so that the NEXT kernel of the stream -- if it was launched through launch_pdl() --
// may start its prologue (barrier init, TMEM allocation, descriptor prefetch) right away; that kernel calls
// pdl_wait() before it touches anything the previous grids produced (returns once they have completed and
// their writes are visible).  Both are no-ops for plainly launched kernels.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------------------------------------
// Covariance functions.  r2 is the squared lengthscale-scaled distance computed by DIRECT
// DIFFERENCE (the reference expands X2 - 2XZ^T + Z2, pyro Isotropy._square_scaled_dist; the
// direct form is what keeps fp32 within tolerance).
// ---------------------------------------------------------------------------------------------
template <typename T> struct Theta {
    T variance, noise, alpha;
    T inv_ls[GPG_MAX_D];
};

template <typename T, int D>
__device__ __forceinline__ Theta<T> load_theta(const T *__restrict__ theta) {
    Theta<T> t;
    t.variance = theta[0];
    t.noise = theta[1];
    t.alpha = theta[2];
#pragma unroll
    for (int k = 0; k < D; ++k) t.inv_ls[k] = T(1) / theta[3 + k];
    return t;
}

__device__ __forceinline__ float gpg_exp(float x) { return expf(x); }
__device__ __forceinline__ double gpg_exp(double x) { return exp(x); }
__device__ __forceinline__ float gpg_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double gpg_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float gpg_log(float x) { return logf(x); }
__device__ __forceinline__ double gpg_log(double x) { return log(x); }
__device__ __forceinline__ float gpg_pow(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double gpg_pow(double x, double y) { return pow(x, y); }

// k(r2) for kernel KID.  Matern52 follows pyro's _torch_sqrt(r2 + 1e-12).
template <typename T, int KID>
__device__ __forceinline__ T cov_from_r2(T r2, const Theta<T> &t) {
    if (KID == GPG_RBF) {
        return t.variance * gpg_exp(T(-0.5) * r2);
    } else if (KID == GPG_MATERN52) {
        T r = gpg_sqrt(r2 + T(1e-12));
        T s = T(2.23606797749978969641) * r;
        return t.variance * (T(1) + s + (T(5) / T(3)) * r * r) * gpg_exp(-s);
    } else {
        T base = T(1) + (T(0.5) / t.alpha) * r2;
        return t.variance * gpg_pow(base, -t.alpha);
    }
}

template <typename T, int D>
__device__ __forceinline__ T scaled_r2(const T *__restrict__ x, const T *__restrict__ z, const Theta<T> &t) {
    T r2 = T(0);
#pragma unroll
    for (int k = 0; k < D; ++k) {
        T dlt = (x[k] - z[k]) * t.inv_ls[k];
        r2 += dlt * dlt;
    }
    return r2;
}

// dispatch helpers ------------------------------------------------------------------------
#define GPG_DISPATCH_D(d, ...)                                       \
    switch (d) {                                                     \
        case 1: { constexpr int D = 1; __VA_ARGS__; } break;         \
        case 2: { constexpr int D = 2; __VA_ARGS__; } break;         \
        case 3: { constexpr int D = 3; __VA_ARGS__; } break;         \
        case 4: { constexpr int D = 4; __VA_ARGS__; } break;         \
        default: gpg_set_error("input dimension %d not in 1..4", d); return GPG_EINVAL; \
    }

#define GPG_DISPATCH_KID(kid, ...)                                                  \
    switch (kid) {                                                                  \
        case GPG_RBF: { constexpr int KID = GPG_RBF; __VA_ARGS__; } break;          \
        case GPG_MATERN52: { constexpr int KID = GPG_MATERN52; __VA_ARGS__; } break;\
        case GPG_RATQUAD: { constexpr int KID = GPG_RATQUAD; __VA_ARGS__; } break;  \
        default: gpg_set_error("unknown kernel id %d", kid); return GPG_EINVAL;     \
    }

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
