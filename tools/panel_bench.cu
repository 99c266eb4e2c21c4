// Phase timing and correctness of the cooperative Cholesky panel kernel (development aid).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DGPG_PANEL_PROFILE -DGPG_DIAG_PROFILE -o /tmp/panel_bench tools/panel_bench.cu -lcuda
// usage: panel_bench [N = 2048]    (first 512-column panel of a random SPD matrix)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include "../gpim_b200/csrc/chol_panel.cuh"
void gpg_set_error(const char *fmt, ...) { fprintf(stderr, "error: %s\n", fmt); }
int gpg_tc_counter(gpg_handle_s *, cudaStream_t, int **) { return 0; }
int gpg_tc_counters(gpg_handle_s *, cudaStream_t, int, int **) { return 0; }
int gpg_ws_reserve(gpg_handle_s *, size_t, void **) { return 0; }
int gpg_gemv_part_reserve(gpg_handle_s *, size_t, double **) { return 0; }
template <> int gemm_dispatch<float>(gpg_handle_s *, const GemmArgs<float> &, cudaStream_t) { return 0; }
template <> int gemm_dispatch<double>(gpg_handle_s *, const GemmArgs<double> &, cudaStream_t) { return 0; }

int main(int argc, char **argv) {
    const long long N = argc > 1 ? atoll(argv[1]) : 2048, ld = (N + 63) / 64 * 64;
    const int nbp = (int)std::min<long long>(4, (N + 127) / 128);
    if (N < 128 * nbp && N % 128 == 0) return 1;
    std::vector<float> A((size_t)N * ld, 0.f);
    // K = 0.5 exp(-d^2 / 2 l^2) on a line + nugget: SPD, well conditioned
    for (long long i = 0; i < N; ++i)
        for (long long j = 0; j <= i; ++j) {
            const double d = (double)(i - j) / 6.0;
            A[i * ld + j] = A[j * ld + i] = (float)(0.5 * exp(-0.5 * d * d) + (i == j ? 0.05 : 0.0));
        }
    // reference: double Cholesky of the leading 128 nbp columns (panel only)
    std::vector<double> Lr((size_t)N * 128 * nbp, 0.0);
    const int W = 128 * nbp;
    for (int j = 0; j < W && j < N; ++j) {
        double d = A[(size_t)j * ld + j];
        for (int k = 0; k < j; ++k) d -= Lr[(size_t)j * W + k] * Lr[(size_t)j * W + k];
        const double ljj = sqrt(d);
        Lr[(size_t)j * W + j] = ljj;
        for (long long i = j + 1; i < N; ++i) {
            double s = A[(size_t)i * ld + j];
            for (int k = 0; k < j; ++k) s -= Lr[(size_t)i * W + k] * Lr[(size_t)j * W + k];
            Lr[(size_t)i * W + j] = s / ljj;
        }
    }
    float *dA, *dsc; __half *pl; int *flags, *info;
    cudaMalloc(&dA, (size_t)N * ld * 4); cudaMalloc(&pl, (size_t)4 * N * ld * 2); cudaMalloc(&flags, 4096); cudaMalloc(&info, 4);
    cudaMalloc(&dsc, 64 * 4);
    float sc[16] = {0};
    // scales as scales_from_theta_kernel would choose them for v = 0.5, nz = 0.05
    const float v = 0.5f, nz = 0.05f;
    auto p2 = [](float x) { return exp2f(floorf(log2f(x))); };
    sc[1] = p2(16384.f * sqrtf(nz)); sc[3] = p2(16384.f / sqrtf(v + nz)); sc[9] = p2(16384.f / (v + nz));
    sc[4] = 1.f / (sc[3] * sc[3]); sc[10] = 1.f / (sc[9] * sc[1]);
    cudaMemcpy(dsc, sc, sizeof(sc), cudaMemcpyHostToDevice);
    cpanel::Args pa;
    pa.A = dA; pa.ld = ld; pa.N = N; pa.J0 = 0; pa.nbp = nbp;
    pa.Ls_hi = pl; pa.Ls_lo = pl + (size_t)N * ld; pa.Ws_hi = pl + (size_t)2 * N * ld; pa.Ws_lo = pl + (size_t)3 * N * ld;
    pa.scales = dsc; pa.sc_A = 9; pa.sc_W = 1; pa.sc_L = 3; pa.sc_inv_AW = 10; pa.sc_inv_LL = 4;
    pa.flags = flags; pa.info = info;
    CUtensorMap mLhi, mLlo, mWhi, mWlo;
    if (tc::make_tensor_map(&mLhi, pa.Ls_hi, N, N, ld, 128) || tc::make_tensor_map(&mLlo, pa.Ls_lo, N, N, ld, 128) ||
        tc::make_tensor_map(&mWhi, pa.Ws_hi, N, N, ld, 128) || tc::make_tensor_map(&mWlo, pa.Ws_lo, N, N, ld, 128)) return 1;
    cudaFuncSetAttribute(cpanel::chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cpanel::SMEM_BYTES);
    const int rows_blk = (int)((N + 127) / 128), grid = std::min(rows_blk, 148);
    void *kargs[] = {&mLhi, &mLlo, &mWhi, &mWlo, &pa};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(dA, A.data(), (size_t)N * ld * 4, cudaMemcpyHostToDevice);
        cudaMemset(flags, 0, 4096); cudaMemset(info, 0, 4);
        cudaEventRecord(e0);
        cudaError_t err = cudaLaunchCooperativeKernel((const void *)cpanel::chol_panel_kernel, dim3(grid), dim3(cpanel::NUM_THREADS), kargs,
                                                      (size_t)cpanel::SMEM_BYTES, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rep %d: N=%lld grid=%d  %.1f us  (%s / %s)\n", rep, N, grid, ms * 1e3, cudaGetErrorString(err), cudaGetErrorString(cudaGetLastError()));
    }
#ifdef GPG_PANEL_PROFILE
    long long clk[512]; cudaMemcpyFromSymbol(clk, cpanel::g_panel_clk, sizeof(clk));
    const char *cn[7] = {"", "wait diag_ready", "load", "factor + inverse + planes", "-", "sync + flag", "L out"};
    for (int j = 0; j < nbp; ++j) {
        printf("chain block %d (cycles):", j);
        for (int k = 1; k <= 6; ++k) printf("  %s %lld", cn[k], clk[16 * j + k] - clk[16 * j + k - 1]);
        printf("   | globaltimer ns: wait %lld  total %lld\n", clk[256 + 16 * j + 1] - clk[256 + 16 * j], clk[256 + 16 * j + 6] - clk[256 + 16 * j]);
    }
    const char *wn[8] = {"", "X built", "diag_done seen", "W landed (TMA)", "L = X W^T done", "planes in smem + diag update", "diag block out + flag", "L to global + flag"};
    for (int rb = 1; rb < nbp; ++rb) {
        printf("worker of row block %d, step %d (cycles):", rb, rb - 1);
        for (int k = 1; k <= 7; ++k) printf("  %s %lld", wn[k], clk[128 + 16 * rb + k] - clk[128 + 16 * rb + k - 1]);
        printf("\n    globaltimer ns: chain flag(diag_done %d) -> worker saw it %lld;  worker flags -> chain saw diag_ready %lld;  worker critical %lld\n",
               rb - 1, clk[256 + 128 + 16 * rb + 2] - clk[256 + 16 * (rb - 1) + 5], clk[256 + 16 * rb + 1] - clk[256 + 128 + 16 * rb + 6],
               clk[256 + 128 + 16 * rb + 6] - clk[256 + 128 + 16 * rb + 2]);
    }
#endif
#ifdef GPG_PANEL_PROFILE
    {   // inside factor_invert_block of the LAST diagonal block the chain handled (thread 0 = lane 0 of the pivot warp)
        const long long *c = clk + 64;
        printf("last chain block, inside (cycles):");
        for (int p4 = 0; p4 < 4; ++p4) {
            printf("  p%d: pivot32 %lld, wait for the others %lld", p4, c[4 * p4 + 1] - c[4 * p4], c[4 * p4 + 2] - c[4 * p4 + 1]);
            if (p4 < 3) printf(", rows below %lld, next-panel update + sync %lld |", c[4 * p4 + 3] - c[4 * p4 + 2], c[4 * p4 + 4] - c[4 * p4 + 3]);
        }
        printf("\n   warp 6 (shadow jobs of sub-step k, cycles):");
        for (int k = 0; k < 3; ++k) printf("  k%d: starts %lld after the pivot, stage 1 %lld, group barrier %lld, stage 2 %lld, planes %lld |", k, clk[96 + 4 * k] - c[4 * (k + 1)],
                                           clk[96 + 4 * k + 1] - clk[96 + 4 * k], clk[96 + 4 * k + 2] - clk[96 + 4 * k + 1], clk[96 + 4 * k + 3] - clk[96 + 4 * k + 2], clk[96 + 12 + k] - clk[96 + 4 * k + 3]);
        printf("\n");
        printf("  | last sync %lld  tail product %lld  planes of the last 32 rows %lld\n", c[16] - c[14], c[17] - c[16], c[18] - c[17]);
    }
#endif
#ifdef GPG_DIAG_PROFILE_OLD
    {   // phases inside diag_factor_smem / diag_invert_smem of the LAST diagonal block the chain handled
        long long d[64]; cudaMemcpyFromSymbol(d, g_diag_clk, sizeof(d));
        printf("last chain block, inside: ");
        for (int p4 = 0; p4 < 4; ++p4) {
            printf(" f32[%d] %lld", p4, d[3 + 5 * p4] - d[2 + 5 * p4]);
            if (p4 < 3) printf(" solve %lld trail %lld |", d[5 + 5 * p4] - d[3 + 5 * p4], d[6 + 5 * p4] - d[5 + 5 * p4]);
        }
        printf("  | invert32+doubling32 %lld (from the last factor32)\n", d[23] - d[18]);
    }
#endif
    std::vector<float> L((size_t)N * ld);
    cudaMemcpy(L.data(), dA, (size_t)N * ld * 4, cudaMemcpyDeviceToHost);
    int hinfo; cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost);
    double err = 0, mx = 0;
    for (long long i = 0; i < N; ++i)
        for (int j = 0; j < W && j <= i; ++j) { const double e = fabs(L[i * ld + j] - Lr[(size_t)i * W + j]); err = (e > err || e != e) ? e : err; mx = fmax(mx, fabs(Lr[(size_t)i * W + j])); }
    printf("info %d   max |L - Lref| / max |Lref| over the panel = %.2e\n", hinfo, err / mx);
    return 0;
}
