// Phase timing of diag_block_kernel<float,128> (development aid).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DGPG_DIAG_PROFILE -o /tmp/diag_bench tools/diag_bench.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "../gpim_b200/csrc/factor.cuh"
void gpg_set_error(const char *, ...) {}
int gpg_tc_counter(gpg_handle_s *, cudaStream_t, int **) { return 0; }
int gpg_ws_reserve(gpg_handle_s *, size_t, void **) { return 0; }
int main() {
    const int N = 128, ld = 128;
    std::vector<float> A(N * ld);
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) A[i * ld + j] = 0.5f * expf(-0.02f * (i - j) * (i - j)) + (i == j ? 0.05f : 0.f);
    float *dA, *dW; __half *pl; int *info; float *sc;
    cudaMalloc(&dA, N * ld * 4); cudaMalloc(&dW, N * ld * 4); cudaMalloc(&pl, 6 * N * ld * 2); cudaMalloc(&info, 4); cudaMalloc(&sc, 8);
    float hs[2] = {1024.f, 64.f}; cudaMemcpy(sc, hs, 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(diag_block_kernel<float, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_block_smem<float, 128>());
    DiagEmit em; em.Lh = pl; em.Ll = pl + N * ld; em.Wh = pl + 2 * N * ld; em.Wl = pl + 3 * N * ld; em.WTh = pl + 4 * N * ld; em.WTl = pl + 5 * N * ld;
    em.lds = ld; em.scale_L = sc; em.scale_W = sc + 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(dA, A.data(), N * ld * 4, cudaMemcpyHostToDevice); cudaMemset(info, 0, 4);
        cudaEventRecord(e0);
        diag_block_kernel<float, 128><<<1, 256, diag_block_smem<float, 128>()>>>(dA, ld, N, 0, 1, dW, ld, 0, 0, info, em);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rep %d: %.1f us  (%s)\n", rep, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
#ifdef GPG_DIAG_PROFILE
    long long clk[64]; cudaMemcpyFromSymbol(clk, g_diag_clk, sizeof(clk));
    const char *names[26] = {"load", "sync", "p0 wait", "p0 factor32", "-", "p0 panel solve", "p0 trailing", "p1 wait", "p1 factor32", "-",
                             "p1 panel solve", "p1 trailing", "p2 wait", "p2 factor32", "-", "p2 panel solve", "p2 trailing", "p3 wait",
                             "p3 factor32", "-", "-", "-", "factor done", "invert + doubling 32", "doubling 64", "emit"};
    long long prev = clk[0];
    int order[] = {1, 2, 3, 5, 6, 7, 8, 10, 11, 12, 13, 15, 16, 17, 18, 22, 23, 24, 25};
    for (int i : order) { printf("  %-22s %8lld cyc\n", names[i], clk[i] - prev); prev = clk[i]; }
    printf("  total after load %lld cyc\n", clk[25] - clk[0]);
#endif
    std::vector<float> L(N * ld), W(N * ld);
    cudaMemcpy(L.data(), dA, N * ld * 4, cudaMemcpyDeviceToHost); cudaMemcpy(W.data(), dW, N * ld * 4, cudaMemcpyDeviceToHost);
    double err = 0;  // || W L - I ||
    for (int i = 0; i < N; ++i) for (int j = 0; j <= i; ++j) { double s = 0; for (int k = j; k <= i; ++k) s += (double)W[i * ld + k] * L[k * ld + j]; err = fmax(err, fabs(s - (i == j))); }
    printf("max |W L - I| = %.2e\n", err);
    return 0;
}
