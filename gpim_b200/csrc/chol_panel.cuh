// chol_panel.cuh -- the column panel of the blocked Cholesky as ONE cooperative kernel with a device-side
// dependency chain (replaces, per 512-column outer panel, 4 x {diag_block_kernel, panel GEMM, inner-update GEMM}).
//
// Panel = columns [J0, J0 + 128 nbp) of the trailing matrix, rows J0 .. N, in row blocks of 128 ("rb", rb = 0 is the
// block row of the panel's first diagonal block).  Roles:
//
//   CTA 0, the CHAIN: walks the nbp diagonal blocks.  Block j: (j > 0: wait until its owner has applied the updates of
//     columns < j) -> Cholesky + inverse of the 128 x 128 block in shared memory (diag_factor_smem / diag_invert_smem)
//     -> fp16 planes of the inverse W_jj to global -> flag diag_done[j] -> then, off the critical path, L_jj (fp32 +
//     planes).
//   CTAs 1.., the WORKERS: worker w owns row blocks rb = 1 + w, 1 + w + W, ...  For one row block the 128 x 512 strip of
//     update sums lives in TENSOR MEMORY (128 lanes = the rows, 512 columns: one 128-column scratch + three
//     accumulators U_c).  Per column block j:
//       X   = A[rb][j] - U_j                   TMEM -> registers, fp32 subtract, fp16 hi/lo split -> swizzled smem
//       L   = X W_jj^T                         tcgen05 (W_jj planes by TMA once diag_done[j] is up)
//       L  -> global fp32 (in place), global planes (operand of the outer SYRK / trtri), smem planes (A operand below)
//       U_c += L[rb][j] L[c][j]^T, c > j       tcgen05; L[c][j] planes by TMA once its owner has raised row_done[c][j]
//     A row block whose own diagonal block lies in the panel (rb < nbp) finishes with  A[rb][rb] - U_rb  -> global
//     -> flag diag_ready[rb]: that is what the chain waits for.
//
// Only the chain (128 pivots of the diagonal block + the hand-over with ONE worker) is serial; every other product of
// the panel runs under it on the other SMs.  Flags live in global memory (release / acquire at gpu scope); the launch
// is cooperative, so all CTAs are co-resident and the spins cannot starve their producers; every spin is bounded
// (~2 s) and raises an abort flag that lets all CTAs run to the end without waiting (info = -1).
#pragma once
#include "factor.cuh"
#include "gemm_tc.cuh"
#include "chol_chain.cuh"

namespace cpanel {

constexpr int NB = 128;
constexpr int KBLK = 32;                                   // halves per swizzled k-block (64-byte swizzle, as gemm_tc)
constexpr int TILE_BYTES = NB * KBLK * 2;                  // 8 KB: [128 rows][32 halves]
constexpr int PLANE_BYTES = (NB / KBLK) * TILE_BYTES;      // 32 KB: one plane of a 128 x 128 operand
constexpr int OPND_BYTES = 2 * PLANE_BYTES;                // hi + lo
constexpr int SMEM_WORKER = 3 * OPND_BYTES;                // XA, B0, B1
constexpr int SMEM_CHAIN = cchain::SMEM_BYTES;
constexpr int SMEM_BYTES = (SMEM_WORKER > SMEM_CHAIN ? SMEM_WORKER : SMEM_CHAIN) + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 256;
constexpr int NFLAGS = 32;
enum { F_DIAG_DONE = 0, F_DIAG_READY = 4, F_ROW_DONE = 8, F_ABORT = 24 };

struct Args {
    float *A; long long ld, N;
    long long J0; int nbp;                 // first column of the panel; number of 128-column blocks in it (1..4)
    __half *Ls_hi, *Ls_lo, *Ws_hi, *Ws_lo; // planes with the geometry of A
    const float *scales;                   // SC_* layout of factor_tc.cuh
    int sc_A, sc_W, sc_L, sc_inv_AW, sc_inv_LL;
    int *flags;                            // NFLAGS ints, zero on entry
    int32_t *info;
};

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// bounded spin; false = give up (abort raised by somebody, or by this call after ~2 s)
__device__ __noinline__ bool spin_until_set(const int *flag, int *abort_flag) {
    const long long start = clock64();
    int it = 0;
    while (ld_acquire(flag) == 0) {
        if ((++it & 63) == 0) {
            if (ld_acquire(abort_flag) != 0) return false;
            if (clock64() - start > (1LL << 32)) { atomicExch(abort_flag, 1); return false; }
        }
        if (it > 4096) __nanosleep(64);                   // poll tightly at first: the hand-overs of the chain are short
    }
    return true;
}

// byte offset of the 16-byte chunk holding columns [c8, c8 + 8) of row r inside one plane of a 128 x 128 operand
// (k-blocks of 32 halves, rows of 64 bytes, 64-byte swizzle: chunk index XOR bits 7..8 of the address)
__device__ __forceinline__ uint32_t opnd_offset(int r, int c8) {
    const int kb = c8 >> 5, chunk = (c8 & 31) >> 3;
    return (uint32_t)(kb * TILE_BYTES + r * 64 + ((chunk ^ ((r >> 1) & 3)) << 4));
}

// D[128 x 128] (+)= A B^T for split operands staged as {hi plane, lo plane} of 4 swizzled k-blocks each
__device__ __forceinline__ void issue_product(uint32_t d_tmem, uint32_t a_base, uint32_t b_base, uint32_t accumulate) {
    constexpr uint32_t idesc = tc::make_idesc(NB, NB);
    uint32_t acc = accumulate;
#pragma unroll
    for (int kb = 0; kb < NB / KBLK; ++kb) {
        const uint64_t dAhi = tc::make_smem_desc(a_base + kb * TILE_BYTES);
        const uint64_t dAlo = tc::make_smem_desc(a_base + PLANE_BYTES + kb * TILE_BYTES);
        const uint64_t dBhi = tc::make_smem_desc(b_base + kb * TILE_BYTES);
        const uint64_t dBlo = tc::make_smem_desc(b_base + PLANE_BYTES + kb * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < KBLK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            tc::tcgen05_mma_f16(d_tmem, dAhi + adv, dBhi + adv, idesc, acc);
            acc = 1;
            tc::tcgen05_mma_f16(d_tmem, dAhi + adv, dBlo + adv, idesc, 1);
            tc::tcgen05_mma_f16(d_tmem, dAlo + adv, dBhi + adv, idesc, 1);
        }
    }
}

// TMA: rows [row0, row0 + 128) x columns [col0, col0 + 128) of a plane pair -> operand buffer at dst
__device__ __forceinline__ void load_operand(uint32_t dst, const CUtensorMap *mhi, const CUtensorMap *mlo, uint32_t bar,
                                             int col0, int row0) {
    tc::mbar_arrive_expect_tx(bar, OPND_BYTES);
#pragma unroll
    for (int kb = 0; kb < NB / KBLK; ++kb) {
        tc::tma_load_2d(dst + kb * TILE_BYTES, mhi, bar, col0 + kb * KBLK, row0);
        tc::tma_load_2d(dst + PLANE_BYTES + kb * TILE_BYTES, mlo, bar, col0 + kb * KBLK, row0);
    }
}

__device__ __forceinline__ void split8(const float *v, float s, uint4 &hi, uint4 &lo) {
    __half2 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float a = v[2 * e] * s, b = v[2 * e + 1] * s;
        h[e] = __floats2half2_rn(a, b);
        const float2 f = __half22float2(h[e]);
        l[e] = __floats2half2_rn(a - f.x, b - f.y);
    }
    hi = make_uint4(*reinterpret_cast<unsigned *>(&h[0]), *reinterpret_cast<unsigned *>(&h[1]),
                    *reinterpret_cast<unsigned *>(&h[2]), *reinterpret_cast<unsigned *>(&h[3]));
    lo = make_uint4(*reinterpret_cast<unsigned *>(&l[0]), *reinterpret_cast<unsigned *>(&l[1]),
                    *reinterpret_cast<unsigned *>(&l[2]), *reinterpret_cast<unsigned *>(&l[3]));
}

// ---------------------------------------------------------------------------------------------
// the chain (CTA 0)
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void chain_role(const Args &p, unsigned char *smem, int *s_flag) {
    constexpr int LDS = cchain::LDS;
    float *S = reinterpret_cast<float *>(smem);
    float *W = S + NB * LDS;
    float *Q = W + NB * LDS;
    float *ring = Q + 96 * cchain::LDQ;
    const int t = threadIdx.x;
    int *abort_flag = p.flags + F_ABORT;
    cchain::Scales sc;
    sc.sA = p.scales[p.sc_A]; sc.sL = p.scales[p.sc_L]; sc.sW = p.scales[p.sc_W];
    sc.sQ = sc.sL * sc.sW * (1.0f / (16384.0f * 32.0f));        // |Q| = |L_rr W_rc| <= 32 max|L| max|W|
    sc.iAW = p.scales[p.sc_inv_AW]; sc.iLL = p.scales[p.sc_inv_LL];
    sc.iLW = 1.0f / (sc.sL * sc.sW); sc.iWQ = 1.0f / (sc.sW * sc.sQ);
    for (int j = 0; j < p.nbp; ++j) {
        const long long j0 = p.J0 + (long long)NB * j;
        const int nb = (int)min((long long)NB, p.N - j0);
        const bool rows_below = j0 + NB < p.N;
        PANEL_CLK(16 * j + 0);
        if (j > 0) {                                   // the owner of row block j has applied the updates of columns < j
            if (t == 0) *s_flag = spin_until_set(p.flags + F_DIAG_READY + j, abort_flag) ? 1 : 0;
            __syncthreads();
        }
        const float *Ab = p.A + j0 * p.ld + j0;
        PANEL_CLK(16 * j + 1);
        // global -> shared by cp.async (no registers in between: all 16 copies of a thread are in flight at once;
        // .cg = through L2, where the worker's release made the block visible); then the strict upper triangle of the
        // diagonal-crossing groups is cleared and a ragged last block is padded with the identity
#pragma unroll
        for (int q = 0; q < NB * NB / 4 / NUM_THREADS; ++q) {
            const int idx = t + q * NUM_THREADS;
            const int i = idx >> 5, k4 = (idx & 31) << 2;
            float *sdst = S + i * LDS + k4;
            if (i < nb && k4 <= i)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(sdst)), "l"(Ab + (long long)i * p.ld + k4) : "memory");
            else
                *reinterpret_cast<float4 *>(sdst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (t < NB) {
            const int i = t;
            if (i < nb) {
                const int k4 = (i >> 2) << 2;
#pragma unroll
                for (int c = 1; c < 4; ++c)
                    if (k4 + c > i) S[i * LDS + k4 + c] = 0.f;
            } else {
                S[i * LDS + i] = 1.f;
            }
        }
        __syncthreads();
        PANEL_CLK(16 * j + 2);
        // Cholesky + inverse of the block; the planes of W_jj (B operand of every worker's panel product) go to
        // global memory row block by row block while the factorisation is still running (chol_chain.cuh)
        cchain::factor_invert_block(S, W, Q, ring, nb, j0, p.info, rows_below, p.Ws_hi, p.Ws_lo, p.ld, sc);
        PANEL_CLK(16 * j + 3);
        PANEL_CLK(16 * j + 4);
        __syncthreads();
        if (t == 0) st_release(p.flags + F_DIAG_DONE + j, 1);
        PANEL_CLK(16 * j + 5);
        // off the critical path: the factor block itself, fp32 in place.  (Its fp16 planes are never an operand: the
        // trailing updates, the panel products and the triangular inverse only read off-diagonal blocks of Ls.)
#pragma unroll 4
        for (int q = t; q < NB * NB / 4; q += NUM_THREADS) {
            const int i = q >> 5, k4 = (q & 31) << 2;
            if (i >= nb || k4 > i) continue;
            const float4 s4 = *reinterpret_cast<const float4 *>(S + i * LDS + k4);
            float *dst = const_cast<float *>(Ab) + (long long)i * p.ld + k4;
            if (k4 + 3 <= i) *reinterpret_cast<float4 *>(dst) = s4;
            else {
                const float e[4] = {s4.x, s4.y, s4.z, s4.w};
                for (int c = 0; c < 4; ++c) if (k4 + c <= i) dst[c] = e[c];
            }
        }
        __syncthreads();                                // S / W are reused by the next diagonal block
        PANEL_CLK(16 * j + 6);
    }
    if (t == 0 && ld_acquire(abort_flag) != 0) atomicCAS(p.info, 0, -1);
}

// L[rb][j] out of the TMEM scratch (this thread: row r, columns [64 half, 64 half + 64) of the block):
// TO_SMEM: fp16 planes into the swizzled A-operand buffer; TO_GLOBAL: fp32 in place + the global planes.
template <bool TO_SMEM, bool TO_GLOBAL>
__device__ __forceinline__ void emit_L(const Args &p, uint32_t tlane, unsigned char *xa_gen, int r, int half, long long gr,
                                       long long col0, bool valid, float inv_AW, float sL) {
    float *arow = p.A + gr * p.ld + col0 + 64 * half;
    const long long poff = gr * p.ld + col0 + 64 * half;
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 32) {
        uint32_t u[32];
        tc::tmem_ld32(tlane + (uint32_t)(64 * half + cc), u);
        tc::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(u[q]) * inv_AW;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 hi, lo;
            split8(v + 8 * q, valid ? sL : 0.f, hi, lo);
            if (TO_SMEM) {
                const uint32_t off = opnd_offset(r, 64 * half + cc + 8 * q);
                *reinterpret_cast<uint4 *>(xa_gen + off) = hi;
                *reinterpret_cast<uint4 *>(xa_gen + PLANE_BYTES + off) = lo;
            }
            if (TO_GLOBAL && valid) {
                *reinterpret_cast<float4 *>(arow + cc + 8 * q) = make_float4(v[8 * q], v[8 * q + 1], v[8 * q + 2], v[8 * q + 3]);
                *reinterpret_cast<float4 *>(arow + cc + 8 * q + 4) = make_float4(v[8 * q + 4], v[8 * q + 5], v[8 * q + 6], v[8 * q + 7]);
                *reinterpret_cast<uint4 *>(p.Ls_hi + poff + cc + 8 * q) = hi;
                *reinterpret_cast<uint4 *>(p.Ls_lo + poff + cc + 8 * q) = lo;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// a worker (CTAs 1..)
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void worker_role(const Args &p, const CUtensorMap *mLhi, const CUtensorMap *mLlo, const CUtensorMap *mWhi,
                            const CUtensorMap *mWlo, unsigned char *smem, uint64_t *bars, uint32_t *tmem_slot) {
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int quarter = warp & 3, half = warp >> 2;           // TMEM lanes 32 q .., columns 64 h .. of a 128-column block
    const int r = quarter * 32 + lane;                        // row of this thread inside the row block
    const uint32_t xa = tc::smem_u32(smem), b0 = xa + OPND_BYTES, b1 = b0 + OPND_BYTES;
    unsigned char *xa_gen = smem;
    const uint32_t bar0 = tc::smem_u32(bars);
    const uint32_t BAR_TMA0 = bar0, BAR_TMA1 = bar0 + 8, BAR_MMA = bar0 + 16, BAR_AUX = bar0 + 24;
    int *abort_flag = p.flags + F_ABORT;

    if (t == 0) {
        tc::mbar_init(BAR_TMA0, 1); tc::mbar_init(BAR_TMA1, 1); tc::mbar_init(BAR_MMA, 1); tc::mbar_init(BAR_AUX, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(mLhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(mLlo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(mWhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(mWlo) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16);

    const float sA = p.scales[p.sc_A], sL = p.scales[p.sc_L];
    const float inv_AW = p.scales[p.sc_inv_AW], inv_LL = p.scales[p.sc_inv_LL];
    uint32_t ph_tma0 = 0, ph_tma1 = 0, ph_mma = 0, ph_aux = 0;   // barrier phases (t == 0 tracks the TMA / AUX ones)
    bool live = true;                                            // t == 0 only: false once a wait was given up

    const long long R = (p.N - p.J0 + NB - 1) / NB;              // row blocks of the panel, including the diagonal one
    const int W = (int)gridDim.x - 1;
    for (long long rb = 1 + ((int)blockIdx.x - 1); rb < R; rb += W) {
        const long long row0 = p.J0 + rb * NB;
        const long long gr = row0 + r;
        const bool valid = gr < p.N;
        const int ncol = (int)min((long long)p.nbp, rb);
        const int cmax = (int)min((long long)p.nbp - 1, rb);
        const bool near_blk = rb < p.nbp;
        for (int j = 0; j < ncol; ++j) {
            const long long col0 = p.J0 + (long long)NB * j;
            // (1) X = A[rb][j] - U_j  -> fp16 hi/lo, swizzled, into XA (does not need the chain: done before the wait)
#ifdef GPG_PANEL_PROFILE
            const bool prof = rb < p.nbp && j == (int)rb - 1;
            if (prof) PANEL_CLK(128 + 16 * (int)rb + 0);
#endif
            {
                float *arow = p.A + gr * p.ld + col0 + 64 * half;
#pragma unroll 1
                for (int cc = 0; cc < 64; cc += 32) {
                    float v[32];
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 a4 = *reinterpret_cast<const float4 *>(arow + cc + 4 * q);
                            v[4 * q] = a4.x; v[4 * q + 1] = a4.y; v[4 * q + 2] = a4.z; v[4 * q + 3] = a4.w;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] = 0.f;
                    }
                    if (j > 0) {
                        uint32_t u[32];
                        tc::tmem_ld32(tlane + (uint32_t)(NB * j + 64 * half + cc), u);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 32; ++q) v[q] -= __uint_as_float(u[q]) * inv_LL;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 hi, lo;
                        split8(v + 8 * q, valid ? sA : 0.f, hi, lo);
                        const uint32_t off = opnd_offset(r, 64 * half + cc + 8 * q);
                        *reinterpret_cast<uint4 *>(xa_gen + off) = hi;
                        *reinterpret_cast<uint4 *>(xa_gen + PLANE_BYTES + off) = lo;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic smem writes -> visible to the MMA
            tc::tcgen05_fence_before();
            __syncthreads();
            // the row block the chain waits for: its diagonal block A[rb][rb] is fetched now (cp.async into the idle B1
            // buffer, same XOR-swizzled fp32 layout as the staging tile below), off the hand-over path
            const bool final_near = near_blk && j == (int)rb - 1;
            if (final_near) {
                const float *dblk = p.A + row0 * p.ld + row0;
                const int nrows = (int)min((long long)NB, p.N - row0);
#pragma unroll
                for (int q = 0; q < NB * NB / 4 / NUM_THREADS; ++q) {
                    const int idx = t + q * NUM_THREADS, i = idx >> 5, c4 = idx & 31;
                    if (i < nrows && 4 * c4 <= i)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(b1 + (uint32_t)(i * NB + ((c4 ^ (i & 31)) << 2)) * 4u),
                                     "l"(dblk + (long long)i * p.ld + 4 * c4) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            // (2) W_jj planes -> B0 as soon as the chain has published them; (3) scratch = X W_jj^T
#ifdef GPG_PANEL_PROFILE
            if (prof) PANEL_CLK(128 + 16 * (int)rb + 1);
#endif
            if (t == 0) {
                live = live && spin_until_set(p.flags + F_DIAG_DONE + j, abort_flag);
#ifdef GPG_PANEL_PROFILE
                if (prof) PANEL_CLK(128 + 16 * (int)rb + 2);
#endif
                tc::tcgen05_fence_after();
                if (live) {
                    fence_proxy_async();
                    load_operand(b0, mWhi, mWlo, BAR_TMA0, (int)col0, (int)col0);
                    tc::mbar_wait(BAR_TMA0, ph_tma0); ph_tma0 ^= 1;
#ifdef GPG_PANEL_PROFILE
                    if (prof) PANEL_CLK(128 + 16 * (int)rb + 3);
#endif
                    tc::tcgen05_fence_after();
                    issue_product(tmem_base, xa, b0, 0);
                }
                tc::tcgen05_commit(BAR_MMA);
            }
            tc::mbar_wait(BAR_MMA, ph_mma); ph_mma ^= 1;
            tc::tcgen05_fence_after();
#ifdef GPG_PANEL_PROFILE
            if (prof) PANEL_CLK(128 + 16 * (int)rb + 4);
#endif
            // (4) L[rb][j]: fp32 in place, planes to global, planes into XA (operand of the updates)
            if (final_near) {
                // The chain waits for THIS row block.  Critical order: operand planes (smem only) -> the one update that
                // is left, the own diagonal block (both operands L[rb][j]) -> A[rb][rb] - U_rb to global -> flag.  The
                // global copies of L[rb][j] (fp32 + planes) follow afterwards.
                emit_L<true, false>(p, tlane, xa_gen, r, half, gr, col0, valid, inv_AW, sL);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc::tcgen05_fence_before();
                __syncthreads();
                if (t == 0) {
                    tc::tcgen05_fence_after();
                    if (live) issue_product(tmem_base + (uint32_t)(NB * (int)rb), xa, xa, j > 0 ? 1u : 0u);
                    tc::tcgen05_commit(BAR_MMA);
                }
                tc::mbar_wait(BAR_MMA, ph_mma); ph_mma ^= 1;
                tc::tcgen05_fence_after();
#ifdef GPG_PANEL_PROFILE
                PANEL_CLK(128 + 16 * (int)rb + 5);
#endif
                // U_rb: TMEM (one row per thread) -> fp32 tile in XA (16-byte chunks XOR-swizzled by the row, conflict
                // free both ways) -> coalesced read-modify-write of the lower part of the diagonal block
                float *stg = reinterpret_cast<float *>(xa_gen);
#pragma unroll 1
                for (int cc = 0; cc < 64; cc += 32) {
                    uint32_t u[32];
                    tc::tmem_ld32(tlane + (uint32_t)(NB * (int)rb + 64 * half + cc), u);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int c4 = (64 * half + cc) / 4 + q;
                        *reinterpret_cast<float4 *>(stg + r * NB + ((c4 ^ (r & 31)) << 2)) =
                            make_float4(__uint_as_float(u[4 * q]) * inv_LL, __uint_as_float(u[4 * q + 1]) * inv_LL,
                                        __uint_as_float(u[4 * q + 2]) * inv_LL, __uint_as_float(u[4 * q + 3]) * inv_LL);
                    }
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                tc::tcgen05_fence_before();
                __syncthreads();
                {
                    float *dblk = p.A + row0 * p.ld + row0;
                    const int nrows = (int)min((long long)NB, p.N - row0);
                    const float *pre = reinterpret_cast<const float *>(xa_gen + 2 * OPND_BYTES);      // B1: the prefetched block
#pragma unroll
                    for (int q = 0; q < NB * NB / 4 / NUM_THREADS; ++q) {
                        const int idx = t + q * NUM_THREADS, i = idx >> 5, c4 = idx & 31;
                        if (i < nrows && 4 * c4 <= i) {
                            const int so = i * NB + ((c4 ^ (i & 31)) << 2);
                            const float4 a4 = *reinterpret_cast<const float4 *>(pre + so);
                            const float4 u4 = *reinterpret_cast<const float4 *>(stg + so);
                            *reinterpret_cast<float4 *>(dblk + (long long)i * p.ld + 4 * c4) =
                                make_float4(a4.x - u4.x, a4.y - u4.y, a4.z - u4.z, a4.w - u4.w);
                        }
                    }
                }
                // no per-thread fence: the CTA barrier orders these stores before thread 0's release at gpu scope
                __syncthreads();
                if (t == 0) st_release(p.flags + F_DIAG_READY + (int)rb, 1);
#ifdef GPG_PANEL_PROFILE
                PANEL_CLK(128 + 16 * (int)rb + 6);
#endif
                emit_L<false, true>(p, tlane, xa_gen, r, half, gr, col0, valid, inv_AW, sL);
                tc::tcgen05_fence_before();
                __syncthreads();
                if (t == 0) st_release(p.flags + F_ROW_DONE + 4 * (int)rb + j, 1);
#ifdef GPG_PANEL_PROFILE
                PANEL_CLK(128 + 16 * (int)rb + 7);
#endif
            } else {
                emit_L<true, true>(p, tlane, xa_gen, r, half, gr, col0, valid, inv_AW, sL);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc::tcgen05_fence_before();
                __syncthreads();
                if (near_blk && t == 0) st_release(p.flags + F_ROW_DONE + 4 * (int)rb + j, 1);
            }
            // (5) U_c (+)= L[rb][j] L[c][j]^T for the later columns c of the panel
            const int nup = final_near ? 0 : cmax - j;
            if (nup > 0) {
                if (t == 0) {
                    tc::tcgen05_fence_after();
                    int used = 0;                              // B buffers filled in this step
                    for (int c = j + 1; c <= cmax; ++c) {
                        const uint32_t d = tmem_base + (uint32_t)(NB * c);
                        if (c == rb) {                         // own diagonal block: both operands are L[rb][j]
                            if (live) issue_product(d, xa, xa, j > 0 ? 1u : 0u);
                            continue;
                        }
                        const int b = used & 1;
                        if (used >= 2) {                       // the buffer is still being read by an earlier product
                            tc::tcgen05_commit(BAR_AUX);
                            tc::mbar_wait(BAR_AUX, ph_aux); ph_aux ^= 1;
                            tc::tcgen05_fence_after();
                        }
                        live = live && spin_until_set(p.flags + F_ROW_DONE + 4 * c + j, abort_flag);
                        if (live) {
                            fence_proxy_async();
                            load_operand(b ? b1 : b0, mLhi, mLlo, b ? BAR_TMA1 : BAR_TMA0, (int)col0, (int)(p.J0 + (long long)NB * c));
                            if (b) { tc::mbar_wait(BAR_TMA1, ph_tma1); ph_tma1 ^= 1; }
                            else { tc::mbar_wait(BAR_TMA0, ph_tma0); ph_tma0 ^= 1; }
                            tc::tcgen05_fence_after();
                            issue_product(d, xa, b ? b1 : b0, j > 0 ? 1u : 0u);
                        }
                        ++used;
                    }
                    tc::tcgen05_commit(BAR_MMA);
                }
                tc::mbar_wait(BAR_MMA, ph_mma); ph_mma ^= 1;
                tc::tcgen05_fence_after();
            }
            tc::tcgen05_fence_before();
            __syncthreads();                                   // XA / TMEM scratch are rewritten by the next step
            tc::tcgen05_fence_after();
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
chol_panel_kernel(const __grid_constant__ CUtensorMap mLhi, const __grid_constant__ CUtensorMap mLlo,
                  const __grid_constant__ CUtensorMap mWhi, const __grid_constant__ CUtensorMap mWlo, const Args p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = tc::smem_u32(smem_raw);
    unsigned char *base = smem_raw + (((raw + 1023u) & ~1023u) - raw);       // 1024-byte aligned operand buffers
    constexpr int BODY = SMEM_WORKER > SMEM_CHAIN ? SMEM_WORKER : SMEM_CHAIN;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + BODY);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    int *s_flag = reinterpret_cast<int *>(bars + 9);
    if (blockIdx.x == 0) chain_role(p, base, s_flag);
    else worker_role(p, &mLhi, &mLlo, &mWhi, &mWlo, base, bars, tmem_slot);
}

}  // namespace cpanel
