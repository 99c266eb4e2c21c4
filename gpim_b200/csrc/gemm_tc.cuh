// gemm_tc.cuh -- the tensor-core GEMM of the engine: tcgen05.mma (kind::f16, fp32 accumulators in
// TMEM) on fp16 hi/lo split operands, TMA-fed, warp-specialised, persistent with a dynamic tile
// scheduler.
//
//     D[m][n] = sum_k A[m][k] * B[n][k]          (both operands K-major, i.e. row-major [rows][k])
//
// fp32 fidelity from 16-bit tensor cores: every fp32 operand x is stored as two fp16 arrays
// hi = fp16(s x), lo = fp16(s x - hi) (s a power of two that keeps s|x| inside the fp16 range) and
// the product is evaluated as  Ahi Bhi + Ahi Blo + Alo Bhi  (three MMAs per k-step, the lo*lo term
// is below fp32 resolution).  hi+lo carries 22 significant bits, the accumulation is fp32.
//
// CTA layout (192 threads, 1 CTA / SM, persistent):
//   warp 0      producer: claims tiles from a global atomic counter, publishes them through a small
//               shared-memory ring, and issues the TMA loads (4 boxes per k-block: Ahi Alo Bhi Blo,
//               64-byte swizzle, BK = 32) into a 4-stage ring of 48 KB stages guarded by mbarriers
//   warp 1      MMA issuer: owns the 512 TMEM columns (two 128 x 256 fp32 accumulators), one
//               elected lane issues 6 tcgen05.mma per k-block and commits to the mbarriers
//   warps 2..5  epilogue: tcgen05.ld the finished accumulator (one TMEM lane = one output row per
//               thread) while the MMA warp already works on the next tile in the other buffer
//
// Epilogues:
//   ROWSUMSQ  part[nblk][m] = scale^-2 * sum_n D[m][n]^2   -- the diagonal predictive variance
//             diag(K*^T K^-1 K*) = colsum((L^-1 K*)^2) of util.conditional (reached from
//             gpr.py:248), with test points on the TMEM lanes so that the reduction is
//             thread-local; Q = L^-1 K* never leaves the SM.
//   STORE     C = alpha * scale^-1 * D + beta * C (fp32), optionally also emitting the fp16 hi/lo
//             split of the result (plain or transposed) for a following tensor-core GEMM.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include "common.cuh"
#include "gemm_simt.cuh"
#include <vector>

namespace tc {

#ifndef GPG_TC_BK
#define GPG_TC_BK 32
#endif
constexpr int BM = 128, BN = 256, BK = GPG_TC_BK;     // BK = 64: 128-byte swizzle; BK = 32: 64-byte swizzle
constexpr int STAGES = 128 / BK;                       // 192 KB of operand ring either way; more, shorter stages
                                                       // give the TMA more time to land each one
static_assert(BK == 32 || BK == 64, "BK must be 32 or 64");
constexpr int SCHED = 4;
constexpr int SCHED_SLOTS = SCHED;
constexpr int A_TILE_BYTES = BM * BK * 2;            // per hi or lo plane
constexpr int B_TILE_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
// mbarrier indices
constexpr int BAR_FULL = 0, BAR_EMPTY = STAGES, BAR_TFULL = 2 * STAGES, BAR_TEMPTY = 2 * STAGES + 2;
constexpr int BAR_SFULL = 2 * STAGES + 4, BAR_SEMPTY = 2 * STAGES + 4 + SCHED_SLOTS, BAR_COUNT = 2 * STAGES + 4 + 2 * SCHED_SLOTS;
constexpr int EPI_LD = 36;                            // floats per staged row: 16-byte aligned, conflict-free float4 access
constexpr int EPI_STAGE_BYTES = 4 * 32 * EPI_LD * 4;  // one 32 x 32 fp32 transpose buffer per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;

enum { EPI_ROWSUMSQ = 0, EPI_STORE = 1 };

struct Params {
    int M, N, K;                 // logical problem (per batch)
    int tiles_m, tiles_n, batch;
    int m_group;                 // tiles are walked in groups of m_group m-blocks (L2 residency of A); 0 = all
    int kb_mode, ke_mode, tile_mode;
    int n_off;                   // added to the n index in the triangular k-range rules (a launch over a sub-range of n-blocks)
    // operand placement inside the matrices described by the tensor maps (elements)
    int a_row0, a_col0, a_bs;    // batch b adds a_bs to both row and column
    int b_row0, b_col0, b_bs;
    int a_kbs, b_kbs;            // batch b additionally adds b * a_kbs / b * b_kbs to the COLUMN (k) only: split-K batches
    int epi;
    const float *scale_inv;      // device scalar: 1 / (scale_A * scale_B)
    // ROWSUMSQ
    float *part; long long ldpart;
    // STORE
    float *C; long long ldc; long long c_bs;      // batch b adds c_bs elements
    float *C2; long long ldc2;                    // optional second copy of the stored value (no batch stride)
    float alpha, beta;
    __half *S_hi, *S_lo; long long lds; long long s_bs;      // split of the result at [m][n]
    int s_ncols;                 // > 0: emit the split only for columns n < s_ncols
    __half *T_hi, *T_lo; long long ldt; long long t_bs;      // split of the result at [n][m] (transposed)
    const float *scale_out;      // device scalar multiplied into the emitted splits
    int *tile_counter;           // zero before launch
    const int *krange;           // optional: per m-tile [lo, hi) of the k indices where operand A is non-negligible
    unsigned long long *work_counter;   // optional: += number of (tile, k-block) pairs executed (one atomic per CTA)
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major swizzled operand tile: rows of BK halves (128 or 64 bytes = the swizzle span), 8-row groups contiguous.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8 * BK * 2) >> 4) << 32;          // stride byte offset: 8 rows of BK halves
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)(BK == 64 ? 2 : 4) << 61;           // SWIZZLE_128B / SWIZZLE_64B
    return d;
}
// kind::f16 instruction descriptor: D fp32, A/B fp16, both K-major, M x N.
__device__ __forceinline__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

struct Tile { int batch, mblk, nblk, kb_blk, ke_blk; };

__device__ __forceinline__ bool decode_tile(const Params &p, int t, Tile &o) {
    const int per_batch = p.tiles_m * p.tiles_n;
    o.batch = t / per_batch;
    const int r = t - o.batch * per_batch;
    const int gm = p.m_group > 0 ? p.m_group : p.tiles_m;
    const int grp = r / (gm * p.tiles_n);
    const int rr = r - grp * gm * p.tiles_n;
    const int gsize = min(gm, p.tiles_m - grp * gm);
    o.nblk = p.tiles_n - 1 - rr / gsize;
    o.mblk = grp * gm + rr % gsize;
    const int m0 = o.mblk * BM, n0 = o.nblk * BN + p.n_off;
    if (p.tile_mode == GEMM_TILES_LOWER && n0 > m0 + BM - 1) return false;
    int kb = 0, ke = p.K;
    if (p.kb_mode == GEMM_KB_N0) kb = n0;
    else if (p.kb_mode == GEMM_KB_MAXMN) kb = max(m0, n0);
    if (p.ke_mode == GEMM_KE_M) ke = min(p.K, m0 + BM);
    else if (p.ke_mode == GEMM_KE_N) ke = min(p.K, n0 + BN);
    if (p.krange) {              // compact support of the A rows of this m-tile
        kb = max(kb, __ldg(p.krange + 2 * o.mblk));
        ke = min(ke, __ldg(p.krange + 2 * o.mblk + 1));
        if (ke <= kb) return false;
    }
    o.kb_blk = kb / BK;
    o.ke_blk = (ke + BK - 1) / BK;
    return o.ke_blk > o.kb_blk;
}

__device__ __forceinline__ void split_fp16(float x, __half &hi, __half &lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, const Params p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;                   // 1024-byte aligned operand ring
    unsigned char *gen_tiles = smem_raw + (tiles - raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(gen_tiles + STAGES * STAGE_BYTES);
    // barrier map: BAR_FULL / BAR_EMPTY (operand ring), BAR_TFULL / BAR_TEMPTY (TMEM accumulators),
    // BAR_SFULL / BAR_SEMPTY (tile scheduler ring)
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    static_assert(BAR_COUNT * 8 + SCHED * 4 + 4 <= 256, "barrier block too small");
    volatile int *sched_tile = reinterpret_cast<volatile int *>(bars + BAR_COUNT);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + BAR_COUNT) + SCHED;
    float *epi_stage = reinterpret_cast<float *>(gen_tiles + STAGES * STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.tiles_m * p.tiles_n * p.batch;
    pdl_trigger();

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(BAR(BAR_FULL + i), 1); mbar_init(BAR(BAR_EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(BAR(BAR_TFULL + i), 1); mbar_init(BAR(BAR_TEMPTY + i), 4); }
        for (int i = 0; i < SCHED; ++i) { mbar_init(BAR(BAR_SFULL + i), 1); mbar_init(BAR(BAR_SEMPTY + i), 5); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                  // everything above overlapped the tail of the previous kernel

    if (warp == 0) {
        // ===================== producer: scheduler + TMA =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAhi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAlo) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBhi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
            int stage = 0; uint32_t phase = 0;
            int slot = 0; uint32_t sphase = 0;
            unsigned long long kblocks = 0;
            while (true) {
                mbar_wait(BAR(BAR_SEMPTY + slot), sphase ^ 1);
                const int t = atomicAdd(p.tile_counter, 1);
                sched_tile[slot] = t;
                mbar_arrive(BAR(BAR_SFULL + slot));
                if (++slot == SCHED) { slot = 0; sphase ^= 1; }
                if (t >= total_tiles) break;
                Tile ti;
                if (!decode_tile(p, t, ti)) continue;
                const int arow = p.a_row0 + ti.batch * p.a_bs + ti.mblk * BM;
                const int acol = p.a_col0 + ti.batch * (p.a_bs + p.a_kbs);
                const int brow = p.b_row0 + ti.batch * p.b_bs + ti.nblk * BN;
                const int bcol = p.b_col0 + ti.batch * (p.b_bs + p.b_kbs);
                kblocks += (unsigned long long)(ti.ke_blk - ti.kb_blk);
                for (int kb = ti.kb_blk; kb < ti.ke_blk; ++kb) {
                    mbar_wait(BAR(BAR_EMPTY + stage), phase ^ 1);
                    const uint32_t sbase = tiles + (uint32_t)stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(BAR(BAR_FULL + stage), STAGE_BYTES);
                    tma_load_2d(sbase, &tmAhi, BAR(BAR_FULL + stage), acol + kb * BK, arow);
                    tma_load_2d(sbase + A_TILE_BYTES, &tmAlo, BAR(BAR_FULL + stage), acol + kb * BK, arow);
                    tma_load_2d(sbase + 2 * A_TILE_BYTES, &tmBhi, BAR(BAR_FULL + stage), bcol + kb * BK, brow);
                    tma_load_2d(sbase + 2 * A_TILE_BYTES + B_TILE_BYTES, &tmBlo, BAR(BAR_FULL + stage), bcol + kb * BK, brow);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (p.work_counter && kblocks) atomicAdd(p.work_counter, kblocks);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN);
            int stage = 0; uint32_t phase = 0;
            int slot = 0; uint32_t sphase = 0;
            int buf = 0; uint32_t bphase = 0;
            while (true) {
                mbar_wait(BAR(BAR_SFULL + slot), sphase);
                const int t = sched_tile[slot];
                mbar_arrive(BAR(BAR_SEMPTY + slot));
                if (++slot == SCHED) { slot = 0; sphase ^= 1; }
                if (t >= total_tiles) break;
                Tile ti;
                if (!decode_tile(p, t, ti)) continue;
                mbar_wait(BAR(BAR_TEMPTY + buf), bphase ^ 1);         // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * BN;
                uint32_t acc = 0;
                for (int kb = ti.kb_blk; kb < ti.ke_blk; ++kb) {
                    mbar_wait(BAR(BAR_FULL + stage), phase);
                    tcgen05_fence_after();
                    const uint32_t sbase = tiles + (uint32_t)stage * STAGE_BYTES;
                    const uint64_t dAhi = make_smem_desc(sbase);
                    const uint64_t dAlo = make_smem_desc(sbase + A_TILE_BYTES);
                    const uint64_t dBhi = make_smem_desc(sbase + 2 * A_TILE_BYTES);
                    const uint64_t dBlo = make_smem_desc(sbase + 2 * A_TILE_BYTES + B_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);     // 16 fp16 = 32 bytes along K
                        tcgen05_mma_f16(d_tmem, dAhi + adv, dBhi + adv, idesc, acc);
                        acc = 1;
                        tcgen05_mma_f16(d_tmem, dAhi + adv, dBlo + adv, idesc, 1);
                        tcgen05_mma_f16(d_tmem, dAlo + adv, dBhi + adv, idesc, 1);
                    }
                    tcgen05_commit(BAR(BAR_EMPTY + stage));           // smem stage free once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(BAR(BAR_TFULL + buf));                 // accumulator complete
                if (++buf == 2) { buf = 0; bphase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int quarter = warp & 3;                                 // TMEM lanes this warp may touch
        const int row = quarter * 32 + lane;
        int slot = 0; uint32_t sphase = 0;
        int buf = 0; uint32_t bphase = 0;
        const float sinv = *p.scale_inv;
        const float sout = (p.epi == EPI_STORE && (p.S_hi || p.T_hi) && p.scale_out) ? *p.scale_out : 1.0f;
        while (true) {
            mbar_wait(BAR(BAR_SFULL + slot), sphase);
            const int t = sched_tile[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(BAR_SEMPTY + slot));
            if (++slot == SCHED) { slot = 0; sphase ^= 1; }
            if (t >= total_tiles) break;
            Tile ti;
            if (!decode_tile(p, t, ti)) continue;
            mbar_wait(BAR(BAR_TFULL + buf), bphase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * BN;
            const int m = ti.mblk * BM + row;
            const int n0 = ti.nblk * BN;
            if (p.epi == EPI_ROWSUMSQ) {
                float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float a = __uint_as_float(r[i]), b = __uint_as_float(r[i + 1]);
                        s0 = fmaf(a, a, s0);
                        s1 = fmaf(b, b, s1);
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(BAR_TEMPTY + buf));
                if (m < p.M) p.part[(long long)ti.nblk * p.ldpart + m] = (s0 + s1) * sinv * sinv;
            } else {
                // STORE: the accumulator arrives with one output row per lane; a 32 x 32 transpose through
                // shared memory turns the global accesses into 128-byte row segments (8 lanes x float4).
                float *Cb = p.C ? p.C + (long long)ti.batch * p.c_bs : nullptr;
                __half *Sh = p.S_hi ? p.S_hi + (long long)ti.batch * p.s_bs : nullptr;
                __half *Sl = p.S_lo ? p.S_lo + (long long)ti.batch * p.s_bs : nullptr;
                __half *Th = p.T_hi ? p.T_hi + (long long)ti.batch * p.t_bs : nullptr;
                __half *Tl = p.T_lo ? p.T_lo + (long long)ti.batch * p.t_bs : nullptr;
                const float a_eff = p.alpha * sinv;
                const bool c_vec = Cb && ((reinterpret_cast<uintptr_t>(Cb) & 15) == 0) && ((p.ldc & 3) == 0);
                float *C2b = p.C2;
                const bool c2_vec = C2b && ((reinterpret_cast<uintptr_t>(C2b) & 15) == 0) && ((p.ldc2 & 3) == 0);
                const bool s_vec = Sh && ((reinterpret_cast<uintptr_t>(Sh) & 7) == 0) &&
                                   ((reinterpret_cast<uintptr_t>(Sl) & 7) == 0) && ((p.lds & 3) == 0);
                float *stg = epi_stage + quarter * (32 * EPI_LD);
                const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
                const int m_base = ti.mblk * BM + quarter * 32;
                const bool rd = Cb && p.beta != 0.f && c_vec;
                float4 nxt4[8];
                auto load_c = [&](int c, float4 (&dst4)[8]) {     // the 8 float4 of C this lane updates in chunk c
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int mm = m_base + it * 4 + sub_r, nn = n0 + c + sub_c;
                        dst4[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (mm < p.M && nn + 4 <= p.N)
                            dst4[it] = *reinterpret_cast<const float4 *>(Cb + (long long)mm * p.ldc + nn);
                    }
                };
                // Fast path: interior tile whose only output is C <- alpha acc + beta C (the trailing updates of the
                // Cholesky).  Same data flow as the general loop below without its per-chunk option branches: the general
                // body is ~1 k instructions per chunk and the four epilogue warps were fetch-bound on it (ncu: 21 % of all
                // samples in no_inst, 19 us of epilogue per tile against 6 us of MMA).
                if (Cb && c_vec && (rd || p.beta == 0.f) && !C2b && !Sh && !Th && ti.mblk * BM + BM <= p.M && n0 + BN <= p.N) {
                    const float beta = rd ? p.beta : 0.f;
                    const float *crow = Cb + (long long)(m_base + sub_r) * p.ldc + n0 + sub_c;
                    const long long rstep = 4 * (long long)p.ldc;
                    float4 cur[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it)
                        cur[it] = rd ? *reinterpret_cast<const float4 *>(crow + it * rstep) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
                    for (int c = 0; c < BN; c += 32) {
                        uint32_t r[32];
                        tmem_ld32(taddr + c, r);
                        float4 nx[8];
                        if (rd && c + 32 < BN) {
#pragma unroll
                            for (int it = 0; it < 8; ++it) nx[it] = *reinterpret_cast<const float4 *>(crow + it * rstep + c + 32);
                        }
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<float4 *>(stg + lane * EPI_LD + 4 * q) =
                                make_float4(a_eff * __uint_as_float(r[4 * q]), a_eff * __uint_as_float(r[4 * q + 1]),
                                            a_eff * __uint_as_float(r[4 * q + 2]), a_eff * __uint_as_float(r[4 * q + 3]));
                        __syncwarp();
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            float4 x = *reinterpret_cast<const float4 *>(stg + (it * 4 + sub_r) * EPI_LD + sub_c);
                            x.x = fmaf(beta, cur[it].x, x.x); x.y = fmaf(beta, cur[it].y, x.y);
                            x.z = fmaf(beta, cur[it].z, x.z); x.w = fmaf(beta, cur[it].w, x.w);
                            *reinterpret_cast<float4 *>(const_cast<float *>(crow) + it * rstep + c) = x;
                        }
                        __syncwarp();
                        if (rd && c + 32 < BN) {
#pragma unroll
                            for (int it = 0; it < 8; ++it) cur[it] = nx[it];
                        }
                    }
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(BAR_TEMPTY + buf));
                    if (++buf == 2) { buf = 0; bphase ^= 1; }
                    continue;
                }
                // Second fast path: interior tile that is stored three ways and not read -- C, its fp16 planes and the
                // transposed planes (the W21 = -W22 T products of the triangular inverse; same fetch-bound general body).
                if (Cb && !rd && p.beta == 0.f && c_vec && s_vec && Sh && Th && !C2b && p.s_ncols == 0 &&
                    ti.mblk * BM + BM <= p.M && n0 + BN <= p.N) {
                    float *crow = Cb + (long long)(m_base + sub_r) * p.ldc + n0 + sub_c;
                    __half *shrow = Sh + (long long)(m_base + sub_r) * p.lds + n0 + sub_c;
                    __half *slrow = Sl + (long long)(m_base + sub_r) * p.lds + n0 + sub_c;
                    const long long rstep = 4 * (long long)p.ldc, sstep = 4 * (long long)p.lds;
#pragma unroll 1
                    for (int c = 0; c < BN; c += 32) {
                        uint32_t r[32];
                        tmem_ld32(taddr + c, r);
                        tmem_ld_wait();
                        __half *th = Th + (long long)(n0 + c) * p.ldt + m, *tl = Tl + (long long)(n0 + c) * p.ldt + m;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float vi = a_eff * __uint_as_float(r[i]);
                            r[i] = __float_as_uint(vi);
                            __half hi, lo;
                            split_fp16(vi * sout, hi, lo);
                            th[(long long)i * p.ldt] = hi;
                            tl[(long long)i * p.ldt] = lo;
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<float4 *>(stg + lane * EPI_LD + 4 * q) =
                                make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                            __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                        __syncwarp();
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const float4 x = *reinterpret_cast<const float4 *>(stg + (it * 4 + sub_r) * EPI_LD + sub_c);
                            *reinterpret_cast<float4 *>(crow + it * rstep + c) = x;
                            __align__(8) __half h4[4], l4[4];
                            split_fp16(x.x * sout, h4[0], l4[0]);
                            split_fp16(x.y * sout, h4[1], l4[1]);
                            split_fp16(x.z * sout, h4[2], l4[2]);
                            split_fp16(x.w * sout, h4[3], l4[3]);
                            *reinterpret_cast<uint2 *>(shrow + it * sstep + c) = *reinterpret_cast<const uint2 *>(h4);
                            *reinterpret_cast<uint2 *>(slrow + it * sstep + c) = *reinterpret_cast<const uint2 *>(l4);
                        }
                        __syncwarp();
                    }
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(BAR_TEMPTY + buf));
                    if (++buf == 2) { buf = 0; bphase ^= 1; }
                    continue;
                }
                if (rd) load_c(0, nxt4);
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    const int nbase = n0 + c;
                    if (nbase >= p.N) break;
                    uint32_t r[32];
                    tmem_ld32(taddr + c, r);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = a_eff * __uint_as_float(r[i]);
                    if (Th && m < p.M) {         // transposed split (beta == 0 only): lanes = consecutive columns of T
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (nbase + i < p.N) {
                                __half hi, lo;
                                split_fp16(v[i] * sout, hi, lo);
                                const long long off = (long long)(nbase + i) * p.ldt + m;
                                Th[off] = hi;
                                Tl[off] = lo;
                            }
                        }
                    }
                    if (Cb || Sh || C2b) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<float4 *>(stg + lane * EPI_LD + 4 * q) =
                                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        __syncwarp();
                        // C reads were issued one chunk ahead (old4); queue the next chunk's before this one's stores
                        float4 old4[8];
#pragma unroll
                        for (int it = 0; it < 8; ++it) old4[it] = nxt4[it];
                        if (rd && c + 32 < BN) load_c(c + 32, nxt4);
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int rr = it * 4 + sub_r;
                            const int mm = m_base + rr;
                            const int nn = nbase + sub_c;
                            if (mm >= p.M || nn >= p.N) continue;
                            float4 x = *reinterpret_cast<const float4 *>(stg + rr * EPI_LD + sub_c);
                            const bool full = nn + 4 <= p.N;
                            if (Cb) {
                                float *dst = Cb + (long long)mm * p.ldc + nn;
                                if (full && c_vec) {
                                    if (rd) {
                                        const float4 o = old4[it];
                                        x.x += p.beta * o.x; x.y += p.beta * o.y; x.z += p.beta * o.z; x.w += p.beta * o.w;
                                    }
                                    *reinterpret_cast<float4 *>(dst) = x;
                                } else {
                                    float xs[4] = {x.x, x.y, x.z, x.w};
                                    for (int e = 0; e < 4 && nn + e < p.N; ++e) {
                                        if (p.beta != 0.f) xs[e] += p.beta * dst[e];
                                        dst[e] = xs[e];
                                    }
                                    x = make_float4(xs[0], xs[1], xs[2], xs[3]);
                                }
                            }
                            if (C2b) {
                                float *dst2 = C2b + (long long)mm * p.ldc2 + nn;
                                if (full && c2_vec) *reinterpret_cast<float4 *>(dst2) = x;
                                else {
                                    const float xs[4] = {x.x, x.y, x.z, x.w};
                                    for (int e = 0; e < 4 && nn + e < p.N; ++e) dst2[e] = xs[e];
                                }
                            }
                            if (Sh && (p.s_ncols == 0 || nn < p.s_ncols)) {
                                __align__(8) __half h4[4], l4[4];
                                split_fp16(x.x * sout, h4[0], l4[0]);
                                split_fp16(x.y * sout, h4[1], l4[1]);
                                split_fp16(x.z * sout, h4[2], l4[2]);
                                split_fp16(x.w * sout, h4[3], l4[3]);
                                const long long off = (long long)mm * p.lds + nn;
                                if (full && s_vec) {
                                    *reinterpret_cast<uint2 *>(Sh + off) = *reinterpret_cast<const uint2 *>(h4);
                                    *reinterpret_cast<uint2 *>(Sl + off) = *reinterpret_cast<const uint2 *>(l4);
                                } else {
                                    for (int e = 0; e < 4 && nn + e < p.N; ++e) { Sh[off + e] = h4[e]; Sl[off + e] = l4[e]; }
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(BAR_TEMPTY + buf));
            }
            if (++buf == 2) { buf = 0; bphase ^= 1; }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
inline PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

// fp16 row-major matrix [rows][cols] with leading dimension ld (elements); box = 64 columns x box_rows rows.
// Encoded maps are memoised: the factorisation issues hundreds of launches over the same few planes.
struct TmapKey {
    const void *base; long long rows, cols, ld; int box_rows;
    bool operator==(const TmapKey &o) const {
        return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct TmapEntry { TmapKey key; CUtensorMap map; };

inline int make_tensor_map(CUtensorMap *map, const __half *base, long long rows, long long cols, long long ld,
                           int box_rows) {
    static thread_local std::vector<TmapEntry> cache;
    const TmapKey key{base, rows, cols, ld, box_rows};
    for (const TmapEntry &e : cache)
        if (e.key == key) { *map = e.map; return GPG_OK; }
    auto fn = get_encode_fn();
    if (!fn) { gpg_set_error("cuTensorMapEncodeTiled entry point not available"); return GPG_ECUDA; }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16) {
        gpg_set_error("tensor map needs a 16-byte aligned base and row pitch (ld=%lld)", ld);
        return GPG_EINVAL;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half *>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { gpg_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return GPG_ECUDA; }
    if (cache.size() >= 64) cache.erase(cache.begin());
    cache.push_back(TmapEntry{key, *map});
    return GPG_OK;
}

// One split operand: hi and lo fp16 matrices of identical geometry.
struct SplitMat {
    const __half *hi = nullptr, *lo = nullptr;
    long long rows = 0, cols = 0, ld = 0;
};

struct Launch {
    SplitMat A, B;
    Params p;
    int max_ctas = 0;            // > 0: cap of the persistent grid (leave SMs to a kernel running next to this one)
};

inline int launch(gpg_handle_s *h, Launch &L, cudaStream_t stream) {
    Params &p = L.p;
    if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.batch <= 0) return GPG_OK;
    if (p.T_hi && p.beta != 0.f) { gpg_set_error("gemm_tc: transposed split emission needs beta == 0"); return GPG_EINVAL; }
    static bool attr_set = false;
    if (!attr_set) {
        GPG_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    p.tiles_m = (p.M + BM - 1) / BM;
    p.tiles_n = (p.N + BN - 1) / BN;
    CUtensorMap mAhi, mAlo, mBhi, mBlo;
    GPG_TRY(make_tensor_map(&mAhi, L.A.hi, L.A.rows, L.A.cols, L.A.ld, BM));
    GPG_TRY(make_tensor_map(&mAlo, L.A.lo, L.A.rows, L.A.cols, L.A.ld, BM));
    GPG_TRY(make_tensor_map(&mBhi, L.B.hi, L.B.rows, L.B.cols, L.B.ld, BN));
    GPG_TRY(make_tensor_map(&mBlo, L.B.lo, L.B.rows, L.B.cols, L.B.ld, BN));
    GPG_TRY(gpg_tc_counter(h, stream, &p.tile_counter));
    const long long total = (long long)p.tiles_m * p.tiles_n * p.batch;
    const int grid = (int)std::min<long long>(total, L.max_ctas > 0 ? std::min(L.max_ctas, h->sm_count) : h->sm_count);
    GPG_CUDA_CHECK(launch_pdl(gemm_tc_kernel, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, stream, mAhi, mAlo, mBhi, mBlo, p));
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

// ---------------------------------------------------------------------------------------------
// fp32 -> fp16 hi/lo split of a row-major matrix (optionally transposed), scaled by *scale.
// Entries outside [rows) x [cols) of the destination padding are written as zero up to ld_dst.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_kernel(const float *__restrict__ src, long long ld_src, long long rows,
                                                    long long cols, const float *__restrict__ scale,
                                                    __half *__restrict__ hi, __half *__restrict__ lo, long long ld_dst,
                                                    int lower_only) {
    // one thread = 8 consecutive columns of one row
    const long long c8 = ((long long)blockIdx.x * 32 + (threadIdx.x & 31)) * 8;
    const long long r = (long long)blockIdx.y * 8 + (threadIdx.x >> 5);
    if (r >= rows || c8 >= ld_dst) return;
    const float s = *scale;
    __align__(16) __half h8[8], l8[8];
    float x[8];
    const bool zero_row = lower_only && c8 > r;
    if (!zero_row && c8 + 8 <= cols && (ld_src & 3) == 0) {
        const float4 a = *reinterpret_cast<const float4 *>(src + r * ld_src + c8);
        const float4 b = *reinterpret_cast<const float4 *>(src + r * ld_src + c8 + 4);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = (!zero_row && c8 + i < cols) ? src[r * ld_src + c8 + i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = x[i] * s;
        if (lower_only && c8 + i > r) v = 0.f;
        split_fp16(v, h8[i], l8[i]);
    }
    *reinterpret_cast<uint4 *>(hi + r * ld_dst + c8) = *reinterpret_cast<const uint4 *>(h8);
    *reinterpret_cast<uint4 *>(lo + r * ld_dst + c8) = *reinterpret_cast<const uint4 *>(l8);
}

// transposed variant: dst[c][r] = split(src[r][c] * s); 32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256) split_transpose_kernel(const float *__restrict__ src, long long ld_src,
                                                              long long rows, long long cols,
                                                              const float *__restrict__ scale, __half *__restrict__ hi,
                                                              __half *__restrict__ lo, long long ld_dst) {
    __shared__ float tile[32][33];
    const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const long long r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.f;
    }
    __syncthreads();
    const float s = *scale;
    for (int i = ty; i < 32; i += 8) {
        const long long c = c0 + i, r = r0 + tx;          // dst row = src column
        if (c < cols && r < ld_dst) {
            __half a, b;
            split_fp16(tile[tx][i] * s, a, b);
            hi[c * ld_dst + r] = a;
            lo[c * ld_dst + r] = b;
        }
    }
}

inline int split_matrix(gpg_handle_s *h, const float *src, long long ld_src, long long rows, long long cols,
                        const float *scale, __half *hi, __half *lo, long long ld_dst, int lower_only,
                        cudaStream_t stream) {
    if (rows <= 0) return GPG_OK;
    dim3 grid((unsigned)((ld_dst / 8 + 31) / 32), (unsigned)((rows + 7) / 8));
    split_kernel<<<grid, 256, 0, stream>>>(src, ld_src, rows, cols, scale, hi, lo, ld_dst, lower_only);
    GPG_LAUNCH_CHECK(h);
    return GPG_OK;
}

}  // namespace tc
