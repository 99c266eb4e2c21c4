// chol_chain.cuh -- the serial part of the Cholesky panel kernel (chol_panel.cuh): Cholesky AND inverse of one
// 128 x 128 diagonal block by ONE CTA of 8 warps, written so that only the 128 pivots are serial.
//
// The block is walked in four sub-steps of 32 columns (c0 = 32 p).  Sub-step p:
//
//   (A) warp 0, the PIVOT warp: Cholesky of the 32 x 32 sub-block in registers (lane = row).  Every scaled column j
//       goes into slot j of a 32-slot ring in shared memory together with 1 / l_jj, then a progress counter is bumped.
//       warp 5, the INVERSE FOLLOWER, consumes the ring column by column: the forward substitution x_j <- x_j / l_jj,
//       x_k <- x_k - x_j l_kj started from the identity gives L11^-T, so lane i ends up with column i of the inverse
//       W_pp of the sub-block IN LOCKSTEP with the factorisation.  No separate triangular inversion is left.
//       the other six warps meanwhile run the SHADOW tiles of sub-step k = p - 1:
//         (a) the part of the rank-32 trailing update the next sub-step does not need,
//         (b) row block k of the inverse:   W[k][0:k] = -W_kk Q[k][0:k],
//         (c) the running products          Q[r][0:k+1] += L[r][k] W[k][0:k+1],  r > k,
//         (d) fp16 hi/lo planes of row block k of W to global memory (B operand of the workers).
//   (B) all 8 warps: rows below the sub-block  X <- X W_pp^T,  then the rank-32 update of the NEXT 32 columns only.
//
// After the last pivot the only work left is W[3][0:3] = -W_33 Q[3][0:3] and the planes of those 32 rows.
// All block products are 16 x 32 x 32 tiles on mma.sync (fp16 hi/lo split, three products: mma_tile below).
//
// Shared memory: S (the block / its factor) and W (its inverse), 128 x 136 floats each; Q, 96 x 104; the ring.
// Measured on B200 (tools/panel_bench.cu, 512-column panel = 4 blocks): 166 -> 123 us.
#pragma once
#include "factor.cuh"

namespace cpanel {
// development aid (tools/panel_bench.cu): timestamps of the chain (CTA 0) and of the worker on the critical path
#ifdef GPG_PANEL_PROFILE
__device__ long long g_panel_clk[512];
__device__ __forceinline__ long long panel_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PANEL_CLK(slot) do { if (threadIdx.x == 0) { cpanel::g_panel_clk[slot] = clock64(); cpanel::g_panel_clk[256 + (slot)] = cpanel::panel_now(); } } while (0)
#else
#define PANEL_CLK(slot)
#endif

}  // namespace cpanel

namespace cchain {
#define CHAIN_CLK(slot) do { PANEL_CLK(64 + (slot)); __syncwarp(); } while (0)
#ifdef GPG_PANEL_PROFILE
#define SHADOW_CLK(slot) do { if (threadIdx.x == 192) cpanel::g_panel_clk[96 + (slot)] = clock64(); __syncwarp(); } while (0)
#else
#define SHADOW_CLK(slot)
#endif

constexpr int NB = 128;
constexpr int LDS = NB + 8;                    // rows 16-byte aligned, 8 banks of skew: the 64-bit fragment loads of
constexpr int LDQ = 96 + 8;                    // mma16x32 (row = lane / 4, column pair = lane % 4) are conflict-free
constexpr int RING_FLOATS = 32 * 32 + 32 + 32; // 32 columns, 1 / l_jj, progress counter (+ padding)
constexpr int SMEM_FLOATS = 2 * NB * LDS + 96 * LDQ + RING_FLOATS;
constexpr int SMEM_BYTES = SMEM_FLOATS * 4;

__device__ __forceinline__ void st_flag(int *p, int v) { asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_flag(const int *p) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void group_barrier(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- warp 0 ---------------------------------------------------------------------------------------------------
// Cholesky of S[c0.., c0..] (32 x 32), lane = row, fully unrolled (the row lives in registers).  Column j, scaled,
// -> ring[32 j + lane]; 1 / l_jj -> rsv[j]; progress = j + 1 once both are visible.  The next pivot's diagonal entry
// is updated, broadcast and sent through rsqrt before the column is exchanged (see factor_tri32).
// Everything lane-dependent is a SELECT, not a branch: a divergent branch plus its reconvergence costs more than the
// rest of a pivot step (measured: 235 instead of 95 cycles per column when ptxas chose branches).
// Returns 0 or 1 + index of the first bad pivot.
__device__ __forceinline__ void st_flag_lane0(int *p, int v, int lane) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %2, 0;\n\t@q st.volatile.shared.s32 [%0], %1;\n\t}"
                 ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v), "r"(lane) : "memory");
}
__device__ __forceinline__ void st_f32_lane0(float *p, float v, int lane) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %2, 0;\n\t@q st.volatile.shared.f32 [%0], %1;\n\t}"
                 ::"r"((unsigned)__cvta_generic_to_shared(p)), "f"(v), "r"(lane) : "memory");
}

__device__ __forceinline__ int pivot32(float *__restrict__ S, int c0, int lane, int valid, float *__restrict__ ring,
                                       float *__restrict__ rsv, int *progress) {
    constexpr unsigned FULL = 0xffffffffu;
    float row[32];
    __syncwarp();                                      // converged from here on: every shuffle below takes the fast path
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const V4<float> v = ld4(S + (c0 + lane) * LDS + c0 + 4 * q);
#pragma unroll
        for (int e = 0; e < 4; ++e) row[4 * q + e] = v.v[e];
    }
    float djj = __shfl_sync(FULL, row[0], 0);
    int bad = (!(djj > 0.f) && 0 < valid) ? 1 : 0;
    float rs = gpg_rsqrt(djj);                         // warp-uniform
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float lij = row[j] * rs;                 // lane j: djj / sqrt(djj) = l_jj
        row[j] = lij;
        float *cb = ring + 32 * j;
        cb[lane] = lij;
        st_f32_lane0(rsv + j, rs, lane);
        float rsn = 0.f;
        if (j < 31) {
            const bool owner = lane == j + 1;
            row[j + 1] = fmaf(-lij, owner ? lij : 0.f, row[j + 1]);
            const float dn = __shfl_sync(FULL, row[j + 1], j + 1);
            bad = (bad == 0 && !(dn > 0.f) && j + 1 < valid) ? j + 2 : bad;
            rsn = gpg_rsqrt(dn);
        }
        __syncwarp();
        st_flag_lane0(progress, j + 1, lane);
        if (j < 31) {
            const bool owner = lane == j + 1;
#pragma unroll
            for (int k4 = ((j + 1) / 4) * 4; k4 < 32; k4 += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(cb + k4);
                const float c[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (k4 + e == j + 1) row[j + 1] = fmaf(-lij, owner ? 0.f : c[e], row[j + 1]);   // the owner has done its own
                    else if (k4 + e > j + 1) row[k4 + e] = fmaf(-lij, c[e], row[k4 + e]);           // meaningful for lane >= k only
                }
            }
            rs = rsn;
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {                      // right of the diagonal the registers hold scratch: zeros go out
        V4<float> v;
#pragma unroll
        for (int e = 0; e < 4; ++e) v.v[e] = (4 * q + e <= lane) ? row[4 * q + e] : 0.f;
        st4(S + (c0 + lane) * LDS + c0 + 4 * q, v);
    }
    return bad;
}

// ---- followers --------------------------------------------------------------------------------------------------
// X <- X L11^-T for 32 rows (lane = row), consuming the pivot warp's ring as it fills.
//   INVERSE = false: rows [r0, r0 + 32) of S, columns [c0, c0 + 32), in place.
//   INVERSE = true : X = I; the result is L11^-T, written transposed into W[c0.., c0..] (the inverse of the sub-block).
template <bool INVERSE>
__device__ __forceinline__ void follow32(float *__restrict__ M, int r0, int c0, int lane, const float *__restrict__ ring,
                                         const float *__restrict__ rsv, const int *progress) {
    float x[32];
    if (INVERSE) {
#pragma unroll
        for (int k = 0; k < 32; ++k) x[k] = (k == lane) ? 1.f : 0.f;
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const V4<float> v = ld4(M + (r0 + lane) * LDS + c0 + 4 * q);
#pragma unroll
            for (int e = 0; e < 4; ++e) x[4 * q + e] = v.v[e];
        }
    }
    int seen = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if (seen <= j) {                                // warp-uniform
            do { seen = ld_flag(progress); } while (seen <= j);
            // compiler barrier only: shared-memory accesses of a warp are performed in order, and the pivot warp's
            // flag store follows its column stores (a MEMBAR here costs ~76 cycles per column)
            asm volatile("" ::: "memory");
        }
        const float *cb = ring + 32 * j;
        const float l = x[j] * *reinterpret_cast<const volatile float *>(rsv + j);
        x[j] = l;
#pragma unroll
        for (int k4 = ((j + 1) / 4) * 4; k4 < 32; k4 += 4) {
            float4 v;
            asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"((unsigned)__cvta_generic_to_shared(cb + k4)));
            const float c[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (k4 + e > j) x[k4 + e] = fmaf(-l, c[e], x[k4 + e]);
        }
    }
    if (INVERSE) {
#pragma unroll
        for (int k = 0; k < 32; ++k) M[(c0 + k) * LDS + c0 + lane] = x[k];      // x_i[k] = W[k][i]; zero for k < i
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            V4<float> v;
#pragma unroll
            for (int e = 0; e < 4; ++e) v.v[e] = x[4 * q + e];
            st4(M + (r0 + lane) * LDS + c0 + 4 * q, v);
        }
    }
}

// ---- 16 x 32 x 32 products on the tensor cores, one warp each ---------------------------------------------------
// Shared-memory BANDWIDTH is what bounds this CTA: a register-tiled SIMT product re-reads its operands through
// 128-bit loads (~50 KB per 32 x 32 x 32 block) and, with eight warps doing so at once, every phase -- the pivot warp
// included -- queued behind the load/store unit (measured: 3-5 k cycles per block product).  mma.sync reads every
// operand element once.  fp32 accuracy by the same three-product split the tcgen05 kernels use, made in registers:
// x s = hi + lo in fp16 with the power-of-two operand scales of factor_tc.cuh, A B ~ (Alo Bhi + Ahi Blo + Ahi Bhi) / (sa sb).
// (TF32 halves need no scales but cost four times the tensor instructions plus an emulated conversion: measured 2.8 k
// cycles per tile against ~0.5 k.)
enum { OP_SET = 0, OP_ADD = 1, OP_SETNEG = 2, OP_SUB = 3 };

__device__ __forceinline__ void split_h2(float x0, float x1, float s, uint32_t &hi, uint32_t &lo) {
    const float a = x0 * s, b = x1 * s;
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One tile of work: C[16 x 32] (op)= A[16 x 32] * B, B given as B[k][n] (nt = false) or as B[n][k] (nt = true:
// C (op)= A B^T).  sa / sb: operand scales (powers of two), inv = 1 / (sa sb).  A may alias C (everything is in
// registers before anything is written).  lower_row0 >= 0: the tile belongs to a diagonal block and starts at its row
// lower_row0; entries right of the diagonal are left alone.
struct Tile {
    float *C; const float *A, *B;
    int ldc, lda, ldb, op, lower_row0;
    bool nt;
    float sa, sb, inv;
};

// All loads first, then the conversions, then 24 tensor instructions as four independent chains: a single warp has
// nobody to hide its latencies behind.
__device__ __forceinline__ void mma_tile(const Tile &J, int lane) {
    const int g = lane >> 2, t = lane & 3;
    float2 xa[2][4], yb[2][4][2], old[4][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const float *a = J.A + g * J.lda + 16 * ks + 2 * t;
        xa[ks][0] = *reinterpret_cast<const float2 *>(a);
        xa[ks][1] = *reinterpret_cast<const float2 *>(a + 8 * J.lda);
        xa[ks][2] = *reinterpret_cast<const float2 *>(a + 8);
        xa[ks][3] = *reinterpret_cast<const float2 *>(a + 8 * J.lda + 8);
    }
    if (J.nt) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int n = 0; n < 4; ++n) {                // B(k = 2t, 2t+1 ; n = g), B(k = 2t+8, 2t+9 ; n = g)
                const float *b = J.B + (8 * n + g) * J.ldb + 16 * ks + 2 * t;
                yb[ks][n][0] = *reinterpret_cast<const float2 *>(b);
                yb[ks][n][1] = *reinterpret_cast<const float2 *>(b + 8);
            }
    } else {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const float *b = J.B + (16 * ks + 2 * t) * J.ldb + 8 * n + g;
                yb[ks][n][0] = make_float2(b[0], b[J.ldb]);
                yb[ks][n][1] = make_float2(b[8 * J.ldb], b[9 * J.ldb]);
            }
    }
    const bool rmw = J.op == OP_ADD || J.op == OP_SUB;
    if (rmw) {
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) old[n][h] = *reinterpret_cast<const float2 *>(J.C + (g + 8 * h) * J.ldc + 8 * n + 2 * t);
    }
    __syncwarp();
    uint32_t ah[2][4], al[2][4], bh[2][4][2], bl[2][4][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int q = 0; q < 4; ++q) split_h2(xa[ks][q].x, xa[ks][q].y, J.sa, ah[ks][q], al[ks][q]);
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int q = 0; q < 2; ++q) split_h2(yb[ks][n][q].x, yb[ks][n][q].y, J.sb, bh[ks][n][q], bl[ks][n][q]);
    }
    float acc[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[n][q] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {                     // small terms first
#pragma unroll
        for (int n = 0; n < 4; ++n) mma_f16(acc[n], al[ks], bh[ks][n][0], bh[ks][n][1]);
#pragma unroll
        for (int n = 0; n < 4; ++n) mma_f16(acc[n], ah[ks], bl[ks][n][0], bl[ks][n][1]);
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int n = 0; n < 4; ++n) mma_f16(acc[n], ah[ks], bh[ks][n][0], bh[ks][n][1]);
    const float sgn = (J.op == OP_SETNEG || J.op == OP_SUB) ? -J.inv : J.inv;
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = g + 8 * h, col = 8 * n + 2 * t;
            float2 o = make_float2(acc[n][2 * h] * sgn, acc[n][2 * h + 1] * sgn);
            if (rmw) {
                o.x += old[n][h].x;
                o.y += old[n][h].y;
                if (J.lower_row0 >= 0) {                 // keep what lies right of the diagonal
                    if (col > J.lower_row0 + r) o.x = old[n][h].x;
                    if (col + 1 > J.lower_row0 + r) o.y = old[n][h].y;
                }
            }
            *reinterpret_cast<float2 *>(J.C + r * J.ldc + col) = o;
        }
}

// fp16 hi/lo planes of rows [32 k, 32 k + 32) of W (all 128 columns; zero right of the diagonal blocks) -> global
__device__ __forceinline__ void publish_rows(const float *__restrict__ W, int k, int tid, int nthreads, __half *__restrict__ hi,
                                             __half *__restrict__ lo, long long ld, long long j0, float sW) {
    for (int q = tid; q < 32 * 16; q += nthreads) {
        const int i = 32 * k + (q >> 4), k8 = (q & 15) << 3;
        uint4 h = make_uint4(0u, 0u, 0u, 0u), l = make_uint4(0u, 0u, 0u, 0u);
        if (k8 < 32 * (k + 1)) {
            const V4<float> wa = ld4(W + i * LDS + k8), wb = ld4(W + i * LDS + k8 + 4);
            const float v[8] = {wa.v[0], wa.v[1], wa.v[2], wa.v[3], wb.v[0], wb.v[1], wb.v[2], wb.v[3]};
            __half2 hh[4], ll[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a = v[2 * e] * sW, b = v[2 * e + 1] * sW;
                hh[e] = __floats2half2_rn(a, b);
                const float2 f = __half22float2(hh[e]);
                ll[e] = __floats2half2_rn(a - f.x, b - f.y);
            }
            h = make_uint4(*reinterpret_cast<unsigned *>(&hh[0]), *reinterpret_cast<unsigned *>(&hh[1]),
                           *reinterpret_cast<unsigned *>(&hh[2]), *reinterpret_cast<unsigned *>(&hh[3]));
            l = make_uint4(*reinterpret_cast<unsigned *>(&ll[0]), *reinterpret_cast<unsigned *>(&ll[1]),
                           *reinterpret_cast<unsigned *>(&ll[2]), *reinterpret_cast<unsigned *>(&ll[3]));
        }
        const long long off = (j0 + i) * ld + j0 + k8;
        *reinterpret_cast<uint4 *>(hi + off) = h;
        *reinterpret_cast<uint4 *>(lo + off) = l;
    }
    // no fence here: the caller's CTA barrier orders these stores before thread 0's st.release.gpu, which is cumulative
}

// the shadow jobs of sub-step k, run by the G = 6 warps that are free during sub-step k + 1 (g = index in the group), in
// tiles of 16 x 32.  Blocks are named by their 32-row / 32-column block indices; Q[r][c] sits at Q + 32 (r - 1) LDQ + 32 c.
struct Scales { float sA, sL, sW, sQ, iAW, iLL, iLW, iWQ; };    // operand scales and 1 / (product of two) of them

// The tiles of one phase, by index; false past the last one.  Blocks are named by their 32-row / 32-column block
// indices, h = which 16-row half; Q[r][c] sits at Q + 32 (r - 1) LDQ + 32 c.
enum { PH_SOLVE = 0, PH_UPDATE = 1, PH_SHADOW1 = 2, PH_SHADOW2 = 3, PH_TAIL = 4 };
__device__ __forceinline__ bool decode_tile(int phase, int k, int idx, float *S, float *W, float *Q, const Scales &sc, Tile &J) {
    auto Sb = [&](int r, int c, int h) { return S + (32 * r + 16 * h) * LDS + 32 * c; };
    auto Wb = [&](int r, int c, int h) { return W + (32 * r + 16 * h) * LDS + 32 * c; };
    auto Qb = [&](int r, int c, int h) { return Q + (32 * (r - 1) + 16 * h) * LDQ + 32 * c; };
    const int h = idx & 1, b = idx >> 1;
    J.lower_row0 = -1;
    J.lda = LDS; J.ldb = LDS; J.ldc = LDS;
    if (phase == PH_SOLVE) {                             // rows below sub-block k: X <- X W_kk^T
        if (b >= 3 - k) return false;
        J.C = Sb(k + 1 + b, k, h); J.A = J.C; J.B = Wb(k, k, 0);
        J.op = OP_SET; J.nt = true; J.sa = sc.sA; J.sb = sc.sW; J.inv = sc.iAW;
        return true;
    }
    if (phase == PH_UPDATE) {                            // rank-32 update of the NEXT 32 columns: S[r][k+1] -= L[r][k] L[k+1][k]^T
        if (b >= 3 - k) return false;
        const int r = k + 1 + b;
        J.C = Sb(r, k + 1, h); J.A = Sb(r, k, h); J.B = Sb(k + 1, k, 0);
        J.op = OP_SUB; J.nt = true; J.sa = sc.sL; J.sb = sc.sL; J.inv = sc.iLL;
        if (b == 0) J.lower_row0 = 16 * h;
        return true;
    }
    if (phase == PH_SHADOW1) {
        // (b) row block k of the inverse: W[k][c] = -W_kk Q[k][c], c < k
        if (b < k) {
            J.C = Wb(k, b, h); J.A = Wb(k, k, h); J.B = Qb(k, b, 0); J.ldb = LDQ;
            J.op = OP_SETNEG; J.nt = false; J.sa = sc.sW; J.sb = sc.sQ; J.inv = sc.iWQ;
            return true;
        }
        // (a) trailing blocks the next sub-step does not touch: S[r][c] -= L[r][k] L[c][k]^T, r >= c >= k + 2
        const int a = b - k;
        int r, cc;
        if (k == 0) { if (a >= 3) return false; r = a == 0 ? 2 : 3; cc = a == 2 ? 3 : 2; }
        else if (k == 1) { if (a >= 1) return false; r = 3; cc = 3; }
        else return false;
        J.C = Sb(r, cc, h); J.A = Sb(r, k, h); J.B = Sb(cc, k, 0);
        J.op = OP_SUB; J.nt = true; J.sa = sc.sL; J.sb = sc.sL; J.inv = sc.iLL;
        if (r == cc) J.lower_row0 = 16 * h;
        return true;
    }
    if (phase == PH_SHADOW2) {                           // (c) Q[r][c] (+)= L[r][k] W[k][c], r > k, c <= k
        if (b >= (3 - k) * (k + 1)) return false;
        const int r = k + 1 + b / (k + 1), cc = b % (k + 1);
        J.C = Qb(r, cc, h); J.ldc = LDQ; J.A = Sb(r, k, h); J.B = Wb(k, cc, 0);
        J.op = cc == k ? OP_SET : OP_ADD; J.nt = false; J.sa = sc.sL; J.sb = sc.sW; J.inv = sc.iLW;
        return true;
    }
    // PH_TAIL: W[3][c] = -W_33 Q[3][c], c = 0..2
    if (b >= 3) return false;
    J.C = Wb(3, b, h); J.A = Wb(3, 3, h); J.B = Qb(3, b, 0); J.ldb = LDQ;
    J.op = OP_SETNEG; J.nt = false; J.sa = sc.sW; J.sb = sc.sQ; J.inv = sc.iWQ;
    return true;
}

// the single copy of the tile code: tiles first, first + stride, ... of a phase
__device__ __noinline__ void run_tiles(int phase, int k, int first, int stride, int lane, float *S, float *W, float *Q,
                                       const Scales &sc) {
    Tile J;
    for (int idx = first; decode_tile(phase, k, idx, S, W, Q, sc, J); idx += stride) mma_tile(J, lane);
}

// Cholesky of the block in S (lower triangle, strict upper triangle zero, identity padding of a ragged block) and,
// when `publish`, its inverse in W with the fp16 planes written to hi / lo (zero right of the diagonal).  All 256
// threads.  On return S holds L (strict upper triangle zero); the caller synchronises the CTA and raises the flag
// with a release at gpu scope (which orders the plane stores of all threads before it).
__device__ __forceinline__ void factor_invert_block(float *__restrict__ S, float *__restrict__ W, float *__restrict__ Q,
                                                    float *__restrict__ ring, int nb, long long j0, int32_t *info,
                                                    bool publish, __half *hi, __half *lo, long long ld, const Scales &sc) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    float *rsv = ring + 32 * 32;
    int *progress = reinterpret_cast<int *>(rsv + 32);
    // warp 0: pivots; warp 5: inverse follower (another scheduler than warp 0's); shadow group: the other six
#ifdef GPG_CHAIN_G5
    constexpr int G = 5;
    const int g = warp <= 3 ? warp - 1 : warp - 3;
    const bool shadow = warp != 0 && warp != 4 && warp != 5;
#else
    constexpr int G = 6;
    const int g = warp <= 4 ? warp - 1 : warp - 2;
    const bool shadow = warp != 0 && warp != 5;
#endif
    for (int p = 0; p < 4; ++p) {
        const int c0 = 32 * p;
        if (t == 0) st_flag(progress, 0);
        __syncthreads();
        CHAIN_CLK(4 * p);
        if (warp == 0) {
            const int bad = pivot32(S, c0, lane, nb - c0, ring, rsv, progress);
            if (lane == 0 && bad && info) atomicCAS(info, 0, (int32_t)(j0 + c0 + bad));
        } else if (warp == 5) {
            follow32<true>(W, 0, c0, lane, ring, rsv, progress);
        } else if (shadow && p >= 1) {
            const int k = p - 1;
            SHADOW_CLK(4 * k);
            run_tiles(PH_SHADOW1, k, g, G, lane, S, W, Q, sc);
            SHADOW_CLK(4 * k + 1);
            if (k >= 1) group_barrier(1, 32 * G);        // (c) and the planes read the finished row block k of W
            SHADOW_CLK(4 * k + 2);
            run_tiles(PH_SHADOW2, k, g, G, lane, S, W, Q, sc);
            SHADOW_CLK(4 * k + 3);
            if (publish) publish_rows(W, k, 32 * g + lane, 32 * G, hi, lo, ld, j0, sc.sW);
            SHADOW_CLK(12 + k);
        }
        CHAIN_CLK(4 * p + 1);
        __syncthreads();
        CHAIN_CLK(4 * p + 2);
        if (p < 3) {
            run_tiles(PH_SOLVE, p, warp, 8, lane, S, W, Q, sc);
            __syncthreads();
            CHAIN_CLK(4 * p + 3);
            run_tiles(PH_UPDATE, p, warp, 8, lane, S, W, Q, sc);
        }
    }
    __syncthreads();
    CHAIN_CLK(16);
    if (publish) {
        run_tiles(PH_TAIL, 3, warp, 8, lane, S, W, Q, sc);
        __syncthreads();
        CHAIN_CLK(17);
        publish_rows(W, 3, t, 256, hi, lo, ld, j0, sc.sW);
        CHAIN_CLK(18);
    }
}

}  // namespace cchain
