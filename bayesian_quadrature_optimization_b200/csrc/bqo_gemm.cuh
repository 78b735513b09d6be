// FP64 tensor-core (DMMA m8n8k4) tile product used by the Cholesky trailing update and the
// triangular solves.  On sm_100a every f64 mma.sync shape lowers to DMMA.8x8x4 (checked with
// cuobjdump), so m8n8k4 is issued directly.  There is no tcgen05/TMEM path for FP64.
//
// A CTA of 128 threads (4 warps as 2x2) accumulates a 64x64 tile acc += A[64xK] * B[Kx64];
// each warp owns 32x32 = 4x4 m8n8 fragments (2 doubles per thread each).  Operands reach shared
// memory through a cp.async pipeline (bqo_gemm_tile); with long K the tile sustains 34.6 TFLOP/s
// = 0.93 of the measured DMMA peak (777-vector triangular solve, profiles/r01_stage_a.md).
#pragma once
#include "bqo_common.cuh"

#define BQO_TILE 64
#define BQO_BK 16
#define BQO_LDK (BQO_BK + 4)    // tile stored [row][k]: stride 20 doubles -> conflict-free frags
#define BQO_LDR (BQO_TILE + 4)  // tile stored [k][row]: stride 68 doubles -> conflict-free frags

__device__ __forceinline__ void bqo_dmma_884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct BqoAcc {
    double c[4][4][2];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) c[a][b][0] = c[a][b][1] = 0.0;
    }
};

// Shared memory for one operand tile (64 x 16), either layout.
#define BQO_OPER_TILE_DOUBLES (BQO_TILE * BQO_LDK > BQO_BK * BQO_LDR ? BQO_TILE * BQO_LDK : BQO_BK * BQO_LDR)

// Load a 64(row) x 16(k) operand tile.  Element (row, k) lives at
//   base[row * ld + k]  when KCONTIG (k contiguous in memory), stored as sm[row][k];
//   base[k * ld + row]  otherwise (row contiguous), stored as sm[k][row].
// rows_valid / k_valid bound the tile (zero fill outside).
template <bool KCONTIG>
__device__ __forceinline__ void bqo_load_oper(double* __restrict__ sm,
                                              const double* __restrict__ base, int64_t ld,
                                              int rows_valid, int k_valid) {
    const int t = threadIdx.x;  // 128 threads
    if (KCONTIG) {
        // 64 rows x 16 k: each thread loads 8 consecutive k of one row half
        int row = t >> 1, kh = (t & 1) * 8;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            int k = kh + kk;
            double v = 0.0;
            if (row < rows_valid && k < k_valid) v = base[(int64_t)row * ld + k];
            sm[row * BQO_LDK + k] = v;
        }
    } else {
        // 16 k x 64 rows: thread t loads k = t/8, rows (t%8)*8 .. +7
        int k = t >> 3, r0 = (t & 7) * 8;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            int row = r0 + rr;
            double v = 0.0;
            if (row < rows_valid && k < k_valid) v = base[(int64_t)k * ld + row];
            sm[k * BQO_LDR + row] = v;
        }
    }
}

template <bool KCONTIG>
__device__ __forceinline__ double bqo_frag(const double* __restrict__ sm, int row, int k) {
    return KCONTIG ? sm[row * BQO_LDK + k] : sm[k * BQO_LDR + row];
}

// acc += A_tile * B_tile^T-style product over one BK slab held in shared memory.
// A fragment: row = 8a + g, k = t;  B fragment: k = t, col = 8b + g  (g = lane/4, t = lane%4).
template <bool A_KCONTIG, bool B_KCONTIG>
__device__ __forceinline__ void bqo_mma_slab(BqoAcc& acc, const double* __restrict__ As,
                                             const double* __restrict__ Bs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
#pragma unroll
    for (int k0 = 0; k0 < BQO_BK; k0 += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = bqo_frag<A_KCONTIG>(As, wr + 8 * a + g, k0 + t);
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = bqo_frag<B_KCONTIG>(Bs, wc + 8 * b + g, k0 + t);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) bqo_dmma_884(acc.c[a][b][0], acc.c[a][b][1], af[a], bf[b]);
    }
}

// Register-staged variant of bqo_load_oper: global -> 8 registers now, registers -> shared later,
// so that the global loads of slab k+1 are in flight while the DMMAs of slab k execute.
template <bool KCONTIG>
__device__ __forceinline__ void bqo_fetch_oper(double (&r)[8], const double* __restrict__ base,
                                               int64_t ld, int rows_valid, int k_valid) {
    const int t = threadIdx.x;
    if (KCONTIG) {
        int row = t >> 1, kh = (t & 1) * 8;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            int k = kh + kk;
            r[kk] = (row < rows_valid && k < k_valid) ? base[(int64_t)row * ld + k] : 0.0;
        }
    } else {
        int k = t >> 3, r0 = (t & 7) * 8;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            int row = r0 + rr;
            r[rr] = (row < rows_valid && k < k_valid) ? base[(int64_t)k * ld + row] : 0.0;
        }
    }
}

template <bool KCONTIG>
__device__ __forceinline__ void bqo_commit_oper(double* __restrict__ sm, const double (&r)[8]) {
    const int t = threadIdx.x;
    if (KCONTIG) {
        int row = t >> 1, kh = (t & 1) * 8;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) sm[row * BQO_LDK + kh + kk] = r[kk];
    } else {
        int k = t >> 3, r0 = (t & 7) * 8;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) sm[k * BQO_LDR + r0 + rr] = r[rr];
    }
}

// ---- cp.async staging ----------------------------------------------------------------------
// One 64 x 16 operand tile per stage; every thread copies its 8 elements with 8-byte cp.async
// (zero fill outside the valid rows / k), the same element-to-thread map as bqo_fetch_oper.
__device__ __forceinline__ void bqo_cp_async8(double* smem_dst, const double* gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 8 : 0;  // src-size 0: the 8 destination bytes are zero filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(gsrc), "r"(bytes));
}

template <bool KCONTIG>
__device__ __forceinline__ void bqo_stage_oper(double* __restrict__ sm,
                                               const double* __restrict__ base, int64_t ld,
                                               int rows_valid, int k_valid) {
    const int t = threadIdx.x;
    if (KCONTIG) {
        const int row = t >> 1, kh = (t & 1) * 8;
        const bool rv = row < rows_valid;
        const double* src = base + (int64_t)(rv ? row : 0) * ld;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int k = kh + kk;
            const bool ok = rv && k < k_valid;
            bqo_cp_async8(sm + row * BQO_LDK + k, src + (ok ? k : 0), ok);
        }
    } else {
        const int k = t >> 3, r0 = (t & 7) * 8;
        const bool kv = k < k_valid;
        const double* src = base + (int64_t)(kv ? k : 0) * ld;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            const int row = r0 + rr;
            const bool ok = kv && row < rows_valid;
            bqo_cp_async8(sm + k * BQO_LDR + row, src + (ok ? row : 0), ok);
        }
    }
}

// 16-byte variant (base 16-byte aligned, ld even): four cp.async per thread and operand, each warp
// instruction covers whole contiguous lines in global AND shared memory (the 8-byte map above
// writes shared memory with 4-way bank conflicts in the row-contiguous layout and kept the LSU
// at 88 % of its wavefront peak).  Out-of-range pairs fall back to two 8-byte copies.
__device__ __forceinline__ void bqo_cp_async16(double* smem_dst, const double* gsrc) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gsrc));
}

template <bool KCONTIG>
__device__ __forceinline__ void bqo_stage_oper16(double* __restrict__ sm,
                                                 const double* __restrict__ base, int64_t ld,
                                                 int rows_valid, int k_valid) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = t + 128 * i;  // 16-byte chunk index, 512 per tile
        if (KCONTIG) {
            const int row = c >> 3, k = (c & 7) * 2;  // 8 chunks per row of 16 k
            double* dst = sm + row * BQO_LDK + k;
            const double* src = base + (int64_t)row * ld + k;
            if (row < rows_valid && k + 1 < k_valid) {
                bqo_cp_async16(dst, src);
            } else {
                bqo_cp_async8(dst, row < rows_valid ? src : base, row < rows_valid && k < k_valid);
                bqo_cp_async8(dst + 1, base, false);
            }
        } else {
            const int k = c >> 5, row = (c & 31) * 2;  // 32 chunks per k line of 64 rows
            double* dst = sm + k * BQO_LDR + row;
            const double* src = base + (int64_t)k * ld + row;
            if (k < k_valid && row + 1 < rows_valid) {
                bqo_cp_async16(dst, src);
            } else {
                bqo_cp_async8(dst, k < k_valid ? src : base, k < k_valid && row < rows_valid);
                bqo_cp_async8(dst + 1, base, false);
            }
        }
    }
}

#define BQO_GEMM_STAGES 3
#define BQO_OPER_BUF_DOUBLES(ST) ((ST) * BQO_OPER_TILE_DOUBLES)

// Full K loop: acc += sum_k A(row,k) B(col,k).  As/Bs: shared operand buffers of
// BQO_OPER_BUF_DOUBLES(STAGES) doubles each.  A(row,k) at A[row*lda + k] (A_KCONTIG) or
// A[k*lda + row]; likewise for B with its own leading dimension.
// STAGES-deep cp.async pipeline, one barrier per slab: the copies of slab i + STAGES - 1 are in
// flight while the DMMAs of slab i execute (the first version prefetched ONE slab into registers
// and ran the DMMA sub-pipe at 56 % with 0.18 eligible warps per cycle, long-scoreboard bound).
// All copies have landed and been consumed when the function returns; callers that reuse the
// buffers must still separate two calls with a __syncthreads().
template <bool A_KCONTIG, bool B_KCONTIG, int STAGES = BQO_GEMM_STAGES>
__device__ __forceinline__ void bqo_gemm_tile(BqoAcc& acc, const double* __restrict__ A,
                                              int64_t lda, int a_rows, const double* __restrict__ B,
                                              int64_t ldb, int b_rows, int K, double* As,
                                              double* Bs) {
    if (K <= 0) return;
    const int nslab = (K + BQO_BK - 1) / BQO_BK;
    // slab offsets are multiples of 16 doubles, so alignment of the base decides for all slabs
    const bool a16 = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((lda & 1) == 0);
    const bool b16 = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((ldb & 1) == 0);
    auto issue = [&](int slab) {
        if (slab < nslab) {
            const int k0 = slab * BQO_BK;
            const int kv = K - k0 < BQO_BK ? K - k0 : BQO_BK;
            const int st = slab % STAGES;
            double* as = As + st * BQO_OPER_TILE_DOUBLES;
            double* bs = Bs + st * BQO_OPER_TILE_DOUBLES;
            const double* ap = A_KCONTIG ? A + k0 : A + (int64_t)k0 * lda;
            const double* bp = B_KCONTIG ? B + k0 : B + (int64_t)k0 * ldb;
            if (a16) bqo_stage_oper16<A_KCONTIG>(as, ap, lda, a_rows, kv);
            else bqo_stage_oper<A_KCONTIG>(as, ap, lda, a_rows, kv);
            if (b16) bqo_stage_oper16<B_KCONTIG>(bs, bp, ldb, b_rows, kv);
            else bqo_stage_oper<B_KCONTIG>(bs, bp, ldb, b_rows, kv);
        }
        asm volatile("cp.async.commit_group;\n" ::);  // (possibly empty: keeps the group count)
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);
    for (int slab = 0; slab < nslab; ++slab) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(STAGES - 2));  // slab `slab` has landed
        __syncthreads();  // ... for every thread, and slab - 1's buffer is free again
        issue(slab + STAGES - 1);
        const int st = slab % STAGES;
        bqo_mma_slab<A_KCONTIG, B_KCONTIG>(acc, As + st * BQO_OPER_TILE_DOUBLES,
                                           Bs + st * BQO_OPER_TILE_DOUBLES);
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
}

// Visit the accumulator entries of this thread: f(row, col, value_ref) with tile-local row/col.
template <typename F>
__device__ __forceinline__ void bqo_acc_foreach(BqoAcc& acc, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            f(wr + 8 * a + g, wc + 8 * b + 2 * t, acc.c[a][b][0]);
            f(wr + 8 * a + g, wc + 8 * b + 2 * t + 1, acc.c[a][b][1]);
        }
}
