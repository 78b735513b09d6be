// FP64 tensor-core (DMMA m8n8k4) tile product used by the Cholesky trailing update and the
// triangular solves.  On sm_100a every f64 mma.sync shape lowers to DMMA.8x8x4 (checked with
// cuobjdump), so m8n8k4 is issued directly.  There is no tcgen05/TMEM path for FP64.
//
// A CTA of 128 threads (4 warps as 2x2) accumulates a 64x64 tile acc += A[64xK] * B[Kx64];
// each warp owns 32x32 = 4x4 m8n8 fragments (2 doubles per thread each).
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

// Full K loop: acc += sum_k A(row,k) B(col,k).  As/Bs: shared operand buffers of
// BQO_OPER_TILE_DOUBLES doubles each.  A(row,k) at A[row*lda + k] (A_KCONTIG) or A[k*lda + row];
// likewise for B with its own leading dimension.  Software pipelined: slab k+1 is fetched into
// registers before the DMMAs of slab k are issued.
template <bool A_KCONTIG, bool B_KCONTIG>
__device__ __forceinline__ void bqo_gemm_tile(BqoAcc& acc, const double* __restrict__ A,
                                              int64_t lda, int a_rows, const double* __restrict__ B,
                                              int64_t ldb, int b_rows, int K, double* As,
                                              double* Bs) {
    if (K <= 0) return;
    double ra[8], rb[8];
    {
        int kv = K < BQO_BK ? K : BQO_BK;
        bqo_fetch_oper<A_KCONTIG>(ra, A, lda, a_rows, kv);
        bqo_fetch_oper<B_KCONTIG>(rb, B, ldb, b_rows, kv);
    }
    for (int k0 = 0; k0 < K; k0 += BQO_BK) {
        __syncthreads();  // previous slab's fragments have been read
        bqo_commit_oper<A_KCONTIG>(As, ra);
        bqo_commit_oper<B_KCONTIG>(Bs, rb);
        __syncthreads();
        const int kn = k0 + BQO_BK;
        if (kn < K) {
            int kv = K - kn < BQO_BK ? K - kn : BQO_BK;
            bqo_fetch_oper<A_KCONTIG>(ra, A_KCONTIG ? A + kn : A + (int64_t)kn * lda, lda, a_rows, kv);
            bqo_fetch_oper<B_KCONTIG>(rb, B_KCONTIG ? B + kn : B + (int64_t)kn * ldb, ldb, b_rows, kv);
        }
        bqo_mma_slab<A_KCONTIG, B_KCONTIG>(acc, As, Bs);
    }
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
