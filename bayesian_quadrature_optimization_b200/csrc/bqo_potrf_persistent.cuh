// Persistent batched Cholesky (replaces ~110 launches per call by one).
//
// One cooperative launch of (SMs x 2) CTAs factors all S matrices.  Work = one task per 64x64
// tile (i, j), i >= j, left-looking:
//     T      = A(i,j) - sum_{k<j} L(i,k) L(j,k)^T            DMMA tile product, K = 64 j
//     i == j : L(j,j) = chol(T),  W(j) = L(j,j)^-1            64x64 in shared memory
//     i >  j : L(i,j) = T W(j)^T                              one more 64x64x64 DMMA product
// Tasks are handed out by an atomic counter in an order that is a topological order of the
// dependency graph (step j: tile (j+1,j), then the next diagonal tile, then the rest of column j,
// all matrices interleaved), and every CTA is resident (cooperative launch), so a task only ever
// waits for tasks that are already running: no deadlock.  Dependencies are tracked with ONE
// counter per tile row: rowdone[s][r] = number of final tiles in row r -- the tiles of a row
// become final in column order, because tile (r, k) needs the tiles (r, k') with k' < k.  A task consumes its K range
// progressively: whatever prefix of the two operand rows is final is multiplied while the rest
// is still being produced, so the long products are off the critical path, which per column is
// only: last 64-wide chunk + 64x64 factor and inverse + one 64x64x64 product.
// Every wait is bounded (spin limit -> abort flag), so a logic error cannot hang the GPU.
#pragma once
#include "bqo_gemm.cuh"

#define PP_NB 64
#define PP_LD 68            // [64][68] doubles: conflict-free m8n8k4 fragment loads (68 mod 16 = 4)
#define PP_STAGES 2
#define PP_THREADS 128
#define PP_SPIN_LIMIT (1 << 24)
#define PP_SMEM_BYTES ((size_t)(2 * BQO_OPER_BUF_DOUBLES(PP_STAGES) + 2 * PP_NB * PP_LD) * sizeof(double))

__device__ __forceinline__ int pp_ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pp_st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ double pp_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double hd = 0.5 * d;
    double e = fma(-hd * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hd * y, y, 0.5);
    y = fma(y, e, y);
    return y;
}

// ---- 64 x 64 factor + inverse on the critical path --------------------------------------------
// (first version: 8-wide blocks, 24 barriers, thread-per-element loops: 51 us per block measured
// with bqo_potrf_profile; this one: 8 us.)
//
// A 32 x 32 block is factored by ONE warp with one matrix row per lane in registers.  Per pivot
// the lanes put their entry of the scaled column into a 32-double shared-memory line and read it
// back with broadcast 128-bit loads (two entries per load): 31 - p DFMAs apply it to the trailing
// columns (no predication: the strict upper triangle collects garbage that nothing reads).  The
// same warp then inverts the factor, lane c solving L w = e_c by forward substitution with the
// rows of L read from the stored factor by broadcast loads (four partial sums per entry to
// shorten the dependent chain).
// The code is straight-line and runs once per diagonal task, so it is bound by instruction fetch:
// the first version broadcast with __shfl_sync (2 SHFL + WARPSYNC + ENDCOLLECTIVE + moves per
// double, 6400 instructions per 32 x 32 block, 11 us); broadcasting through shared memory needs
// half a load per entry.
// `col`: 2 x 32 doubles of scratch (16-byte aligned), double-buffered by pivot parity.
// Returns 0 or base + (1-based index of the first non-positive pivot); warp-uniform.
__device__ __noinline__ int pp_chol_inv_32_warp(double* T, double* Wm, double* col, int base) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    double a[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) a[q] = T[lane * PP_LD + q];
    int bad = 0;
#pragma unroll
    for (int p = 0; p < 32; ++p) {
        double* cb = col + (p & 1) * 32;
        if (lane == p) cb[p] = a[p];
        __syncwarp(full);
        const double d = cb[p];
        if (!(d > 0.0) && bad == 0) bad = base + p + 1;   // also NaN
        const double rs = pp_rsqrt(d > 0.0 ? d : 1.0);
        double lp = a[p] * rs;                              // L[lane][p] for lane > p
        if (lane == p) {
            const double g = d * rs;                        // sqrt(d) with one correction step
            lp = fma(fma(-g, g, d), 0.5 * rs, g);
        }
        a[p] = lp;
        if (p < 31) {
            // lane > p publishes L[lane][p]; lane p's slot still holds d, which nothing reads again
            if (lane > p) cb[lane] = lp;
            __syncwarp(full);
            if ((p & 1) == 0) a[p + 1] = fma(-lp, cb[p + 1], a[p + 1]);   // odd q = p + 1 first
#pragma unroll
            for (int q = (p + 2) & ~1; q < 32; q += 2) {
                const double2 l2 = *reinterpret_cast<const double2*>(cb + q);   // L[q][p], L[q+1][p]
                a[q] = fma(-lp, l2.x, a[q]);
                a[q + 1] = fma(-lp, l2.y, a[q + 1]);
            }
        }
    }
    if (bad) return bad;
#pragma unroll
    for (int q = 0; q < 32; ++q)
        if (q <= lane) T[lane * PP_LD + q] = a[q];
    double dg = 1.0;
#pragma unroll
    for (int q = 0; q < 32; ++q)
        if (q == lane) dg = a[q];
    col[lane] = 1.0 / dg;                                   // 1 / L[i][i], read back by broadcast
    __syncwarp(full);
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        double acc[4] = {(i == lane) ? 1.0 : 0.0, 0.0, 0.0, 0.0};
        const double* Li = T + i * PP_LD;                   // row i of L (16-byte aligned)
#pragma unroll
        for (int k = 0; k + 1 < i; k += 2) {
            const double2 l2 = *reinterpret_cast<const double2*>(Li + k);   // L[i][k], L[i][k+1]
            acc[k & 3] = fma(-l2.x, x[k], acc[k & 3]);      // x[k] = 0 for k < lane
            acc[(k + 1) & 3] = fma(-l2.y, x[k + 1], acc[(k + 1) & 3]);
        }
        if (i & 1) acc[(i - 1) & 3] = fma(-Li[i - 1], x[i - 1], acc[(i - 1) & 3]);
        x[i] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) * col[i];
        Wm[i * PP_LD + lane] = x[i];                        // W[i][c], zero above the diagonal
    }
    return 0;
}

// One warp's 16 x 16 block of a 32 x 32 x 32 product from shared memory on DMMA m8n8k4:
// acc[a][b] += sum_k A(r0 + 8a + g, k) B(c0 + 8b + g', k).  A(row, k) at A[row * lda + k] when AK
// (k contiguous) else A[k * lda + row]; likewise B(col, k).
template <bool AK, bool BK>
__device__ __forceinline__ void pp_mma_32(double (&acc)[2][2][2], const double* __restrict__ A, int lda,
                                          const double* __restrict__ B, int ldb, int r0, int c0) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int k0 = 0; k0 < 32; k0 += 4) {
        double af[2], bf[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            const int row = r0 + 8 * x + g, col = c0 + 8 * x + g;
            af[x] = AK ? A[row * lda + k0 + t] : A[(k0 + t) * lda + row];
            bf[x] = BK ? B[col * ldb + k0 + t] : B[(k0 + t) * ldb + col];
        }
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int y = 0; y < 2; ++y) bqo_dmma_884(acc[x][y][0], acc[x][y][1], af[x], bf[y]);
    }
}

template <typename F>
__device__ __forceinline__ void pp_acc32_foreach(double (&acc)[2][2][2], int r0, int c0, F f) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 2; ++y) {
            f(r0 + 8 * x + g, c0 + 8 * y + 2 * t, acc[x][y][0]);
            f(r0 + 8 * x + g, c0 + 8 * y + 2 * t + 1, acc[x][y][1]);
        }
}

// Cholesky factor and inverse of the 64 x 64 block T (shared memory, stride PP_LD, lower triangle
// used), 128 threads.  On return T holds L (lower; strict upper untouched) and Wm = L^-1 (lower,
// strict upper zero).  Rows / columns beyond the matrix must hold the identity.  Returns 0 or the
// 1-based index of the first non-positive pivot (to every thread).
//   [ A11     ]     L11 = chol(A11), W11 = L11^-1          warp 0
//   [ A21 A22 ]     L21 = A21 W11^T ; A22 -= L21 L21^T      DMMA, 4 warps
//                   L22 = chol(A22), W22 = L22^-1          warp 0   | P = L21 W11   warps 1-3
//                   W21 = -W22 P                            DMMA, 4 warps
__device__ int pp_potrf_inv_64(double* T, double* Wm, int* flag_sm) {
    __shared__ __align__(16) double pp_col[64];
    const int t = threadIdx.x, warp = t >> 5;
    double* T21 = T + 32 * PP_LD;
    double* T22 = T21 + 32;
    double* W21 = Wm + 32 * PP_LD;
    double* W22 = W21 + 32;
    for (int e = t; e < 32 * 32; e += PP_THREADS) Wm[(e >> 5) * PP_LD + 32 + (e & 31)] = 0.0;
    if (warp == 0) {
        const int bad = pp_chol_inv_32_warp(T, Wm, pp_col, 0);
        if ((t & 31) == 0) *flag_sm = bad;
    }
    __syncthreads();
    if (*flag_sm) return *flag_sm;
    const int r0 = (warp >> 1) * 16, c0 = (warp & 1) * 16;
    {
        double acc[2][2][2] = {};
        pp_mma_32<true, true>(acc, T21, PP_LD, Wm, PP_LD, r0, c0);          // L21 = A21 W11^T
        __syncthreads();
        pp_acc32_foreach(acc, r0, c0, [&](int r, int c, double v) { T21[r * PP_LD + c] = v; });
        __syncthreads();
        double upd[2][2][2] = {};
        pp_mma_32<true, true>(upd, T21, PP_LD, T21, PP_LD, r0, c0);         // L21 L21^T
        pp_acc32_foreach(upd, r0, c0, [&](int r, int c, double v) { T22[r * PP_LD + c] -= v; });
    }
    __syncthreads();
    if (warp == 0) {
        const int bad = pp_chol_inv_32_warp(T22, W22, pp_col, 32);
        if ((t & 31) == 0) *flag_sm = bad;
    } else {
        // P = L21 W11 (into W21's place), the four 16 x 16 blocks over warps 1..3
        for (int blk = warp - 1; blk < 4; blk += 3) {
            const int pr = (blk >> 1) * 16, pc = (blk & 1) * 16;
            double acc[2][2][2] = {};
            pp_mma_32<true, false>(acc, T21, PP_LD, Wm, PP_LD, pr, pc);
            pp_acc32_foreach(acc, pr, pc, [&](int r, int c, double v) { W21[r * PP_LD + c] = v; });
        }
    }
    __syncthreads();
    if (*flag_sm) return *flag_sm;
    {
        double acc[2][2][2] = {};
        pp_mma_32<true, false>(acc, W22, PP_LD, W21, PP_LD, r0, c0);        // W22 P
        __syncthreads();
        pp_acc32_foreach(acc, r0, c0, [&](int r, int c, double v) { W21[r * PP_LD + c] = -v; });
    }
    __syncthreads();
    return 0;
}

// acc += A[64 x 64] * B with both tiles in shared memory (stride PP_LD): A[row][k]; B[col][k] when BK
// (the product A B^T of two row-major tiles) or B[k][col] (the plain product A B).
template <bool BK = true>
__device__ __forceinline__ void pp_mma_smem(BqoAcc& acc, const double* __restrict__ Ts,
                                            const double* __restrict__ Ws) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
#pragma unroll 4
    for (int k0 = 0; k0 < PP_NB; k0 += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = Ts[(wr + 8 * a + g) * PP_LD + k0 + t];
#pragma unroll
        for (int b = 0; b < 4; ++b)
            bf[b] = BK ? Ws[(wc + 8 * b + g) * PP_LD + k0 + t] : Ws[(k0 + t) * PP_LD + wc + 8 * b + g];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) bqo_dmma_884(acc.c[a][b][0], acc.c[a][b][1], af[a], bf[b]);
    }
}

struct PotrfPersistArgs {
    double* A;          // [S][n][n], factored in place
    int n, S, nb;
    int* info;          // [S]
    int* rowdone;       // [S][nb]  final tiles per tile row
    int* failed;        // [S]      matrix s stopped (not positive definite)
    int* counters;      // [0] next task, [1] abort (spin limit reached: a bug, never expected)
    double* Wd;         // [S][nb][64*64] inverses of the diagonal blocks
    double* Wmat;       // optional [S][n][n]: W = L^-1 (lower tiles), produced by extra tasks
    int* wdone;         // [S][nb]  final tiles of W per tile column (rows j, j+1, ...)
    unsigned long long* prof;  // optional [nb][2][8] %globaltimer stamps of matrix 0's diagonal task
                               // (index 0) and of tile (j+1, j) (index 1) of every column j
};

__device__ __forceinline__ unsigned long long pp_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define PP_STAMP(slot)                                                                   \
    do {                                                                                 \
        if (a.prof && s == 0 && t == 0 && (i == j || i == j + 1))                         \
            a.prof[((size_t)j * 2 + (i == j ? 0 : 1)) * 8 + (slot)] = pp_now();           \
    } while (0)

// Task order (a topological order of the dependency graph; priority = position):
//   diag(0); tile(1,0); diag(1);                                     all matrices interleaved
//   step j = 0 .. nb-2:  tile (j+2, j)                                the chain's input from column j
//                        tile (j+2, j+1), diag (j+2)                  the critical chain, issued one
//                                                                     step AHEAD of its column's bulk
//                        tiles (i >= j+3, j)                          the bulk of column j
//                        W(j, j'), j' < j                             (inverse requested) background
//   W(nb-1, .)                                                        (inverse requested)
// The chain tasks are grabbed early and wait (at most 2 S CTAs do), so the bulk and the inverse
// tiles never stand between two links of the chain.  kind: 0 = factor tile, 1 = inverse tile.
__device__ __forceinline__ void pp_decode(int tsk, int S, int nb, bool with_w, int& kind, int& s,
                                          int& i, int& j) {
    kind = 0;
    if (tsk < 3 * S) {
        s = tsk % S;
        const int w = tsk / S;       // 0: diag(0), 1: tile(1,0), 2: diag(1)
        i = (w == 0) ? 0 : 1;
        j = (w == 2) ? 1 : 0;
        return;
    }
    int rem = tsk - 3 * S;
    for (int jj = 0; jj <= nb - 2; ++jj) {
        const bool chain = jj + 2 <= nb - 1;
        const int c1a = chain ? S : 0;               // tile (jj+2, jj): the chain's own input
        if (rem < c1a) {
            s = rem;
            i = jj + 2;
            j = jj;
            return;
        }
        rem -= c1a;
        const int c2 = chain ? 2 * S : 0;
        if (rem < c2) {
            s = rem % S;
            i = jj + 2;
            j = (rem < S) ? jj + 1 : jj + 2;
            return;
        }
        rem -= c2;
        const int c1b = chain ? S * (nb - jj - 3) : 0;   // the rest of column jj
        if (rem < c1b) {
            s = rem % S;
            i = jj + 3 + rem / S;
            j = jj;
            return;
        }
        rem -= c1b;
        const int c3 = with_w ? S * jj : 0;
        if (rem < c3) {
            kind = 1;
            s = rem % S;
            i = jj;
            j = rem / S;
            return;
        }
        rem -= c3;
    }
    kind = 1;
    s = rem % S;
    i = nb - 1;
    j = rem / S;
}

// Bounded wait until *p >= target (thread 0 polls, the result is broadcast): false when the
// matrix was stopped or the spin limit was hit.
__device__ __forceinline__ bool pp_wait_ge(const int* p, int target, const int* failed_s, int* abort_flag,
                                           int* sh_word) {
    if (threadIdx.x == 0) {
        int spins = 0, ok = 1;
        while (pp_ld_acquire(p) < target) {
            if (pp_ld_acquire(failed_s) || pp_ld_acquire(abort_flag)) {
                ok = 0;
                break;
            }
            if (++spins > PP_SPIN_LIMIT) {
                atomicExch(abort_flag, 1);
                ok = 0;
                break;
            }
            __nanosleep(40);
        }
        *sh_word = ok;
    }
    __syncthreads();
    const int ok = *sh_word;
    __syncthreads();
    return ok != 0;
}

__global__ void __launch_bounds__(PP_THREADS, 2) potrf_persistent_kernel(PotrfPersistArgs a) {
    extern __shared__ __align__(16) double dyn_smem[];
    double* Asm = dyn_smem;
    double* Bsm = Asm + BQO_OPER_BUF_DOUBLES(PP_STAGES);
    double* Ts = Bsm + BQO_OPER_BUF_DOUBLES(PP_STAGES);
    double* Ws = Ts + PP_NB * PP_LD;
    __shared__ int sh_task, sh_avail, sh_flag;
    const int t = threadIdx.x;
    const int n = a.n, nb = a.nb, S = a.S;
    const bool with_w = a.Wmat != nullptr;
    const int total = with_w ? S * nb * nb : S * (nb * (nb + 1) / 2);
    for (;;) {
        if (t == 0) sh_task = atomicAdd(&a.counters[0], 1);
        __syncthreads();
        const int tsk = sh_task;
        if (tsk >= total) break;
        int kind, s, i, j;
        pp_decode(tsk, S, nb, with_w, kind, s, i, j);
        double* As = a.A + (int64_t)s * n * n;
        int* rd = a.rowdone + (int64_t)s * nb;
        const int i0 = i * PP_NB, j0 = j * PP_NB;
        const int ri = (n - i0 < PP_NB) ? n - i0 : PP_NB;
        const int rj = (n - j0 < PP_NB) ? n - j0 : PP_NB;
        if (kind == 1) {
            // ---- W(i,j) = -W(i,i) sum_{k=j}^{i-1} L(i,k) W(k,j),  i > j ------------------------------
            double* Wm_ = a.Wmat + (int64_t)s * n * n;
            int* wd_col = a.wdone + (int64_t)s * nb + j;
            if (!pp_wait_ge(rd + i, i + 1, a.failed + s, a.counters + 1, &sh_avail)) continue;
            // W(i,i) from the diagonal task, fetched while the product runs
            const double* wdi = a.Wd + ((int64_t)s * nb + i) * (PP_NB * PP_NB);
            for (int e = t; e < PP_NB * PP_NB; e += PP_THREADS)
                Ws[(e >> 6) * PP_LD + (e & 63)] = __ldcg(wdi + e);
            BqoAcc acc;
            acc.zero();
            int kdone = j;      // tile rows of column j consumed so far
            bool dead = false;
            while (kdone < i) {
                if (t == 0) {
                    int spins = 0, av = 0;
                    for (;;) {
                        av = j + pp_ld_acquire(wd_col);     // rows j .. av-1 of column j are final
                        if (av > kdone) break;
                        if (pp_ld_acquire(a.failed + s) || pp_ld_acquire(a.counters + 1)) {
                            av = -1;
                            break;
                        }
                        if (++spins > PP_SPIN_LIMIT) {
                            atomicExch(a.counters + 1, 1);
                            av = -1;
                            break;
                        }
                        __nanosleep(40);
                    }
                    sh_avail = av;
                }
                __syncthreads();
                int av = sh_avail;
                __syncthreads();
                if (av < 0) {
                    dead = true;
                    break;
                }
                if (av > i) av = i;
                bqo_gemm_tile<true, false, PP_STAGES>(
                    acc, As + (int64_t)i0 * n + (int64_t)kdone * PP_NB, n, ri,
                    Wm_ + (int64_t)kdone * PP_NB * n + j0, n, rj, (av - kdone) * PP_NB, Asm, Bsm);
                __syncthreads();
                kdone = av;
            }
            if (dead) continue;
            bqo_acc_foreach(acc, [&](int r, int c, double& v) { Ts[r * PP_LD + c] = v; });
            __syncthreads();
            BqoAcc x;
            x.zero();
            pp_mma_smem<false>(x, Ws, Ts);
            bqo_acc_foreach(x, [&](int r, int c, double& v) {
                if (r < ri && c < rj) __stcg(Wm_ + (int64_t)(i0 + r) * n + (j0 + c), -v);
            });
            __threadfence();
            __syncthreads();
            if (t == 0) pp_st_release(wd_col, i - j + 1);
            continue;
        }
        BqoAcc acc;
        acc.zero();
        bool dead = false;
        PP_STAMP(0);
        // the tile's own entries do not depend on anything: fetched now (coalesced 16-byte L2 loads,
        // identity padding outside the matrix), the product is subtracted from them afterwards
        for (int e = t; e < PP_NB * (PP_NB / 2); e += PP_THREADS) {
            const int r = e >> 5, c = (e & 31) * 2;
            double2 v = make_double2((r == c && i == j) ? 1.0 : 0.0, (r == c + 1 && i == j) ? 1.0 : 0.0);
            if (r < ri && c + 1 < rj) {
                v = __ldcg(reinterpret_cast<const double2*>(As + (int64_t)(i0 + r) * n + (j0 + c)));
            } else if (r < ri && c < rj) {
                v.x = __ldcg(As + (int64_t)(i0 + r) * n + (j0 + c));
            }
            Ts[r * PP_LD + c] = v.x;
            Ts[r * PP_LD + c + 1] = v.y;
        }
        // ---- progressive left-looking product over the final prefix of rows i and j ------------
        int kdone = 0;
        while (kdone < j) {
            if (t == 0) {
                int spins = 0, av = 0;
                for (;;) {
                    const int a1 = pp_ld_acquire(rd + i);
                    const int a2 = (i == j) ? a1 : pp_ld_acquire(rd + j);
                    av = a1 < a2 ? a1 : a2;
                    if (av > kdone) break;
                    if (pp_ld_acquire(a.failed + s) || pp_ld_acquire(a.counters + 1)) {
                        av = -1;
                        break;
                    }
                    if (++spins > PP_SPIN_LIMIT) {
                        atomicExch(a.counters + 1, 1);
                        av = -1;
                        break;
                    }
                    __nanosleep(40);
                }
                sh_avail = av;
            }
            __syncthreads();
            int av = sh_avail;
            __syncthreads();
            if (av < 0) {
                dead = true;
                break;
            }
            if (av > j) av = j;
            bqo_gemm_tile<true, true, PP_STAGES>(acc, As + (int64_t)i0 * n + (int64_t)kdone * PP_NB, n, ri,
                                                 As + (int64_t)j0 * n + (int64_t)kdone * PP_NB, n, rj,
                                                 (av - kdone) * PP_NB, Asm, Bsm);
            __syncthreads();
            kdone = av;
        }
        if (dead) continue;
        PP_STAMP(1);
        // ---- T = A(i,j) - acc in shared memory ---------------------------------------------------
        __syncthreads();
        bqo_acc_foreach(acc, [&](int r, int c, double& v) { Ts[r * PP_LD + c] -= v; });
        __syncthreads();
        PP_STAMP(2);
        if (i == j) {
            const int bad = pp_potrf_inv_64(Ts, Ws, &sh_flag);
            PP_STAMP(3);
            if (bad) {
                if (t == 0) {
                    a.info[s] = j0 + bad;
                    __threadfence();
                    pp_st_release(a.failed + s, 1);
                }
                __syncthreads();
                continue;
            }
            double* wd = a.Wd + ((int64_t)s * nb + j) * (PP_NB * PP_NB);
            for (int e = t; e < PP_NB * PP_NB; e += PP_THREADS) {
                const int r = e >> 6, c = e & 63;
                __stcg(wd + e, Ws[r * PP_LD + c]);
                if (r < ri && c < rj)
                    __stcg(As + (int64_t)(i0 + r) * n + (j0 + c), (c <= r) ? Ts[r * PP_LD + c] : 0.0);
            }
            if (with_w) {
                double* Wm_ = a.Wmat + (int64_t)s * n * n;
                for (int e = t; e < PP_NB * PP_NB; e += PP_THREADS) {
                    const int r = e >> 6, c = e & 63;
                    if (r < ri && c < rj) __stcg(Wm_ + (int64_t)(i0 + r) * n + (j0 + c), Ws[r * PP_LD + c]);
                }
            }
            __threadfence();
            __syncthreads();
            if (t == 0) {
                pp_st_release(rd + j, j + 1);
                if (with_w) pp_st_release(a.wdone + (int64_t)s * nb + j, 1);
            }
            PP_STAMP(4);
        } else {
            // wait for the diagonal block of column j, then L(i,j) = T W(j)^T
            if (t == 0) {
                int spins = 0, ok = 1;
                while (pp_ld_acquire(rd + j) < j + 1) {
                    if (pp_ld_acquire(a.failed + s) || pp_ld_acquire(a.counters + 1)) {
                        ok = 0;
                        break;
                    }
                    if (++spins > PP_SPIN_LIMIT) {
                        atomicExch(a.counters + 1, 1);
                        ok = 0;
                        break;
                    }
                    __nanosleep(40);
                }
                sh_avail = ok;
            }
            __syncthreads();
            const int ok = sh_avail;
            __syncthreads();
            if (!ok) continue;
            PP_STAMP(3);
            const double* wd = a.Wd + ((int64_t)s * nb + j) * (PP_NB * PP_NB);
            for (int e = t; e < PP_NB * PP_NB; e += PP_THREADS)
                Ws[(e >> 6) * PP_LD + (e & 63)] = __ldcg(wd + e);
            __syncthreads();
            PP_STAMP(4);
            BqoAcc x;
            x.zero();
            pp_mma_smem<true>(x, Ts, Ws);
            PP_STAMP(5);
            bqo_acc_foreach(x, [&](int r, int c, double& v) {
                if (r < ri && c < rj) {
                    __stcg(As + (int64_t)(i0 + r) * n + (j0 + c), v);
                    __stcg(As + (int64_t)(j0 + c) * n + (i0 + r), 0.0);   // strict upper triangle
                }
            });
            __threadfence();
            __syncthreads();
            if (t == 0) pp_st_release(rd + i, j + 1);
            PP_STAMP(6);
        }
    }
}
