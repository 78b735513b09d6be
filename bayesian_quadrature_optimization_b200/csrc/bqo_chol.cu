// Batched FP64 Cholesky, triangular solves, alpha and log marginal likelihood
// (SURVEY 8a rows a10, a11, a12, a17).
//
//   potrf: right-looking blocked factorisation, NB = 64.  Per block column two launches:
//     (1) panel kernel: every CTA factors the 64x64 diagonal block in shared memory
//         (redundantly -- it is tiny -- so no inter-CTA dependency exists) and solves its own 64
//         rows of the panel against it; (2) trailing update A22 -= L21 L21^T on DMMA tiles.
//   potrs: RHS-major layout Bt[rhs][n].  A CTA owns 64 right-hand sides and walks the block
//     columns itself (left-looking): tile = B_tile - Y[:, :k0] L[kb, :k0]^T on DMMA, then a
//     64x64 substitution in shared memory.  No cross-CTA dependency, one launch per direction.
#include "bqo_gemm.cuh"
#include "bqo_potrf_persistent.cuh"
#include <cstdlib>
#include <vector>

#define NB BQO_TILE
#define LDT (NB + 1)

// ---------------------------------------------------------------------------------------------
// 1/sqrt(d) to full precision: MUFU.RSQ64H seed (2^-22) and two Newton steps.
__device__ __forceinline__ double bqo_rsqrt_full(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double hd = 0.5 * d;
    double e = fma(-hd * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hd * y, y, 0.5);
    y = fma(y, e, y);
    return y;
}

// Factor the NB x NB block held in shared memory `Ls` (stride LDT), lower triangle, in place.
// 128 threads.  Returns (to all threads) 0 or the 1-based index of the first non-positive pivot.
// Rows / columns >= nv must hold the identity (the loaders do that), so the loops need no tails.
//
// Blocked right-looking, 8 block steps of width 8 (a column-by-column version needs 64 x 2
// barriers and a latency-bound dot product per column: 56 us per block on B200):
//   (1) the 8x8 diagonal block: lanes 0..7 of warp 0 hold one row each in registers; per pivot
//       1/sqrt by MUFU.RSQ64H + Newton, the pivot column goes through shared memory;
//   (2) the panel below it: one thread per row, 8 right-looking steps in registers;
//   (3) the trailing update A[i][j] -= sum_p L[i][c0+p] L[j][c0+p], j <= i: row i's eight panel
//       values in registers, the (row, column-group) pairs spread over all 128 threads.
// `rdiag` (NB doubles of shared memory) receives 1 / l_kk.
#define PBS 8
__device__ int potrf_block_smem(double* Ls, int* flag_sm, double* rdiag) {
    __shared__ double colbuf[PBS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) *flag_sm = 0;
    __syncthreads();
    for (int c0 = 0; c0 < NB; c0 += PBS) {
        // ---- (1) diagonal 8x8 block ----------------------------------------------------------
        if (warp == 0) {
            double a[PBS];
            const int row = c0 + (lane & (PBS - 1));
#pragma unroll
            for (int q = 0; q < PBS; ++q) a[q] = Ls[row * LDT + c0 + q];
            int bad = 0;
#pragma unroll
            for (int p = 0; p < PBS; ++p) {
                const double d = __shfl_sync(0xffffffffu, a[p], p);
                if (!(d > 0.0) && bad == 0) bad = c0 + p + 1;  // also catches NaN
                const double r = bqo_rsqrt_full(d > 0.0 ? d : 1.0);
                if (lane == p) {
                    double g = d * r;                     // sqrt(d), one correction step
                    a[p] = fma(fma(-g, g, d), 0.5 * r, g);
                    rdiag[c0 + p] = r;
                } else if (lane > p) {
                    a[p] *= r;
                }
                if (lane < PBS) colbuf[lane] = a[p];
                __syncwarp();
#pragma unroll
                for (int q = p + 1; q < PBS; ++q)
                    if (lane >= q) a[q] = fma(-a[p], colbuf[q], a[q]);
                __syncwarp();
            }
            if (lane < PBS) {
#pragma unroll
                for (int q = 0; q < PBS; ++q)
                    if (q <= lane) Ls[row * LDT + c0 + q] = a[q];
            }
            if (bad && lane == 0) *flag_sm = bad;
        }
        __syncthreads();
        if (*flag_sm) return *flag_sm;
        const int m = NB - c0 - PBS;  // rows below the diagonal block
        if (m <= 0) break;
        // ---- (2) panel: rows c0+8 .. 63, x L_d^T = a -------------------------------------------
        if (t < m) {
            double x[PBS];
            double* ar = Ls + (c0 + PBS + t) * LDT + c0;
#pragma unroll
            for (int q = 0; q < PBS; ++q) x[q] = ar[q];
#pragma unroll
            for (int p = 0; p < PBS; ++p) {
                x[p] *= rdiag[c0 + p];
#pragma unroll
                for (int q = p + 1; q < PBS; ++q)
                    x[q] = fma(-x[p], Ls[(c0 + q) * LDT + c0 + p], x[q]);
            }
#pragma unroll
            for (int q = 0; q < PBS; ++q) ar[q] = x[q];
        }
        __syncthreads();
        // ---- (3) trailing update ----------------------------------------------------------------
        {
            const int groups = 128 / m;          // column groups per row (>= 2)
            const int ri = t % m, g = t / m;
            if (g < groups) {
                const int i = c0 + PBS + ri;
                double li[PBS];
#pragma unroll
                for (int q = 0; q < PBS; ++q) li[q] = Ls[i * LDT + c0 + q];
                for (int j = c0 + PBS + g; j <= i; j += groups) {
                    const double* lj = Ls + j * LDT + c0;
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < PBS; ++q) acc = fma(li[q], lj[q], acc);
                    Ls[i * LDT + j] -= acc;
                }
            }
        }
        __syncthreads();
    }
    return 0;
}

// Panel kernel for block column k0: factor the 64x64 diagonal block, then solve the rows below
// it, X L11^T = A21, one thread per row with L11 broadcast from shared memory.
//   MODE 0 (fused): grid = (row tiles from the diagonal tile down, S).  Every CTA factors the
//     diagonal block itself -- redundant, but it removes a launch and a global synchronisation;
//     this is the latency-optimal form (45 us per block column) and is used while the grid fits
//     the machine in about one wave.
//   MODE 1 + MODE 2 (split): one CTA per matrix factors the block and stores it in
//     `diag_scratch`; a second launch solves the row tiles against the stored factor.  Used when
//     the fused grid would be several waves of redundant factorisations (20 matrices of
//     n = 2048: 620 CTAs, 102 us fused).
// The diagonal tile's factor goes to `diag_scratch` instead of in place in MODE 0 because the
// other CTAs of that launch still read the unfactored block from A; potrf_finish_kernel copies
// it back.
// (Two right-looking rewrites of the row solve were measured and dropped: rows in registers with
// the factor fully unrolled too -- 16 000 straight-line instructions, instruction-fetch bound;
// tiles in shared memory with rolled loops -- shared-memory latency and bank conflicts.)
template <int MODE>
__global__ void __launch_bounds__(128) potrf_panel_kernel(double* __restrict__ A, int n, int k0,
                                                          int* __restrict__ info,
                                                          double* __restrict__ diag_scratch) {
    extern __shared__ __align__(16) double dyn_smem[];
    double* Ls = dyn_smem;
    __shared__ int flag;
    __shared__ double rdiag[NB];
    const int s = blockIdx.y;
    if (info[s] != 0) return;  // this matrix already failed
    double* As = A + (int64_t)s * n * n;
    const int nv = (n - k0 < NB) ? n - k0 : NB;
    const int nb = (n + NB - 1) / NB;
    const int t = threadIdx.x;
    const int tile = (MODE == 2) ? blockIdx.x + 1 : blockIdx.x;  // 0 = the diagonal tile
    double* scratch = diag_scratch + ((int64_t)s * nb + k0 / NB) * NB * NB;
    if (MODE == 2) {
        // the factor was stored by the MODE 1 launch
        for (int e = t; e < NB * NB; e += blockDim.x) {
            int r = e / NB, c = e % NB;
            double v = scratch[e];
            if (r == c) {
                if (r >= nv) v = 1.0;
                rdiag[r] = 1.0 / v;
            }
            Ls[r * LDT + c] = v;
        }
        __syncthreads();
    } else {
        for (int e = t; e < NB * NB; e += blockDim.x) {
            int r = e / NB, c = e % NB;
            double v = (r == c) ? 1.0 : 0.0;  // rows beyond the matrix: identity
            if (r < nv && c <= r) v = As[(int64_t)(k0 + r) * n + (k0 + c)];
            Ls[r * LDT + c] = v;
        }
        __syncthreads();
        int bad = potrf_block_smem(Ls, &flag, rdiag);
        if (bad) {
            if (tile == 0 && t == 0) info[s] = k0 + bad;
            return;
        }
        if (tile == 0) {
            for (int e = t; e < NB * NB; e += blockDim.x) {
                int r = e / NB, c = e % NB;
                scratch[e] = (r < nv && c <= r) ? Ls[r * LDT + c] : 0.0;
            }
            return;
        }
    }
    const int row = k0 + tile * NB + t;
    if (t < NB && row < n) {
        double x[NB];
        double* ap = As + (int64_t)row * n + k0;
#pragma unroll
        for (int j = 0; j < NB; ++j) x[j] = ap[j];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            double acc = x[j];
#pragma unroll
            for (int k = 0; k < j; ++k) acc = fma(-x[k], Ls[j * LDT + k], acc);
            x[j] = acc * rdiag[j];
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) ap[j] = x[j];
    }
}

#ifndef POTRF_SYRK_STAGES
#define POTRF_SYRK_STAGES 2  // K <= 256 here: a short pipeline and one more resident CTA per SM
#endif
#define SMEM_POTRF_SYRK ((size_t)(2 * BQO_OPER_BUF_DOUBLES(POTRF_SYRK_STAGES)) * sizeof(double))
// Trailing update A[i][j] -= sum_{k < K} L[i][kc+k] L[j][kc+k] on the lower tiles of the square
// that starts at row/column `base`; only the first `ncolblk` block columns of it are touched
// (1 = the strip that the second panel of a 128-wide step needs, all = the full update).
__global__ void __launch_bounds__(128) potrf_syrk_kernel(double* __restrict__ A, int n, int kc, int K,
                                                         int base, int ncolblk,
                                                         const int* __restrict__ info) {
    extern __shared__ __align__(16) double dyn_smem[];
    double* Asm = dyn_smem;
    double* Bsm = Asm + BQO_OPER_BUF_DOUBLES(POTRF_SYRK_STAGES);
    const int s = blockIdx.z;
    if (info[s] != 0) return;
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi || bj >= ncolblk) return;
    double* As = A + (int64_t)s * n * n;
    const int i0 = base + bi * NB, j0 = base + bj * NB;
    const int ri = (n - i0 < NB) ? n - i0 : NB;
    const int rj = (n - j0 < NB) ? n - j0 : NB;
    BqoAcc acc;
    acc.zero();
    bqo_gemm_tile<true, true, POTRF_SYRK_STAGES>(acc, As + (int64_t)i0 * n + kc, n, ri,
                                                 As + (int64_t)j0 * n + kc, n, rj, K, Asm, Bsm);
    bqo_acc_foreach(acc, [&](int r, int c, double& v) {
        if (r < ri && c < rj) As[(int64_t)(i0 + r) * n + (j0 + c)] -= v;
    });
}

// Copies the factored diagonal blocks into place and zeroes the strict upper triangle
// (scipy dpotrf clean=1).
__global__ void potrf_finish_kernel(double* __restrict__ A, int n, const int* __restrict__ info,
                                    const double* __restrict__ diag_scratch) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    int s = blockIdx.z;
    if (j >= n) return;
    if (j > i) {
        A[((int64_t)s * n + i) * n + j] = 0.0;
    } else if (i / NB == j / NB && info[s] == 0) {
        const int nb = (n + NB - 1) / NB;
        const int kb = i / NB;
        A[((int64_t)s * n + i) * n + j] =
            diag_scratch[(((int64_t)s * nb + kb) * NB + (i - kb * NB)) * NB + (j - kb * NB)];
    }
}

#define SMEM_LS ((size_t)NB * LDT * sizeof(double))
#define SMEM_SYRK ((size_t)(2 * BQO_OPER_BUF_DOUBLES(BQO_GEMM_STAGES)) * sizeof(double))
// the solves keep two CTAs per SM: two pipeline stages next to their two 64x64 work tiles
#define TRSM_STAGES 2
#define SMEM_TRSM ((size_t)(2 * BQO_OPER_BUF_DOUBLES(TRSM_STAGES) + 2 * NB * LDT) * sizeof(double))

#ifndef POTRF_W
#define POTRF_W 4
#endif
// Workspace of one factorisation call: S * nb * 64 * 64 doubles (inverses of the diagonal blocks;
// the multi-launch path keeps its factored diagonal blocks there) followed by S * nb + S + 4 ints
// (tile-row progress counters, per-matrix stop flags, task counter and abort flag).
static size_t potrf_workspace_doubles(int n, int S) {
    const size_t nb = (size_t)(n + NB - 1) / NB;
    return (size_t)S * nb * NB * NB;
}
static size_t potrf_workspace_ints(int n, int S) {
    const size_t nb = (size_t)(n + NB - 1) / NB;
    return 2 * (size_t)S * nb + (size_t)S + 4;
}
extern "C" int64_t bqo_potrf_workspace_bytes(int32_t n, int32_t S) {
    if (n <= 0 || S <= 0) return 0;
    return (int64_t)(potrf_workspace_doubles(n, S) * sizeof(double) +
                     ((potrf_workspace_ints(n, S) * sizeof(int) + 15) / 16) * 16);
}

// Multi-launch right-looking factorisation (round 1): kept for odd n (rows not 16-byte aligned:
// the persistent kernel reads finished tiles with L2-only 16-byte copies) and as the reference
// the persistent kernel is tested against (environment variable BQO_POTRF_LEGACY=1).
static int potrf_launch_multi(double* A, int n, int S, int* info, double* diag_scratch,
                              cudaStream_t cs) {
    const int nb = (n + NB - 1) / NB;
    bqo_smem_opt_in<potrf_syrk_kernel>(SMEM_POTRF_SYRK);
    bqo_smem_opt_in<potrf_panel_kernel<0>>(SMEM_LS);
    bqo_smem_opt_in<potrf_panel_kernel<1>>(SMEM_LS);
    bqo_smem_opt_in<potrf_panel_kernel<2>>(SMEM_LS);
    cudaMemsetAsync(info, 0, sizeof(int) * S, cs);
    auto panel = [&](int k0) {
        const int tiles = (n - k0 + NB - 1) / NB;  // diagonal tile + row tiles below it
        if ((int64_t)tiles * S <= 3 * 148 || tiles == 1) {
            dim3 pg(tiles, S);
            BQO_L, potrf_panel_kernel<0><<<pg, 128, SMEM_LS, cs>>>(A, n, k0, info, diag_scratch);
        } else {
            dim3 fg(1, S), sg(tiles - 1, S);
            BQO_L, potrf_panel_kernel<1><<<fg, 128, SMEM_LS, cs>>>(A, n, k0, info, diag_scratch);
            BQO_L, potrf_panel_kernel<2><<<sg, 128, SMEM_LS, cs>>>(A, n, k0, info, diag_scratch);
        }
    };
    auto syrk = [&](int kc, int K, int base, int ncolblk) {
        const int tr = (n - base + NB - 1) / NB;
        if (tr <= 0) return;
        dim3 sg(ncolblk < tr ? ncolblk : tr, tr, S);
        BQO_L, potrf_syrk_kernel<<<sg, 128, SMEM_POTRF_SYRK, cs>>>(A, n, kc, K, base, ncolblk, info);
    };
    // POTRF_W 64-wide panels per step.  Panel p of a step first receives, as a one-block-column
    // strip, the update of the p panels before it (K = 64 p); the full trailing update then runs
    // once per step with K = 64 POTRF_W: 1/POTRF_W of the passes over the trailing matrix and
    // POTRF_W times the work per operand tile (K = 64 updates ran at 12 TFLOP/s).
    for (int k0 = 0; k0 < n; k0 += POTRF_W * NB) {
        for (int p = 0; p < POTRF_W; ++p) {
            const int kp = k0 + p * NB;
            if (kp >= n) break;
            if (p > 0) syrk(k0, p * NB, kp, 1);
            panel(kp);
        }
        const int kend = k0 + POTRF_W * NB;
        if (kend < n) syrk(k0, POTRF_W * NB, kend, 1 << 30);
    }
    dim3 zg((n + 255) / 256, n, S);
    BQO_L, potrf_finish_kernel<<<zg, 256, 0, cs>>>(A, n, info, diag_scratch);
    return (int)cudaGetLastError();
}

// One persistent cooperative launch (bqo_potrf_persistent.cuh).
static unsigned long long* g_potrf_prof = nullptr;   // debug: bqo_potrf_profile
static int potrf_launch_persistent(double* A, int n, int S, int* info, double* Wd, int* ints,
                                   double* Winv, cudaStream_t cs) {
    const int nb = (n + NB - 1) / NB;
    static int ctas_per_sm[64] = {0}, sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return -100;
    bqo_smem_opt_in<potrf_persistent_kernel>(PP_SMEM_BYTES);
    if (ctas_per_sm[dev] == 0) {
        int per = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, potrf_persistent_kernel, PP_THREADS,
                                                      PP_SMEM_BYTES);
        cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        ctas_per_sm[dev] = per > 0 ? per : -1;
    }
    if (ctas_per_sm[dev] < 1) return -101;
    const int64_t total = Winv ? (int64_t)S * nb * nb : (int64_t)S * ((int64_t)nb * (nb + 1) / 2);
    int64_t grid = (int64_t)sms[dev] * ctas_per_sm[dev];
    if (grid > total) grid = total;
    cudaMemsetAsync(info, 0, sizeof(int) * S, cs);
    cudaMemsetAsync(ints, 0, sizeof(int) * potrf_workspace_ints(n, S), cs);
    PotrfPersistArgs a;
    a.A = A;
    a.n = n;
    a.S = S;
    a.nb = nb;
    a.info = info;
    a.rowdone = ints;
    a.wdone = ints + (size_t)S * nb;
    a.failed = a.wdone + (size_t)S * nb;
    a.counters = a.failed + S;
    a.Wd = Wd;
    a.Wmat = Winv;
    a.prof = g_potrf_prof;
    void* args[] = {&a};
    BQO_L;
    cudaError_t e = cudaLaunchCooperativeKernel((void*)potrf_persistent_kernel, dim3((unsigned)grid),
                                                dim3(PP_THREADS), args, PP_SMEM_BYTES, cs);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}

int bqo_internal_assemble(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                          const double* Xs, const int32_t* tid, int32_t n, const double* noise_obs,
                          int32_t add_noise, const double* extra_diag, double* K, int lower_only,
                          void* stream);

static int g_potrf_legacy = -1;
extern "C" int bqo_potrf_set_multi_launch(int on) {
    const int old = g_potrf_legacy > 0 ? 1 : 0;
    g_potrf_legacy = on ? 1 : 0;
    return old;
}

int bqo_internal_tri_inverse(const double* L, int n, int S, double* Wbuf, double* scratch,
                              cudaStream_t cs);

// Winv (optional, [S][n][n]): also W = L^-1 on the lower tiles -- by extra tasks of the persistent
// kernel, or by the level kernels after the multi-launch path (scratch = the S n^2 doubles of T).
static int potrf_launch(double* A, int n, int S, int* info, void* workspace, cudaStream_t cs,
                        double* Winv = nullptr, double* Wscratch = nullptr) {
    double* wd = (double*)workspace;
    int* ints = (int*)(wd + potrf_workspace_doubles(n, S));
    if (g_potrf_legacy < 0) {
        const char* e = getenv("BQO_POTRF_LEGACY");
        g_potrf_legacy = (e && e[0] == '1') ? 1 : 0;
    }
    const int legacy = g_potrf_legacy;
    // 16-byte aligned rows (n even, base aligned) are what the persistent kernel's L2-only copies
    // of finished tiles need
    const bool aligned = (n % 2 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    if (legacy || !aligned) {
        int rc = potrf_launch_multi(A, n, S, info, wd, cs);
        if (rc == 0 && Winv) rc = bqo_internal_tri_inverse(A, n, S, Winv, Wscratch, cs);
        return rc;
    }
    return potrf_launch_persistent(A, n, S, info, wd, ints, Winv, cs);
}

// Tuning aid: the same factorisation with %globaltimer stamps of the critical-path tasks of matrix 0
// written to prof[nb][2][8] (device, zero-initialised by the caller).  Not thread safe.
extern "C" int bqo_potrf_profile(double* A, int32_t n, int32_t S, int32_t* info, void* workspace,
                                 unsigned long long* prof, void* stream) {
    g_potrf_prof = prof;
    int rc = potrf_launch(A, n, S, info, workspace, (cudaStream_t)stream);
    g_potrf_prof = nullptr;
    return rc;
}

extern "C" int bqo_potrf_batched(double* A, int32_t n, int32_t S, int32_t* info, void* workspace,
                                 void* stream) {
    BQO_CHECK_ARG(A != nullptr, 1);
    BQO_CHECK_ARG(n > 0, 2);
    BQO_CHECK_ARG(S > 0 && S <= 65535, 3);
    BQO_CHECK_ARG(info != nullptr, 4);
    BQO_CHECK_ARG(workspace != nullptr, 5);
    return potrf_launch(A, n, S, info, workspace, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Triangular solves, RHS-major.  CTA = 64 right-hand sides x all block columns.
//   forward : L y = b      y[kb] = (b[kb] - sum_{j<kb} L[kb][j] y[j]) / L[kb][kb]
//   backward: L^T x = y    x[kb] = (y[kb] - sum_{j>kb} L[j][kb]^T x[j]) / L[kb][kb]^T
// `first_block_is_tile`: the right-hand sides are rows of the identity, so for RHS tile rt the
// forward solution vanishes before block rt and the walk can start there.
template <bool FORWARD>
__global__ void __launch_bounds__(128) trsm_kernel(const double* __restrict__ L, int n,
                                                   double* __restrict__ Bt, int m,
                                                   int identity_rhs) {
    extern __shared__ __align__(16) double dyn_smem[];
    double* Asm = dyn_smem;
    double* Bsm = Asm + BQO_OPER_BUF_DOUBLES(TRSM_STAGES);
    double* Ts = Bsm + BQO_OPER_BUF_DOUBLES(TRSM_STAGES);
    double* Ls = Ts + NB * LDT;
    const int s = blockIdx.y;
    const double* Lm = L + (int64_t)s * n * n;
    const int r0 = blockIdx.x * NB;
    const int rv = (m - r0 < NB) ? m - r0 : NB;
    double* Bs_ = Bt + ((int64_t)s * m + r0) * n;
    const int nb = (n + NB - 1) / NB;
    const int t = threadIdx.x;
    int kb_begin = FORWARD ? 0 : nb - 1;
    if (FORWARD && identity_rhs) kb_begin = r0 / NB;
    for (int kb = kb_begin; FORWARD ? (kb < nb) : (kb >= 0); kb += FORWARD ? 1 : -1) {
        const int k0 = kb * NB;
        const int kv = (n - k0 < NB) ? n - k0 : NB;
        BqoAcc acc;
        acc.zero();
        if (FORWARD) {
            int kstart = identity_rhs ? kb_begin * NB : 0;
            // acc[r][c] = sum_{k<k0} Y[r][k] L[k0+c][k]
            if (k0 > kstart)
                bqo_gemm_tile<true, true, TRSM_STAGES>(acc, Bs_ + kstart, n, rv,
                                                       Lm + (int64_t)k0 * n + kstart, n, kv,
                                                       k0 - kstart, Asm, Bsm);
        } else {
            // acc[r][c] = sum_{k>=k0+NB} X[r][k] L[k][k0+c]
            int kk = k0 + NB;
            if (kk < n)
                bqo_gemm_tile<true, false, TRSM_STAGES>(acc, Bs_ + kk, n, rv,
                                                        Lm + (int64_t)kk * n + k0, n, kv, n - kk,
                                                        Asm, Bsm);
        }
        __syncthreads();
        bqo_acc_foreach(acc, [&](int r, int c, double& v) {
            double b = (r < rv && c < kv) ? Bs_[(int64_t)r * n + k0 + c] : 0.0;
            Ts[r * LDT + c] = b - v;
        });
        for (int e = t; e < NB * NB; e += blockDim.x) {
            int r = e / NB, c = e % NB;
            double v = 0.0;
            if (r < kv && c <= r) v = Lm[(int64_t)(k0 + r) * n + (k0 + c)];
            if (r >= kv && c == r) v = 1.0;
            Ls[r * LDT + c] = v;
        }
        __syncthreads();
        if (t < rv) {
            double* x = Ts + t * LDT;
            if (FORWARD) {
                for (int j = 0; j < kv; ++j) {
                    double a0 = x[j], a1 = 0.0;
                    int k = 0;
                    for (; k + 1 < j; k += 2) {
                        a0 = fma(-x[k], Ls[j * LDT + k], a0);
                        a1 = fma(-x[k + 1], Ls[j * LDT + k + 1], a1);
                    }
                    if (k < j) a0 = fma(-x[k], Ls[j * LDT + k], a0);
                    x[j] = (a0 + a1) / Ls[j * LDT + j];
                }
            } else {
                for (int j = kv - 1; j >= 0; --j) {
                    double a0 = x[j], a1 = 0.0;
                    int k = j + 1;
                    for (; k + 1 < kv; k += 2) {
                        a0 = fma(-x[k], Ls[k * LDT + j], a0);
                        a1 = fma(-x[k + 1], Ls[(k + 1) * LDT + j], a1);
                    }
                    if (k < kv) a0 = fma(-x[k], Ls[k * LDT + j], a0);
                    x[j] = (a0 + a1) / Ls[j * LDT + j];
                }
            }
        }
        __syncthreads();
        for (int e = t; e < NB * NB; e += blockDim.x) {
            int r = e / NB, c = e % NB;
            if (r < rv && c < kv) Bs_[(int64_t)r * n + k0 + c] = Ts[r * LDT + c];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Few right-hand sides (alpha = Sigma^-1 (y - mean): m = 1): the tile path above would leave one
// CTA per matrix walking 2 x 32 block columns with a 64x64x64 DMMA product and a serial
// substitution each (6 ms for 20 matrices of n = 2048).  A vector solve is bandwidth work: one
// CTA per (right-hand side, matrix) keeps the vector in shared memory and streams L once per
// direction with coalesced row reads:
//   forward  (L y = b)   left-looking : rows of block kb take dot products with y[0:k0]
//                                       (warp per row group, lanes along the row);
//   backward (L^T x = y) right-looking: after block kb is solved, y[0:k0] -= L[kb, 0:k0]^T x[kb]
//                                       (thread per column, rows of L read contiguously).
// The 64x64 diagonal systems are solved by one warp with shuffles (reciprocal diagonal).
#define TRSV_THREADS 512
#define TRSV_WARPS (TRSV_THREADS / 32)
#define TRSV_ROWS (NB / TRSV_WARPS)  // rows of a block handled by one warp

template <bool FORWARD>
__global__ void __launch_bounds__(TRSV_THREADS) trsv_kernel(const double* __restrict__ L, int n,
                                                            double* __restrict__ Bt, int m, int r0,
                                                            int r1) {
    extern __shared__ __align__(16) double dyn_smem[];
    double* xs = dyn_smem;                         // [nb * NB] the vector (zero padded)
    const int nb = (n + NB - 1) / NB;
    double* Ls = xs + (size_t)nb * NB;             // [NB][LDT] diagonal block
    double* rd = Ls + NB * LDT;                    // [NB] reciprocal diagonal
    const int s = blockIdx.y;
    const double* Lm = L + (int64_t)s * n * n;
    double* b = Bt + ((int64_t)s * m + blockIdx.x) * n;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // rows [r0, r1) of the system (r0 a multiple of 64): everything outside was already
    // eliminated by the caller (trsv_gemv_*), see potrs_launch
    const int kb0 = r0 / NB, kb1 = (r1 + NB - 1) / NB;
    for (int i = r0 + t; i < kb1 * NB; i += TRSV_THREADS) xs[i] = (i < r1) ? b[i] : 0.0;
    __syncthreads();
    for (int kb = FORWARD ? kb0 : kb1 - 1; FORWARD ? (kb < kb1) : (kb >= kb0);
         kb += FORWARD ? 1 : -1) {
        const int k0 = kb * NB;
        const int kv = (n - k0 < NB) ? n - k0 : NB;
        // diagonal block (and the reciprocals of its diagonal) into shared memory; the backward
        // walk loads it one block ahead, behind the elimination step of the previous block
        if (FORWARD || kb == kb1 - 1) {
#pragma unroll
            for (int e = t; e < NB * NB; e += TRSV_THREADS) {
                int r = e / NB, c = e % NB;
                double v = 0.0;
                if (r < kv && c <= r) v = Lm[(int64_t)(k0 + r) * n + (k0 + c)];
                Ls[r * LDT + c] = v;
                if (r == c) rd[r] = (r < kv) ? 1.0 / v : 1.0;
            }
        }
        if (FORWARD && k0 > r0) {
            // rows k0 + warp*TRSV_ROWS + q: acc_q = sum_{c < k0} L[row][c] y[c]
            double acc[TRSV_ROWS];
#pragma unroll
            for (int q = 0; q < TRSV_ROWS; ++q) acc[q] = 0.0;
            const int rbase = k0 + warp * TRSV_ROWS;
            // k0 is a multiple of 64: two columns per lane and trip, no bounds checks; the loads
            // of four trips are issued together (the walk is latency bound otherwise)
            const double* lrow[TRSV_ROWS];
#pragma unroll
            for (int q = 0; q < TRSV_ROWS; ++q) {
                const int row = (rbase + q < n) ? rbase + q : n - 1;  // clamped rows are discarded
                lrow[q] = Lm + (int64_t)row * n;
            }
#pragma unroll 4
            for (int c = r0 + lane; c < k0; c += 64) {
                const double x0 = xs[c], x1 = xs[c + 32];
#pragma unroll
                for (int q = 0; q < TRSV_ROWS; ++q) {
                    acc[q] = fma(lrow[q][c], x0, acc[q]);
                    acc[q] = fma(lrow[q][c + 32], x1, acc[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < TRSV_ROWS; ++q) {
                double a = bqo_warp_sum(acc[q]);
                // rows >= k0 are not read by the dots above
                if (lane == 0 && rbase + q < n) xs[rbase + q] -= a;
            }
        }
        __syncthreads();
        if (warp == 0) {
            // lanes own entries lane and lane + 32 of the block
            double v0 = xs[k0 + lane], v1 = xs[k0 + lane + 32];
            if (FORWARD) {
                for (int j = 0; j < kv; ++j) {
                    double cand = ((j < 32) ? v0 : v1) * rd[j];
                    const double yj = __shfl_sync(0xffffffffu, cand, j & 31);
                    if (lane == (j & 31)) {
                        if (j < 32) v0 = yj; else v1 = yj;
                    }
                    if (lane > j) v0 = fma(-Ls[lane * LDT + j], yj, v0);
                    if (lane + 32 > j) v1 = fma(-Ls[(lane + 32) * LDT + j], yj, v1);
                }
            } else {
                for (int j = kv - 1; j >= 0; --j) {
                    double cand = ((j < 32) ? v0 : v1) * rd[j];
                    const double xj = __shfl_sync(0xffffffffu, cand, j & 31);
                    if (lane == (j & 31)) {
                        if (j < 32) v0 = xj; else v1 = xj;
                    }
                    // (L^T)[c][j] = L[j][c] for c < j: row j of the block, contiguous in c
                    if (lane < j) v0 = fma(-Ls[j * LDT + lane], xj, v0);
                    if (lane + 32 < j) v1 = fma(-Ls[j * LDT + lane + 32], xj, v1);
                }
            }
            xs[k0 + lane] = v0;
            xs[k0 + lane + 32] = v1;
        }
        __syncthreads();
        if (!FORWARD && k0 > r0) {
            // next diagonal block: its global loads go out first, the stores follow the
            // elimination below (the block just solved is not read any more)
            const int kn0 = k0 - NB;  // a full block: kn0 >= r0
            double nv[NB * NB / TRSV_THREADS];
#pragma unroll
            for (int q = 0; q < NB * NB / TRSV_THREADS; ++q) {
                const int e = t + q * TRSV_THREADS;
                const int r = e / NB, c = e % NB;
                nv[q] = (c <= r) ? Lm[(int64_t)(kn0 + r) * n + (kn0 + c)] : 0.0;
            }
            // y[c] -= sum_j L[k0 + j][c] x[k0 + j] for every column r0 <= c < k0: the CTA
            // covers 128 columns x 4 row groups at a time, all loads of a thread in flight
            __shared__ double part[4][128];
            const int tx = t & 127, ty = t >> 7;
            for (int c0 = r0; c0 < k0; c0 += 128) {
                const int c = c0 + tx;
                double a0 = 0.0, a1 = 0.0;
                if (c < k0) {
                    const double* lc = Lm + (int64_t)(k0 + ty * 16) * n + c;
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const int jj = ty * 16 + j;
                        if (jj < kv) a0 = fma(lc[(int64_t)j * n], xs[k0 + jj], a0);
                        if (jj + 1 < kv) a1 = fma(lc[(int64_t)(j + 1) * n], xs[k0 + jj + 1], a1);
                    }
                }
                part[ty][tx] = a0 + a1;
                __syncthreads();
                if (ty == 0 && c < k0)
                    xs[c] -= (part[0][tx] + part[1][tx]) + (part[2][tx] + part[3][tx]);
                __syncthreads();
            }
#pragma unroll
            for (int q = 0; q < NB * NB / TRSV_THREADS; ++q) {
                const int e = t + q * TRSV_THREADS;
                const int r = e / NB, c = e % NB;
                Ls[r * LDT + c] = nv[q];
                if (r == c) rd[r] = 1.0 / nv[q];
            }
        }
    }
    for (int i = r0 + t; i < r1; i += TRSV_THREADS) b[i] = xs[i];
}

// Elimination of a solved block from the rest of the vector, spread over many CTAs (a single CTA
// per vector streams L at the bandwidth of one SM):
//   forward : b[r] -= sum_{c in [c0, c1)} L[r][c] y[c]   for r >= c1   (warp per 8 rows)
//   backward: y[c] -= sum_{j in [r0, r1)} L[j][c] x[j]   for c <  r0   (thread per column)
#define TRSV_BLK 256
__global__ void __launch_bounds__(256) trsv_gemv_fwd_kernel(const double* __restrict__ L, int n,
                                                            double* __restrict__ Bt, int m, int c0,
                                                            int c1) {
    __shared__ double xb[TRSV_BLK];
    const int s = blockIdx.z, v = blockIdx.y;
    const double* Lm = L + (int64_t)s * n * n;
    double* b = Bt + ((int64_t)s * m + v) * n;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < TRSV_BLK; i += 256) xb[i] = (c0 + i < c1) ? b[c0 + i] : 0.0;
    __syncthreads();
    const int rbase = c1 + blockIdx.x * 64 + warp * 8;
    double acc[8];
    const double* lrow[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        acc[q] = 0.0;
        const int row = (rbase + q < n) ? rbase + q : n - 1;
        lrow[q] = Lm + (int64_t)row * n + c0;
    }
    const int w = c1 - c0;  // multiple of 64 (c1 < n here)
    for (int c = lane; c < w; c += 32) {
        const double x = xb[c];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fma(lrow[q][c], x, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const double a = bqo_warp_sum(acc[q]);
        if (lane == 0 && rbase + q < n) b[rbase + q] -= a;
    }
}

__global__ void __launch_bounds__(256) trsv_gemv_bwd_kernel(const double* __restrict__ L, int n,
                                                            double* __restrict__ Bt, int m, int r0,
                                                            int r1) {
    // CTA = 64 columns x 4 row groups (64 rows of the block each); partial sums meet in shared memory
    __shared__ double xb[TRSV_BLK];
    __shared__ double part[4][64];
    const int s = blockIdx.z, v = blockIdx.y;
    const double* Lm = L + (int64_t)s * n * n;
    double* b = Bt + ((int64_t)s * m + v) * n;
    const int t = threadIdx.x, tx = t & 63, ty = t >> 6;
    for (int i = t; i < TRSV_BLK; i += 256) xb[i] = (r0 + i < r1) ? b[r0 + i] : 0.0;
    __syncthreads();
    const int c = blockIdx.x * 64 + tx;
    const int h = r1 - r0;
    const int j0 = ty * 64, j1 = (j0 + 64 < h) ? j0 + 64 : h;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (c < r0) {
        const double* lc = Lm + (int64_t)r0 * n + c;
        int j = j0;
#pragma unroll 4
        for (; j + 3 < j1; j += 4) {
            a0 = fma(lc[(int64_t)j * n], xb[j], a0);
            a1 = fma(lc[(int64_t)(j + 1) * n], xb[j + 1], a1);
            a2 = fma(lc[(int64_t)(j + 2) * n], xb[j + 2], a2);
            a3 = fma(lc[(int64_t)(j + 3) * n], xb[j + 3], a3);
        }
        for (; j < j1; ++j) a0 = fma(lc[(int64_t)j * n], xb[j], a0);
    }
    part[ty][tx] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (ty == 0 && c < r0) b[c] -= (part[0][tx] + part[1][tx]) + (part[2][tx] + part[3][tx]);
}

// The vector path reads L once per right-hand side (n^2 x 8 bytes per direction), the tile path
// once per 64 of them but needs ~6 ms of serial block steps at n = 2048: cross-over near 500 vectors.
#define TRSV_MAX_VECTORS 512
static size_t trsv_smem_bytes(int n) {
    const int nb = (n + NB - 1) / NB;
    return ((size_t)nb * NB + (size_t)NB * LDT + NB) * sizeof(double);
}

static int potrs_launch(const double* L, int n, int S, double* Bt, int m, int identity_rhs,
                        cudaStream_t cs, bool forward_only = false) {
    if ((int64_t)m * S <= TRSV_MAX_VECTORS && m <= 65535 && !identity_rhs &&
        trsv_smem_bytes(n) <= 200 * 1024) {
        const size_t sb = trsv_smem_bytes(n);
        bqo_smem_opt_in<trsv_kernel<true>>(sb);
        bqo_smem_opt_in<trsv_kernel<false>>(sb);
        dim3 grid(m, S);
        // blocks of TRSV_BLK rows: one CTA per vector solves the block, then the block is
        // eliminated from the remaining rows by (rows / 64) x m x S CTAs
        for (int r0 = 0; r0 < n; r0 += TRSV_BLK) {
            const int r1 = (r0 + TRSV_BLK < n) ? r0 + TRSV_BLK : n;
            BQO_L, trsv_kernel<true><<<grid, TRSV_THREADS, sb, cs>>>(L, n, Bt, m, r0, r1);
            if (r1 < n) {
                dim3 gg((n - r1 + 63) / 64, m, S);
                BQO_L, trsv_gemv_fwd_kernel<<<gg, 256, 0, cs>>>(L, n, Bt, m, r0, r1);
            }
        }
        if (forward_only) return (int)cudaGetLastError();
        const int last = ((n - 1) / TRSV_BLK) * TRSV_BLK;
        for (int r0 = last; r0 >= 0; r0 -= TRSV_BLK) {
            const int r1 = (r0 + TRSV_BLK < n) ? r0 + TRSV_BLK : n;
            BQO_L, trsv_kernel<false><<<grid, TRSV_THREADS, sb, cs>>>(L, n, Bt, m, r0, r1);
            if (r0 > 0) {
                dim3 gg((r0 + 63) / 64, m, S);
                BQO_L, trsv_gemv_bwd_kernel<<<gg, 256, 0, cs>>>(L, n, Bt, m, r0, r1);
            }
        }
        return (int)cudaGetLastError();
    }
    bqo_smem_opt_in<trsm_kernel<true>>(SMEM_TRSM);
    bqo_smem_opt_in<trsm_kernel<false>>(SMEM_TRSM);
    dim3 grid((m + NB - 1) / NB, S);
    BQO_L, trsm_kernel<true><<<grid, 128, SMEM_TRSM, cs>>>(L, n, Bt, m, identity_rhs);
    if (forward_only) return (int)cudaGetLastError();
    BQO_L, trsm_kernel<false><<<grid, 128, SMEM_TRSM, cs>>>(L, n, Bt, m, 0);
    return (int)cudaGetLastError();
}

extern "C" int bqo_potrs_batched(const double* L, int32_t n, int32_t S, double* Bt, int32_t m,
                                 void* stream) {
    BQO_CHECK_ARG(L != nullptr, 1);
    BQO_CHECK_ARG(n > 0, 2);
    BQO_CHECK_ARG(S > 0 && S <= 65535, 3);
    BQO_CHECK_ARG(Bt != nullptr, 4);
    BQO_CHECK_ARG(m > 0, 5);
    return potrs_launch(L, n, S, Bt, m, 0, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Internal: Sigma^{-1} = L^-T L^-1 (used by the log-likelihood gradient).
//
// W = L^-1 is built by recursive doubling so that all but n*64^2 of its n^3/3 flops are DMMA tile
// products (a column-walking triangular solve on the identity is a serial chain of 32 small
// steps per CTA and ran at 7 TFLOP/s):
//   level 0 : the 64x64 diagonal blocks are inverted in shared memory (one thread per column);
//   level b = 64, 128, ...: for every aligned pair of b x b diagonal blocks
//       [ L11  0  ]^-1   [ W11              0  ]
//       [ L21 L22 ]    = [ -W22 L21 W11    W22 ],   T = L21 W11, then W21 = -W22 T,
//     two launches per level; the triangular zeros of W11 / W22 shorten the K ranges.
// Then Inv = W^T W on lower tiles (mirrored on store), K range from the tile's first row.
int bqo_internal_inverse_from_chol(const double* L, int n, int S, double* Inv, double* Wbuf,
                                   cudaStream_t cs);

__global__ void __launch_bounds__(NB) tri_inv_diag_kernel(const double* __restrict__ L, int n,
                                                          double* __restrict__ W) {
    extern __shared__ __align__(16) double dyn_smem[];
    double* Ls = dyn_smem;
    double* Ws = Ls + NB * LDT;
    const int s = blockIdx.y;
    const int k0 = blockIdx.x * NB;
    const int kv = (n - k0 < NB) ? n - k0 : NB;
    const double* Lm = L + (int64_t)s * n * n;
    double* Wm = W + (int64_t)s * n * n;
    const int t = threadIdx.x;
    for (int e = t; e < NB * NB; e += NB) {
        int r = e / NB, c = e % NB;
        Ls[r * LDT + c] = (r < kv && c <= r) ? Lm[(int64_t)(k0 + r) * n + (k0 + c)] : 0.0;
        Ws[r * LDT + c] = 0.0;
    }
    __syncthreads();
    if (t < kv) {
        // column t of the inverse: forward substitution on the unit vector e_t
        Ws[t * LDT + t] = 1.0 / Ls[t * LDT + t];
        for (int i = t + 1; i < kv; ++i) {
            double a0 = 0.0, a1 = 0.0;
            int k = t;
            for (; k + 1 < i; k += 2) {
                a0 = fma(Ls[i * LDT + k], Ws[k * LDT + t], a0);
                a1 = fma(Ls[i * LDT + k + 1], Ws[(k + 1) * LDT + t], a1);
            }
            if (k < i) a0 = fma(Ls[i * LDT + k], Ws[k * LDT + t], a0);
            Ws[i * LDT + t] = -(a0 + a1) / Ls[i * LDT + i];
        }
    }
    __syncthreads();
    for (int e = t; e < NB * NB; e += NB) {
        int r = e / NB, c = e % NB;
        if (r < kv && c < kv) Wm[(int64_t)(k0 + r) * n + (k0 + c)] = Ws[r * LDT + c];
    }
}

// PASS 1: T[R, C] = L[R, C] W[C, C];  PASS 2: W[R, C] = -W[R, R] T[R, C]
// for the pair whose top-left block starts at o = pair * 2b: C = [o, o+b), R = [o+b, o+2b).
// grid = (b/64 column tiles, b/64 row tiles, pairs * S).
template <int PASS>
__global__ void __launch_bounds__(128) tri_inv_level_kernel(const double* __restrict__ L,
                                                            double* __restrict__ W,
                                                            double* __restrict__ T, int n, int b,
                                                            int pairs) {
    extern __shared__ __align__(16) double dyn_smem[];
    double* Asm = dyn_smem;
    double* Bsm = Asm + BQO_OPER_BUF_DOUBLES(BQO_GEMM_STAGES);
    const int s = blockIdx.z / pairs;
    const int pair = blockIdx.z % pairs;
    const int o = pair * 2 * b;
    const int i0 = o + b + blockIdx.y * NB;
    const int j0 = o + blockIdx.x * NB;
    if (i0 >= n) return;
    const int ri = (n - i0 < NB) ? n - i0 : NB;
    const int rj = NB;  // o + b < n, so every column of C exists
    const double* Lm = L + (int64_t)s * n * n;
    double* Wm = W + (int64_t)s * n * n;
    double* Tm = T + (int64_t)s * n * n;
    BqoAcc acc;
    acc.zero();
    if (PASS == 1) {
        // sum over k in [j0, o+b): W11 is lower triangular, W[k][j] = 0 for k < j
        bqo_gemm_tile<true, false>(acc, Lm + (int64_t)i0 * n + j0, n, ri,
                                   Wm + (int64_t)j0 * n + j0, n, rj, o + b - j0, Asm, Bsm);
        bqo_acc_foreach(acc, [&](int r, int c, double& v) {
            if (r < ri) Tm[(int64_t)(i0 + r) * n + (j0 + c)] = v;
        });
    } else {
        // sum over k in [o+b, i0+64): W22 is lower triangular, W[i][k] = 0 for k > i
        const int kend = (i0 + NB < n) ? i0 + NB : n;
        bqo_gemm_tile<true, false>(acc, Wm + (int64_t)i0 * n + (o + b), n, ri,
                                   Tm + (int64_t)(o + b) * n + j0, n, rj, kend - (o + b), Asm, Bsm);
        bqo_acc_foreach(acc, [&](int r, int c, double& v) {
            if (r < ri) Wm[(int64_t)(i0 + r) * n + (j0 + c)] = -v;
        });
    }
}

// Inv = W^T W with W = L^-1 row-major lower triangular:
// Inv[a][b] = sum_{i >= max(a,b)} W[i][a] W[i][b].  Lower tiles only, K range starts at the
// tile's first row (W vanishes above it; the diagonal blocks carry explicit zeros).  The strict
// upper tiles of Inv are NOT written.
__global__ void __launch_bounds__(128) inv_syrk_kernel(const double* __restrict__ W, int n,
                                                       double* __restrict__ Inv) {
    extern __shared__ __align__(16) double dyn_smem[];
    double* Asm = dyn_smem;
    double* Bsm = Asm + BQO_OPER_BUF_DOUBLES(BQO_GEMM_STAGES);
    const int s = blockIdx.z;
    const int ba = blockIdx.y, bb = blockIdx.x;
    if (bb > ba) return;
    const double* Wm = W + (int64_t)s * n * n;
    double* Iv = Inv + (int64_t)s * n * n;
    const int a0 = ba * NB, b0 = bb * NB;
    const int ra = (n - a0 < NB) ? n - a0 : NB;
    const int rb = (n - b0 < NB) ? n - b0 : NB;
    BqoAcc acc;
    acc.zero();
    bqo_gemm_tile<false, false>(acc, Wm + (int64_t)a0 * n + a0, n, ra, Wm + (int64_t)a0 * n + b0, n,
                                rb, n - a0, Asm, Bsm);
    // only the lower tiles are stored (the diagonal tiles are complete squares): the one consumer,
    // the gradient contraction, visits exactly those
    bqo_acc_foreach(acc, [&](int r, int c, double& v) {
        if (r < ra && c < rb) Iv[(int64_t)(a0 + r) * n + (b0 + c)] = v;
    });
}

// W = L^-1 by recursive doubling (multi-launch path); `scratch`: S n^2 doubles (T of the levels).
int bqo_internal_tri_inverse(const double* L, int n, int S, double* Wbuf, double* scratch,
                             cudaStream_t cs) {
    const int nb = (n + NB - 1) / NB;
    const size_t smem_diag = 2 * (size_t)NB * LDT * sizeof(double);
    bqo_smem_opt_in<tri_inv_diag_kernel>(smem_diag);
    bqo_smem_opt_in<tri_inv_level_kernel<1>>(SMEM_SYRK);
    bqo_smem_opt_in<tri_inv_level_kernel<2>>(SMEM_SYRK);
    dim3 dg(nb, S);
    BQO_L, tri_inv_diag_kernel<<<dg, NB, smem_diag, cs>>>(L, n, Wbuf);
    for (int b = NB; b < n; b *= 2) {
        const int pairs = (n + 2 * b - 1) / (2 * b);
        if ((int64_t)pairs * S > 65535) return -3;
        dim3 lg(b / NB, b / NB, pairs * S);
        BQO_L, tri_inv_level_kernel<1><<<lg, 128, SMEM_SYRK, cs>>>(L, Wbuf, scratch, n, b, pairs);
        BQO_L, tri_inv_level_kernel<2><<<lg, 128, SMEM_SYRK, cs>>>(L, Wbuf, scratch, n, b, pairs);
    }
    return (int)cudaGetLastError();
}

// Inv = W^T W on the lower tiles.
int bqo_internal_inv_syrk(const double* Wbuf, int n, int S, double* Inv, cudaStream_t cs) {
    const int nb = (n + NB - 1) / NB;
    bqo_smem_opt_in<inv_syrk_kernel>(SMEM_SYRK);
    dim3 sg(nb, nb, S);
    BQO_L, inv_syrk_kernel<<<sg, 128, SMEM_SYRK, cs>>>(Wbuf, n, Inv);
    return (int)cudaGetLastError();
}

int bqo_internal_inverse_from_chol(const double* L, int n, int S, double* Inv, double* Wbuf,
                                   cudaStream_t cs) {
    // Inv doubles as the scratch T of the levels; the final product overwrites all of it
    int rc = bqo_internal_tri_inverse(L, n, S, Wbuf, Inv, cs);
    if (rc) return rc;
    return bqo_internal_inv_syrk(Wbuf, n, S, Inv, cs);
}

// ---------------------------------------------------------------------------------------------
// Sigma assembly + Cholesky with the reference's jitter policy (sbo/lib/la_functions.py:8-38).
__global__ void diag_stats_kernel(const double* __restrict__ A, int n, double* __restrict__ mean,
                                  int* __restrict__ nonpos) {
    __shared__ double red[32];
    __shared__ int anyflag;
    int s = blockIdx.x;
    if (threadIdx.x == 0) anyflag = 0;
    __syncthreads();
    double acc = 0.0;
    int bad = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double d = A[((int64_t)s * n + i) * n + i];
        acc += d;
        if (d <= 0.0) bad = 1;
    }
    if (bad) atomicOr(&anyflag, 1);
    double tot = bqo_block_sum(acc, red);
    __syncthreads();
    if (threadIdx.x == 0) {
        mean[s] = tot / (double)n;
        nonpos[s] = anyflag;
    }
}

// No allocation and no host synchronisation: Sigma is assembled into L, its diagonal statistics
// are taken (they are needed only if a factorisation fails, but the factorisation overwrites the
// diagonal), and all S matrices are factored.  `status` (device, 2 S doubles + S ints, see
// bqo_chol_cov_status_bytes) receives mean(diag Sigma_s), a flag for a non-positive diagonal entry
// and -- in `info` -- dpotrf's code per matrix.  The caller reads them together with whatever
// result it needs anyway and, for the matrices that failed, drives the reference's jitter policy
// (sbo/lib/la_functions.py:19-38) with bqo_chol_cov_retry.
extern "C" int bqo_chol_cov_batched(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                                    const double* Xs, const int32_t* tid, int32_t n,
                                    const double* noise_obs, double* L, int32_t* info,
                                    double* diag_mean, int32_t* diag_nonpos, double* Winv,
                                    double* Winv_scratch, void* workspace, void* stream) {
    BQO_CHECK_ARG(desc != nullptr, 1);
    BQO_CHECK_ARG(S > 0 && S <= 65535, 3);
    BQO_CHECK_ARG(L != nullptr, 8);
    BQO_CHECK_ARG(info != nullptr, 9);
    BQO_CHECK_ARG(diag_mean != nullptr, 10);
    BQO_CHECK_ARG(diag_nonpos != nullptr, 11);
    BQO_CHECK_ARG(workspace != nullptr, 14);
    BQO_CHECK_ARG(Winv == nullptr || Winv_scratch != nullptr || n % 2 == 0, 13);
    cudaStream_t cs = (cudaStream_t)stream;
    int rc = bqo_internal_assemble(desc, hyp, S, Xs, tid, n, noise_obs, 1, nullptr, L, 1, stream);
    if (rc) return rc;
    BQO_L, diag_stats_kernel<<<S, 256, 0, cs>>>(L, n, diag_mean, diag_nonpos);
    return potrf_launch(L, n, S, info, workspace, cs, Winv, Winv_scratch);
}

// One step of the jitter policy for matrix s: Sigma_s + jitter I is re-assembled into its slot of
// L and factored again; info_one (device, 1 int) receives the code.  `jitter_dev`: one double of
// device scratch (the value is copied there).  The workspace of bqo_potrf_workspace_bytes(n, 1)
// is enough.
extern "C" int bqo_chol_cov_retry(const bqo_kernel_desc* desc, const double* hyp, int32_t s,
                                  const double* Xs, const int32_t* tid, int32_t n,
                                  const double* noise_obs, double jitter, double* jitter_dev,
                                  double* L, int32_t* info_one, void* workspace, void* stream) {
    BQO_CHECK_ARG(desc != nullptr, 1);
    BQO_CHECK_ARG(s >= 0, 3);
    BQO_CHECK_ARG(jitter_dev != nullptr, 9);
    BQO_CHECK_ARG(L != nullptr, 10);
    BQO_CHECK_ARG(info_one != nullptr, 11);
    BQO_CHECK_ARG(workspace != nullptr, 12);
    cudaStream_t cs = (cudaStream_t)stream;
    const int64_t st = bqo_hyper_stride_impl(*desc);
    cudaMemcpyAsync(jitter_dev, &jitter, sizeof(double), cudaMemcpyHostToDevice, cs);
    int rc = bqo_internal_assemble(desc, hyp + (int64_t)s * st, 1, Xs + (int64_t)s * desc->dim_m * n,
                                   tid, n, noise_obs, 1, jitter_dev, L + (int64_t)s * n * n, 1,
                                   stream);
    if (rc) return rc;
    return potrf_launch(L + (int64_t)s * n * n, n, 1, info_one, workspace, cs);
}

// ---------------------------------------------------------------------------------------------
// Appending one training point to S factors (the reference refactorises, see
// the TODO at sbo/models/gp_fitting_gaussian.py:506-508):
//   Sigma' = [ Sigma  k ]      L' = [ L    0      ]    L l = k,   lambda = sqrt(kappa - l.l)
//            [ k^T  kappa ]         [ l^T  lambda ]
// O(n^2) per factor instead of n^3/3.  `knew` is the new row of Sigma' ([S][n+1], noise and any
// jitter of the old factor included in its last entry by the caller).
__global__ void chol_append_copy_kernel(const double* __restrict__ Lold, int n,
                                        double* __restrict__ Lnew) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y, s = blockIdx.z;
    const int n1 = n + 1;
    if (j >= n1) return;
    double v = 0.0;
    if (i < n && j < n) v = Lold[((int64_t)s * n + i) * n + j];
    Lnew[((int64_t)s * n1 + i) * n1 + j] = v;
}

__global__ void chol_append_row_kernel(const double* __restrict__ lrow, const double* __restrict__ knew,
                                       int n, double* __restrict__ Lnew, int* __restrict__ info) {
    __shared__ double red[32];
    const int s = blockIdx.x;
    const int n1 = n + 1;
    const double* l = lrow + (int64_t)s * n;
    double* dst = Lnew + ((int64_t)s * n1 + n) * n1;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = l[i];
        dst[i] = v;
        acc = fma(v, v, acc);
    }
    const double ll = bqo_block_sum(acc, red);
    if (threadIdx.x == 0) {
        const double d = knew[(int64_t)s * n1 + n] - ll;
        if (d > 0.0) {
            dst[n] = sqrt(d);
            info[s] = 0;
        } else {
            dst[n] = 0.0;
            info[s] = n1;  // like dpotrf: order of the leading minor that is not positive definite
        }
    }
}

extern "C" int bqo_chol_append_batched(const double* L_old, int32_t n, int32_t S, const double* knew,
                                       double* L_new, int32_t* info, void* workspace, void* stream) {
    BQO_CHECK_ARG(L_old != nullptr, 1);
    BQO_CHECK_ARG(n > 0, 2);
    BQO_CHECK_ARG(S > 0 && S <= 65535, 3);
    BQO_CHECK_ARG(knew != nullptr, 4);
    BQO_CHECK_ARG(L_new != nullptr, 5);
    BQO_CHECK_ARG(info != nullptr, 6);
    BQO_CHECK_ARG(workspace != nullptr, 7);  // S * n doubles
    cudaStream_t cs = (cudaStream_t)stream;
    double* lrow = (double*)workspace;
    cudaMemcpy2DAsync(lrow, sizeof(double) * n, knew, sizeof(double) * (n + 1), sizeof(double) * n, S,
                      cudaMemcpyDeviceToDevice, cs);
    int rc = potrs_launch(L_old, n, S, lrow, 1, 0, cs, true);  // forward solve only: L l = k
    if (rc) return rc;
    dim3 cg((n + 1 + 255) / 256, n + 1, S);
    BQO_L, chol_append_copy_kernel<<<cg, 256, 0, cs>>>(L_old, n, L_new);
    BQO_L, chol_append_row_kernel<<<S, 256, 0, cs>>>(lrow, knew, n, L_new, info);
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
__global__ void residual_kernel(const double* __restrict__ hyp, int64_t stride,
                                const double* __restrict__ y, int n, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int s = blockIdx.y;
    if (i < n) out[(int64_t)s * n + i] = y[i] - hyp[(int64_t)s * stride + BQO_H_MEAN];
}

extern "C" int bqo_alpha_batched(const double* hyp, int64_t hyp_stride, const double* L, int32_t n,
                                 int32_t S, const double* y, double* alpha, void* stream) {
    BQO_CHECK_ARG(hyp != nullptr, 1);
    BQO_CHECK_ARG(L != nullptr, 3);
    BQO_CHECK_ARG(n > 0, 4);
    BQO_CHECK_ARG(S > 0 && S <= 65535, 5);
    BQO_CHECK_ARG(y != nullptr, 6);
    BQO_CHECK_ARG(alpha != nullptr, 7);
    cudaStream_t cs = (cudaStream_t)stream;
    dim3 grid((n + 255) / 256, S);
    BQO_L, residual_kernel<<<grid, 256, 0, cs>>>(hyp, hyp_stride, y, n, alpha);
    // alpha is [S][1][n]: one right-hand side per matrix
    return potrs_launch(L, n, S, alpha, 1, 0, cs);
}

// ---------------------------------------------------------------------------------------------
// alpha = Sigma^-1 (y - mean) from W = L^-1 (when the factorisation was asked for it):
// alpha = W^T (W r), two passes over the lower triangle of W at HBM speed (n^2 doubles read per
// matrix in total) instead of the two triangular vector solves, which are chains of ~70 small
// launches (0.9 ms for 20 matrices of n = 2048 against 0.15 ms).  Only tiles on or below the
// diagonal of W are defined, so both kernels stop at the diagonal by index.
__global__ void __launch_bounds__(256) alpha_w_forward_kernel(const double* __restrict__ hyp,
                                                              int64_t stride,
                                                              const double* __restrict__ W, int n,
                                                              const double* __restrict__ y,
                                                              double* __restrict__ t) {
    // one warp per row i: t[i] = sum_{j <= i} W[i][j] (y[j] - mean)
    const int s = blockIdx.y;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const double mean = hyp[(int64_t)s * stride + BQO_H_MEAN];
    const double* row = W + ((int64_t)s * n + i) * n;
    double a0 = 0.0, a1 = 0.0;
    int j = lane;
    for (; j + 32 <= i; j += 64) {
        a0 = fma(row[j], y[j] - mean, a0);
        a1 = fma(row[j + 32], y[j + 32] - mean, a1);
    }
    for (; j <= i; j += 32) a0 = fma(row[j], y[j] - mean, a0);
    const double tot = bqo_warp_sum(a0 + a1);
    if (lane == 0) t[(int64_t)s * n + i] = tot;
}

__global__ void __launch_bounds__(256) alpha_w_backward_kernel(const double* __restrict__ W, int n,
                                                               const double* __restrict__ t,
                                                               double* __restrict__ alpha) {
    // CTA = 32 columns j, 8 warps striding over the rows i >= j: alpha[j] = sum_i W[i][j] t[i]
    __shared__ double part[8][32];
    const int s = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j0 = blockIdx.x * 32, j = j0 + lane;
    const double* Wm = W + (int64_t)s * n * n;
    const double* ts = t + (int64_t)s * n;
    double a0 = 0.0, a1 = 0.0;
    if (j < n) {
        int i = j0 + warp;
        for (; i + 8 < n; i += 16) {
            if (i >= j) a0 = fma(Wm[(int64_t)i * n + j], ts[i], a0);
            if (i + 8 >= j) a1 = fma(Wm[(int64_t)(i + 8) * n + j], ts[i + 8], a1);
        }
        for (; i < n; i += 8)
            if (i >= j) a0 = fma(Wm[(int64_t)i * n + j], ts[i], a0);
    }
    part[warp][lane] = a0 + a1;
    __syncthreads();
    if (warp == 0 && j < n) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += part[w][lane];
        alpha[(int64_t)s * n + j] = tot;
    }
}

extern "C" int bqo_alpha_from_inverse_batched(const double* hyp, int64_t hyp_stride,
                                              const double* Winv, int32_t n, int32_t S,
                                              const double* y, double* alpha, double* work,
                                              void* stream) {
    BQO_CHECK_ARG(hyp != nullptr, 1);
    BQO_CHECK_ARG(Winv != nullptr, 3);
    BQO_CHECK_ARG(n > 0, 4);
    BQO_CHECK_ARG(S > 0 && S <= 65535, 5);
    BQO_CHECK_ARG(y != nullptr, 6);
    BQO_CHECK_ARG(alpha != nullptr, 7);
    BQO_CHECK_ARG(work != nullptr, 8);   // S * n doubles
    cudaStream_t cs = (cudaStream_t)stream;
    dim3 g1((n + 7) / 8, S), g2((n + 31) / 32, S);
    BQO_L, alpha_w_forward_kernel<<<g1, 256, 0, cs>>>(hyp, hyp_stride, Winv, n, y, work);
    BQO_L, alpha_w_backward_kernel<<<g2, 256, 0, cs>>>(Winv, n, work, alpha);
    BQO_RETURN_LAST_ERROR();
}

__global__ void loglik_kernel(const double* __restrict__ hyp, int64_t stride,
                              const double* __restrict__ L, const double* __restrict__ alpha,
                              const double* __restrict__ y, int n, double* __restrict__ llh) {
    __shared__ double red[32];
    int s = blockIdx.x;
    double mean = hyp[(int64_t)s * stride + BQO_H_MEAN];
    double sl = 0.0, sq = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        sl += log(L[((int64_t)s * n + i) * n + i]);
        sq = fma(y[i] - mean, alpha[(int64_t)s * n + i], sq);
    }
    double a = bqo_block_sum(sl, red);
    double b = bqo_block_sum(sq, red);
    if (threadIdx.x == 0) llh[s] = -a - 0.5 * b;
}

extern "C" int bqo_loglik_batched(const double* hyp, int64_t hyp_stride, const double* L,
                                  const double* alpha, const double* y, int32_t n, int32_t S,
                                  double* llh, void* stream) {
    BQO_CHECK_ARG(hyp != nullptr, 1);
    BQO_CHECK_ARG(L != nullptr, 3);
    BQO_CHECK_ARG(alpha != nullptr, 4);
    BQO_CHECK_ARG(y != nullptr, 5);
    BQO_CHECK_ARG(n > 0, 6);
    BQO_CHECK_ARG(S > 0, 7);
    BQO_CHECK_ARG(llh != nullptr, 8);
    BQO_L, loglik_kernel<<<S, 256, 0, (cudaStream_t)stream>>>(hyp, hyp_stride, L, alpha, y, n, llh);
    BQO_RETURN_LAST_ERROR();
}
