// FP64 pipe micro-benchmark for B200 (sm_100a): dependent-issue latency and throughput of DFMA
// as a function of resident warps per SM sub-partition and independent chains per thread, with
// all three source operands in distinct registers (the shape of the stage-C inner loop).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu && ./fp64_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, int MODE>
__global__ void k(double* out, const double* in, int iters) {
    double x[CH], y[CH], z[CH];
    int ix[CH];
    for (int c = 0; c < CH; ++c) ix[c] = threadIdx.x + c;
    for (int c = 0; c < CH; ++c) {
        x[c] = in[threadIdx.x + c];
        y[c] = in[threadIdx.x + 64 + c];
        z[c] = in[threadIdx.x + 128 + c];
    }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (MODE == 0) x[c] = fma(x[c], y[c], z[c]);          // 3 distinct registers
            if (MODE == 1) x[c] = fma(x[c], x[c], z[c]);          // 2 distinct registers
            if (MODE == 2) x[c] = fma(x[c], 1.0000001, 1e-7);     // 1 register + constants
            if (MODE == 3) x[c] = x[c] * y[c];                    // DMUL
            if (MODE == 4) x[c] = x[c] + y[c];                    // DADD
            if (MODE == 5 || MODE == 6 || MODE == 7) x[c] = fma(x[c], x[c], z[c]);
            // accumulator forms: does the operand-reuse cache hide the third register read?
            if (MODE == 8) x[c] = fma(y[c], z[0], x[c]);        // one multiplicand shared by all chains
            if (MODE == 9) x[c] = fma(y[c], z[c >> 1], x[c]);   // shared by pairs (E1 += u e; H += u2 e)
            if (MODE == 10) x[c] = fma(y[c], z[c], x[c]);       // accumulator, three distinct registers
        }
        // MODE 5/6/7: CH two-register DFMAs followed by CH/2CH/4CH independent integer ops: does a
        // non-FP64 instruction issue "for free" in the second cycle of an FP64 instruction?
        if (MODE >= 5 && MODE <= 7) {
            const int reps = (MODE == 5) ? 1 : (MODE == 6 ? 2 : 4);
#pragma unroll
            for (int r = 0; r < reps; ++r)
#pragma unroll
                for (int c = 0; c < CH; ++c) ix[c] = (ix[c] ^ (ix[c] >> 3)) + i;
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int c = 0; c < CH; ++c) s += x[c] + (double)ix[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) ((long long*)out)[gridDim.x * blockDim.x] = t1 - t0;
}

template <int CH, int MODE>
void run(int warps_per_sm, double* out, double* in, int sms) {
    const int iters = 20000;
    int threads = warps_per_sm * 32;
    k<CH, MODE><<<sms, threads>>>(out, in, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<CH, MODE><<<sms, threads>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc;
    cudaMemcpy(&cyc, ((long long*)out) + sms * threads, 8, cudaMemcpyDeviceToHost);
    double per_smsp_instr = (double)iters * CH * warps_per_sm / 4.0;
    printf("mode %d chains %d warps/SM %2d : %6.2f cycles per warp-instr per SMSP, %7.2f TFLOP/s (fma=2)\n",
           MODE, CH, warps_per_sm, (double)cyc / per_smsp_instr,
           2.0 * iters * CH * threads * (double)sms / (ms * 1e-3) / 1e12);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out, *in;
    cudaMalloc(&out, sizeof(double) * (sms * 1024 + 16));
    cudaMalloc(&in, sizeof(double) * 4096);
    cudaMemset(in, 0, sizeof(double) * 4096);
    printf("latency (1 warp per SMSP, 1 chain):\n");
    run<1, 0>(4, out, in, sms);
    run<1, 3>(4, out, in, sms);
    run<1, 4>(4, out, in, sms);
    printf("ILP sweep, 1 warp per SMSP:\n");
    run<2, 0>(4, out, in, sms);
    run<4, 0>(4, out, in, sms);
    run<8, 0>(4, out, in, sms);
    run<16, 0>(4, out, in, sms);
    printf("3 warps per SMSP (12 per SM):\n");
    run<1, 0>(12, out, in, sms);
    run<2, 0>(12, out, in, sms);
    run<5, 0>(12, out, in, sms);
    run<8, 0>(12, out, in, sms);
    printf("4 warps per SMSP (16 per SM):\n");
    run<1, 0>(16, out, in, sms);
    run<2, 0>(16, out, in, sms);
    run<4, 0>(16, out, in, sms);
    printf("8 warps per SMSP (32 per SM):\n");
    run<1, 0>(32, out, in, sms);
    run<2, 0>(32, out, in, sms);
    printf("operand shapes at 16 warps/SM x 8 chains:\n");
    run<8, 0>(16, out, in, sms);
    run<8, 1>(16, out, in, sms);
    run<8, 2>(16, out, in, sms);
    run<8, 3>(16, out, in, sms);
    run<8, 4>(16, out, in, sms);
    printf("accumulator forms at 16 warps/SM x 8 chains (8: one shared multiplicand, 9: shared by pairs, 10: all distinct):\n");
    run<8, 8>(16, out, in, sms);
    run<8, 9>(16, out, in, sms);
    run<8, 10>(16, out, in, sms);
    printf("mixing: 8 two-register DFMAs + {2,4,8 x (SHF,LOP3,IADD)} integer ops per chain, 16 warps/SM\n");
    run<8, 1>(16, out, in, sms);
    run<8, 5>(16, out, in, sms);
    run<8, 6>(16, out, in, sms);
    run<8, 7>(16, out, in, sms);
    return 0;
}
