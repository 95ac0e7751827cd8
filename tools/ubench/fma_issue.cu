// Single-warp issue-rate microbenchmark for the FP32 pipe of sm_100a: how many cycles does ONE warp need per FFMA / FMUL /
// FADD / FFMA2 when 8 independent dependency chains are available (i.e. latency is hidden and only the pipe's issue
// rate binds)?  Motivation: the rollout kernel runs one warp per scheduler and 60 % of its instructions are FP32.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_issue fma_issue.cu && ./fma_issue
#include <cstdio>
#include <cuda_runtime.h>

#define N_ITER 4096
#define CHAINS 8

template <int MODE>
__global__ void bench(float* out, long long* cycles, float a, float b, int active_lanes) {
    if ((int)threadIdx.x % 32 >= active_lanes) return;
    float x[CHAINS]; float2 y[CHAINS];
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) { x[j] = threadIdx.x * 0.001f + j; y[j] = make_float2(x[j], x[j] + 0.5f); }
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_ITER; ++i) {
#pragma unroll
        for (int j = 0; j < CHAINS; ++j) {
            if (MODE == 0) x[j] = fmaf(x[j], a, b);                 // FFMA, 3 register operands
            if (MODE == 1) x[j] = x[j] * a;                          // FMUL
            if (MODE == 2) x[j] = x[j] + b;                          // FADD
            if (MODE == 3) y[j] = __ffma2_rn(y[j], a2, b2);          // FFMA2 (two FMAs per instruction)
            if (MODE == 4) x[j] = fmaf(x[j], 1.0001f, b);            // FFMA with an immediate operand
            if (MODE == 5) { x[j] = fmaf(x[j], a, b); y[j].x = y[j].x + b2.y; }   // FFMA + FADD interleaved
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) s += x[j] + y[j].x + y[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x % 32 == 0) cycles[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_block, int lanes, int per_iter) {
    float* out; long long* cyc;
    cudaMalloc(&out, 1024 * sizeof(float)); cudaMalloc(&cyc, 64 * sizeof(long long));
    bench<MODE><<<1, 32 * warps_per_block>>>(out, cyc, 1.0001f, 0.0001f, lanes);
    bench<MODE><<<1, 32 * warps_per_block>>>(out, cyc, 1.0001f, 0.0001f, lanes);
    long long h[64];
    cudaMemcpy(h, cyc, sizeof(long long) * warps_per_block, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-34s warps/CTA=%d lanes=%2d  cycles per instruction = %.3f  (%s)\n", name, warps_per_block, lanes,
           (double)h[0] / ((double)N_ITER * per_iter), cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 4, 8}) {
        run<0>("FFMA r,r,r", w, 32, CHAINS);
        run<1>("FMUL", w, 32, CHAINS);
        run<2>("FADD", w, 32, CHAINS);
        run<3>("FFMA2 (2 fma / instr)", w, 32, CHAINS);
        run<4>("FFMA r,imm,r", w, 32, CHAINS);
        run<5>("FFMA + FADD pairs", w, 32, 2 * CHAINS);
    }
    run<0>("FFMA r,r,r half warp", 1, 16, CHAINS);
    run<3>("FFMA2 half warp", 1, 16, CHAINS);
    return 0;
}
