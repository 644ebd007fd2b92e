// Micro-benchmark: issue rate of packed FFMA2 vs scalar FFMA on sm_100a (B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2_bench.cu -o ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096;
template <int MODE>
__global__ void k(float* out, float a, float b) {
    float2 x[8];
    for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 A = make_float2(a, a), Bv = make_float2(b, b);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
            else if (MODE == 1) x[i] = __ffma2_rn(x[i], A, Bv);
            else if (MODE == 2) { x[i].x = x[i].x + b; x[i].y = x[i].y + b; }
            else if (MODE == 3) x[i] = __fadd2_rn(x[i], Bv);
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops_elem = (double)148 * 8 * 256 * ITER * 16;  // fp32 element-ops (fma counted once)
    printf("%-8s %.3f ms  %.1f G elem-op/s  (per SM per clk @1.965GHz: %.1f)\n", name, ms, flops_elem / ms / 1e6,
           flops_elem / ms / 1e6 / 148 / 1.965);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("FFMA", out); run<1>("FFMA2", out); run<2>("FADD", out); run<3>("FADD2", out);
    return 0;
}
