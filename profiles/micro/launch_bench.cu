// Micro-benchmark: cost of one kernel node in a replayed CUDA graph on B200 as a function of the
// parameter-block size and grid size (the fused loss passes a ~1.3 KB __grid_constant__ struct).
#include <cstdio>
#include <cuda_runtime.h>
template <int N> struct P { int v[N]; };
template <int N> __global__ void k(const __grid_constant__ P<N> p, int* out) {
    if (p.v[N - 1] == 12345 && threadIdx.x == 0) out[blockIdx.x] = p.v[0];
}
template <int N> float run(int grid, int block, int n, cudaStream_t st, int* out) {
    P<N> p; for (int i = 0; i < N; ++i) p.v[i] = i;
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < n; ++i) k<N><<<grid, block, 0, st>>>(p, out);
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int r = 0; r < 20; ++r) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st); cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / (20 * n);
}
int main() {
    cudaStream_t st; cudaStreamCreate(&st);
    int* out; cudaMalloc(&out, 1 << 20);
    const int grids[] = {1, 148, 444, 1184};
    for (int gi = 0; gi < 4; ++gi) {
        const int g = grids[gi];
        printf("grid %4d x 256: params 16B %.2f us | 256B %.2f us | 1.3KB %.2f us | 4KB %.2f us\n", g,
               run<4>(g, 256, 64, st, out), run<64>(g, 256, 64, st, out), run<336>(g, 256, 64, st, out),
               run<1000>(g, 256, 64, st, out));
    }
    return 0;
}
