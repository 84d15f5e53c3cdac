// Microbenchmark: HBM read bandwidth of the (channels, pixels) tile access pattern of the pointwise kernel.
// Each CTA walks pixel tiles (tile = blockIdx.x + i * gridDim.x); per tile it reads C rows of TILE_PX floats (row stride P).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int TILE_PX>
__global__ void __launch_bounds__(512) tile_read(const float* __restrict__ x, float* __restrict__ sink, int C, long P) {
    const long ntiles = P / TILE_PX;
    float acc = 0.f;
    constexpr int V = TILE_PX / 4;                 // float4 per row
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const float* base = x + t * TILE_PX;
        for (int i = threadIdx.x; i < C * V; i += blockDim.x) {
            const int c = i / V, v = i - c * V;
            const float4 q = __ldcs(reinterpret_cast<const float4*>(base + (long)c * P) + v);
            acc += q.x + q.y + q.z + q.w;
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}
template <int TILE_PX>
void run(const float* x, float* sink, int C, long P, int ctas_per_sm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        tile_read<TILE_PX><<<148 * ctas_per_sm, 512>>>(x, sink, C, P);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("C %4d tile %5d px ctas/sm %d: %.1f us  %.0f GB/s\n", C, TILE_PX, ctas_per_sm, best * 1e3, C * P * 4.0 / best / 1e6);
}
int main() {
    const long P = 240000;
    float *x, *sink; cudaMalloc(&x, 640 * P * 4); cudaMalloc(&sink, 4); cudaMemset(x, 0, 640 * P * 4);
    for (int C : {160, 640}) for (int k : {1, 2, 4}) {
        run<128>(x, sink, C, P, k); run<256>(x, sink, C, P, k); run<512>(x, sink, C, P, k); run<1024>(x, sink, C, P, k);
    }
    return 0;
}
