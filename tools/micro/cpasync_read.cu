// Microbenchmark: how much HBM bandwidth can one CTA per SM pull with cp.async (LDGSTS) 512-byte rows, as a function
// of producer warps and groups in flight; and with 1-D bulk TMA copies of the same rows.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int TILE_PX = 128, KC = 16;
template <int DEPTH>
__global__ void __launch_bounds__(1024, 1) cpasync_read(const float* __restrict__ x, int C, long P, int rows_per_warp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const long ntiles = P / TILE_PX;
    const int nk = C / KC;
    // each warp owns `rows_per_warp` rows of every chunk; warps with row0 >= KC idle
    const int row0 = warp * rows_per_warp;
    if (row0 >= KC) return;
    unsigned char* mine = smem + warp * (DEPTH + 1) * rows_per_warp * 512;
    uint32_t slot = 0;
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const float* base = x + t * TILE_PX + lane * 4;
        for (int kc = 0; kc < nk; ++kc) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(mine + slot * rows_per_warp * 512) + lane * 16;
            const float* src = base + (long)(kc * KC + row0) * P;
            for (int r = 0; r < rows_per_warp; ++r, src += P)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + r * 512), "l"(src) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH) : "memory");
            slot = slot == DEPTH ? 0 : slot + 1;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
template <int DEPTH>
void run(const float* x, int C, long P, int warps, int rows_per_warp) {
    const int smem = warps * (DEPTH + 1) * rows_per_warp * 512;
    cudaFuncSetAttribute(cpasync_read<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(e0);
        cpasync_read<DEPTH><<<148, warps * 32, smem>>>(x, C, P, rows_per_warp);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("cp.async C %4d warps %2d rows/warp %2d depth %2d (%3d KB in flight): %.1f us  %.0f GB/s  %s\n", C, warps, rows_per_warp, DEPTH,
           warps * (DEPTH + 1) * rows_per_warp * 512 / 1024, best * 1e3, C * P * 4.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const long P = 240000;
    float* x; cudaMalloc(&x, 640 * P * 4); cudaMemset(x, 0, 640 * P * 4);
    const int C = 160;
    run<2>(x, C, P, 1, 16); run<6>(x, C, P, 1, 16); run<12>(x, C, P, 1, 16);
    run<2>(x, C, P, 2, 8);  run<6>(x, C, P, 2, 8);  run<12>(x, C, P, 2, 8);  run<24>(x, C, P, 2, 8);
    run<2>(x, C, P, 4, 4);  run<6>(x, C, P, 4, 4);  run<12>(x, C, P, 4, 4);  run<24>(x, C, P, 4, 4);
    run<6>(x, C, P, 8, 2);  run<12>(x, C, P, 8, 2); run<24>(x, C, P, 8, 2);
    run<6>(x, C, P, 16, 1); run<12>(x, C, P, 16, 1); run<24>(x, C, P, 16, 1);
    return 0;
}
