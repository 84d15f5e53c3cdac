// Microbenchmark: HBM read bandwidth of 1-D TMA bulk copies (cp.async.bulk) from one producer thread per SM,
// as a function of copy size and ring depth. Rows of `COPY` bytes, consecutive rows `stride` bytes apart.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(64, 1) bulk_read(const char* __restrict__ x, long total_bytes, int copy, int per_stage, int stages, long stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    unsigned char* data = smem + 1024;
    const int stage_bytes = copy * per_stage;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const long ncopies = total_bytes / copy;
    const long nstage_total = ncopies / per_stage;
    long issued = 0, waited = 0;
    for (long st = blockIdx.x; st < nstage_total; st += gridDim.x) {
        if (issued - waited == stages) {   // oldest stage must land before its slot is reused
            const int s = waited % stages; const uint32_t ph = (waited / stages) & 1;
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(&full[s])), "r"(ph) : "memory");
            ++waited;
        }
        const int s = issued % stages;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(stage_bytes) : "memory");
        const char* src0 = x + st * (long)stage_bytes;
        for (int c = 0; c < per_stage; ++c) {
            const char* src = src0 + c * copy;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(data + s * stage_bytes + c * copy)), "l"(src), "r"(copy), "r"(s32(&full[s])) : "memory");
        }
        ++issued;
    }
    while (waited < issued) {
        const int s = waited % stages; const uint32_t ph = (waited / stages) & 1;
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(s32(&full[s])), "r"(ph) : "memory");
        ++waited;
    }
}
int main() {
    const long total = 460l << 20;
    char* x; cudaMalloc(&x, total); cudaMemset(x, 0, total);
    cudaFuncSetAttribute(bulk_read, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int copies[] = {512, 1024, 2048, 3072, 4096, 8192, 16384};
    for (int copy : copies) for (int inflight_kb : {32, 64, 128}) {
        const int per_stage = copy >= 8192 ? 1 : 8192 / copy;          // 8 KB stages
        const int stages = inflight_kb * 1024 / (copy * per_stage);
        const int smem = 1024 + stages * copy * per_stage;
        float best = 1e9;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(e0);
            bulk_read<<<148, 64, smem>>>(x, total, copy, per_stage, stages, copy);   // contiguous stream of copies
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("bulk copy %5d B x%2d per stage, %2d stages (%3d KB in flight): %.1f us %.0f GB/s  %s\n", copy, per_stage, stages, inflight_kb, best * 1e3,
               total / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
