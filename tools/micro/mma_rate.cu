// Issue / completion rate of tcgen05.mma kind::tf32 M128 x N x K8 on one SM, as the pointwise kernel issues it:
// groups of 6 MMAs into one accumulator followed by a tcgen05.commit, by one elected lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W;\n\t}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}

// mode 0: A from TMEM, 1: A from smem; alt: alternate between two accumulators per group; per_commit: MMAs per commit
__global__ void __launch_bounds__(128) mma_rate(int N, int mode, int alt, int per_commit, int groups, int issuers, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar_mid4[4], bar_end4[4];
    __shared__ uint32_t slot;
    float* sa = reinterpret_cast<float*>(sm);              // 128 x 16 fp32
    float* sb = sa + 128 * 16;                             // 256 x 16 fp32
    for (int i = threadIdx.x; i < (128 + 256) * 16; i += blockDim.x) sa[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&bar_mid4[i], 1); mbar_init(&bar_end4[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < issuers) {
        const uint32_t wi = threadIdx.x >> 5;
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t dA = make_desc(smem_u32(sa), 128, 512), dB = make_desc(smem_u32(sb), 128, 512);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            const uint32_t d = tmem + ((alt && (g & 1)) ? 256 : 0) + wi * 64;
            for (int i = 0; i < per_commit; ++i) {
                const uint64_t adv = (uint64_t)((i & 1) * 16);
                if (mode == 0) mma_ts(d, tmem + 480 + (i & 1) * 8, dB + adv, idesc, 1);
                else mma_ss(d, dA + adv, dB + adv, idesc, 1);
            }
            commit(&bar_mid4[wi]);
        }
        const long long t1 = clock64();
        commit(&bar_end4[wi]);
        mbar_wait(&bar_end4[wi], 0);
        const long long t2 = clock64();
        out[2 * wi] = t1 - t0; out[2 * wi + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    long long* out; cudaMallocManaged(&out, 64);
    const int smem = (128 + 256) * 16 * 4 + 1024;
    cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    printf("%5s %5s %4s %4s %4s | %10s %10s\n", "mode", "N", "alt", "/cmt", "thr", "issue clk/mma", "done clk/mma (per issuing thread, 600 MMAs each)");
    for (int mode = 0; mode < 2; ++mode)
        for (int N : {48})
            for (int issuers : {1, 2, 4})
                for (int pc : {1, 6, 24, 600}) {
                    const int groups = 600 / pc;
                    for (int rep = 0; rep < 2; ++rep) { mma_rate<<<1, 128, smem>>>(N, mode, 0, pc, groups, issuers, out); cudaDeviceSynchronize(); }
                    cudaError_t e = cudaGetLastError();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    printf("%5s %5d %4d %4d %4d | %10.1f %10.1f\n", mode ? "SS" : "TS", N, 0, pc, issuers, out[2 * (issuers - 1)] / 600.0, out[2 * (issuers - 1) + 1] / 600.0);
                }
    return 0;
}
