// bem_kernels.h — internal declarations shared by the .cu files of libbem_b200.so (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bem_b200.h"

namespace bem {

// Programmatic dependent launch. Every kernel of the library is launched with the programmatic-stream-serialization
// attribute and starts with pdl_trigger() — the next kernel in the stream may be scheduled as soon as every CTA of this one
// is running — and pdl_wait(), which returns once the previous kernel has completed and its writes are visible. Nothing
// before pdl_wait() may touch global memory. What this buys is the launch latency and the CTA ramp of each kernel (its
// blocks take the place of the previous kernel's blocks as those retire): ~350 kernels of 5-100 us per Monte-Carlo sample.
// BEM_NO_PDL=1 launches without the attribute (the two instructions are then no-ops).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// scan tiling: consumer warps per CTA, positions per lane (=> tile length 32*ITEMS)
constexpr int kScanWarps = 8;
constexpr int kItemsF32 = 12;        // 48 B per lane: conflict-free LDS.128 with a blocked lane layout
constexpr int kItems16 = 16;         // 32 B per lane for fp16 / bf16 (2-way LDS conflict, but fits 2 CTAs/SM in registers)
constexpr int kFwdItemsF32N1 = 24;   // forward, fp32, dstate 1: 768-position tiles amortise the per-tile overhead
constexpr int kCarryF32 = 32 * kItemsF32;   // positions per carry chunk of `x` (fp32): the backward kernel's tile
constexpr int kCarry16 = 32 * kItems16;     // ... for fp16 / bf16
constexpr int kMaxDstate = 16;       // states staged per pass (B/C chunk lives in shared memory)

inline int scan_items(int dtype) { return dtype == BEM_F32 ? kItemsF32 : kItems16; }

// workspace layout: [0,4) ticket, [4,8) error word, [128, ...) look-back descriptors (16 B each): aggregates, then inclusives
constexpr int64_t kWsHeader = 128;

constexpr int kMaxDtRank = 8;   // fused dt_proj: low-rank rows per group that fit the delta halves of a tile's row slots

struct ScanFwdArgs {
    const void* u;
    const void* delta;
    const void* Bm;
    const void* Cm;
    const float* A;
    const float* D;
    const float* bias;
    void* out;
    float* x;
    int64_t u_bs, u_ds, dl_bs, dl_ds, A_ds, A_ns, B_bs, B_gs, B_ns, C_bs, C_gs, C_ns, out_bs, out_ds;
    int batch, dim, L, N, G, Dg;
    int nchunks;       // tiles per row = ceil(L / CL)            (filled by the launcher)
    int nxchunks;      // carry chunks per row of `x` = ceil(L / kCarry)
    int RB;            // row blocks (of kScanWarps rows) per group
    int RT;            // row tiles per chunk = batch * G * RB
    int total_tiles;   // nchunks * RT
    int softplus;
    int stages;
    int lb_dynamic;    // A/B knob: classic timing-dependent look-back instead of the deterministic one
    int trace;         // record CTA 0's stage timeline (debug, tools/trace_scan.py)
    int R;             // fused dt_proj rank (0: `delta` is given per channel row)
    const float* dt_w; // (dim, R) dt_proj weight, row-major
    int64_t dl_gs;     // group stride of the low-rank delta (B, G, R, L); dl_ds is then the stride between its R rows
    uint4* desc;
    uint4* desc_incl;
    unsigned int* ticket;
    unsigned int* err;
};

struct ScanBwdArgs {
    const void* u;
    const void* delta;
    const void* Bm;
    const void* Cm;
    const float* A;
    const float* D;
    const float* bias;
    const void* dout;
    const float* x;
    void* du;
    void* ddelta;
    float* dA;
    float* dB;
    float* dC;
    float* dD;
    float* dbias;
    int64_t u_bs, u_ds, dl_bs, dl_ds, A_ds, A_ns, B_bs, B_gs, B_ns, C_bs, C_gs, C_ns, do_bs, do_ds, du_bs, du_ds, dd_bs,
        dd_ds;
    int batch, dim, L, N, G, Dg;
    int nchunks;
    int RS;            // row splits per group (CTAs sharing one (b, g, chunk) dB/dC slab)
    int rows_per_split;
    int RBS;           // row blocks per split
    int ST;            // super tiles per chunk = batch * G * RS
    int total_tiles;   // nchunks * ST
    int softplus;
    int stages;
    int atomic_bc;     // 1: dB/dC accumulated with atomics (RS > 1 or general dstate), 0: plain stores
    uint4* desc;
    uint4* desc_incl;
    unsigned int* ticket;
    unsigned int* err;
};

// traversal-aware SS2D core (ss2d_fused.cu)
struct Ss2dFusedArgs {
    const float* x;      // (B, D, H, W)
    const float* xdbl;   // (B, 4, R + 2, H, W): x_proj output in image order, [dt rows | B | C] per direction
    const float* dt_w;   // (4 * D, R)
    const float* A;      // (4 * D) (dstate 1)
    const float* Ds;     // (4 * D) or nullptr
    const float* bias;   // (4 * D) or nullptr
    float* y;            // (B, D, H, W)
    float2* agg;         // segment maps (P, V)
    float* hin;          // state entering each segment
    int B, D, H, W, R, softplus;
    int NTH, NTW;        // tiles along H / W
};
bool ss2d_fused_supported(int dstate, int dt_rank);
int64_t ss2d_fused_workspace(int B, int D, int H, int W);
int ss2d_fused_dispatch(Ss2dFusedArgs a, void* workspace, cudaStream_t stream);

int scan_fwd_dispatch(const ScanFwdArgs& a, int dtype, int out_dtype, int sm_count, cudaStream_t stream);
int scan_fwd_deferred_dispatch(const ScanFwdArgs& a, int sm_count, cudaStream_t stream);   // fp32, N = 1: deferred-finish schedule
int scan_bwd_dispatch(ScanBwdArgs& a, int dtype, int dout_dtype, int sm_count, cudaStream_t stream);
// dstate >= 2: one CTA per channel row, chunks walked in order (scan_rows.cu)
bool scan_rows_preferred(int batch, int dim, int N, int sm_count);
int scan_rows_fwd_dispatch(const ScanFwdArgs& a, int dtype, int out_dtype, int sm_count, cudaStream_t stream);
int scan_rows_bwd_dispatch(const ScanBwdArgs& a, int dtype, int dout_dtype, int sm_count, cudaStream_t stream);

int bayes_pointwise_tc_launch(const BemBayesPointwiseParams& p, cudaStream_t stream);
int64_t bayes_pointwise_tc_workspace(int n_samples, int cin, int cout);
int64_t bayes_pointwise_pack_table_bytes(int n);
int bayes_pointwise_pack_table(const BemBayesPointwiseParams* params, int n, void* table_host, int32_t* total_blocks);
int bayes_pointwise_pack_run(const void* table_dev, int n, int total_blocks, cudaStream_t stream);

int device_sm_count();

}  // namespace bem
