// capi.cu — the extern "C" surface of libbem_b200.so (include/bem_b200.h): argument validation, workspace carving,
// launch parameter packing. No torch types, no allocation, no synchronisation.
#include <cuda_runtime.h>

#include <cstdlib>

#include "bem_kernels.h"

namespace bem {

bool pdl_enabled() {
    static const bool on = [] {
        const char* v = getenv("BEM_NO_PDL");
        return !(v && atoi(v) != 0);
    }();
    return on;
}

int device_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}


static bool dtype_ok(int dt) { return dt == BEM_F32 || dt == BEM_F16 || dt == BEM_BF16; }

}  // namespace bem

using namespace bem;

extern "C" {

int bem_abi_version(void) { return BEM_ABI_VERSION; }

const char* bem_error_string(int code) {
    switch (code) {
        case BEM_OK: return "ok";
        case BEM_ERR_BAD_ARG: return "bem: bad argument (shape / dtype / NULL contract)";
        case BEM_ERR_WORKSPACE: return "bem: workspace missing or too small";
        case BEM_ERR_UNSUPPORTED: return "bem: configuration not built (see DESIGN.md)";
        default: return cudaGetErrorString((cudaError_t)code);
    }
}

int bem_scan_chunk_len(int dtype) { return dtype_ok(dtype) ? 32 * scan_items(dtype) : 0; }

int64_t bem_scan_workspace_bytes(int batch, int dim, int seqlen, int dstate, int dtype) {
    if (!dtype_ok(dtype) || batch <= 0 || dim <= 0 || seqlen <= 0 || dstate <= 0) return 0;
    const int CL = bem_scan_chunk_len(dtype);   // the finest tiling any kernel uses
    const int64_t nchunks = (seqlen + CL - 1) / CL;
    const int64_t ns = dstate < kMaxDstate ? dstate : kMaxDstate;   // larger dstate runs in passes of kMaxDstate states
    return kWsHeader + 2 * (int64_t)batch * dim * nchunks * ns * 16;   // aggregate + inclusive descriptors
}

int bem_scan_fwd(const BemScanFwdParams* q, void* stream_) {
    if (!q) return BEM_ERR_BAD_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!dtype_ok(q->dtype) || !(q->out_dtype == BEM_F32 || q->out_dtype == q->dtype)) return BEM_ERR_BAD_ARG;
    if (q->batch <= 0 || q->dim <= 0 || q->seqlen <= 0 || q->dstate <= 0 || q->n_groups <= 0) return BEM_ERR_BAD_ARG;
    if (q->dim % q->n_groups != 0) return BEM_ERR_BAD_ARG;
    if (q->dstate > 256) return BEM_ERR_BAD_ARG;   // MAX_DSTATE of the reference (selective_scan_oflex.cpp:190)
    if (!q->u || !q->delta || !q->A || !q->B || !q->C || !q->out) return BEM_ERR_BAD_ARG;
    const int64_t need = bem_scan_workspace_bytes(q->batch, q->dim, q->seqlen, q->dstate, q->dtype);
    if (!q->workspace || q->workspace_bytes < need || (reinterpret_cast<uintptr_t>(q->workspace) & 15)) return BEM_ERR_WORKSPACE;

    const int CL = bem_scan_chunk_len(q->dtype);
    ScanFwdArgs a{};
    a.u = q->u; a.delta = q->delta; a.Bm = q->B; a.Cm = q->C; a.A = q->A; a.D = q->D; a.bias = q->delta_bias;
    a.out = q->out; a.x = q->x;
    a.u_bs = q->u_bs; a.u_ds = q->u_ds; a.dl_bs = q->delta_bs; a.dl_ds = q->delta_ds;
    a.A_ds = q->A_ds; a.A_ns = q->A_ns;
    a.B_bs = q->B_bs; a.B_gs = q->B_gs; a.B_ns = q->B_ns;
    a.C_bs = q->C_bs; a.C_gs = q->C_gs; a.C_ns = q->C_ns;
    a.out_bs = q->out_bs; a.out_ds = q->out_ds;
    a.batch = q->batch; a.dim = q->dim; a.L = q->seqlen; a.N = q->dstate; a.G = q->n_groups; a.Dg = q->dim / q->n_groups;
    a.nxchunks = (q->seqlen + CL - 1) / CL;   // carries of `x`; the tile count is set by the launcher
    a.softplus = q->delta_softplus ? 1 : 0;
    if (q->dt_rank < 0 || (q->dt_rank > 0 && (!q->dt_weight || q->dtype != BEM_F32))) return BEM_ERR_BAD_ARG;
    a.R = q->dt_rank; a.dt_w = q->dt_weight; a.dl_gs = q->delta_gs;
    unsigned char* ws = reinterpret_cast<unsigned char*>(q->workspace);
    a.ticket = reinterpret_cast<unsigned int*>(ws);
    a.err = reinterpret_cast<unsigned int*>(ws + 4);
    a.desc = reinterpret_cast<uint4*>(ws + kWsHeader);
    return scan_fwd_dispatch(a, q->dtype, q->out_dtype, device_sm_count(), stream);
}

int bem_scan_bwd(const BemScanBwdParams* q, void* stream_) {
    if (!q) return BEM_ERR_BAD_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!dtype_ok(q->dtype) || !(q->dout_dtype == BEM_F32 || q->dout_dtype == q->dtype)) return BEM_ERR_BAD_ARG;
    if (q->batch <= 0 || q->dim <= 0 || q->seqlen <= 0 || q->dstate <= 0 || q->n_groups <= 0) return BEM_ERR_BAD_ARG;
    if (q->dim % q->n_groups != 0 || q->dstate > 256) return BEM_ERR_BAD_ARG;
    if (!q->u || !q->delta || !q->A || !q->B || !q->C || !q->dout || !q->du || !q->ddelta || !q->dA || !q->dB || !q->dC)
        return BEM_ERR_BAD_ARG;
    const int CL = bem_scan_chunk_len(q->dtype);
    const int nchunks = (q->seqlen + CL - 1) / CL;
    if (nchunks > 1 && !q->x) return BEM_ERR_BAD_ARG;   // selective_scan_oflex.cpp:315
    const int64_t need = bem_scan_workspace_bytes(q->batch, q->dim, q->seqlen, q->dstate, q->dtype);
    if (!q->workspace || q->workspace_bytes < need || (reinterpret_cast<uintptr_t>(q->workspace) & 15)) return BEM_ERR_WORKSPACE;

    ScanBwdArgs a{};
    a.u = q->u; a.delta = q->delta; a.Bm = q->B; a.Cm = q->C; a.A = q->A; a.D = q->D; a.bias = q->delta_bias;
    a.dout = q->dout; a.x = q->x; a.du = q->du; a.ddelta = q->ddelta;
    a.dA = q->dA; a.dB = q->dB; a.dC = q->dC; a.dD = q->D ? q->dD : nullptr; a.dbias = q->delta_bias ? q->ddelta_bias : nullptr;
    a.u_bs = q->u_bs; a.u_ds = q->u_ds; a.dl_bs = q->delta_bs; a.dl_ds = q->delta_ds;
    a.A_ds = q->A_ds; a.A_ns = q->A_ns;
    a.B_bs = q->B_bs; a.B_gs = q->B_gs; a.B_ns = q->B_ns;
    a.C_bs = q->C_bs; a.C_gs = q->C_gs; a.C_ns = q->C_ns;
    a.do_bs = q->dout_bs; a.do_ds = q->dout_ds;
    a.du_bs = q->du_bs; a.du_ds = q->du_ds; a.dd_bs = q->ddelta_bs; a.dd_ds = q->ddelta_ds;
    a.batch = q->batch; a.dim = q->dim; a.L = q->seqlen; a.N = q->dstate; a.G = q->n_groups; a.Dg = q->dim / q->n_groups;
    a.nchunks = nchunks;
    // Row splits: one CTA walks all rows of a (b, group, chunk) slab when that still fills the machine; otherwise the
    // group's rows are split over RS CTAs whose dB/dC partial sums meet in global atomics.
    const int sms = device_sm_count();
    const int64_t slabs = (int64_t)a.batch * a.G * a.nchunks;
    const int max_rs = (a.Dg + kScanWarps - 1) / kScanWarps;
    int rs = 1;
    while (rs < max_rs && slabs * rs < 4LL * sms) rs *= 2;
    if (rs > max_rs) rs = max_rs;
    a.RS = rs;
    a.rows_per_split = ((a.Dg + rs - 1) / rs + kScanWarps - 1) / kScanWarps * kScanWarps;   // multiple of the warp count
    a.RS = (a.Dg + a.rows_per_split - 1) / a.rows_per_split;
    a.RBS = a.rows_per_split / kScanWarps;
    a.ST = a.batch * a.G * a.RS;
    const int64_t total = (int64_t)a.nchunks * a.ST;
    if (total > 0x7fffffff) return BEM_ERR_UNSUPPORTED;
    a.total_tiles = (int)total;
    a.softplus = q->delta_softplus ? 1 : 0;
    a.atomic_bc = (a.RS > 1 || a.N > 1) ? 1 : 0;
    unsigned char* ws = reinterpret_cast<unsigned char*>(q->workspace);
    a.ticket = reinterpret_cast<unsigned int*>(ws);
    a.err = reinterpret_cast<unsigned int*>(ws + 4);
    a.desc = reinterpret_cast<uint4*>(ws + kWsHeader);
    return scan_bwd_dispatch(a, q->dtype, q->dout_dtype, sms, stream);
}

// SS2D core in one call. Default: the traversal-aware scan of ss2d_fused.cu (three launches, no materialised traversal).
// Composed form (BEM_SS2D_COMPOSED=1, or a dt_rank the fused kernels are not instantiated for): cross_scan(x),
// cross_scan(xdbl, one_by_one), scan with dt_proj fused, cross_merge — four launches with intermediates in the workspace.
static int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }
static bool ss2d_use_fused(int dstate, int dt_rank) {
    static const bool composed = [] {
        const char* v = getenv("BEM_SS2D_COMPOSED");
        return v && atoi(v) != 0;
    }();
    return !composed && ss2d_fused_supported(dstate, dt_rank);
}
struct Ss2dLayout {
    int64_t scan_ws, xs, xdbl_s, ys, fused, total;
};
static Ss2dLayout ss2d_layout(int batch, int d_inner, int H, int W, int dstate, int dt_rank) {
    const int64_t L = (int64_t)H * W, Cx = dt_rank + 2 * dstate;
    Ss2dLayout l;
    l.scan_ws = 0;
    if (ss2d_use_fused(dstate, dt_rank)) {
        l.xs = l.xdbl_s = l.ys = l.fused = 256;          // header kept (and left zero) so both forms may share one buffer
        l.total = l.fused + align256(ss2d_fused_workspace(batch, d_inner, H, W));
        return l;
    }
    l.xs = align256(bem_scan_workspace_bytes(batch, 4 * d_inner, (int)L, dstate, BEM_F32));
    l.xdbl_s = l.xs + align256((int64_t)batch * 4 * d_inner * L * 4);
    l.ys = l.xdbl_s + align256((int64_t)batch * 4 * Cx * L * 4);
    l.fused = l.total = l.ys + align256((int64_t)batch * 4 * d_inner * L * 4);
    return l;
}
int bem_ss2d_supported(int dstate, int dt_rank) {
    if (dstate != 1 || dt_rank <= 0) return 0;
    if (ss2d_use_fused(dstate, dt_rank)) return 2;
    return dt_rank <= kMaxDtRank ? 1 : 0;
}
int64_t bem_ss2d_workspace_bytes(int batch, int d_inner, int H, int W, int dstate, int dt_rank) {
    if (batch <= 0 || d_inner <= 0 || H <= 0 || W <= 0 || dstate <= 0 || dt_rank <= 0) return 0;
    return ss2d_layout(batch, d_inner, H, W, dstate, dt_rank).total;
}
int bem_ss2d_fwd(const BemSs2dFwdParams* p, void* stream) {
    if (!p || !p->x || !p->xdbl || !p->dt_weight || !p->A || !p->y) return BEM_ERR_BAD_ARG;
    if (p->batch <= 0 || p->d_inner <= 0 || p->H <= 0 || p->W <= 0 || p->dstate <= 0 || p->dt_rank <= 0) return BEM_ERR_BAD_ARG;
    const bool fused = ss2d_use_fused(p->dstate, p->dt_rank);
    if (p->dstate != 1 || (!fused && p->dt_rank > kMaxDtRank)) return BEM_ERR_UNSUPPORTED;
    const Ss2dLayout l = ss2d_layout(p->batch, p->d_inner, p->H, p->W, p->dstate, p->dt_rank);
    if (!p->workspace || p->workspace_bytes < l.total || (reinterpret_cast<uintptr_t>(p->workspace) & 255)) return BEM_ERR_WORKSPACE;
    unsigned char* ws = reinterpret_cast<unsigned char*>(p->workspace);
    if (fused) {
        Ss2dFusedArgs a{};
        a.x = p->x; a.xdbl = p->xdbl; a.dt_w = p->dt_weight; a.A = p->A; a.Ds = p->Dskip; a.bias = p->delta_bias; a.y = p->y;
        a.B = p->batch; a.D = p->d_inner; a.H = p->H; a.W = p->W; a.R = p->dt_rank; a.softplus = p->delta_softplus ? 1 : 0;
        return ss2d_fused_dispatch(a, ws + l.fused, (cudaStream_t)stream);
    }
    float* xs = reinterpret_cast<float*>(ws + l.xs);
    float* xdbl_s = reinterpret_cast<float*>(ws + l.xdbl_s);
    float* ys = reinterpret_cast<float*>(ws + l.ys);
    const int64_t L = (int64_t)p->H * p->W;
    const int D = p->d_inner, N = p->dstate, R = p->dt_rank, Cx = R + 2 * N;

    BemCsmParams c{};
    c.B = p->batch; c.C = D; c.H = p->H; c.W = p->W; c.dtype = BEM_F32;
    c.img_channel_first = 1; c.seq_channel_first = 1; c.one_by_one = 0; c.scans = 0;
    c.src = p->x; c.dst = xs;
    int rc = bem_cross_scan(&c, stream);                       // xs : (B, 4, D, L)
    if (rc) return rc;
    c.C = Cx; c.one_by_one = 1; c.src = p->xdbl; c.dst = xdbl_s;
    rc = bem_cross_scan(&c, stream);                           // xdbl_s : (B, 4, Cx, L), direction k traverses its own block
    if (rc) return rc;

    BemScanFwdParams q{};
    q.batch = p->batch; q.dim = 4 * D; q.seqlen = (int)L; q.dstate = N; q.n_groups = 4;
    q.dtype = BEM_F32; q.out_dtype = BEM_F32; q.delta_softplus = p->delta_softplus;
    q.u = xs; q.delta = xdbl_s; q.A = p->A; q.B = xdbl_s + (int64_t)R * L; q.C = xdbl_s + (int64_t)(R + N) * L;
    q.D = p->Dskip; q.delta_bias = p->delta_bias; q.out = ys; q.x = nullptr;
    q.u_bs = 4 * (int64_t)D * L; q.u_ds = L;
    q.delta_bs = 4 * (int64_t)Cx * L; q.delta_gs = (int64_t)Cx * L; q.delta_ds = L;
    q.A_ds = N; q.A_ns = 1;
    q.B_bs = q.C_bs = 4 * (int64_t)Cx * L; q.B_gs = q.C_gs = (int64_t)Cx * L; q.B_ns = q.C_ns = L;
    q.out_bs = 4 * (int64_t)D * L; q.out_ds = L;
    q.workspace = ws + l.scan_ws; q.workspace_bytes = l.xs;
    q.dt_rank = R; q.dt_weight = p->dt_weight;
    rc = bem_scan_fwd(&q, stream);                             // ys : (B, 4*D, L)
    if (rc) return rc;

    c.C = D; c.one_by_one = 0; c.src = ys; c.dst = p->y;
    return bem_cross_merge(&c, stream);                        // y : (B, D, L)
}

}  // extern "C"
