/*
 * bem_b200.h — C-ABI of libbem_b200.so, the sm_100a implementation of the Bayesian-Enhancement-Model
 * hot path (SS2D selective scan + CrossScan/CrossMerge + reparameterised Bayesian layers + MC selection).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in `_host`
 *   - no entry point allocates, synchronises or throws; each returns 0 or a cudaError_t / BEM_ERR_* code
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - strides are in ELEMENTS and 64-bit (the reference's SSMParamsBase uses uint32 strides and overflows
 *     at B*KD*L >= 2^32, kernels/selective_scan/csrc/selective_scan/selective_scan.h:27)
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference root).
 */
#ifndef BEM_B200_H_
#define BEM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BEM_ABI_VERSION 14

/* element types of u/delta/B/C/x-activations */
enum { BEM_F32 = 0, BEM_F16 = 1, BEM_BF16 = 2 };

/* error codes outside the cudaError_t range */
enum {
    BEM_OK = 0,
    BEM_ERR_BAD_ARG = 10001,      /* shape/dtype/NULL contract violated                     */
    BEM_ERR_WORKSPACE = 10002,    /* workspace missing or too small                          */
    BEM_ERR_UNSUPPORTED = 10003   /* legal in the reference, not built here  */
};

int bem_abi_version(void);
/* human-readable name for a return code (static storage) */
const char* bem_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * Selective scan, boundary form.
 * Replaces selective_scan_cuda_oflex.fwd / .bwd
 *   (kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_oflex.cpp:157-243, 245-358)
 * and the kernels behind them (cusoflex/selective_scan_fwd_kernel_oflex.cuh:67-211,
 * cusoflex/selective_scan_bwd_kernel_oflex.cuh:73-322).
 *
 *   u, delta : (batch, dim, seqlen)           dtype `dtype`, last-dim stride 1
 *   A        : (dim, dstate)                  fp32
 *   B, C     : (batch, n_groups, dstate, seqlen)  dtype `dtype`, last-dim stride 1, dim % n_groups == 0
 *   D, delta_bias : (dim) fp32 or NULL
 *   out      : (batch, dim, seqlen)           fp32 (out_dtype == BEM_F32, the "oflex" mode) or `dtype`
 *   x        : (batch, dim, n_chunks, 2*dstate) fp32 contiguous, n_chunks = ceil(seqlen / bem_scan_chunk_len(dtype)).
 *              x[b,d,c,2n]   = prod_{t < end(c)} exp(delta_t A_n)   (running decay from t = 0)
 *              x[b,d,c,2n+1] = h_n at the last position of chunk c   (the recurrence state)
 *              so `last_state = x[:, :, -1, 1::2]` holds exactly as in
 *              kernels/selective_scan/test_selective_scan.py:79. The reference fixes the chunk at 2048;
 *              here the chunk length depends on the element type (the tensor is opaque to callers).
 *   workspace: bem_scan_workspace_bytes() bytes, 16-byte aligned, ZERO-FILLED BY THE CALLER BEFORE ITS FIRST USE and
 *              from then on owned by the scan entry points of one stream (they re-arm it at the end of every launch:
 *              ticket counter back to zero, descriptor epoch advanced — no per-launch clearing, which keeps a launch a
 *              single kernel node under CUDA-graph capture). Word 1 is a sticky watchdog flag (0 = clean).
 * ---------------------------------------------------------------------------------------------- */
typedef struct BemScanFwdParams {
    int32_t batch, dim, seqlen, dstate, n_groups;
    int32_t dtype;           /* BEM_F32 | BEM_F16 | BEM_BF16 */
    int32_t out_dtype;       /* BEM_F32 or == dtype          */
    int32_t delta_softplus;  /* 0 | 1                        */
    const void* u;
    const void* delta;
    const float* A;
    const void* B;
    const void* C;
    const float* D;          /* may be NULL */
    const float* delta_bias; /* may be NULL */
    void* out;
    float* x;                /* may be NULL: carries are then not returned */
    int64_t u_bs, u_ds;          /* batch / dim strides of u      */
    int64_t delta_bs, delta_ds;
    int64_t A_ds, A_ns;
    int64_t B_bs, B_gs, B_ns;    /* batch / group / dstate strides */
    int64_t C_bs, C_gs, C_ns;
    int64_t out_bs, out_ds;
    void* workspace;
    int64_t workspace_bytes;
    /* Fused dt_proj (SS2Dv2.forward_corev2, basicsr/vmamba/models/vmamba.py:660-661: `dts = F.conv1d(dts, dt_projs_weight,
     * groups=K)` feeding the scan). dt_rank > 0: `delta` is the LOW-RANK dt of shape (batch, n_groups, dt_rank, seqlen)
     * with strides delta_bs (batch), delta_gs (group), delta_ds (rank row), and the kernel forms
     * delta[b, d, l] = sum_r dt_weight[d][r] * delta_lowrank[b, g(d), r, l] itself — the (batch, dim, seqlen) delta tensor is
     * never written or read. fp32, dstate = 1, dt_rank <= 8 (else BEM_ERR_UNSUPPORTED: run the projection separately). */
    int32_t dt_rank;             /* 0: delta is given per channel row (the plain selective_scan_fn contract) */
    const float* dt_weight;      /* (dim, dt_rank) row-major, or NULL */
    int64_t delta_gs;
} BemScanFwdParams;

typedef struct BemScanBwdParams {
    int32_t batch, dim, seqlen, dstate, n_groups;
    int32_t dtype;           /* u/delta/B/C and du/ddelta */
    int32_t dout_dtype;      /* BEM_F32 or == dtype       */
    int32_t delta_softplus;
    const void* u;
    const void* delta;
    const float* A;
    const void* B;
    const void* C;
    const float* D;          /* may be NULL */
    const float* delta_bias; /* may be NULL */
    const void* dout;        /* (batch, dim, seqlen), last-dim stride 1 */
    const float* x;          /* carries written by bem_scan_fwd; may be NULL iff n_chunks == 1 */
    void* du;                /* (batch, dim, seqlen) dtype, contiguous rows */
    void* ddelta;
    float* dA;               /* (dim, dstate) fp32, ACCUMULATED INTO: caller zero-fills */
    float* dB;               /* (batch, n_groups, dstate, seqlen) fp32 contiguous, accumulated into: caller zero-fills */
    float* dC;
    float* dD;               /* (dim) fp32 accumulated into, or NULL */
    float* ddelta_bias;      /* (dim) fp32 accumulated into, or NULL */
    int64_t u_bs, u_ds;
    int64_t delta_bs, delta_ds;
    int64_t A_ds, A_ns;
    int64_t B_bs, B_gs, B_ns;
    int64_t C_bs, C_gs, C_ns;
    int64_t dout_bs, dout_ds;
    int64_t du_bs, du_ds;
    int64_t ddelta_bs, ddelta_ds;
    void* workspace;
    int64_t workspace_bytes;
} BemScanBwdParams;

/* positions per carry chunk for an element type (multiple of 32) */
int bem_scan_chunk_len(int dtype);
/* scratch bytes needed by bem_scan_fwd / bem_scan_bwd for these sizes */
int64_t bem_scan_workspace_bytes(int batch, int dim, int seqlen, int dstate, int dtype);
int bem_scan_fwd(const BemScanFwdParams* p, void* stream);
int bem_scan_bwd(const BemScanBwdParams* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CrossScan / CrossMerge (four-direction traversal).
 * Replaces triton_cross_scan_flex and its torch twin
 *   (basicsr/vmamba/models/csm_triton.py:22-85 torch, :278-390 Triton, API :491-505).
 *
 * Image side  : one_by_one == 0 : (B, C, H, W)   or channel-last (B, H, W, C)
 *               one_by_one == 1 : (B, 4, C, H, W) or channel-last (B, H, W, 4, C)
 * Sequence side: (B, 4, C, L) or channel-last (B, L, 4, C), L = H*W
 * scans: 0 = cross2d (k0 row-major, k1 column-major, k2 = flip k0, k3 = flip k1), 1 = unidirectional
 *        (4 copies of k0), 2 = bidirectional (k0,k0,flip,flip)                 (csm_triton.py:25-34)
 * bem_cross_scan  : image -> sequences (a copy)
 * bem_cross_merge : sequences -> image; one_by_one == 0 sums the four directions (csm_triton.py:60-67)
 * All tensors contiguous, same dtype.
 * ---------------------------------------------------------------------------------------------- */
typedef struct BemCsmParams {
    int32_t B, C, H, W;
    int32_t dtype;
    int32_t img_channel_first;   /* layout of the image-side tensor    */
    int32_t seq_channel_first;   /* layout of the sequence-side tensor */
    int32_t one_by_one;
    int32_t scans;               /* 0 | 1 | 2 */
    const void* src;
    void* dst;
} BemCsmParams;
int bem_cross_scan(const BemCsmParams* p, void* stream);
int bem_cross_merge(const BemCsmParams* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SS2D core in one call: the part of SS2Dv2.forward_corev2 after x_proj (basicsr/vmamba/models/vmamba.py:656-684, scan_mode
 * "cross2d", `no_einsum`): split dt/B/C -> dt_proj -> cross_scan -> selective scan -> cross_merge.
 * x_proj is pointwise in l and commutes with the traversal, so its output is taken in IMAGE order, one
 * (dt_rank + 2*dstate)-channel block per direction — produced by ONE bem_bayes_pointwise call on x with the concatenated
 * x_proj weights (bem_b200/ss2d.py). This entry point then runs the traversal-aware scan (csrc/ss2d_fused.cu): the cross-scan
 * gather and the cross-merge scatter are part of the scan itself — every direction reads x and xdbl where they lie (k0 / k2
 * walk image rows forward / backward, k1 / k3 image columns), the four results of a pixel are summed on chip in the
 * reference's association (y0 + y2) + (y1 + y3) (csm_triton.py:60-62) and y is written once. No (B, 4, D, L) tensor exists.
 * Three launches (tile maps, carry scan, tile outputs) on the caller's stream, 12 bytes of workspace per 32-pixel segment.
 *   x         : (B, D, H, W)                         fp32
 *   xdbl      : (B, 4, dt_rank + 2*dstate, H, W)     fp32, channel order [dt | B | C] per direction
 *   dt_weight : (4*D, dt_rank) fp32; A : (4*D, dstate) fp32; Dskip, delta_bias : (4*D) fp32 or NULL
 *   y         : (B, D, H*W)  fp32 = sum over the 4 directions, un-traversed (before out_norm)
 * fp32, dstate = 1. bem_ss2d_supported(dstate, dt_rank) tells whether the call is built for a configuration: dt_rank 3 / 5 / 10
 * (the BEM levels) run the traversal-aware kernels; other ranks <= 8 (and everything under BEM_SS2D_COMPOSED=1, the A/B knob)
 * run the composed form bem_cross_scan x2 -> bem_scan_fwd (dt_proj fused) -> bem_cross_merge with intermediates in `workspace`;
 * anything else returns BEM_ERR_UNSUPPORTED (compose the entry points with a separate dt_proj).
 * workspace : bem_ss2d_workspace_bytes() bytes, 256-byte aligned, ZERO-FILLED BY THE CALLER BEFORE ITS FIRST USE (the composed
 *             form starts with the scan look-back workspace, see bem_scan_fwd), then owned by the calls of one stream.
 * ---------------------------------------------------------------------------------------------- */
typedef struct BemSs2dFwdParams {
    int32_t batch, d_inner, H, W, dstate, dt_rank;
    int32_t delta_softplus;
    const float* x;
    const float* xdbl;
    const float* dt_weight;
    const float* A;
    const float* Dskip;
    const float* delta_bias;
    float* y;
    void* workspace;
    int64_t workspace_bytes;
} BemSs2dFwdParams;
int bem_ss2d_supported(int dstate, int dt_rank);   /* 0: not built; 1: composed form; 2: traversal-aware kernels */
int64_t bem_ss2d_workspace_bytes(int batch, int d_inner, int H, int W, int dstate, int dt_rank);
int bem_ss2d_fwd(const BemSs2dFwdParams* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Bayesian reparameterised layers.
 * Replaces Conv2dReparameterization / Linear2dReparameterization / LinearReparameterization
 *   (basicsr/bayesian/conv.py:91-128, basicsr/bayesian/linear.py:67-104, 165-203):
 *     sigma = log1p(exp(rho)); w = mu + sigma * eps; out = conv(x, w, b)
 *
 * bem_bayes_sample : w[s, i] = mu[i] + log1p(exp(rho[i])) * eps[s, i]      (conv.py:106-107)
 *     eps == NULL  -> eps is generated in-kernel: Philox4x32-10 keyed (seed, stream_id), counter
 *                     (sample0 + s, i / 4), Box-Muller; element i takes lane i % 4.  The generator is restated in
 *                     oracle/philox.py so the same eps can be fed to the reference layer.
 *     eps_out != NULL -> the eps used is also written there (what the reference leaves in eps_weight).
 * bem_bayes_pointwise : 1x1 convolution / Linear2d with per-sample weights on tcgen05 (3xTF32: fp32-accurate)
 *     x : (S*Bx, Cin, P)  w : (S or 1, Cout, Cin)  bias : (S or 1, Cout) or NULL  out : (S*Bx, Cout, P), fp32
 *     mu/sigma/eps given instead of w  -> the sample step w = mu + sigma * eps is fused into the weight pack;
 *     ln_gamma given -> the LayerNorm2d that precedes the layer (vmamba.py:59-64, eps = ln_eps) is fused in as well;
 *     residual given -> out = residual + conv (the block's skip connection); prepacked -> weights packed by an earlier call.
 * bem_bayes_depthwise : depthwise KxK (groups == channels, stride 1, dilation 1, zero padding K/2), K in {3}
 *     x : (S*Bx, C, H, W)  w : (S or 1, C, K, K)  bias : (S or 1, C) or NULL; `act` fuses the SiLU / gated GELU that follows
 * ---------------------------------------------------------------------------------------------- */
typedef struct BemBayesSampleParams {
    int64_t numel;       /* elements of one weight tensor */
    int32_t n_samples;   /* S */
    const float* mu;
    const float* rho;    /* NULL -> deterministic: w = mu (conv.py:118-128) */
    const float* eps;    /* (S, numel) or NULL */
    float* w;            /* (S, numel) */
    float* eps_out;      /* (S, numel) or NULL */
    uint64_t seed;
    uint64_t stream_id;  /* layer / tensor id */
    int64_t sample0;     /* global index of sample 0 of this call */
} BemBayesSampleParams;
int bem_bayes_sample(const BemBayesSampleParams* p, void* stream);

/* One launch that samples every Bayesian tensor of a network for one Monte-Carlo draw: the per-layer
 * `eps.normal_(); w = mu + sigma * eps` pairs of one forward (conv.py:105-111 x ~60 layers) collapsed into a single
 * kernel. Entry i yields exactly what bem_bayes_sample gives for (seed, stream_id_i, sample) with the Philox source.
 * `entries` and `blocks` are DEVICE arrays. blocks[2b] = entry index, blocks[2b+1] = first 4-element Philox block of
 * CUDA block b, which covers 256 such blocks (1024 elements) of that entry.
 * sample = sample0 + (sample0_dev ? *sample0_dev : 0): the device word lets a captured CUDA graph be replayed for
 * another sample index. */
typedef struct BemBayesSampleEntry {
    const float* mu;
    const float* rho;
    float* w;
    int64_t numel;
    int64_t stream_id;
} BemBayesSampleEntry;
typedef struct BemBayesSampleBatchedParams {
    const BemBayesSampleEntry* entries;
    const int32_t* blocks;
    int32_t n_blocks;
    uint64_t seed;
    int64_t sample0;
    const int64_t* sample0_dev;
} BemBayesSampleBatchedParams;
int bem_bayes_sample_batched(const BemBayesSampleBatchedParams* p, void* stream);

typedef struct BemBayesPointwiseParams {
    int32_t n_samples;   /* S: weight sets; 1 = shared weights */
    int32_t batch;       /* total images = S * Bx */
    int32_t cin, cout;
    int64_t P;           /* pixels per image */
    const float* x;
    const float* w;      /* (S, cout, cin) or NULL when mu (+ sigma|rho, eps) are given */
    const float* mu;     /* (cout, cin) */
    const float* rho;    /* (cout, cin): sigma = log1p(exp(rho)) evaluated in the kernel (slow; prefer `sigma`) */
    const float* eps;    /* (S, cout, cin) */
    const float* bias;   /* (S, cout) or NULL */
    float* out;
    const float* sigma;     /* (cout, cin) precomputed log1p(exp(rho)), or NULL */
    const float* ln_gamma;  /* (cin) or NULL: LayerNorm over the input channels of every pixel is applied to x first */
    const float* ln_beta;   /* (cin) or NULL */
    float ln_eps;
    int32_t force_simt;     /* 1: fp32 CUDA-core tiles instead of the tcgen05 path (A/B measurements, no LayerNorm) */
    int64_t x_img_stride;   /* elements between consecutive images of x; 0 = cin * P (channel stride is always P) */
    int32_t sample_interleave; /* 0: image i uses weight set i / (batch / S); 1: i % S (grouped 1x1: x_proj / dt_proj,
                                  basicsr/vmamba/models/vmamba.py:659-661, with the K directions as weight sets) */
    void* workspace;        /* bem_bayes_pointwise_workspace_bytes() bytes, 16-byte aligned: packed weight tiles */
    int64_t workspace_bytes;
    const float* residual;  /* (batch, cout, P) or NULL: out = residual + conv(x) — the block's skip connection
                               (vmamba.py:1331-1333) folded into the epilogue; may alias `out` */
    const float* prelu_slope; /* NULL, or the negative slope(s) of the nn.PReLU that follows the conv (DualUpSample,
                                 basicsr/archs/UNet_arch.py:113-135): out = v > 0 ? v : slope * v, applied last */
    int32_t prelu_n;          /* number of slopes: 1 (nn.PReLU() default) or cout */
    int32_t prepacked;      /* 1: `workspace` still holds the packed tiles written by an earlier call with the same weights,
                               bias, LayerNorm parameters, n_samples / cin / cout and the same alignment class of x
                               (deterministic layers: pack once, reuse) — the pack kernel is skipped */
} BemBayesPointwiseParams;
int64_t bem_bayes_pointwise_workspace_bytes(int n_samples, int cin, int cout);
int bem_bayes_pointwise(const BemBayesPointwiseParams* p, void* stream);

/* The pack step of many bem_bayes_pointwise calls in ONE launch. A Monte-Carlo forward re-draws every Bayesian weight first
 * (bem_bayes_sample_batched; the reference samples inside each layer's forward, bayesian/conv.py:105-111), so the weight
 * tiles of all its 1x1 layers can be packed right after the draw; each layer's own call then passes `prepacked = 1` and the
 * same workspace, and launches one kernel instead of two.
 *   bem_bayes_pointwise_pack_table : host only. `params[i]` is the parameter block of the i-th later call (its x / out
 *       pointers are not dereferenced; only the alignment class of x, P and x_img_stride must be the later call's, and
 *       `workspace` must be that call's own buffer, not shared with another entry). Fills `table_host`
 *       (bem_bayes_pointwise_pack_table_bytes(n) bytes), which the caller copies to device memory once, and
 *       `*total_blocks`.
 *   bem_bayes_pointwise_pack_run   : one launch that packs every entry of the device copy of the table. */
int64_t bem_bayes_pointwise_pack_table_bytes(int n);
int bem_bayes_pointwise_pack_table(const BemBayesPointwiseParams* params, int n, void* table_host, int32_t* total_blocks);
int bem_bayes_pointwise_pack_run(const void* table_dev, int n, int total_blocks, void* stream);

typedef struct BemBayesDepthwiseParams {
    int32_t n_samples;
    int32_t batch;
    int32_t C, H, W, K;
    const float* x;
    const float* w;      /* (S, C, K, K) */
    const float* bias;   /* (S, C) or NULL */
    float* out;
    int32_t act;         /* what follows the convolution in the reference, fused:
                            0 none; 1 SiLU (SS2D.act, vmamba.py:708-710); 2 gated GELU: out has C/2 channels,
                            out[c] = gelu(y[c]) * y[c + C/2] (gdMlp: chunk -> act(x1) * x2, vmamba.py:129-131) */
} BemBayesDepthwiseParams;
int bem_bayes_depthwise(const BemBayesDepthwiseParams* p, void* stream);

/* Dense 3x3 convolution, stride 1, zero padding 1, groups 1, fp32 — the two full-resolution stems of the stage-1 network
 * (`first_conv` 3 -> 40 and `proj` 40 -> 3, basicsr/archs/UNet_arch.py:423-431; nn.Conv2d in the reference).
 *   x : (batch, cin, H, W)   w : (cout, cin, 3, 3)   bias : (cout) or NULL   out : (batch, cout, H, W)
 * Direct register-tiled kernel meant for small channel counts (cin * 9 * 8 floats of weights must fit 48 KB). */
typedef struct BemConv3x3Params {
    int32_t batch, cin, cout, H, W;
    const float* x;
    const float* w;
    const float* bias;
    float* out;
} BemConv3x3Params;
int bem_conv3x3(const BemConv3x3Params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Monte-Carlo best-sample selection.
 * Replaces `_idx = one_clip_list.index(max(one_clip_list))` (Enhancement/eval.py:270-271; NIQE uses min, :273-274):
 * first index attaining the extremum with Python's comparison semantics: every comparison with NaN is false, so a NaN at
 * index 0 stays selected (`max([nan, 1, 2])` is nan -> index 0) while a NaN anywhere else is never selected.
 *   scores : (n) fp32;  out_index : int32[1];  out_value : fp32[1] or NULL
 * ---------------------------------------------------------------------------------------------- */
int bem_select_best(const float* scores, int32_t n, int32_t take_min, int32_t* out_index, float* out_value, void* stream);

/* ------------------------------------------------------------------------------------------------
 * No-reference scorer of the Monte-Carlo predictions: the per-pixel part of NIQE, batched over images, fp64.
 * Replaces the inner loops of basicsr/metrics/niqe.py (called per prediction by Enhancement/eval.py:249-250):
 * bem_niqe_mscn        : mean / deviation with the 7x7 Gaussian window (scipy.ndimage.convolve, mode 'nearest') and the
 *                        normalised image (img - mu) / (sigma + 1)                                   (niqe.py:104-108)
 *     img : (n, H, W) fp32 (the rounded luma, or the half-scale image); window7x7 : 49 doubles; out : (n, H, W) fp64
 * bem_niqe_block_stats : per (block x block) tile of the normalised image, for the tile and its products with the four
 *                        circular shifts (0,1), (1,0), (1,1), (1,-1) of niqe.py:55-57, the moments the AGGD fit of
 *                        niqe.py:13-38 needs: sum_{v<0} v^2, #{v<0}, sum_{v>0} v^2, #{v>0}, sum |v|, sum v^2
 *     normalized : (n, H, W) fp64, H and W multiples of `block`; out : (n, (H/block)*(W/block), 5, 6) fp64, tiles row-major
 * The gamma-table matching, the 36-d Gaussian fit and the pseudo-inverse (niqe.py:126-139) are host-language code over
 * these few kilobytes (bem_b200/niqe.py).
 * ---------------------------------------------------------------------------------------------- */
int bem_niqe_mscn(const float* img, const double* window7x7, double* out, int32_t n_images, int32_t H, int32_t W, void* stream);
int bem_niqe_block_stats(const double* normalized, double* out, int32_t n_images, int32_t H, int32_t W, int32_t block, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over the channels of a channel-first tensor, forward and backward (training path).
 * Replaces LayerNorm2d.forward (basicsr/vmamba/models/vmamba.py:58-63: permute, F.layer_norm, permute) and its autograd.
 *     x, y, dy, dx : (batch, channels, hw) fp32 contiguous;  weight, bias : (channels) fp32 or NULL (no affine)
 *     mean, rstd   : (batch, hw) fp32, written by the forward for the backward; both NULL for an inference-only forward
 *     dweight, dbias : (channels) fp32, ACCUMULATED with atomics (zero-fill before the call); NULL to skip; dx NULL to skip
 * Statistics are the biased variance and rstd = rsqrt(var + eps) of torch.nn.functional.layer_norm.
 * ---------------------------------------------------------------------------------------------- */
int bem_layernorm2d_fwd(const float* x, const float* weight, const float* bias, float* y, float* mean, float* rstd,
                        int32_t batch, int32_t channels, int64_t hw, float eps, void* stream);
int bem_layernorm2d_bwd(const float* dy, const float* x, const float* weight, const float* mean, const float* rstd, float* dx,
                        float* dweight, float* dbias, int32_t batch, int32_t channels, int64_t hw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BEM_B200_H_ */
