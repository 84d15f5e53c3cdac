// csm.cu — CrossScan / CrossMerge (four-direction traversal) for sm_100a.
//
// Replaces triton_cross_scan_flex (basicsr/vmamba/models/csm_triton.py:278-390) and its torch twins (:22-179).
//   * channel-first <-> channel-first, scans == 0 (the only combination the BEM archs use, vmamba.py:657,684):
//     32x32 shared-memory tile transposes, every global access coalesced in both the row-major and the column-major
//     direction (the Triton kernel writes the transposed directions with stride-H scatter, :330-345)
//   * every other layout / scan mode: an index-mapped kernel whose threads follow the DESTINATION layout
// Pure data movement; the merge adds in the reference's association (y0 + y2) + (y1 + y3) (:60-62) and in the tensor's
// own precision, so results are bit-exact against the torch path.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "bem_kernels.h"

namespace bem {

template <typename T> __device__ __forceinline__ float csm_to_f(T v);
template <> __device__ __forceinline__ float csm_to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float csm_to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float csm_to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T csm_from_f(float v);
template <> __device__ __forceinline__ float csm_from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half csm_from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 csm_from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
// a + b rounded to T (what a T-typed torch add produces)
template <typename T> __device__ __forceinline__ T csm_add(T a, T b) { return csm_from_f<T>(csm_to_f<T>(a) + csm_to_f<T>(b)); }

struct CsmArgs {
    int B, C, H, W;
    int img_cf, seq_cf, obo, scans;
    const void* src;
    void* dst;
};

// position of pixel (h, w) in direction k's sequence
__device__ __forceinline__ int64_t csm_pos(int scans, int k, int h, int w, int H, int W) {
    const int64_t L = (int64_t)H * W;
    int64_t pos;
    if (scans == 0) {
        pos = (k & 1) ? (int64_t)w * H + h : (int64_t)h * W + w;
        if (k >= 2) pos = L - 1 - pos;
    } else if (scans == 1) {
        pos = (int64_t)h * W + w;
    } else {
        pos = (int64_t)h * W + w;
        if (k >= 2) pos = L - 1 - pos;
    }
    return pos;
}
// inverse: sequence position -> pixel
__device__ __forceinline__ void csm_inv(int scans, int k, int64_t l, int H, int W, int& h, int& w) {
    const int64_t L = (int64_t)H * W;
    const bool rev = (scans == 0 || scans == 2) && k >= 2;
    const int64_t q = rev ? L - 1 - l : l;
    if (scans == 0 && (k & 1)) {
        w = (int)(q / H);
        h = (int)(q - (int64_t)w * H);
    } else {
        h = (int)(q / W);
        w = (int)(q - (int64_t)h * W);
    }
}
__device__ __forceinline__ int64_t seq_off(const CsmArgs& p, int b, int k, int c, int64_t l) {
    const int64_t L = (int64_t)p.H * p.W;
    return p.seq_cf ? (((int64_t)b * 4 + k) * p.C + c) * L + l : (((int64_t)b * L + l) * 4 + k) * p.C + c;
}
__device__ __forceinline__ int64_t img_off(const CsmArgs& p, int b, int k, int c, int h, int w) {
    if (p.obo)
        return p.img_cf ? ((((int64_t)b * 4 + k) * p.C + c) * p.H + h) * p.W + w
                        : ((((int64_t)b * p.H + h) * p.W + w) * 4 + k) * p.C + c;
    return p.img_cf ? (((int64_t)b * p.C + c) * p.H + h) * p.W + w : (((int64_t)b * p.H + h) * p.W + w) * p.C + c;
}

// ---------------------------------------------------------------------------------------------------
// generic kernels: one thread per destination element, thread order = destination memory order
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void csm_scan_generic(const CsmArgs p) {
    pdl_trigger();
    pdl_wait();
    const int64_t L = (int64_t)p.H * p.W;
    const int64_t total = (int64_t)p.B * 4 * p.C * L;
    const T* src = reinterpret_cast<const T*>(p.src);
    T* dst = reinterpret_cast<T*>(p.dst);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int b, k, c;
        int64_t l;
        if (p.seq_cf) {   // (B,4,C,L)
            l = i % L;
            int64_t r = i / L;
            c = (int)(r % p.C);
            r /= p.C;
            k = (int)(r % 4);
            b = (int)(r / 4);
        } else {          // (B,L,4,C)
            c = (int)(i % p.C);
            int64_t r = i / p.C;
            k = (int)(r % 4);
            r /= 4;
            l = r % L;
            b = (int)(r / L);
        }
        int h, w;
        csm_inv(p.scans, k, l, p.H, p.W, h, w);
        dst[i] = src[img_off(p, b, k, c, h, w)];
    }
}

template <typename T>
__global__ void csm_merge_generic(const CsmArgs p) {
    pdl_trigger();
    pdl_wait();
    const int64_t L = (int64_t)p.H * p.W;
    const int K = p.obo ? 4 : 1;
    const int64_t total = (int64_t)p.B * K * p.C * L;
    const T* src = reinterpret_cast<const T*>(p.src);
    T* dst = reinterpret_cast<T*>(p.dst);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int b, k, c, h, w;
        if (p.img_cf) {   // (B,[4],C,H,W)
            w = (int)(i % p.W);
            int64_t r = i / p.W;
            h = (int)(r % p.H);
            r /= p.H;
            c = (int)(r % p.C);
            r /= p.C;
            k = (int)(r % K);
            b = (int)(r / K);
        } else {          // (B,H,W,[4],C)
            c = (int)(i % p.C);
            int64_t r = i / p.C;
            k = (int)(r % K);
            r /= K;
            w = (int)(r % p.W);
            r /= p.W;
            h = (int)(r % p.H);
            b = (int)(r / p.H);
        }
        if (p.obo) {
            dst[i] = src[seq_off(p, b, k, c, csm_pos(p.scans, k, h, w, p.H, p.W))];
        } else {
            const T v0 = src[seq_off(p, b, 0, c, csm_pos(p.scans, 0, h, w, p.H, p.W))];
            const T v1 = src[seq_off(p, b, 1, c, csm_pos(p.scans, 1, h, w, p.H, p.W))];
            const T v2 = src[seq_off(p, b, 2, c, csm_pos(p.scans, 2, h, w, p.H, p.W))];
            const T v3 = src[seq_off(p, b, 3, c, csm_pos(p.scans, 3, h, w, p.H, p.W))];
            T r;
            if (p.scans == 1) {   // y.sum(1) (csm_triton.py:64): torch accumulates in fp32 and rounds once
                r = csm_from_f<T>(((csm_to_f<T>(v0) + csm_to_f<T>(v1)) + csm_to_f<T>(v2)) + csm_to_f<T>(v3));
            } else {              // (y0 + y2) + (y1 + y3)  (csm_triton.py:60-62, 66-67)
                r = csm_add<T>(csm_add<T>(v0, v2), csm_add<T>(v1, v3));
            }
            dst[i] = r;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// tiled kernels: channel-first both sides, scans == 0. grid = (tilesW * tilesH, C, B), block = (32, 8)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) csm_scan_tiled(const CsmArgs p) {
    pdl_trigger();
    pdl_wait();
    __shared__ T tile[4][32][33];   // one_by_one: a tile per direction, all loads in flight before the one barrier
    const int H = p.H, W = p.W;
    const int64_t L = (int64_t)H * W;
    const int tilesW = (W + 31) / 32;
    const int h0 = (blockIdx.x / tilesW) * 32, w0 = (blockIdx.x % tilesW) * 32;
    const int c = blockIdx.y, b = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const T* src = reinterpret_cast<const T*>(p.src);
    T* dst = reinterpret_cast<T*>(p.dst);
    const int nsrc = p.obo ? 4 : 1;
    T r[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < nsrc) {
            const T* s = src + (p.obo ? (((int64_t)b * 4 + k) * p.C + c) : ((int64_t)b * p.C + c)) * L;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int h = h0 + ty + 8 * i, w = w0 + tx;
                r[k][i] = (h < H && w < W) ? s[(int64_t)h * W + w] : csm_from_f<T>(0.f);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < nsrc) {
#pragma unroll
            for (int i = 0; i < 4; ++i) tile[k][ty + 8 * i][tx] = r[k][i];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ks = p.obo ? k : 0;
        T* d = dst + (((int64_t)b * 4 + k) * p.C + c) * L;
        if ((k & 1) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int h = h0 + ty + 8 * i, w = w0 + tx;
                if (h < H && w < W) {
                    int64_t pos = (int64_t)h * W + w;
                    if (k == 2) pos = L - 1 - pos;
                    d[pos] = tile[ks][ty + 8 * i][tx];
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int w = w0 + ty + 8 * i, h = h0 + tx;
                if (h < H && w < W) {
                    int64_t pos = (int64_t)w * H + h;
                    if (k == 3) pos = L - 1 - pos;
                    d[pos] = tile[ks][tx][ty + 8 * i];
                }
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) csm_merge_tiled(const CsmArgs p) {
    pdl_trigger();
    pdl_wait();
    __shared__ T tile[2][32][33];   // the two column-major directions are transposed through shared memory
    const int H = p.H, W = p.W;
    const int64_t L = (int64_t)H * W;
    const int tilesW = (W + 31) / 32;
    const int h0 = (blockIdx.x / tilesW) * 32, w0 = (blockIdx.x % tilesW) * 32;
    const int c = blockIdx.y, b = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const T* src = reinterpret_cast<const T*>(p.src);
    T* dst = reinterpret_cast<T*>(p.dst);
    T v[4][4];   // [k][i]; all 16 loads of a thread are in flight before the one barrier
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const T* s = src + (((int64_t)b * 4 + k) * p.C + c) * L;
        if ((k & 1) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int h = h0 + ty + 8 * i, w = w0 + tx;
                int64_t pos = (int64_t)h * W + w;
                if (k == 2) pos = L - 1 - pos;
                v[k][i] = (h < H && w < W) ? s[pos] : csm_from_f<T>(0.f);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int w = w0 + ty + 8 * i, h = h0 + tx;
                int64_t pos = (int64_t)w * H + h;
                if (k == 3) pos = L - 1 - pos;
                v[k][i] = (h < H && w < W) ? s[pos] : csm_from_f<T>(0.f);   // element (w, h) of the transposed tile
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        tile[0][ty + 8 * i][tx] = v[1][i];   // tile[w][h]
        tile[1][ty + 8 * i][tx] = v[3][i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[1][i] = tile[0][tx][ty + 8 * i];   // (h = ty + 8i, w = tx)
        v[3][i] = tile[1][tx][ty + 8 * i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int h = h0 + ty + 8 * i, w = w0 + tx;
        if (h < H && w < W) {
            if (p.obo) {
#pragma unroll
                for (int k = 0; k < 4; ++k) dst[(((int64_t)b * 4 + k) * p.C + c) * L + (int64_t)h * W + w] = v[k][i];
            } else {
                dst[((int64_t)b * p.C + c) * L + (int64_t)h * W + w] =
                    csm_add<T>(csm_add<T>(v[0][i], v[2][i]), csm_add<T>(v[1][i], v[3][i]));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// fp32, H % 4 == 0 and W % 4 == 0 (levels 0 and 1 of the BEM nets): the same tile transposes with 16-byte global accesses.
// Tile = 64 x 64; a thread owns four float4: along w for the row-major directions, along h for the column-major ones
// (position w * H + h is contiguous in h); a reversed direction stores its float4 mirrored at L - 4 - pos.
// grid = (tilesW * tilesH, C, B), block = 256
// ---------------------------------------------------------------------------------------------------
constexpr int CSM_T = 64, CSM_LD = 65;
__device__ __forceinline__ float4 csm_rev4(float4 v) { return make_float4(v.w, v.z, v.y, v.x); }

__global__ void __launch_bounds__(256) csm_scan_tiled4(const CsmArgs p) {
    pdl_trigger();
    pdl_wait();
    __shared__ float tile[2][CSM_T][CSM_LD];
    const int H = p.H, W = p.W;
    const int64_t L = (int64_t)H * W;
    const int tilesW = (W + CSM_T - 1) / CSM_T;
    const int h0 = (blockIdx.x / tilesW) * CSM_T, w0 = (blockIdx.x % tilesW) * CSM_T;
    const int c = blockIdx.y, b = blockIdx.z;
    const int q = threadIdx.x & 15, r = threadIdx.x >> 4;   // float4 column / row within a 16-row band
    const float* src = reinterpret_cast<const float*>(p.src);
    float* dst = reinterpret_cast<float*>(p.dst);
    const int nsrc = p.obo ? 4 : 1;
    float4 v[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < nsrc) {
            const float* s = src + (p.obo ? (((int64_t)b * 4 + k) * p.C + c) : ((int64_t)b * p.C + c)) * L;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int h = h0 + r + 16 * i, w = w0 + 4 * q;
                v[k][i] = (h < H && w < W) ? *reinterpret_cast<const float4*>(s + (int64_t)h * W + w) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    // the column-major directions go through shared memory: direction 1 from tile 0, direction 3 from tile 1 (one_by_one) or 0
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        if (t == 0 || p.obo) {
            const int k = p.obo ? 1 + 2 * t : 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float* row = &tile[t][r + 16 * i][4 * q];
                row[0] = v[k][i].x; row[1] = v[k][i].y; row[2] = v[k][i].z; row[3] = v[k][i].w;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k += 2) {   // row-major directions straight from registers
        const int ks = p.obo ? k : 0;
        float* d = dst + (((int64_t)b * 4 + k) * p.C + c) * L;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int h = h0 + r + 16 * i, w = w0 + 4 * q;
            if (h < H && w < W) {
                const int64_t pos = (int64_t)h * W + w;
                if (k == 0) *reinterpret_cast<float4*>(d + pos) = v[ks][i];
                else *reinterpret_cast<float4*>(d + (L - 4 - pos)) = csm_rev4(v[ks][i]);
            }
        }
    }
#pragma unroll
    for (int k = 1; k < 4; k += 2) {
        const int t = (p.obo && k == 3) ? 1 : 0;
        float* d = dst + (((int64_t)b * 4 + k) * p.C + c) * L;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int wl = r + 16 * i, hl = 4 * q;
            const int w = w0 + wl, h = h0 + hl;
            if (h < H && w < W) {
                const float4 o = make_float4(tile[t][hl][wl], tile[t][hl + 1][wl], tile[t][hl + 2][wl], tile[t][hl + 3][wl]);
                const int64_t pos = (int64_t)w * H + h;
                if (k == 1) *reinterpret_cast<float4*>(d + pos) = o;
                else *reinterpret_cast<float4*>(d + (L - 4 - pos)) = csm_rev4(o);
            }
        }
    }
}

__global__ void __launch_bounds__(256) csm_merge_tiled4(const CsmArgs p) {
    pdl_trigger();
    pdl_wait();
    __shared__ float tile[2][CSM_T][CSM_LD];
    const int H = p.H, W = p.W;
    const int64_t L = (int64_t)H * W;
    const int tilesW = (W + CSM_T - 1) / CSM_T;
    const int h0 = (blockIdx.x / tilesW) * CSM_T, w0 = (blockIdx.x % tilesW) * CSM_T;
    const int c = blockIdx.y, b = blockIdx.z;
    const int q = threadIdx.x & 15, r = threadIdx.x >> 4;
    const float* src = reinterpret_cast<const float*>(p.src);
    float* dst = reinterpret_cast<float*>(p.dst);
    float4 v[4][4];   // all 16 loads of a thread in flight before the one barrier
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float* s = src + (((int64_t)b * 4 + k) * p.C + c) * L;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if ((k & 1) == 0) {
                const int h = h0 + r + 16 * i, w = w0 + 4 * q;
                const int64_t pos = (int64_t)h * W + w;
                if (h < H && w < W) v[k][i] = k == 0 ? *reinterpret_cast<const float4*>(s + pos) : csm_rev4(*reinterpret_cast<const float4*>(s + (L - 4 - pos)));
                else v[k][i] = zero;
            } else {
                const int w = w0 + r + 16 * i, h = h0 + 4 * q;
                const int64_t pos = (int64_t)w * H + h;
                if (h < H && w < W) v[k][i] = k == 1 ? *reinterpret_cast<const float4*>(s + pos) : csm_rev4(*reinterpret_cast<const float4*>(s + (L - 4 - pos)));
                else v[k][i] = zero;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // element j of the float4 is pixel (h = 4q + j, w = r + 16 i)
            const int wl = r + 16 * i, hl = 4 * q;
            const float4 o = v[1 + 2 * t][i];
            tile[t][hl][wl] = o.x; tile[t][hl + 1][wl] = o.y; tile[t][hl + 2][wl] = o.z; tile[t][hl + 3][wl] = o.w;
        }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int hl = r + 16 * i, wl = 4 * q;
        const int h = h0 + hl, w = w0 + wl;
        if (h < H && w < W) {
            const float4 a1 = make_float4(tile[0][hl][wl], tile[0][hl][wl + 1], tile[0][hl][wl + 2], tile[0][hl][wl + 3]);
            const float4 a3 = make_float4(tile[1][hl][wl], tile[1][hl][wl + 1], tile[1][hl][wl + 2], tile[1][hl][wl + 3]);
            const int64_t pos = (int64_t)h * W + w;
            if (p.obo) {
                *reinterpret_cast<float4*>(dst + (((int64_t)b * 4 + 0) * p.C + c) * L + pos) = v[0][i];
                *reinterpret_cast<float4*>(dst + (((int64_t)b * 4 + 1) * p.C + c) * L + pos) = a1;
                *reinterpret_cast<float4*>(dst + (((int64_t)b * 4 + 2) * p.C + c) * L + pos) = v[2][i];
                *reinterpret_cast<float4*>(dst + (((int64_t)b * 4 + 3) * p.C + c) * L + pos) = a3;
            } else {   // (y0 + y2) + (y1 + y3), csm_triton.py:60-62
                const float4 a0 = v[0][i], a2 = v[2][i];
                *reinterpret_cast<float4*>(dst + ((int64_t)b * p.C + c) * L + pos) =
                    make_float4((a0.x + a2.x) + (a1.x + a3.x), (a0.y + a2.y) + (a1.y + a3.y), (a0.z + a2.z) + (a1.z + a3.z),
                                (a0.w + a2.w) + (a1.w + a3.w));
            }
        }
    }
}

template <typename T>
static int csm_launch(const CsmArgs& a, bool merge, cudaStream_t stream) {
    const bool tiled = a.img_cf && a.seq_cf && a.scans == 0 && a.C <= 65535 && a.B <= 65535;
    const bool vec4 = tiled && sizeof(T) == 4 && a.H % 4 == 0 && a.W % 4 == 0 &&
                      ((reinterpret_cast<uintptr_t>(a.src) | reinterpret_cast<uintptr_t>(a.dst)) & 15) == 0;
    if (vec4) {
        dim3 grid(((a.W + CSM_T - 1) / CSM_T) * ((a.H + CSM_T - 1) / CSM_T), a.C, a.B);
        if (merge) launch_pdl(csm_merge_tiled4, grid, dim3(256), 0, stream, a);
        else launch_pdl(csm_scan_tiled4, grid, dim3(256), 0, stream, a);
    } else if (tiled) {
        dim3 grid(((a.W + 31) / 32) * ((a.H + 31) / 32), a.C, a.B), block(32, 8);
        if (merge) launch_pdl(csm_merge_tiled<T>, dim3(grid), dim3(block), 0, stream, a);
        else launch_pdl(csm_scan_tiled<T>, dim3(grid), dim3(block), 0, stream, a);
    } else {
        const int64_t total = (int64_t)a.B * ((merge && !a.obo) ? 1 : 4) * a.C * a.H * a.W;
        const int threads = 256;
        int64_t blocks = (total + threads - 1) / threads;
        const int64_t cap = (int64_t)device_sm_count() * 32;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        if (merge) launch_pdl(csm_merge_generic<T>, dim3((int)blocks), dim3(threads), 0, stream, a);
        else launch_pdl(csm_scan_generic<T>, dim3((int)blocks), dim3(threads), 0, stream, a);
    }
    return (int)cudaGetLastError();
}

static int csm_entry(const BemCsmParams* q, bool merge, void* stream_) {
    if (!q || !q->src || !q->dst) return BEM_ERR_BAD_ARG;
    if (q->B <= 0 || q->C <= 0 || q->H <= 0 || q->W <= 0) return BEM_ERR_BAD_ARG;
    if (q->scans < 0 || q->scans > 2) return BEM_ERR_BAD_ARG;
    CsmArgs a{q->B, q->C, q->H, q->W, q->img_channel_first ? 1 : 0, q->seq_channel_first ? 1 : 0, q->one_by_one ? 1 : 0,
              q->scans, q->src, q->dst};
    cudaStream_t stream = (cudaStream_t)stream_;
    switch (q->dtype) {
        case BEM_F32: return csm_launch<float>(a, merge, stream);
        case BEM_F16: return csm_launch<__half>(a, merge, stream);
        case BEM_BF16: return csm_launch<__nv_bfloat16>(a, merge, stream);
        default: return BEM_ERR_BAD_ARG;
    }
}

}  // namespace bem

extern "C" {
int bem_cross_scan(const BemCsmParams* p, void* stream) { return bem::csm_entry(p, false, stream); }
int bem_cross_merge(const BemCsmParams* p, void* stream) { return bem::csm_entry(p, true, stream); }
}
