// select.cu — Monte-Carlo best-sample selection.
// Replaces `_idx = lst.index(max(lst))` / `lst.index(min(lst))` (Enhancement/eval.py:270-274) with the exact semantics
// Python gives it: the FIRST index attaining the extremum; a NaN is only ever returned when it sits at index 0
// (max()/min() keep their first element unless a later one compares strictly better, and NaN never does).
#include "bem_kernels.h"

namespace bem {

__device__ __forceinline__ bool sel_better(float v, int i, float bv, int bi, bool take_min) {
    // strictly better value, or equal value at a lower index (NaN compares false both ways)
    const bool strictly = take_min ? (v < bv) : (v > bv);
    return strictly || (v == bv && i < bi);
}

__global__ void __launch_bounds__(256) select_best_kernel(const float* __restrict__ scores, int n, int take_min_,
                                                          int* __restrict__ out_index, float* __restrict__ out_value) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sv[256];
    __shared__ int si[256];
    const bool take_min = take_min_ != 0;
    const int tid = threadIdx.x;
    const float first = scores[0];
    if (first != first) {   // NaN at index 0 wins, exactly as in Python
        if (tid == 0) {
            *out_index = 0;
            if (out_value) *out_value = first;
        }
        return;
    }
    float bv = first;
    int bi = 0;
    for (int i = tid; i < n; i += 256) {
        const float v = scores[i];
        if (sel_better(v, i, bv, bi, take_min)) {
            bv = v;
            bi = i;
        }
    }
    sv[tid] = bv;
    si[tid] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
            if (sel_better(sv[tid + o], si[tid + o], sv[tid], si[tid], take_min)) {
                sv[tid] = sv[tid + o];
                si[tid] = si[tid + o];
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        *out_index = si[0];
        if (out_value) *out_value = sv[0];
    }
}

}  // namespace bem

extern "C" int bem_select_best(const float* scores, int32_t n, int32_t take_min, int32_t* out_index, float* out_value,
                               void* stream) {
    if (!scores || !out_index || n <= 0) return BEM_ERR_BAD_ARG;
    bem::launch_pdl(bem::select_best_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, scores, n, take_min, out_index, out_value);
    return (int)cudaGetLastError();
}
