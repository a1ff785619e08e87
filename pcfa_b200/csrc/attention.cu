// GMA global-motion attention, row softmax in half precision storage.
// Reference: models/gma/gma.py:54-76 — sim = einsum(q, k) (fp16 under the shipped autocast config), attn = softmax(sim)
// (autocast runs softmax in fp32 and returns fp32), then every iteration's aggregation einsum(attn, v)
// (gma.py:102-115) casts attn back to fp16 for the fp16 GEMM.  Materialised as the reference does, that is per
// sample: 99 MB (sim) -> 198 MB (fp32 copy) -> 198 MB (attn fp32) -> 6 x (198 MB read + 99 MB written) casts.
// Here: one pass, fp16 sim in, fp16 attn out, fp32 arithmetic inside — the values every aggregation GEMM of the
// reference consumes are exactly these (round-to-nearest fp16 of the fp32 softmax), so results are unchanged while
// the traffic drops from ~2.3 GB to 198 MB per sample; the backward is the matching single pass
//   dsim = attn * (dattn - sum_j attn_j * dattn_j).
// One CTA per row (7040 columns at 436x1024): the row lives in registers (8 halves per 128-bit load), two block
// reductions (max, sum).  HBM-bound: 2 bytes in + 2 bytes out per element forward, 4 + 2 backward.
#include "common.cuh"
#include <cuda_fp16.h>

namespace pcfa {

constexpr int SM_THREADS = 256;
constexpr int SM_MAX_VEC = 8;                 // up to 8 x 8 halves per thread = 16384 columns per row (template NV <= 8)

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, u) : v + u;
    }
    __syncthreads();                           // red[] may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < SM_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

struct alignas(16) Half8 { __half2 h[4]; };

// rows x cols, cols % 8 == 0, cols <= SM_THREADS * 8 * NV
template <int NV>
__global__ void __launch_bounds__(SM_THREADS)
softmax_rows_f16_kernel(const __half* __restrict__ sim, __half* __restrict__ attn, int cols) {
    __shared__ float red[SM_THREADS / 32];
    const int64_t row = blockIdx.x;
    const Half8* src = reinterpret_cast<const Half8*>(sim + row * cols);
    Half8* dst = reinterpret_cast<Half8*>(attn + row * cols);
    const int nvec = cols >> 3;
    float x[NV][8];
    float mx = -3.0e38f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = threadIdx.x + i * SM_THREADS;
        if (v < nvec) {
            const Half8 h = src[v];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __half22float2(h.h[k]);
                x[i][2 * k] = f.x; x[i][2 * k + 1] = f.y;
                mx = fmaxf(mx, fmaxf(f.x, f.y));
            }
        }
    }
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (threadIdx.x + i * SM_THREADS < nvec) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { x[i][k] = __expf(x[i][k] - mx); sum += x[i][k]; }
        }
    }
    sum = block_reduce(sum, red, false);
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = threadIdx.x + i * SM_THREADS;
        if (v < nvec) {
            Half8 h;
#pragma unroll
            for (int k = 0; k < 4; ++k) h.h[k] = __floats2half2_rn(x[i][2 * k] * inv, x[i][2 * k + 1] * inv);
            dst[v] = h;
        }
    }
}

template <int NV>
__global__ void __launch_bounds__(SM_THREADS)
softmax_rows_f16_bwd_kernel(const __half* __restrict__ attn, const __half* __restrict__ dattn, __half* __restrict__ dsim, int cols) {
    __shared__ float red[SM_THREADS / 32];
    const int64_t row = blockIdx.x;
    const Half8* pa = reinterpret_cast<const Half8*>(attn + row * cols);
    const Half8* pg = reinterpret_cast<const Half8*>(dattn + row * cols);
    Half8* dst = reinterpret_cast<Half8*>(dsim + row * cols);
    const int nvec = cols >> 3;
    float a[NV][8], g[NV][8];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = threadIdx.x + i * SM_THREADS;
        if (v < nvec) {
            const Half8 ha = pa[v], hg = pg[v];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 fa = __half22float2(ha.h[k]), fg = __half22float2(hg.h[k]);
                a[i][2 * k] = fa.x; a[i][2 * k + 1] = fa.y; g[i][2 * k] = fg.x; g[i][2 * k + 1] = fg.y;
                dot = fmaf(fa.x, fg.x, fmaf(fa.y, fg.y, dot));
            }
        }
    }
    dot = block_reduce(dot, red, false);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int v = threadIdx.x + i * SM_THREADS;
        if (v < nvec) {
            Half8 h;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                h.h[k] = __floats2half2_rn(a[i][2 * k] * (g[i][2 * k] - dot), a[i][2 * k + 1] * (g[i][2 * k + 1] - dot));
            dst[v] = h;
        }
    }
}

}  // namespace pcfa

using namespace pcfa;

static int sm_check(const void* a, const void* b, int64_t rows, int cols) {
    if (!a || !b || rows <= 0 || cols <= 0) return PCFA_E_BADARG;
    if (cols % 8 != 0 || cols > SM_THREADS * 8 * SM_MAX_VEC) return PCFA_E_BADARG;     // the caller pads / falls back
    if (rows > 0x7fffffffLL) return PCFA_E_TOOLARGE;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) return PCFA_E_BADARG;
    return PCFA_OK;
}

extern "C" int pcfa_softmax_rows_f16_forward(const void* sim, void* attn, int64_t rows, int cols, pcfa_stream_t stream) {
    PCFA_TRY(sm_check(sim, attn, rows, cols));
    const __half* a = reinterpret_cast<const __half*>(sim);
    __half* o = reinterpret_cast<__half*>(attn);
    const int nv = ceil_div(cols, SM_THREADS * 8);
    if (nv <= 2)      softmax_rows_f16_kernel<2><<<(unsigned)rows, SM_THREADS, 0, as_stream(stream)>>>(a, o, cols);
    else if (nv <= 4) softmax_rows_f16_kernel<4><<<(unsigned)rows, SM_THREADS, 0, as_stream(stream)>>>(a, o, cols);
    else              softmax_rows_f16_kernel<8><<<(unsigned)rows, SM_THREADS, 0, as_stream(stream)>>>(a, o, cols);
    return after_launch();
}

extern "C" int pcfa_softmax_rows_f16_backward(const void* attn, const void* grad_attn, void* grad_sim, int64_t rows, int cols,
                                              pcfa_stream_t stream) {
    PCFA_TRY(sm_check(attn, grad_attn, rows, cols));
    if (!grad_sim || (reinterpret_cast<uintptr_t>(grad_sim) & 15)) return PCFA_E_BADARG;
    const __half* a = reinterpret_cast<const __half*>(attn);
    const __half* g = reinterpret_cast<const __half*>(grad_attn);
    __half* o = reinterpret_cast<__half*>(grad_sim);
    const int nv = ceil_div(cols, SM_THREADS * 8);
    if (nv <= 2)      softmax_rows_f16_bwd_kernel<2><<<(unsigned)rows, SM_THREADS, 0, as_stream(stream)>>>(a, g, o, cols);
    else if (nv <= 4) softmax_rows_f16_bwd_kernel<4><<<(unsigned)rows, SM_THREADS, 0, as_stream(stream)>>>(a, g, o, cols);
    else              softmax_rows_f16_bwd_kernel<8><<<(unsigned)rows, SM_THREADS, 0, as_stream(stream)>>>(a, g, o, cols);
    return after_launch();
}
