// Bias + ReLU epilogue of the frozen convolutions (glue around the cuDNN convolutions).
// torch's cuDNN convolution adds the bias in a separate broadcasting ATen kernel (non-vectorised: 8.5 us for a 3.6 MB
// activation, 122 launches = 1.0 ms of one RAFT closure) and the ReLU in another (clamp, 99 launches, 0.47 ms).  Here the
// convolution runs without bias and ONE in-place pass does y = act(x + b[c]), act = ReLU or LeakyReLU(slope) (PWCNet and
// FlowNet2 use slope 0.1); the backward is the mask y > 0 ? g : slope * g in one pass (what threshold_backward does).  fp32 and fp16 (GMA under autocast), channels-last or NCHW.
// HBM/L2-bound: 8 bytes per element forward (in place), 12 backward.
#include "common.cuh"
#include <cuda_fp16.h>

namespace pcfa {

constexpr int BA_THREADS = 256;

// channel of flat element i:  channels-last: i % C   |   NCHW: (i / inner) % C     (inner = H*W, 1 for channels-last)
template <bool RELU>
__global__ void __launch_bounds__(BA_THREADS)
bias_act_f32_kernel(float* __restrict__ x, const float* __restrict__ bias, int64_t n4, int C, int64_t inner, float slope) {
    for (int64_t v = (int64_t)blockIdx.x * BA_THREADS + threadIdx.x; v < n4; v += (int64_t)gridDim.x * BA_THREADS) {
        float4 a = reinterpret_cast<float4*>(x)[v];
        const int64_t i = v * 4;
        if (inner == 1) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + (int)(i % C)));
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        } else {
            const float b = bias ? __ldg(bias + (int)((i / inner) % C)) : 0.f;     // inner % 4 == 0: one channel per vector
            a.x += b; a.y += b; a.z += b; a.w += b;
        }
        if (RELU) { a.x = fmaxf(a.x, slope * a.x); a.y = fmaxf(a.y, slope * a.y); a.z = fmaxf(a.z, slope * a.z); a.w = fmaxf(a.w, slope * a.w); }   // 0 <= slope < 1
        reinterpret_cast<float4*>(x)[v] = a;
    }
}

template <bool RELU>
__global__ void __launch_bounds__(BA_THREADS)
bias_act_f16_kernel(__half* __restrict__ x, const __half* __restrict__ bias, int64_t n8, int C, int64_t inner, float slope) {
    for (int64_t v = (int64_t)blockIdx.x * BA_THREADS + threadIdx.x; v < n8; v += (int64_t)gridDim.x * BA_THREADS) {
        uint4 raw = reinterpret_cast<uint4*>(x)[v];
        __half2* h = reinterpret_cast<__half2*>(&raw);
        const int64_t i = v * 8;
        const __half2 sl = __float2half2_rn(slope);
        if (inner == 1) {
            const uint4 braw = __ldg(reinterpret_cast<const uint4*>(bias + (int)(i % C)));
            const __half2* b = reinterpret_cast<const __half2*>(&braw);
#pragma unroll
            for (int k = 0; k < 4; ++k) { h[k] = __hadd2(h[k], b[k]); if (RELU) h[k] = __hmax2(h[k], __hmul2(h[k], sl)); }
        } else {
            const __half2 b = __half2half2(__ldg(bias + (int)((i / inner) % C)));
#pragma unroll
            for (int k = 0; k < 4; ++k) { h[k] = __hadd2(h[k], b); if (RELU) h[k] = __hmax2(h[k], __hmul2(h[k], sl)); }
        }
        reinterpret_cast<uint4*>(x)[v] = raw;
    }
}

__global__ void __launch_bounds__(BA_THREADS)
relu_mask_f32_kernel(const float* __restrict__ y, const float* __restrict__ gy, float* __restrict__ gx, int64_t n4, float slope) {
    for (int64_t v = (int64_t)blockIdx.x * BA_THREADS + threadIdx.x; v < n4; v += (int64_t)gridDim.x * BA_THREADS) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(y) + v), g = __ldg(reinterpret_cast<const float4*>(gy) + v);
        reinterpret_cast<float4*>(gx)[v] = make_float4(a.x > 0.f ? g.x : slope * g.x, a.y > 0.f ? g.y : slope * g.y, a.z > 0.f ? g.z : slope * g.z, a.w > 0.f ? g.w : slope * g.w);
    }
}

__global__ void __launch_bounds__(BA_THREADS)
relu_mask_f16_kernel(const __half* __restrict__ y, const __half* __restrict__ gy, __half* __restrict__ gx, int64_t n8, float slope) {
    for (int64_t v = (int64_t)blockIdx.x * BA_THREADS + threadIdx.x; v < n8; v += (int64_t)gridDim.x * BA_THREADS) {
        const uint4 ar = __ldg(reinterpret_cast<const uint4*>(y) + v);
        uint4 gr = __ldg(reinterpret_cast<const uint4*>(gy) + v);
        const __half* a = reinterpret_cast<const __half*>(&ar);
        __half* g = reinterpret_cast<__half*>(&gr);
#pragma unroll
        for (int k = 0; k < 8; ++k) if (!(__half2float(a[k]) > 0.f)) g[k] = __float2half_rn(slope * __half2float(g[k]));
        reinterpret_cast<uint4*>(gx)[v] = gr;
    }
}

// grad_x = y > 0 ? g1 + g2 : 0 — ReLU mask applied to the SUM of two gradients (a residual tail whose output feeds the next
// block's convolution and its skip branch): one pass instead of autograd's add followed by the mask
__global__ void __launch_bounds__(BA_THREADS)
relu_mask2_f32_kernel(const float* __restrict__ y, const float* __restrict__ g1, const float* __restrict__ g2, float* __restrict__ gx, int64_t n4) {
    for (int64_t v = (int64_t)blockIdx.x * BA_THREADS + threadIdx.x; v < n4; v += (int64_t)gridDim.x * BA_THREADS) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(y) + v), p = __ldg(reinterpret_cast<const float4*>(g1) + v),
                     q = __ldg(reinterpret_cast<const float4*>(g2) + v);
        reinterpret_cast<float4*>(gx)[v] = make_float4(a.x > 0.f ? p.x + q.x : 0.f, a.y > 0.f ? p.y + q.y : 0.f,
                                                       a.z > 0.f ? p.z + q.z : 0.f, a.w > 0.f ? p.w + q.w : 0.f);
    }
}

// grad_y is a channel slice of a wider channels-last tensor (the gradient of a concatenation): rows of C elements, ld apart
__global__ void __launch_bounds__(BA_THREADS)
relu_mask_rows_f32_kernel(const float* __restrict__ y, const float* __restrict__ gy, float* __restrict__ gx, int64_t n4, int c4, int64_t ld,
                          float slope) {
    for (int64_t v = (int64_t)blockIdx.x * BA_THREADS + threadIdx.x; v < n4; v += (int64_t)gridDim.x * BA_THREADS) {
        const int64_t r = v / c4;
        const int c = (int)(v - r * c4);
        const float4 a = __ldg(reinterpret_cast<const float4*>(y) + v), g = __ldg(reinterpret_cast<const float4*>(gy + r * ld) + c);
        reinterpret_cast<float4*>(gx)[v] = make_float4(a.x > 0.f ? g.x : slope * g.x, a.y > 0.f ? g.y : slope * g.y, a.z > 0.f ? g.z : slope * g.z, a.w > 0.f ? g.w : slope * g.w);
    }
}

__global__ void __launch_bounds__(BA_THREADS)
relu_mask_rows_f16_kernel(const __half* __restrict__ y, const __half* __restrict__ gy, __half* __restrict__ gx, int64_t n8, int c8, int64_t ld,
                          float slope) {
    for (int64_t v = (int64_t)blockIdx.x * BA_THREADS + threadIdx.x; v < n8; v += (int64_t)gridDim.x * BA_THREADS) {
        const int64_t r = v / c8;
        const int c = (int)(v - r * c8);
        const uint4 ar = __ldg(reinterpret_cast<const uint4*>(y) + v);
        uint4 gr = __ldg(reinterpret_cast<const uint4*>(gy + r * ld) + c);
        const __half* a = reinterpret_cast<const __half*>(&ar);
        __half* g = reinterpret_cast<__half*>(&gr);
#pragma unroll
        for (int k = 0; k < 8; ++k) if (!(__half2float(a[k]) > 0.f)) g[k] = __float2half_rn(slope * __half2float(g[k]));
        reinterpret_cast<uint4*>(gx)[v] = gr;
    }
}

// dst[r][c] += src[r * ld + c]: a dense gradient plus a channel slice of a wider channels-last gradient (the skip branch of a
// DenseNet-style concatenation) in one vectorised pass — ATen adds a strided operand with its non-vectorised kernel
__global__ void __launch_bounds__(BA_THREADS)
add_rows_f32_kernel(float* dst, const float* a, const float* __restrict__ src, int64_t n4, int c4, int64_t ld) {      // dst may be a
    for (int64_t v = (int64_t)blockIdx.x * BA_THREADS + threadIdx.x; v < n4; v += (int64_t)gridDim.x * BA_THREADS) {
        const int64_t r = v / c4;
        const int c = (int)(v - r * c4);
        float4 d = reinterpret_cast<const float4*>(a)[v];
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + r * ld) + c);
        d.x += a.x; d.y += a.y; d.z += a.z; d.w += a.w;
        reinterpret_cast<float4*>(dst)[v] = d;
    }
}

// out = relu(a + b): the tail of a residual block (models/raft/extractor.py:56) in one pass instead of add + clamp
__global__ void __launch_bounds__(BA_THREADS)
add_relu_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n4) {
    for (int64_t v = (int64_t)blockIdx.x * BA_THREADS + threadIdx.x; v < n4; v += (int64_t)gridDim.x * BA_THREADS) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(a) + v), y = __ldg(reinterpret_cast<const float4*>(b) + v);
        reinterpret_cast<float4*>(out)[v] = make_float4(fmaxf(x.x + y.x, 0.f), fmaxf(x.y + y.y, 0.f), fmaxf(x.z + y.z, 0.f), fmaxf(x.w + y.w, 0.f));
    }
}

__global__ void __launch_bounds__(BA_THREADS)
add_relu_f16_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ out, int64_t n8) {
    for (int64_t v = (int64_t)blockIdx.x * BA_THREADS + threadIdx.x; v < n8; v += (int64_t)gridDim.x * BA_THREADS) {
        const uint4 xr = __ldg(reinterpret_cast<const uint4*>(a) + v), yr = __ldg(reinterpret_cast<const uint4*>(b) + v);
        const __half* x = reinterpret_cast<const __half*>(&xr);
        const __half* y = reinterpret_cast<const __half*>(&yr);
        uint4 o;
        __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
        for (int k = 0; k < 8; ++k) oh[k] = __float2half_rn(fmaxf(__half2float(__hadd(x[k], y[k])), 0.f));   // fp16 add like ATen
        reinterpret_cast<uint4*>(out)[v] = o;
    }
}

// One launch for the coordinate bookkeeping of a RAFT/GMA iteration (models/raft/raft.py:123-131): coords1 += delta_flow,
// flow = coords1 - coords0.  coords: [B,2,H,W]; delta: channels-last with `ld` channels per pixel (the flow head's padded
// output), channels 0..1 used; flow_cl: [B,H,W,2] (channels-last of [B,2,H,W]) for the next iteration's motion encoder.
template <typename TD>
__global__ void __launch_bounds__(256)
flow_step_kernel(const float* __restrict__ coords1, const float* __restrict__ coords0, const TD* __restrict__ delta, int ld,
                 float* __restrict__ new_coords1, float* __restrict__ flow_cl, int fld, int B, int N) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= B * N) return;
    const int b = i / N, p = i - b * N;
    float2 d;
    if (sizeof(TD) == 4) d = __ldg(reinterpret_cast<const float2*>(delta + (int64_t)i * ld));
    else d = __half22float2(*reinterpret_cast<const __half2*>(delta + (int64_t)i * ld));       // GMA: fp16 flow head
    const int64_t o = (int64_t)b * 2 * N + p;
    const float x = coords1[o] + d.x, y = coords1[o + N] + d.y;
    new_coords1[o] = x; new_coords1[o + N] = y;
    float2* f = reinterpret_cast<float2*>(flow_cl + (int64_t)i * fld);
    f[0] = make_float2(x - coords0[o], y - coords0[o + N]);
    for (int k = 1; k < fld / 2; ++k) f[k] = make_float2(0.f, 0.f);          // zero channels for a tensor-core convf1
}

static int ba_grid(int64_t nvec) {
    int64_t b = (nvec + BA_THREADS - 1) / BA_THREADS;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace pcfa

using namespace pcfa;

// x: n elements, in place.  dtype 0 = fp32 (vector 4), 1 = fp16 (vector 8).  channels-last: inner = 1 and C % vector == 0;
// NCHW: inner = H*W with inner % vector == 0.  n % vector == 0; x (and bias for channels-last) 16-byte aligned.
extern "C" int pcfa_bias_act_forward(void* x, const void* bias, int64_t n, int C, int64_t inner, int relu, float slope, int dtype,
                                     pcfa_stream_t stream) {
    if (!x || !bias || n <= 0 || C <= 0 || inner <= 0 || dtype < 0 || dtype > 1 || !(slope >= 0.f && slope < 1.f)) return PCFA_E_BADARG;
    const int vec = dtype == 0 ? 4 : 8;
    if (n % vec || (reinterpret_cast<uintptr_t>(x) & 15)) return PCFA_E_BADARG;
    if (inner == 1 ? (C % vec || (reinterpret_cast<uintptr_t>(bias) & 15)) : (inner % vec != 0)) return PCFA_E_BADARG;
    const int64_t nv = n / vec;
    cudaStream_t s = as_stream(stream);
    if (dtype == 0) {
        if (relu) bias_act_f32_kernel<true><<<ba_grid(nv), BA_THREADS, 0, s>>>((float*)x, (const float*)bias, nv, C, inner, slope);
        else      bias_act_f32_kernel<false><<<ba_grid(nv), BA_THREADS, 0, s>>>((float*)x, (const float*)bias, nv, C, inner, slope);
    } else {
        if (relu) bias_act_f16_kernel<true><<<ba_grid(nv), BA_THREADS, 0, s>>>((__half*)x, (const __half*)bias, nv, C, inner, slope);
        else      bias_act_f16_kernel<false><<<ba_grid(nv), BA_THREADS, 0, s>>>((__half*)x, (const __half*)bias, nv, C, inner, slope);
    }
    return after_launch();
}

// grad_x = y > 0 ? grad_y : slope * grad_y; any dense layout (element-wise), n % vector == 0, 16-byte aligned pointers
extern "C" int pcfa_relu_mask_backward(const void* y, const void* grad_y, void* grad_x, int64_t n, float slope, int dtype,
                                       pcfa_stream_t stream) {
    if (!y || !grad_y || !grad_x || n <= 0 || dtype < 0 || dtype > 1 || !(slope >= 0.f && slope < 1.f)) return PCFA_E_BADARG;
    const int vec = dtype == 0 ? 4 : 8;
    if (n % vec || ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(grad_y) | reinterpret_cast<uintptr_t>(grad_x)) & 15))
        return PCFA_E_BADARG;
    const int64_t nv = n / vec;
    if (dtype == 0) relu_mask_f32_kernel<<<ba_grid(nv), BA_THREADS, 0, as_stream(stream)>>>((const float*)y, (const float*)grad_y, (float*)grad_x, nv, slope);
    else            relu_mask_f16_kernel<<<ba_grid(nv), BA_THREADS, 0, as_stream(stream)>>>((const __half*)y, (const __half*)grad_y, (__half*)grad_x, nv, slope);
    return after_launch();
}

// grad_y rows `ld` elements apart (a channel slice of a channels-last tensor); y and grad_x dense [rows][C]
extern "C" int pcfa_relu_mask_backward_rows(const void* y, const void* grad_y, void* grad_x, int64_t rows, int C, int64_t ld, float slope,
                                            int dtype, pcfa_stream_t stream) {
    if (!y || !grad_y || !grad_x || rows <= 0 || C <= 0 || ld < C || dtype < 0 || dtype > 1 || !(slope >= 0.f && slope < 1.f)) return PCFA_E_BADARG;
    const int vec = dtype == 0 ? 4 : 8;
    if (C % vec || ld % vec || ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(grad_y) | reinterpret_cast<uintptr_t>(grad_x)) & 15))
        return PCFA_E_BADARG;
    const int64_t nv = rows * (C / vec);
    if (dtype == 0) relu_mask_rows_f32_kernel<<<ba_grid(nv), BA_THREADS, 0, as_stream(stream)>>>((const float*)y, (const float*)grad_y, (float*)grad_x, nv, C / vec, ld, slope);
    else            relu_mask_rows_f16_kernel<<<ba_grid(nv), BA_THREADS, 0, as_stream(stream)>>>((const __half*)y, (const __half*)grad_y, (__half*)grad_x, nv, C / vec, ld, slope);
    return after_launch();
}

extern "C" int pcfa_add_relu_forward(const void* a, const void* b, void* out, int64_t n, int dtype, pcfa_stream_t stream) {
    if (!a || !b || !out || n <= 0 || dtype < 0 || dtype > 1) return PCFA_E_BADARG;
    const int vec = dtype == 0 ? 4 : 8;
    if (n % vec || ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15)) return PCFA_E_BADARG;
    const int64_t nv = n / vec;
    if (dtype == 0) add_relu_f32_kernel<<<ba_grid(nv), BA_THREADS, 0, as_stream(stream)>>>((const float*)a, (const float*)b, (float*)out, nv);
    else            add_relu_f16_kernel<<<ba_grid(nv), BA_THREADS, 0, as_stream(stream)>>>((const __half*)a, (const __half*)b, (__half*)out, nv);
    return after_launch();
}

extern "C" int pcfa_flow_step(const float* coords1, const float* coords0, const void* delta, int delta_ld, int delta_dtype,
                              float* new_coords1, float* flow_cl, int flow_ld, int B, int H, int W, pcfa_stream_t stream) {
    if (delta_dtype < 0 || delta_dtype > 1) return PCFA_E_BADARG;
    if (!coords1 || !coords0 || !delta || !new_coords1 || !flow_cl || B <= 0 || H <= 0 || W <= 0 || delta_ld < 2 || (delta_ld & 1) ||
        flow_ld < 2 || (flow_ld & 1)) return PCFA_E_BADARG;
    if ((int64_t)B * H * W > 0x7fffffffLL) return PCFA_E_TOOLARGE;
    if ((reinterpret_cast<uintptr_t>(delta) & (delta_dtype ? 3 : 7)) || (reinterpret_cast<uintptr_t>(flow_cl) & 7)) return PCFA_E_BADARG;
    const int n = B * H * W;
    if (delta_dtype == 0)
        flow_step_kernel<float><<<(n + 255) / 256, 256, 0, as_stream(stream)>>>((const float*)coords1, coords0, (const float*)delta, delta_ld, new_coords1, flow_cl, flow_ld, B, H * W);
    else
        flow_step_kernel<__half><<<(n + 255) / 256, 256, 0, as_stream(stream)>>>((const float*)coords1, coords0, (const __half*)delta, delta_ld, new_coords1, flow_cl, flow_ld, B, H * W);
    return after_launch();
}

extern "C" int pcfa_add_rows_inplace(float* dst, const float* src, int64_t rows, int C, int64_t ld, pcfa_stream_t stream) {
    if (!dst || !src || rows <= 0 || C <= 0 || ld < C || C % 4 || ld % 4 ||
        ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15)) return PCFA_E_BADARG;
    const int64_t nv = rows * (C / 4);
    add_rows_f32_kernel<<<ba_grid(nv), BA_THREADS, 0, as_stream(stream)>>>(dst, dst, src, nv, C / 4, ld);
    return after_launch();
}

// out[r][c] = a[r][c] + src[r*ld + c]  (out and a dense [rows][C]): the sum of a dense gradient and a channel slice of a wider
// channels-last gradient where a tensor feeds both a convolution and a later concatenation (FlowNet's encoder skips)
extern "C" int pcfa_add_rows(float* out, const float* a, const float* src, int64_t rows, int C, int64_t ld, pcfa_stream_t stream) {
    if (!out || !a || !src || rows <= 0 || C <= 0 || ld < C || C % 4 || ld % 4 ||
        ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(src)) & 15)) return PCFA_E_BADARG;
    const int64_t nv = rows * (C / 4);
    add_rows_f32_kernel<<<ba_grid(nv), BA_THREADS, 0, as_stream(stream)>>>(out, a, src, nv, C / 4, ld);
    return after_launch();
}

extern "C" int pcfa_relu_mask2_backward(const float* y, const float* g1, const float* g2, float* grad_x, int64_t n, pcfa_stream_t stream) {
    if (!y || !g1 || !g2 || !grad_x || n <= 0 || n % 4 ||
        ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(g1) | reinterpret_cast<uintptr_t>(g2) | reinterpret_cast<uintptr_t>(grad_x)) & 15))
        return PCFA_E_BADARG;
    relu_mask2_f32_kernel<<<ba_grid(n / 4), BA_THREADS, 0, as_stream(stream)>>>(y, g1, g2, grad_x, n / 4);
    return after_launch();
}
