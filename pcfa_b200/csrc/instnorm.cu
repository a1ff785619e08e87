// Instance normalisation (+ optional fused ReLU), forward and backward, NCHW fp32.
//
// RAFT's feature encoder normalises every convolution output with nn.InstanceNorm2d (affine=False, eps 1e-5,
// biased variance; models/raft/extractor.py:13-55,118-150) followed by ReLU.  ATen runs that as batch-norm kernels on a
// [1, B*C, H, W] view plus separate ReLU kernels: 15 norms cost 3.1 ms of the 16.8 ms RAFT closure at 436x1024
// (scripts/profile_closure.py), at ~0.9 TB/s.  These kernels are plain streaming passes:
//   forward : stats (read x) ; apply y = relu((x - mean) * rstd) (read x, write y)            3 passes
//   backward: sums  s1 = sum g, s2 = sum g*xhat with g = dy * [xhat > 0] (read x, dy) ;
//             dx = rstd * (g - s1/n - xhat * s2/n) (read x, dy, write dx)                      5 passes
// A plane (b, c) is cut into `splits` contiguous chunks, one CTA each, so small batches still fill the GPU; the chunk
// partials are combined in double by every CTA of the second kernel (deterministic, no atomics).
#include "common.cuh"
#include <cuda_fp16.h>

namespace pcfa {

constexpr int IN_THREADS = 256;
constexpr int IN_MAX_SPLITS = 64;

struct InChunk { int64_t lo, hi; };
__device__ __forceinline__ InChunk in_chunk(int64_t hw, int splits, int split) {
    int64_t per = (hw + splits - 1) / splits;
    per = (per + 3) & ~(int64_t)3;                       // chunks start on 16-byte boundaries when the plane does
    InChunk c;
    c.lo = per * split; c.hi = c.lo + per;
    if (c.lo > hw) c.lo = hw;
    if (c.hi > hw) c.hi = hw;
    return c;
}

__device__ __forceinline__ float2 block_sum2(float a, float b, float2* sh) {
    a = warp_sum(a); b = warp_sum(b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = make_float2(a, b);
    __syncthreads();
    float2 r = make_float2(0.f, 0.f);
    if (warp == 0) {
        float2 v = lane < IN_THREADS / 32 ? sh[lane] : make_float2(0.f, 0.f);
        r.x = warp_sum(v.x); r.y = warp_sum(v.y);
    }
    __syncthreads();
    return r;                                            // valid in warp 0
}

// MODE 0: (sum x, sum x^2).  MODE 1: (sum g, sum g*xhat), g = dy * mask.
template <int MODE>
__global__ void __launch_bounds__(IN_THREADS)
instnorm_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float2* __restrict__ stats,
                        float2* __restrict__ part, int64_t hw, int splits, int relu, int vec) {
    __shared__ float2 sh[IN_THREADS / 32];
    const int64_t plane = blockIdx.y;
    const InChunk c = in_chunk(hw, splits, blockIdx.x);
    const float* xp = x + plane * hw;
    const float* gp = MODE == 1 ? dy + plane * hw : nullptr;
    float mean = 0.f, rstd = 1.f;
    if (MODE == 1) { const float2 st = stats[plane]; mean = st.x; rstd = st.y; }
    float a = 0.f, b = 0.f;
    auto acc = [&](float xv, float gv) {
        if (MODE == 0) { a += xv; b = fmaf(xv, xv, b); }
        else {
            const float xh = (xv - mean) * rstd;
            const float g = (relu && !(xh > 0.f)) ? 0.f : gv;
            a += g; b = fmaf(g, xh, b);
        }
    };
    if (vec) {
        for (int64_t i = c.lo + 4 * (int64_t)threadIdx.x; i < c.hi; i += 4 * IN_THREADS) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + i));
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE == 1) gv = __ldg(reinterpret_cast<const float4*>(gp + i));
            acc(xv.x, gv.x); acc(xv.y, gv.y); acc(xv.z, gv.z); acc(xv.w, gv.w);
        }
    } else {
        for (int64_t i = c.lo + threadIdx.x; i < c.hi; i += IN_THREADS) acc(__ldg(xp + i), MODE == 1 ? __ldg(gp + i) : 0.f);
    }
    const float2 r = block_sum2(a, b, sh);
    if (threadIdx.x == 0) part[plane * splits + blockIdx.x] = r;
}

// MODE 0: y = relu((x - mean) * rstd), writes stats (mean, rstd).  MODE 1: dx.
template <int MODE>
__global__ void __launch_bounds__(IN_THREADS)
instnorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float2* __restrict__ part,
                      float2* __restrict__ stats, float* __restrict__ out, int64_t hw, int splits, float eps, int relu,
                      int vec) {
    __shared__ float2 sh;
    const int64_t plane = blockIdx.y;
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int k = 0; k < splits; ++k) { const float2 p = part[plane * splits + k]; s0 += (double)p.x; s1 += (double)p.y; }
        const double n = (double)hw;
        if (MODE == 0) {
            const double m = s0 / n;
            double var = s1 / n - m * m;                 // biased variance (F.instance_norm, use_input_stats)
            if (var < 0.0) var = 0.0;
            sh = make_float2((float)m, (float)(1.0 / sqrt(var + (double)eps)));
            if (blockIdx.x == 0) stats[plane] = sh;
        } else {
            sh = make_float2((float)(s0 / n), (float)(s1 / n));
        }
    }
    __syncthreads();
    const InChunk c = in_chunk(hw, splits, blockIdx.x);
    const float* xp = x + plane * hw;
    float* op = out + plane * hw;
    if (MODE == 0) {
        const float mean = sh.x, rstd = sh.y;
        auto f = [&](float xv) { const float v = (xv - mean) * rstd; return (relu && !(v > 0.f)) ? 0.f : v; };
        if (vec) {
            for (int64_t i = c.lo + 4 * (int64_t)threadIdx.x; i < c.hi; i += 4 * IN_THREADS) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + i));
                *reinterpret_cast<float4*>(op + i) = make_float4(f(xv.x), f(xv.y), f(xv.z), f(xv.w));
            }
        } else {
            for (int64_t i = c.lo + threadIdx.x; i < c.hi; i += IN_THREADS) op[i] = f(__ldg(xp + i));
        }
    } else {
        const float2 st = stats[plane];
        const float mean = st.x, rstd = st.y, m1 = sh.x, m2 = sh.y;
        const float* gp = dy + plane * hw;
        auto f = [&](float xv, float gv) {
            const float xh = (xv - mean) * rstd;
            const float g = (relu && !(xh > 0.f)) ? 0.f : gv;
            return rstd * (g - m1 - xh * m2);
        };
        if (vec) {
            for (int64_t i = c.lo + 4 * (int64_t)threadIdx.x; i < c.hi; i += 4 * IN_THREADS) {
                const float4 xv = __ldg(reinterpret_cast<const float4*>(xp + i));
                const float4 gv = __ldg(reinterpret_cast<const float4*>(gp + i));
                *reinterpret_cast<float4*>(op + i) = make_float4(f(xv.x, gv.x), f(xv.y, gv.y), f(xv.z, gv.z), f(xv.w, gv.w));
            }
        } else {
            for (int64_t i = c.lo + threadIdx.x; i < c.hi; i += IN_THREADS) op[i] = f(__ldg(xp + i), __ldg(gp + i));
        }
    }
}

// ------------------------------------------------------------------------------------ channels-last (NHWC) memory
// x is [B][H*W][C] in memory (torch.channels_last): a plane's statistics are a column reduction.  Thread = (row slot r,
// channel quad cg): C/4 threads cover one pixel with 128-bit accesses, 1024/(C/4) pixels per iteration and U rows in
// flight per thread (64 KB of loads in flight per SM); the row slots are combined through shared memory.  About one
// CTA per SM: few chunk partials ([B][splits][C] float2), so every CTA of the second kernel can afford to reduce them
// itself (deterministic, in double) instead of a third launch.  Requires C % 4 == 0, C <= 1024.
constexpr int IN_MAX_C = 1024;
constexpr int INL_THREADS = 1024;
constexpr int INL_U = 4;

// 4 consecutive channels as float4 from fp32 or fp16 storage (the fp16 variant serves GMA's autocast encoder: the
// convolution output is normalised straight from half precision with fp32 arithmetic, which is what autocast makes of
// F.instance_norm(conv_out.float()) without the 58 MB conversion copies around it)
template <typename T> struct In4;
template <> struct In4<float> {
    static __device__ __forceinline__ float4 ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct In4<__half> {
    static __device__ __forceinline__ float4 ld(const __half* p) {
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    static __device__ __forceinline__ void st(__half* p, float4 v) {
        uint2 r;
        *reinterpret_cast<__half2*>(&r.x) = __floats2half2_rn(v.x, v.y);
        *reinterpret_cast<__half2*>(&r.y) = __floats2half2_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(p) = r;
    }
};

template <int MODE, typename TX>
__global__ void __launch_bounds__(INL_THREADS)
instnorm_partial_nhwc_kernel(const TX* __restrict__ x, const float* __restrict__ dy, const float2* __restrict__ stats,
                             float2* __restrict__ part, int64_t hw, int C, int splits, int relu) {
    extern __shared__ float sm[];                         // [2][rpi][C]
    const int b = blockIdx.y, tpr = C >> 2, rpi = INL_THREADS / tpr;
    const int r = threadIdx.x / tpr, cg = threadIdx.x - r * tpr;
    const int64_t per = (hw + splits - 1) / splits;
    const int64_t lo = per * blockIdx.x, hi = (lo + per < hw) ? lo + per : hw;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < rpi) {
        float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {1.f, 1.f, 1.f, 1.f};
        if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 st = stats[(int64_t)b * C + 4 * cg + k]; mean[k] = st.x; rstd[k] = st.y; }
        }
        const TX* xp = x + ((int64_t)b * hw) * C + 4 * cg;
        const float* gp = MODE == 1 ? dy + ((int64_t)b * hw) * C + 4 * cg : nullptr;
        for (int64_t row0 = lo + r; row0 < hi; row0 += (int64_t)INL_U * rpi) {
            float4 xv4[INL_U], gv4[INL_U];
#pragma unroll
            for (int u = 0; u < INL_U; ++u) {
                const int64_t row = row0 + (int64_t)u * rpi;
                xv4[u] = make_float4(0.f, 0.f, 0.f, 0.f); gv4[u] = xv4[u];
                if (row < hi) {
                    xv4[u] = In4<TX>::ld(xp + row * C);
                    if (MODE == 1) gv4[u] = __ldg(reinterpret_cast<const float4*>(gp + row * C));
                }
            }
#pragma unroll
            for (int u = 0; u < INL_U; ++u) {
                if (row0 + (int64_t)u * rpi >= hi) break;
                const float xv[4] = {xv4[u].x, xv4[u].y, xv4[u].z, xv4[u].w};
                const float gv[4] = {gv4[u].x, gv4[u].y, gv4[u].z, gv4[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (MODE == 0) { a[k] += xv[k]; q[k] = fmaf(xv[k], xv[k], q[k]); }
                    else {
                        const float xh = (xv[k] - mean[k]) * rstd[k];
                        const float g = (relu && !(xh > 0.f)) ? 0.f : gv[k];
                        a[k] += g; q[k] = fmaf(g, xh, q[k]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { sm[r * C + 4 * cg + k] = a[k]; sm[(rpi + r) * C + 4 * cg + k] = q[k]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += INL_THREADS) {
        float s0 = 0.f, s1 = 0.f;
        for (int k = 0; k < rpi; ++k) { s0 += sm[k * C + c]; s1 += sm[(rpi + k) * C + c]; }
        part[((int64_t)b * splits + blockIdx.x) * C + c] = make_float2(s0, s1);
    }
}

// TO: forward (MODE 0) writes fp32 (autocast's instance norm returns fp32); backward (MODE 1) writes the gradient in x's type
template <int MODE, typename TX, typename TO>
__global__ void __launch_bounds__(INL_THREADS)
instnorm_apply_nhwc_kernel(const TX* __restrict__ x, const float* __restrict__ dy, const float2* __restrict__ part,
                           float2* __restrict__ stats, TO* __restrict__ out, int64_t hw, int C, int splits, float eps,
                           int relu, int reverse) {
    extern __shared__ float sm[];                         // [4][C] mean, rstd, m1, m2 ; then [slots][C] double2 scratch
    float* s_mean = sm; float* s_rstd = sm + C; float* s_m1 = sm + 2 * C; float* s_m2 = sm + 3 * C;
    double* scratch = reinterpret_cast<double*>(sm + 4 * C);
    const int b = blockIdx.y;
    {   // every CTA reduces the chunk partials of its image: slot j sums chunks j, j + slots, ... (coalesced over c)
        const int slots = INL_THREADS / C > 0 ? INL_THREADS / C : 1;
        const int j = threadIdx.x / C, c = threadIdx.x - j * C;
        if (j < slots) {
            double s0 = 0.0, s1 = 0.0;
            for (int k = j; k < splits; k += slots) { const float2 p = part[((int64_t)b * splits + k) * C + c]; s0 += (double)p.x; s1 += (double)p.y; }
            scratch[2 * (j * C + c)] = s0; scratch[2 * (j * C + c) + 1] = s1;
        }
        __syncthreads();
        for (int cc = threadIdx.x; cc < C; cc += INL_THREADS) {
            double s0 = 0.0, s1 = 0.0;
            for (int k = 0; k < slots; ++k) { s0 += scratch[2 * (k * C + cc)]; s1 += scratch[2 * (k * C + cc) + 1]; }
            const double n = (double)hw;
            if (MODE == 0) {
                const double m = s0 / n;
                double var = s1 / n - m * m;
                if (var < 0.0) var = 0.0;
                const float2 st = make_float2((float)m, (float)(1.0 / sqrt(var + (double)eps)));
                s_mean[cc] = st.x; s_rstd[cc] = st.y;
                if (blockIdx.x == 0) stats[(int64_t)b * C + cc] = st;
            } else {
                const float2 st = stats[(int64_t)b * C + cc];
                s_mean[cc] = st.x; s_rstd[cc] = st.y; s_m1[cc] = (float)(s0 / n); s_m2[cc] = (float)(s1 / n);
            }
        }
        __syncthreads();
    }
    const int tpr = C >> 2, rpi = INL_THREADS / tpr;
    const int r = threadIdx.x / tpr, cg = threadIdx.x - r * tpr;
    if (r >= rpi) return;
    const int64_t per = (hw + splits - 1) / splits;
    const int64_t lo = per * blockIdx.x, hi = (lo + per < hw) ? lo + per : hw;
    float mean[4], rstd[4], m1[4], m2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        mean[k] = s_mean[4 * cg + k]; rstd[k] = s_rstd[4 * cg + k];
        m1[k] = MODE == 1 ? s_m1[4 * cg + k] : 0.f; m2[k] = MODE == 1 ? s_m2[4 * cg + k] : 0.f;
    }
    const TX* xp = x + ((int64_t)b * hw) * C + 4 * cg;
    const float* gp = MODE == 1 ? dy + ((int64_t)b * hw) * C + 4 * cg : nullptr;
    TO* op = out + ((int64_t)b * hw) * C + 4 * cg;
    // The chunk is walked BACKWARDS: the statistics kernel that ran just before walked it forwards, so the chunk's tail is
    // what is most likely still in L2 (a 58 MB plane set against 126 MB of L2 shared with the output being written).
    const int64_t step = (int64_t)INL_U * rpi;
    const int64_t iters = (hi - lo - r + step - 1) / step;
    for (int64_t it = iters - 1; it >= 0; --it) {
        const int64_t row0 = lo + r + (reverse ? it : iters - 1 - it) * step;
        float4 xv4[INL_U], gv4[INL_U];
#pragma unroll
        for (int u = 0; u < INL_U; ++u) {
            const int64_t row = row0 + (int64_t)u * rpi;
            xv4[u] = make_float4(0.f, 0.f, 0.f, 0.f); gv4[u] = xv4[u];
            if (row < hi) {
                xv4[u] = In4<TX>::ld(xp + row * C);
                if (MODE == 1) gv4[u] = __ldg(reinterpret_cast<const float4*>(gp + row * C));
            }
        }
#pragma unroll
        for (int u = 0; u < INL_U; ++u) {
            const int64_t row = row0 + (int64_t)u * rpi;
            if (row >= hi) break;
            const float xv[4] = {xv4[u].x, xv4[u].y, xv4[u].z, xv4[u].w};
            const float gv[4] = {gv4[u].x, gv4[u].y, gv4[u].z, gv4[u].w};
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float xh = (xv[k] - mean[k]) * rstd[k];
                if (MODE == 0) o[k] = (relu && !(xh > 0.f)) ? 0.f : xh;
                else {
                    const float g = (relu && !(xh > 0.f)) ? 0.f : gv[k];
                    o[k] = rstd[k] * (g - m1[k] - xh * m2[k]);
                }
            }
            In4<TO>::st(op + row * C, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

static int in_reverse() { static const int v = [] { const char* e = getenv("PCFA_IN_REVERSE"); return e ? atoi(e) : 1; }(); return v; }

static int in_splits_nhwc(int B, int64_t hw) {
    int64_t s = ((int64_t)kNumSMs + B - 1) / B;                      // about one 1024-thread CTA per SM in total
    const int64_t max_by_size = hw / 256 > 0 ? hw / 256 : 1;
    if (s > max_by_size) s = max_by_size;
    return (int)(s < 1 ? 1 : s);
}

static bool nhwc_ok(int C, const void* a, const void* b, const void* c) {
    return C % 4 == 0 && C >= 4 && C <= IN_MAX_C &&
           ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

static int in_splits(int64_t planes, int64_t hw) {
    int64_t s = ((int64_t)kNumSMs * 4 + planes - 1) / planes;       // ~4 CTAs per SM in total
    const int64_t max_by_size = hw / 2048 > 0 ? hw / 2048 : 1;      // at least 2048 elements per chunk
    if (s > max_by_size) s = max_by_size;
    if (s > IN_MAX_SPLITS) s = IN_MAX_SPLITS;
    return (int)(s < 1 ? 1 : s);
}

}  // namespace pcfa

using namespace pcfa;

extern "C" int64_t pcfa_instnorm_workspace_bytes(int B, int C, int H, int W) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    const int64_t nchw = (int64_t)B * C * IN_MAX_SPLITS, nhwc = (int64_t)B * C * (in_splits_nhwc(B, (int64_t)H * W) + 1);
    return (nchw > nhwc ? nchw : nhwc) * (int64_t)sizeof(float2);
}

extern "C" int pcfa_instnorm_forward(const float* x, float* y, float* stats, void* workspace, int B, int C, int H, int W,
                                     float eps, int relu, int channels_last, pcfa_stream_t stream) {
    if (!x || !y || !stats || !workspace || B <= 0 || C <= 0 || H <= 0 || W <= 0) return PCFA_E_BADARG;
    const int64_t planes = (int64_t)B * C, hw = (int64_t)H * W;
    if (channels_last) {
        if (!nhwc_ok(C, x, y, y) || B > 65535) return PCFA_E_BADARG;
        const int splits = in_splits_nhwc(B, hw);
        const int rpi = INL_THREADS / (C / 4);
        const int slots = INL_THREADS / C > 0 ? INL_THREADS / C : 1;
        const size_t sm_apply = 4 * C * sizeof(float) + (size_t)slots * C * 2 * sizeof(double);
        cudaStream_t s = as_stream(stream);
        float2* part = reinterpret_cast<float2*>(workspace);
        dim3 grid(splits, B);
        instnorm_partial_nhwc_kernel<0, float><<<grid, INL_THREADS, 2 * rpi * C * sizeof(float), s>>>(x, nullptr, nullptr, part, hw, C, splits, relu);
        PCFA_TRY(after_launch());
        instnorm_apply_nhwc_kernel<0, float, float><<<grid, INL_THREADS, sm_apply, s>>>(x, nullptr, part, reinterpret_cast<float2*>(stats),
                                                                          y, hw, C, splits, eps, relu, in_reverse());
        return after_launch();
    }
    if (planes > 65535) return PCFA_E_TOOLARGE;
    const int splits = in_splits(planes, hw);
    const int vec = (hw % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) ? 1 : 0;
    cudaStream_t s = as_stream(stream);
    float2* part = reinterpret_cast<float2*>(workspace);
    dim3 grid(splits, (unsigned)planes);
    instnorm_partial_kernel<0><<<grid, IN_THREADS, 0, s>>>(x, nullptr, nullptr, part, hw, splits, relu, vec);
    PCFA_TRY(after_launch());
    instnorm_apply_kernel<0><<<grid, IN_THREADS, 0, s>>>(x, nullptr, part, reinterpret_cast<float2*>(stats), y, hw, splits,
                                                         eps, relu, vec);
    return after_launch();
}

extern "C" int pcfa_instnorm_backward(const float* x, const float* grad_y, const float* stats, float* grad_x,
                                      void* workspace, int B, int C, int H, int W, int relu, int channels_last,
                                      pcfa_stream_t stream) {
    if (!x || !grad_y || !stats || !grad_x || !workspace || B <= 0 || C <= 0 || H <= 0 || W <= 0) return PCFA_E_BADARG;
    const int64_t planes = (int64_t)B * C, hw = (int64_t)H * W;
    if (channels_last) {
        if (!nhwc_ok(C, x, grad_y, grad_x) || B > 65535) return PCFA_E_BADARG;
        const int splits = in_splits_nhwc(B, hw);
        const int rpi = INL_THREADS / (C / 4);
        const int slots = INL_THREADS / C > 0 ? INL_THREADS / C : 1;
        const size_t sm_apply = 4 * C * sizeof(float) + (size_t)slots * C * 2 * sizeof(double);
        cudaStream_t s = as_stream(stream);
        float2* part = reinterpret_cast<float2*>(workspace);
        float2* st = reinterpret_cast<float2*>(const_cast<float*>(stats));
        dim3 grid(splits, B);
        instnorm_partial_nhwc_kernel<1, float><<<grid, INL_THREADS, 2 * rpi * C * sizeof(float), s>>>(x, grad_y, st, part, hw, C, splits, relu);
        PCFA_TRY(after_launch());
        instnorm_apply_nhwc_kernel<1, float, float><<<grid, INL_THREADS, sm_apply, s>>>(x, grad_y, part, st, grad_x, hw, C, splits, 0.f, relu, in_reverse());
        return after_launch();
    }
    if (planes > 65535) return PCFA_E_TOOLARGE;
    const int splits = in_splits(planes, hw);
    const int vec = (hw % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad_y) |
                                      reinterpret_cast<uintptr_t>(grad_x)) & 15) == 0) ? 1 : 0;
    cudaStream_t s = as_stream(stream);
    float2* part = reinterpret_cast<float2*>(workspace);
    const float2* st = reinterpret_cast<const float2*>(stats);
    dim3 grid(splits, (unsigned)planes);
    instnorm_partial_kernel<1><<<grid, IN_THREADS, 0, s>>>(x, grad_y, st, part, hw, splits, relu, vec);
    PCFA_TRY(after_launch());
    instnorm_apply_kernel<1><<<grid, IN_THREADS, 0, s>>>(x, grad_y, part, const_cast<float2*>(st), grad_x, hw, splits, 0.f,
                                                         relu, vec);
    return after_launch();
}

// Channels-last, x in fp16 (GMA's autocast encoder): y fp32; backward: grad_y fp32, grad_x fp16.  Same kernels, fp32 arithmetic.
extern "C" int pcfa_instnorm_forward_h(const void* x_half, float* y, float* stats, void* workspace, int B, int C, int H, int W,
                                       float eps, int relu, pcfa_stream_t stream) {
    if (!x_half || !y || !stats || !workspace || B <= 0 || C <= 0 || H <= 0 || W <= 0) return PCFA_E_BADARG;
    if (!nhwc_ok(C, y, y, y) || (reinterpret_cast<uintptr_t>(x_half) & 7) || B > 65535) return PCFA_E_BADARG;
    const int64_t hw = (int64_t)H * W;
    const int splits = in_splits_nhwc(B, hw);
    const int rpi = INL_THREADS / (C / 4);
    const int slots = INL_THREADS / C > 0 ? INL_THREADS / C : 1;
    const size_t sm_apply = 4 * C * sizeof(float) + (size_t)slots * C * 2 * sizeof(double);
    cudaStream_t s = as_stream(stream);
    float2* part = reinterpret_cast<float2*>(workspace);
    const __half* x = reinterpret_cast<const __half*>(x_half);
    dim3 grid(splits, B);
    instnorm_partial_nhwc_kernel<0, __half><<<grid, INL_THREADS, 2 * rpi * C * sizeof(float), s>>>(x, nullptr, nullptr, part, hw, C, splits, relu);
    PCFA_TRY(after_launch());
    instnorm_apply_nhwc_kernel<0, __half, float><<<grid, INL_THREADS, sm_apply, s>>>(x, nullptr, part, reinterpret_cast<float2*>(stats), y, hw, C,
                                                                                     splits, eps, relu, in_reverse());
    return after_launch();
}

extern "C" int pcfa_instnorm_backward_h(const void* x_half, const float* grad_y, const float* stats, void* grad_x_half, void* workspace,
                                        int B, int C, int H, int W, int relu, pcfa_stream_t stream) {
    if (!x_half || !grad_y || !stats || !grad_x_half || !workspace || B <= 0 || C <= 0 || H <= 0 || W <= 0) return PCFA_E_BADARG;
    if (!nhwc_ok(C, grad_y, grad_y, grad_y) || ((reinterpret_cast<uintptr_t>(x_half) | reinterpret_cast<uintptr_t>(grad_x_half)) & 7) || B > 65535)
        return PCFA_E_BADARG;
    const int64_t hw = (int64_t)H * W;
    const int splits = in_splits_nhwc(B, hw);
    const int rpi = INL_THREADS / (C / 4);
    const int slots = INL_THREADS / C > 0 ? INL_THREADS / C : 1;
    const size_t sm_apply = 4 * C * sizeof(float) + (size_t)slots * C * 2 * sizeof(double);
    cudaStream_t s = as_stream(stream);
    float2* part = reinterpret_cast<float2*>(workspace);
    float2* st = reinterpret_cast<float2*>(const_cast<float*>(stats));
    const __half* x = reinterpret_cast<const __half*>(x_half);
    dim3 grid(splits, B);
    instnorm_partial_nhwc_kernel<1, __half><<<grid, INL_THREADS, 2 * rpi * C * sizeof(float), s>>>(x, grad_y, st, part, hw, C, splits, relu);
    PCFA_TRY(after_launch());
    instnorm_apply_nhwc_kernel<1, __half, __half><<<grid, INL_THREADS, sm_apply, s>>>(x, grad_y, part, st, reinterpret_cast<__half*>(grad_x_half),
                                                                                      hw, C, splits, 0.f, relu, in_reverse());
    return after_launch();
}
