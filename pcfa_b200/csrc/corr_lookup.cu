// Multi-level bilinear correlation lookup (forward gather, backward scatter).
// Replaces CorrBlock.__call__ / bilinear_sampler / F.grid_sample(align_corners=True)
// (reference: models/raft/corr.py:29-50, models/raft/utils/utils.py:57-71).
//
// Geometry per (query q, level l): all (2r+1)^2 taps share one fractional offset (fx, fy), so the
// taps read a (2r+2)x(2r+2) footprint of level l's row q.  A CTA owns QB=32 consecutive queries of
// one sample at one level: it stages the 32 footprints in shared memory (coalesced-ish 40-byte row
// segments), then each warp walks output channels with lane == query, which makes every global
// store a full 128-byte line of out[b, ch, q0:q0+32].
// HBM-bound gather: algorithmic bytes per query per level = (2r+2)^2*4 read + (2r+1)^2*4 written.
#include "common.cuh"

namespace pcfa {

constexpr int QB = 32;          // queries per CTA
constexpr int LOOKUP_THREADS = 256;

struct LookupGeom {
    int   ix0, iy0;   // top-left cell of the footprint (may be out of range)
    float fx, fy;
};

__device__ __forceinline__ LookupGeom lookup_geom(float cx, float cy, int level, int r) {
    // centroid_lvl = coords / 2**i  (corr.py:41) — exact power-of-two scaling
    const float s  = 1.0f / (float)(1 << level);
    const float sx = cx * s, sy = cy * s;
    const float flx = floorf(sx), fly = floorf(sy);
    LookupGeom g;
    // clamp so that absurd coordinates cannot overflow int arithmetic; anything this far out
    // samples only zero padding anyway.
    g.ix0 = (int)fminf(fmaxf(flx, -1.0e6f), 1.0e6f) - r;
    g.iy0 = (int)fminf(fmaxf(fly, -1.0e6f), 1.0e6f) - r;
    g.fx  = sx - flx;
    g.fy  = sy - fly;
    return g;
}

// grid: (ceil(N/QB) * B, levels).  RT > 0: radius known at compile time (all index divisions become
// multiply-shift; the runtime-radius instantiation spent most of its issue slots on integer division).
template <int RT>
__global__ void __launch_bounds__(LOOKUP_THREADS)
corr_lookup_fwd_kernel(const float* __restrict__ pyramid, const float* __restrict__ coords,
                       float* __restrict__ out, PyramidLayout L, int B, int H, int W, int r_rt, int out_cl) {
    extern __shared__ float smem[];
    const int r = RT > 0 ? RT : r_rt;
    const int D = 2 * r + 1, F = D + 1, FP = F * F;
    const int FS = FP | 1;                       // odd stride → conflict-free lane==query reads
    float*      S    = smem;                     // [QB][FS]
    LookupGeom* geom = reinterpret_cast<LookupGeom*>(smem + QB * FS);   // [QB]

    const int N = H * W;
    const int groups = ceil_div(N, QB);
    const int b  = blockIdx.x / groups;
    const int q0 = (blockIdx.x % groups) * QB;
    const int l  = blockIdx.y;
    const int Hl = L.h[l], Wl = L.w[l];
    const int nq = min(QB, N - q0);

    if (threadIdx.x < QB) {
        const int q = min(q0 + (int)threadIdx.x, N - 1);
        const float cx = coords[((int64_t)b * 2 + 0) * N + q];
        const float cy = coords[((int64_t)b * 2 + 1) * N + q];
        geom[threadIdx.x] = lookup_geom(cx, cy, l, r);
    }
    __syncthreads();

    // gather: one warp per query, lanes walk the footprint cells (cell -> (row, col) is the same for every
    // query, so it is computed once per thread; everything per query is warp-uniform)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* lvl = pyramid + L.off[l];
    const int64_t plane = (int64_t)Hl * Wl;
    constexpr int kMaxPass = 8;                   // (2r+2)^2 <= 256 cells for r <= 7; larger radii loop more
    const int npass = ceil_div(FP, 32);
    for (int qi = warp; qi < nq; qi += LOOKUP_THREADS / 32) {
        const LookupGeom g = geom[qi];
        const float* src = lvl + ((int64_t)b * N + q0 + qi) * plane;
        float* dstq = S + qi * FS;
#pragma unroll
        for (int ps = 0; ps < kMaxPass; ++ps) {
            if (ps >= npass) break;
            const int cell = ps * 32 + lane;
            const int rr = cell / F, cc = cell - rr * F;
            const int y = g.iy0 + rr, x = g.ix0 + cc;
            if (cell < FP) {
                float v = 0.f;
                if ((unsigned)y < (unsigned)Hl && (unsigned)x < (unsigned)Wl) v = __ldg(src + y * Wl + x);
                dstq[cell] = v;
            }
        }
        for (int cell = kMaxPass * 32 + lane; cell < FP; cell += 32) {     // r > 7 only
            const int rr = cell / F, cc = cell - rr * F;
            const int y = g.iy0 + rr, x = g.ix0 + cc;
            dstq[cell] = ((unsigned)y < (unsigned)Hl && (unsigned)x < (unsigned)Wl) ? __ldg(src + y * Wl + x) : 0.f;
        }
    }
    __syncthreads();

    if (out_cl) {
        // channels-last output [B][N][levels*D*D]: warp = query, lanes = channels -> 324-byte contiguous runs
        const int nch = D * D, CT = L.levels * nch;
        for (int qi = warp; qi < nq; qi += LOOKUP_THREADS / 32) {
            const LookupGeom g = geom[qi];
            const float w00 = (1.f - g.fx) * (1.f - g.fy), w01 = g.fx * (1.f - g.fy);
            const float w10 = (1.f - g.fx) * g.fy,         w11 = g.fx * g.fy;
            const float* Sq = S + qi * FS;
            float* o = out + ((int64_t)b * N + q0 + qi) * CT + (int64_t)l * nch;
            for (int ch = lane; ch < nch; ch += 32) {
                const int a = ch / D, bb = ch - a * D;
                const float* p = Sq + bb * F + a;
                o[ch] = w00 * p[0] + w01 * p[1] + w10 * p[F] + w11 * p[F + 1];
            }
        }
        return;
    }
    if (lane < nq) {
        const LookupGeom g = geom[lane];
        const float w00 = (1.f - g.fx) * (1.f - g.fy), w01 = g.fx * (1.f - g.fy);
        const float w10 = (1.f - g.fx) * g.fy,         w11 = g.fx * g.fy;
        const float* Sq = S + lane * FS;
        const int nch = D * D;
        float* o = out + ((int64_t)b * (L.levels * nch) + (int64_t)l * nch) * N + q0 + lane;
        for (int ch = warp; ch < nch; ch += LOOKUP_THREADS / 32) {
            const int a = ch / D, bb = ch - a * D;    // a shifts x, bb shifts y (corr.py:37-43)
            const float* p = Sq + bb * F + a;
            const float v = w00 * p[0] + w01 * p[1] + w10 * p[F] + w11 * p[F + 1];
            o[(int64_t)ch * N] = v;
        }
    }
}

// grid: (ceil(N/QB) * B, levels).  Every (q, l, cell) address is touched by exactly one thread of
// one CTA per launch, so the accumulation is race-free within a launch; RED (no return) is used
// because successive lookups of one forward pass accumulate into the same buffer in stream order.
template <int RT>
__global__ void __launch_bounds__(LOOKUP_THREADS)
corr_lookup_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ coords,
                       float* __restrict__ gpyr, PyramidLayout L, int B, int H, int W, int r_rt, int out_cl) {
    extern __shared__ float smem[];
    const int r = RT > 0 ? RT : r_rt;
    const int D = 2 * r + 1, F = D + 1, FP = F * F, nch = D * D;
    const int GS = nch | 1;                      // odd stride
    float*      G    = smem;                     // [QB][GS]  gout of this (group, level)
    LookupGeom* geom = reinterpret_cast<LookupGeom*>(smem + QB * GS);

    const int N = H * W;
    const int groups = ceil_div(N, QB);
    const int b  = blockIdx.x / groups;
    const int q0 = (blockIdx.x % groups) * QB;
    const int l  = blockIdx.y;
    const int Hl = L.h[l], Wl = L.w[l];
    const int nq = min(QB, N - q0);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < QB) {
        const int q = min(q0 + (int)threadIdx.x, N - 1);
        const float cx = coords[((int64_t)b * 2 + 0) * N + q];
        const float cy = coords[((int64_t)b * 2 + 1) * N + q];
        geom[threadIdx.x] = lookup_geom(cx, cy, l, r);
    }
    if (out_cl) {                                  // grad_out is [B][N][levels*D*D]: warp = query, lanes = channels
        const int CT = L.levels * nch;
        for (int qi = warp; qi < nq; qi += LOOKUP_THREADS / 32) {
            const float* gi = gout + ((int64_t)b * N + q0 + qi) * CT + (int64_t)l * nch;
            for (int ch = lane; ch < nch; ch += 32) G[qi * GS + ch] = __ldg(gi + ch);
        }
    } else if (lane < nq) {
        const float* gi = gout + ((int64_t)b * (L.levels * nch) + (int64_t)l * nch) * N + q0 + lane;
        for (int ch = warp; ch < nch; ch += LOOKUP_THREADS / 32)
            G[lane * GS + ch] = __ldg(gi + (int64_t)ch * N);
    }
    __syncthreads();

    // scatter: one warp per query, lanes = footprint cells (adjoint gather of <= 4 taps per cell)
    float* lvl = gpyr + L.off[l];
    const int64_t plane = (int64_t)Hl * Wl;
    for (int qi = warp; qi < nq; qi += LOOKUP_THREADS / 32) {
        const LookupGeom g = geom[qi];
        float* dst = lvl + ((int64_t)b * N + q0 + qi) * plane;
        const float* Gq = G + qi * GS;
        const float w00 = (1.f - g.fx) * (1.f - g.fy), w01 = g.fx * (1.f - g.fy);
        const float w10 = (1.f - g.fx) * g.fy, w11 = g.fx * g.fy;
        for (int cell = lane; cell < FP; cell += 32) {
            const int rr = cell / F, cc = cell - rr * F;
            const int y = g.iy0 + rr, x = g.ix0 + cc;
            if ((unsigned)y >= (unsigned)Hl || (unsigned)x >= (unsigned)Wl) continue;
            // cell (rr,cc) receives tap (a=cc,b=rr)*w00 + (cc-1,rr)*w01 + (cc,rr-1)*w10 + (cc-1,rr-1)*w11
            const bool a0 = cc < D, a1 = cc > 0, b0 = rr < D, b1 = rr > 0;
            float acc = 0.f;
            if (a0 && b0) acc = w00 * Gq[cc * D + rr];
            if (a1 && b0) acc = fmaf(w01, Gq[(cc - 1) * D + rr], acc);
            if (a0 && b1) acc = fmaf(w10, Gq[cc * D + rr - 1], acc);
            if (a1 && b1) acc = fmaf(w11, Gq[(cc - 1) * D + rr - 1], acc);
            red_add(dst + y * Wl + x, acc);
        }
    }
}

}  // namespace pcfa

using namespace pcfa;

static int lookup_check(const void* a, const void* b, const void* c, int B, int H, int W,
                        int levels, int radius) {
    if (!a || !b || !c || B <= 0 || H <= 0 || W <= 0 || levels <= 0 || levels > 8 || radius < 0 ||
        radius > 15)
        return PCFA_E_BADARG;
    if ((int64_t)B * ceil_div(H * W, QB) > 0x7fffffffLL) return PCFA_E_TOOLARGE;
    return PCFA_OK;
}

static int lookup_forward(const float* pyramid, const float* coords, float* out, int B, int H, int W, int num_levels,
                          int radius, int out_cl, pcfa_stream_t stream) {
    PCFA_TRY(lookup_check(pyramid, coords, out, B, H, W, num_levels, radius));
    const PyramidLayout L = make_pyramid_layout(B, H, W, num_levels);
    const int D = 2 * radius + 1, FP = (D + 1) * (D + 1);
    const size_t smem = (size_t)QB * (FP | 1) * sizeof(float) + QB * sizeof(LookupGeom);
    dim3 grid(B * ceil_div(H * W, QB), num_levels);
    cudaStream_t s = as_stream(stream);
    if (radius == 4)      corr_lookup_fwd_kernel<4><<<grid, LOOKUP_THREADS, smem, s>>>(pyramid, coords, out, L, B, H, W, radius, out_cl);
    else if (radius == 3) corr_lookup_fwd_kernel<3><<<grid, LOOKUP_THREADS, smem, s>>>(pyramid, coords, out, L, B, H, W, radius, out_cl);
    else {
        if (smem > 48 * 1024)
            PCFA_CUDA_TRY(cudaFuncSetAttribute(corr_lookup_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        corr_lookup_fwd_kernel<0><<<grid, LOOKUP_THREADS, smem, s>>>(pyramid, coords, out, L, B, H, W, radius, out_cl);
    }
    return after_launch();
}

static int lookup_backward(const float* grad_out, const float* coords, float* grad_pyramid, int B, int H, int W,
                           int num_levels, int radius, int out_cl, pcfa_stream_t stream) {
    PCFA_TRY(lookup_check(grad_out, coords, grad_pyramid, B, H, W, num_levels, radius));
    const PyramidLayout L = make_pyramid_layout(B, H, W, num_levels);
    const int D = 2 * radius + 1;
    const size_t smem = (size_t)QB * ((D * D) | 1) * sizeof(float) + QB * sizeof(LookupGeom);
    dim3 grid(B * ceil_div(H * W, QB), num_levels);
    cudaStream_t s = as_stream(stream);
    if (radius == 4)      corr_lookup_bwd_kernel<4><<<grid, LOOKUP_THREADS, smem, s>>>(grad_out, coords, grad_pyramid, L, B, H, W, radius, out_cl);
    else if (radius == 3) corr_lookup_bwd_kernel<3><<<grid, LOOKUP_THREADS, smem, s>>>(grad_out, coords, grad_pyramid, L, B, H, W, radius, out_cl);
    else {
        if (smem > 48 * 1024)
            PCFA_CUDA_TRY(cudaFuncSetAttribute(corr_lookup_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        corr_lookup_bwd_kernel<0><<<grid, LOOKUP_THREADS, smem, s>>>(grad_out, coords, grad_pyramid, L, B, H, W, radius, out_cl);
    }
    return after_launch();
}

extern "C" int pcfa_corr_lookup_forward(const float* pyramid, const float* coords, float* out, int B, int H, int W,
                                        int num_levels, int radius, pcfa_stream_t stream) {
    return lookup_forward(pyramid, coords, out, B, H, W, num_levels, radius, 0, stream);
}
extern "C" int pcfa_corr_lookup_backward(const float* grad_out, const float* coords, float* grad_pyramid, int B, int H,
                                         int W, int num_levels, int radius, pcfa_stream_t stream) {
    return lookup_backward(grad_out, coords, grad_pyramid, B, H, W, num_levels, radius, 0, stream);
}
// channels-last variants: out / grad_out are [B][H][W][levels*(2r+1)^2] in memory (torch.channels_last of [B,C,H,W])
extern "C" int pcfa_corr_lookup_forward_cl(const float* pyramid, const float* coords, float* out, int B, int H, int W,
                                           int num_levels, int radius, pcfa_stream_t stream) {
    return lookup_forward(pyramid, coords, out, B, H, W, num_levels, radius, 1, stream);
}
extern "C" int pcfa_corr_lookup_backward_cl(const float* grad_out, const float* coords, float* grad_pyramid, int B, int H,
                                            int W, int num_levels, int radius, pcfa_stream_t stream) {
    return lookup_backward(grad_out, coords, grad_pyramid, B, H, W, num_levels, radius, 1, stream);
}
