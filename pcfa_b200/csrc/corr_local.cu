// Local-window correlation (cost volume over a displacement patch) — forward and backward.
// One kernel family serves both reference operators:
//   * spatial_correlation_sample (Correlation_Module/correlation.cpp:9-178)  — PWCNet 9x9 patch
//   * FlowNet2 correlation_cuda (correlation_cuda_kernel.cu:73-334)          — 21x21, stride2 = 2
// Channel order is dy-major in both (ph*patchW + pw ; tc = (tj+R)*D + (ti+R)).
//
// Fast path (kH = kW = 1, the only kernel size either network uses): each thread owns one output
// pixel and a whole row of patchW displacements in registers, so every in1 value is loaded once per
// patch row; loads are coalesced along w.  The backward stages the gout tile of one output row in
// shared memory and walks channels, turning both input gradients into gathers (no atomics).
// A fully generic (any kernel/stride/dilation) gather kernel covers the remaining parameter space.
#include "common.cuh"

namespace pcfa {

struct LocalCorrGeom {
    int B, C, iH, iW, oH, oW;
    int kH, kW, patchH, patchW, padH, padW, dilH, dilW, dpH, dpW, dH, dW;
    int radH, radW;      // (patch-1)/2, CPU form (correlation.cpp:88-89)
    float scale;
};

// ------------------------------------------------------------------ generic forward (any params)
__global__ void local_corr_fwd_generic(const float* __restrict__ in1, const float* __restrict__ in2,
                                       float* __restrict__ out, LocalCorrGeom g) {
    const int64_t total = (int64_t)g.B * g.patchH * g.patchW * g.oH * g.oW;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = idx;
        const int w = (int)(t % g.oW); t /= g.oW;
        const int h = (int)(t % g.oH); t /= g.oH;
        const int pw = (int)(t % g.patchW); t /= g.patchW;
        const int ph = (int)(t % g.patchH);
        const int n = (int)(t / g.patchH);
        const int u = -g.padH + h * g.dH, v = -g.padW + w * g.dW;
        const int su = (ph - g.radH) * g.dpH, sv = (pw - g.radW) * g.dpW;
        const float* a = in1 + (int64_t)n * g.C * g.iH * g.iW;
        const float* b = in2 + (int64_t)n * g.C * g.iH * g.iW;
        float acc = 0.f;
        for (int c = 0; c < g.C; ++c) {
            for (int i = 0; i < g.kH; ++i) {
                const int i1 = u + i * g.dilH, i2 = i1 + su;
                if (i1 < 0 || i1 >= g.iH || i2 < 0 || i2 >= g.iH) continue;
                for (int j = 0; j < g.kW; ++j) {
                    const int j1 = v + j * g.dilW, j2 = j1 + sv;
                    if (j1 < 0 || j1 >= g.iW || j2 < 0 || j2 >= g.iW) continue;
                    acc = fmaf(__ldg(a + ((int64_t)c * g.iH + i1) * g.iW + j1),
                               __ldg(b + ((int64_t)c * g.iH + i2) * g.iW + j2), acc);
                }
            }
        }
        out[idx] = acc * g.scale;
    }
}

// ------------------------------------------------------------------ k=1 forward, patch row in regs
// grid: (ceil(oW/32), oH, B*patchH) ; block: 32 x CS (x, channel slice) ; partial sums over the CS
// channel slices are combined through shared memory.  This kernel serves the small maps (the tiled kernel takes the
// large ones), where the serial channel loop is the critical path: hence many slices.
template <int PW, int CS>
__global__ void __launch_bounds__(32 * CS)
local_corr_fwd_k1(const float* __restrict__ in1, const float* __restrict__ in2,
                  float* __restrict__ out, LocalCorrGeom g) {
    __shared__ float red[CS][PW][33];
    const int w = blockIdx.x * 32 + threadIdx.x;
    const int h = blockIdx.y;
    const int n = blockIdx.z / g.patchH, ph = blockIdx.z % g.patchH;
    const int cs = threadIdx.y;
    const int i1 = -g.padH + h * g.dH, j1 = -g.padW + w * g.dW;
    const int i2 = i1 + (ph - g.radH) * g.dpH;
    float acc[PW];
#pragma unroll
    for (int p = 0; p < PW; ++p) acc[p] = 0.f;
    const bool ok = (w < g.oW) && i1 >= 0 && i1 < g.iH && j1 >= 0 && j1 < g.iW && i2 >= 0 && i2 < g.iH;
    if (ok) {
        const int64_t plane = (int64_t)g.iH * g.iW;
        const float* a = in1 + (int64_t)n * g.C * plane + (int64_t)i1 * g.iW + j1;
        const float* b = in2 + (int64_t)n * g.C * plane + (int64_t)i2 * g.iW;
        for (int c = cs; c < g.C; c += CS) {
            const float av = __ldg(a + c * plane);
            const float* br = b + c * plane;
#pragma unroll
            for (int p = 0; p < PW; ++p) {
                const int j2 = j1 + (p - g.radW) * g.dpW;
                if (p < g.patchW && j2 >= 0 && j2 < g.iW) acc[p] = fmaf(av, __ldg(br + j2), acc[p]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < PW; ++p) red[cs][p][threadIdx.x] = acc[p];
    __syncthreads();
    if (w < g.oW) {
        for (int p = cs; p < g.patchW; p += CS) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < CS; ++k) v += red[k][p][threadIdx.x];
            out[((((int64_t)n * g.patchH + ph) * g.patchW + p) * g.oH + h) * g.oW + w] = v * g.scale;
        }
    }
}

// ------------------------------------------------------------------ generic backward (gather form)
// WHICH = 1: grad wrt in1, WHICH = 2: grad wrt in2.  One thread per input element.
template <int WHICH>
__global__ void local_corr_bwd_generic(const float* __restrict__ other, const float* __restrict__ gout,
                                       float* __restrict__ gin, LocalCorrGeom g) {
    const int64_t total = (int64_t)g.B * g.C * g.iH * g.iW;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = idx;
        const int x = (int)(t % g.iW); t /= g.iW;
        const int y = (int)(t % g.iH); t /= g.iH;
        const int c = (int)(t % g.C);
        const int n = (int)(t / g.C);
        const float* o = other + ((int64_t)n * g.C + c) * g.iH * g.iW;
        float acc = 0.f;
        for (int ph = 0; ph < g.patchH; ++ph) {
            const int su = (ph - g.radH) * g.dpH;
            for (int pw = 0; pw < g.patchW; ++pw) {
                const int sv = (pw - g.radW) * g.dpW;
                // position in in1 / in2 of the product this element takes part in
                const int y1 = (WHICH == 1) ? y : y - su, x1 = (WHICH == 1) ? x : x - sv;
                const int y2 = y1 + su, x2 = x1 + sv;
                if (y1 < 0 || y1 >= g.iH || x1 < 0 || x1 >= g.iW || y2 < 0 || y2 >= g.iH || x2 < 0 ||
                    x2 >= g.iW)
                    continue;
                const float ov = (WHICH == 1) ? __ldg(o + (int64_t)y2 * g.iW + x2)
                                              : __ldg(o + (int64_t)y1 * g.iW + x1);
                const float* go = gout + (((int64_t)n * g.patchH + ph) * g.patchW + pw) * g.oH * g.oW;
                float gs = 0.f;
                for (int i = 0; i < g.kH; ++i) {
                    const int hh = y1 + g.padH - i * g.dilH;
                    if (hh < 0 || hh % g.dH) continue;
                    const int h = hh / g.dH;
                    if (h >= g.oH) continue;
                    for (int j = 0; j < g.kW; ++j) {
                        const int ww = x1 + g.padW - j * g.dilW;
                        if (ww < 0 || ww % g.dW) continue;
                        const int w = ww / g.dW;
                        if (w >= g.oW) continue;
                        gs += __ldg(go + (int64_t)h * g.oW + w);
                    }
                }
                acc = fmaf(gs, ov, acc);
            }
        }
        gin[idx] = acc * g.scale;
    }
}

// ------------------------------------------------------------------ k=1, stride-1 backward
// grid: (ceil(iW/32), iH, B) ; block 32 x 8.  Shared: gout[patchH*patchW][32] for this input row,
// already shifted for WHICH==2 so that both variants are "sum_p Gs[p][x] * other[c, y+-su, x+-sv]".
template <int WHICH>
__global__ void __launch_bounds__(256)
local_corr_bwd_k1s1(const float* __restrict__ other, const float* __restrict__ gout,
                    float* __restrict__ gin, LocalCorrGeom g) {
    extern __shared__ float Gs[];     // [P][32]
    const int P = g.patchH * g.patchW;
    const int x0 = blockIdx.x * 32, y = blockIdx.y, n = blockIdx.z;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int e = tid; e < P * 32; e += 256) {
        const int p = e >> 5, xi = e & 31;
        const int ph = p / g.patchW, pw = p - ph * g.patchW;
        const int su = (ph - g.radH) * g.dpH, sv = (pw - g.radW) * g.dpW;
        const int x = x0 + xi;
        const int y1 = (WHICH == 1) ? y : y - su, x1 = (WHICH == 1) ? x : x - sv;
        const int y2 = y1 + su, x2 = x1 + sv;
        const int h = y1 + g.padH, w = x1 + g.padW;     // dH = dW = 1, k = 1
        float v = 0.f;
        if (x < g.iW && y1 >= 0 && y1 < g.iH && x1 >= 0 && x1 < g.iW && y2 >= 0 && y2 < g.iH &&
            x2 >= 0 && x2 < g.iW && h >= 0 && h < g.oH && w >= 0 && w < g.oW)
            v = __ldg(gout + ((((int64_t)n * P + p) * g.oH) + h) * g.oW + w);
        Gs[e] = v;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= g.iW) return;
    const int64_t plane = (int64_t)g.iH * g.iW;
    for (int c = threadIdx.y; c < g.C; c += 8) {
        const float* o = other + ((int64_t)n * g.C + c) * plane;
        float acc = 0.f;
        for (int ph = 0; ph < g.patchH; ++ph) {
            const int su = (ph - g.radH) * g.dpH;
            const int yo = (WHICH == 1) ? y + su : y - su;
            if (yo < 0 || yo >= g.iH) continue;
            const float* orow = o + (int64_t)yo * g.iW;
            const float* grow = Gs + (ph * g.patchW) * 32 + threadIdx.x;
            for (int pw = 0; pw < g.patchW; ++pw) {
                const int sv = (pw - g.radW) * g.dpW;
                const int xo = (WHICH == 1) ? x + sv : x - sv;
                if (xo >= 0 && xo < g.iW) acc = fmaf(grow[pw * 32], __ldg(orow + xo), acc);
            }
        }
        gin[((int64_t)n * g.C + c) * plane + (int64_t)y * g.iW + x] = acc * g.scale;
    }
}

// opaque 16-byte shared-memory load: keeps register-resident windows from being re-materialised as scalar LDS
__device__ __forceinline__ float4 lds128(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

// ------------------------------------------------------------------ tiled fast paths (k = 1, stride 1)
// Forward: CTA = (4-row x 32-col output tile, chunk of DYC displacement rows); channel chunks of CC are staged in
// shared memory (in1 tile + the in2 halo tile), each thread owns 4 adjacent pixels x all D horizontal
// displacements of one displacement row: one LDS.128 of in1 and NB4 LDS.128 of in2 feed 4*D FMAs per channel.
template <int D, int DIL, int DYC, int TH, int CC>
__global__ void __launch_bounds__(8 * DYC * TH)
lc_fwd_tiled(const float* __restrict__ in1, const float* __restrict__ in2, float* __restrict__ out, LocalCorrGeom g) {
    constexpr int TW = 32, RAD = (D - 1) / 2, R = RAD * DIL, BW = TW + 2 * R, BH = TH + (DYC - 1) * DIL;
    constexpr int WIN = 4 + (D - 1) * DIL, NB4 = WIN / 4, NT = 8 * DYC * TH;
    static_assert(WIN % 4 == 0 && R % 4 == 0 && BW % 4 == 0, "window must be float4-tileable");
    extern __shared__ __align__(16) float lc_smem[];
    float* As = lc_smem;                       // [CC][TH][TW]
    float* Bs = lc_smem + CC * TH * TW;        // [CC][BH][BW]
    const int tiles_x = ceil_div(g.oW, TW);
    const int tx0 = (blockIdx.x % tiles_x) * TW, ty0 = (blockIdx.x / tiles_x) * TH;
    constexpr int NDYC = (D + DYC - 1) / DYC;
    const int n = blockIdx.y / NDYC, dy0 = (blockIdx.y % NDYC) * DYC;
    const int tid = threadIdx.x;
    const int gx = tid & 7, tdy = (tid >> 3) % DYC, ty = tid / (8 * DYC);
    const int64_t plane = (int64_t)g.iH * g.iW;
    const float* a = in1 + (int64_t)n * g.C * plane;
    const float* b = in2 + (int64_t)n * g.C * plane;
    const int i2base = ty0 - g.padH + (dy0 - RAD) * DIL, j2base = tx0 - g.padW - R;
    const bool vec = (g.iW % 4 == 0) && (g.padW % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) % 16 == 0);

    float acc[4][D];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int d = 0; d < D; ++d) acc[p][d] = 0.f;

    for (int c0 = 0; c0 < g.C; c0 += CC) {
        if (vec) {      // rows are 16-byte tileable: float4 loads, a chunk is entirely inside or outside the image
            for (int e = tid; e < CC * TH * (TW / 4); e += NT) {
                const int c = e / (TH * (TW / 4)), r = (e / (TW / 4)) % TH, x = (e % (TW / 4)) * 4;
                const int i1 = ty0 + r - g.padH, j1 = tx0 + x - g.padW;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + c < g.C && (unsigned)i1 < (unsigned)g.iH && (unsigned)j1 < (unsigned)g.iW)
                    v = __ldg(reinterpret_cast<const float4*>(a + (c0 + c) * plane + (int64_t)i1 * g.iW + j1));
                *reinterpret_cast<float4*>(As + (c * TH + r) * TW + x) = v;
            }
            for (int e = tid; e < CC * BH * (BW / 4); e += NT) {
                const int c = e / (BH * (BW / 4)), r = (e / (BW / 4)) % BH, x = (e % (BW / 4)) * 4;
                const int i2 = i2base + r, j2 = j2base + x;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + c < g.C && (unsigned)i2 < (unsigned)g.iH && (unsigned)j2 < (unsigned)g.iW)
                    v = __ldg(reinterpret_cast<const float4*>(b + (c0 + c) * plane + (int64_t)i2 * g.iW + j2));
                *reinterpret_cast<float4*>(Bs + (c * BH + r) * BW + x) = v;
            }
        } else {
        for (int e = tid; e < CC * TH * TW; e += NT) {
            const int c = e / (TH * TW), r = (e / TW) % TH, x = e % TW;
            const int i1 = ty0 + r - g.padH, j1 = tx0 + x - g.padW;
            float v = 0.f;
            if (c0 + c < g.C && (unsigned)i1 < (unsigned)g.iH && (unsigned)j1 < (unsigned)g.iW)
                v = __ldg(a + (c0 + c) * plane + (int64_t)i1 * g.iW + j1);
            As[e] = v;
        }
        for (int e = tid; e < CC * BH * BW; e += NT) {
            const int c = e / (BH * BW), r = (e / BW) % BH, x = e % BW;
            const int i2 = i2base + r, j2 = j2base + x;
            float v = 0.f;
            if (c0 + c < g.C && (unsigned)i2 < (unsigned)g.iH && (unsigned)j2 < (unsigned)g.iW)
                v = __ldg(b + (c0 + c) * plane + (int64_t)i2 * g.iW + j2);
            Bs[e] = v;
        }
        }
        __syncthreads();
        if (dy0 + tdy < D) {
#pragma unroll 2
            for (int c = 0; c < CC; ++c) {
                const float4 a4 = *reinterpret_cast<const float4*>(As + (c * TH + ty) * TW + gx * 4);
                const float av[4] = {a4.x, a4.y, a4.z, a4.w};
                const float* brow = Bs + (c * BH + ty + tdy * DIL) * BW + gx * 4;
                float bv[WIN];
#pragma unroll
                for (int k = 0; k < NB4; ++k) {
                    const float4 t = lds128(brow + 4 * k);
                    bv[4 * k] = t.x; bv[4 * k + 1] = t.y; bv[4 * k + 2] = t.z; bv[4 * k + 3] = t.w;
                }
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int d = 0; d < D; ++d) acc[p][d] = fmaf(av[p], bv[p + d * DIL], acc[p][d]);
            }
        }
        __syncthreads();
    }
    const int dyi = dy0 + tdy, h = ty0 + ty, w = tx0 + gx * 4;
    if (dyi < D && h < g.oH && w < g.oW) {
        float* o = out + ((((int64_t)n * D + dyi) * D) * g.oH + h) * g.oW + w;
        const int64_t dstride = (int64_t)g.oH * g.oW;
        const bool vst = (w + 3 < g.oW) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0) && ((dstride & 3) == 0);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            if (vst) {
                *reinterpret_cast<float4*>(o + d * dstride) =
                    make_float4(acc[0][d] * g.scale, acc[1][d] * g.scale, acc[2][d] * g.scale, acc[3][d] * g.scale);
            } else {
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    if (w + p < g.oW) o[d * dstride + p] = acc[p][d] * g.scale;
            }
        }
    }
}

// Backward: CTA = (one input row y, 32 columns, 32 channels); for each displacement row the matching row of
// the other input (with halo) and the D gout rows (pre-shifted for WHICH == 2) are staged; each thread owns
// 4 adjacent pixels of one channel, keeps the 4+(D-1)*DIL halo values in registers and does 4*D FMAs per
// displacement row against one LDS.128 of gout per displacement.  Both gradients are gathers: no atomics.
template <int D, int DIL, int WHICH, int CPT>
__global__ void __launch_bounds__(256)
lc_bwd_rows(const float* __restrict__ other, const float* __restrict__ gout, float* __restrict__ gin, LocalCorrGeom g) {
    // CPT = channels per thread (2 amortises every gout LDS.128 over 8 FMAs; used when C >= 64)
    constexpr int TW = 32, CHB = 32 * CPT, RAD = (D - 1) / 2, R = RAD * DIL, BW = TW + 2 * R;
    constexpr int WIN = 4 + (D - 1) * DIL, NB4 = WIN / 4;
    __shared__ __align__(16) float Os[CHB][BW];
    __shared__ __align__(16) float Gs[D][TW];
    const int tiles_x = ceil_div(g.iW, TW);
    const int x0 = (blockIdx.x % tiles_x) * TW, y = blockIdx.x / tiles_x;
    const int cblocks = ceil_div(g.C, CHB);
    const int n = blockIdx.y / cblocks, c0 = (blockIdx.y % cblocks) * CHB;
    const int tid = threadIdx.x, gx = tid & 7, cl = tid >> 3;
    const int64_t plane = (int64_t)g.iH * g.iW, oplane = (int64_t)g.oH * g.oW;
    const float* o = other + ((int64_t)n * g.C + c0) * plane;
    const float* go = gout + (int64_t)n * D * D * oplane;
    float acc[CPT][4];
#pragma unroll
    for (int k = 0; k < CPT; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
    const bool vec = (g.iW % 4 == 0) && (reinterpret_cast<uintptr_t>(o) % 16 == 0);

    for (int dyi = 0; dyi < D; ++dyi) {
        const int su = (dyi - RAD) * DIL;
        const int yo = (WHICH == 1) ? y + su : y - su;              // row of the other input
        const int h = ((WHICH == 1) ? y : y - su) + g.padH;          // gout row (out pixel of the in1 pixel)
        __syncthreads();
        if (vec) {
            for (int e = tid; e < CHB * (BW / 4); e += 256) {
                const int c = e / (BW / 4), q = (e % (BW / 4)) * 4;
                const int xo = x0 - R + q;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + c < g.C && (unsigned)yo < (unsigned)g.iH && (unsigned)xo < (unsigned)g.iW)
                    v = __ldg(reinterpret_cast<const float4*>(o + c * plane + (int64_t)yo * g.iW + xo));
                *reinterpret_cast<float4*>(&Os[c][q]) = v;
            }
        } else {
            for (int e = tid; e < CHB * BW; e += 256) {
                const int c = e / BW, q = e % BW;
                const int xo = x0 - R + q;
                float v = 0.f;
                if (c0 + c < g.C && (unsigned)yo < (unsigned)g.iH && (unsigned)xo < (unsigned)g.iW)
                    v = __ldg(o + c * plane + (int64_t)yo * g.iW + xo);
                Os[c][q] = v;
            }
        }
        for (int e = tid; e < D * TW; e += 256) {
            const int dxi = e / TW, xi = e % TW;
            const int sv = (dxi - RAD) * DIL;
            const int w = ((WHICH == 1) ? x0 + xi : x0 + xi - sv) + g.padW;
            float v = 0.f;
            if ((unsigned)h < (unsigned)g.oH && (unsigned)w < (unsigned)g.oW)
                v = __ldg(go + ((int64_t)dyi * D + dxi) * oplane + (int64_t)h * g.oW + w);
            Gs[dxi][xi] = v;
        }
        __syncthreads();
        float bv[CPT][WIN];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const float* orow = &Os[cl + 32 * k][gx * 4];
#pragma unroll
            for (int q = 0; q < NB4; ++q) {
                const float4 t = lds128(orow + 4 * q);
                bv[k][4 * q] = t.x; bv[k][4 * q + 1] = t.y; bv[k][4 * q + 2] = t.z; bv[k][4 * q + 3] = t.w;
            }
        }
#pragma unroll
        for (int dxi = 0; dxi < D; ++dxi) {
            const float4 g4 = lds128(&Gs[dxi][gx * 4]);
            const int sh = (WHICH == 1) ? dxi * DIL : (D - 1 - dxi) * DIL;
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                acc[k][0] = fmaf(g4.x, bv[k][sh + 0], acc[k][0]);
                acc[k][1] = fmaf(g4.y, bv[k][sh + 1], acc[k][1]);
                acc[k][2] = fmaf(g4.z, bv[k][sh + 2], acc[k][2]);
                acc[k][3] = fmaf(g4.w, bv[k][sh + 3], acc[k][3]);
            }
        }
    }
    const int x = x0 + gx * 4;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = c0 + cl + 32 * k;
        if (c >= g.C || x >= g.iW) continue;
        float* dst = gin + ((int64_t)n * g.C + c) * plane + (int64_t)y * g.iW + x;
        if (x + 3 < g.iW && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            *reinterpret_cast<float4*>(dst) = make_float4(acc[k][0] * g.scale, acc[k][1] * g.scale, acc[k][2] * g.scale, acc[k][3] * g.scale);
        } else {
#pragma unroll
            for (int p = 0; p < 4; ++p)
                if (x + p < g.iW) dst[p] = acc[k][p] * g.scale;
        }
    }
}

template <int D, int DIL, int DYC, int TH, int CC>
static int launch_lc_fwd(const float* in1, const float* in2, float* out, const LocalCorrGeom& g, cudaStream_t s) {
    constexpr int R = (D - 1) / 2 * DIL, BW = 32 + 2 * R, BH = TH + (DYC - 1) * DIL;
    constexpr int smem = (CC * TH * 32 + CC * BH * BW) * (int)sizeof(float);
    static bool set = false;
    if (!set && smem > 48 * 1024) {
        PCFA_CUDA_TRY(cudaFuncSetAttribute(lc_fwd_tiled<D, DIL, DYC, TH, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        set = true;
    }
    constexpr int NDYC = (D + DYC - 1) / DYC;
    dim3 grid(ceil_div(g.oW, 32) * ceil_div(g.oH, TH), g.B * NDYC);
    lc_fwd_tiled<D, DIL, DYC, TH, CC><<<grid, 8 * DYC * TH, smem, s>>>(in1, in2, out, g);
    return after_launch();
}

template <int D, int DIL, int CPT>
static int launch_lc_bwd_cpt(const float* in1, const float* in2, const float* gout, float* g1, float* g2,
                             const LocalCorrGeom& g, cudaStream_t s) {
    dim3 grid(ceil_div(g.iW, 32) * g.iH, g.B * ceil_div(g.C, 32 * CPT));
    lc_bwd_rows<D, DIL, 1, CPT><<<grid, 256, 0, s>>>(in2, gout, g1, g);
    PCFA_TRY(after_launch());
    lc_bwd_rows<D, DIL, 2, CPT><<<grid, 256, 0, s>>>(in1, gout, g2, g);
    return after_launch();
}

template <int D, int DIL>
static int launch_lc_bwd(const float* in1, const float* in2, const float* gout, float* g1, float* g2,
                         const LocalCorrGeom& g, cudaStream_t s) {
    if (g.C >= 64 && g.C % 64 == 0) return launch_lc_bwd_cpt<D, DIL, 2>(in1, in2, gout, g1, g2, g, s);
    return launch_lc_bwd_cpt<D, DIL, 1>(in1, in2, gout, g1, g2, g, s);
}

// ------------------------------------------------------------------ FlowNet2 backward, literal form
// Follows correlation_cuda_kernel.cu:150-334 including the truncating integer divisions; used when
// stride1 != 1 or kernel_size != 1 (for stride1 == 1, k == 1 it equals the fast path above).
template <int WHICH>
__global__ void fn2corr_bwd_literal(const float* __restrict__ other, const float* __restrict__ gout,
                                    float* __restrict__ gin, int B, int C, int H, int W, int outC,
                                    int oH, int oW, int pad, int ks, int md, int s1, int s2) {
    const int64_t total = (int64_t)B * C * H * W;
    const int kr = (ks - 1) / 2, R = md / s2, D = 2 * R + 1;
    const float nelems = (float)(ks * ks * C);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = idx;
        const int xw = (int)(t % W); t /= W;
        const int yh = (int)(t % H); t /= H;
        const int c = (int)(t % C);
        const int n = (int)(t / C);
        // the reference launches one block per (h, w) with y = h*stride1 + pad — it only ever
        // writes the elements reachable that way (kernel.cu:165-166); with stride1 == 1 all of them
        if (yh % s1 || xw % s1) { gin[idx] = 0.f; continue; }   // never written by the reference
        const int y = yh + pad, x = xw + pad;    // padded coordinates
        float acc = 0.f;
        const float* o = other + ((int64_t)n * C + c) * H * W;
        for (int tc = 0; tc < outC; ++tc) {
            const int i2 = (tc % D - R) * s2, j2 = (tc / D - R) * s2;
            int xmin, ymin, xmax, ymax, yo, xo;
            if (WHICH == 1) {
                xmin = (x - kr - md) / s1; ymin = (y - kr - md) / s1;
                xmax = (x + kr - md) / s1; ymax = (y + kr - md) / s1;
                yo = y + j2; xo = x + i2;
            } else {
                xmin = (x - kr - md - i2) / s1; ymin = (y - kr - md - j2) / s1;
                xmax = (x + kr - md - i2) / s1; ymax = (y + kr - md - j2) / s1;
                yo = y - j2; xo = x - i2;
            }
            if (xmax < 0 || ymax < 0 || xmin >= oW || ymin >= oH) continue;
            if (xmin > xmax || ymin > ymax) continue;
            xmin = max(0, xmin); xmax = min(oW - 1, xmax);
            ymin = max(0, ymin); ymax = min(oH - 1, ymax);
            // value of the zero-padded other input at padded (yo, xo)
            const int yu = yo - pad, xu = xo - pad;
            if (yu < 0 || yu >= H || xu < 0 || xu >= W) continue;
            const float ov = __ldg(o + (int64_t)yu * W + xu);
            const float* go = gout + ((int64_t)n * outC + tc) * oH * oW;
            float gs = 0.f;
            for (int j = ymin; j <= ymax; ++j)
                for (int i = xmin; i <= xmax; ++i) gs += __ldg(go + (int64_t)j * oW + i);
            acc = fmaf(gs, ov, acc);
        }
        gin[idx] = acc / nelems;
    }
}

static int grid1d(int64_t total, int threads) {
    int64_t b = ceil_div<int64_t>(total, threads);
    const int64_t cap = (int64_t)kNumSMs * 32;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace pcfa

using namespace pcfa;

static int scs_geom(LocalCorrGeom& g, int B, int C, int iH, int iW, const pcfa_scs_params* p,
                    float scale) {
    if (!p || B <= 0 || C <= 0 || iH <= 0 || iW <= 0) return PCFA_E_BADARG;
    if (p->kH <= 0 || p->kW <= 0 || p->patchH <= 0 || p->patchW <= 0 || p->padH < 0 || p->padW < 0 ||
        p->dilH <= 0 || p->dilW <= 0 || p->dilPatchH <= 0 || p->dilPatchW <= 0 || p->dH <= 0 ||
        p->dW <= 0)
        return PCFA_E_BADARG;
    const int dkH = (p->kH - 1) * p->dilH + 1, dkW = (p->kW - 1) * p->dilW + 1;
    if (iH + 2 * p->padH < dkH || iW + 2 * p->padW < dkW) return PCFA_E_BADARG;
    g.B = B; g.C = C; g.iH = iH; g.iW = iW;
    g.oH = (iH + 2 * p->padH - dkH) / p->dH + 1;     // correlation.cpp:93-95
    g.oW = (iW + 2 * p->padW - dkW) / p->dW + 1;
    g.kH = p->kH; g.kW = p->kW; g.patchH = p->patchH; g.patchW = p->patchW;
    g.padH = p->padH; g.padW = p->padW; g.dilH = p->dilH; g.dilW = p->dilW;
    g.dpH = p->dilPatchH; g.dpW = p->dilPatchW; g.dH = p->dH; g.dW = p->dW;
    g.radH = (p->patchH - 1) / 2; g.radW = (p->patchW - 1) / 2;
    g.scale = scale;
    if ((int64_t)B * p->patchH > 65535 || g.oH > 65535) return PCFA_E_TOOLARGE;
    return PCFA_OK;
}

static bool tiled_ok(const LocalCorrGeom& g, int D, int DIL) {
    return g.kH == 1 && g.kW == 1 && g.dH == 1 && g.dW == 1 && g.patchH == D && g.patchW == D && g.dpH == DIL &&
           g.dpW == DIL && (int64_t)g.B * 7 <= 65535 && (int64_t)g.B * ceil_div(g.C, 32) <= 65535;
}

static int local_forward(const float* in1, const float* in2, float* out, const LocalCorrGeom& g,
                         cudaStream_t s) {
    // the tiled kernel needs enough tiles to fill the machine; small maps (PWCNet levels 6..4) stay on the
    // register-row kernel below, which parallelises over displacement rows instead
    if (tiled_ok(g, 9, 1) && (int64_t)g.oH * g.oW >= 4096) return launch_lc_fwd<9, 1, 3, 4, 16>(in1, in2, out, g, s);   // PWCNet
    if (tiled_ok(g, 21, 2) && (int64_t)g.oH * g.oW >= 2048) return launch_lc_fwd<21, 2, 3, 4, 16>(in1, in2, out, g, s); // FlowNet2
    if (g.kH == 1 && g.kW == 1 && g.patchW <= 21) {
        dim3 grid(ceil_div(g.oW, 32), g.oH, g.B * g.patchH);
        if (g.patchW <= 9) local_corr_fwd_k1<9, 16><<<grid, dim3(32, 16), 0, s>>>(in1, in2, out, g);
        else               local_corr_fwd_k1<21, 8><<<grid, dim3(32, 8), 0, s>>>(in1, in2, out, g);
        return after_launch();
    }
    const int64_t total = (int64_t)g.B * g.patchH * g.patchW * g.oH * g.oW;
    local_corr_fwd_generic<<<grid1d(total, 128), 128, 0, s>>>(in1, in2, out, g);
    return after_launch();
}

static int local_backward(const float* in1, const float* in2, const float* gout, float* g1, float* g2,
                          const LocalCorrGeom& g, cudaStream_t s) {
    if (tiled_ok(g, 9, 1)) return launch_lc_bwd<9, 1>(in1, in2, gout, g1, g2, g, s);
    if (tiled_ok(g, 21, 2)) return launch_lc_bwd<21, 2>(in1, in2, gout, g1, g2, g, s);
    const size_t smem = (size_t)g.patchH * g.patchW * 32 * sizeof(float);
    if (g.kH == 1 && g.kW == 1 && g.dH == 1 && g.dW == 1 && smem <= 200 * 1024 && g.iH <= 65535 &&
        g.B <= 65535) {
        if (smem > 48 * 1024) {
            PCFA_CUDA_TRY(cudaFuncSetAttribute(local_corr_bwd_k1s1<1>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            PCFA_CUDA_TRY(cudaFuncSetAttribute(local_corr_bwd_k1s1<2>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        dim3 grid(ceil_div(g.iW, 32), g.iH, g.B), block(32, 8);
        local_corr_bwd_k1s1<1><<<grid, block, smem, s>>>(in2, gout, g1, g);
        PCFA_TRY(after_launch());
        local_corr_bwd_k1s1<2><<<grid, block, smem, s>>>(in1, gout, g2, g);
        return after_launch();
    }
    const int64_t total = (int64_t)g.B * g.C * g.iH * g.iW;
    local_corr_bwd_generic<1><<<grid1d(total, 128), 128, 0, s>>>(in2, gout, g1, g);
    PCFA_TRY(after_launch());
    local_corr_bwd_generic<2><<<grid1d(total, 128), 128, 0, s>>>(in1, gout, g2, g);
    return after_launch();
}

extern "C" int pcfa_scs_output_size(int iH, int iW, const pcfa_scs_params* p, int* oH, int* oW) {
    LocalCorrGeom g;
    PCFA_TRY(scs_geom(g, 1, 1, iH, iW, p, 1.f));
    if (oH) *oH = g.oH;
    if (oW) *oW = g.oW;
    return PCFA_OK;
}

extern "C" int pcfa_scs_forward(const float* in1, const float* in2, float* out, int B, int C, int iH,
                                int iW, const pcfa_scs_params* p, float scale, pcfa_stream_t stream) {
    if (!in1 || !in2 || !out) return PCFA_E_BADARG;
    LocalCorrGeom g;
    PCFA_TRY(scs_geom(g, B, C, iH, iW, p, scale));
    return local_forward(in1, in2, out, g, as_stream(stream));
}

extern "C" int pcfa_scs_backward(const float* in1, const float* in2, const float* grad_out,
                                 float* grad_in1, float* grad_in2, int B, int C, int iH, int iW,
                                 const pcfa_scs_params* p, float scale, pcfa_stream_t stream) {
    if (!in1 || !in2 || !grad_out || !grad_in1 || !grad_in2) return PCFA_E_BADARG;
    LocalCorrGeom g;
    PCFA_TRY(scs_geom(g, B, C, iH, iW, p, scale));
    return local_backward(in1, in2, grad_out, grad_in1, grad_in2, g, as_stream(stream));
}

// ---- FlowNet2 parameterisation ---------------------------------------------------------------
static int fn2_sizes(int H, int W, int pad, int ks, int md, int s1, int s2, int* outC, int* oH,
                     int* oW) {
    if (H <= 0 || W <= 0 || pad < 0 || ks <= 0 || (ks & 1) == 0 || md < 0 || s1 <= 0 || s2 <= 0)
        return PCFA_E_BADARG;
    const int kr = (ks - 1) / 2, border = kr + md;
    const int ph = H + 2 * pad - 2 * border, pw = W + 2 * pad - 2 * border;
    if (ph <= 0 || pw <= 0) return PCFA_E_BADARG;
    // the reference reads the padded buffers at y1 + tj*stride2 + j without bounds checks
    // (correlation_cuda_kernel.cu:119-126); that is only defined when pad >= max_disp + kernel_rad
    if (pad < md + kr) return PCFA_E_BADARG;
    const int D = 2 * (md / s2) + 1;
    *outC = D * D;
    *oH = (ph + s1 - 1) / s1;        // ceil(float / float), correlation_cuda.cc:33-34
    *oW = (pw + s1 - 1) / s1;
    return PCFA_OK;
}

extern "C" int pcfa_fn2corr_output_size(int H, int W, int pad_size, int kernel_size, int max_disp,
                                        int stride1, int stride2, int* outC, int* oH, int* oW) {
    int c, h, w;
    PCFA_TRY(fn2_sizes(H, W, pad_size, kernel_size, max_disp, stride1, stride2, &c, &h, &w));
    if (outC) *outC = c;
    if (oH) *oH = h;
    if (oW) *oW = w;
    return PCFA_OK;
}

// Map FlowNet2's parameters onto the sampler geometry (see DESIGN.md "local-window correlation"):
// window top-left in unpadded coordinates = y*s1 + md - kr - pad  ⇒  padH = pad - md + kr.
static int fn2_geom(LocalCorrGeom& g, int B, int C, int H, int W, int pad, int ks, int md, int s1,
                    int s2) {
    int outC, oH, oW;
    PCFA_TRY(fn2_sizes(H, W, pad, ks, md, s1, s2, &outC, &oH, &oW));
    if (B <= 0 || C <= 0) return PCFA_E_BADARG;
    const int kr = (ks - 1) / 2, D = 2 * (md / s2) + 1;
    g.B = B; g.C = C; g.iH = H; g.iW = W; g.oH = oH; g.oW = oW;
    g.kH = ks; g.kW = ks; g.patchH = D; g.patchW = D;
    g.padH = pad - md + kr; g.padW = pad - md + kr;
    g.dilH = 1; g.dilW = 1; g.dpH = s2; g.dpW = s2; g.dH = s1; g.dW = s1;
    g.radH = md / s2; g.radW = md / s2;
    g.scale = 1.0f / (float)(ks * ks * C);       // kernel.cu:104,143
    if ((int64_t)B * D > 65535 || oH > 65535) return PCFA_E_TOOLARGE;
    return PCFA_OK;
}

extern "C" int pcfa_fn2corr_forward(const float* in1, const float* in2, float* out, int B, int C,
                                    int H, int W, int pad_size, int kernel_size, int max_disp,
                                    int stride1, int stride2, pcfa_stream_t stream) {
    if (!in1 || !in2 || !out) return PCFA_E_BADARG;
    LocalCorrGeom g;
    PCFA_TRY(fn2_geom(g, B, C, H, W, pad_size, kernel_size, max_disp, stride1, stride2));
    return local_forward(in1, in2, out, g, as_stream(stream));
}

extern "C" int pcfa_fn2corr_backward(const float* in1, const float* in2, const float* grad_out,
                                     float* grad_in1, float* grad_in2, int B, int C, int H, int W,
                                     int pad_size, int kernel_size, int max_disp, int stride1,
                                     int stride2, pcfa_stream_t stream) {
    if (!in1 || !in2 || !grad_out || !grad_in1 || !grad_in2) return PCFA_E_BADARG;
    LocalCorrGeom g;
    PCFA_TRY(fn2_geom(g, B, C, H, W, pad_size, kernel_size, max_disp, stride1, stride2));
    cudaStream_t s = as_stream(stream);
    if (kernel_size == 1 && stride1 == 1)
        return local_backward(in1, in2, grad_out, grad_in1, grad_in2, g, s);
    const int64_t total = (int64_t)B * C * H * W;
    fn2corr_bwd_literal<1><<<grid1d(total, 128), 128, 0, s>>>(in2, grad_out, grad_in1, B, C, H, W,
                                                               g.patchH * g.patchW, g.oH, g.oW,
                                                               pad_size, kernel_size, max_disp,
                                                               stride1, stride2);
    PCFA_TRY(after_launch());
    fn2corr_bwd_literal<2><<<grid1d(total, 128), 128, 0, s>>>(in1, grad_out, grad_in2, B, C, H, W,
                                                               g.patchH * g.patchW, g.oH, g.oW,
                                                               pad_size, kernel_size, max_disp,
                                                               stride1, stride2);
    return after_launch();
}
