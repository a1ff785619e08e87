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
#include <stdlib.h>

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


// ---------------------------------------------------------------------------------------------------------------
// Channels-last lookups, second generation (the variants the RAFT closure launches 12x forward + 12x backward).
//
// ncu of the first generation at 1x55x128 (profiles/ncu_corr_kernels_r1.txt): 5.4 M warp instructions (190 per
// (query, level) task) = 60 % issue utilisation while active, 4 dependent DRAM round trips per warp (its four queries
// are gathered one after the other) and a CTA-wide barrier between gather and output; 0.74 waves, so every CTA's
// latency chain is the kernel's duration (12-14 us for 20 MB).  Here a WARP owns a (query, level) task end to end:
//   * no CTA barrier (a warp consumes only what it staged itself, __syncwarp);
//   * LK_TPW tasks per warp with all of their loads issued before the first use (12 gathers / 9 gradient loads in
//     flight per lane) and 1174 CTAs x 8 warps = ONE wave at 8 CTAs per SM for B = 1;
//   * per-lane cell -> (row, col) maps, tap offsets and validity are task-independent and computed once; footprints
//     that lie completely inside the level (the common case) skip the per-cell bounds tests;
//   * the footprint loads / gradient REDs carry an L2 evict_last policy: successive GRU iterations look up almost the
//     same 10x10 footprints (the flow changes by a fraction of a cell), 11 MB per launch, which then survive in the
//     126 MB L2 between iterations while the update block's activations stream past with normal priority.
constexpr int LK_WARPS = 8;      // warps per CTA
constexpr int LK_TPW = 5;        // consecutive queries per warp (at one level)
constexpr int LK_MIN_CTAS = 5;   // 704 CTAs at 1x55x128 = one wave at 5 CTAs per SM (<= 51 registers)

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float ldg_policy(const float* p, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void red_add_policy(float* p, float v, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

// A warp owns LK_TPW consecutive queries at ONE level, so everything that depends on the level (width, per-lane cell
// offsets, scale) is set up once per warp; warp g of sample b handles level g % levels of query group g / levels.
struct LkWarp { int l, q0, nq, Hl, Wl, plane; float scale; int64_t lvl_off; };

template <int LV>
__device__ __forceinline__ LkWarp lk_warp(int gw, int b, int N, const PyramidLayout& L) {
    LkWarp w;
    const int levels = LV ? LV : L.levels;
    const int qg = LV == 4 ? (gw >> 2) : gw / levels;
    w.l = gw - qg * levels;
    w.q0 = qg * LK_TPW;
    w.nq = min(LK_TPW, N - w.q0);                         // <= 0: nothing to do
    w.Hl = L.h[w.l]; w.Wl = L.w[w.l];
    w.plane = w.Hl * w.Wl;
    w.scale = __int_as_float((127 - w.l) << 23);          // 2^-l exactly: centroid_lvl = coords / 2**i (corr.py:41)
    w.lvl_off = L.off[w.l] + (int64_t)(b * N + w.q0) * w.plane;
    return w;
}

// base + idx as ONE IMAD.WIDE (the compiler otherwise re-derives 64-bit sums with carry chains for every access)
template <typename T>
__device__ __forceinline__ T* ptr_add(T* base, int idx) {
    uint64_t r;
    asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(r) : "r"(idx), "l"(reinterpret_cast<uint64_t>(base)));
    return reinterpret_cast<T*>(r);
}

struct LkGeom { int o; float fx, fy; bool interior; int iy0, ix0; };

template <int R>
__device__ __forceinline__ LkGeom lk_geom(float cx, float cy, const LkWarp& w) {
    constexpr int F = 2 * R + 2;
    const float sx = cx * w.scale, sy = cy * w.scale;
    const float flx = floorf(sx), fly = floorf(sy);
    LkGeom g;
    g.ix0 = (int)fminf(fmaxf(flx, -1.0e6f), 1.0e6f) - R;  // same clamp as lookup_geom
    g.iy0 = (int)fminf(fmaxf(fly, -1.0e6f), 1.0e6f) - R;
    g.fx = sx - flx; g.fy = sy - fly;
    g.o = g.iy0 * w.Wl + g.ix0;
    g.interior = g.iy0 >= 0 && g.ix0 >= 0 && g.iy0 + F <= w.Hl && g.ix0 + F <= w.Wl;
    return g;
}

// grid: (ceil(ceil(N/LK_TPW)*levels / LK_WARPS), B); out is [B][N][levels*(2R+1)^2].
// Addressing: one 64-bit base per array and 32-bit element offsets (a single IMAD.WIDE per access).
template <int R, int LV>
__global__ void __launch_bounds__(LK_WARPS * 32, LK_MIN_CTAS)
corr_lookup_fwd_cl2_kernel(const float* __restrict__ pyramid, const float* __restrict__ coords, float* __restrict__ out,
                           const PyramidLayout L, int N, int hint) {
    constexpr int D = 2 * R + 1, F = D + 1, FP = F * F, NCH = D * D;
    constexpr int NLD = (FP + 31) / 32, NOUT = (NCH + 31) / 32;
    __shared__ float S[LK_WARPS][LK_TPW][FP + 2];           // [FP], [FP+1] = fx, fy of the task
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.y;
    const int CT = (LV ? LV : L.levels) * NCH;
    const LkWarp w = lk_warp<LV>(blockIdx.x * LK_WARPS + warp, b, N, L);
    if (w.nq <= 0) return;
    const uint64_t pol = hint ? l2_policy_evict_last() : l2_policy_evict_normal();
    const float* lvl = pyramid + w.lvl_off;
    const float* cxp = coords + ((int64_t)b * 2 * N + w.q0);

    int rr[NLD], cc[NLD], off[NLD];                // per-lane (row, col) and element offset of cell lane+32i inside a footprint
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        const int cell = lane + 32 * i;
        rr[i] = cell / F; cc[i] = cell - rr[i] * F;
        if (cell >= FP) rr[i] = 1 << 20;           // never in range
        off[i] = rr[i] * w.Wl + cc[i];
    }
    float cx[LK_TPW], cy[LK_TPW];
#pragma unroll
    for (int j = 0; j < LK_TPW; ++j) {
        const int jj = min(j, w.nq - 1);
        cx[j] = *ptr_add(cxp, jj);
        cy[j] = *ptr_add(cxp, N + jj);
    }
    // Levels 2 and 3 (13x32, 6x16 at 55x128) never contain a whole 10x10 footprint, so every load is bounds-checked:
    // rows [rlo, rhi) and columns [clo, chi) of the footprint lie inside the level (one subtract + one compare each).
    float v[LK_TPW][NLD];
#pragma unroll
    for (int j = 0; j < LK_TPW; ++j) {
        const LkGeom g = lk_geom<R>(cx[j], cy[j], w);
        if (lane == 0) { S[warp][j][FP] = g.fx; S[warp][j][FP + 1] = g.fy; }
        const int e = min(j, w.nq - 1) * w.plane + g.o;
        const int rlo = max(0, -g.iy0), clo = max(0, -g.ix0);
        const unsigned rn = (unsigned)max(0, min(F, w.Hl - g.iy0) - rlo), cn = (unsigned)max(0, min(F, w.Wl - g.ix0) - clo);
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            const bool ok = (unsigned)(rr[i] - rlo) < rn && (unsigned)(cc[i] - clo) < cn;
            v[j][i] = ok ? ldg_policy(ptr_add(lvl, e + off[i]), pol) : 0.f;
        }
    }
#pragma unroll
    for (int j = 0; j < LK_TPW; ++j)
#pragma unroll
        for (int i = 0; i < NLD; ++i)
            if (32 * i + 31 < FP || lane + 32 * i < FP) S[warp][j][lane + 32 * i] = v[j][i];
    __syncwarp();

    int toff[NOUT];                                // output channel -> top-left cell of its 2x2 taps
#pragma unroll
    for (int i = 0; i < NOUT; ++i) {
        const int ch = lane + 32 * i;
        const int a = ch / D, bb = ch - a * D;     // a shifts x, bb shifts y (corr.py:37-43)
        toff[i] = ch < NCH ? bb * F + a : 0;
    }
    float* o = out + ((int64_t)(b * N + w.q0) * CT + (w.l * NCH + lane));
#pragma unroll
    for (int j = 0; j < LK_TPW; ++j) {
        if (j >= w.nq) break;                      // warp-uniform
        const float* Sq = S[warp][j];
        const float fx = Sq[FP], fy = Sq[FP + 1];
        const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
#pragma unroll
        for (int i = 0; i < NOUT; ++i) {
            if (32 * i + 31 < NCH || lane + 32 * i < NCH) {
                const float* p = Sq + toff[i];
                ptr_add(o, j * CT)[32 * i] = w00 * p[0] + w01 * p[1] + w10 * p[F] + w11 * p[F + 1];
            }
        }
    }
}

// grid as above; gout is [B][N][levels*(2R+1)^2].  Every (q, l, cell) address is touched by one lane of one warp per
// launch (race-free); RED because successive lookups accumulate into the same buffer in stream order.
template <int R, int LV>
__global__ void __launch_bounds__(LK_WARPS * 32, LK_MIN_CTAS)
corr_lookup_bwd_cl2_kernel(const float* __restrict__ gout, const float* __restrict__ coords, float* __restrict__ gpyr,
                           const PyramidLayout L, int N, int hint) {
    constexpr int D = 2 * R + 1, F = D + 1, FP = F * F, NCH = D * D;
    constexpr int NLD = (FP + 31) / 32, NOUT = (NCH + 31) / 32;
    constexpr int ZERO = NCH;                      // G[..][ZERO] == 0: target of taps that do not exist
    __shared__ float G[LK_WARPS][LK_TPW][NCH + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.y;
    const int CT = (LV ? LV : L.levels) * NCH;
    const LkWarp w = lk_warp<LV>(blockIdx.x * LK_WARPS + warp, b, N, L);
    if (w.nq <= 0) return;
    const uint64_t pol = hint ? l2_policy_evict_last() : l2_policy_evict_normal();
    const float* cxp = coords + ((int64_t)b * 2 * N + w.q0);
    const float* gi = gout + ((int64_t)(b * N + w.q0) * CT + (w.l * NCH + lane));
    float* lvl = gpyr + w.lvl_off;

    float cx[LK_TPW], cy[LK_TPW];
#pragma unroll
    for (int j = 0; j < LK_TPW; ++j) {
        const int jj = min(j, w.nq - 1);
        cx[j] = *ptr_add(cxp, jj);
        cy[j] = *ptr_add(cxp, N + jj);
#pragma unroll
        for (int i = 0; i < NOUT; ++i)
            if (32 * i + 31 < NCH || lane + 32 * i < NCH) G[warp][j][lane + 32 * i] = __ldg(ptr_add(gi, jj * CT) + 32 * i);
        if (lane == 0) G[warp][j][ZERO] = 0.f;
    }
    // cell (rr, cc) receives tap (a=cc, b=rr)*w00 + (cc-1, rr)*w01 + (cc, rr-1)*w10 + (cc-1, rr-1)*w11; tap (a, b) is channel a*D + b
    int rr[NLD], cc[NLD], off[NLD], i00[NLD], i01[NLD], i10[NLD], i11[NLD];
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        const int cell = lane + 32 * i;
        rr[i] = cell / F; cc[i] = cell - rr[i] * F;
        off[i] = rr[i] * w.Wl + cc[i];
        const bool a0 = cc[i] < D, a1 = cc[i] > 0, b0 = rr[i] < D, b1 = rr[i] > 0, in = cell < FP;
        i00[i] = (in && a0 && b0) ? cc[i] * D + rr[i] : ZERO;
        i01[i] = (in && a1 && b0) ? (cc[i] - 1) * D + rr[i] : ZERO;
        i10[i] = (in && a0 && b1) ? cc[i] * D + rr[i] - 1 : ZERO;
        i11[i] = (in && a1 && b1) ? (cc[i] - 1) * D + rr[i] - 1 : ZERO;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < LK_TPW; ++j) {
        if (j >= w.nq) break;                      // warp-uniform
        const LkGeom g = lk_geom<R>(cx[j], cy[j], w);
        const float w00 = (1.f - g.fx) * (1.f - g.fy), w01 = g.fx * (1.f - g.fy);
        const float w10 = (1.f - g.fx) * g.fy, w11 = g.fx * g.fy;
        const float* Gq = G[warp][j];
        const int e = j * w.plane + g.o;
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            const bool ok = (32 * i + 31 < FP || lane + 32 * i < FP) &&
                            (g.interior || ((unsigned)(g.iy0 + rr[i]) < (unsigned)w.Hl && (unsigned)(g.ix0 + cc[i]) < (unsigned)w.Wl));
            if (ok) {
                const float acc = fmaf(w11, Gq[i11[i]], fmaf(w10, Gq[i10[i]], fmaf(w01, Gq[i01[i]], w00 * Gq[i00[i]])));
                red_add_policy(ptr_add(lvl, e + off[i]), acc, pol);
            }
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

// PCFA_LOOKUP_IMPL=1 selects the first-generation channels-last kernels; PCFA_LOOKUP_L2HINT=0 drops the evict_last policy.
static int lookup_impl() { static const int v = [] { const char* e = getenv("PCFA_LOOKUP_IMPL"); return e ? atoi(e) : 2; }(); return v; }
static int lookup_hint() { static const int v = [] { const char* e = getenv("PCFA_LOOKUP_L2HINT"); return e ? atoi(e) : 1; }(); return v; }

static int lookup_forward(const float* pyramid, const float* coords, float* out, int B, int H, int W, int num_levels,
                          int radius, int out_cl, pcfa_stream_t stream) {
    PCFA_TRY(lookup_check(pyramid, coords, out, B, H, W, num_levels, radius));
    const PyramidLayout L = make_pyramid_layout(B, H, W, num_levels);
    if (out_cl && (radius == 4 || radius == 3) && lookup_impl() == 2 && B <= 65535 &&
        (int64_t)B * H * W * num_levels * (2 * radius + 1) * (2 * radius + 1) < 0x7fffffffLL) {
        const int N = H * W;
        dim3 grid(ceil_div(ceil_div(N, LK_TPW) * num_levels, LK_WARPS), B);
        cudaStream_t s = as_stream(stream);
        const int hint = lookup_hint();
#define PCFA_LK_FWD(RR, LL) corr_lookup_fwd_cl2_kernel<RR, LL><<<grid, LK_WARPS * 32, 0, s>>>(pyramid, coords, out, L, N, hint)
        if (radius == 4) { if (num_levels == 4) PCFA_LK_FWD(4, 4); else PCFA_LK_FWD(4, 0); }
        else             { if (num_levels == 4) PCFA_LK_FWD(3, 4); else PCFA_LK_FWD(3, 0); }
#undef PCFA_LK_FWD
        return after_launch();
    }
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
    if (out_cl && (radius == 4 || radius == 3) && lookup_impl() == 2 && B <= 65535 &&
        (int64_t)B * H * W * num_levels * (2 * radius + 1) * (2 * radius + 1) < 0x7fffffffLL) {
        const int N = H * W;
        dim3 grid(ceil_div(ceil_div(N, LK_TPW) * num_levels, LK_WARPS), B);
        cudaStream_t s = as_stream(stream);
        const int hint = lookup_hint();
#define PCFA_LK_BWD(RR, LL) corr_lookup_bwd_cl2_kernel<RR, LL><<<grid, LK_WARPS * 32, 0, s>>>(grad_out, coords, grad_pyramid, L, N, hint)
        if (radius == 4) { if (num_levels == 4) PCFA_LK_BWD(4, 4); else PCFA_LK_BWD(4, 0); }
        else             { if (num_levels == 4) PCFA_LK_BWD(3, 4); else PCFA_LK_BWD(3, 0); }
#undef PCFA_LK_BWD
        return after_launch();
    }
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

// ---------------------------------------------------------------------------------------------------------------
// Occupancy bitmap of the gradient pyramid for the sparse build backward (common.cuh: OccLayout; pcfa_b200.h).
// ONE launch for all lookups of a backward pass: a warp owns (32 consecutive queries = one bitmap row, one level) and
// walks every lookup's footprint [floor(x/2^l) - r, +2r+2) x [floor(y/2^l) - r, +2r+2) — exactly the cells the
// lookup-backward kernels above may write — setting the bits of the 32-cell chunks each clipped footprint row touches
// in a shared-memory copy of its part of the row, which is then OR-ed into the global bitmap (<= a few words per warp).
namespace pcfa {
constexpr int OCC_MAX_LOOKUPS = 32;
struct OccCoords { const float* p[OCC_MAX_LOOKUPS]; int n; };

// grid (ceil(qgroups*levels/8), B, splits): slice z handles lookups z, z+splits, ...  Lanes are consecutive queries, whose
// footprints mostly fall into the same chunks: a lane first collects its footprint's chunks in a 64-bit register mask, the
// warp then OR-reduces the lanes' masks (REDUX) and four lanes write the result — per-lane, per-row shared-memory atomics on
// the same few words serialise on the SM's shared-memory pipe (12-18 us for the whole launch).
__global__ void __launch_bounds__(256)
occ_mark_kernel(const OccCoords cl, unsigned* __restrict__ occ, const PyramidLayout L, const OccLayout OL, int N, int radius) {
    extern __shared__ unsigned occ_rows[];                      // [8 warps][OL.words]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.y;
    const int gw = blockIdx.x * 8 + warp;
    const int qg = gw / L.levels, l = gw - qg * L.levels;
    if (qg >= OL.qgroups) return;
    unsigned* row = occ_rows + warp * OL.words;
    const int w_lo = OL.chunk_off[l] >> 5, w_hi = (OL.chunk_off[l + 1] + 31) >> 5;      // words holding this level's columns
    for (int w = w_lo + lane; w < w_hi; w += 32) row[w] = 0u;
    __syncwarp();
    const int q = min(qg * 32 + lane, N - 1);                   // tail lanes repeat the last query (same bitmap row)
    const int Hl = L.h[l], Wl = L.w[l], F = 2 * radius + 2, colbase = OL.chunk_off[l];
    const float scale = __int_as_float((127 - l) << 23);
    const float* base = nullptr;
    constexpr int U = 4;                                        // lookups whose coordinates are in flight together
    for (int i0 = blockIdx.z; i0 < cl.n; i0 += U * gridDim.z) {
        float sx[U], sy[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * gridDim.z;
            base = cl.p[i < cl.n ? i : i0] + (int64_t)b * 2 * N;
            sx[u] = __ldg(base + q) * scale; sy[u] = __ldg(base + N + q) * scale;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + u * (int)gridDim.z >= cl.n) break;         // warp-uniform
            const int ix0 = (int)fminf(fmaxf(floorf(sx[u]), -1.0e6f), 1.0e6f) - radius;   // same clamp as lookup_geom / lk_geom
            const int iy0 = (int)fminf(fmaxf(floorf(sy[u]), -1.0e6f), 1.0e6f) - radius;
            const int xlo = max(ix0, 0), xhi = min(ix0 + F, Wl) - 1;
            const int ylo = max(iy0, 0), yhi = min(iy0 + F, Hl);
            // this query's chunks as a 64-bit mask relative to its first chunk (a footprint spans (F-1)*Wl/32 + 2 chunks)
            const bool act = xhi >= xlo && yhi > ylo;
            const int cfirst = colbase + ((ylo * Wl + xlo) >> 5);
            unsigned long long mask = 0ull;
            if (act) {
                for (int y = ylo; y < yhi; ++y) {
                    const int d0 = colbase + ((y * Wl + xlo) >> 5) - cfirst, d1 = colbase + ((y * Wl + xhi) >> 5) - cfirst;
                    if (d1 < 64) mask |= (1ull << d0) | (1ull << d1);
                    else { atomicOr(row + ((cfirst + d0) >> 5), 1u << ((cfirst + d0) & 31)); atomicOr(row + ((cfirst + d1) >> 5), 1u << ((cfirst + d1) & 31)); }
                }
            }
            // merge the 32 queries: masks whose first chunk lies within 64 chunks of the warp's (word-aligned) minimum are
            // OR-reduced into four words (REDUX) and written by four lanes; scattered ones (noisy flow) go in directly
            const int cmin = __reduce_min_sync(0xffffffffu, act ? cfirst : 0x7fffffff);
            if (cmin == 0x7fffffff) continue;                                               // warp-uniform
            const int cbase = cmin & ~31, off = cfirst - cbase;
            const bool near = act && off < 64;
            unsigned v[4] = {0u, 0u, 0u, 0u};
            if (near) {
                const unsigned long long lo = mask << (off & 63), hi = (off & 63) ? (mask >> (64 - (off & 63))) : 0ull;
                v[0] = (unsigned)lo; v[1] = (unsigned)(lo >> 32); v[2] = (unsigned)hi; v[3] = (unsigned)(hi >> 32);
            }
            unsigned mine = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) { const unsigned t = __reduce_or_sync(0xffffffffu, v[k]); if (lane == k) mine = t; }
            if (lane < 4 && mine) atomicOr(row + (cbase >> 5) + lane, mine);
            if (act && !near && mask) {
                const int w = cfirst >> 5, sh = cfirst & 31;
                const unsigned long long lo = mask << sh;
                atomicOr(row + w, (unsigned)lo);
                if ((unsigned)(lo >> 32)) atomicOr(row + w + 1, (unsigned)(lo >> 32));
                if (sh && (unsigned)(mask >> (64 - sh))) atomicOr(row + w + 2, (unsigned)(mask >> (64 - sh)));
            }
        }
    }
    __syncwarp();
    unsigned* orow = occ + ((int64_t)b * OL.qgroups + qg) * OL.words;
    for (int w = w_lo + lane; w < w_hi; w += 32) {
        const unsigned v = row[w];
        if (v) atomicOr(orow + w, v);
    }
}
}  // namespace pcfa

extern "C" int64_t pcfa_corr_occupancy_bytes(int B, int H, int W, int num_levels) {
    if (B <= 0 || H <= 0 || W <= 0 || num_levels <= 0 || num_levels > 8) return 0;
    const OccLayout OL = make_occ_layout(H, W, num_levels);
    return (int64_t)B * OL.qgroups * OL.words * 4;
}

extern "C" int pcfa_corr_occupancy_mark(const float* const* coords_list, int n_lookups, uint32_t* occupancy, int B, int H, int W,
                                        int num_levels, int radius, pcfa_stream_t stream) {
    if (!coords_list || n_lookups < 0 || !occupancy || (reinterpret_cast<uintptr_t>(occupancy) & 3) || B <= 0 || B > 65535 ||
        H <= 0 || W <= 0 || num_levels <= 0 || num_levels > 8 || radius < 0 || radius > 15)
        return PCFA_E_BADARG;
    const PyramidLayout L = make_pyramid_layout(B, H, W, num_levels);
    const OccLayout OL = make_occ_layout(H, W, num_levels);
    const size_t smem = (size_t)8 * OL.words * sizeof(unsigned);
    if (smem > 48 * 1024) return PCFA_E_TOOLARGE;
    for (int i0 = 0; i0 < n_lookups; i0 += OCC_MAX_LOOKUPS) {
        OccCoords cl{};
        cl.n = n_lookups - i0 < OCC_MAX_LOOKUPS ? n_lookups - i0 : OCC_MAX_LOOKUPS;
        for (int i = 0; i < cl.n; ++i) {
            if (!coords_list[i0 + i]) return PCFA_E_BADARG;
            cl.p[i] = coords_list[i0 + i];
        }
        dim3 grid(ceil_div(OL.qgroups * num_levels, 8), B, cl.n >= 9 ? 4 : cl.n >= 3 ? 2 : 1);
        occ_mark_kernel<<<grid, 256, smem, as_stream(stream)>>>(cl, reinterpret_cast<unsigned*>(occupancy), L, OL, H * W, radius);
        PCFA_TRY(after_launch());
    }
    return PCFA_OK;
}
