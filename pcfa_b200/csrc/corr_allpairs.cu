// All-pairs correlation pyramid: host-side orchestration + the exact-fp32 SIMT variant.
// Replaces CorrBlock.__init__/CorrBlock.corr (models/raft/corr.py:13-27,52-60) and the autograd
// backward of torch.matmul + F.avg_pool2d.  The tensor-core variant lives in
// corr_allpairs_tc.cu; this file owns the C-ABI entry points and dispatches.
#include "common.cuh"
#include "sgemm_simt.cuh"
#include <math.h>

namespace pcfa {

// --- tcgen05 variant (corr_allpairs_tc.cu) -------------------------------------------------
int corr_pyramid_forward_tc(const float* fmap1, const float* fmap2, float* pyramid, void* ws,
                            int64_t ws_bytes, int B, int C, int H, int W, int levels,
                            cudaStream_t s, int two_cta);
int64_t corr_pyramid_tc_workspace_bytes(int B, int C, int H, int W, int levels);
bool corr_pyramid_tc_supported(int B, int C, int H, int W, int levels);
// --- tcgen05 backward (corr_allpairs_bwd_tc.cu) ----------------------------------------------
int corr_pyramid_backward_tc(const float* gpyr, const float* f1, const float* f2, float* gf1, float* gf2, void* ws,
                             int64_t ws_bytes, int B, int C, int H, int W, int levels, cudaStream_t s, int two_cta, const unsigned* occ);
int64_t corr_pyramid_bwd_tc_workspace_bytes(int B, int C, int H, int W, int levels);
bool corr_pyramid_bwd_tc_supported(int B, int C, int H, int W, int levels);

// 2x2 average pooling with floor output size over R independent images (F.avg_pool2d(x,2,2)).
__global__ void avgpool2_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t R,
                                int Hi, int Wi, int Ho, int Wo) {
    const int64_t total = R * Ho * Wo;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % Wo);
        const int y = (int)((idx / Wo) % Ho);
        const int64_t r = idx / ((int64_t)Wo * Ho);
        const float* p = in + (r * Hi + 2 * y) * (int64_t)Wi + 2 * x;
        out[idx] = 0.25f * ((p[0] + p[1]) + (p[Wi] + p[Wi + 1]));
    }
}

// grad_fmap2[r, y, x] += sum_{l>=1} gP_l[r, y>>l, x>>l] * 0.25^l   (adjoint of successive pooling)
struct UnpoolArgs { const float* g[8]; int h[8], w[8]; int levels; };
__global__ void unpool_accumulate_kernel(float* __restrict__ gf2, UnpoolArgs a, int64_t R, int H,
                                         int W) {
    const int64_t total = R * H * W;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % W);
        const int y = (int)((idx / W) % H);
        const int64_t r = idx / ((int64_t)W * H);
        float acc = gf2[idx];
        float sc = 1.f;
        for (int l = 1; l < a.levels; ++l) {
            sc *= 0.25f;
            const int yy = y >> l, xx = x >> l;
            if (yy < a.h[l] && xx < a.w[l])
                acc += sc * a.g[l][(r * a.h[l] + yy) * (int64_t)a.w[l] + xx];
        }
        gf2[idx] = acc;
    }
}

static int grid_for(int64_t total, int threads) {
    int64_t b = ceil_div<int64_t>(total, threads);
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// Scratch: pooled fmap2 levels 1.. and their gradients (backward only).
static int64_t pooled_floats(int B, int C, int H, int W, int levels) {
    int64_t n = 0; int h = H, w = W;
    for (int l = 1; l < levels; ++l) { h /= 2; w /= 2; n += (int64_t)B * C * h * w; }
    return n;
}

static int forward_simt(const float* f1, const float* f2, float* pyr, int B, int C, int H, int W,
                        int levels, cudaStream_t s) {
    const PyramidLayout L = make_pyramid_layout(B, H, W, levels);
    const int N = H * W;
    GemmArgs g{};
    g.A = f1; g.B = f2; g.C = pyr;
    g.M = N; g.N = N; g.K = C;
    g.a_sm = 1; g.a_sk = N; g.b_sn = 1; g.b_sk = N; g.c_sm = N; g.c_sn = 1;
    g.a_sb = (int64_t)C * N; g.b_sb = (int64_t)C * N; g.c_sb = (int64_t)N * N;
    g.alpha = 1.0f / sqrtf((float)C);     // corr / torch.sqrt(torch.tensor(dim).float())  corr.py:60
    g.accumulate = 0;
    PCFA_TRY(launch_sgemm(g, B, s));
    for (int l = 1; l < levels; ++l) {
        if (L.h[l] == 0 || L.w[l] == 0) break;
        const int64_t R = (int64_t)B * N;
        avgpool2_kernel<<<grid_for(R * L.h[l] * L.w[l], 256), 256, 0, s>>>(
            pyr + L.off[l - 1], pyr + L.off[l], R, L.h[l - 1], L.w[l - 1], L.h[l], L.w[l]);
        PCFA_TRY(after_launch());
    }
    return PCFA_OK;
}

static int backward_simt(const float* gpyr, const float* f1, const float* f2, float* gf1, float* gf2,
                         float* ws, int B, int C, int H, int W, int levels, cudaStream_t s) {
    const PyramidLayout L = make_pyramid_layout(B, H, W, levels);
    const int N = H * W;
    const float alpha = 1.0f / sqrtf((float)C);
    const int64_t npool = pooled_floats(B, C, H, W, levels);
    float* P  = ws;            // pooled fmap2, levels 1..
    float* gP = ws + npool;    // their gradients

    // pooled copies of fmap2 (level l from level l-1)
    const float* Pl[8]; float* gPl[8];
    Pl[0] = f2; gPl[0] = gf2;
    {
        int64_t o = 0;
        for (int l = 1; l < levels; ++l) {
            Pl[l] = P + o; gPl[l] = gP + o;
            const int64_t cnt = (int64_t)B * C * L.h[l] * L.w[l];
            if (cnt > 0) {
                avgpool2_kernel<<<grid_for(cnt, 256), 256, 0, s>>>(
                    Pl[l - 1], P + o, (int64_t)B * C, L.h[l - 1], L.w[l - 1], L.h[l], L.w[l]);
                PCFA_TRY(after_launch());
            }
            o += cnt;
        }
    }
    for (int l = 0; l < levels; ++l) {
        const int Nl = L.h[l] * L.w[l];
        if (Nl == 0) continue;
        const float* gl = gpyr + L.off[l];
        // grad_fmap1[c, i] (+)= alpha * sum_j g_l[i, j] * P_l[c, j]
        GemmArgs a{};
        a.A = gl; a.B = Pl[l]; a.C = gf1;
        a.M = N; a.N = C; a.K = Nl;
        a.a_sm = Nl; a.a_sk = 1; a.b_sn = Nl; a.b_sk = 1; a.c_sm = 1; a.c_sn = N;
        a.a_sb = (int64_t)N * Nl; a.b_sb = (int64_t)C * Nl; a.c_sb = (int64_t)C * N;
        a.alpha = alpha; a.accumulate = (l > 0);
        PCFA_TRY(launch_sgemm(a, B, s));
        // grad_P_l[c, j] = alpha * sum_i g_l[i, j] * fmap1[c, i]
        GemmArgs b{};
        b.A = gl; b.B = f1; b.C = gPl[l];
        b.M = Nl; b.N = C; b.K = N;
        b.a_sm = 1; b.a_sk = Nl; b.b_sn = N; b.b_sk = 1; b.c_sm = 1; b.c_sn = Nl;
        b.a_sb = (int64_t)N * Nl; b.b_sb = (int64_t)C * N; b.c_sb = (int64_t)C * Nl;
        b.alpha = alpha; b.accumulate = 0;
        PCFA_TRY(launch_sgemm(b, B, s));
    }
    if (levels > 1) {
        UnpoolArgs u{};
        u.levels = levels;
        for (int l = 0; l < levels; ++l) { u.g[l] = gPl[l]; u.h[l] = L.h[l]; u.w[l] = L.w[l]; }
        const int64_t total = (int64_t)B * C * N;
        unpool_accumulate_kernel<<<grid_for(total, 256), 256, 0, s>>>(gf2, u, (int64_t)B * C, H, W);
        PCFA_TRY(after_launch());
    }
    return PCFA_OK;
}

}  // namespace pcfa

using namespace pcfa;

extern "C" int pcfa_corr_pyramid_layout(int B, int H, int W, int num_levels, int64_t* offsets,
                                        int* hs, int* ws) {
    if (B <= 0 || H <= 0 || W <= 0 || num_levels <= 0 || num_levels > 8 || !offsets || !hs || !ws)
        return PCFA_E_BADARG;
    const PyramidLayout L = make_pyramid_layout(B, H, W, num_levels);
    for (int l = 0; l < num_levels; ++l) { offsets[l] = L.off[l]; hs[l] = L.h[l]; ws[l] = L.w[l]; }
    offsets[num_levels] = L.off[num_levels];
    return PCFA_OK;
}

extern "C" int64_t pcfa_corr_pyramid_workspace_bytes(int B, int C, int H, int W, int num_levels) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || num_levels <= 0 || num_levels > 8) return 0;
    const int64_t simt = 2 * pooled_floats(B, C, H, W, num_levels) * (int64_t)sizeof(float);
    const int64_t tcf  = corr_pyramid_tc_workspace_bytes(B, C, H, W, num_levels);
    const int64_t tcb  = corr_pyramid_bwd_tc_workspace_bytes(B, C, H, W, num_levels);
    const int64_t tc   = tcf > tcb ? tcf : tcb;
    const int64_t m    = simt > tc ? simt : tc;
    return m > 256 ? m : 256;
}

static int pyramid_check(int B, int C, int H, int W, int levels) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || levels <= 0 || levels > 8) return PCFA_E_BADARG;
    if ((int64_t)H * W > 46340) return PCFA_E_TOOLARGE;   // N*N must fit the int32 tile maps
    return PCFA_OK;
}

extern "C" int pcfa_corr_pyramid_forward(const float* fmap1, const float* fmap2, float* pyramid,
                                         void* workspace, int64_t workspace_bytes, int B, int C,
                                         int H, int W, int num_levels, int impl,
                                         pcfa_stream_t stream) {
    if (!fmap1 || !fmap2 || !pyramid) return PCFA_E_BADARG;
    PCFA_TRY(pyramid_check(B, C, H, W, num_levels));
    const bool tc_ok = corr_pyramid_tc_supported(B, C, H, W, num_levels);
    if ((impl == 2 || impl == 3) && !tc_ok) return PCFA_E_BADARG;
    if (impl == 2 || impl == 3 || (impl == 0 && tc_ok)) {
        if (!workspace || workspace_bytes < corr_pyramid_tc_workspace_bytes(B, C, H, W, num_levels))
            return PCFA_E_WORKSPACE;
        return corr_pyramid_forward_tc(fmap1, fmap2, pyramid, workspace, workspace_bytes, B, C, H, W,
                                       num_levels, as_stream(stream), impl == 2 ? 0 : 1);   // auto and 3: cta_group::2 pairs
    }
    return forward_simt(fmap1, fmap2, pyramid, B, C, H, W, num_levels, as_stream(stream));
}

static int pyramid_backward(const float* grad_pyramid, const unsigned* occ, const float* fmap1,
                                          const float* fmap2, float* grad_fmap1, float* grad_fmap2,
                                          void* workspace, int64_t workspace_bytes, int B, int C,
                                          int H, int W, int num_levels, int impl,
                                          pcfa_stream_t stream) {
    if (!grad_pyramid || !fmap1 || !fmap2 || !grad_fmap1 || !grad_fmap2) return PCFA_E_BADARG;
    PCFA_TRY(pyramid_check(B, C, H, W, num_levels));
    const bool tc_ok = corr_pyramid_bwd_tc_supported(B, C, H, W, num_levels);
    if ((impl == 2 || impl == 3) && !tc_ok) return PCFA_E_BADARG;
    if (impl == 2 || impl == 3 || (impl == 0 && tc_ok)) {
        if (!workspace || workspace_bytes < corr_pyramid_bwd_tc_workspace_bytes(B, C, H, W, num_levels))
            return PCFA_E_WORKSPACE;
        return corr_pyramid_backward_tc(grad_pyramid, fmap1, fmap2, grad_fmap1, grad_fmap2, workspace,
                                        workspace_bytes, B, C, H, W, num_levels, as_stream(stream), impl == 2 ? 0 : 1, occ);
    }
    const int64_t need = 2 * pooled_floats(B, C, H, W, num_levels) * (int64_t)sizeof(float);
    if (need > 0 && (!workspace || workspace_bytes < need)) return PCFA_E_WORKSPACE;
    return backward_simt(grad_pyramid, fmap1, fmap2, grad_fmap1, grad_fmap2,
                         reinterpret_cast<float*>(workspace), B, C, H, W, num_levels,
                         as_stream(stream));
}

extern "C" int pcfa_corr_pyramid_backward(const float* grad_pyramid, const float* fmap1, const float* fmap2, float* grad_fmap1,
                                          float* grad_fmap2, void* workspace, int64_t workspace_bytes, int B, int C, int H,
                                          int W, int num_levels, int impl, pcfa_stream_t stream) {
    return pyramid_backward(grad_pyramid, nullptr, fmap1, fmap2, grad_fmap1, grad_fmap2, workspace, workspace_bytes, B, C, H, W,
                            num_levels, impl, stream);
}
// Sparse variant: `occupancy` (pcfa_corr_lookup_backward_cl_occ) names the 32x32 blocks of grad_pyramid that may be non-zero;
// the tensor-core CTA-pair path skips every K-chunk without marked blocks, the other paths ignore the bitmap (dense).
extern "C" int pcfa_corr_pyramid_backward_occ(const float* grad_pyramid, const uint32_t* occupancy, const float* fmap1,
                                              const float* fmap2, float* grad_fmap1, float* grad_fmap2, void* workspace,
                                              int64_t workspace_bytes, int B, int C, int H, int W, int num_levels, int impl,
                                              pcfa_stream_t stream) {
    if (!occupancy || (reinterpret_cast<uintptr_t>(occupancy) & 3)) return PCFA_E_BADARG;
    return pyramid_backward(grad_pyramid, occupancy, fmap1, fmap2, grad_fmap1, grad_fmap2, workspace, workspace_bytes, B, C, H, W,
                            num_levels, impl, stream);
}
