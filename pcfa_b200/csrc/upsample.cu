// Convex up-sampling of the flow field: RAFT.upsample_flow, models/raft/raft.py:72-83
//   mask = softmax(mask.view(N,1,9,8,8,H,W), dim=2);  up = unfold(8*flow, 3x3, pad 1).view(N,2,9,1,1,H,W)
//   out  = sum(mask*up, dim=2).permute(0,1,4,2,5,3).reshape(N,2,8H,8W)
// (~10 ATen launches forward and ~12 backward, plus two layout conversions of the 16 MB mask when the update block
// runs channels-last) as ONE forward kernel and two backward kernels.
//
// Layout: `mask` is the mask head's output in channels-last memory, [N][H][W][576] with channel = k*64 + i*8 + j
// (k = 3x3 tap in unfold's row-major order, (i, j) = sub-pixel); `mask_scale` folds the 0.25 of update.py:135.
// Thread = (coarse pixel, sub-pixel): the 64 threads of a pixel read 9 x 256-byte channel runs (coalesced), the taps'
// flow values are broadcast loads, and each store instruction writes four 32-byte row segments of `up`.
// HBM-bound: 16.2 MB mask + 3.6 MB up per sample forward; mask + grad_mask + grad_up backward (36 MB).
// Backward of the flow is a gather over per-pixel tap sums (no atomics, bit-reproducible).
#include "common.cuh"

namespace pcfa {

constexpr int UP_PIX = 4;             // coarse pixels per CTA (256 threads)

struct UpTaps { float p[9]; float f0[9], f1[9]; };

__device__ __forceinline__ void up_softmax_and_taps(const float* __restrict__ flow, const float* __restrict__ mp, int n, int h,
                                                    int w, int H, int W, float mscale, UpTaps& t) {
    float mx = -3.0e38f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { t.p[k] = __ldg(mp + k * 64) * mscale; mx = fmaxf(mx, t.p[k]); }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { t.p[k] = __expf(t.p[k] - mx); sum += t.p[k]; }
    const float inv = 1.f / sum;
    const int64_t HW = (int64_t)H * W;
    const float* f0p = flow + (int64_t)n * 2 * HW;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        t.p[k] *= inv;
        const int y = h + k / 3 - 1, x = w + k % 3 - 1;
        const bool ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;      // unfold pads with zeros
        t.f0[k] = ok ? 8.f * __ldg(f0p + y * W + x) : 0.f;
        t.f1[k] = ok ? 8.f * __ldg(f0p + HW + y * W + x) : 0.f;
    }
}

__global__ void __launch_bounds__(64 * UP_PIX)
convex_up_fwd_kernel(const float* __restrict__ flow, const float* __restrict__ mask, float* __restrict__ up, int N, int H, int W,
                     float mscale) {
    const int ij = threadIdx.x & 63;
    const int64_t pix = (int64_t)blockIdx.x * UP_PIX + (threadIdx.x >> 6);
    const int64_t HW = (int64_t)H * W;
    if (pix >= N * HW) return;
    const int n = (int)(pix / HW), hw = (int)(pix - n * HW), h = hw / W, w = hw - h * W;
    UpTaps t;
    up_softmax_and_taps(flow, mask + pix * 576 + ij, n, h, w, H, W, mscale, t);
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { o0 = fmaf(t.p[k], t.f0[k], o0); o1 = fmaf(t.p[k], t.f1[k], o1); }
    const int i = ij >> 3, j = ij & 7;
    const int64_t W8 = 8 * (int64_t)W, plane = 64 * HW;
    float* o = up + (int64_t)n * 2 * plane + (8 * h + i) * W8 + 8 * w + j;
    o[0] = o0;
    o[plane] = o1;
}

// grad_mask (channels-last, like mask) and the per-pixel tap sums  taps[pix][k*2+c] = 8 * sum_ij p_k * g_c.
__global__ void __launch_bounds__(64 * UP_PIX)
convex_up_bwd_kernel(const float* __restrict__ flow, const float* __restrict__ mask, const float* __restrict__ gup,
                     float* __restrict__ gmask, float* __restrict__ taps, int N, int H, int W, float mscale) {
    __shared__ float red[UP_PIX][2][18];
    const int ij = threadIdx.x & 63, slot = threadIdx.x >> 6, half = (threadIdx.x >> 5) & 1, lane = threadIdx.x & 31;
    const int64_t pix = (int64_t)blockIdx.x * UP_PIX + slot;
    const int64_t HW = (int64_t)H * W;
    const bool live = pix < N * HW;
    float tsum[18];
    if (live) {
        const int n = (int)(pix / HW), hw = (int)(pix - n * HW), h = hw / W, w = hw - h * W;
        UpTaps t;
        up_softmax_and_taps(flow, mask + pix * 576 + ij, n, h, w, H, W, mscale, t);
        const int i = ij >> 3, j = ij & 7;
        const int64_t W8 = 8 * (int64_t)W, plane = 64 * HW;
        const float* g = gup + (int64_t)n * 2 * plane + (8 * h + i) * W8 + 8 * w + j;
        const float g0 = __ldg(g), g1 = __ldg(g + plane);
        float dp[9], s = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) { dp[k] = t.f0[k] * g0 + t.f1[k] * g1; s = fmaf(t.p[k], dp[k], s); }
        float* gm = gmask + pix * 576 + ij;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            gm[k * 64] = mscale * t.p[k] * (dp[k] - s);
            tsum[2 * k] = 8.f * t.p[k] * g0;
            tsum[2 * k + 1] = 8.f * t.p[k] * g1;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 18; ++k) tsum[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 18; ++k) {
        const float v = warp_sum(tsum[k]);
        if (lane == 0) red[slot][half][k] = v;
    }
    __syncthreads();
    if (live && ij < 18) taps[pix * 18 + ij] = red[slot][0][ij] + red[slot][1][ij];
}

// grad_flow[n,c,y,x] = sum_k taps[pixel (y-ky+1, x-kx+1)][k*2+c]   (adjoint of unfold's zero-padded 3x3 gather)
__global__ void __launch_bounds__(256)
convex_up_bwd_flow_kernel(const float* __restrict__ taps, float* __restrict__ gflow, int N, int H, int W) {
    const int64_t HW = (int64_t)H * W, total = (int64_t)N * 2 * HW;
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= total) return;
    const int n = (int)(e / (2 * HW)), c = (int)((e / HW) & 1), hw = (int)(e % HW), y = hw / W, x = hw - y * W;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int h = y - (k / 3) + 1, w = x - (k % 3) + 1;
        if ((unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W)
            acc += __ldg(taps + ((int64_t)n * HW + h * W + w) * 18 + 2 * k + c);
    }
    gflow[e] = acc;
}

}  // namespace pcfa

using namespace pcfa;

static int up_check(const void* a, const void* b, const void* c, int N, int H, int W) {
    if (!a || !b || !c || N <= 0 || H <= 0 || W <= 0) return PCFA_E_BADARG;
    if ((int64_t)N * H * W * 576 > 0x7fffffffffLL || ceil_div((int64_t)N * H * W, (int64_t)UP_PIX) > 0x7fffffffLL) return PCFA_E_TOOLARGE;
    return PCFA_OK;
}

extern "C" int64_t pcfa_convex_upsample_workspace_bytes(int N, int H, int W) { return (int64_t)N * H * W * 18 * 4; }

extern "C" int pcfa_convex_upsample_forward(const float* flow, const float* mask_cl, float* up, int N, int H, int W, float mask_scale,
                                            pcfa_stream_t stream) {
    PCFA_TRY(up_check(flow, mask_cl, up, N, H, W));
    const unsigned grid = (unsigned)ceil_div((int64_t)N * H * W, (int64_t)UP_PIX);
    convex_up_fwd_kernel<<<grid, 64 * UP_PIX, 0, as_stream(stream)>>>(flow, mask_cl, up, N, H, W, mask_scale);
    return after_launch();
}

extern "C" int pcfa_convex_upsample_backward(const float* flow, const float* mask_cl, const float* grad_up, float* grad_flow,
                                             float* grad_mask_cl, void* workspace, int64_t workspace_bytes, int N, int H, int W,
                                             float mask_scale, pcfa_stream_t stream) {
    PCFA_TRY(up_check(flow, mask_cl, grad_up, N, H, W));
    if (!grad_flow || !grad_mask_cl) return PCFA_E_BADARG;
    if (!workspace || workspace_bytes < pcfa_convex_upsample_workspace_bytes(N, H, W)) return PCFA_E_WORKSPACE;
    float* taps = reinterpret_cast<float*>(workspace);
    const unsigned grid = (unsigned)ceil_div((int64_t)N * H * W, (int64_t)UP_PIX);
    convex_up_bwd_kernel<<<grid, 64 * UP_PIX, 0, as_stream(stream)>>>(flow, mask_cl, grad_up, grad_mask_cl, taps, N, H, W, mask_scale);
    PCFA_TRY(after_launch());
    const unsigned g2 = (unsigned)ceil_div((int64_t)N * 2 * H * W, (int64_t)256);
    convex_up_bwd_flow_kernel<<<g2, 256, 0, as_stream(stream)>>>(taps, grad_flow, N, H, W);
    return after_launch();
}
