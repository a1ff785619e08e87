// Backward warping operators (memory-bound gathers forward, scatters backward).
//   * Resample2d  — FlowNet2 (resample2d_package/resample2d_kernel.cu:15-198): border-clamped
//     bilinear, with the reference's floor-vs-trunc asymmetry between forward and image-gradient.
//   * PWC warp    — PWCDCNet.warp (models/PWCNet/PWCNet.py:166-206): grid_sample(bilinear, zeros,
//     align_corners=False) of x and of a ones tensor, thresholded mask, product.
// One thread owns one output pixel and walks the channels, so the flow / coordinates / weights are
// computed once per pixel instead of once per element as in the reference.  Image gradients are
// scattered with RED atomics; lanes of a warp hit neighbouring addresses (the flow field is smooth)
// so the L2 sees coalesced atomic sectors.
#include "common.cuh"

namespace pcfa {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return max(min(v, hi), lo); }

// ---------------------------------------------------------------------------------- Resample2d
__global__ void resample2d_fwd_kernel(const float* __restrict__ img, const float* __restrict__ flow,
                                      float* __restrict__ out, int B, int C, int H, int W, int oH,
                                      int oW, int bilinear) {
    const int64_t npix = (int64_t)B * oH * oW;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(p % oW);
        const int y = (int)((p / oW) % oH);
        const int b = (int)(p / ((int64_t)oW * oH));
        const int64_t opl = (int64_t)oH * oW, ipl = (int64_t)H * W;
        const float dx = flow[((int64_t)b * 2 + 0) * opl + (int64_t)y * oW + x];
        const float dy = flow[((int64_t)b * 2 + 1) * opl + (int64_t)y * oW + x];
        const float xf = (float)x + dx, yf = (float)y + dy;
        const float* ib = img + (int64_t)b * C * ipl;
        float* ob = out + (int64_t)b * C * opl + (int64_t)y * oW + x;
        if (bilinear) {
            const float fxf = floorf(xf), fyf = floorf(yf);
            const float alpha = xf - fxf, beta = yf - fyf;                 // kernel.cu:45-46
            const int xL = clampi((int)fxf, 0, W - 1), xR = clampi((int)(fxf + 1.f), 0, W - 1);
            const int yT = clampi((int)fyf, 0, H - 1), yB = clampi((int)(fyf + 1.f), 0, H - 1);
            const float w00 = (1.f - alpha) * (1.f - beta), w01 = alpha * (1.f - beta);
            const float w10 = (1.f - alpha) * beta, w11 = alpha * beta;
            for (int c = 0; c < C; ++c) {
                const float* ic = ib + c * ipl;
                float v = w00 * __ldg(ic + (int64_t)yT * W + xL);
                v += w01 * __ldg(ic + (int64_t)yT * W + xR);
                v += w10 * __ldg(ic + (int64_t)yB * W + xL);
                v += w11 * __ldg(ic + (int64_t)yB * W + xR);
                ob[c * opl] = v;
            }
        } else {
            const int xN = clampi((int)floorf(xf + 0.5f), 0, W - 1);       // kernel.cu:64-65
            const int yN = clampi((int)floorf(yf + 0.5f), 0, H - 1);
            for (int c = 0; c < C; ++c) ob[c * opl] = __ldg(ib + c * ipl + (int64_t)yN * W + xN);
        }
    }
}

__global__ void resample2d_bwd_kernel(const float* __restrict__ img, const float* __restrict__ flow,
                                      const float* __restrict__ gout, float* __restrict__ gimg,
                                      float* __restrict__ gflow, int B, int C, int H, int W, int oH,
                                      int oW) {
    const int64_t npix = (int64_t)B * oH * oW;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(p % oW);
        const int y = (int)((p / oW) % oH);
        const int b = (int)(p / ((int64_t)oW * oH));
        const int64_t opl = (int64_t)oH * oW, ipl = (int64_t)H * W;
        const float dx = flow[((int64_t)b * 2 + 0) * opl + (int64_t)y * oW + x];
        const float dy = flow[((int64_t)b * 2 + 1) * opl + (int64_t)y * oW + x];
        const float xf = (float)x + dx, yf = (float)y + dy;
        const float fxf = floorf(xf), fyf = floorf(yf);
        const int xL = clampi((int)fxf, 0, W - 1), xR = clampi((int)(fxf + 1.f), 0, W - 1);
        const int yT = clampi((int)fyf, 0, H - 1), yB = clampi((int)(fyf + 1.f), 0, H - 1);
        // image gradient weights: alpha = xf - int(xf)  (truncation, kernel.cu:105-106)
        const float at = xf - (float)(int)xf, bt = yf - (float)(int)yf;
        const float w00 = (1.f - at) * (1.f - bt), w01 = at * (1.f - bt);
        const float w10 = (1.f - at) * bt, w11 = at * bt;
        // flow gradient weights use floor (kernel.cu:163-193): gamma = 1 - (xf - floor(xf))
        const float gx = 1.f - (xf - fxf), gy = 1.f - (yf - fyf);
        const float* ib = img + (int64_t)b * C * ipl;
        float* gb = gimg + (int64_t)b * C * ipl;
        const float* go = gout + (int64_t)b * C * opl + (int64_t)y * oW + x;
        float gfx = 0.f, gfy = 0.f;
        for (int c = 0; c < C; ++c) {
            const float g = __ldg(go + c * opl);
            const float* ic = ib + c * ipl;
            float* gc = gb + c * ipl;
            const float vTL = __ldg(ic + (int64_t)yT * W + xL), vTR = __ldg(ic + (int64_t)yT * W + xR);
            const float vBL = __ldg(ic + (int64_t)yB * W + xL), vBR = __ldg(ic + (int64_t)yB * W + xR);
            red_add(gc + (int64_t)yT * W + xL, w00 * g);
            red_add(gc + (int64_t)yT * W + xR, w01 * g);
            red_add(gc + (int64_t)yB * W + xL, w10 * g);
            red_add(gc + (int64_t)yB * W + xR, w11 * g);
            // c%2 == 0 (d/dflow_x): gamma = 1-(yf-floor yf):  g*[gy*(TR-TL) + (1-gy)*(BR-BL)]
            gfx += g * (gy * (vTR - vTL) + (1.f - gy) * (vBR - vBL));
            // c%2 == 1 (d/dflow_y): gamma = 1-(xf-floor xf):  g*[gx*(BL-TL) + (1-gx)*(BR-TR)]
            gfy += g * (gx * (vBL - vTL) + (1.f - gx) * (vBR - vTR));
        }
        gflow[((int64_t)b * 2 + 0) * opl + (int64_t)y * oW + x] = gfx;
        gflow[((int64_t)b * 2 + 1) * opl + (int64_t)y * oW + x] = gfy;
    }
}

// ---------------------------------------------------------------------------------- PWC warp
struct WarpTaps {
    int x0, y0;            // north-west corner
    float wx1, wy1;        // weight of the +1 neighbour along x / y (fractional part)
    bool vx0, vx1, vy0, vy1;
    float mask;
};

__device__ __forceinline__ WarpTaps pwc_taps(int x, int y, float fx, float fy, int H, int W) {
    // vgrid = 2*(x+flo)/max(W-1,1) - 1  (PWCNet.py:189-190), then grid_sample's align_corners=False
    // un-normalisation ix = ((g+1)*W - 1)/2.
    const float gxn = 2.0f * ((float)x + fx) / (float)max(W - 1, 1) - 1.0f;
    const float gyn = 2.0f * ((float)y + fy) / (float)max(H - 1, 1) - 1.0f;
    const float ix = ((gxn + 1.f) * (float)W - 1.f) * 0.5f;
    const float iy = ((gyn + 1.f) * (float)H - 1.f) * 0.5f;
    const float flx = floorf(ix), fly = floorf(iy);
    WarpTaps t;
    t.x0 = (int)fminf(fmaxf(flx, -2.0e6f), 2.0e6f);
    t.y0 = (int)fminf(fmaxf(fly, -2.0e6f), 2.0e6f);
    t.wx1 = ix - flx; t.wy1 = iy - fly;
    t.vx0 = t.x0 >= 0 && t.x0 < W;  t.vx1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    t.vy0 = t.y0 >= 0 && t.y0 < H;  t.vy1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    // grid_sample(ones): sum of the in-bounds corner weights, then (mask >= 0.0001)  PWCNet.py:195-204
    float m = 0.f;
    if (t.vy0 && t.vx0) m += (1.f - t.wx1) * (1.f - t.wy1);
    if (t.vy0 && t.vx1) m += t.wx1 * (1.f - t.wy1);
    if (t.vy1 && t.vx0) m += (1.f - t.wx1) * t.wy1;
    if (t.vy1 && t.vx1) m += t.wx1 * t.wy1;
    t.mask = (m >= 0.0001f) ? 1.f : 0.f;
    return t;
}

// block = (pixels, channel slices): threadIdx.y walks channels c = y, y + blockDim.y, ...  Small pyramid levels (PWCNet's
// 6x20 .. 24x80 maps with 96-196 channels) have too few pixels to fill the GPU with one thread per pixel.
__global__ void pwc_warp_fwd_kernel(const float* __restrict__ xin, const float* __restrict__ flow,
                                    float* __restrict__ out, int B, int C, int H, int W) {
    const int64_t npix = (int64_t)B * H * W, pl = (int64_t)H * W;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(p % W);
        const int y = (int)((p / W) % H);
        const int b = (int)(p / pl);
        const float fx = flow[((int64_t)b * 2 + 0) * pl + (int64_t)y * W + x];
        const float fy = flow[((int64_t)b * 2 + 1) * pl + (int64_t)y * W + x];
        const WarpTaps t = pwc_taps(x, y, fx, fy, H, W);
        const float w00 = (t.vy0 && t.vx0) ? (1.f - t.wx1) * (1.f - t.wy1) * t.mask : 0.f;
        const float w01 = (t.vy0 && t.vx1) ? t.wx1 * (1.f - t.wy1) * t.mask : 0.f;
        const float w10 = (t.vy1 && t.vx0) ? (1.f - t.wx1) * t.wy1 * t.mask : 0.f;
        const float w11 = (t.vy1 && t.vx1) ? t.wx1 * t.wy1 * t.mask : 0.f;
        const int xa = clampi(t.x0, 0, W - 1), xb = clampi(t.x0 + 1, 0, W - 1);
        const int ya = clampi(t.y0, 0, H - 1), yb = clampi(t.y0 + 1, 0, H - 1);
        const float* ib = xin + (int64_t)b * C * pl;
        float* ob = out + (int64_t)b * C * pl + (int64_t)y * W + x;
        for (int c = threadIdx.y; c < C; c += blockDim.y) {
            const float* ic = ib + c * pl;
            float v = w00 * __ldg(ic + (int64_t)ya * W + xa);
            v += w01 * __ldg(ic + (int64_t)ya * W + xb);
            v += w10 * __ldg(ic + (int64_t)yb * W + xa);
            v += w11 * __ldg(ic + (int64_t)yb * W + xb);
            ob[c * pl] = v;
        }
    }
}

__global__ void pwc_warp_bwd_kernel(const float* __restrict__ xin, const float* __restrict__ flow,
                                    const float* __restrict__ gout, float* __restrict__ gx,
                                    float* __restrict__ gflow, int B, int C, int H, int W) {
    extern __shared__ float red[];                    // [2][blockDim.y][blockDim.x] flow-gradient partials (blockDim.y > 1)
    const int64_t npix = (int64_t)B * H * W, pl = (int64_t)H * W;
    // every thread of a block runs the same number of iterations (the reduction below has block barriers)
    for (int64_t p0 = (int64_t)blockIdx.x * blockDim.x; p0 < npix; p0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = p0 + threadIdx.x;
        const bool live = p < npix;
        float gix = 0.f, giy = 0.f;
        int x = 0, y = 0, b = 0;
        if (live) {
        x = (int)(p % W);
        y = (int)((p / W) % H);
        b = (int)(p / pl);
        const float fx = flow[((int64_t)b * 2 + 0) * pl + (int64_t)y * W + x];
        const float fy = flow[((int64_t)b * 2 + 1) * pl + (int64_t)y * W + x];
        const WarpTaps t = pwc_taps(x, y, fx, fy, H, W);
        const bool v00 = t.vy0 && t.vx0, v01 = t.vy0 && t.vx1, v10 = t.vy1 && t.vx0,
                   v11 = t.vy1 && t.vx1;
        const float ax = 1.f - t.wx1, ay = 1.f - t.wy1;
        const int xa = clampi(t.x0, 0, W - 1), xb = clampi(t.x0 + 1, 0, W - 1);
        const int ya = clampi(t.y0, 0, H - 1), yb = clampi(t.y0 + 1, 0, H - 1);
        const float* ib = xin + (int64_t)b * C * pl;
        float* gb = gx + (int64_t)b * C * pl;
        const float* go = gout + (int64_t)b * C * pl + (int64_t)y * W + x;
        if (t.mask != 0.f) {
            for (int c = threadIdx.y; c < C; c += blockDim.y) {
                const float g = __ldg(go + c * pl);
                const float* ic = ib + c * pl;
                float* gc = gb + c * pl;
                const float a = v00 ? __ldg(ic + (int64_t)ya * W + xa) : 0.f;
                const float bq = v01 ? __ldg(ic + (int64_t)ya * W + xb) : 0.f;
                const float cq = v10 ? __ldg(ic + (int64_t)yb * W + xa) : 0.f;
                const float d = v11 ? __ldg(ic + (int64_t)yb * W + xb) : 0.f;
                if (v00) red_add(gc + (int64_t)ya * W + xa, ax * ay * g);
                if (v01) red_add(gc + (int64_t)ya * W + xb, t.wx1 * ay * g);
                if (v10) red_add(gc + (int64_t)yb * W + xa, ax * t.wy1 * g);
                if (v11) red_add(gc + (int64_t)yb * W + xb, t.wx1 * t.wy1 * g);
                gix += g * (ay * (bq - a) + t.wy1 * (d - cq));
                giy += g * (ax * (cq - a) + t.wx1 * (d - bq));
            }
        }
        }
        if (blockDim.y > 1) {                                           // ordered sum over the channel slices
            float* rx = red + threadIdx.y * blockDim.x + threadIdx.x;
            float* ry = rx + blockDim.y * blockDim.x;
            *rx = gix; *ry = giy;
            __syncthreads();
            if (threadIdx.y == 0) {
                for (int k = 1; k < (int)blockDim.y; ++k) { gix += rx[k * blockDim.x]; giy += ry[k * blockDim.x]; }
            }
            __syncthreads();
        }
        // d ix / d flow_x = (2/max(W-1,1)) * (W/2)
        if (live && threadIdx.y == 0) {
            gflow[((int64_t)b * 2 + 0) * pl + (int64_t)y * W + x] = gix * ((float)W / (float)max(W - 1, 1));
            gflow[((int64_t)b * 2 + 1) * pl + (int64_t)y * W + x] = giy * ((float)H / (float)max(H - 1, 1));
        }
    }
}

static int grid_pix(int64_t npix, int threads) {
    int64_t b = ceil_div<int64_t>(npix, threads);
    const int64_t cap = (int64_t)kNumSMs * 32;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace pcfa

using namespace pcfa;

extern "C" int pcfa_resample2d_forward(const float* img, const float* flow, float* out, int B, int C,
                                       int H, int W, int oH, int oW, int kernel_size, int bilinear,
                                       pcfa_stream_t stream) {
    if (!img || !flow || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || oH <= 0 || oW <= 0)
        return PCFA_E_BADARG;
    if (kernel_size != 1) return PCFA_E_BADARG;
    const int64_t npix = (int64_t)B * oH * oW;
    resample2d_fwd_kernel<<<grid_pix(npix, 128), 128, 0, as_stream(stream)>>>(img, flow, out, B, C, H,
                                                                              W, oH, oW, bilinear);
    return after_launch();
}

extern "C" int pcfa_resample2d_backward(const float* img, const float* flow, const float* grad_out,
                                        float* grad_img, float* grad_flow, int B, int C, int H, int W,
                                        int oH, int oW, int kernel_size, int bilinear,
                                        pcfa_stream_t stream) {
    if (!img || !flow || !grad_out || !grad_img || !grad_flow || B <= 0 || C <= 0 || H <= 0 ||
        W <= 0 || oH <= 0 || oW <= 0)
        return PCFA_E_BADARG;
    if (kernel_size != 1) return PCFA_E_BADARG;
    (void)bilinear;   // the reference's backward kernels ignore it (kernel.cu:75-198)
    const int64_t npix = (int64_t)B * oH * oW;
    resample2d_bwd_kernel<<<grid_pix(npix, 128), 128, 0, as_stream(stream)>>>(
        img, flow, grad_out, grad_img, grad_flow, B, C, H, W, oH, oW);
    return after_launch();
}

extern "C" int pcfa_pwc_warp_forward(const float* x, const float* flow, float* out, int B, int C,
                                     int H, int W, pcfa_stream_t stream) {
    if (!x || !flow || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return PCFA_E_BADARG;
    const int64_t npix = (int64_t)B * H * W;
    const dim3 block = npix >= 32768 ? dim3(128, 1) : dim3(32, 8);
    pwc_warp_fwd_kernel<<<grid_pix(npix, block.x), block, 0, as_stream(stream)>>>(x, flow, out, B, C, H, W);
    return after_launch();
}

extern "C" int pcfa_pwc_warp_backward(const float* x, const float* flow, const float* grad_out,
                                      float* grad_x, float* grad_flow, int B, int C, int H, int W,
                                      pcfa_stream_t stream) {
    if (!x || !flow || !grad_out || !grad_x || !grad_flow || B <= 0 || C <= 0 || H <= 0 || W <= 0)
        return PCFA_E_BADARG;
    const int64_t npix = (int64_t)B * H * W;
    const dim3 block = npix >= 32768 ? dim3(128, 1) : dim3(32, 8);
    pwc_warp_bwd_kernel<<<grid_pix(npix, block.x), block, 2 * block.x * block.y * sizeof(float), as_stream(stream)>>>(
        x, flow, grad_out, grad_x, grad_flow, B, C, H, W);
    return after_launch();
}
