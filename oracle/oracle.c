/*
 * oracle.c — TEST INFRASTRUCTURE ONLY.  Scalar CPU restatement of the reference's algorithms for
 * PCFA's hot-path operators.  Nothing under pcfa_b200/ may link, load or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do, as the checker.
 *
 * Each function follows the cited reference source (paths relative to cv-stuttgart/PCFA) loop for
 * loop; accumulations are done in double where the reference accumulates in float so that the
 * oracle is the more accurate side of every comparison (documented per function).
 *
 * Parity pinning: the reference ships no tests or golden vectors.  This oracle is
 * pinned instead against outputs of the reference code itself executed in the authoring container
 * (oracle/make_golden.py → tests/golden/*.npz) and against the reference's own CPU extension
 * compiled from its sources (oracle/_ref, see oracle/build_ref.py).
 *
 * Build: gcc -O2 -fPIC -shared -o oracle/_build/liboracle.so oracle/oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX4(n, c, y, x, C, H, W) ((((int64_t)(n) * (C) + (c)) * (H) + (y)) * (int64_t)(W) + (x))

/* ------------------------------------------------------------------------------------------
 * Spatial correlation sampler, CPU semantics.
 * Correlation_Module/correlation.cpp:9-37 (correlate_patch), :75-124 (forward).
 * out[n,ph,pw,h,w]; accumulates in double (reference: scalar_t).
 * ------------------------------------------------------------------------------------------ */
void oracle_scs_forward(const float* in1, const float* in2, float* out, int B, int C, int iH, int iW,
                        int kH, int kW, int patchH, int patchW, int padH, int padW, int dilH, int dilW,
                        int dpH, int dpW, int dH, int dW) {
    const int radH = (patchH - 1) / 2, radW = (patchW - 1) / 2;           /* :88-89 */
    const int dkH = (kH - 1) * dilH + 1, dkW = (kW - 1) * dilW + 1;       /* :90-91 */
    const int oH = (iH + 2 * padH - dkH) / dH + 1, oW = (iW + 2 * padW - dkW) / dW + 1; /* :93-94 */
    for (int n = 0; n < B; ++n)
        for (int ph = 0; ph < patchH; ++ph)
            for (int pw = 0; pw < patchW; ++pw)
                for (int h = 0; h < oH; ++h)
                    for (int w = 0; w < oW; ++w) {
                        const int u = -padH + h * dH, v = -padW + w * dW;
                        const int su = (ph - radH) * dpH, sv = (pw - radW) * dpW;
                        double acc = 0.0;
                        for (int c = 0; c < C; ++c)
                            for (int i = 0; i < kH; ++i) {
                                const int i1 = u + i * dilH, i2 = i1 + su;
                                if (!(i1 >= 0 && i1 < iH && i2 >= 0 && i2 < iH)) continue;
                                for (int j = 0; j < kW; ++j) {
                                    const int j1 = v + j * dilW, j2 = j1 + sv;
                                    if (!(j1 >= 0 && j1 < iW && j2 >= 0 && j2 < iW)) continue;
                                    acc += (double)in1[IDX4(n, c, i1, j1, C, iH, iW)] *
                                           (double)in2[IDX4(n, c, i2, j2, C, iH, iW)];
                                }
                            }
                        out[((((int64_t)n * patchH + ph) * patchW + pw) * oH + h) * oW + w] = (float)acc;
                    }
}

/* Correlation_Module/correlation.cpp:40-73 (correlate_patch_grad), :126-178 (backward).
 * Scatter form exactly as the reference; accumulation buffers are double. */
void oracle_scs_backward(const float* in1, const float* in2, const float* gout, float* g1, float* g2,
                         int B, int C, int iH, int iW, int kH, int kW, int patchH, int patchW,
                         int padH, int padW, int dilH, int dilW, int dpH, int dpW, int dH, int dW) {
    const int radH = (patchH - 1) / 2, radW = (patchW - 1) / 2;
    const int dkH = (kH - 1) * dilH + 1, dkW = (kW - 1) * dilW + 1;
    const int oH = (iH + 2 * padH - dkH) / dH + 1, oW = (iW + 2 * padW - dkW) / dW + 1;
    const int64_t numel = (int64_t)B * C * iH * iW;
    double* a1 = (double*)calloc(numel, sizeof(double));
    double* a2 = (double*)calloc(numel, sizeof(double));
    for (int n = 0; n < B; ++n)
        for (int ph = 0; ph < patchH; ++ph)
            for (int pw = 0; pw < patchW; ++pw)
                for (int h = 0; h < oH; ++h)
                    for (int w = 0; w < oW; ++w) {
                        const double go =
                            gout[((((int64_t)n * patchH + ph) * patchW + pw) * oH + h) * oW + w];
                        const int u = -padH + h * dH, v = -padW + w * dW;
                        const int su = (ph - radH) * dpH, sv = (pw - radW) * dpW;
                        for (int c = 0; c < C; ++c)
                            for (int i = 0; i < kH; ++i) {
                                const int i1 = u + i * dilH, i2 = i1 + su;
                                if (!(i1 >= 0 && i1 < iH && i2 >= 0 && i2 < iH)) continue;
                                for (int j = 0; j < kW; ++j) {
                                    const int j1 = v + j * dilW, j2 = j1 + sv;
                                    if (!(j1 >= 0 && j1 < iW && j2 >= 0 && j2 < iW)) continue;
                                    const int64_t p1 = IDX4(n, c, i1, j1, C, iH, iW);
                                    const int64_t p2 = IDX4(n, c, i2, j2, C, iH, iW);
                                    a2[p2] += go * in1[p1];
                                    a1[p1] += go * in2[p2];
                                }
                            }
                    }
    for (int64_t i = 0; i < numel; ++i) { g1[i] = (float)a1[i]; g2[i] = (float)a2[i]; }
    free(a1); free(a2);
}

/* ------------------------------------------------------------------------------------------
 * FlowNet2 correlation.  models/FlowNet/correlation_package/correlation_cuda.cc:25-38 (shapes),
 * correlation_cuda_kernel.cu:46-70 (channels_first: zero-padded NHWC copy), :73-147 (forward),
 * :150-241 / :243-334 (backward input1 / input2, truncating integer division).
 * The padded NHWC buffers are materialised here exactly like the reference does.
 * ------------------------------------------------------------------------------------------ */
static float* fn2_pad_nhwc(const float* in, int B, int C, int H, int W, int pad) {
    const int pH = H + 2 * pad, pW = W + 2 * pad;
    float* r = (float*)calloc((size_t)B * pH * pW * C, sizeof(float));
    for (int n = 0; n < B; ++n)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x)
                    r[(((int64_t)n * pH + (y + pad)) * pW + (x + pad)) * C + c] = in[IDX4(n, c, y, x, C, H, W)];
    return r;
}

void oracle_fn2corr_sizes(int H, int W, int pad, int ks, int md, int s1, int s2, int* outC, int* oH, int* oW) {
    const int kr = (ks - 1) / 2, border = kr + md;
    const int D = (md / s2) * 2 + 1;
    *outC = D * D;
    *oH = (int)ceil((float)(H + 2 * pad - 2 * border) / (float)s1);
    *oW = (int)ceil((float)(W + 2 * pad - 2 * border) / (float)s1);
}

void oracle_fn2corr_forward(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                            int pad, int ks, int md, int s1, int s2) {
    int outC, oH, oW;
    oracle_fn2corr_sizes(H, W, pad, ks, md, s1, s2, &outC, &oH, &oW);
    const int pH = H + 2 * pad, pW = W + 2 * pad;
    const int kr = (ks - 1) / 2, R = md / s2, D = 2 * R + 1;
    float* r1 = fn2_pad_nhwc(in1, B, C, H, W, pad);
    float* r2 = fn2_pad_nhwc(in2, B, C, H, W, pad);
    const double nelems = (double)ks * ks * C;
    for (int n = 0; n < B; ++n)
        for (int y = 0; y < oH; ++y)
            for (int x = 0; x < oW; ++x) {
                const int y1 = y * s1 + md, x1 = x * s1 + md;
                for (int tj = -R; tj <= R; ++tj)
                    for (int ti = -R; ti <= R; ++ti) {
                        const int x2 = x1 + ti * s2, y2 = y1 + tj * s2;
                        double acc = 0.0;
                        for (int j = -kr; j <= kr; ++j)
                            for (int i = -kr; i <= kr; ++i)
                                for (int ch = 0; ch < C; ++ch) {
                                    const int64_t a = (((int64_t)n * pH + (y1 + j)) * pW + (x1 + i)) * C + ch;
                                    const int64_t b = (((int64_t)n * pH + (y2 + j)) * pW + (x2 + i)) * C + ch;
                                    acc += (double)r1[a] * (double)r2[b];
                                }
                        const int tc = (tj + R) * D + (ti + R);
                        out[IDX4(n, tc, y, x, outC, oH, oW)] = (float)(acc / nelems);
                    }
            }
    free(r1); free(r2);
}

void oracle_fn2corr_backward(const float* in1, const float* in2, const float* gout, float* g1, float* g2,
                             int B, int C, int H, int W, int pad, int ks, int md, int s1, int s2) {
    int outC, oH, oW;
    oracle_fn2corr_sizes(H, W, pad, ks, md, s1, s2, &outC, &oH, &oW);
    const int pH = H + 2 * pad, pW = W + 2 * pad;
    const int kr = (ks - 1) / 2, R = md / s2, D = 2 * R + 1;
    float* r1 = fn2_pad_nhwc(in1, B, C, H, W, pad);
    float* r2 = fn2_pad_nhwc(in2, B, C, H, W, pad);
    const double nelems = (double)ks * ks * C;
    memset(g1, 0, sizeof(float) * (size_t)B * C * H * W);
    memset(g2, 0, sizeof(float) * (size_t)B * C * H * W);
    /* grid (H, W, C): y = blockIdx.x*stride1 + pad  (kernel.cu:165-166, 257-258).  Blocks whose
     * target element would fall outside the tensor are skipped (the reference would write out of
     * bounds; FlowNet2 only uses stride1 == 1 where this never happens). */
    for (int n = 0; n < B; ++n)
        for (int by = 0; by < H; ++by)
            for (int bx = 0; bx < W; ++bx) {
                const int y = by * s1 + pad, x = bx * s1 + pad;
                if (y - pad >= H || x - pad >= W) continue;
                for (int c = 0; c < C; ++c) {
                    /* input1 */
                    {
                        int xmin = (x - kr - md) / s1, ymin = (y - kr - md) / s1;
                        int xmax = (x + kr - md) / s1, ymax = (y + kr - md) / s1;
                        if (!(xmax < 0 || ymax < 0 || xmin >= oW || ymin >= oH) && !(xmin > xmax || ymin > ymax)) {
                            if (xmin < 0) xmin = 0; if (xmax > oW - 1) xmax = oW - 1;
                            if (ymin < 0) ymin = 0; if (ymax > oH - 1) ymax = oH - 1;
                            double sum = 0.0;
                            for (int tc = 0; tc < outC; ++tc) {
                                const int i2 = (tc % D - R) * s2, j2 = (tc / D - R) * s2;
                                const int yy = y + j2, xx = x + i2;
                                if (yy < 0 || yy >= pH || xx < 0 || xx >= pW) continue; /* out of buffer */
                                const double v2 = r2[(((int64_t)n * pH + yy) * pW + xx) * C + c];
                                for (int j = ymin; j <= ymax; ++j)
                                    for (int i = xmin; i <= xmax; ++i)
                                        sum += (double)gout[IDX4(n, tc, j, i, outC, oH, oW)] * v2;
                            }
                            g1[IDX4(n, c, y - pad, x - pad, C, H, W)] = (float)(sum / nelems);
                        }
                    }
                    /* input2 */
                    {
                        double sum = 0.0;
                        for (int tc = 0; tc < outC; ++tc) {
                            const int i2 = (tc % D - R) * s2, j2 = (tc / D - R) * s2;
                            int xmin = (x - kr - md - i2) / s1, ymin = (y - kr - md - j2) / s1;
                            int xmax = (x + kr - md - i2) / s1, ymax = (y + kr - md - j2) / s1;
                            if (xmax < 0 || ymax < 0 || xmin >= oW || ymin >= oH) continue;
                            if (xmin > xmax || ymin > ymax) continue;
                            if (xmin < 0) xmin = 0; if (xmax > oW - 1) xmax = oW - 1;
                            if (ymin < 0) ymin = 0; if (ymax > oH - 1) ymax = oH - 1;
                            const int yy = y - j2, xx = x - i2;
                            if (yy < 0 || yy >= pH || xx < 0 || xx >= pW) continue;
                            const double v1 = r1[(((int64_t)n * pH + yy) * pW + xx) * C + c];
                            for (int j = ymin; j <= ymax; ++j)
                                for (int i = xmin; i <= xmax; ++i)
                                    sum += (double)gout[IDX4(n, tc, j, i, outC, oH, oW)] * v1;
                        }
                        g2[IDX4(n, c, y - pad, x - pad, C, H, W)] = (float)(sum / nelems);
                    }
                }
            }
    free(r1); free(r2);
}

/* ------------------------------------------------------------------------------------------
 * Resample2d.  models/FlowNet/resample2d_package/resample2d_kernel.cu:15-72 (forward),
 * :75-125 (backward image, alpha = xf - int(xf)), :127-198 (backward flow).  kernel_size == 1.
 * Arithmetic kept in float like the reference (the weights are float products; the forward
 * promotes (1. - alpha) to double, reproduced here).
 * ------------------------------------------------------------------------------------------ */
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void oracle_resample2d_forward(const float* img, const float* flow, float* out, int B, int C, int H,
                               int W, int oH, int oW, int bilinear) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < oH; ++y)
                for (int x = 0; x < oW; ++x) {
                    const float dx = flow[IDX4(b, 0, y, x, 2, oH, oW)], dy = flow[IDX4(b, 1, y, x, 2, oH, oW)];
                    const float xf = (float)x + dx, yf = (float)y + dy;
                    const float alpha = xf - floorf(xf), beta = yf - floorf(yf);
                    float val = 0.0f;
                    if (bilinear) {
                        const int xL = clampi((int)floorf(xf), 0, W - 1), xR = clampi((int)(floorf(xf) + 1), 0, W - 1);
                        const int yT = clampi((int)floorf(yf), 0, H - 1), yB = clampi((int)(floorf(yf) + 1), 0, H - 1);
                        val += (float)((1. - alpha) * (1. - beta) * img[IDX4(b, c, yT, xL, C, H, W)]);
                        val += (float)((alpha) * (1. - beta) * img[IDX4(b, c, yT, xR, C, H, W)]);
                        val += (float)((1. - alpha) * (beta)*img[IDX4(b, c, yB, xL, C, H, W)]);
                        val += (float)((alpha) * (beta)*img[IDX4(b, c, yB, xR, C, H, W)]);
                    } else {
                        const int xN = clampi((int)floor(xf + 0.5), 0, W - 1), yN = clampi((int)floor(yf + 0.5), 0, H - 1);
                        val = img[IDX4(b, c, yN, xN, C, H, W)];
                    }
                    out[IDX4(b, c, y, x, C, oH, oW)] = val;
                }
}

void oracle_resample2d_backward(const float* img, const float* flow, const float* gout, float* gimg,
                                float* gflow, int B, int C, int H, int W, int oH, int oW) {
    const int64_t n_img = (int64_t)B * C * H * W;
    double* acc = (double*)calloc(n_img, sizeof(double));
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < oH; ++y)
                for (int x = 0; x < oW; ++x) {
                    const float dx = flow[IDX4(b, 0, y, x, 2, oH, oW)], dy = flow[IDX4(b, 1, y, x, 2, oH, oW)];
                    const float xf = (float)x + dx, yf = (float)y + dy;
                    const float alpha = xf - (float)(int)xf, beta = yf - (float)(int)yf;   /* :105-106 */
                    const int xL = clampi((int)floorf(xf), 0, W - 1), xR = clampi((int)(floorf(xf) + 1), 0, W - 1);
                    const int yT = clampi((int)floorf(yf), 0, H - 1), yB = clampi((int)(floorf(yf) + 1), 0, H - 1);
                    const float g = gout[IDX4(b, c, y, x, C, oH, oW)];
                    acc[IDX4(b, c, yT, xL, C, H, W)] += (1 - alpha) * (1 - beta) * g;
                    acc[IDX4(b, c, yT, xR, C, H, W)] += (alpha) * (1 - beta) * g;
                    acc[IDX4(b, c, yB, xL, C, H, W)] += (1 - alpha) * (beta)*g;
                    acc[IDX4(b, c, yB, xR, C, H, W)] += (alpha) * (beta)*g;
                }
    for (int64_t i = 0; i < n_img; ++i) gimg[i] = (float)acc[i];
    free(acc);
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < 2; ++c)
            for (int y = 0; y < oH; ++y)
                for (int x = 0; x < oW; ++x) {
                    const float dx = flow[IDX4(b, 0, y, x, 2, oH, oW)], dy = flow[IDX4(b, 1, y, x, 2, oH, oW)];
                    const float xf = (float)x + dx, yf = (float)y + dy;
                    const int xL = clampi((int)floorf(xf), 0, W - 1), xR = clampi((int)(floorf(xf) + 1), 0, W - 1);
                    const int yT = clampi((int)floorf(yf), 0, H - 1), yB = clampi((int)(floorf(yf) + 1), 0, H - 1);
                    double o = 0.0;
                    if (c % 2) {
                        const float gamma = 1 - (xf - floorf(xf));
                        for (int ch = 0; ch < C; ++ch) {
                            const float g = gout[IDX4(b, ch, y, x, C, oH, oW)];
                            o += (gamma)*g * img[IDX4(b, ch, yB, xL, C, H, W)];
                            o -= (gamma)*g * img[IDX4(b, ch, yT, xL, C, H, W)];
                            o += (1 - gamma) * g * img[IDX4(b, ch, yB, xR, C, H, W)];
                            o -= (1 - gamma) * g * img[IDX4(b, ch, yT, xR, C, H, W)];
                        }
                    } else {
                        const float gamma = 1 - (yf - floorf(yf));
                        for (int ch = 0; ch < C; ++ch) {
                            const float g = gout[IDX4(b, ch, y, x, C, oH, oW)];
                            o += (gamma)*g * img[IDX4(b, ch, yT, xR, C, H, W)];
                            o -= (gamma)*g * img[IDX4(b, ch, yT, xL, C, H, W)];
                            o += (1 - gamma) * g * img[IDX4(b, ch, yB, xR, C, H, W)];
                            o -= (1 - gamma) * g * img[IDX4(b, ch, yB, xL, C, H, W)];
                        }
                    }
                    gflow[IDX4(b, c, y, x, 2, oH, oW)] = (float)o;
                }
}

/* ------------------------------------------------------------------------------------------
 * ChannelNorm.  models/FlowNet/channelnorm_package/channelnorm_kernel.cu:18-60, :63-96.
 * ------------------------------------------------------------------------------------------ */
void oracle_channelnorm_forward(const float* x, float* out, int B, int C, int H, int W) {
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int xx = 0; xx < W; ++xx) {
                float s = 0.f;
                for (int c = 0; c < C; ++c) { const float v = x[IDX4(b, c, y, xx, C, H, W)]; s += v * v; }
                out[((int64_t)b * H + y) * W + xx] = sqrtf(s);
            }
}

void oracle_channelnorm_backward(const float* x, const float* out, const float* gout, float* gx, int B,
                                 int C, int H, int W) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y)
                for (int xx = 0; xx < W; ++xx) {
                    const int64_t o = ((int64_t)b * H + y) * W + xx;
                    gx[IDX4(b, c, y, xx, C, H, W)] = gout[o] * x[IDX4(b, c, y, xx, C, H, W)] / (out[o] + 1e-9f);
                }
}

/* ------------------------------------------------------------------------------------------
 * All-pairs correlation pyramid.  models/raft/corr.py:52-60 (corr: matmul / sqrt(dim)),
 * :25-27 (successive F.avg_pool2d(corr, 2, stride=2), floor sizes).  The matmul and the pooling
 * live in PyTorch (pinned torch==1.7.1, scripts/requirements.txt:2); the published algorithms are
 * restated here with double accumulation.  Flat layout: level l at off[l], [B*N, H_l, W_l].
 * ------------------------------------------------------------------------------------------ */
int64_t oracle_pyramid_layout(int B, int H, int W, int levels, int64_t* off, int* hs, int* ws) {
    int64_t o = 0; int h = H, w = W;
    for (int l = 0; l < levels; ++l) { off[l] = o; hs[l] = h; ws[l] = w; o += (int64_t)B * H * W * h * w; h /= 2; w /= 2; }
    off[levels] = o;
    return o;
}

void oracle_corr_pyramid_forward(const float* f1, const float* f2, float* pyr, int B, int C, int H, int W,
                                 int levels) {
    int64_t off[9]; int hs[8], ws[8];
    oracle_pyramid_layout(B, H, W, levels, off, hs, ws);
    const int N = H * W;
    const float inv = 1.0f / sqrtf((float)C);
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                double acc = 0.0;
                for (int c = 0; c < C; ++c)
                    acc += (double)f1[((int64_t)b * C + c) * N + i] * (double)f2[((int64_t)b * C + c) * N + j];
                pyr[off[0] + ((int64_t)b * N + i) * N + j] = (float)acc * inv;
            }
    for (int l = 1; l < levels; ++l) {
        const int Hi = hs[l - 1], Wi = ws[l - 1], Ho = hs[l], Wo = ws[l];
        for (int64_t r = 0; r < (int64_t)B * N; ++r)
            for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x) {
                    const float* p = pyr + off[l - 1] + (r * Hi + 2 * y) * (int64_t)Wi + 2 * x;
                    pyr[off[l] + (r * Ho + y) * (int64_t)Wo + x] =
                        (float)(((double)p[0] + p[1] + p[Wi] + p[Wi + 1]) * 0.25);
                }
    }
}

/* autograd of the above: avg_pool2d_backward level by level (fold to level 0), then the two matmul
 * gradients.  gpyr holds dL/d(level l) for every level (flat layout). */
void oracle_corr_pyramid_backward(const float* gpyr, const float* f1, const float* f2, float* g1, float* g2,
                                  int B, int C, int H, int W, int levels) {
    int64_t off[9]; int hs[8], ws[8];
    const int64_t total = oracle_pyramid_layout(B, H, W, levels, off, hs, ws);
    const int N = H * W;
    double* G = (double*)malloc(sizeof(double) * total);
    for (int64_t i = 0; i < total; ++i) G[i] = gpyr[i];
    for (int l = levels - 1; l >= 1; --l) {
        const int Hi = hs[l - 1], Wi = ws[l - 1], Ho = hs[l], Wo = ws[l];
        for (int64_t r = 0; r < (int64_t)B * N; ++r)
            for (int y = 0; y < Ho; ++y)
                for (int x = 0; x < Wo; ++x) {
                    const double g = 0.25 * G[off[l] + (r * Ho + y) * (int64_t)Wo + x];
                    double* p = G + off[l - 1] + (r * Hi + 2 * y) * (int64_t)Wi + 2 * x;
                    p[0] += g; p[1] += g; p[Wi] += g; p[Wi + 1] += g;
                }
    }
    const double inv = 1.0 / sqrt((double)C);
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            for (int i = 0; i < N; ++i) {
                double a = 0.0;
                for (int j = 0; j < N; ++j) a += G[((int64_t)b * N + i) * N + j] * f2[((int64_t)b * C + c) * N + j];
                g1[((int64_t)b * C + c) * N + i] = (float)(a * inv);
            }
            for (int j = 0; j < N; ++j) {
                double a = 0.0;
                for (int i = 0; i < N; ++i) a += G[((int64_t)b * N + i) * N + j] * f1[((int64_t)b * C + c) * N + i];
                g2[((int64_t)b * C + c) * N + j] = (float)(a * inv);
            }
        }
    free(G);
}

/* ------------------------------------------------------------------------------------------
 * Multi-level lookup.  models/raft/corr.py:29-50 + models/raft/utils/utils.py:57-71
 * (bilinear_sampler → F.grid_sample(align_corners=True), zeros padding).  The normalise /
 * un-normalise round trip of grid_sample is reproduced in float:
 *     g = 2*s/(W-1) - 1 ;  s' = ((g + 1) / 2) * (W-1)
 * window: delta = stack(meshgrid(dy, dx), -1) added to (x, y) → first window index shifts x.
 * ------------------------------------------------------------------------------------------ */
static void lookup_sample_pos(float c, int level, int k, int r, int size, float* pos) {
    const float s = c / (float)(1 << level) + (float)(k - r);     /* centroid_lvl + delta_lvl */
    const float g = 2.0f * s / (float)(size - 1) - 1.0f;          /* utils.py:61-62 */
    *pos = ((g + 1.0f) / 2.0f) * (float)(size - 1);               /* grid_sampler_unnormalize, align_corners */
}

void oracle_corr_lookup_forward(const float* pyr, const float* coords, float* out, int B, int H, int W,
                                int levels, int r) {
    int64_t off[9]; int hs[8], ws[8];
    oracle_pyramid_layout(B, H, W, levels, off, hs, ws);
    const int N = H * W, D = 2 * r + 1;
    for (int b = 0; b < B; ++b)
        for (int q = 0; q < N; ++q) {
            const float cx = coords[((int64_t)b * 2 + 0) * N + q], cy = coords[((int64_t)b * 2 + 1) * N + q];
            for (int l = 0; l < levels; ++l) {
                const int Hl = hs[l], Wl = ws[l];
                const float* img = pyr + off[l] + ((int64_t)b * N + q) * Hl * Wl;
                for (int a = 0; a < D; ++a)
                    for (int bb = 0; bb < D; ++bb) {
                        float sx, sy;
                        lookup_sample_pos(cx, l, a, r, Wl, &sx);
                        lookup_sample_pos(cy, l, bb, r, Hl, &sy);
                        const float fx0 = floorf(sx), fy0 = floorf(sy);
                        const int x0 = (int)fx0, y0 = (int)fy0;
                        const float wx = sx - fx0, wy = sy - fy0;
                        double v = 0.0;
                        for (int dy = 0; dy < 2; ++dy)
                            for (int dx = 0; dx < 2; ++dx) {
                                const int yy = y0 + dy, xx = x0 + dx;
                                if (yy < 0 || yy >= Hl || xx < 0 || xx >= Wl) continue;
                                const float wgt = (dx ? wx : 1.f - wx) * (dy ? wy : 1.f - wy);
                                v += (double)wgt * img[(int64_t)yy * Wl + xx];
                            }
                        out[((int64_t)b * (levels * D * D) + l * D * D + a * D + bb) * N + q] = (float)v;
                    }
            }
        }
}

/* grid_sampler_2d_backward w.r.t. the input only (coords are detached, raft.py:123);
 * gpyr must be zero-initialised by the caller or hold a running sum (accumulates). */
void oracle_corr_lookup_backward(const float* gout, const float* coords, float* gpyr, int B, int H, int W,
                                 int levels, int r) {
    int64_t off[9]; int hs[8], ws[8];
    oracle_pyramid_layout(B, H, W, levels, off, hs, ws);
    const int N = H * W, D = 2 * r + 1;
    for (int b = 0; b < B; ++b)
        for (int q = 0; q < N; ++q) {
            const float cx = coords[((int64_t)b * 2 + 0) * N + q], cy = coords[((int64_t)b * 2 + 1) * N + q];
            for (int l = 0; l < levels; ++l) {
                const int Hl = hs[l], Wl = ws[l];
                float* gimg = gpyr + off[l] + ((int64_t)b * N + q) * Hl * Wl;
                for (int a = 0; a < D; ++a)
                    for (int bb = 0; bb < D; ++bb) {
                        float sx, sy;
                        lookup_sample_pos(cx, l, a, r, Wl, &sx);
                        lookup_sample_pos(cy, l, bb, r, Hl, &sy);
                        const float fx0 = floorf(sx), fy0 = floorf(sy);
                        const int x0 = (int)fx0, y0 = (int)fy0;
                        const float wx = sx - fx0, wy = sy - fy0;
                        const float g = gout[((int64_t)b * (levels * D * D) + l * D * D + a * D + bb) * N + q];
                        for (int dy = 0; dy < 2; ++dy)
                            for (int dx = 0; dx < 2; ++dx) {
                                const int yy = y0 + dy, xx = x0 + dx;
                                if (yy < 0 || yy >= Hl || xx < 0 || xx >= Wl) continue;
                                gimg[(int64_t)yy * Wl + xx] += (dx ? wx : 1.f - wx) * (dy ? wy : 1.f - wy) * g;
                            }
                    }
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * PWCNet warp.  models/PWCNet/PWCNet.py:166-206: vgrid normalisation (:189-190), two
 * F.grid_sample calls (bilinear, zeros, align_corners=False — the torch default since 1.3),
 * mask = (grid_sample(ones) >= 0.0001), output*mask.  Backward = grid_sampler_2d_backward for x and
 * the grid, chained through the normalisation; the mask carries no gradient.
 * ------------------------------------------------------------------------------------------ */
static void pwc_pos(int x, float f, int size, float* pos) {
    const float g = 2.0f * ((float)x + f) / (float)(size - 1 > 1 ? size - 1 : 1) - 1.0f;
    *pos = ((g + 1.f) * (float)size - 1.f) / 2.f;                 /* unnormalize, align_corners=False */
}

void oracle_pwc_warp_forward(const float* xin, const float* flow, float* out, int B, int C, int H, int W) {
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float ix, iy;
                pwc_pos(x, flow[IDX4(b, 0, y, x, 2, H, W)], W, &ix);
                pwc_pos(y, flow[IDX4(b, 1, y, x, 2, H, W)], H, &iy);
                const float fx0 = floorf(ix), fy0 = floorf(iy);
                const int x0 = (int)fx0, y0 = (int)fy0;
                const float wx = ix - fx0, wy = iy - fy0;
                float m = 0.f;
                for (int dy = 0; dy < 2; ++dy)
                    for (int dx = 0; dx < 2; ++dx) {
                        const int yy = y0 + dy, xx = x0 + dx;
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                        m += (dx ? wx : 1.f - wx) * (dy ? wy : 1.f - wy);
                    }
                const float mask = (m >= 0.0001f) ? 1.f : 0.f;
                for (int c = 0; c < C; ++c) {
                    float v = 0.f;
                    for (int dy = 0; dy < 2; ++dy)
                        for (int dx = 0; dx < 2; ++dx) {
                            const int yy = y0 + dy, xx = x0 + dx;
                            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                            v += (dx ? wx : 1.f - wx) * (dy ? wy : 1.f - wy) * xin[IDX4(b, c, yy, xx, C, H, W)];
                        }
                    out[IDX4(b, c, y, x, C, H, W)] = v * mask;
                }
            }
}

void oracle_pwc_warp_backward(const float* xin, const float* flow, const float* gout, float* gx,
                              float* gflow, int B, int C, int H, int W) {
    const int64_t n = (int64_t)B * C * H * W;
    double* acc = (double*)calloc(n, sizeof(double));
    for (int b = 0; b < B; ++b)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float ix, iy;
                pwc_pos(x, flow[IDX4(b, 0, y, x, 2, H, W)], W, &ix);
                pwc_pos(y, flow[IDX4(b, 1, y, x, 2, H, W)], H, &iy);
                const float fx0 = floorf(ix), fy0 = floorf(iy);
                const int x0 = (int)fx0, y0 = (int)fy0;
                const float wx = ix - fx0, wy = iy - fy0;
                float m = 0.f;
                for (int dy = 0; dy < 2; ++dy)
                    for (int dx = 0; dx < 2; ++dx) {
                        const int yy = y0 + dy, xx = x0 + dx;
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                        m += (dx ? wx : 1.f - wx) * (dy ? wy : 1.f - wy);
                    }
                const float mask = (m >= 0.0001f) ? 1.f : 0.f;
                double gix = 0.0, giy = 0.0;
                for (int c = 0; c < C; ++c) {
                    const float g = gout[IDX4(b, c, y, x, C, H, W)] * mask;
                    for (int dy = 0; dy < 2; ++dy)
                        for (int dx = 0; dx < 2; ++dx) {
                            const int yy = y0 + dy, xx = x0 + dx;
                            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                            const float v = xin[IDX4(b, c, yy, xx, C, H, W)];
                            acc[IDX4(b, c, yy, xx, C, H, W)] += (double)((dx ? wx : 1.f - wx) * (dy ? wy : 1.f - wy)) * g;
                            gix += (double)g * v * (dx ? 1.f : -1.f) * (dy ? wy : 1.f - wy);
                            giy += (double)g * v * (dy ? 1.f : -1.f) * (dx ? wx : 1.f - wx);
                        }
                }
                /* d ix / d flow_x = (W/2) * 2/max(W-1,1) */
                gflow[IDX4(b, 0, y, x, 2, H, W)] = (float)(gix * ((double)W / (double)(W - 1 > 1 ? W - 1 : 1)));
                gflow[IDX4(b, 1, y, x, 2, H, W)] = (float)(giy * ((double)H / (double)(H - 1 > 1 ? H - 1 : 1)));
            }
    for (int64_t i = 0; i < n; ++i) gx[i] = (float)acc[i];
    free(acc);
}
