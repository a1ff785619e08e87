// Element-wise halves of RAFT/GMA's convolutional GRU (models/raft/update.py:16-60, SepConvGRU / ConvGRU):
//     z = sigmoid(convz(hx)); r = sigmoid(convr(hx)); q = tanh(convq([r*h, x])); h' = (1-z)*h + z*q
// The convolutions stay in cuDNN (convz and convr share their input and run as ONE convolution with concatenated
// output channels); the eight element-wise ATen launches per GRU step forward (and ~ten backward) become two each.
// At 55x128 features every launch is ~4 us of latency for 3.6 MB of data, and RAFT runs 24 GRU steps per closure, so
// the launch count, not the bytes, is what these kernels remove (row f-4 of SURVEY.md section 8).
//   gates : zr = [B][2C][HW] pre-activations (z first), h = [B][C][HW]  ->  z, r, rh = r*h
//   blend : z, qc (pre-activation), h                                   ->  q = tanh(qc), h' = (1-z)*h + z*q
#include "common.cuh"
#include <math.h>

namespace pcfa {

constexpr int GRU_THREADS = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename F>
__device__ __forceinline__ void gru_for_each4(int64_t n, int vec, F f) {
    if (vec) {
        for (int64_t i = 4 * ((int64_t)blockIdx.x * GRU_THREADS + threadIdx.x); i < n; i += 4 * (int64_t)gridDim.x * GRU_THREADS) f(i, 4);
    } else {
        for (int64_t i = (int64_t)blockIdx.x * GRU_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * GRU_THREADS) f(i, 1);
    }
}

__global__ void __launch_bounds__(GRU_THREADS)
gru_gates_fwd_kernel(const float* __restrict__ zr, const float* __restrict__ h, float* __restrict__ z, float* __restrict__ r,
                     float* __restrict__ rh, int64_t n, int vec, int cl_C) {
    // NCHW (cl_C == 0): sample b = blockIdx.y, z pre-activations at zr[b][0:n], r at zr[b][n:2n].
    // channels-last (cl_C = C): one flat range of B*n elements, pixel p = i / C: z at zr[p*2C + c], r at + C.
    const int64_t b = blockIdx.y;
    const float* zc = zr + b * 2 * n; const float* rc = zc + n;
    const float* hb = h + b * n; float* zb = z + b * n; float* rb = r + b * n; float* rhb = rh + b * n;
    gru_for_each4(n, vec, [&](int64_t i, int w) {
        int64_t zi = i;
        if (cl_C) { const int64_t p = i / cl_C; zi = p * 2 * cl_C + (i - p * cl_C); }
        const float* zc_ = zc + zi; const float* rc_ = cl_C ? zc_ + cl_C : rc + i;
        if (w == 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(zc_)), c = __ldg(reinterpret_cast<const float4*>(rc_));
            const float4 hv = __ldg(reinterpret_cast<const float4*>(hb + i));
            const float4 zz = make_float4(sigmoidf_(a.x), sigmoidf_(a.y), sigmoidf_(a.z), sigmoidf_(a.w));
            const float4 rr = make_float4(sigmoidf_(c.x), sigmoidf_(c.y), sigmoidf_(c.z), sigmoidf_(c.w));
            *reinterpret_cast<float4*>(zb + i) = zz;
            *reinterpret_cast<float4*>(rb + i) = rr;
            *reinterpret_cast<float4*>(rhb + i) = make_float4(rr.x * hv.x, rr.y * hv.y, rr.z * hv.z, rr.w * hv.w);
        } else {
            const float zz = sigmoidf_(zc_[0]), rr = sigmoidf_(rc_[0]);
            zb[i] = zz; rb[i] = rr; rhb[i] = rr * hb[i];
        }
    });
}

// dz, drh (either may be NULL = zero)  ->  dzr = [B][2C][HW] (dz*z*(1-z) | drh*h*r*(1-r)),  dh = drh * r
__global__ void __launch_bounds__(GRU_THREADS)
gru_gates_bwd_kernel(const float* __restrict__ z, const float* __restrict__ r, const float* __restrict__ h,
                     const float* __restrict__ dz, const float* __restrict__ drh, float* __restrict__ dzr,
                     float* __restrict__ dh, int64_t n, int vec, int cl_C) {
    const int64_t b = blockIdx.y;
    const float* zb = z + b * n; const float* rb = r + b * n; const float* hb = h + b * n;
    const float* dzb = dz ? dz + b * n : nullptr; const float* drhb = drh ? drh + b * n : nullptr;
    float* dzc = dzr + b * 2 * n; float* drc = dzc + n; float* dhb = dh + b * n;
    gru_for_each4(n, vec, [&](int64_t i, int w) {
        int64_t zi = i;
        if (cl_C) { const int64_t p = i / cl_C; zi = p * 2 * cl_C + (i - p * cl_C); }
        float* dz_ = dzc + zi; float* dr_ = cl_C ? dz_ + cl_C : drc + i;
        for (int k = 0; k < w; ++k) {
            const float zz = zb[i + k], rr = rb[i + k], hv = hb[i + k];
            const float gz = dzb ? dzb[i + k] : 0.f, grh = drhb ? drhb[i + k] : 0.f;
            dz_[k] = gz * zz * (1.f - zz);
            dr_[k] = grh * hv * rr * (1.f - rr);
            dhb[i + k] = grh * rr;
        }
    });
}

__global__ void __launch_bounds__(GRU_THREADS)
gru_blend_fwd_kernel(const float* __restrict__ z, const float* __restrict__ qc, const float* __restrict__ h,
                     float* __restrict__ q, float* __restrict__ hn, int64_t n, int vec) {
    gru_for_each4(n, vec, [&](int64_t i, int w) {
        for (int k = 0; k < w; ++k) {
            const float zz = z[i + k], qq = tanhf(qc[i + k]), hv = h[i + k];
            q[i + k] = qq;
            hn[i + k] = (1.f - zz) * hv + zz * qq;
        }
    });
}

// dhn -> dz = dhn*(q-h), dqc = dhn*z*(1-q^2), dh = dhn*(1-z)
__global__ void __launch_bounds__(GRU_THREADS)
gru_blend_bwd_kernel(const float* __restrict__ z, const float* __restrict__ q, const float* __restrict__ h,
                     const float* __restrict__ dhn, float* __restrict__ dz, float* __restrict__ dqc, float* __restrict__ dh,
                     int64_t n, int vec) {
    gru_for_each4(n, vec, [&](int64_t i, int w) {
        for (int k = 0; k < w; ++k) {
            const float zz = z[i + k], qq = q[i + k], hv = h[i + k], g = dhn[i + k];
            dz[i + k] = g * (qq - hv);
            dqc[i + k] = g * zz * (1.f - qq * qq);
            dh[i + k] = g * (1.f - zz);
        }
    });
}

// Channel concatenation of channels-last tensors: out[p][0:C0 | C0:C0+C1 | ...] = in_k[p][:].  ATen's cat takes a
// non-vectorised path for torch.channels_last inputs (15 us instead of 5 us per call at 55x128); the update block
// concatenates seven times per GRU iteration.  All channel counts must be multiples of 4 for the 128-bit path.
struct CatArgs { const float* in[4]; int c[4]; int n; int ctot; };
__global__ void __launch_bounds__(GRU_THREADS)
cat_cl_kernel(const CatArgs a, float* __restrict__ out, int64_t npix, int vec) {
    const int w = vec ? 4 : 1;
    const int64_t per_pix = a.ctot / w, total = npix * per_pix;
    for (int64_t e = (int64_t)blockIdx.x * GRU_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GRU_THREADS) {
        const int64_t p = e / per_pix;
        int c = (int)(e - p * per_pix) * w, k = 0;
        while (k < a.n - 1 && c >= a.c[k]) { c -= a.c[k]; ++k; }
        const float* src = a.in[k] + p * a.c[k] + c;
        float* dst = out + e * w;
        if (vec) *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(src));
        else *dst = __ldg(src);
    }
}

// Same with arbitrary channel counts per input and zero channels appended up to `cpad` (a multiple of 4): every output
// float4 is assembled from scalar loads (each element may come from a different input) and stored with one 128-bit store.
// Used where the consumer is a convolution whose input-channel count cuDNN would otherwise pad itself (PWCNet's and
// FlowNet2's 213-, 473-, 1026-channel concatenations: nhwcAddPaddingKernel was 0.9 / 1.4 ms of their closures).
__global__ void __launch_bounds__(GRU_THREADS)
cat_cl_pad_kernel(const CatArgs a, float* __restrict__ out, int64_t npix, int cpad) {
    const int per_pix = cpad >> 2;
    const int64_t total = npix * per_pix;
    for (int64_t e = (int64_t)blockIdx.x * GRU_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GRU_THREADS) {
        const int64_t p = e / per_pix;
        const int c0 = (int)(e - p * per_pix) * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = c0 + j, k = 0;
            v[j] = 0.f;
            if (c < a.ctot) {
                while (k < a.n - 1 && c >= a.c[k]) { c -= a.c[k]; ++k; }
                v[j] = __ldg(a.in[k] + p * a.c[k] + c);
            }
        }
        reinterpret_cast<float4*>(out)[e] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// ------------------------------------------------------------------------------------ channels-last "x" variants
// The NHWC update block hoists the iteration-invariant third of every GRU convolution (the context features `inp`)
// out of the loop: P = conv(inp, W[:, inp slice]) + bias is computed once and enters here as an addend of the
// pre-activations.  The kernels also write the next convolution's concatenated input directly — [r*h | m] from the
// gates, [h' | m] from the blend (m = motion features, Cm channels) — so the two torch.cat launches per GRU step
// disappear.  Pixel-major indexing: item = (pixel, channel quad) over C + Cm channels; quads >= C/4 copy m.
struct GruX {
    int C, Cm;
    int64_t npix;
};

__global__ void __launch_bounds__(GRU_THREADS)
gru_gates_x_fwd_kernel(const float* __restrict__ zr, const float* __restrict__ P, const float* __restrict__ h,
                       const float* __restrict__ m, float* __restrict__ z, float* __restrict__ r, float* __restrict__ rhm,
                       GruX g) {
    const int Q = (g.C + g.Cm) >> 2, QC = g.C >> 2;
    const int64_t total = g.npix * Q;
    for (int64_t e = (int64_t)blockIdx.x * GRU_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GRU_THREADS) {
        const int64_t p = e / Q;
        const int qd = (int)(e - p * Q);
        float4* dst = reinterpret_cast<float4*>(rhm + p * (g.C + g.Cm)) + qd;
        if (qd >= QC) { *dst = __ldg(reinterpret_cast<const float4*>(m + p * g.Cm) + (qd - QC)); continue; }
        const float4 a = __ldg(reinterpret_cast<const float4*>(zr + p * 2 * g.C) + qd), pa = __ldg(reinterpret_cast<const float4*>(P + p * 2 * g.C) + qd);
        const float4 c = __ldg(reinterpret_cast<const float4*>(zr + p * 2 * g.C + g.C) + qd), pc = __ldg(reinterpret_cast<const float4*>(P + p * 2 * g.C + g.C) + qd);
        const float4 hv = __ldg(reinterpret_cast<const float4*>(h + p * g.C) + qd);
        const float4 zz = make_float4(sigmoidf_(a.x + pa.x), sigmoidf_(a.y + pa.y), sigmoidf_(a.z + pa.z), sigmoidf_(a.w + pa.w));
        const float4 rr = make_float4(sigmoidf_(c.x + pc.x), sigmoidf_(c.y + pc.y), sigmoidf_(c.z + pc.z), sigmoidf_(c.w + pc.w));
        reinterpret_cast<float4*>(z + p * g.C)[qd] = zz;
        reinterpret_cast<float4*>(r + p * g.C)[qd] = rr;
        *dst = make_float4(rr.x * hv.x, rr.y * hv.y, rr.z * hv.z, rr.w * hv.w);
    }
}

// grad_z [npix][C] (may be NULL), grad_rhm [npix][C+Cm] (may be NULL; its tail is the caller's grad of m, a view)
__global__ void __launch_bounds__(GRU_THREADS)
gru_gates_x_bwd_kernel(const float* __restrict__ z, const float* __restrict__ r, const float* __restrict__ h,
                       const float* __restrict__ gz, const float* __restrict__ grhm, float* __restrict__ gzr,
                       float* __restrict__ gh, GruX g) {
    const int QC = g.C >> 2;
    const int64_t total = g.npix * QC;
    for (int64_t e = (int64_t)blockIdx.x * GRU_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GRU_THREADS) {
        const int64_t p = e / QC;
        const int qd = (int)(e - p * QC);
        const float4 zz = __ldg(reinterpret_cast<const float4*>(z + p * g.C) + qd), rr = __ldg(reinterpret_cast<const float4*>(r + p * g.C) + qd);
        const float4 hv = __ldg(reinterpret_cast<const float4*>(h + p * g.C) + qd);
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 a = gz ? __ldg(reinterpret_cast<const float4*>(gz + p * g.C) + qd) : zero;
        const float4 b = grhm ? __ldg(reinterpret_cast<const float4*>(grhm + p * (g.C + g.Cm)) + qd) : zero;
        reinterpret_cast<float4*>(gzr + p * 2 * g.C)[qd] =
            make_float4(a.x * zz.x * (1.f - zz.x), a.y * zz.y * (1.f - zz.y), a.z * zz.z * (1.f - zz.z), a.w * zz.w * (1.f - zz.w));
        reinterpret_cast<float4*>(gzr + p * 2 * g.C + g.C)[qd] =
            make_float4(b.x * hv.x * rr.x * (1.f - rr.x), b.y * hv.y * rr.y * (1.f - rr.y), b.z * hv.z * rr.z * (1.f - rr.z),
                        b.w * hv.w * rr.w * (1.f - rr.w));
        reinterpret_cast<float4*>(gh + p * g.C)[qd] = make_float4(b.x * rr.x, b.y * rr.y, b.z * rr.z, b.w * rr.w);
    }
}

// hm (may be NULL): [npix][C+Cm] = [h_new | m]
__global__ void __launch_bounds__(GRU_THREADS)
gru_blend_x_fwd_kernel(const float* __restrict__ z, const float* __restrict__ qc, const float* __restrict__ P,
                       const float* __restrict__ h, const float* __restrict__ m, float* __restrict__ q,
                       float* __restrict__ hn, float* __restrict__ hm, GruX g) {
    const int QC = g.C >> 2, Q = hm ? (g.C + g.Cm) >> 2 : QC;
    const int64_t total = g.npix * Q;
    for (int64_t e = (int64_t)blockIdx.x * GRU_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GRU_THREADS) {
        const int64_t p = e / Q;
        const int qd = (int)(e - p * Q);
        if (qd >= QC) {
            reinterpret_cast<float4*>(hm + p * (g.C + g.Cm))[qd] = __ldg(reinterpret_cast<const float4*>(m + p * g.Cm) + (qd - QC));
            continue;
        }
        const float4 zz = __ldg(reinterpret_cast<const float4*>(z + p * g.C) + qd), a = __ldg(reinterpret_cast<const float4*>(qc + p * g.C) + qd);
        const float4 pa = __ldg(reinterpret_cast<const float4*>(P + p * g.C) + qd), hv = __ldg(reinterpret_cast<const float4*>(h + p * g.C) + qd);
        const float4 qq = make_float4(tanhf(a.x + pa.x), tanhf(a.y + pa.y), tanhf(a.z + pa.z), tanhf(a.w + pa.w));
        const float4 o = make_float4((1.f - zz.x) * hv.x + zz.x * qq.x, (1.f - zz.y) * hv.y + zz.y * qq.y,
                                     (1.f - zz.z) * hv.z + zz.z * qq.z, (1.f - zz.w) * hv.w + zz.w * qq.w);
        reinterpret_cast<float4*>(q + p * g.C)[qd] = qq;
        reinterpret_cast<float4*>(hn + p * g.C)[qd] = o;
        if (hm) reinterpret_cast<float4*>(hm + p * (g.C + g.Cm))[qd] = o;
    }
}

// total grad of h_new = ghn (may be NULL) + ghm[:, :C] (may be NULL)
__global__ void __launch_bounds__(GRU_THREADS)
gru_blend_x_bwd_kernel(const float* __restrict__ z, const float* __restrict__ q, const float* __restrict__ h,
                       const float* __restrict__ ghn, const float* __restrict__ ghm, float* __restrict__ gz,
                       float* __restrict__ gq, float* __restrict__ gh, GruX g) {
    const int QC = g.C >> 2;
    const int64_t total = g.npix * QC;
    for (int64_t e = (int64_t)blockIdx.x * GRU_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GRU_THREADS) {
        const int64_t p = e / QC;
        const int qd = (int)(e - p * QC);
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 a = ghn ? __ldg(reinterpret_cast<const float4*>(ghn + p * g.C) + qd) : zero;
        const float4 b = ghm ? __ldg(reinterpret_cast<const float4*>(ghm + p * (g.C + g.Cm)) + qd) : zero;
        const float4 d = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        const float4 zz = __ldg(reinterpret_cast<const float4*>(z + p * g.C) + qd), qq = __ldg(reinterpret_cast<const float4*>(q + p * g.C) + qd);
        const float4 hv = __ldg(reinterpret_cast<const float4*>(h + p * g.C) + qd);
        reinterpret_cast<float4*>(gz + p * g.C)[qd] = make_float4(d.x * (qq.x - hv.x), d.y * (qq.y - hv.y), d.z * (qq.z - hv.z), d.w * (qq.w - hv.w));
        reinterpret_cast<float4*>(gq + p * g.C)[qd] = make_float4(d.x * zz.x * (1.f - qq.x * qq.x), d.y * zz.y * (1.f - qq.y * qq.y),
                                                                  d.z * zz.z * (1.f - qq.z * qq.z), d.w * zz.w * (1.f - qq.w * qq.w));
        reinterpret_cast<float4*>(gh + p * g.C)[qd] = make_float4(d.x * (1.f - zz.x), d.y * (1.f - zz.y), d.z * (1.f - zz.z), d.w * (1.f - zz.w));
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Whole-GRU-step backward helpers (pcfa_b200/gru_ops.py::gru_step_x): the two half steps of SepConvGRU share h and the
// motion features m between five consumers each, and autograd sums their gradients one strided ATen add at a time
// (~14 launches, 6 us each, per GRU iteration; 1.0 ms of the 8.6 ms RAFT closure).  With one autograd node per step
// the sums happen inside these kernels and every gradient is written once.
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// acc_mode: 0 no accumulator, 1 acc = value, 2 acc += value (gradient of the hoisted addend, summed over iterations)
__device__ __forceinline__ void acc4(float* acc, int mode, float4 v) {
    if (mode == 1) *reinterpret_cast<float4*>(acc) = v;
    else if (mode == 2) { float4* a = reinterpret_cast<float4*>(acc); *a = f4add(*a, v); }
}

// like gru_gates_x_bwd_kernel, plus the addend accumulator acc [npix][2C]
__global__ void __launch_bounds__(GRU_THREADS)
gru_gates_x_bwd2_kernel(const float* __restrict__ z, const float* __restrict__ r, const float* __restrict__ h,
                        const float* __restrict__ gz, const float* __restrict__ grhm, float* __restrict__ gzr,
                        float* __restrict__ gh, float* __restrict__ acc, int acc_mode, GruX g) {
    const int QC = g.C >> 2;
    const int64_t total = g.npix * QC;
    for (int64_t e = (int64_t)blockIdx.x * GRU_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GRU_THREADS) {
        const int64_t p = e / QC;
        const int qd = (int)(e - p * QC);
        const float4 zz = ld4(z + p * g.C + 4 * qd), rr = ld4(r + p * g.C + 4 * qd), hv = ld4(h + p * g.C + 4 * qd);
        const float4 a = ld4(gz + p * g.C + 4 * qd), b = ld4(grhm + p * (g.C + g.Cm) + 4 * qd);
        const float4 o1 = make_float4(a.x * zz.x * (1.f - zz.x), a.y * zz.y * (1.f - zz.y), a.z * zz.z * (1.f - zz.z), a.w * zz.w * (1.f - zz.w));
        const float4 o2 = make_float4(b.x * hv.x * rr.x * (1.f - rr.x), b.y * hv.y * rr.y * (1.f - rr.y), b.z * hv.z * rr.z * (1.f - rr.z),
                                      b.w * hv.w * rr.w * (1.f - rr.w));
        reinterpret_cast<float4*>(gzr + p * 2 * g.C)[qd] = o1;
        reinterpret_cast<float4*>(gzr + p * 2 * g.C + g.C)[qd] = o2;
        acc4(acc + p * 2 * g.C + 4 * qd, acc_mode, o1);
        acc4(acc + p * 2 * g.C + g.C + 4 * qd, acc_mode, o2);
        reinterpret_cast<float4*>(gh + p * g.C)[qd] = make_float4(b.x * rr.x, b.y * rr.y, b.z * rr.z, b.w * rr.w);
    }
}

// total grad of h_new = ghn_a + ghn_b (may be NULL) + ghm[:, :C] (may be NULL); acc [npix][C] accumulates grad_q_pre
__global__ void __launch_bounds__(GRU_THREADS)
gru_blend_x_bwd2_kernel(const float* __restrict__ z, const float* __restrict__ q, const float* __restrict__ h,
                        const float* __restrict__ ghn_a, const float* __restrict__ ghn_b, const float* __restrict__ ghm,
                        float* __restrict__ gz, float* __restrict__ gq, float* __restrict__ gh, float* __restrict__ acc, int acc_mode,
                        GruX g) {
    const int QC = g.C >> 2;
    const int64_t total = g.npix * QC;
    for (int64_t e = (int64_t)blockIdx.x * GRU_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GRU_THREADS) {
        const int64_t p = e / QC;
        const int qd = (int)(e - p * QC);
        float4 d = ld4(ghn_a + p * g.C + 4 * qd);
        if (ghn_b) d = f4add(d, ld4(ghn_b + p * g.C + 4 * qd));
        if (ghm) d = f4add(d, ld4(ghm + p * (g.C + g.Cm) + 4 * qd));
        const float4 zz = ld4(z + p * g.C + 4 * qd), qq = ld4(q + p * g.C + 4 * qd), hv = ld4(h + p * g.C + 4 * qd);
        reinterpret_cast<float4*>(gz + p * g.C)[qd] = make_float4(d.x * (qq.x - hv.x), d.y * (qq.y - hv.y), d.z * (qq.z - hv.z), d.w * (qq.w - hv.w));
        const float4 o = make_float4(d.x * zz.x * (1.f - qq.x * qq.x), d.y * zz.y * (1.f - qq.y * qq.y),
                                     d.z * zz.z * (1.f - qq.z * qq.z), d.w * zz.w * (1.f - qq.w * qq.w));
        reinterpret_cast<float4*>(gq + p * g.C)[qd] = o;
        acc4(acc + p * g.C + 4 * qd, acc_mode, o);
        reinterpret_cast<float4*>(gh + p * g.C)[qd] = make_float4(d.x * (1.f - zz.x), d.y * (1.f - zz.y), d.z * (1.f - zz.z), d.w * (1.f - zz.w));
    }
}

// grad_h = gh_a + gh_b + cat0[:, :C];  grad_m = cat0[:, C:] + cat1[:, C:] + cat2[:, C:] + cat3[:, C:]
// (cat_k are the [npix][C+Cm] gradients of the concatenated convolution inputs of the step)
__global__ void __launch_bounds__(GRU_THREADS)
gru_step_combine_kernel(const float* __restrict__ gh_a, const float* __restrict__ gh_b, const float* __restrict__ cat0,
                        const float* __restrict__ cat1, const float* __restrict__ cat2, const float* __restrict__ cat3,
                        float* __restrict__ gh, float* __restrict__ gm, GruX g) {
    const int QC = g.C >> 2, Q = (g.C + g.Cm) >> 2, CT = g.C + g.Cm;
    const int64_t total = g.npix * Q;
    for (int64_t e = (int64_t)blockIdx.x * GRU_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GRU_THREADS) {
        const int64_t p = e / Q;
        const int qd = (int)(e - p * Q);
        const float4 c0 = ld4(cat0 + p * CT + 4 * qd);
        if (qd < QC) {
            reinterpret_cast<float4*>(gh + p * g.C)[qd] = f4add(f4add(ld4(gh_a + p * g.C + 4 * qd), ld4(gh_b + p * g.C + 4 * qd)), c0);
        } else {
            const float4 s = f4add(f4add(c0, ld4(cat1 + p * CT + 4 * qd)), f4add(ld4(cat2 + p * CT + 4 * qd), ld4(cat3 + p * CT + 4 * qd)));
            reinterpret_cast<float4*>(gm + p * g.Cm)[qd - QC] = s;
        }
    }
}

static int gru_grid(int64_t n, int vec) {
    const int64_t per = (int64_t)GRU_THREADS * (vec ? 4 : 1);
    int64_t b = (n + per - 1) / per;
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
static int aligned16(std::initializer_list<const void*> ps) {
    uintptr_t a = 0;
    for (const void* p : ps) a |= reinterpret_cast<uintptr_t>(p);
    return (a & 15) == 0;
}

}  // namespace pcfa

using namespace pcfa;

extern "C" int pcfa_gru_gates_forward(const float* zr, const float* h, float* z, float* r, float* rh, int B, int64_t n,
                                      int channels_last_C, pcfa_stream_t stream) {
    if (!zr || !h || !z || !r || !rh || B <= 0 || n <= 0 || B > 65535 || channels_last_C < 0) return PCFA_E_BADARG;
    if (channels_last_C) {
        if (n % channels_last_C != 0) return PCFA_E_BADARG;
        const int64_t tot = (int64_t)B * n;
        const int vec = (channels_last_C % 4 == 0 && aligned16({zr, h, z, r, rh})) ? 1 : 0;
        gru_gates_fwd_kernel<<<dim3(gru_grid(tot, vec), 1), GRU_THREADS, 0, as_stream(stream)>>>(zr, h, z, r, rh, tot, vec, channels_last_C);
        return after_launch();
    }
    const int vec = (n % 4 == 0 && aligned16({zr, h, z, r, rh})) ? 1 : 0;
    gru_gates_fwd_kernel<<<dim3(gru_grid(n, vec), B), GRU_THREADS, 0, as_stream(stream)>>>(zr, h, z, r, rh, n, vec, 0);
    return after_launch();
}

extern "C" int pcfa_gru_gates_backward(const float* z, const float* r, const float* h, const float* grad_z, const float* grad_rh,
                                       float* grad_zr, float* grad_h, int B, int64_t n, int channels_last_C,
                                       pcfa_stream_t stream) {
    if (!z || !r || !h || !grad_zr || !grad_h || B <= 0 || n <= 0 || B > 65535 || channels_last_C < 0) return PCFA_E_BADARG;
    if (channels_last_C) {
        if (n % channels_last_C != 0) return PCFA_E_BADARG;
        const int64_t tot = (int64_t)B * n;
        const int vec = (channels_last_C % 4 == 0) ? 1 : 0;
        gru_gates_bwd_kernel<<<dim3(gru_grid(tot, vec), 1), GRU_THREADS, 0, as_stream(stream)>>>(z, r, h, grad_z, grad_rh, grad_zr,
                                                                                                grad_h, tot, vec, channels_last_C);
        return after_launch();
    }
    const int vec = (n % 4 == 0) ? 1 : 0;
    gru_gates_bwd_kernel<<<dim3(gru_grid(n, vec), B), GRU_THREADS, 0, as_stream(stream)>>>(z, r, h, grad_z, grad_rh, grad_zr,
                                                                                          grad_h, n, vec, 0);
    return after_launch();
}

extern "C" int pcfa_gru_blend_forward(const float* z, const float* q_pre, const float* h, float* q, float* h_new, int64_t numel,
                                      pcfa_stream_t stream) {
    if (!z || !q_pre || !h || !q || !h_new || numel <= 0) return PCFA_E_BADARG;
    const int vec = (numel % 4 == 0) ? 1 : 0;
    gru_blend_fwd_kernel<<<gru_grid(numel, vec), GRU_THREADS, 0, as_stream(stream)>>>(z, q_pre, h, q, h_new, numel, vec);
    return after_launch();
}

extern "C" int pcfa_gru_blend_backward(const float* z, const float* q, const float* h, const float* grad_h_new, float* grad_z,
                                       float* grad_q_pre, float* grad_h, int64_t numel, pcfa_stream_t stream) {
    if (!z || !q || !h || !grad_h_new || !grad_z || !grad_q_pre || !grad_h || numel <= 0) return PCFA_E_BADARG;
    const int vec = (numel % 4 == 0) ? 1 : 0;
    gru_blend_bwd_kernel<<<gru_grid(numel, vec), GRU_THREADS, 0, as_stream(stream)>>>(z, q, h, grad_h_new, grad_z, grad_q_pre,
                                                                                     grad_h, numel, vec);
    return after_launch();
}

extern "C" int pcfa_cat_channels_last(const float* const* inputs, const int* channels, int n_inputs, float* out, int64_t npix,
                                      pcfa_stream_t stream) {
    if (!inputs || !channels || !out || n_inputs < 1 || n_inputs > 4 || npix <= 0) return PCFA_E_BADARG;
    CatArgs a{};
    a.n = n_inputs;
    uintptr_t al = reinterpret_cast<uintptr_t>(out);
    int vec = 1;
    for (int k = 0; k < n_inputs; ++k) {
        if (!inputs[k] || channels[k] <= 0) return PCFA_E_BADARG;
        a.in[k] = inputs[k]; a.c[k] = channels[k]; a.ctot += channels[k];
        al |= reinterpret_cast<uintptr_t>(inputs[k]);
        if (channels[k] % 4) vec = 0;
    }
    if (al & 15) vec = 0;
    const int64_t total = npix * (a.ctot / (vec ? 4 : 1));
    int64_t blocks = (total + GRU_THREADS - 1) / GRU_THREADS;
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    cat_cl_kernel<<<(int)blocks, GRU_THREADS, 0, as_stream(stream)>>>(a, out, npix, vec);
    return after_launch();
}

extern "C" int pcfa_cat_channels_last_pad(const float* const* inputs, const int* channels, int n_inputs, float* out, int64_t npix,
                                          int out_channels, pcfa_stream_t stream) {
    if (!inputs || !channels || !out || n_inputs < 1 || n_inputs > 4 || npix <= 0 || out_channels <= 0 || out_channels % 4) return PCFA_E_BADARG;
    CatArgs a{};
    a.n = n_inputs;
    for (int k = 0; k < n_inputs; ++k) {
        if (!inputs[k] || channels[k] <= 0) return PCFA_E_BADARG;
        a.in[k] = inputs[k]; a.c[k] = channels[k]; a.ctot += channels[k];
    }
    if (a.ctot > out_channels || (reinterpret_cast<uintptr_t>(out) & 15)) return PCFA_E_BADARG;
    const int64_t total = npix * (out_channels / 4);
    int64_t blocks = (total + GRU_THREADS - 1) / GRU_THREADS;
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    cat_cl_pad_kernel<<<(int)blocks, GRU_THREADS, 0, as_stream(stream)>>>(a, out, npix, out_channels);
    return after_launch();
}

static int grux_check(int C, int Cm, int64_t npix, std::initializer_list<const void*> ps) {
    if (C <= 0 || Cm < 0 || C % 4 || Cm % 4 || npix <= 0) return PCFA_E_BADARG;
    for (const void* p : ps) if (p && (reinterpret_cast<uintptr_t>(p) & 15)) return PCFA_E_BADARG;
    return PCFA_OK;
}

extern "C" int pcfa_gru_gates_x_forward(const float* zr, const float* addend, const float* h, const float* m, float* z, float* r,
                                        float* rhm, int C, int Cm, int64_t npix, pcfa_stream_t stream) {
    if (!zr || !addend || !h || !m || !z || !r || !rhm) return PCFA_E_BADARG;
    PCFA_TRY(grux_check(C, Cm, npix, {zr, addend, h, m, z, r, rhm}));
    const GruX g{C, Cm, npix};
    gru_gates_x_fwd_kernel<<<gru_grid(npix * (C + Cm), 1), GRU_THREADS, 0, as_stream(stream)>>>(zr, addend, h, m, z, r, rhm, g);
    return after_launch();
}

extern "C" int pcfa_gru_gates_x_backward(const float* z, const float* r, const float* h, const float* grad_z, const float* grad_rhm,
                                         float* grad_zr, float* grad_h, int C, int Cm, int64_t npix, pcfa_stream_t stream) {
    if (!z || !r || !h || !grad_zr || !grad_h) return PCFA_E_BADARG;
    PCFA_TRY(grux_check(C, Cm, npix, {z, r, h, grad_z, grad_rhm, grad_zr, grad_h}));
    const GruX g{C, Cm, npix};
    gru_gates_x_bwd_kernel<<<gru_grid(npix * C, 1), GRU_THREADS, 0, as_stream(stream)>>>(z, r, h, grad_z, grad_rhm, grad_zr, grad_h, g);
    return after_launch();
}

extern "C" int pcfa_gru_blend_x_forward(const float* z, const float* q_pre, const float* addend, const float* h, const float* m,
                                        float* q, float* h_new, float* hm /* may be NULL */, int C, int Cm, int64_t npix,
                                        pcfa_stream_t stream) {
    if (!z || !q_pre || !addend || !h || !q || !h_new || (hm && !m)) return PCFA_E_BADARG;
    PCFA_TRY(grux_check(C, Cm, npix, {z, q_pre, addend, h, m, q, h_new, hm}));
    const GruX g{C, Cm, npix};
    gru_blend_x_fwd_kernel<<<gru_grid(npix * (hm ? C + Cm : C), 1), GRU_THREADS, 0, as_stream(stream)>>>(z, q_pre, addend, h, m, q,
                                                                                                         h_new, hm, g);
    return after_launch();
}

extern "C" int pcfa_gru_blend_x_backward(const float* z, const float* q, const float* h, const float* grad_h_new, const float* grad_hm,
                                         float* grad_z, float* grad_q_pre, float* grad_h, int C, int Cm, int64_t npix,
                                         pcfa_stream_t stream) {
    if (!z || !q || !h || !grad_z || !grad_q_pre || !grad_h) return PCFA_E_BADARG;
    PCFA_TRY(grux_check(C, Cm, npix, {z, q, h, grad_h_new, grad_hm, grad_z, grad_q_pre, grad_h}));
    const GruX g{C, Cm, npix};
    gru_blend_x_bwd_kernel<<<gru_grid(npix * C, 1), GRU_THREADS, 0, as_stream(stream)>>>(z, q, h, grad_h_new, grad_hm, grad_z, grad_q_pre,
                                                                                        grad_h, g);
    return after_launch();
}

extern "C" int pcfa_gru_gates_x_backward_acc(const float* z, const float* r, const float* h, const float* grad_z, const float* grad_rhm,
                                             float* grad_zr, float* grad_h, float* acc, int acc_mode, int C, int Cm, int64_t npix,
                                             pcfa_stream_t stream) {
    if (!z || !r || !h || !grad_z || !grad_rhm || !grad_zr || !grad_h || acc_mode < 0 || acc_mode > 2 || (acc_mode && !acc)) return PCFA_E_BADARG;
    PCFA_TRY(grux_check(C, Cm, npix, {z, r, h, grad_z, grad_rhm, grad_zr, grad_h, acc}));
    const GruX g{C, Cm, npix};
    gru_gates_x_bwd2_kernel<<<gru_grid(npix * C, 1), GRU_THREADS, 0, as_stream(stream)>>>(z, r, h, grad_z, grad_rhm, grad_zr, grad_h, acc, acc_mode, g);
    return after_launch();
}

extern "C" int pcfa_gru_blend_x_backward_acc(const float* z, const float* q, const float* h, const float* grad_h_new_a,
                                             const float* grad_h_new_b, const float* grad_hm, float* grad_z, float* grad_q_pre,
                                             float* grad_h, float* acc, int acc_mode, int C, int Cm, int64_t npix, pcfa_stream_t stream) {
    if (!z || !q || !h || !grad_h_new_a || !grad_z || !grad_q_pre || !grad_h || acc_mode < 0 || acc_mode > 2 || (acc_mode && !acc)) return PCFA_E_BADARG;
    PCFA_TRY(grux_check(C, Cm, npix, {z, q, h, grad_h_new_a, grad_h_new_b, grad_hm, grad_z, grad_q_pre, grad_h, acc}));
    const GruX g{C, Cm, npix};
    gru_blend_x_bwd2_kernel<<<gru_grid(npix * C, 1), GRU_THREADS, 0, as_stream(stream)>>>(z, q, h, grad_h_new_a, grad_h_new_b, grad_hm, grad_z,
                                                                                         grad_q_pre, grad_h, acc, acc_mode, g);
    return after_launch();
}

extern "C" int pcfa_gru_step_combine(const float* gh_a, const float* gh_b, const float* cat0, const float* cat1, const float* cat2,
                                     const float* cat3, float* grad_h, float* grad_m, int C, int Cm, int64_t npix, pcfa_stream_t stream) {
    if (!gh_a || !gh_b || !cat0 || !cat1 || !cat2 || !cat3 || !grad_h || !grad_m || Cm <= 0) return PCFA_E_BADARG;
    PCFA_TRY(grux_check(C, Cm, npix, {gh_a, gh_b, cat0, cat1, cat2, cat3, grad_h, grad_m}));
    const GruX g{C, Cm, npix};
    gru_step_combine_kernel<<<gru_grid(npix * (C + Cm), 1), GRU_THREADS, 0, as_stream(stream)>>>(gh_a, gh_b, cat0, cat1, cat2, cat3, grad_h, grad_m, g);
    return after_launch();
}
