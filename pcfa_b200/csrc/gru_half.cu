// fp16-storage variants of the whole-GRU-step kernels (csrc/gru.cu) for GMA under its shipped fp16 autocast
// (models/_config/gma_config.json:5; models/gma/update.py:33-60 SepConvGRU inside the autocast region of network.py:104).
// Storage is half (what autocast's convolutions produce and consume), arithmetic is fp32 per element: each output is the
// correctly rounded fp16 of the fp32 expression, whereas the reference's chain of fp16 ATen ops rounds after every
// sigmoid / mul / add.  Same pixel-major channels-last indexing as the fp32 kernels; 4 channels (8 bytes) per access.
#include "common.cuh"
#include <cuda_fp16.h>
#include <initializer_list>

namespace pcfa {

constexpr int GH_THREADS = 256;
struct GruXh { int C, Cm; int64_t npix; };

__device__ __forceinline__ float4 ldh4(const __half* p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void sth4(__half* p, float4 v) {
    uint2 r;
    *reinterpret_cast<__half2*>(&r.x) = __floats2half2_rn(v.x, v.y);
    *reinterpret_cast<__half2*>(&r.y) = __floats2half2_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = r;
}
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// out[p] = [a[p] (Ca) | b[p] (Cb)]
__global__ void __launch_bounds__(GH_THREADS)
cat2_h_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ out, int Ca, int Cb, int64_t npix) {
    const int Q = (Ca + Cb) >> 2, QA = Ca >> 2;
    const int64_t total = npix * Q;
    for (int64_t e = (int64_t)blockIdx.x * GH_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GH_THREADS) {
        const int64_t p = e / Q;
        const int qd = (int)(e - p * Q);
        const uint2 v = qd < QA ? __ldg(reinterpret_cast<const uint2*>(a + p * Ca) + qd) : __ldg(reinterpret_cast<const uint2*>(b + p * Cb) + (qd - QA));
        reinterpret_cast<uint2*>(out + p * (Ca + Cb))[qd] = v;
    }
}

__global__ void __launch_bounds__(GH_THREADS)
gates_x_fwd_h_kernel(const __half* __restrict__ zr, const __half* __restrict__ P, const __half* __restrict__ h, const __half* __restrict__ m,
                     __half* __restrict__ z, __half* __restrict__ r, __half* __restrict__ rhm, GruXh g) {
    const int Q = (g.C + g.Cm) >> 2, QC = g.C >> 2;
    const int64_t total = g.npix * Q;
    for (int64_t e = (int64_t)blockIdx.x * GH_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GH_THREADS) {
        const int64_t p = e / Q;
        const int qd = (int)(e - p * Q);
        __half* dst = rhm + p * (g.C + g.Cm) + 4 * qd;
        if (qd >= QC) { *reinterpret_cast<uint2*>(dst) = __ldg(reinterpret_cast<const uint2*>(m + p * g.Cm) + (qd - QC)); continue; }
        const float4 a = add4(ldh4(zr + p * 2 * g.C + 4 * qd), ldh4(P + p * 2 * g.C + 4 * qd));
        const float4 c = add4(ldh4(zr + p * 2 * g.C + g.C + 4 * qd), ldh4(P + p * 2 * g.C + g.C + 4 * qd));
        const float4 hv = ldh4(h + p * g.C + 4 * qd);
        const float4 zz = make_float4(sigm(a.x), sigm(a.y), sigm(a.z), sigm(a.w));
        const float4 rr = make_float4(sigm(c.x), sigm(c.y), sigm(c.z), sigm(c.w));
        sth4(z + p * g.C + 4 * qd, zz);
        sth4(r + p * g.C + 4 * qd, rr);
        sth4(dst, make_float4(rr.x * hv.x, rr.y * hv.y, rr.z * hv.z, rr.w * hv.w));
    }
}

// q = tanh(q_pre + P), h' = (1-z) h + z q ; hm (may be NULL) = [h' | m]
__global__ void __launch_bounds__(GH_THREADS)
blend_x_fwd_h_kernel(const __half* __restrict__ z, const __half* __restrict__ qc, const __half* __restrict__ P, const __half* __restrict__ h,
                     const __half* __restrict__ m, __half* __restrict__ q, __half* __restrict__ hn, __half* __restrict__ hm, GruXh g) {
    const int QC = g.C >> 2, Q = hm ? (g.C + g.Cm) >> 2 : QC;
    const int64_t total = g.npix * Q;
    for (int64_t e = (int64_t)blockIdx.x * GH_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GH_THREADS) {
        const int64_t p = e / Q;
        const int qd = (int)(e - p * Q);
        if (qd >= QC) {
            reinterpret_cast<uint2*>(hm + p * (g.C + g.Cm))[qd] = __ldg(reinterpret_cast<const uint2*>(m + p * g.Cm) + (qd - QC));
            continue;
        }
        const float4 zz = ldh4(z + p * g.C + 4 * qd), a = add4(ldh4(qc + p * g.C + 4 * qd), ldh4(P + p * g.C + 4 * qd)), hv = ldh4(h + p * g.C + 4 * qd);
        const float4 qq = make_float4(tanhf(a.x), tanhf(a.y), tanhf(a.z), tanhf(a.w));
        const float4 o = make_float4((1.f - zz.x) * hv.x + zz.x * qq.x, (1.f - zz.y) * hv.y + zz.y * qq.y,
                                     (1.f - zz.z) * hv.z + zz.z * qq.z, (1.f - zz.w) * hv.w + zz.w * qq.w);
        sth4(q + p * g.C + 4 * qd, qq);
        sth4(hn + p * g.C + 4 * qd, o);
        if (hm) sth4(hm + p * (g.C + g.Cm) + 4 * qd, o);
    }
}

__device__ __forceinline__ void acc4h(__half* acc, int mode, float4 v) {
    if (mode == 1) sth4(acc, v);
    else if (mode == 2) sth4(acc, add4(ldh4(acc), v));
}

__global__ void __launch_bounds__(GH_THREADS)
gates_x_bwd_h_kernel(const __half* __restrict__ z, const __half* __restrict__ r, const __half* __restrict__ h, const __half* __restrict__ gz,
                     const __half* __restrict__ grhm, __half* __restrict__ gzr, __half* __restrict__ gh, __half* __restrict__ acc, int acc_mode,
                     GruXh g) {
    const int QC = g.C >> 2;
    const int64_t total = g.npix * QC;
    for (int64_t e = (int64_t)blockIdx.x * GH_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GH_THREADS) {
        const int64_t p = e / QC;
        const int qd = (int)(e - p * QC);
        const float4 zz = ldh4(z + p * g.C + 4 * qd), rr = ldh4(r + p * g.C + 4 * qd), hv = ldh4(h + p * g.C + 4 * qd);
        const float4 a = ldh4(gz + p * g.C + 4 * qd), b = ldh4(grhm + p * (g.C + g.Cm) + 4 * qd);
        const float4 o1 = make_float4(a.x * zz.x * (1.f - zz.x), a.y * zz.y * (1.f - zz.y), a.z * zz.z * (1.f - zz.z), a.w * zz.w * (1.f - zz.w));
        const float4 o2 = make_float4(b.x * hv.x * rr.x * (1.f - rr.x), b.y * hv.y * rr.y * (1.f - rr.y), b.z * hv.z * rr.z * (1.f - rr.z),
                                      b.w * hv.w * rr.w * (1.f - rr.w));
        sth4(gzr + p * 2 * g.C + 4 * qd, o1);
        sth4(gzr + p * 2 * g.C + g.C + 4 * qd, o2);
        acc4h(acc + p * 2 * g.C + 4 * qd, acc_mode, o1);
        acc4h(acc + p * 2 * g.C + g.C + 4 * qd, acc_mode, o2);
        sth4(gh + p * g.C + 4 * qd, make_float4(b.x * rr.x, b.y * rr.y, b.z * rr.z, b.w * rr.w));
    }
}

__global__ void __launch_bounds__(GH_THREADS)
blend_x_bwd_h_kernel(const __half* __restrict__ z, const __half* __restrict__ q, const __half* __restrict__ h, const __half* __restrict__ ghn_a,
                     const __half* __restrict__ ghn_b, const __half* __restrict__ ghm, __half* __restrict__ gz, __half* __restrict__ gq,
                     __half* __restrict__ gh, __half* __restrict__ acc, int acc_mode, GruXh g) {
    const int QC = g.C >> 2;
    const int64_t total = g.npix * QC;
    for (int64_t e = (int64_t)blockIdx.x * GH_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GH_THREADS) {
        const int64_t p = e / QC;
        const int qd = (int)(e - p * QC);
        float4 d = ldh4(ghn_a + p * g.C + 4 * qd);
        if (ghn_b) d = add4(d, ldh4(ghn_b + p * g.C + 4 * qd));
        if (ghm) d = add4(d, ldh4(ghm + p * (g.C + g.Cm) + 4 * qd));
        const float4 zz = ldh4(z + p * g.C + 4 * qd), qq = ldh4(q + p * g.C + 4 * qd), hv = ldh4(h + p * g.C + 4 * qd);
        sth4(gz + p * g.C + 4 * qd, make_float4(d.x * (qq.x - hv.x), d.y * (qq.y - hv.y), d.z * (qq.z - hv.z), d.w * (qq.w - hv.w)));
        const float4 o = make_float4(d.x * zz.x * (1.f - qq.x * qq.x), d.y * zz.y * (1.f - qq.y * qq.y), d.z * zz.z * (1.f - qq.z * qq.z),
                                     d.w * zz.w * (1.f - qq.w * qq.w));
        sth4(gq + p * g.C + 4 * qd, o);
        acc4h(acc + p * g.C + 4 * qd, acc_mode, o);
        sth4(gh + p * g.C + 4 * qd, make_float4(d.x * (1.f - zz.x), d.y * (1.f - zz.y), d.z * (1.f - zz.z), d.w * (1.f - zz.w)));
    }
}

__global__ void __launch_bounds__(GH_THREADS)
step_combine_h_kernel(const __half* __restrict__ gh_a, const __half* __restrict__ gh_b, const __half* __restrict__ cat0, const __half* __restrict__ cat1,
                      const __half* __restrict__ cat2, const __half* __restrict__ cat3, __half* __restrict__ gh, __half* __restrict__ gm, GruXh g) {
    const int QC = g.C >> 2, Q = (g.C + g.Cm) >> 2, CT = g.C + g.Cm;
    const int64_t total = g.npix * Q;
    for (int64_t e = (int64_t)blockIdx.x * GH_THREADS + threadIdx.x; e < total; e += (int64_t)gridDim.x * GH_THREADS) {
        const int64_t p = e / Q;
        const int qd = (int)(e - p * Q);
        const float4 c0 = ldh4(cat0 + p * CT + 4 * qd);
        if (qd < QC) sth4(gh + p * g.C + 4 * qd, add4(add4(ldh4(gh_a + p * g.C + 4 * qd), ldh4(gh_b + p * g.C + 4 * qd)), c0));
        else sth4(gm + p * g.Cm + 4 * (qd - QC),
                  add4(add4(c0, ldh4(cat1 + p * CT + 4 * qd)), add4(ldh4(cat2 + p * CT + 4 * qd), ldh4(cat3 + p * CT + 4 * qd))));
    }
}

static int gh_grid(int64_t items) {
    int64_t b = (items + GH_THREADS - 1) / GH_THREADS;
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
static int gh_check(int C, int Cm, int64_t npix, std::initializer_list<const void*> ps) {
    if (C <= 0 || Cm < 0 || C % 4 || Cm % 4 || npix <= 0) return PCFA_E_BADARG;
    for (const void* p : ps) if (p && (reinterpret_cast<uintptr_t>(p) & 7)) return PCFA_E_BADARG;
    return PCFA_OK;
}

}  // namespace pcfa

using namespace pcfa;
#define H(p) reinterpret_cast<const __half*>(p)
#define HM(p) reinterpret_cast<__half*>(p)

extern "C" int pcfa_cat2_channels_last_h(const void* a, const void* b, void* out, int Ca, int Cb, int64_t npix, pcfa_stream_t stream) {
    if (!a || !b || !out) return PCFA_E_BADARG;
    PCFA_TRY(gh_check(Ca, Cb, npix, {a, b, out}));
    cat2_h_kernel<<<gh_grid(npix * ((Ca + Cb) / 4)), GH_THREADS, 0, as_stream(stream)>>>(H(a), H(b), HM(out), Ca, Cb, npix);
    return after_launch();
}
extern "C" int pcfa_gru_gates_x_forward_h(const void* zr, const void* addend, const void* h, const void* m, void* z, void* r, void* rhm, int C,
                                          int Cm, int64_t npix, pcfa_stream_t stream) {
    if (!zr || !addend || !h || !m || !z || !r || !rhm) return PCFA_E_BADARG;
    PCFA_TRY(gh_check(C, Cm, npix, {zr, addend, h, m, z, r, rhm}));
    gates_x_fwd_h_kernel<<<gh_grid(npix * ((C + Cm) / 4)), GH_THREADS, 0, as_stream(stream)>>>(H(zr), H(addend), H(h), H(m), HM(z), HM(r), HM(rhm),
                                                                                              GruXh{C, Cm, npix});
    return after_launch();
}
extern "C" int pcfa_gru_blend_x_forward_h(const void* z, const void* q_pre, const void* addend, const void* h, const void* m, void* q, void* h_new,
                                          void* hm, int C, int Cm, int64_t npix, pcfa_stream_t stream) {
    if (!z || !q_pre || !addend || !h || !q || !h_new || (hm && !m)) return PCFA_E_BADARG;
    PCFA_TRY(gh_check(C, Cm, npix, {z, q_pre, addend, h, m, q, h_new, hm}));
    blend_x_fwd_h_kernel<<<gh_grid(npix * ((hm ? C + Cm : C) / 4)), GH_THREADS, 0, as_stream(stream)>>>(H(z), H(q_pre), H(addend), H(h), H(m), HM(q),
                                                                                                       HM(h_new), HM(hm), GruXh{C, Cm, npix});
    return after_launch();
}
extern "C" int pcfa_gru_gates_x_backward_acc_h(const void* z, const void* r, const void* h, const void* grad_z, const void* grad_rhm, void* grad_zr,
                                               void* grad_h, void* acc, int acc_mode, int C, int Cm, int64_t npix, pcfa_stream_t stream) {
    if (!z || !r || !h || !grad_z || !grad_rhm || !grad_zr || !grad_h || acc_mode < 0 || acc_mode > 2 || (acc_mode && !acc)) return PCFA_E_BADARG;
    PCFA_TRY(gh_check(C, Cm, npix, {z, r, h, grad_z, grad_rhm, grad_zr, grad_h, acc}));
    gates_x_bwd_h_kernel<<<gh_grid(npix * (C / 4)), GH_THREADS, 0, as_stream(stream)>>>(H(z), H(r), H(h), H(grad_z), H(grad_rhm), HM(grad_zr), HM(grad_h),
                                                                                       HM(acc), acc_mode, GruXh{C, Cm, npix});
    return after_launch();
}
extern "C" int pcfa_gru_blend_x_backward_acc_h(const void* z, const void* q, const void* h, const void* grad_h_new_a, const void* grad_h_new_b,
                                               const void* grad_hm, void* grad_z, void* grad_q_pre, void* grad_h, void* acc, int acc_mode, int C,
                                               int Cm, int64_t npix, pcfa_stream_t stream) {
    if (!z || !q || !h || !grad_h_new_a || !grad_z || !grad_q_pre || !grad_h || acc_mode < 0 || acc_mode > 2 || (acc_mode && !acc)) return PCFA_E_BADARG;
    PCFA_TRY(gh_check(C, Cm, npix, {z, q, h, grad_h_new_a, grad_h_new_b, grad_hm, grad_z, grad_q_pre, grad_h, acc}));
    blend_x_bwd_h_kernel<<<gh_grid(npix * (C / 4)), GH_THREADS, 0, as_stream(stream)>>>(H(z), H(q), H(h), H(grad_h_new_a), H(grad_h_new_b), H(grad_hm),
                                                                                       HM(grad_z), HM(grad_q_pre), HM(grad_h), HM(acc), acc_mode,
                                                                                       GruXh{C, Cm, npix});
    return after_launch();
}
extern "C" int pcfa_gru_step_combine_h(const void* gh_a, const void* gh_b, const void* cat0, const void* cat1, const void* cat2, const void* cat3,
                                       void* grad_h, void* grad_m, int C, int Cm, int64_t npix, pcfa_stream_t stream) {
    if (!gh_a || !gh_b || !cat0 || !cat1 || !cat2 || !cat3 || !grad_h || !grad_m || Cm <= 0) return PCFA_E_BADARG;
    PCFA_TRY(gh_check(C, Cm, npix, {gh_a, gh_b, cat0, cat1, cat2, cat3, grad_h, grad_m}));
    step_combine_h_kernel<<<gh_grid(npix * ((C + Cm) / 4)), GH_THREADS, 0, as_stream(stream)>>>(H(gh_a), H(gh_b), H(cat0), H(cat1), H(cat2), H(cat3),
                                                                                               HM(grad_h), HM(grad_m), GruXh{C, Cm, npix});
    return after_launch();
}
