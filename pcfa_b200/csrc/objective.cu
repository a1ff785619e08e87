// Fused PCFA objective: box-constraint input transform (forward/backward) and loss + penalty.
// Replaces ~40 elementwise/reduction launches of the reference per closure:
//   ScaledInputModel.forward pre-processing   helper_functions/own_models.py:62-85
//   extract_deltas / extract_deltas_joint     attack_PCFA.py:20-37
//   loss_delta_constraint and friends         helper_functions/losses.py:3-44,76-88,110-126,177-230
//   InputPadder.unpad (+ the .cpu() hop)      helper_functions/ownutilities.py:51-62,297
// All reductions use a fixed grid and a fixed summation order, so results are bit-reproducible
// run to run and identical on every rank of the universal-perturbation mode.
#include "common.cuh"

namespace pcfa {

constexpr int BOX_THREADS = 256;
constexpr int LOSS_BLOCKS = 256;
constexpr int LOSS_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.f;
    if (warp == 0) {
        r = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.f;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;   // valid in warp 0
}

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// Carlini-Wagner change of variables, own_models.py:73-75 / attack_PCFA.py:23-24:
//   (1./2.) * 1./(1.-eps) * (tanh(w) + (1-eps))   with python-float scalars rounded to fp32
struct CovConsts { float half_inv; float one_m_eps; };
__device__ __forceinline__ float cov_u(float w, CovConsts k, float* th) {
    const float t = tanhf(w);
    *th = t;
    return k.half_inv * (t + k.one_m_eps);
}

__global__ void __launch_bounds__(BOX_THREADS)
box_forward_kernel(const float* __restrict__ var, const float* __restrict__ image,
                   const float* __restrict__ amax, const float* __restrict__ amin,
                   float* __restrict__ net_in, float* __restrict__ delta_out,
                   float* __restrict__ partials, int mode, int64_t B, int64_t chw, CovConsts k,
                   float scale) {
    __shared__ float sh[BOX_THREADS / 32];
    const int64_t total = B * chw;
    float ss = 0.f;
    for (int64_t idx = (int64_t)blockIdx.x * BOX_THREADS + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * BOX_THREADS) {
        const float I = image[idx];
        float x, d;
        if (mode == PCFA_BOX_COV) {
            float t;
            const float u = cov_u(var[idx], k, &t);
            x = clamp01(u);
            d = u - I;
            ss += d * d;
        } else if (mode == PCFA_BOX_CLIP) {
            x = clamp01(var[idx]);
            d = x - I;
            ss += d * d;
        } else if (mode == PCFA_BOX_JOINT) {
            const float dl = var[idx];
            x = clamp01(I + dl);
            const float mx = amax[idx], mn = amin[idx];
            const float up = clamp01(dl + mx) - mx;          // attack_PCFA.py:34
            d = clamp01(up + mn) - mn;                       // attack_PCFA.py:35
            ss += d * d;
        } else {   // UNIVERSAL: var is [chw], broadcast over the batch; penalty on the raw delta
            const int64_t e = idx % chw;
            d = var[e];
            x = clamp01(I + d);
            if (idx < chw) ss += d * d;
        }
        net_in[idx] = scale * x;
        if (delta_out && (mode != PCFA_BOX_UNIVERSAL || idx < chw)) delta_out[idx] = d;
    }
    const float r = block_sum(ss, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

// grad wrt the optimisation variable.  coef = loss_terms[2] = mu*2/numel_total if the penalty is
// active else 0 (device-resident, so no host round trip between loss and this kernel).
__global__ void __launch_bounds__(BOX_THREADS)
box_backward_kernel(const float* __restrict__ var, const float* __restrict__ image,
                    const float* __restrict__ amax, const float* __restrict__ amin,
                    const float* __restrict__ gnet, const float* __restrict__ loss_terms,
                    float* __restrict__ gvar, int accumulate, int mode, int64_t B, int64_t chw,
                    CovConsts k, float scale) {
    const float coef = loss_terms ? loss_terms[2] : 0.f;
    if (mode == PCFA_BOX_UNIVERSAL) {
        for (int64_t e = (int64_t)blockIdx.x * BOX_THREADS + threadIdx.x; e < chw;
             e += (int64_t)gridDim.x * BOX_THREADS) {
            const float d = var[e];
            float g = coef * d;
            if (gnet) {
                for (int64_t b = 0; b < B; ++b) {
                    const float s = image[b * chw + e] + d;
                    if (s >= 0.f && s <= 1.f) g += scale * gnet[b * chw + e];   // inclusive clamp mask
                }
            }
            gvar[e] = accumulate ? gvar[e] + g : g;
        }
        return;
    }
    const int64_t total = B * chw;
    for (int64_t idx = (int64_t)blockIdx.x * BOX_THREADS + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * BOX_THREADS) {
        const float I = image[idx];
        const float gn = gnet ? scale * gnet[idx] : 0.f;
        float g;
        if (mode == PCFA_BOX_COV) {
            float t;
            const float u = cov_u(var[idx], k, &t);
            const float du = k.half_inv * (1.f - t * t);
            const float m = (u >= 0.f && u <= 1.f) ? 1.f : 0.f;
            g = (gn * m + coef * (u - I)) * du;
        } else if (mode == PCFA_BOX_CLIP) {
            const float v = var[idx];
            const float m = (v >= 0.f && v <= 1.f) ? 1.f : 0.f;
            g = m * (gn + coef * (clamp01(v) - I));
        } else {   // JOINT
            const float dl = var[idx];
            const float s = I + dl;
            const float m = (s >= 0.f && s <= 1.f) ? 1.f : 0.f;
            const float mx = amax[idx], mn = amin[idx];
            const float a = dl + mx;
            const float up = clamp01(a) - mx;
            const float bq = up + mn;
            const float d = clamp01(bq) - mn;
            const float md = ((a >= 0.f && a <= 1.f) && (bq >= 0.f && bq <= 1.f)) ? 1.f : 0.f;
            g = gn * m + coef * d * md;
        }
        gvar[idx] = accumulate ? gvar[idx] + g : g;
    }
}


// d delta / d var applied to an arbitrary incoming gradient (autograd of extract_deltas /
// extract_deltas_joint, attack_PCFA.py:20-37, when the deltas are consumed outside the fused loss).
__global__ void __launch_bounds__(BOX_THREADS)
box_delta_backward_kernel(const float* __restrict__ var, const float* __restrict__ amax,
                          const float* __restrict__ amin, const float* __restrict__ gdelta,
                          float* __restrict__ gvar, int accumulate, int mode, int64_t total,
                          CovConsts k) {
    for (int64_t idx = (int64_t)blockIdx.x * BOX_THREADS + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * BOX_THREADS) {
        float j;
        if (mode == PCFA_BOX_COV) {
            const float t = tanhf(var[idx]);
            j = k.half_inv * (1.f - t * t);
        } else if (mode == PCFA_BOX_CLIP) {
            const float v = var[idx];
            j = (v >= 0.f && v <= 1.f) ? 1.f : 0.f;
        } else if (mode == PCFA_BOX_JOINT) {
            const float a = var[idx] + amax[idx];
            const float bq = clamp01(a) - amax[idx] + amin[idx];
            j = ((a >= 0.f && a <= 1.f) && (bq >= 0.f && bq <= 1.f)) ? 1.f : 0.f;
        } else {
            j = 1.f;
        }
        const float g = gdelta[idx] * j;
        gvar[idx] = accumulate ? gvar[idx] + g : g;
    }
}

__global__ void __launch_bounds__(BOX_THREADS)
sumsq_partials_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ partials) {
    __shared__ float sh[BOX_THREADS / 32];
    float ss = 0.f;
    for (int64_t idx = (int64_t)blockIdx.x * BOX_THREADS + threadIdx.x; idx < n;
         idx += (int64_t)gridDim.x * BOX_THREADS) {
        const float v = x[idx];
        ss += v * v;
    }
    const float r = block_sum(ss, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
}

// ---- loss ---------------------------------------------------------------------------------
// pass 1: per-block partial sums over the (padded) flow domain; writes the similarity gradient for
// AEE / MSE directly (it needs no global quantity), zero in the padding.
//   ws layout: [3][LOSS_BLOCKS] floats  (sum0, sum1, sum2)
//     AEE: sum0 = sum ||f - t||            MSE: sum0 = sum (f-t)^2
//     COSIM: sum0 = sum f*t, sum1 = sum f*f, sum2 = sum t*t
__global__ void __launch_bounds__(LOSS_THREADS)
loss_partial_kernel(const float* __restrict__ flow, const float* __restrict__ target,
                    float* __restrict__ gflow, float* __restrict__ ws, int loss_type, int B, int H,
                    int W, int Hp, int Wp, int pt, int pl) {
    __shared__ float sh[LOSS_THREADS / 32];
    const int64_t ppl = (int64_t)Hp * Wp, upl = (int64_t)H * W;
    const float inv_n_aee = 1.0f / (float)((int64_t)B * H * W);
    const float inv_n_mse = 1.0f / (float)((int64_t)B * 2 * H * W);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    // block k owns the padded rows [k*R/G, (k+1)*R/G): a fixed partition (deterministic partial sums) without any
    // per-pixel 64-bit division
    const int64_t rows = (int64_t)B * Hp;
    const int64_t r_lo = rows * blockIdx.x / gridDim.x, r_hi = rows * (blockIdx.x + 1) / gridDim.x;
    for (int64_t row = r_lo; row < r_hi; ++row) {
      const int b = (int)(row / Hp), yp = (int)(row - (int64_t)b * Hp);
      const int y = yp - pt;
      for (int xp = threadIdx.x; xp < Wp; xp += LOSS_THREADS) {
        const int x = xp - pl;
        float gu = 0.f, gv = 0.f;
        if (x >= 0 && x < W && y >= 0 && y < H) {
            const float fu = flow[((int64_t)b * 2 + 0) * ppl + (int64_t)yp * Wp + xp];
            const float fv = flow[((int64_t)b * 2 + 1) * ppl + (int64_t)yp * Wp + xp];
            const float tu = target[((int64_t)b * 2 + 0) * upl + (int64_t)y * W + x];
            const float tv = target[((int64_t)b * 2 + 1) * upl + (int64_t)y * W + x];
            if (loss_type == PCFA_LOSS_AEE) {
                const float du = fu - tu, dv = fv - tv;
                const float n = sqrtf(du * du + dv * dv);
                s0 += n;
                if (n > 0.f) { gu = du / n * inv_n_aee; gv = dv / n * inv_n_aee; }   // 0 at the kink
            } else if (loss_type == PCFA_LOSS_MSE) {
                const float du = fu - tu, dv = fv - tv;
                s0 += du * du + dv * dv;
                gu = 2.f * du * inv_n_mse; gv = 2.f * dv * inv_n_mse;
            } else {
                s0 += fu * tu + fv * tv;
                s1 += fu * fu + fv * fv;
                s2 += tu * tu + tv * tv;
            }
        }
        if (gflow && loss_type != PCFA_LOSS_COSIM) {
            gflow[((int64_t)b * 2 + 0) * ppl + (int64_t)yp * Wp + xp] = gu;
            gflow[((int64_t)b * 2 + 1) * ppl + (int64_t)yp * Wp + xp] = gv;
        }
      }
    }
    const float r0 = block_sum(s0, sh);
    const float r1 = block_sum(s1, sh);
    const float r2 = block_sum(s2, sh);
    if (threadIdx.x == 0) {
        ws[blockIdx.x] = r0;
        ws[LOSS_BLOCKS + blockIdx.x] = r1;
        ws[2 * LOSS_BLOCKS + blockIdx.x] = r2;
    }
}

// pass 2 (one block): ordered final sums in double, loss terms, penalty switch.
__global__ void __launch_bounds__(256)
loss_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ p1,
                     const float* __restrict__ p2, float w1, float w2, float* __restrict__ terms,
                     float* __restrict__ cos_sums, int loss_type, int B, int H, int W,
                     double numel_total, float delta_bound, float mu) {
    __shared__ double red[256];
    auto ordered_sum = [&](const float* v, int n) -> double {
        double a = 0.0;
        if (v) for (int i = threadIdx.x; i < n; i += 256) a += (double)v[i];
        red[threadIdx.x] = a;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        const double r = red[0];
        __syncthreads();
        return r;
    };
    const double s0 = ordered_sum(ws, LOSS_BLOCKS);
    const double s1 = ordered_sum(ws + LOSS_BLOCKS, LOSS_BLOCKS);
    const double s2 = ordered_sum(ws + 2 * LOSS_BLOCKS, LOSS_BLOCKS);
    const double q1 = ordered_sum(p1, PCFA_BOX_PARTIALS);
    const double q2 = ordered_sum(p2, PCFA_BOX_PARTIALS);
    if (threadIdx.x == 0) {
        float sim;
        if (loss_type == PCFA_LOSS_AEE)      sim = (float)(s0 / ((double)B * H * W));
        else if (loss_type == PCFA_LOSS_MSE) sim = (float)(s0 / ((double)B * 2 * H * W));
        else {
            // 1 - sum(p*t) / sqrt(sum(p*p)) * sqrt(sum(t*t))      (sic, losses.py:88)
            sim = 1.0f - (float)s0 / sqrtf((float)s1) * sqrtf((float)s2);
            cos_sums[0] = (float)s0; cos_sums[1] = (float)s1; cos_sums[2] = (float)s2;
        }
        const float mean_sq = (float)(((double)w1 * q1 + (double)w2 * q2) / numel_total);
        const float excess = mean_sq - delta_bound * delta_bound;       // losses.py:196
        const float pen = fmaxf(0.f, excess);
        terms[0] = sim + mu * pen;
        terms[1] = sim;
        terms[2] = (excess > 0.f) ? (float)((double)mu * 2.0 / numel_total) : 0.f;
        terms[3] = mean_sq;
    }
}

// pass 3 (cosim only): gradient needs the global sums.
__global__ void __launch_bounds__(LOSS_THREADS)
loss_cosim_grad_kernel(const float* __restrict__ flow, const float* __restrict__ target,
                       const float* __restrict__ cos_sums, float* __restrict__ gflow, int B, int H,
                       int W, int Hp, int Wp, int pt, int pl) {
    const int64_t ppl = (int64_t)Hp * Wp, upl = (int64_t)H * W;
    const float A = cos_sums[0], P = cos_sums[1], T = cos_sums[2];
    const float sqT = sqrtf(T), sqP = sqrtf(P);
    // d/dp [1 - A/sqrt(P)*sqrt(T)] = -sqrt(T) * ( t/sqrt(P) - A*p/P^{3/2} )
    const float c_t = -sqT / sqP, c_p = sqT * A / (P * sqP);
    const int64_t rows = (int64_t)B * Hp;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
      const int b = (int)(row / Hp), yp = (int)(row - (int64_t)b * Hp);
      const int y = yp - pt;
      for (int xp = threadIdx.x; xp < Wp; xp += LOSS_THREADS) {
        const int x = xp - pl;
        float gu = 0.f, gv = 0.f;
        if (x >= 0 && x < W && y >= 0 && y < H) {
            const float fu = flow[((int64_t)b * 2 + 0) * ppl + (int64_t)yp * Wp + xp];
            const float fv = flow[((int64_t)b * 2 + 1) * ppl + (int64_t)yp * Wp + xp];
            const float tu = target[((int64_t)b * 2 + 0) * upl + (int64_t)y * W + x];
            const float tv = target[((int64_t)b * 2 + 1) * upl + (int64_t)y * W + x];
            gu = c_t * tu + c_p * fu;
            gv = c_t * tv + c_p * fv;
        }
        gflow[((int64_t)b * 2 + 0) * ppl + (int64_t)yp * Wp + xp] = gu;
        gflow[((int64_t)b * 2 + 1) * ppl + (int64_t)yp * Wp + xp] = gv;
      }
    }
}

static CovConsts cov_consts(float eps_box) {
    // python: (1./2.)*1./(1.-eps) and (1-eps) are float64 scalars that torch rounds to fp32
    CovConsts k;
    k.half_inv = (float)(0.5 * 1.0 / (1.0 - (double)eps_box));
    k.one_m_eps = (float)(1.0 - (double)eps_box);
    return k;
}

}  // namespace pcfa

using namespace pcfa;

extern "C" int pcfa_box_forward(const float* var, const float* image, const float* aux_max,
                                const float* aux_min, float* net_in, float* delta_out,
                                float* sumsq_partials, int mode, int B, int64_t chw, float eps_box,
                                float scale, pcfa_stream_t stream) {
    if (!var || !image || !net_in || !sumsq_partials || B <= 0 || chw <= 0) return PCFA_E_BADARG;
    if (mode < PCFA_BOX_COV || mode > PCFA_BOX_UNIVERSAL) return PCFA_E_BADARG;
    if (mode == PCFA_BOX_JOINT && (!aux_max || !aux_min)) return PCFA_E_BADARG;
    box_forward_kernel<<<PCFA_BOX_PARTIALS, BOX_THREADS, 0, as_stream(stream)>>>(
        var, image, aux_max, aux_min, net_in, delta_out, sumsq_partials, mode, B, chw,
        cov_consts(eps_box), scale);
    return after_launch();
}

extern "C" int pcfa_box_backward(const float* var, const float* image, const float* aux_max,
                                 const float* aux_min, const float* grad_net_in,
                                 const float* loss_terms, float* grad_var, int accumulate, int mode,
                                 int B, int64_t chw, float eps_box, float scale,
                                 pcfa_stream_t stream) {
    if (!var || !image || !grad_var || B <= 0 || chw <= 0) return PCFA_E_BADARG;
    if (mode < PCFA_BOX_COV || mode > PCFA_BOX_UNIVERSAL) return PCFA_E_BADARG;
    if (mode == PCFA_BOX_JOINT && (!aux_max || !aux_min)) return PCFA_E_BADARG;
    const int64_t work = (mode == PCFA_BOX_UNIVERSAL) ? chw : (int64_t)B * chw;
    int64_t blocks = ceil_div<int64_t>(work, BOX_THREADS);
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    box_backward_kernel<<<(int)blocks, BOX_THREADS, 0, as_stream(stream)>>>(
        var, image, aux_max, aux_min, grad_net_in, loss_terms, grad_var, accumulate, mode, B, chw,
        cov_consts(eps_box), scale);
    return after_launch();
}


extern "C" int pcfa_box_delta_backward(const float* var, const float* aux_max, const float* aux_min,
                                       const float* grad_delta, float* grad_var, int accumulate,
                                       int mode, int64_t numel, float eps_box, pcfa_stream_t stream) {
    if (!var || !grad_delta || !grad_var || numel <= 0) return PCFA_E_BADARG;
    if (mode < PCFA_BOX_COV || mode > PCFA_BOX_UNIVERSAL) return PCFA_E_BADARG;
    if (mode == PCFA_BOX_JOINT && (!aux_max || !aux_min)) return PCFA_E_BADARG;
    int64_t blocks = ceil_div<int64_t>(numel, BOX_THREADS);
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    box_delta_backward_kernel<<<(int)blocks, BOX_THREADS, 0, as_stream(stream)>>>(
        var, aux_max, aux_min, grad_delta, grad_var, accumulate, mode, numel, cov_consts(eps_box));
    return after_launch();
}

extern "C" int pcfa_sumsq_partials(const float* x, int64_t numel, float* sumsq_partials,
                                   pcfa_stream_t stream) {
    if (!x || !sumsq_partials || numel <= 0) return PCFA_E_BADARG;
    sumsq_partials_kernel<<<PCFA_BOX_PARTIALS, BOX_THREADS, 0, as_stream(stream)>>>(x, numel,
                                                                                   sumsq_partials);
    return after_launch();
}

extern "C" int64_t pcfa_objective_workspace_bytes(void) {
    return (int64_t)(3 * LOSS_BLOCKS + 4) * sizeof(float);
}

extern "C" int pcfa_objective_loss(const float* flow, const float* target,
                                   const float* sumsq_partials1, const float* sumsq_partials2,
                                   float sumsq_weight1, float sumsq_weight2, float* loss_terms,
                                   float* grad_flow, void* workspace, int loss_type, int B, int H,
                                   int W, int Hp, int Wp, int pad_top, int pad_left,
                                   double numel_total, float delta_bound, float mu,
                                   pcfa_stream_t stream) {
    if (!flow || !target || !loss_terms || !workspace || B <= 0 || H <= 0 || W <= 0) return PCFA_E_BADARG;
    if (Hp < H || Wp < W || pad_top < 0 || pad_left < 0 || pad_top + H > Hp || pad_left + W > Wp)
        return PCFA_E_BADARG;
    if (loss_type < PCFA_LOSS_AEE || loss_type > PCFA_LOSS_COSIM || numel_total <= 0.0)
        return PCFA_E_BADARG;
    cudaStream_t s = as_stream(stream);
    float* ws = reinterpret_cast<float*>(workspace);
    loss_partial_kernel<<<LOSS_BLOCKS, LOSS_THREADS, 0, s>>>(flow, target, grad_flow, ws, loss_type, B,
                                                            H, W, Hp, Wp, pad_top, pad_left);
    PCFA_TRY(after_launch());
    loss_finalize_kernel<<<1, 256, 0, s>>>(ws, sumsq_partials1, sumsq_partials2, sumsq_weight1,
                                           sumsq_weight2, loss_terms, ws + 3 * LOSS_BLOCKS, loss_type,
                                           B, H, W, numel_total, delta_bound, mu);
    PCFA_TRY(after_launch());
    if (loss_type == PCFA_LOSS_COSIM && grad_flow) {
        loss_cosim_grad_kernel<<<LOSS_BLOCKS, LOSS_THREADS, 0, s>>>(
            flow, target, ws + 3 * LOSS_BLOCKS, grad_flow, B, H, W, Hp, Wp, pad_top, pad_left);
        PCFA_TRY(after_launch());
    }
    return PCFA_OK;
}
