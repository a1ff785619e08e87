// L-BFGS direction in the compact (Byrd-Nocedal-Schnabel) representation — same H_k as the two-loop recursion of
// torch.optim.LBFGS (torch/optim/lbfgs.py; the reference's optimiser, attack_PCFA.py:97,114), without its sequential
// chain.  The two-loop recursion needs 2*h dependent (dot -> axpy) steps; as one cooperative launch that is 2*h grid
// barriers, 16 us per history entry: 2.0 ms at h = 59 and 3.3 ms at h = 100 for the 2.7 M-float Sintel variables, nine
// times per outer step (scripts/attack_breakdown.py) — latency, not bandwidth (the same 4*h*n*4 bytes at 1.3 TB/s).
//
//   H g = gamma g + [S  gamma Y] [ R^-T (D + gamma Y^T Y) R^-1   -R^-T ] [ S^T g       ]
//                                [ -R^-1                           0    ] [ gamma Y^T g ]
//   R = upper triangle of S^T Y (R_ij = s_i^T y_j, i older than or equal to j), D = diag(s_i^T y_i), gamma = H_diag.
//
// Per iteration: (1) ONE pass over S and Y for u = S^T g, v = Y^T g, g^T g — all 2h+1 dot products independent;
// (2) a single-CTA kernel that maintains S^T Y and Y^T Y incrementally in double precision — the new pair's column is
// S^T y_new = S^T g - S^T g_prev, i.e. the difference of this iteration's and the previous iteration's u, so no extra
// pass — and solves the two triangular systems; (3) ONE pass  d = -gamma g - S q_top - gamma Y q_bot  that also applies
// torch's pre-update test (<g,d> is known from the small vectors) and the parameter update, and reduces max|d|.
// Traffic as before (4*h*n*4 bytes), three launches, no grid barrier.
#include "common.cuh"

namespace pcfa {

constexpr int LC_THREADS = 256;
constexpr int LC_CHUNKS = 64;              // partial sums per dot product (fixed: deterministic reduction order)
constexpr int LC_MAX_H = 128;

struct LcState {                           // one device buffer (pcfa_lbfgs_compact_workspace_bytes)
    double* SY;  double* YY;               // [m][m] by ring slot
    double* uprev; double* vprev;          // [m] by ring slot: S^T g_prev, Y^T g_prev
    float* cs; float* cy;                  // [m] by age (0 = oldest): coefficients of s_k and y_k in d
    float* cg;                             // coefficient of g
    float* partials;                       // [(2m+1)][LC_CHUNKS]
};

__host__ __device__ inline LcState lc_state(void* ws, int m) {
    LcState s;
    uint8_t* p = reinterpret_cast<uint8_t*>(ws);
    s.SY = reinterpret_cast<double*>(p); p += sizeof(double) * m * m;
    s.YY = reinterpret_cast<double*>(p); p += sizeof(double) * m * m;
    s.uprev = reinterpret_cast<double*>(p); p += sizeof(double) * m;
    s.vprev = reinterpret_cast<double*>(p); p += sizeof(double) * m;
    s.cs = reinterpret_cast<float*>(p); p += sizeof(float) * m;
    s.cy = reinterpret_cast<float*>(p); p += sizeof(float) * m;
    s.cg = reinterpret_cast<float*>(p); p += sizeof(float) * 4;
    s.partials = reinterpret_cast<float*>(p);
    return s;
}

__device__ __forceinline__ float lc_block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < LC_THREADS / 32; ++w) r += sh[w];
    return r;
}

// grid (LC_CHUNKS, m + 1): row k < num_old -> <s_k, g>, <y_k, g> over this chunk (k by age); row m -> <g, g>
__global__ void __launch_bounds__(LC_THREADS)
lbc_dots_kernel(const float* __restrict__ S, const float* __restrict__ Y, const float* __restrict__ g, const int* __restrict__ ring,
                float* __restrict__ partials, int64_t n, int m) {
    __shared__ float sh[LC_THREADS / 32];
    const int start = ring[0], num_old = ring[1];
    const int k = blockIdx.y;
    if (k < m && k >= num_old) return;
    const int64_t per = ((n + LC_CHUNKS - 1) / LC_CHUNKS + 3) & ~(int64_t)3;
    const int64_t lo = min(per * blockIdx.x, n), hi = min(lo + per, n);
    if (k == m) {
        float a = 0.f;
        for (int64_t i = lo + threadIdx.x; i < hi; i += LC_THREADS) { const float v = __ldg(g + i); a = fmaf(v, v, a); }
        a = lc_block_sum(a, sh);
        if (threadIdx.x == 0) partials[(int64_t)(2 * m) * LC_CHUNKS + blockIdx.x] = a;
        return;
    }
    const int slot = (start + k) % m;
    const float* s = S + (int64_t)slot * n;
    const float* y = Y + (int64_t)slot * n;
    float a = 0.f, b = 0.f;
    const bool vec = ((n & 3) == 0) && (((uintptr_t)s | (uintptr_t)y | (uintptr_t)g) & 15) == 0;
    if (vec) {
        for (int64_t i = lo + 4 * threadIdx.x; i + 3 < hi; i += 4 * LC_THREADS) {
            const float4 gv = __ldg(reinterpret_cast<const float4*>(g + i));
            const float4 sv = __ldg(reinterpret_cast<const float4*>(s + i)), yv = __ldg(reinterpret_cast<const float4*>(y + i));
            a = fmaf(sv.x, gv.x, fmaf(sv.y, gv.y, fmaf(sv.z, gv.z, fmaf(sv.w, gv.w, a))));
            b = fmaf(yv.x, gv.x, fmaf(yv.y, gv.y, fmaf(yv.z, gv.z, fmaf(yv.w, gv.w, b))));
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += LC_THREADS) { const float gv = __ldg(g + i); a = fmaf(__ldg(s + i), gv, a); b = fmaf(__ldg(y + i), gv, b); }
    }
    a = lc_block_sum(a, sh);
    b = lc_block_sum(b, sh);
    if (threadIdx.x == 0) {
        partials[(int64_t)(2 * k) * LC_CHUNKS + blockIdx.x] = a;
        partials[(int64_t)(2 * k + 1) * LC_CHUNKS + blockIdx.x] = b;
    }
}

// One CTA of LC_MAX_H threads.  ring = {start, num_old, accepted}; pair_scalars = {<y,s>, <y,y>} of the pair just committed.
// out_scalars = {<g,d>, max|d| (zeroed here, reduced by lbc_combine_kernel)}.
__global__ void __launch_bounds__(LC_MAX_H)
lbc_solve_kernel(LcState st, const int* __restrict__ ring, const float* __restrict__ hdiag, const float* __restrict__ pair_scalars,
                 float* __restrict__ out_scalars, int m) {
    __shared__ double u[LC_MAX_H], v[LC_MAX_H], a[LC_MAX_H], w[LC_MAX_H], red[LC_MAX_H];
    __shared__ double gg_s;
    const int t = threadIdx.x;
    const int start = ring[0], h = ring[1], accepted = ring[2];
    const int slot = (start + t) % m;                          // slot of the entry of age t
    if (t < h) {
        double su = 0.0, sv = 0.0;
        for (int c = 0; c < LC_CHUNKS; ++c) { su += (double)st.partials[(int64_t)(2 * t) * LC_CHUNKS + c]; sv += (double)st.partials[(int64_t)(2 * t + 1) * LC_CHUNKS + c]; }
        u[t] = su; v[t] = sv;
    }
    if (t == 0) {
        double s = 0.0;
        for (int c = 0; c < LC_CHUNKS; ++c) s += (double)st.partials[(int64_t)(2 * m) * LC_CHUNKS + c];
        gg_s = s;
    }
    __syncthreads();
    if (accepted && h > 0) {                                   // column / row of the newest pair (age h-1)
        const int sn = (start + h - 1) % m;
        if (t < h - 1) {
            st.SY[(int64_t)slot * m + sn] = u[t] - st.uprev[slot];             // s_t^T y_new
            const double yy = v[t] - st.vprev[slot];                            // y_t^T y_new
            st.YY[(int64_t)slot * m + sn] = yy; st.YY[(int64_t)sn * m + slot] = yy;
        } else if (t == h - 1) {
            st.SY[(int64_t)sn * m + sn] = (double)pair_scalars[0];
            st.YY[(int64_t)sn * m + sn] = (double)pair_scalars[1];
        }
    }
    if (t < h) { st.uprev[slot] = u[t]; st.vprev[slot] = v[t]; }
    __syncthreads();
    const double gamma = (double)*hdiag;
    auto R = [&](int i, int j) { return st.SY[(int64_t)((start + i) % m) * m + (start + j) % m]; };     // i <= j by age
    // a = R^-1 u  (back substitution, column oriented)
    if (t < h) a[t] = u[t];
    __syncthreads();
    for (int j = h - 1; j >= 0; --j) {
        if (t == j) a[j] = a[j] / R(j, j);
        __syncthreads();
        if (t < j) a[t] -= R(t, j) * a[j];
        __syncthreads();
    }
    // w = (D + gamma Y^T Y) a - gamma v
    if (t < h) {
        double acc = R(t, t) * a[t];
        for (int j = 0; j < h; ++j) acc += gamma * st.YY[(int64_t)slot * m + (start + j) % m] * a[j];
        w[t] = acc - gamma * v[t];
    }
    __syncthreads();
    // q_top = R^-T w  (forward substitution with the transpose), in place in w
    for (int j = 0; j < h; ++j) {
        if (t == j) w[j] = w[j] / R(j, j);
        __syncthreads();
        if (t > j && t < h) w[t] -= R(j, t) * w[j];
        __syncthreads();
    }
    // d = -gamma g - S q_top - gamma Y q_bot,  q_bot = -a ;  <g,d> = -gamma g.g - q_top.u + gamma a.v
    if (t < h) { st.cs[t] = (float)(-w[t]); st.cy[t] = (float)(gamma * a[t]); }
    red[t] = t < h ? (-w[t] * u[t] + gamma * a[t] * v[t]) : 0.0;
    __syncthreads();
    if (t == 0) {
        double s = -gamma * gg_s;
        for (int i = 0; i < h; ++i) s += red[i];
        st.cg[0] = (float)(-gamma);
        out_scalars[0] = (float)s;
        out_scalars[1] = 0.f;
    }
}

// d = cg g + sum_k cs[k] s_k + cy[k] y_k ; param += t d unless <g,d> > -tol_change ; max|d| (non-negative floats order as ints)
__global__ void __launch_bounds__(LC_THREADS)
lbc_combine_kernel(const float* __restrict__ S, const float* __restrict__ Y, const float* __restrict__ g, LcState st,
                   const int* __restrict__ ring, float* __restrict__ d, float* __restrict__ param, float t, float tol_change,
                   float* __restrict__ scalars, int64_t n, int m) {
    __shared__ float cs[LC_MAX_H], cy[LC_MAX_H];
    __shared__ float sh[LC_THREADS / 32];
    const int start = ring[0], h = ring[1];
    for (int k = threadIdx.x; k < h; k += LC_THREADS) { cs[k] = st.cs[k]; cy[k] = st.cy[k]; }
    __syncthreads();
    const float cg = st.cg[0];
    const bool update = param && !(scalars[0] > -tol_change);
    float mx = 0.f;
    const bool vec = ((n & 3) == 0) && (((uintptr_t)S | (uintptr_t)Y | (uintptr_t)g | (uintptr_t)d | (uintptr_t)param) & 15) == 0;
    if (vec) {
        const int64_t n4 = n >> 2;
        for (int64_t i = (int64_t)blockIdx.x * LC_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * LC_THREADS) {
            const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
            float4 acc = make_float4(cg * gv.x, cg * gv.y, cg * gv.z, cg * gv.w);
            for (int k = 0; k < h; ++k) {
                const int64_t row = (int64_t)((start + k) % m) * n4 + i;
                const float4 sv = __ldg(reinterpret_cast<const float4*>(S) + row), yv = __ldg(reinterpret_cast<const float4*>(Y) + row);
                const float a = cs[k], b = cy[k];
                acc.x = fmaf(a, sv.x, fmaf(b, yv.x, acc.x)); acc.y = fmaf(a, sv.y, fmaf(b, yv.y, acc.y));
                acc.z = fmaf(a, sv.z, fmaf(b, yv.z, acc.z)); acc.w = fmaf(a, sv.w, fmaf(b, yv.w, acc.w));
            }
            reinterpret_cast<float4*>(d)[i] = acc;
            mx = fmaxf(mx, fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w))));
            if (update) {
                float4 pv = reinterpret_cast<float4*>(param)[i];
                pv.x = fmaf(t, acc.x, pv.x); pv.y = fmaf(t, acc.y, pv.y); pv.z = fmaf(t, acc.z, pv.z); pv.w = fmaf(t, acc.w, pv.w);
                reinterpret_cast<float4*>(param)[i] = pv;
            }
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * LC_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * LC_THREADS) {
            float acc = cg * __ldg(g + i);
            for (int k = 0; k < h; ++k) {
                const int64_t row = (int64_t)((start + k) % m) * n + i;
                acc = fmaf(cs[k], __ldg(S + row), fmaf(cy[k], __ldg(Y + row), acc));
            }
            d[i] = acc;
            mx = fmaxf(mx, fabsf(acc));
            if (update) param[i] = fmaf(t, acc, param[i]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < LC_THREADS / 32; ++w) mx = fmaxf(mx, sh[w]);
        atomicMax(reinterpret_cast<int*>(scalars + 1), __float_as_int(mx));
    }
}

}  // namespace pcfa

using namespace pcfa;

extern "C" int64_t pcfa_lbfgs_compact_workspace_bytes(int history_capacity) {
    if (history_capacity <= 0 || history_capacity > LC_MAX_H) return 0;
    const int64_t m = history_capacity;
    return (int64_t)sizeof(double) * (2 * m * m + 2 * m) + (int64_t)sizeof(float) * (2 * m + 4) +
           (int64_t)sizeof(float) * (2 * m + 1) * LC_CHUNKS + 64;
}

// d = -H grad over ring = {start, num_old, accepted} (as written by pcfa_lbfgs_update_history), then param += t*d unless
// <grad,d> > -tol_change.  pair_scalars = {<y,s>, <y,y>} of that update; scalars_out = {<grad,d>, max|d|}.  `state` keeps
// S^T Y, Y^T Y and the previous iteration's S^T g, Y^T g (pcfa_lbfgs_compact_workspace_bytes, 8-byte aligned); it must be
// used with every history update since the ring was last reset.
extern "C" int pcfa_lbfgs_direction_compact(const float* S, const float* Y, const float* grad, const float* h_diag, float* d,
                                            const int* ring, const float* pair_scalars, float* param, float t, float tol_change,
                                            float* scalars_out, void* state, int64_t n, int history_capacity, pcfa_stream_t stream) {
    if (!S || !Y || !grad || !h_diag || !d || !ring || !pair_scalars || !scalars_out || !state || n <= 0 || history_capacity <= 0 ||
        history_capacity > LC_MAX_H || (reinterpret_cast<uintptr_t>(state) & 7))
        return PCFA_E_BADARG;
    const int m = history_capacity;
    const LcState st = lc_state(state, m);
    cudaStream_t s = as_stream(stream);
    lbc_dots_kernel<<<dim3(LC_CHUNKS, m + 1), LC_THREADS, 0, s>>>(S, Y, grad, ring, st.partials, n, m);
    PCFA_TRY(after_launch());
    lbc_solve_kernel<<<1, LC_MAX_H, 0, s>>>(st, ring, h_diag, pair_scalars, scalars_out, m);
    PCFA_TRY(after_launch());
    int64_t blocks = (n / 4 + LC_THREADS - 1) / LC_THREADS;
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    lbc_combine_kernel<<<(unsigned)blocks, LC_THREADS, 0, s>>>(S, Y, grad, st, ring, d, param, t, tol_change, scalars_out, n, m);
    return after_launch();
}
