// On-device L-BFGS direction (row f-1 of SURVEY.md section 8): the two-loop recursion of torch.optim.LBFGS.step
// (torch/optim/lbfgs.py, the reference's optimiser: attack_PCFA.py:97,114; history 100, no line search) as ONE
// cooperative launch instead of ~4*history tiny ATen launches driven from Python.
//
//   q = -g ; for i = newest..oldest: al_i = ro_i * <s_i, q> ; q -= al_i * y_i
//   r = H_diag * q ; for i = oldest..newest: be_i = ro_i * <y_i, r> ; r += (al_i - be_i) * s_i ;  d = r
//
// History lives in two ring buffers S, Y of [m][n] floats (s = t*d steps, y = gradient differences).  Every CTA owns
// a contiguous slice of the vector; a step is "partial dot over my slice -> grid barrier -> every CTA sums the per-CTA
// partials in the same order (deterministic) -> axpy over my slice".  q/r live in the output buffer d (10.8 MB for the
// Sintel pair: L2-resident), so HBM traffic is one read of s_i and one of y_i per loop step.
// Also: lbfgs_pair_kernel writes the new (s, y) pair into a ring slot and returns <y,s>, <y,y>.
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace pcfa {

constexpr int LB_THREADS = 512;
constexpr int LB_MAX_HISTORY = 128;
constexpr int LB_MAX_CTAS = 1024;

__device__ __forceinline__ float lb_block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.f;
    if (warp == 0) {
        r = lane < LB_THREADS / 32 ? sh[lane] : 0.f;
        r = warp_sum(r);
        if (lane == 0) sh[0] = r;
    }
    __syncthreads();
    r = sh[0];
    __syncthreads();
    return r;
}

// partials: [2][gridDim.x] floats (double-buffered per step).  scalars out: [0] = <g, d>, [1] = max |d|
__global__ void __launch_bounds__(LB_THREADS)
lbfgs_two_loop_kernel(const float* __restrict__ S, const float* __restrict__ Y, const float* __restrict__ ro,
                      const float* __restrict__ g, const float* __restrict__ hdiag, float* __restrict__ d,
                      float* __restrict__ partials, float* __restrict__ scalars, int64_t n, int m, int start, int num_old,
                      const int* __restrict__ ring, float* __restrict__ param, float t, float tol_change) {
    cg::grid_group grid = cg::this_grid();
    if (ring) { start = ring[0]; num_old = ring[1]; }      // history bookkeeping kept on the device (lbfgs_pair_commit_kernel)
    __shared__ float sh[LB_THREADS / 32];
    __shared__ float al[LB_MAX_HISTORY];
    const int64_t per = ((n + gridDim.x - 1) / gridDim.x + 3) & ~(int64_t)3;
    const int64_t lo = per * blockIdx.x < n ? per * blockIdx.x : n, hi = lo + per < n ? lo + per : n;
    int buf = 0;
    auto total = [&](float part) -> float {                 // same ordered sum in every CTA
        if (threadIdx.x == 0) partials[buf * gridDim.x + blockIdx.x] = part;
        grid.sync();
        float acc = 0.f;
        for (int c = threadIdx.x; c < (int)gridDim.x; c += LB_THREADS) acc += partials[buf * gridDim.x + c];
        // fixed association: per-thread strided sums, then the block tree — identical in every CTA
        acc = lb_block_sum(acc, sh);
        buf ^= 1;
        return acc;
    };
    for (int64_t i = lo + threadIdx.x; i < hi; i += LB_THREADS) d[i] = -g[i];
    __syncthreads();
    for (int k = num_old - 1; k >= 0; --k) {
        const int slot = (start + k) % m;
        const float* s = S + (int64_t)slot * n;
        const float* y = Y + (int64_t)slot * n;
        float p = 0.f;
        for (int64_t i = lo + threadIdx.x; i < hi; i += LB_THREADS) p = fmaf(__ldg(s + i), d[i], p);
        const float a = ro[slot] * total(lb_block_sum(p, sh));
        if (threadIdx.x == 0) al[k] = a;
        for (int64_t i = lo + threadIdx.x; i < hi; i += LB_THREADS) d[i] = fmaf(-a, __ldg(y + i), d[i]);
        __syncthreads();
    }
    const float hd = *hdiag;
    for (int64_t i = lo + threadIdx.x; i < hi; i += LB_THREADS) d[i] *= hd;
    __syncthreads();
    for (int k = 0; k < num_old; ++k) {
        const int slot = (start + k) % m;
        const float* s = S + (int64_t)slot * n;
        const float* y = Y + (int64_t)slot * n;
        float p = 0.f;
        for (int64_t i = lo + threadIdx.x; i < hi; i += LB_THREADS) p = fmaf(__ldg(y + i), d[i], p);
        const float be = ro[slot] * total(lb_block_sum(p, sh));
        const float c = al[k] - be;
        for (int64_t i = lo + threadIdx.x; i < hi; i += LB_THREADS) d[i] = fmaf(c, __ldg(s + i), d[i]);
        __syncthreads();
    }
    // <g, d> and max |d|
    float gd = 0.f, mx = 0.f;
    for (int64_t i = lo + threadIdx.x; i < hi; i += LB_THREADS) { const float v = d[i]; gd = fmaf(__ldg(g + i), v, gd); mx = fmaxf(mx, fabsf(v)); }
    const float gtd = total(lb_block_sum(gd, sh));
    // torch: `if gtd > -tolerance_change: break` BEFORE the parameter update — here a device-side predicate, identical in
    // every CTA (same ordered sum), so the host needs no round trip between the direction and the update
    if (param && !(gtd > -tol_change))
        for (int64_t i = lo + threadIdx.x; i < hi; i += LB_THREADS) param[i] = fmaf(t, d[i], param[i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < LB_THREADS / 32; ++w) mx = fmaxf(mx, sh[w]);
        partials[buf * gridDim.x + blockIdx.x] = mx;
    }
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float m2 = 0.f;
        for (int c = 0; c < (int)gridDim.x; ++c) m2 = fmaxf(m2, partials[buf * gridDim.x + c]);
        scalars[0] = gtd; scalars[1] = m2;
    }
}

// y = g - g_prev -> Y[slot], s = t * d -> S[slot]; partial sums of <y,s> and <y,y> per CTA (fixed grid: deterministic).
// Also g_prev <- g.  partials: [2][gridDim.x]
__global__ void __launch_bounds__(LB_THREADS)
lbfgs_pair_kernel(const float* __restrict__ g, float* __restrict__ g_prev, const float* __restrict__ d, float t,
                  float* __restrict__ s_out, float* __restrict__ y_out, float* __restrict__ partials, int64_t n) {
    __shared__ float sh[LB_THREADS / 32];
    float ys = 0.f, yy = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * LB_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * LB_THREADS) {
        const float gi = g[i];
        const float y = gi - g_prev[i], s = t * d[i];
        y_out[i] = y; s_out[i] = s;
        g_prev[i] = gi;
        ys = fmaf(y, s, ys); yy = fmaf(y, y, yy);
    }
    ys = lb_block_sum(ys, sh);
    yy = lb_block_sum(yy, sh);
    if (threadIdx.x == 0) { partials[blockIdx.x] = ys; partials[gridDim.x + blockIdx.x] = yy; }
}

// <y,s>, <y,y> of the candidate pair WITHOUT writing it (the slot it would take may still hold live history)
__global__ void __launch_bounds__(LB_THREADS)
lbfgs_pair_reduce_kernel(const float* __restrict__ g, const float* __restrict__ g_prev, const float* __restrict__ d, float t,
                         float* __restrict__ partials, int64_t n) {
    __shared__ float sh[LB_THREADS / 32];
    float ys = 0.f, yy = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * LB_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * LB_THREADS) {
        const float y = g[i] - g_prev[i], s = t * d[i];
        ys = fmaf(y, s, ys); yy = fmaf(y, y, yy);
    }
    ys = lb_block_sum(ys, sh);
    yy = lb_block_sum(yy, sh);
    if (threadIdx.x == 0) { partials[blockIdx.x] = ys; partials[gridDim.x + blockIdx.x] = yy; }
}

// torch/optim/lbfgs.py: `if ys > 1e-10:` pop the oldest pair when the history is full, append (y, s), ro, H_diag = ys / yy.
// Decided on the device from the reduced scalars; ring = {start, num_old} is double-buffered (ring_in read by every
// thread, ring_out written by one) so that the launch needs no grid barrier.  Always: g_prev <- g.
__global__ void __launch_bounds__(LB_THREADS)
lbfgs_pair_commit_kernel(const float* __restrict__ g, float* __restrict__ g_prev, const float* __restrict__ d, float t,
                         float* __restrict__ S, float* __restrict__ Y, float* __restrict__ ro, float* __restrict__ hdiag,
                         const int* __restrict__ ring_in, int* __restrict__ ring_out, const float* __restrict__ scalars,
                         int64_t n, int m) {
    const float ys = scalars[0], yy = scalars[1];
    const int start = ring_in[0], num_old = ring_in[1];
    const bool accept = ys > 1e-10f;
    const int slot = num_old < m ? (start + num_old) % m : start;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (accept) {
            ring_out[0] = num_old < m ? start : (start + 1) % m;
            ring_out[1] = num_old < m ? num_old + 1 : m;
            ro[slot] = 1.0f / ys;
            *hdiag = ys / yy;
        } else { ring_out[0] = start; ring_out[1] = num_old; }
        ring_out[2] = accept ? 1 : 0;                          // read by the compact direction (lbfgs_compact.cu)
    }
    float* s_out = S + (int64_t)slot * n;
    float* y_out = Y + (int64_t)slot * n;
    for (int64_t i = (int64_t)blockIdx.x * LB_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * LB_THREADS) {
        const float gi = g[i];
        if (accept) { y_out[i] = gi - g_prev[i]; s_out[i] = t * d[i]; }
        g_prev[i] = gi;
    }
}

__global__ void lbfgs_pair_finalize_kernel(const float* __restrict__ partials, int nblocks, float* __restrict__ out) {
    __shared__ double red[2][256];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) { a += (double)partials[i]; b += (double)partials[nblocks + i]; }
    red[0][threadIdx.x] = a; red[1][threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { red[0][threadIdx.x] += red[0][threadIdx.x + s]; red[1][threadIdx.x] += red[1][threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = (float)red[0][0]; out[1] = (float)red[1][0]; }
}

}  // namespace pcfa

using namespace pcfa;

extern "C" int64_t pcfa_lbfgs_workspace_bytes(void) { return (int64_t)(2 * LB_MAX_CTAS + 16) * (int64_t)sizeof(float); }

// scalars_out (device): [0] = <y,s>, [1] = <y,y>
extern "C" int pcfa_lbfgs_store_pair(const float* grad, float* grad_prev, const float* d, float t, float* s_slot, float* y_slot,
                                     float* scalars_out, void* workspace, int64_t n, pcfa_stream_t stream) {
    if (!grad || !grad_prev || !d || !s_slot || !y_slot || !scalars_out || !workspace || n <= 0) return PCFA_E_BADARG;
    int64_t blocks = (n + LB_THREADS - 1) / LB_THREADS;
    const int64_t cap = (int64_t)kNumSMs * 4;
    if (blocks > cap) blocks = cap;
    float* part = reinterpret_cast<float*>(workspace);
    cudaStream_t s = as_stream(stream);
    lbfgs_pair_kernel<<<(int)blocks, LB_THREADS, 0, s>>>(grad, grad_prev, d, t, s_slot, y_slot, part, n);
    PCFA_TRY(after_launch());
    lbfgs_pair_finalize_kernel<<<1, 256, 0, s>>>(part, (int)blocks, scalars_out);
    return after_launch();
}

// scalars_out (device): [0] = <grad, d>, [1] = max |d|
extern "C" int pcfa_lbfgs_direction(const float* S, const float* Y, const float* ro, const float* grad, const float* h_diag,
                                    float* d, float* scalars_out, void* workspace, int64_t n, int history_capacity, int start,
                                    int num_old, pcfa_stream_t stream) {
    if (!S || !Y || !ro || !grad || !h_diag || !d || !scalars_out || !workspace || n <= 0 || history_capacity <= 0 ||
        num_old < 0 || num_old > history_capacity || num_old > LB_MAX_HISTORY || start < 0 || start >= history_capacity)
        return PCFA_E_BADARG;
    static int max_ctas = 0;
    if (max_ctas == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        PCFA_CUDA_TRY(cudaGetDevice(&dev));
        PCFA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PCFA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbfgs_two_loop_kernel, LB_THREADS, 0));
        max_ctas = sms * (per_sm > 2 ? 2 : per_sm);
        if (max_ctas > LB_MAX_CTAS) max_ctas = LB_MAX_CTAS;
        if (max_ctas < 1) return PCFA_E_NODEVICE;
    }
    int64_t want = (n + 4 * LB_THREADS - 1) / (4 * LB_THREADS);
    int grid = (int)(want < max_ctas ? (want < 1 ? 1 : want) : max_ctas);
    float* part = reinterpret_cast<float*>(workspace);
    const int* ring = nullptr; float* param = nullptr; float t = 0.f, tol = 0.f;
    void* args[] = {(void*)&S, (void*)&Y, (void*)&ro, (void*)&grad, (void*)&h_diag, (void*)&d, (void*)&part, (void*)&scalars_out,
                    (void*)&n, (void*)&history_capacity, (void*)&start, (void*)&num_old, (void*)&ring, (void*)&param, (void*)&t, (void*)&tol};
    PCFA_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)lbfgs_two_loop_kernel, dim3(grid), dim3(LB_THREADS), args, 0, as_stream(stream)));
    return after_launch();
}

static int lb_coop_grid(int64_t n, int* grid_out) {
    static int max_ctas = 0;
    if (max_ctas == 0) {
        int dev = 0, sms = 0, per_sm = 0;
        PCFA_CUDA_TRY(cudaGetDevice(&dev));
        PCFA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PCFA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbfgs_two_loop_kernel, LB_THREADS, 0));
        max_ctas = sms * (per_sm > 2 ? 2 : per_sm);
        if (max_ctas > LB_MAX_CTAS) max_ctas = LB_MAX_CTAS;
        if (max_ctas < 1) return PCFA_E_NODEVICE;
    }
    const int64_t want = (n + 4 * LB_THREADS - 1) / (4 * LB_THREADS);
    *grid_out = (int)(want < max_ctas ? (want < 1 ? 1 : want) : max_ctas);
    return PCFA_OK;
}

// History update decided on the device: scalars_out = {<y,s>, <y,y>}; ring_in / ring_out: int[4] = {start, num_old, accepted, -}.
extern "C" int pcfa_lbfgs_update_history(const float* grad, float* grad_prev, const float* d, float t, float* S, float* Y, float* ro,
                                         float* h_diag, const int* ring_in, int* ring_out, float* scalars_out, void* workspace,
                                         int64_t n, int history_capacity, pcfa_stream_t stream) {
    if (!grad || !grad_prev || !d || !S || !Y || !ro || !h_diag || !ring_in || !ring_out || ring_in == ring_out || !scalars_out ||
        !workspace || n <= 0 || history_capacity <= 0 || history_capacity > LB_MAX_HISTORY)
        return PCFA_E_BADARG;
    int64_t blocks = (n + LB_THREADS - 1) / LB_THREADS;
    const int64_t cap = (int64_t)kNumSMs * 4;
    if (blocks > cap) blocks = cap;
    float* part = reinterpret_cast<float*>(workspace);
    cudaStream_t s = as_stream(stream);
    lbfgs_pair_reduce_kernel<<<(int)blocks, LB_THREADS, 0, s>>>(grad, grad_prev, d, t, part, n);
    PCFA_TRY(after_launch());
    lbfgs_pair_finalize_kernel<<<1, 256, 0, s>>>(part, (int)blocks, scalars_out);
    PCFA_TRY(after_launch());
    lbfgs_pair_commit_kernel<<<(int)blocks, LB_THREADS, 0, s>>>(grad, grad_prev, d, t, S, Y, ro, h_diag, ring_in, ring_out, scalars_out,
                                                               n, history_capacity);
    return after_launch();
}

// Two-loop recursion over the device-side ring, then param += t * d unless <grad, d> > -tol_change (torch's break, as a
// device predicate).  scalars_out = {<grad, d>, max |d|}.
extern "C" int pcfa_lbfgs_direction_step(const float* S, const float* Y, const float* ro, const float* grad, const float* h_diag,
                                         float* d, const int* ring, float* param, float t, float tol_change, float* scalars_out,
                                         void* workspace, int64_t n, int history_capacity, pcfa_stream_t stream) {
    if (!S || !Y || !ro || !grad || !h_diag || !d || !ring || !param || !scalars_out || !workspace || n <= 0 ||
        history_capacity <= 0 || history_capacity > LB_MAX_HISTORY)
        return PCFA_E_BADARG;
    int grid = 0;
    PCFA_TRY(lb_coop_grid(n, &grid));
    float* part = reinterpret_cast<float*>(workspace);
    int start = 0, num_old = 0;
    void* args[] = {(void*)&S, (void*)&Y, (void*)&ro, (void*)&grad, (void*)&h_diag, (void*)&d, (void*)&part, (void*)&scalars_out,
                    (void*)&n, (void*)&history_capacity, (void*)&start, (void*)&num_old, (void*)&ring, (void*)&param, (void*)&t,
                    (void*)&tol_change};
    PCFA_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)lbfgs_two_loop_kernel, dim3(grid), dim3(LB_THREADS), args, 0, as_stream(stream)));
    return after_launch();
}
