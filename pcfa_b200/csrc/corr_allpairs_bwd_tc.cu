// Backward of the all-pairs correlation pyramid on the tensor cores (tcgen05, kind::tf32).
//
// With G_l = dL/d(level l)  ([B*N, N_l] fp32, accumulated by the lookup-backward scatters) and
// P_l = pool_l(fmap2) the adjoint of "pool the operand, then correlate" needs no fold to level 0:
//
//   pass I :  grad_fmap1[b,c,i]  = (1/sqrt C) * sum_l sum_j G_l[i,j] * P_l[b,c,j]
//   pass II:  grad_P_l[b,c,j]    = (1/sqrt C) * sum_i G_l[i,j] * fmap1[b,c,i]      (then un-pooled into grad_fmap2)
//
// (autograd of torch.matmul + F.avg_pool2d in the reference, models/raft/corr.py:25-27,52-60.)
// Both passes are "128-row block of G" x "all C channels", contracted over the other index of G, and
// read the 261 MB gradient pyramid exactly once each, straight from HBM through TMA:
//   pass I  uses G tiles K-major   (box {32 j, 128 i}: rows = queries, 128-byte swizzle rows along j);
//   pass II uses the SAME memory MN-major (box {32 j, 32 i} x4, 128B swizzle with 32-byte atoms — the only
//           MN-major layout tcgen05 takes for tf32): the transposed operand is expressed by the
//           shared-memory descriptor, nothing is transposed in memory.
// Precision: G enters the tensor core as TF32 (the hardware ignores the low 13 mantissa bits); the
// small operands (fmap1, pooled fmap2) are split into TF32 hi + lo and both products are accumulated
// in fp32, so the only rounding beyond fp32 is the 2^-11 relative truncation of G — the same
// precision class as the cuDNN TF32 convolutions that produce and consume these gradients.
// One CTA owns two 128-row blocks (two TMEM accumulators of C columns) so each streamed channel
// chunk is used twice; the contraction range is split across CTAs (RED.ADD into zeroed outputs).
#include "tc_common.cuh"
#include <math.h>

namespace pcfa {

constexpr int BW_BM = 128, BW_BK = 32, BW_STAGES = 2, BW_THREADS = 384;
constexpr int BW_A_BYTES = BW_BM * 128;      // [128 x 32 tf32]
constexpr int BW_MAX_LEVELS = 4;

struct BwMaps {
    CUtensorMap a[BW_MAX_LEVELS];    // G_l  [N_l, N, B]   pass I box {32,128,1} ; pass II box {32,32,1}
    CUtensorMap b[BW_MAX_LEVELS];    // pass I: split P_l [N_l, C, 2B] box {32,C,1} ; pass II: b[0] = split fmap1 [N, C, 2B]
};

struct BwParams {
    int B, C, N, levels, pass, splits, chunks_total, units_per_sample;
    int nl[BW_MAX_LEVELS];
    int chunk_off[BW_MAX_LEVELS + 1];    // pass I : K-chunk prefix over levels
    int unit_off[BW_MAX_LEVELS + 1];     // pass II: block-pair prefix over levels
    float* out[BW_MAX_LEVELS];           // pass I: out[0] = grad_fmap1 ; pass II: grad of P_l (out[0] = grad_fmap2)
};

__global__ void __launch_bounds__(BW_THREADS, 1)
corr_pyramid_bwd_tc_kernel(const __grid_constant__ BwMaps maps, const BwParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)P.C * 128u;                       // one [C x 32 tf32] tile
    const uint32_t stage_bytes = 2 * BW_A_BYTES + 2 * b_bytes;
    const uint32_t bars = base + BW_STAGES * stage_bytes;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * BW_STAGES, bar_done = bar_empty + 8 * BW_STAGES;
    const uint32_t tmem_slot = bar_done + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- decode the work unit
    const int split = blockIdx.x % P.splits;
    const int rest = blockIdx.x / P.splits;
    const int b = rest / P.units_per_sample;
    const int unit = rest - b * P.units_per_sample;
    int level = 0, pair = unit;
    if (P.pass == 2) {
#pragma unroll
        for (int l = 1; l < BW_MAX_LEVELS; ++l)
            if (l < P.levels && unit >= P.unit_off[l]) level = l;
        pair = unit - P.unit_off[level];
    }
    const int rows_total = (P.pass == 1) ? P.N : P.nl[level];             // extent of the blocked (M) index
    const int m0[2] = {pair * 2 * BW_BM, (pair * 2 + 1) * BW_BM};
    const int nblk = (m0[1] < rows_total) ? 2 : 1;
    const int k_begin = (int)((long long)P.chunks_total * split / P.splits);
    const int k_end = (int)((long long)P.chunks_total * (split + 1) / P.splits);
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2u * P.C) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < BW_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0 && lane == 0) {
        // ===================================================================== TMA producer
        int stage = 0; uint32_t phase = 0;
        for (int kc = k_begin; kc < k_end; ++kc) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t dst = base + stage * stage_bytes;
            const uint32_t full = bar_full + 8 * stage;
            mbar_expect_tx(full, nblk * BW_A_BYTES + 2 * b_bytes);
            if (P.pass == 1) {
                int l = 0;
#pragma unroll
                for (int i = 1; i < BW_MAX_LEVELS; ++i)
                    if (i < P.levels && kc >= P.chunk_off[i]) l = i;
                const int j0 = (kc - P.chunk_off[l]) * BW_BK;
                for (int r = 0; r < nblk; ++r) tma_load_3d(dst + r * BW_A_BYTES, &maps.a[l], full, j0, m0[r], b);
                tma_load_3d(dst + 2 * BW_A_BYTES, &maps.b[l], full, j0, 0, b);
                tma_load_3d(dst + 2 * BW_A_BYTES + b_bytes, &maps.b[l], full, j0, 0, P.B + b);
            } else {
                const int i0 = kc * BW_BK;
                for (int r = 0; r < nblk; ++r)
                    for (int g = 0; g < 4; ++g)
                        tma_load_3d(dst + r * BW_A_BYTES + g * 4096, &maps.a[level], full, m0[r] + 32 * g, i0, b);
                tma_load_3d(dst + 2 * BW_A_BYTES, &maps.b[0], full, i0, 0, b);
                tma_load_3d(dst + 2 * BW_A_BYTES + b_bytes, &maps.b[0], full, i0, 0, P.B + b);
            }
            if (++stage == BW_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && lane == 0) {
        // ===================================================================== MMA issuer
        // D[128 x C] (fp32, TMEM) += A[128 x 8] (tf32) * B[C x 8]^T ; A K-major (pass I) or MN-major (pass II)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((P.pass == 2 ? 1u : 0u) << 15) |
                               ((uint32_t)(P.C >> 3) << 17) | ((uint32_t)(BW_BM >> 4) << 24);
        int stage = 0; uint32_t phase = 0;
        for (int kc = k_begin; kc < k_end; ++kc) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sa = base + stage * stage_bytes, sb_hi = sa + 2 * BW_A_BYTES, sb_lo = sb_hi + b_bytes;
            for (int r = 0; r < nblk; ++r) {
                const uint32_t d_tmem = tmem_base + r * P.C;
#pragma unroll
                for (int kk = 0; kk < BW_BK / 8; ++kk) {
                    const uint64_t ad = (P.pass == 1) ? umma_desc_sw128(sa + r * BW_A_BYTES + kk * 32)
                                                      : umma_desc_mn_tf32(sa + r * BW_A_BYTES + kk * 1024, 4096, 512);
                    tc_mma_tf32(d_tmem, ad, umma_desc_sw128(sb_hi + kk * 32), idesc, (kc != k_begin) || (kk != 0));
                    tc_mma_tf32(d_tmem, ad, umma_desc_sw128(sb_lo + kk * 32), idesc, 1);
                }
            }
            tc_commit(bar_empty + 8 * stage);
            if (++stage == BW_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_done);
    } else if (warp >= 4) {
        // ===================================================================== epilogue (once)
        const int ew = warp & 3, r = (warp - 4) >> 2;            // TMEM lane group, accumulator block
        if (r < nblk && k_end > k_begin) {
            mbar_wait(bar_done, 0);
            tc_fence_after();
            const int m = m0[r] + ew * 32 + lane;                 // row of G's blocked index (query or cell)
            float* out = P.out[(P.pass == 1) ? 0 : level] + (long long)b * P.C * rows_total + m;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + r * P.C;
            for (int c0 = 0; c0 < P.C; c0 += 32) {
                uint32_t v[32];
                tc_ld32(taddr + c0, v);
                tc_wait_ld();
                if (m < rows_total) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < P.C) red_add(out + (long long)(c0 + j) * rows_total, __uint_as_float(v[j]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

// src fp32 [n] * scale  ->  hi = round-to-nearest TF32 (low 13 bits zero), lo = v - hi (exact in fp32)
__global__ void split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
                                  long long n, float scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = src[i] * scale;
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
        const float hf = __uint_as_float(h);
        hi[i] = hf;
        lo[i] = v - hf;
    }
}

__global__ void bw_avgpool2_kernel(const float* __restrict__ in, float* __restrict__ out, long long R, int Hi, int Wi,
                                   int Ho, int Wo) {
    const long long total = R * Ho * Wo;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % Wo);
        const int y = (int)((idx / Wo) % Ho);
        const long long r = idx / ((long long)Wo * Ho);
        const float* p = in + (r * Hi + 2 * y) * (long long)Wi + 2 * x;
        out[idx] = 0.25f * ((p[0] + p[1]) + (p[Wi] + p[Wi + 1]));
    }
}

// grad_fmap2[r, y, x] += sum_{l>=1} gP_l[r, y>>l, x>>l] * 0.25^l
struct BwUnpool { const float* g[BW_MAX_LEVELS]; int h[BW_MAX_LEVELS], w[BW_MAX_LEVELS]; int levels; };
__global__ void bw_unpool_kernel(float* __restrict__ gf2, BwUnpool a, long long R, int H, int W) {
    const long long total = R * H * W;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % W);
        const int y = (int)((idx / W) % H);
        const long long r = idx / ((long long)W * H);
        float acc = gf2[idx], sc = 1.f;
        for (int l = 1; l < a.levels; ++l) {
            sc *= 0.25f;
            const int yy = y >> l, xx = x >> l;
            if (yy < a.h[l] && xx < a.w[l]) acc += sc * a.g[l][(r * a.h[l] + yy) * (long long)a.w[l] + xx];
        }
        gf2[idx] = acc;
    }
}

struct BwWorkspace { int64_t f1_split, p_split[BW_MAX_LEVELS], pooled[BW_MAX_LEVELS], gp[BW_MAX_LEVELS], total; };

static BwWorkspace bw_workspace(int B, int C, int H, int W, int levels) {
    BwWorkspace w{};
    auto align = [](int64_t v) { return (v + 1023) & ~(int64_t)1023; };
    int64_t o = 0;
    w.f1_split = o; o = align(o + (int64_t)2 * B * C * H * W * 4);
    int h = H, ww = W;
    for (int l = 0; l < levels; ++l) {
        const int64_t n = (int64_t)B * C * h * ww * 4;
        w.p_split[l] = o; o = align(o + 2 * n);
        if (l > 0) { w.pooled[l] = o; o = align(o + n); w.gp[l] = o; o = align(o + n); }
        h /= 2; ww /= 2;
    }
    w.total = o;
    return w;
}

bool corr_pyramid_bwd_tc_supported(int B, int C, int H, int W, int levels) {
    if (C % 16 != 0 || C < 16 || C > 256 || levels < 1 || levels > BW_MAX_LEVELS || B < 1) return false;
    int h = H, w = W;
    for (int l = 0; l < levels; ++l) {
        if (h < 1 || w < 1) return false;
        if (((long long)h * w) % 4 != 0) return false;        // TMA global strides must be multiples of 16 bytes
        h /= 2; w /= 2;
    }
    return tc_encode_fn() != nullptr;
}

int64_t corr_pyramid_bwd_tc_workspace_bytes(int B, int C, int H, int W, int levels) {
    if (C % 16 != 0 || C > 256 || levels < 1 || levels > BW_MAX_LEVELS) return 0;
    return bw_workspace(B, C, H, W, levels).total;
}

static int enc3(EncodeTiledFn enc, CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
                uint32_t b1, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * 4, d0 * d1 * 4};
    cuuint32_t box[3] = {b0, b1, 1}, es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? PCFA_OK : PCFA_E_BADARG;
}

static int grid1(long long total) {
    long long b = (total + 255) / 256;
    const long long cap = (long long)kNumSMs * 8;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

int corr_pyramid_backward_tc(const float* gpyr, const float* f1, const float* f2, float* gf1, float* gf2, void* ws,
                             int64_t ws_bytes, int B, int C, int H, int W, int levels, cudaStream_t s) {
    EncodeTiledFn enc = tc_encode_fn();
    if (!enc) return PCFA_E_NODEVICE;
    const BwWorkspace wl = bw_workspace(B, C, H, W, levels);
    if (!ws || ws_bytes < wl.total || (reinterpret_cast<uintptr_t>(ws) & 15)) return PCFA_E_WORKSPACE;
    uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
    const int N = H * W;
    const PyramidLayout L = make_pyramid_layout(B, H, W, levels);
    const float alpha = 1.0f / sqrtf((float)C);

    // ---- zero the RED.ADD targets
    PCFA_CUDA_TRY(cudaMemsetAsync(gf1, 0, (size_t)B * C * N * 4, s));
    PCFA_CUDA_TRY(cudaMemsetAsync(gf2, 0, (size_t)B * C * N * 4, s));
    for (int l = 1; l < levels; ++l)
        PCFA_CUDA_TRY(cudaMemsetAsync(wsb + wl.gp[l], 0, (size_t)B * C * L.h[l] * L.w[l] * 4, s));

    // ---- operand prep: alpha * fmap1 and alpha * pool_l(fmap2), split into TF32 hi / lo planes
    {
        const long long n = (long long)B * C * N;
        float* hi = reinterpret_cast<float*>(wsb + wl.f1_split);
        split_tf32_kernel<<<grid1(n), 256, 0, s>>>(f1, hi, hi + n, n, alpha);
        PCFA_TRY(after_launch());
    }
    const float* prev = f2;
    for (int l = 0; l < levels; ++l) {
        const long long n = (long long)B * C * L.h[l] * L.w[l];
        const float* cur = prev;
        if (l > 0) {
            float* pooled = reinterpret_cast<float*>(wsb + wl.pooled[l]);
            bw_avgpool2_kernel<<<grid1(n), 256, 0, s>>>(prev, pooled, (long long)B * C, L.h[l - 1], L.w[l - 1], L.h[l], L.w[l]);
            PCFA_TRY(after_launch());
            cur = pooled;
        }
        float* hi = reinterpret_cast<float*>(wsb + wl.p_split[l]);
        split_tf32_kernel<<<grid1(n), 256, 0, s>>>(cur, hi, hi + n, n, alpha);
        PCFA_TRY(after_launch());
        prev = cur;
    }

    const int smem = BW_STAGES * (2 * BW_A_BYTES + 2 * C * 128) + 1024 + 256;
    static int smem_set = 0;
    if (smem > smem_set) {
        PCFA_CUDA_TRY(cudaFuncSetAttribute(corr_pyramid_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set = smem;
    }
    const int sms = tc_num_sms();

    // ---- pass I: grad_fmap1
    {
        BwMaps maps;
        BwParams P{};
        P.B = B; P.C = C; P.N = N; P.levels = levels; P.pass = 1;
        int off = 0;
        for (int l = 0; l < levels; ++l) {
            const int nl = L.h[l] * L.w[l];
            P.nl[l] = nl; P.chunk_off[l] = off; off += ceil_div(nl, BW_BK);
            PCFA_TRY(enc3(enc, &maps.a[l], gpyr + L.off[l], nl, N, B, BW_BK, BW_BM));
            PCFA_TRY(enc3(enc, &maps.b[l], wsb + wl.p_split[l], nl, C, 2 * B, BW_BK, C));
        }
        for (int l = levels; l <= BW_MAX_LEVELS; ++l) P.chunk_off[l] = off;
        for (int l = levels; l < BW_MAX_LEVELS; ++l) { maps.a[l] = maps.a[0]; maps.b[l] = maps.b[0]; }
        P.chunks_total = off;
        P.units_per_sample = ceil_div(ceil_div(N, BW_BM), 2);
        const int units = B * P.units_per_sample;
        P.splits = units >= sms ? 1 : sms / units;
        if (P.splits > off) P.splits = off;
        P.out[0] = gf1;
        corr_pyramid_bwd_tc_kernel<<<units * P.splits, BW_THREADS, smem, s>>>(maps, P);
        PCFA_TRY(after_launch());
    }
    // ---- pass II: grad of P_l (level 0 goes straight into grad_fmap2)
    {
        BwMaps maps;
        BwParams P{};
        P.B = B; P.C = C; P.N = N; P.levels = levels; P.pass = 2;
        int off = 0;
        for (int l = 0; l < levels; ++l) {
            const int nl = L.h[l] * L.w[l];
            P.nl[l] = nl; P.unit_off[l] = off; off += ceil_div(ceil_div(nl, BW_BM), 2);
            PCFA_TRY(enc3(enc, &maps.a[l], gpyr + L.off[l], nl, N, B, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
            P.out[l] = (l == 0) ? gf2 : reinterpret_cast<float*>(wsb + wl.gp[l]);
        }
        for (int l = levels; l <= BW_MAX_LEVELS; ++l) P.unit_off[l] = off;
        PCFA_TRY(enc3(enc, &maps.b[0], wsb + wl.f1_split, N, C, 2 * B, BW_BK, C));
        for (int l = levels; l < BW_MAX_LEVELS; ++l) maps.a[l] = maps.a[0];
        for (int l = 1; l < BW_MAX_LEVELS; ++l) maps.b[l] = maps.b[0];
        P.chunks_total = ceil_div(N, BW_BK);
        P.units_per_sample = off;
        const int units = B * off;
        P.splits = units >= sms ? 1 : sms / units;
        if (P.splits > P.chunks_total) P.splits = P.chunks_total;
        corr_pyramid_bwd_tc_kernel<<<units * P.splits, BW_THREADS, smem, s>>>(maps, P);
        PCFA_TRY(after_launch());
    }
    if (levels > 1) {
        BwUnpool u{};
        u.levels = levels;
        for (int l = 0; l < levels; ++l) {
            u.g[l] = (l == 0) ? gf2 : reinterpret_cast<const float*>(wsb + wl.gp[l]);
            u.h[l] = L.h[l]; u.w[l] = L.w[l];
        }
        const long long total = (long long)B * C * N;
        bw_unpool_kernel<<<grid1(total), 256, 0, s>>>(gf2, u, (long long)B * C, H, W);
        PCFA_TRY(after_launch());
    }
    return PCFA_OK;
}

}  // namespace pcfa
