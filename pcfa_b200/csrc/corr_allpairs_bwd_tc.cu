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
// chunk is used twice; the linear (row-block pair, K-chunk) work space is cut into equal contiguous shares,
// one per SM (a CTA finishes a block pair's partial sum with RED.ADD into zeroed outputs and moves on).
#include "tc_common.cuh"
#include <math.h>
#include <stdlib.h>

namespace pcfa {

constexpr int BW_BM = 128, BW_BK = 32, BW_STAGES = 2, BW_THREADS = 384;
constexpr int BW_A_BYTES = BW_BM * 128;      // [128 x 32 tf32]
constexpr int BW_MAX_LEVELS = 4;

struct BwMaps {
    CUtensorMap a[BW_MAX_LEVELS];    // G_l  [N_l, N, B]   pass I box {32,128,1} ; pass II box {32,32,1}
    CUtensorMap b[BW_MAX_LEVELS];    // pass I: split P_l [N_l, C, 2B] box {32,C,1} ; pass II: b[0] = split fmap1 [N, C, 2B]
    CUtensorMap o[BW_MAX_LEVELS];    // 2-CTA kernel outputs [rows, C, B] box {128,16,1}: pass I o[0] = grad_fmap1 ; pass II grad P_l
};

struct BwParams {
    int B, C, N, levels, pass, chunks_total, units_per_sample;
    int terms, stages;                   // 2-CTA kernel: feature operand terms (1: TF32 hi, 2: hi + lo), ring depth
    long long work_total;                // B * units_per_sample * chunks_total  (linear (unit, K-chunk) space)
    int nl[BW_MAX_LEVELS];
    int chunk_off[BW_MAX_LEVELS + 1];    // pass I : K-chunk prefix over levels
    int unit_off[BW_MAX_LEVELS + 1];     // pass II: block-pair prefix over levels
    float* out[BW_MAX_LEVELS];           // pass I: out[0] = grad_fmap1 ; pass II: grad of P_l (out[0] = grad_fmap2)
    unsigned long long* trace;           // optional per-CTA timeline (8 globaltimer stamps), set by pcfa_debug_set_trace
    // sparse mode (CTA-pair kernel, occupancy bitmap given): the K-chunks of a unit are the set bits of its live mask
    int sparse, det, nunits, mask_words; // nunits = B * units_per_sample
    const unsigned* wl_mask;             // [nunits][mask_words]  bit k: K-chunk k of the unit has marked blocks
    const int* wl_cnt;                   // [nunits]              popcount of the mask
};

struct BwSegment { int b, level, rows_total, m0a, m0b, nblk, k0, k1, ug; long long next; };

// The CTA's share [w, w_end) of the linear work space is cut at unit boundaries into segments; every warp role
// walks the same segments.
__device__ __forceinline__ BwSegment bw_segment(long long w, long long w_end, const BwParams& P) {
    BwSegment sg;
    const long long ug = w / P.chunks_total;                       // global unit index
    sg.k0 = (int)(w - ug * P.chunks_total);
    const long long unit_end = (ug + 1) * P.chunks_total;
    sg.next = unit_end < w_end ? unit_end : w_end;
    sg.k1 = sg.k0 + (int)(sg.next - w);
    sg.b = (int)(ug / P.units_per_sample);
    const int unit = (int)(ug - (long long)sg.b * P.units_per_sample);
    int level = 0, pair = unit;
    if (P.pass == 2) {
#pragma unroll
        for (int l = 1; l < BW_MAX_LEVELS; ++l)
            if (l < P.levels && unit >= P.unit_off[l]) level = l;
        pair = unit - P.unit_off[level];
    }
    sg.level = level;
    sg.rows_total = (P.pass == 1) ? P.N : P.nl[level];
    sg.m0a = pair * 2 * BW_BM;
    sg.m0b = sg.m0a + BW_BM;
    sg.nblk = (sg.m0b < sg.rows_total) ? 2 : 1;
    return sg;
}

__global__ void __launch_bounds__(BW_THREADS, 1)
corr_pyramid_bwd_tc_kernel(const __grid_constant__ BwMaps maps, const BwParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)P.C * 128u;                       // one [C x 32 tf32] tile
    const uint32_t stage_bytes = 2 * BW_A_BYTES + 2 * b_bytes;
    const uint32_t bars = base + BW_STAGES * stage_bytes;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * BW_STAGES;
    const uint32_t bar_accfull = bar_empty + 8 * BW_STAGES, bar_accempty = bar_accfull + 8;
    const uint32_t tmem_slot = bar_accempty + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const long long w_begin = P.work_total * blockIdx.x / gridDim.x;
    const long long w_end = P.work_total * (blockIdx.x + 1) / gridDim.x;
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2u * P.C) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < BW_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_accfull, 1);
        mbar_init(bar_accempty, 8);
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
        for (long long w = w_begin; w < w_end;) {
            const BwSegment sg = bw_segment(w, w_end, P);
            for (int kc = sg.k0; kc < sg.k1; ++kc) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t dst = base + stage * stage_bytes;
                const uint32_t full = bar_full + 8 * stage;
                mbar_expect_tx(full, sg.nblk * BW_A_BYTES + 2 * b_bytes);
                if (P.pass == 1) {
                    int l = 0;
#pragma unroll
                    for (int i = 1; i < BW_MAX_LEVELS; ++i)
                        if (i < P.levels && kc >= P.chunk_off[i]) l = i;
                    const int j0 = (kc - P.chunk_off[l]) * BW_BK;
                    tma_load_3d(dst, &maps.a[l], full, j0, sg.m0a, sg.b);
                    if (sg.nblk == 2) tma_load_3d(dst + BW_A_BYTES, &maps.a[l], full, j0, sg.m0b, sg.b);
                    tma_load_3d(dst + 2 * BW_A_BYTES, &maps.b[l], full, j0, 0, sg.b);
                    tma_load_3d(dst + 2 * BW_A_BYTES + b_bytes, &maps.b[l], full, j0, 0, P.B + sg.b);
                } else {
                    const int i0 = kc * BW_BK;
                    for (int g = 0; g < 4; ++g)
                        tma_load_3d(dst + g * 4096, &maps.a[sg.level], full, sg.m0a + 32 * g, i0, sg.b);
                    if (sg.nblk == 2)
                        for (int g = 0; g < 4; ++g)
                            tma_load_3d(dst + BW_A_BYTES + g * 4096, &maps.a[sg.level], full, sg.m0b + 32 * g, i0, sg.b);
                    tma_load_3d(dst + 2 * BW_A_BYTES, &maps.b[0], full, i0, 0, sg.b);
                    tma_load_3d(dst + 2 * BW_A_BYTES + b_bytes, &maps.b[0], full, i0, 0, P.B + sg.b);
                }
                if (++stage == BW_STAGES) { stage = 0; phase ^= 1; }
            }
            w = sg.next;
        }
    } else if (warp == 1 && lane == 0) {
        // ===================================================================== MMA issuer
        // D[128 x C] (fp32, TMEM) += A[128 x 8] (tf32) * B[C x 8]^T ; A K-major (pass I) or MN-major (pass II)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((P.pass == 2 ? 1u : 0u) << 15) |
                               ((uint32_t)(P.C >> 3) << 17) | ((uint32_t)(BW_BM >> 4) << 24);
        int stage = 0; uint32_t phase = 0, accphase = 0;
        for (long long w = w_begin; w < w_end;) {
            const BwSegment sg = bw_segment(w, w_end, P);
            mbar_wait(bar_accempty, accphase ^ 1);            // epilogue has drained the previous segment
            tc_fence_after();
            for (int kc = sg.k0; kc < sg.k1; ++kc) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint32_t sa = base + stage * stage_bytes, sb_hi = sa + 2 * BW_A_BYTES, sb_lo = sb_hi + b_bytes;
                for (int r = 0; r < sg.nblk; ++r) {
                    const uint32_t d_tmem = tmem_base + r * P.C;
#pragma unroll
                    for (int kk = 0; kk < BW_BK / 8; ++kk) {
                        const uint64_t ad = (P.pass == 1) ? umma_desc_sw128(sa + r * BW_A_BYTES + kk * 32)
                                                          : umma_desc_mn_tf32(sa + r * BW_A_BYTES + kk * 1024, 4096, 512);
                        tc_mma_tf32(d_tmem, ad, umma_desc_sw128(sb_hi + kk * 32), idesc, (kc != sg.k0) || (kk != 0));
                        tc_mma_tf32(d_tmem, ad, umma_desc_sw128(sb_lo + kk * 32), idesc, 1);
                    }
                }
                tc_commit(bar_empty + 8 * stage);
                if (++stage == BW_STAGES) { stage = 0; phase ^= 1; }
            }
            tc_commit(bar_accfull);
            accphase ^= 1;
            w = sg.next;
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue (per segment)
        const int ew = warp & 3, r = (warp - 4) >> 2;            // TMEM lane group, accumulator block
        uint32_t accphase = 0;
        for (long long w = w_begin; w < w_end;) {
            const BwSegment sg = bw_segment(w, w_end, P);
            mbar_wait(bar_accfull, accphase);
            tc_fence_after();
            if (r < sg.nblk) {
                const int m = (r == 0 ? sg.m0a : sg.m0b) + ew * 32 + lane;     // query (pass I) or cell (pass II)
                float* out = P.out[(P.pass == 1) ? 0 : sg.level] + (long long)sg.b * P.C * sg.rows_total + m;
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + r * P.C;
                for (int c0 = 0; c0 < P.C; c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(taddr + c0, v);
                    tc_wait_ld();
                    if (m < sg.rows_total) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c0 + j < P.C) red_add(out + (long long)(c0 + j) * sg.rows_total, __uint_as_float(v[j]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accempty);
            accphase ^= 1;
            w = sg.next;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

// ------------------------------------------------------------------------------------ CTA-pair variant
// cta_group::2: one MMA covers M = 256 rows (one 128-row block per CTA) x N = C channels, whose operand is split
// across the pair (each CTA streams C/2 channel rows).  With two accumulators per CTA a pair owns FOUR row blocks
// per streamed channel chunk: per MMA cycle each CTA streams 64 KB / 2048 cyc instead of 96 KB / 2048 cyc, and
// the smaller stage makes room for a 3-deep ring.  Barrier protocol as in corr_pyramid_tc2_kernel.
__device__ __forceinline__ void bw_stamp(unsigned long long* tr, int slot) {
    if (tr) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tr[(size_t)blockIdx.x * 8 + slot] = t; }
}
constexpr int BW2_MAX_STAGES = 4;
constexpr int BW2_STG_BYTES = 32 * BW_BM * 4;     // epilogue staging tile for the bulk reduce-add

// Walks the CTA pair's share of the linear (unit, K-chunk) space.  Dense: every unit has chunks_total chunks.  Sparse: unit
// ug has the live chunks [pre[ug], pre[ug+1]) (prefix sums of wl_cnt in shared memory); k0/k1 are then ORDINALS among the
// unit's live chunks, which the TMA producer maps to chunk indices by walking the set bits of the unit's mask.
__device__ __forceinline__ BwSegment bw2_segment(long long w, long long w_end, const BwParams& P, int rank, const int* pre,
                                                 int& cursor) {
    BwSegment sg;
    long long ug;
    if (P.sparse) {
        while (pre[cursor + 1] <= w) ++cursor;
        ug = cursor;
        sg.k0 = (int)(w - pre[cursor]);
        const long long unit_end = pre[cursor + 1];
        sg.next = unit_end < w_end ? unit_end : w_end;
    } else {
        ug = w / P.chunks_total;
        sg.k0 = (int)(w - ug * P.chunks_total);
        const long long unit_end = (ug + 1) * P.chunks_total;
        sg.next = unit_end < w_end ? unit_end : w_end;
    }
    sg.k1 = sg.k0 + (int)(sg.next - w);
    sg.b = (int)(ug / P.units_per_sample);
    const int unit = (int)(ug - (long long)sg.b * P.units_per_sample);
    int level = 0, quad = unit;
    if (P.pass == 2) {
#pragma unroll
        for (int l = 1; l < BW_MAX_LEVELS; ++l)
            if (l < P.levels && unit >= P.unit_off[l]) level = l;
        quad = unit - P.unit_off[level];
    }
    sg.level = level;
    sg.rows_total = (P.pass == 1) ? P.N : P.nl[level];
    sg.m0a = (quad * 4 + rank * 2) * BW_BM;          // this CTA's two row blocks (may lie past rows_total: zero-filled)
    sg.m0b = sg.m0a + BW_BM;
    sg.nblk = 2;
    sg.ug = (int)ug;
    return sg;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BW_THREADS, 1)
corr_pyramid_bwd_tc2_kernel(const __grid_constant__ BwMaps maps, const BwParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bh_bytes = (uint32_t)(P.C / 2) * 128u;               // this CTA's half of a [C x 32 tf32] tile
    const uint32_t stage_bytes = 2 * BW_A_BYTES + (uint32_t)P.terms * bh_bytes;
    const int nst = P.stages;
    const uint32_t stg = base + nst * stage_bytes;                      // 2 x [32 channels][128 rows] fp32 staging tiles
    const uint32_t bars = stg + 2 * BW2_STG_BYTES;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * BW2_MAX_STAGES;
    const uint32_t bar_accfull = bar_empty + 8 * BW2_MAX_STAGES, bar_accempty = bar_accfull + 8;
    const uint32_t tmem_slot = bar_accempty + 8;
    int* pre = reinterpret_cast<int*>(smem_raw + (tmem_slot + 8 - smem_u32(smem_raw)));     // sparse: [nunits + 1] prefix sums
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const long long ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2u * P.C) tmem_cols <<= 1;

    // Programmatic dependent launch: pass II does not read pass I's output, so its CTAs may take over an SM as soon as
    // the pass-I CTA there has exited (the tail of one pass overlaps the head of the next).  No-op without a dependent.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) bw_stamp(P.trace, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_accfull, 1);
        mbar_init(bar_accempty, 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    if (P.sparse && warp == 3) {                                        // prefix sums of the units' live-chunk counts
        int run = 0;
        for (int u0 = 0; u0 < P.nunits; u0 += 32) {
            int v = (u0 + lane < P.nunits) ? P.wl_cnt[u0 + lane] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
            if (u0 + lane < P.nunits) pre[u0 + lane + 1] = run + v;
            run += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) pre[0] = 0;
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (threadIdx.x == 0) bw_stamp(P.trace, 1);
    // this pair's share: equal counts of (live) chunks; deterministic mode cuts at unit boundaries (one RED.ADD per output)
    long long w_begin, w_end;
    if (P.sparse) {
        if (P.det) { w_begin = pre[(long long)P.nunits * cid / ncl]; w_end = pre[(long long)P.nunits * (cid + 1) / ncl]; }
        else { const long long total = pre[P.nunits]; w_begin = total * cid / ncl; w_end = total * (cid + 1) / ncl; }
    } else { w_begin = P.work_total * cid / ncl; w_end = P.work_total * (cid + 1) / ncl; }

    if (warp == 0 && lane == 0) {
        // ===================================================================== TMA producer (both CTAs)
        int stage = 0, cursor = 0; uint32_t phase = 0;
        for (long long w = w_begin; w < w_end;) {
            const BwSegment sg = bw2_segment(w, w_end, P, rank, pre, cursor);
            const unsigned* mk = P.wl_mask + (size_t)sg.ug * P.mask_words;
            int wi = 0; unsigned live = 0;
            if (P.sparse) {                                         // position on the k0-th set bit of the unit's mask
                int skip = sg.k0;
                live = __ldg(mk);
                for (int c = __popc(live); skip >= c; c = __popc(live)) { skip -= c; live = __ldg(mk + ++wi); }
                while (skip-- > 0) live &= live - 1;
            }
            for (int ord = sg.k0; ord < sg.k1; ++ord) {
                int kc = ord;
                if (P.sparse) {
                    while (!live) live = __ldg(mk + ++wi);
                    kc = wi * 32 + __ffs(live) - 1;
                    live &= live - 1;
                }
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t dst = base + stage * stage_bytes;
                const uint32_t full = leader_bar(bar_full + 8 * stage);
                if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * stage_bytes);       // both CTAs' bytes
                const int crow = rank * (P.C / 2);
                if (P.pass == 1) {
                    int l = 0;
#pragma unroll
                    for (int i = 1; i < BW_MAX_LEVELS; ++i)
                        if (i < P.levels && kc >= P.chunk_off[i]) l = i;
                    const int j0 = (kc - P.chunk_off[l]) * BW_BK;
                    tma2_load_3d(dst, &maps.a[l], full, j0, sg.m0a, sg.b);
                    tma2_load_3d(dst + BW_A_BYTES, &maps.a[l], full, j0, sg.m0b, sg.b);
                    tma2_load_3d(dst + 2 * BW_A_BYTES, &maps.b[l], full, j0, crow, sg.b);
                    if (P.terms == 2) tma2_load_3d(dst + 2 * BW_A_BYTES + bh_bytes, &maps.b[l], full, j0, crow, P.B + sg.b);
                } else {
                    const int i0 = kc * BW_BK;
                    for (int g = 0; g < 4; ++g) {
                        tma2_load_3d(dst + g * 4096, &maps.a[sg.level], full, sg.m0a + 32 * g, i0, sg.b);
                        tma2_load_3d(dst + BW_A_BYTES + g * 4096, &maps.a[sg.level], full, sg.m0b + 32 * g, i0, sg.b);
                    }
                    tma2_load_3d(dst + 2 * BW_A_BYTES, &maps.b[0], full, i0, crow, sg.b);
                    if (P.terms == 2) tma2_load_3d(dst + 2 * BW_A_BYTES + bh_bytes, &maps.b[0], full, i0, crow, P.B + sg.b);
                }
                if (++stage == nst) { stage = 0; phase ^= 1; }
            }
            w = sg.next;
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ===================================================================== MMA issuer (leader only)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((P.pass == 2 ? 1u : 0u) << 15) |
                               ((uint32_t)(P.C >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        int stage = 0, cursor = 0; uint32_t phase = 0, accphase = 0;
        for (long long w = w_begin; w < w_end;) {
            const BwSegment sg = bw2_segment(w, w_end, P, 0, pre, cursor);
            mbar_wait(bar_accempty, accphase ^ 1);
            tc_fence_after();
            for (int kc = sg.k0; kc < sg.k1; ++kc) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (w == w_begin && kc == sg.k0) bw_stamp(P.trace, 2);
                const uint32_t sa = base + stage * stage_bytes, sb_hi = sa + 2 * BW_A_BYTES, sb_lo = sb_hi + bh_bytes;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint32_t d_tmem = tmem_base + r * P.C;
#pragma unroll
                    for (int kk = 0; kk < BW_BK / 8; ++kk) {
                        const uint64_t ad = (P.pass == 1) ? umma_desc_sw128(sa + r * BW_A_BYTES + kk * 32)
                                                          : umma_desc_mn_tf32(sa + r * BW_A_BYTES + kk * 1024, 4096, 512);
                        tc2_mma_tf32(d_tmem, ad, umma_desc_sw128(sb_hi + kk * 32), idesc, (kc != sg.k0) || (kk != 0));
                        if (P.terms == 2) tc2_mma_tf32(d_tmem, ad, umma_desc_sw128(sb_lo + kk * 32), idesc, 1);
                    }
                }
                tc2_commit_mc(bar_empty + 8 * stage);
                if (++stage == nst) { stage = 0; phase ^= 1; }
            }
            tc2_commit_mc(bar_accfull);
            accphase ^= 1;
            w = sg.next;
        }
        bw_stamp(P.trace, 3);
    } else if (warp >= 4) {
        // ===================================================================== epilogue (per segment, both CTAs)
        // The SM issues REDs at ~1.3 cycles per lane (64 K lanes per CTA and segment would cost ~45 us, fully exposed
        // at the end of the pass).  Instead the four warps of a row block stage [16 channels][128 rows] half tiles in shared
        // memory (lanes = consecutive rows: conflict-free) and one thread issues a bulk tensor reduce-add; rows past
        // the end of the tensor are clipped by the TMA unit.
        const int ew = warp & 3, r = (warp - 4) >> 2;
        const uint32_t my_stg = stg + r * BW2_STG_BYTES;
        const bool issuer = (ew == 0 && lane == 0);
        uint32_t accphase = 0, gg = 0;
        int cursor = 0;
        for (long long w = w_begin; w < w_end;) {
            const BwSegment sg = bw2_segment(w, w_end, P, rank, pre, cursor);
            mbar_wait(bar_accfull, accphase);
            tc_fence_after();
            if (warp == 4 && lane == 0) bw_stamp(P.trace, sg.next >= w_end ? 5 : 4);
            const int m0 = (r == 0 ? sg.m0a : sg.m0b);
            const bool live = m0 < sg.rows_total;                        // uniform over the four warps of this block
            const CUtensorMap* omap = &maps.o[(P.pass == 1) ? 0 : sg.level];
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + r * P.C;
            for (int c0 = 0; c0 < P.C; c0 += 32) {
                uint32_t v[32];
                tc_ld32(taddr + c0, v);
                tc_wait_ld();
                if (c0 + 32 >= P.C) {                                   // accumulator fully read: release it to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta0(bar_accempty);
                }
                if (live) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {                       // two 16-channel half tiles, double-buffered
                        const uint32_t buf = my_stg + h * (BW2_STG_BYTES / 2);
                        if (gg >= 2) {
                            if (issuer) tma_wait_group_read1();         // the reduction that last used `buf` has read it
                            named_bar_sync(1 + r, 128);
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(buf + (uint32_t)((j * BW_BM + ew * 32 + lane) * 4)), "r"(v[16 * h + j]) : "memory");
                        fence_proxy_async_smem();
                        named_bar_sync(1 + r, 128);
                        if (issuer) {
                            tma_reduce_add_3d(omap, buf, m0, c0 + 16 * h, sg.b);
                            tma_commit_group();
                        }
                        ++gg;
                    }
                }
            }
            accphase ^= 1;
            w = sg.next;
        }
        if (issuer) tma_wait_group0();                                   // all reductions performed before the CTA exits
        if (warp == 4 && lane == 0) bw_stamp(P.trace, 6);
    }
    tc_fence_before();
    cluster_sync_all();
    if (threadIdx.x == 0) bw_stamp(P.trace, 7);
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

// Fused backward prep, one launch: TF32 hi/lo planes of alpha*fmap1 and of alpha*pool_l(fmap2) for every level
// (successive 2x2 floor pooling), and zero-fill of every RED.ADD target (grad_fmap1, grad_fmap2, grad P_l).
// CTA = 32 channels x (8 rows x 32 cols) of level 0; blockIdx.z selects (which fmap, sample, channel block).
struct BwPrepArgs {
    float* f1_hi; long long f1_plane;                       // [2][B][C][N]
    float* p_hi[BW_MAX_LEVELS]; long long p_plane[BW_MAX_LEVELS];   // [2][B][C][N_l]
    float* zero1; float* zero2; float* zero_gp[BW_MAX_LEVELS];     // grad_fmap1, grad_fmap2, grad P_l (l >= 1)
    int h[BW_MAX_LEVELS], w[BW_MAX_LEVELS];
    int levels, B, C, write_lo;
    float alpha;
    // sparse mode: per-unit live masks of both passes, built from the occupancy bitmap by extra CTAs of the prep launch
    const unsigned* occ; OccLayout OL;
    unsigned* mask1; int* cnt1; unsigned* mask2; int* cnt2;
    int units1, units2, words1, words2;
    int unit_off2[BW_MAX_LEVELS + 1];
};
__device__ __forceinline__ void tf32_split(float v, float* hi, float* lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    *hi = __uint_as_float(h);
    *lo = v - *hi;
}
constexpr int BP_CH = 8;        // channels per CTA of the pooled path: 4x the CTAs of a 32-channel tile, all loads in flight
template <int LVL>
__device__ __forceinline__ void bw_prep_drain(const float* __restrict__ t, int cstride, int rstride, const BwPrepArgs& a,
                                              int b, int c0, int nc, int y0, int x0) {
    constexpr int ww = 32 >> LVL, cells = (8 >> LVL) * ww;
    const int Hl = a.h[LVL], Wl = a.w[LVL];
    for (int e = threadIdx.x; e < BP_CH * cells; e += 256) {
        const int c = e / cells, cell = e % cells, r = cell / ww, q = cell % ww;      // powers of two: shifts
        const int yy = (y0 >> LVL) + r, xx = (x0 >> LVL) + q;
        if (yy >= Hl || xx >= Wl || c >= nc) continue;
        const long long o = (((long long)b * a.C + c0 + c) * Hl + yy) * Wl + xx;
        float hi, lo;
        tf32_split(t[c * cstride + r * rstride + q], &hi, &lo);
        a.p_hi[LVL][o] = hi;
        if (a.write_lo) a.p_hi[LVL][a.p_plane[LVL] + o] = lo;
        if (LVL > 0) a.zero_gp[LVL][o] = 0.f;
    }
}

// One CTA per (sample, unit) of either pass.  Pass I unit u = queries [512u, +512) = bitmap rows [16u, +16): its mask over
// the K-chunks (32-cell chunks of all levels) is the OR of those rows.  Pass II unit (level l, quad) = cells [512 quad, +512)
// of level l = bitmap columns [chunk_off[l] + 16 quad, +16): its mask over the K-chunks (32-query groups) has bit g set
// when row g has any of those columns.
__device__ __forceinline__ void bw_mask_block(const BwPrepArgs& a, int id) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const int per = a.units1 + a.units2;
    const int b = id / per, u = id - b * per;
    const unsigned* ob = a.occ + (long long)b * a.OL.qgroups * a.OL.words;
    int local = 0;
    if (u < a.units1) {
        unsigned* mo = a.mask1 + ((long long)b * a.units1 + u) * a.words1;
        const int g0 = u * 16, g1 = min(g0 + 16, a.OL.qgroups);
        for (int wi = threadIdx.x; wi < a.words1; wi += 256) {
            unsigned m = 0;
            for (int g = g0; g < g1; ++g) m |= __ldg(ob + g * a.OL.words + wi);
            mo[wi] = m;
            local += __popc(m);
        }
    } else {
        const int v = u - a.units1;
        int l = 0;
#pragma unroll
        for (int i = 1; i < BW_MAX_LEVELS; ++i)
            if (i < a.levels && v >= a.unit_off2[i]) l = i;
        const int c0 = a.OL.chunk_off[l] + 16 * (v - a.unit_off2[l]), c1 = min(c0 + 16, a.OL.chunk_off[l + 1]);   // [c0, c1)
        const int w0 = c0 >> 5, w1 = (c1 - 1) >> 5;
        const unsigned lo = 0xffffffffu << (c0 & 31), hi = 0xffffffffu >> (31 - ((c1 - 1) & 31));
        const unsigned mA = (w0 == w1) ? (lo & hi) : lo, mB = (w0 == w1) ? 0u : hi;
        unsigned* mo = a.mask2 + ((long long)b * a.units2 + v) * a.words2;
        for (int g = threadIdx.x; g < a.words2 * 32; g += 256) {
            bool on = false;
            if (g < a.OL.qgroups) {
                on = (__ldg(ob + g * a.OL.words + w0) & mA) != 0;
                if (mB) on = on || (__ldg(ob + g * a.OL.words + w1) & mB) != 0;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, on);
            if ((threadIdx.x & 31) == 0) { mo[g >> 5] = bal; local += __popc(bal); }
        }
    }
    if (local) atomicAdd(&cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (u < a.units1) a.cnt1[b * a.units1 + u] = cnt;
        else a.cnt2[b * a.units2 + (u - a.units1)] = cnt;
    }
}

// blockIdx.x < flat_blocks: elementwise part (fmap1 -> TF32 planes, zero-fill of grad_fmap1 / grad_fmap2), 4 elements
// per thread (128-bit accesses when `vec`).  Remaining CTAs: BP_CH channels x (8 rows x 32 cols) of fmap2, pooled.
__global__ void __launch_bounds__(256)
bw_prep_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const BwPrepArgs a, int flat_blocks, int vec,
               int nx, int ny, int first_mask_block) {
    const int H = a.h[0], W = a.w[0];
    const long long hw = (long long)H * W;
    if ((int)blockIdx.x >= first_mask_block) { bw_mask_block(a, (int)blockIdx.x - first_mask_block); return; }
    if ((int)blockIdx.x < flat_blocks) {
        const long long n = (long long)a.B * a.C * hw;
        const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
        if (i >= n) return;
        if (vec) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(f1 + i));
            float4 h, l;
            tf32_split(v.x * a.alpha, &h.x, &l.x); tf32_split(v.y * a.alpha, &h.y, &l.y);
            tf32_split(v.z * a.alpha, &h.z, &l.z); tf32_split(v.w * a.alpha, &h.w, &l.w);
            *reinterpret_cast<float4*>(a.f1_hi + i) = h;
            if (a.write_lo) *reinterpret_cast<float4*>(a.f1_hi + a.f1_plane + i) = l;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(a.zero1 + i) = z;
            *reinterpret_cast<float4*>(a.zero2 + i) = z;
        } else {
            for (long long k = i; k < i + 4 && k < n; ++k) {
                float hi, lo;
                tf32_split(__ldg(f1 + k) * a.alpha, &hi, &lo);
                a.f1_hi[k] = hi;
                if (a.write_lo) a.f1_hi[a.f1_plane + k] = lo;
                a.zero1[k] = 0.f;
                a.zero2[k] = 0.f;
            }
        }
        return;
    }
    __shared__ float t0[BP_CH * (8 * 33 + 1)];       // odd per-channel strides: conflict-free fills and drains
    __shared__ float t1[BP_CH * (4 * 17 + 1)];
    __shared__ float t2[BP_CH * (2 * 9 + 1)];
    __shared__ float t3[BP_CH * 5];
    constexpr int S0 = 8 * 33 + 1, S1 = 4 * 17 + 1, S2 = 2 * 9 + 1, S3 = 5;
    int pb = (int)blockIdx.x - flat_blocks;
    const int x0 = (pb % nx) * 32; pb /= nx;
    const int y0 = (pb % ny) * 8; pb /= ny;
    const int cblocks = ceil_div(a.C, BP_CH);
    const int b = pb / cblocks, c0 = (pb % cblocks) * BP_CH;
    const int nc = (a.C - c0) < BP_CH ? (a.C - c0) : BP_CH;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    {
        const int y = y0 + ty, x = x0 + tx;
        const bool inb = y < H && x < W;
        const float* p = f2 + ((long long)b * a.C + c0) * hw + (long long)y * W + x;
        float v[BP_CH];
#pragma unroll
        for (int c = 0; c < BP_CH; ++c) v[c] = (inb && c < nc) ? __ldg(p + c * hw) * a.alpha : 0.f;
#pragma unroll
        for (int c = 0; c < BP_CH; ++c) t0[c * S0 + ty * 33 + tx] = v[c];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < BP_CH * 64; e += 256) {
        const int c = e >> 6, r = (e >> 4) & 3, q = e & 15;
        const float* s0 = t0 + c * S0 + (2 * r) * 33 + 2 * q;
        t1[c * S1 + r * 17 + q] = 0.25f * ((s0[0] + s0[1]) + (s0[33] + s0[34]));
    }
    __syncthreads();
    if (threadIdx.x < BP_CH * 16) {
        const int e = threadIdx.x, c = e >> 4, r = (e >> 3) & 1, q = e & 7;
        const float* s1 = t1 + c * S1 + (2 * r) * 17 + 2 * q;
        t2[c * S2 + r * 9 + q] = 0.25f * ((s1[0] + s1[1]) + (s1[17] + s1[18]));
    }
    __syncthreads();
    if (threadIdx.x < BP_CH * 4) {
        const int c = threadIdx.x >> 2, q = threadIdx.x & 3;
        const float* s2 = t2 + c * S2 + 2 * q;
        t3[c * S3 + q] = 0.25f * ((s2[0] + s2[1]) + (s2[9] + s2[10]));
    }
    __syncthreads();
    bw_prep_drain<0>(t0, S0, 33, a, b, c0, nc, y0, x0);
    if (a.levels > 1) bw_prep_drain<1>(t1, S1, 17, a, b, c0, nc, y0, x0);
    if (a.levels > 2) bw_prep_drain<2>(t2, S2, 9, a, b, c0, nc, y0, x0);
    if (a.levels > 3) bw_prep_drain<3>(t3, S3, 5, a, b, c0, nc, y0, x0);
}

// grad_fmap2[r, y, x] += sum_{l>=1} gP_l[r, y>>l, x>>l] * 0.25^l   (adjoint of the successive floor pooling).
// grid: (ceil(W/128), H, R) with 128 threads: no index divisions, coalesced rows.
struct BwUnpool { const float* g[BW_MAX_LEVELS]; int h[BW_MAX_LEVELS], w[BW_MAX_LEVELS]; int levels; };
// CTA = (plane r, 8 rows); thread = 4 consecutive x of one row (128-bit accesses when W % 4 == 0): no index divisions.
__global__ void __launch_bounds__(256)
bw_unpool_kernel(float* __restrict__ gf2, const BwUnpool a, int H, int W, int vec) {
    const long long r = blockIdx.y;
    const int xq = ceil_div(W, 4);                       // 4-wide column groups per row
    for (int item = threadIdx.x; item < 8 * xq; item += 256) {
        const int y = blockIdx.x * 8 + item / xq, x = (item % xq) * 4;
        if (y >= H) break;
        float* p = gf2 + (r * H + y) * W + x;
        float v[4];
        if (vec) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (x + k < W) ? p[k] : 0.f;
        }
        float sc = 1.f;
#pragma unroll
        for (int l = 1; l < BW_MAX_LEVELS; ++l) {
            if (l >= a.levels) break;
            sc *= 0.25f;
            const int yy = y >> l;
            if (yy >= a.h[l]) continue;
            const float* gl = a.g[l] + (r * a.h[l] + yy) * a.w[l];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xx = (x + k) >> l;
                if (xx < a.w[l]) v[k] = fmaf(sc, __ldg(gl + xx), v[k]);
            }
        }
        if (vec) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (x + k < W) p[k] = v[k];
        }
    }
}

struct BwWorkspace {
    int64_t f1_split, p_split[BW_MAX_LEVELS], pooled[BW_MAX_LEVELS], gp[BW_MAX_LEVELS], total;
    int64_t mask1, cnt1, mask2, cnt2;            // sparse mode work lists
    int units1, units2, words1, words2, unit_off2[BW_MAX_LEVELS + 1];
};

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
    {   // sparse mode (CTA-pair kernel: 4 row blocks = 512 rows per unit)
        const int N = H * W;
        const OccLayout OL = make_occ_layout(H, W, levels);
        w.units1 = ceil_div(ceil_div(N, BW_BM), 4);
        int off = 0, hh = H, w2 = W;
        for (int l = 0; l < levels; ++l) { w.unit_off2[l] = off; off += ceil_div(ceil_div(hh * w2, BW_BM), 4); hh /= 2; w2 /= 2; }
        for (int l = levels; l <= BW_MAX_LEVELS; ++l) w.unit_off2[l] = off;
        w.units2 = off;
        w.words1 = OL.words; w.words2 = ceil_div(OL.qgroups, 32);
        w.mask1 = o; o = align(o + (int64_t)B * w.units1 * w.words1 * 4);
        w.cnt1 = o;  o = align(o + (int64_t)B * w.units1 * 4);
        w.mask2 = o; o = align(o + (int64_t)B * w.units2 * w.words2 * 4);
        w.cnt2 = o;  o = align(o + (int64_t)B * w.units2 * 4);
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

static unsigned long long* g_bw_trace = nullptr;
void corr_pyramid_bwd_set_trace(void* p) { g_bw_trace = reinterpret_cast<unsigned long long*>(p); }

int corr_pyramid_backward_tc(const float* gpyr, const float* f1, const float* f2, float* gf1, float* gf2, void* ws,
                             int64_t ws_bytes, int B, int C, int H, int W, int levels, cudaStream_t s, int two_cta,
                             const unsigned* occ) {
    if (C % 32 != 0) two_cta = 0;                       // each CTA of a pair streams C/2 channel rows
    EncodeTiledFn enc = tc_encode_fn();
    if (!enc) return PCFA_E_NODEVICE;
    const BwWorkspace wl = bw_workspace(B, C, H, W, levels);
    if (!ws || ws_bytes < wl.total || (reinterpret_cast<uintptr_t>(ws) & 15)) return PCFA_E_WORKSPACE;
    uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
    const int N = H * W;
    const PyramidLayout L = make_pyramid_layout(B, H, W, levels);
    // Feature operand terms.  The gradient pyramid enters the tensor core truncated to TF32 (relative error uniform in
    // [0, 2^-10)), which dominates the result error; the lo term of the small operands then buys ~10 % accuracy for 2x
    // the tensor work, so the CTA-pair kernel uses hi only.  The truncation's mean relative shrink (2^-11 / (2 ln 2)
    // ... 2^-11 ln 2 for Benford ... uniform mantissas, ~3.5e-4) is folded into alpha, which leaves the zero-mean
    // part of the error only: measured rel-L2 vs fp32 3.0e-4, below the 2-term result without it (4.1e-4).
    static const int env_terms = [] { const char* e = getenv("PCFA_BWD_TERMS"); return e ? atoi(e) : 1; }();
    static const int env_debias = [] { const char* e = getenv("PCFA_BWD_DEBIAS"); return e ? atoi(e) : 1; }();
    const int terms = (two_cta && env_terms == 1) ? 1 : 2;
    const float alpha = (1.0f / sqrtf((float)C)) * ((terms == 1 && env_debias) ? (1.0f + 3.5e-4f) : 1.0f);

    const int stage2 = 2 * BW_A_BYTES + terms * (C / 2) * 128;
    const int smem_fixed = 1024 + 256 + 2 * BW2_STG_BYTES;
    int stages2 = (227 * 1024 - smem_fixed) / stage2;
    if (stages2 > BW2_MAX_STAGES) stages2 = BW2_MAX_STAGES;
    // Sparse mode needs the prefix sums of the units' live-chunk counts in shared memory; it is used when they fit without
    // costing a ring stage (B * 20 units at 55x128: up to B = 22 with C = 256).
    static const int env_sparse = [] { const char* e = getenv("PCFA_BWD_SPARSE"); return e ? atoi(e) : 1; }();
    const int max_units = B * (wl.units1 > wl.units2 ? wl.units1 : wl.units2);
    const int pre_bytes = 4 * (max_units + 2);
    const bool sparse = occ && two_cta && env_sparse && levels <= BW_MAX_LEVELS &&
                        (227 * 1024 - smem_fixed - pre_bytes) / stage2 >= stages2;

    // ---- one launch: operand prep (alpha folded in, pooling, TF32 hi/lo split) + zero-fill of the RED.ADD targets
    // (+ in sparse mode the per-unit live masks of both passes)
    {
        BwPrepArgs pa{};
        pa.levels = levels; pa.B = B; pa.C = C; pa.alpha = alpha; pa.write_lo = terms == 2;
        pa.f1_hi = reinterpret_cast<float*>(wsb + wl.f1_split);
        pa.f1_plane = (long long)B * C * N;
        pa.zero1 = gf1; pa.zero2 = gf2;
        for (int l = 0; l < levels; ++l) {
            pa.p_hi[l] = reinterpret_cast<float*>(wsb + wl.p_split[l]);
            pa.p_plane[l] = (long long)B * C * L.h[l] * L.w[l];
            pa.zero_gp[l] = (l == 0) ? nullptr : reinterpret_cast<float*>(wsb + wl.gp[l]);
            pa.h[l] = L.h[l]; pa.w[l] = L.w[l];
        }
        const long long n = (long long)B * C * N;
        const long long flat_blocks = (n + 1023) / 1024;
        const int nx = ceil_div(W, 32), ny = ceil_div(H, 8);
        const long long pooled_blocks = (long long)nx * ny * B * ceil_div(C, BP_CH);
        if (flat_blocks + pooled_blocks > 0x7fffffffLL) return PCFA_E_TOOLARGE;
        const uintptr_t al = reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(gf1) |
                             reinterpret_cast<uintptr_t>(gf2) | reinterpret_cast<uintptr_t>(pa.f1_hi);
        const int vec = ((al & 15) == 0 && n % 4 == 0 && pa.f1_plane % 4 == 0) ? 1 : 0;
        long long mask_blocks = 0;
        if (sparse) {
            pa.occ = occ; pa.OL = make_occ_layout(H, W, levels);
            pa.mask1 = reinterpret_cast<unsigned*>(wsb + wl.mask1); pa.cnt1 = reinterpret_cast<int*>(wsb + wl.cnt1);
            pa.mask2 = reinterpret_cast<unsigned*>(wsb + wl.mask2); pa.cnt2 = reinterpret_cast<int*>(wsb + wl.cnt2);
            pa.units1 = wl.units1; pa.units2 = wl.units2; pa.words1 = wl.words1; pa.words2 = wl.words2;
            for (int l = 0; l <= BW_MAX_LEVELS; ++l) pa.unit_off2[l] = wl.unit_off2[l];
            mask_blocks = (long long)B * (wl.units1 + wl.units2);
        }
        if (flat_blocks + pooled_blocks + mask_blocks > 0x7fffffffLL) return PCFA_E_TOOLARGE;
        bw_prep_kernel<<<(unsigned)(flat_blocks + pooled_blocks + mask_blocks), 256, 0, s>>>(f1, f2, pa, (int)flat_blocks, vec, nx, ny,
                                                                                           (int)(flat_blocks + pooled_blocks));
        PCFA_TRY(after_launch());
    }

    const int smem = two_cta ? stages2 * stage2 + smem_fixed + (sparse ? pre_bytes : 0)
                             : BW_STAGES * (2 * BW_A_BYTES + 2 * C * 128) + 1024 + 256;
    static int smem_set = 0, smem2_set = 0;
    if (!two_cta && smem > smem_set) {
        PCFA_CUDA_TRY(cudaFuncSetAttribute(corr_pyramid_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set = smem;
    }
    if (two_cta && smem > smem2_set) {
        PCFA_CUDA_TRY(cudaFuncSetAttribute(corr_pyramid_bwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem2_set = smem;
    }
    const int sms = tc_num_sms();
    const int group = two_cta ? 4 : 2;                  // row blocks per work unit
    static const int env_pdl = [] { const char* e = getenv("PCFA_BWD_PDL"); return e ? atoi(e) : 1; }();
    // PCFA_DETERMINISTIC=1: shares are cut at unit boundaries only (the share count is a divisor of the unit count), so
    // every output element receives exactly ONE reduce-add into its zeroed target and the result is bit-reproducible;
    // the default split-K shares finish partial sums with RED.ADD in arrival order (fp32 addition order varies).
    static const int env_det = [] { const char* e = getenv("PCFA_DETERMINISTIC"); return e ? atoi(e) : 0; }();
    auto det_shares = [&](const BwParams& P, long long cap) -> long long {
        const long long units = P.work_total / P.chunks_total;
        long long n = units < cap ? units : cap;
        while (n > 1 && units % n != 0) --n;
        return n < 1 ? 1 : n;
    };
    auto launch = [&](const BwMaps& maps, const BwParams& P) -> int {
        if (two_cta) {
            long long clusters = sms / 2;
            if (clusters > P.work_total) clusters = P.work_total;
            if (env_det) clusters = det_shares(P, sms / 2);
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(2 * clusters)); cfg.blockDim = dim3(BW_THREADS);
            cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = (P.pass == 2 && env_pdl) ? 1 : 0;      // pass II may overlap pass I's tail
            PCFA_CUDA_TRY(cudaLaunchKernelEx(&cfg, corr_pyramid_bwd_tc2_kernel, maps, P));
        } else {
            const int grid = env_det ? (int)det_shares(P, sms) : (int)(P.work_total < sms ? P.work_total : sms);
            corr_pyramid_bwd_tc_kernel<<<grid, BW_THREADS, smem, s>>>(maps, P);
        }
        return after_launch();
    };

    // ---- pass I: grad_fmap1
    {
        BwMaps maps{};
        BwParams P{};
        P.B = B; P.C = C; P.N = N; P.levels = levels; P.pass = 1; P.terms = terms; P.stages = stages2; P.trace = g_bw_trace;
        int off = 0;
        for (int l = 0; l < levels; ++l) {
            const int nl = L.h[l] * L.w[l];
            P.nl[l] = nl; P.chunk_off[l] = off; off += ceil_div(nl, BW_BK);
            PCFA_TRY(enc3(enc, &maps.a[l], gpyr + L.off[l], nl, N, B, BW_BK, BW_BM));
            PCFA_TRY(enc3(enc, &maps.b[l], wsb + wl.p_split[l], nl, C, 2 * B, BW_BK, two_cta ? C / 2 : C));
        }
        for (int l = levels; l <= BW_MAX_LEVELS; ++l) P.chunk_off[l] = off;
        for (int l = levels; l < BW_MAX_LEVELS; ++l) { maps.a[l] = maps.a[0]; maps.b[l] = maps.b[0]; }
        P.chunks_total = off;
        P.units_per_sample = ceil_div(ceil_div(N, BW_BM), group);
        P.work_total = (long long)B * P.units_per_sample * P.chunks_total;
        P.out[0] = gf1;
        if (sparse) {
            P.sparse = 1; P.det = env_det; P.nunits = B * wl.units1; P.mask_words = wl.words1;
            P.wl_mask = reinterpret_cast<const unsigned*>(wsb + wl.mask1); P.wl_cnt = reinterpret_cast<const int*>(wsb + wl.cnt1);
        }
        if (two_cta) PCFA_TRY(enc3(enc, &maps.o[0], gf1, N, C, B, BW_BM, 16, CU_TENSOR_MAP_SWIZZLE_NONE));
        for (int l = 1; l < BW_MAX_LEVELS; ++l) maps.o[l] = maps.o[0];
        PCFA_TRY(launch(maps, P));
    }
    // ---- pass II: grad of P_l (level 0 goes straight into grad_fmap2)
    {
        BwMaps maps{};
        BwParams P{};
        P.B = B; P.C = C; P.N = N; P.levels = levels; P.pass = 2; P.terms = terms; P.stages = stages2; P.trace = g_bw_trace ? g_bw_trace + 8 * 1024 : nullptr;
        int off = 0;
        for (int l = 0; l < levels; ++l) {
            const int nl = L.h[l] * L.w[l];
            P.nl[l] = nl; P.unit_off[l] = off; off += ceil_div(ceil_div(nl, BW_BM), group);
            PCFA_TRY(enc3(enc, &maps.a[l], gpyr + L.off[l], nl, N, B, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
            P.out[l] = (l == 0) ? gf2 : reinterpret_cast<float*>(wsb + wl.gp[l]);
            if (two_cta) PCFA_TRY(enc3(enc, &maps.o[l], P.out[l], nl, C, B, BW_BM, 16, CU_TENSOR_MAP_SWIZZLE_NONE));
        }
        for (int l = levels; l <= BW_MAX_LEVELS; ++l) P.unit_off[l] = off;
        PCFA_TRY(enc3(enc, &maps.b[0], wsb + wl.f1_split, N, C, 2 * B, BW_BK, two_cta ? C / 2 : C));
        for (int l = levels; l < BW_MAX_LEVELS; ++l) { maps.a[l] = maps.a[0]; maps.o[l] = maps.o[0]; }
        for (int l = 1; l < BW_MAX_LEVELS; ++l) maps.b[l] = maps.b[0];
        P.chunks_total = ceil_div(N, BW_BK);
        P.units_per_sample = off;
        P.work_total = (long long)B * off * P.chunks_total;
        if (sparse) {
            P.sparse = 1; P.det = env_det; P.nunits = B * wl.units2; P.mask_words = wl.words2;
            P.wl_mask = reinterpret_cast<const unsigned*>(wsb + wl.mask2); P.wl_cnt = reinterpret_cast<const int*>(wsb + wl.cnt2);
        }
        PCFA_TRY(launch(maps, P));
    }
    if (levels > 1) {
        BwUnpool u{};
        u.levels = levels;
        for (int l = 0; l < levels; ++l) {
            u.g[l] = (l == 0) ? gf2 : reinterpret_cast<const float*>(wsb + wl.gp[l]);
            u.h[l] = L.h[l]; u.w[l] = L.w[l];
        }
        dim3 ugrid(ceil_div(H, 8), B * C);
        if (ugrid.y > 65535) return PCFA_E_TOOLARGE;
        const int vec = (W % 4 == 0 && (reinterpret_cast<uintptr_t>(gf2) & 15) == 0) ? 1 : 0;
        bw_unpool_kernel<<<ugrid, 256, 0, s>>>(gf2, u, H, W, vec);
        PCFA_TRY(after_launch());
    }
    return PCFA_OK;
}

}  // namespace pcfa
