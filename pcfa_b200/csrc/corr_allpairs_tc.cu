// All-pairs correlation pyramid on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   level_l[b*N + q, y, x] = (1/sqrt C) * <fmap1[b,:,q], pool_l(fmap2)[b,:,y,x]>      l = 0..levels-1
//
// (models/raft/corr.py:13-27,52-60).  avg_pool2d is linear, so pooling the *operand* (fmap2) and
// correlating gives the same pyramid as correlating and pooling the 261 MB volume; every level is
// therefore a plain GEMM tile and ONE persistent launch writes the whole pyramid exactly once —
// no read-back of level 0, no separate pooling passes.
//
// Precision: fp32 features are split into bf16 hi + bf16 mid (a = hi + mid + O(2^-18 a)); the three
// products hi*hi + hi*mid + mid*hi are accumulated in fp32 in TMEM, giving ~1e-5 relative error per
// product — inside the 1e-3 parity bar with two orders of margin, at 1.5x the tensor work of TF32.
//
// Tile mapping: UMMA M = 128 *targets* (a 4 x 32 pixel patch of level l, fetched by one 4-D TMA box from
// the channel-last split copy of pool_l(fmap2), out-of-range rows/cols zero-filled), UMMA N = 128
// *queries*.  TMEM lane == target, so an epilogue warp's 32 lanes are one 32-pixel row segment and
// every global store instruction writes one full 128-byte line of one query's row: coalesced without
// staging through shared memory (round 1 used 8 x 16 patches = two 64-byte half lines per store, which
// alone cost 63.5 instead of 44.6 us, profiles/fwd_store_pattern_r2.txt).  The 128-query operand (hi+mid, all
// channels: 128 KB) stays resident in shared memory while the CTA sweeps target tiles; the target operand streams
// through a TMA/mbarrier ring.  Warp roles: 0 = TMA producer, 1 = MMA issuer (one thread),
// 2 = TMEM allocator, 4..11 = epilogue (TMEM -> registers -> global), double-buffered accumulators.
// The 1/sqrt(C) factor is folded into the query operand before the bf16 split.
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

namespace pcfa {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64;
#ifndef PCFA_TC_PH
#define PCFA_TC_PH 4
#define PCFA_TC_PW 32
#endif
constexpr int TC_PH = PCFA_TC_PH, TC_PW = PCFA_TC_PW;          // target patch (4 x 32; -DPCFA_TC_PH=8 -DPCFA_TC_PW=16: round-1 shape)
constexpr bool TC_WIDE = (TC_PH == 4 && TC_PW == 32);          // a TMEM lane quarter (one warp) = one 32-pixel patch row
constexpr int TC_STAGES = 3;
constexpr int TC2_SLOTS = 5;                // CTA-pair kernel: ring of single [128 x 64 bf16] target tiles (hi, mid, hi, ...)
constexpr int TC2_XBUF_BYTES = 16 * 1024;   // CTA-pair kernel: 1 KB per epilogue warp for the level-1 pooling exchange
constexpr int TC_THREADS = 384;             // 4 control warps + 8 epilogue warps
constexpr int TC2_THREADS = 640;            // CTA-pair kernel: 4 control warps + 16 epilogue warps (lane quarter x column quarter)
constexpr int TC_TILE_BYTES = 128 * 128;      // one [128 rows x 64 bf16] SW128 tile
constexpr int TC_MAX_KCHUNKS = 4;             // C <= 256
constexpr int TC_MAX_LEVELS = 4;

struct TcMaps {
    CUtensorMap q;                  // [C, N, 2B]            box {64, 128, 1}
    CUtensorMap t[TC_MAX_LEVELS];   // [C, W_l, H_l, 2B]     box {64, 16, 8, 1}
};

struct TcParams {
    int B, N, levels, kchunks, qblocks, tiles_per_qb;
    int lh[TC_MAX_LEVELS], lw[TC_MAX_LEVELS], tiles_x[TC_MAX_LEVELS], tile_off[TC_MAX_LEVELS + 1];
    long long lvl_off[TC_MAX_LEVELS];
    long long total_tiles;
    float scale;
    // 2-CTA variant: work item = (sample, pair of query blocks, pair of target tiles)
    int qb2blocks, pairs_per_qb2;
    int fuse_l1;                    // CTA-pair kernel: level 1 is pooled in the epilogue (no level-1 tiles)
    long long total_pairs;
};

// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = 128.
constexpr uint32_t kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                                ((uint32_t)(TC_BM >> 4) << 24);

struct TileCoord { int b, qb, level, ty, tx; long long key; };
__device__ __forceinline__ TileCoord decode_tile(long long t, const TcParams& P) {
    TileCoord c;
    c.key = t / P.tiles_per_qb;
    int nt = (int)(t - c.key * P.tiles_per_qb);
    c.b = (int)(c.key / P.qblocks);
    c.qb = (int)(c.key - (long long)c.b * P.qblocks);
    int l = 0;
#pragma unroll
    for (int i = 1; i < TC_MAX_LEVELS; ++i)
        if (i < P.levels && nt >= P.tile_off[i]) l = i;
    nt -= P.tile_off[l];
    c.level = l;
    c.ty = nt / P.tiles_x[l];
    c.tx = nt - c.ty * P.tiles_x[l];
    return c;
}

// ------------------------------------------------------------------------------------ main kernel
__global__ void __launch_bounds__(TC_THREADS, 1)
corr_pyramid_tc_kernel(const __grid_constant__ TcMaps maps, const TcParams P, float* __restrict__ pyr) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_q = base;                                        // [2][kchunks] tiles
    const uint32_t smem_a = base + 2 * TC_MAX_KCHUNKS * TC_TILE_BYTES;   // [STAGES][2] tiles
    const uint32_t bars = smem_a + TC_STAGES * 2 * TC_TILE_BYTES;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * TC_STAGES;
    const uint32_t bar_qfull = bars + 16 * TC_STAGES, bar_qempty = bar_qfull + 8;
    const uint32_t bar_tfull = bar_qempty + 8, bar_tempty = bar_tfull + 16;
    const uint32_t tmem_slot = bar_tempty + 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t_begin = P.total_tiles * blockIdx.x / gridDim.x;
    const long long t_end = P.total_tiles * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_qfull, 1); mbar_init(bar_qempty, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0 && lane == 0) {
        // ===================================================================== TMA producer
        int stage = 0; uint32_t phase = 0, qphase = 0;
        long long cur = -1;
        for (long long t = t_begin; t < t_end; ++t) {
            const TileCoord c = decode_tile(t, P);
            if (c.key != cur) {
                mbar_wait(bar_qempty, qphase ^ 1);           // MMAs reading the previous query block retired
                mbar_expect_tx(bar_qfull, 2 * P.kchunks * TC_TILE_BYTES);
                for (int part = 0; part < 2; ++part)
                    for (int kc = 0; kc < P.kchunks; ++kc)
                        tma_load_3d(smem_q + (part * TC_MAX_KCHUNKS + kc) * TC_TILE_BYTES, &maps.q, bar_qfull,
                                    kc * TC_BK, c.qb * TC_BN, part * P.B + c.b);
                qphase ^= 1;
                cur = c.key;
            }
            for (int kc = 0; kc < P.kchunks; ++kc) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                mbar_expect_tx(bar_full + 8 * stage, 2 * TC_TILE_BYTES);
                const uint32_t dst = smem_a + stage * 2 * TC_TILE_BYTES;
                tma_load_4d(dst, &maps.t[c.level], bar_full + 8 * stage, kc * TC_BK, c.tx * TC_PW, c.ty * TC_PH, c.b);
                tma_load_4d(dst + TC_TILE_BYTES, &maps.t[c.level], bar_full + 8 * stage, kc * TC_BK, c.tx * TC_PW,
                            c.ty * TC_PH, P.B + c.b);
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===================================================================== MMA issuer
        int stage = 0; uint32_t phase = 0, qphase = 0, acc = 0, accphase = 0;
        long long cur = -1;
        for (long long t = t_begin; t < t_end; ++t) {
            const long long key = t / P.tiles_per_qb;
            if (key != cur) { mbar_wait(bar_qfull, qphase); qphase ^= 1; cur = key; }
            mbar_wait(bar_tempty + 8 * acc, accphase ^ 1);   // epilogue drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * TC_BN;
            for (int kc = 0; kc < P.kchunks; ++kc) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint32_t a_hi = smem_a + stage * 2 * TC_TILE_BYTES, a_mid = a_hi + TC_TILE_BYTES;
                const uint32_t b_hi = smem_q + kc * TC_TILE_BYTES, b_mid = smem_q + (TC_MAX_KCHUNKS + kc) * TC_TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < TC_BK / 16; ++kk) {
                    const uint32_t ko = kk * 32;             // 16 bf16 = 32 bytes along K inside the swizzle row
                    tc_mma_bf16(d_tmem, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_hi + ko), kIdescBf16, (kc | kk) != 0);
                    tc_mma_bf16(d_tmem, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_mid + ko), kIdescBf16, 1);
                    tc_mma_bf16(d_tmem, umma_desc_sw128(a_mid + ko), umma_desc_sw128(b_hi + ko), kIdescBf16, 1);
                }
                tc_commit(bar_empty + 8 * stage);            // frees the smem slot once these MMAs retire
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
            tc_commit(bar_tfull + 8 * acc);
            const bool last_of_block = (t + 1 == t_end) || ((t + 1) / P.tiles_per_qb != key);
            if (last_of_block) tc_commit(bar_qempty);
            acc ^= 1;
            if (acc == 0) accphase ^= 1;
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue (8 warps)
        // warp -> TMEM lanes [32*(warp%4), +32) (hardware rule) and query columns [64*half, +64).
        const int ew = warp & 3, half = (warp - 4) >> 2;
        const int p = ew * 32 + lane;                          // target index inside the 8x16 patch
        const int yl = p / TC_PW, xl = p % TC_PW;
        uint32_t acc = 0, accphase = 0;
        for (long long t = t_begin; t < t_end; ++t) {
            const TileCoord c = decode_tile(t, P);
            const int Hl = P.lh[c.level], Wl = P.lw[c.level];
            const int y = c.ty * TC_PH + yl, x = c.tx * TC_PW + xl;
            const bool ok = (y < Hl) && (x < Wl);
            const long long qstride = (long long)Hl * Wl;
            const int q0 = c.qb * TC_BN + half * 64;
            float* out = pyr + P.lvl_off[c.level] + ((long long)c.b * P.N + q0) * qstride + (long long)y * Wl + x;
            const int nq = min(64, P.N - q0);                  // valid queries in this half (<= 0: none)
            mbar_wait(bar_tfull + 8 * acc, accphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * TC_BN + half * 64;
            uint32_t v0[32], v1[32];
            tc_ld32(taddr, v0);
            tc_ld32(taddr + 32, v1);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);   // accumulator is in registers: release it early
            if (ok) {
                if (nq >= 64) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) { __stcs(out, __uint_as_float(v0[j])); out += qstride; }
#pragma unroll
                    for (int j = 0; j < 32; ++j) { __stcs(out, __uint_as_float(v1[j])); out += qstride; }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) { if (j < nq) __stcs(out, __uint_as_float(v0[j])); out += qstride; }
#pragma unroll
                    for (int j = 0; j < 32; ++j) { if (32 + j < nq) __stcs(out, __uint_as_float(v1[j])); out += qstride; }
                }
            }
            acc ^= 1;
            if (acc == 0) accphase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }
}

// ------------------------------------------------------------------------------------ 2-CTA kernel
// Same algorithm on CTA pairs (cta_group::2): one tcgen05.mma covers M = 256 targets (two 8x16 patches, one per
// CTA) x N = 256 queries (two query blocks, one resident in each CTA's shared memory).  Per MMA cycle each CTA
// now streams HALF the target bytes of the single-CTA kernel, which was limited by the L2->SM operand stream
// (tensor pipe 49 % active).  Both CTAs run a TMA producer (loads signal the LEADER's mbarriers through the
// .cta_group::2 form); only the leader issues MMAs; tcgen05.commit multicasts "slot free"/"accumulator full"
// to both CTAs; both CTAs' epilogue warps release the accumulator on the leader's barrier (remote arrive).
constexpr uint32_t kIdescBf16_2cta = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

struct PairCoord { int b, qb2, level, ty, tx; bool valid; long long key; };
__device__ __forceinline__ PairCoord decode_pair(long long t, int rank, const TcParams& P) {
    PairCoord c;
    c.key = t / P.pairs_per_qb2;
    const int pt = (int)(t - c.key * P.pairs_per_qb2);
    c.b = (int)(c.key / P.qb2blocks);
    c.qb2 = (int)(c.key - (long long)c.b * P.qb2blocks);
    int nt = 2 * pt + rank;
    c.valid = nt < P.tiles_per_qb;
    if (!c.valid) nt = 0;
    int l = 0;
#pragma unroll
    for (int i = 1; i < TC_MAX_LEVELS; ++i)
        if (i < P.levels && nt >= P.tile_off[i]) l = i;
    nt -= P.tile_off[l];
    c.level = l;
    c.ty = nt / P.tiles_x[l];
    c.tx = nt - c.ty * P.tiles_x[l];
    return c;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
corr_pyramid_tc2_kernel(const __grid_constant__ TcMaps maps, const TcParams P, float* __restrict__ pyr) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_q = base;
    const uint32_t smem_a = base + 2 * TC_MAX_KCHUNKS * TC_TILE_BYTES;
    const uint32_t xbuf = smem_a + TC2_SLOTS * TC_TILE_BYTES;
    const uint32_t bars = xbuf + TC2_XBUF_BYTES;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * TC2_SLOTS;
    const uint32_t bar_qfull = bars + 16 * TC2_SLOTS, bar_qempty = bar_qfull + 8;
    const uint32_t bar_tfull = bar_qempty + 8, bar_tempty = bar_tfull + 16;
    const uint32_t tmem_slot = bar_tempty + 16;

    // broadcast from lane 0 so the compiler can prove the role branches warp-uniform (plain SHFL in the epilogue)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const long long ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
    const long long t_begin = P.total_pairs * cid / ncl, t_end = P.total_pairs * (cid + 1) / ncl;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC2_SLOTS; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_qfull, 1); mbar_init(bar_qempty, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0 && lane == 0) {
        // ===================================================================== TMA producer (both CTAs)
        // launched as a programmatic dependent of the operand-prep kernel: everything above overlapped its tail
        asm volatile("griddepcontrol.wait;" ::: "memory");
        int stage = 0; uint32_t phase = 0, qphase = 0;
        long long cur = -1;
        for (long long t = t_begin; t < t_end; ++t) {
            const PairCoord c = decode_pair(t, rank, P);
            if (c.key != cur) {
                mbar_wait(bar_qempty, qphase ^ 1);
                if (rank == 0) mbar_expect_tx(bar_qfull, 2 * 2 * P.kchunks * TC_TILE_BYTES);     // both CTAs' bytes
                for (int part = 0; part < 2; ++part)
                    for (int kc = 0; kc < P.kchunks; ++kc)
                        tma2_load_3d(smem_q + (part * TC_MAX_KCHUNKS + kc) * TC_TILE_BYTES, &maps.q, leader_bar(bar_qfull),
                                     kc * TC_BK, (2 * c.qb2 + rank) * TC_BN, part * P.B + c.b);
                qphase ^= 1;
                cur = c.key;
            }
            // an absent second tile of an odd tile count reads a fully out-of-range box: zero-filled
            const int y0 = c.valid ? c.ty * TC_PH : (P.lh[0] + TC_PH), x0 = c.tx * TC_PW;
            for (int kc = 0; kc < P.kchunks; ++kc) {
#pragma unroll
                for (int part = 0; part < 2; ++part) {          // hi tile, then mid tile: one ring slot each
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * TC_TILE_BYTES);        // both CTAs' bytes
                    tma2_load_4d(smem_a + stage * TC_TILE_BYTES, &maps.t[c.level], leader_bar(bar_full + 8 * stage), kc * TC_BK, x0, y0,
                                 part * P.B + c.b);
                    if (++stage == TC2_SLOTS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ===================================================================== MMA issuer (leader CTA only)
        int stage = 0; uint32_t phase = 0, qphase = 0, acc = 0, accphase = 0;
        long long cur = -1;
        for (long long t = t_begin; t < t_end; ++t) {
            const long long key = t / P.pairs_per_qb2;
            if (key != cur) { mbar_wait(bar_qfull, qphase); qphase ^= 1; cur = key; }
            mbar_wait(bar_tempty + 8 * acc, accphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int kc = 0; kc < P.kchunks; ++kc) {
                const uint32_t b_hi = smem_q + kc * TC_TILE_BYTES, b_mid = smem_q + (TC_MAX_KCHUNKS + kc) * TC_TILE_BYTES;
                {   // targets' hi tile: hi*hi + hi*mid
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t a_hi = smem_a + stage * TC_TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 16; ++kk) {
                        const uint32_t ko = kk * 32;
                        tc2_mma_bf16(d_tmem, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_hi + ko), kIdescBf16_2cta, (kc | kk) != 0);
                        tc2_mma_bf16(d_tmem, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_mid + ko), kIdescBf16_2cta, 1);
                    }
                    tc2_commit_mc(bar_empty + 8 * stage);
                    if (++stage == TC2_SLOTS) { stage = 0; phase ^= 1; }
                }
                {   // targets' mid tile: mid*hi
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t a_mid = smem_a + stage * TC_TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 16; ++kk)
                        tc2_mma_bf16(d_tmem, umma_desc_sw128(a_mid + kk * 32), umma_desc_sw128(b_hi + kk * 32), kIdescBf16_2cta, 1);
                    tc2_commit_mc(bar_empty + 8 * stage);
                    if (++stage == TC2_SLOTS) { stage = 0; phase ^= 1; }
                }
            }
            tc2_commit_mc(bar_tfull + 8 * acc);
            const bool last_of_block = (t + 1 == t_end) || ((t + 1) / P.pairs_per_qb2 != key);
            if (last_of_block) tc2_commit_mc(bar_qempty);
            acc ^= 1;
            if (acc == 0) accphase ^= 1;
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue (16 warps, both CTAs)
        // warp = (TMEM lane quarter ew, query-column quarter cq).  Level 0 is stored straight from the accumulator
        // (lane == target: every store instruction writes one 128-byte line).  Level 1 = 2x2 floor pooling of level 0
        // (F.avg_pool2d, corr.py:25-27) is derived here instead of spending level-1 GEMM tiles (-21 % tensor work).
        // 4 x 32 patches (TC_WIDE): see the exchange below.  8 x 16 patches (round-1 shape, compile-time option):
        // a warp's 32 lanes are two 16-pixel patch rows, so the 2x2 blocks live on lane bits 0 (x) and 4 (y).  Four
        // columns (queries) are reduced together by a transpose-reduce — 3 shuffles per 4 columns — after which the
        // four lanes of a block hold the sums of four different queries: one store writes 4 queries x 8 cells.
        const int ew = warp & 3, cq = (warp - 4) >> 2;         // TMEM lanes [32*ew,+32), query columns [64*cq,+64)
        const int p = ew * 32 + lane;
        const int yl = p / TC_PW, xl = p % TC_PW;
        const bool bx = (lane & 1) != 0, by = (lane & 16) != 0;
        const int sub = (bx ? 2 : 0) + (by ? 1 : 0);           // which of the 4 columns this lane ends up holding
        uint32_t acc = 0, accphase = 0;
        for (long long t = t_begin; t < t_end; ++t) {
            const PairCoord c = decode_pair(t, rank, P);
            const int Hl = P.lh[c.level], Wl = P.lw[c.level];
            const int y = c.ty * TC_PH + yl, x = c.tx * TC_PW + xl;
            const bool ok = c.valid && (y < Hl) && (x < Wl);
            const long long qstride = (long long)Hl * Wl;
            const int q0 = c.qb2 * 256 + cq * 64;
            float* out = pyr + P.lvl_off[c.level] + ((long long)c.b * P.N + q0) * qstride + (long long)y * Wl + x;
            const int nq = P.N - q0;                                // valid queries from q0 on (may be <= 0)
            const bool pool1 = P.fuse_l1 && c.valid && c.level == 0;        // warp-uniform
            const int y1 = c.ty * (TC_PH / 2) + (TC_WIDE ? (ew >> 1) : ew), x1 = c.tx * (TC_PW / 2) + (xl >> 1);
            const bool ok1 = pool1 && y1 < P.lh[1] && x1 < P.lw[1];
            const long long q1stride = (long long)P.lh[1] * P.lw[1];
            float* out1 = pyr + P.lvl_off[1] + ((long long)c.b * P.N + q0 + (TC_WIDE ? 0 : sub)) * q1stride + (long long)y1 * P.lw[1] + x1;
            mbar_wait(bar_tfull + 8 * acc, accphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 256 + cq * 64;
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t v[32];
                tc_ld32(taddr + c0, v);
                tc_wait_ld();
                if (c0 == 32) {                                     // this warp's accumulator slice is in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cta0(bar_tempty + 8 * acc);
                }
                if (ok) {
                    if (nq >= c0 + 32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { __stcs(out, __uint_as_float(v[j])); out += qstride; }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { if (c0 + j < nq) __stcs(out, __uint_as_float(v[j])); out += qstride; }
                    }
                }
                __syncwarp();
                if (TC_WIDE && pool1) {
                    // 4 x 32 patches: this warp holds ONE patch row, its partner (ew ^ 1, same CTA) the other row of the 2x2
                    // blocks.  x pairs are reduced in the warp (the two lanes of a pair end up with the sums of two different
                    // queries), then each warp keeps half of the 16 query pairs and hands the other half to its partner
                    // through 1 KB of shared memory (two 64-thread named barriers per 32 columns).
                    float xr[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float a0 = __uint_as_float(v[2 * k]), a1 = __uint_as_float(v[2 * k + 1]);
                        xr[k] = (bx ? a1 : a0) + __shfl_xor_sync(0xffffffffu, bx ? a0 : a1, 1);     // query c0 + 2k + bx
                    }
                    const bool even = (ew & 1) == 0;
                    const int bar_id = 1 + (warp - 4) / 2;                                          // 8 warp pairs: ids 1..8
                    const uint32_t mine = xbuf + (uint32_t)(warp - 4) * 1024u + lane * 4u;
                    const uint32_t theirs = xbuf + (uint32_t)((warp - 4) ^ 1) * 1024u + lane * 4u;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(mine + i * 128u), "f"(even ? xr[8 + i] : xr[i]) : "memory");
                    named_bar_sync(bar_id, 64);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float o;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o) : "r"(theirs + i * 128u) : "memory");
                        const float r = 0.25f * ((even ? xr[i] : xr[8 + i]) + o);
                        const int qrel = c0 + (even ? 0 : 16) + 2 * i + (bx ? 1 : 0);
                        if (ok1 && qrel < nq) __stcs(out1 + (long long)qrel * q1stride, r);
                    }
                    named_bar_sync(bar_id, 64);                      // the partner has read this chunk: the buffer may be rewritten
                } else if (pool1) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float a0 = __uint_as_float(v[j]), a1 = __uint_as_float(v[j + 1]);
                        const float a2 = __uint_as_float(v[j + 2]), a3 = __uint_as_float(v[j + 3]);
                        // x pairs: lanes with bx keep columns 2,3 and hand over 0,1 (and vice versa)
                        const float r0 = (bx ? a2 : a0) + __shfl_xor_sync(0xffffffffu, bx ? a0 : a2, 1);
                        const float r1 = (bx ? a3 : a1) + __shfl_xor_sync(0xffffffffu, bx ? a1 : a3, 1);
                        // y pairs: lanes with by keep the second column of their pair
                        const float r = (by ? r1 : r0) + __shfl_xor_sync(0xffffffffu, by ? r0 : r1, 16);
                        if (ok1 && c0 + j + sub < nq) __stcs(out1 + (long long)(c0 + j) * q1stride, 0.25f * r);
                    }
                }
            }
            acc ^= 1;
            if (acc == 0) accphase ^= 1;
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------------------------ operand prep
// Fused operand prep, ONE launch for both operands.  Target side: one pass over fmap2 produces the channel-last bf16 hi/mid planes of
// pool_l(fmap2) for every level (successive 2x2 floor pooling, as F.avg_pool2d applied level by level).
// CTA = 32 channels x (8 rows x 32 cols) of level 0: coalesced 128-byte row reads; each thread then owns a
// channel PAIR of one cell and writes bf16x2, so a warp stores two 64-byte channel runs (full sectors).
// Shared tiles use an odd per-channel stride (conflict-free for both the row-wise fill and the channel-wise drain).
// Query side (blockIdx.z >= B*C/32): the same transpose + split of fmap1 * (1/sqrt C), level 0 only.
struct TcPrepArgs {
    __nv_bfloat16* dst[TC_MAX_LEVELS];     // hi plane base per level; mid plane at + plane[l]
    long long plane[TC_MAX_LEVELS];        // B * H_l * W_l * C
    int h[TC_MAX_LEVELS], w[TC_MAX_LEVELS];
    int levels, B, C, skip_l1;
    const float* qsrc;                     // fmap1: blockIdx.z >= B*C/32 handles the query operand (level 0 only)
    __nv_bfloat16* qdst; long long qplane; float qscale;
};
constexpr int PT_S0 = 8 * 33 + 1, PT_S1 = 4 * 17 + 1, PT_S2 = 2 * 9 + 1, PT_S3 = 5;

template <int LVL>
__device__ __forceinline__ void prep_targets_drain(const float* __restrict__ t, int cstride, int rstride,
                                                   __nv_bfloat16* __restrict__ dst, long long plane, int Hl, int Wl,
                                                   int C, int b, int c0, int y0, int x0) {
    constexpr int hh = 8 >> LVL, ww = 32 >> LVL, cells = hh * ww;
    for (int item = threadIdx.x; item < cells * 16; item += 256) {
        const int cp = item & 15, cell = item >> 4;
        const int r = cell / ww, q = cell % ww;
        const int y = (y0 >> LVL) + r, x = (x0 >> LVL) + q;
        if (y >= Hl || x >= Wl) continue;
        const float v0 = t[(2 * cp) * cstride + r * rstride + q], v1 = t[(2 * cp + 1) * cstride + r * rstride + q];
        const __nv_bfloat162 hi = __floats2bfloat162_rn(v0, v1);
        const __nv_bfloat162 mid = __floats2bfloat162_rn(v0 - __low2float(hi), v1 - __high2float(hi));
        const long long o = (((long long)b * Hl + y) * Wl + x) * C + c0 + 2 * cp;
        *reinterpret_cast<__nv_bfloat162*>(dst + o) = hi;
        *reinterpret_cast<__nv_bfloat162*>(dst + plane + o) = mid;
    }
}

__global__ void __launch_bounds__(256)
prep_targets_kernel(const float* __restrict__ src, const TcPrepArgs a) {
    __shared__ float t0[32 * PT_S0];
    __shared__ float t1[32 * PT_S1];
    __shared__ float t2[32 * PT_S2];
    __shared__ float t3[32 * PT_S3];
    const int H = a.h[0], W = a.w[0];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // the pyramid kernel's prologue may start now
    const int cblocks = a.C / 32;
    const bool qside = (int)blockIdx.z >= a.B * cblocks;
    const int zz = qside ? (int)blockIdx.z - a.B * cblocks : (int)blockIdx.z;
    const int b = zz / cblocks, c0 = (zz % cblocks) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
    {
        const int y = y0 + ty, x = x0 + tx;
        const bool inb = y < H && x < W;
        const float* p = (qside ? a.qsrc : src) + (((long long)b * a.C + c0) * H + y) * W + x;
        const long long cs = (long long)H * W;
        const float sc = qside ? a.qscale : 1.f;
        float v[32];                                                  // all 32 loads in flight
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = inb ? __ldg(p + c * cs) * sc : 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) t0[c * PT_S0 + ty * 33 + tx] = v[c];
    }
    __syncthreads();
    if (qside) {
        prep_targets_drain<0>(t0, PT_S0, 33, a.qdst, a.qplane, H, W, a.C, b, c0, y0, x0);
        return;
    }
    for (int e = threadIdx.x; e < 32 * 64; e += 256) {
        const int c = e >> 6, r = (e >> 4) & 3, q = e & 15;
        const float* s0 = t0 + c * PT_S0 + (2 * r) * 33 + 2 * q;
        t1[c * PT_S1 + r * 17 + q] = 0.25f * ((s0[0] + s0[1]) + (s0[33] + s0[34]));
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * 16; e += 256) {
        const int c = e >> 4, r = (e >> 3) & 1, q = e & 7;
        const float* s1 = t1 + c * PT_S1 + (2 * r) * 17 + 2 * q;
        t2[c * PT_S2 + r * 9 + q] = 0.25f * ((s1[0] + s1[1]) + (s1[17] + s1[18]));
    }
    __syncthreads();
    if (threadIdx.x < 32 * 4) {
        const int c = threadIdx.x >> 2, q = threadIdx.x & 3;
        const float* s2 = t2 + c * PT_S2 + 2 * q;
        t3[c * PT_S3 + q] = 0.25f * ((s2[0] + s2[1]) + (s2[9] + s2[10]));
    }
    __syncthreads();
    prep_targets_drain<0>(t0, PT_S0, 33, a.dst[0], a.plane[0], a.h[0], a.w[0], a.C, b, c0, y0, x0);
    if (a.levels > 1 && !a.skip_l1) prep_targets_drain<1>(t1, PT_S1, 17, a.dst[1], a.plane[1], a.h[1], a.w[1], a.C, b, c0, y0, x0);
    if (a.levels > 2) prep_targets_drain<2>(t2, PT_S2, 9, a.dst[2], a.plane[2], a.h[2], a.w[2], a.C, b, c0, y0, x0);
    if (a.levels > 3) prep_targets_drain<3>(t3, PT_S3, 5, a.dst[3], a.plane[3], a.h[3], a.w[3], a.C, b, c0, y0, x0);
}

// ------------------------------------------------------------------------------------ host side
EncodeTiledFn tc_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

int tc_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            cudaGetLastError();
            n = kNumSMs;
        }
    }
    return n;
}

struct TcWorkspace {       // byte offsets into the caller's workspace (all 1 KB aligned)
    int64_t q_split, t_split[TC_MAX_LEVELS], pooled[TC_MAX_LEVELS], total;
};

static TcWorkspace tc_workspace(int B, int C, int H, int W, int levels) {
    TcWorkspace w{};
    auto align = [](int64_t v) { return (v + 1023) & ~(int64_t)1023; };
    int64_t o = 0;
    w.q_split = o; o = align(o + (int64_t)2 * B * H * W * C * 2);
    int h = H, ww = W;
    for (int l = 0; l < levels; ++l) {
        w.t_split[l] = o; o = align(o + (int64_t)2 * B * h * ww * C * 2);
        if (l > 0) { w.pooled[l] = o; o = align(o + (int64_t)B * C * h * ww * 4); }
        h /= 2; ww /= 2;
    }
    w.total = o;
    return w;
}

bool corr_pyramid_tc_supported(int B, int C, int H, int W, int levels) {
    if (C % TC_BK != 0 || C / TC_BK > TC_MAX_KCHUNKS || levels > TC_MAX_LEVELS || levels < 1) return false;
    if (B < 1 || 2 * (long long)B > 0x7fffffff) return false;
    int h = H, w = W;
    for (int l = 0; l < levels; ++l) { if (h < 1 || w < 1) return false; h /= 2; w /= 2; }
    return tc_encode_fn() != nullptr;
}

int64_t corr_pyramid_tc_workspace_bytes(int B, int C, int H, int W, int levels) {
    if (C % TC_BK != 0 || C / TC_BK > TC_MAX_KCHUNKS || levels > TC_MAX_LEVELS || levels < 1) return 0;
    return tc_workspace(B, C, H, W, levels).total;
}

int corr_pyramid_forward_tc(const float* fmap1, const float* fmap2, float* pyramid, void* ws, int64_t ws_bytes,
                            int B, int C, int H, int W, int levels, cudaStream_t s, int two_cta) {
    EncodeTiledFn enc = tc_encode_fn();
    if (!enc) return PCFA_E_NODEVICE;
    const TcWorkspace wl = tc_workspace(B, C, H, W, levels);
    if (!ws || ws_bytes < wl.total || (reinterpret_cast<uintptr_t>(ws) & 15)) return PCFA_E_WORKSPACE;
    uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
    const int N = H * W;
    const PyramidLayout L = make_pyramid_layout(B, H, W, levels);

    // ---- operand preparation: channel-last bf16 hi/mid copies of fmap1 and of pool_l(fmap2)
    static const int env_fuse = [] { const char* e = getenv("PCFA_FWD_FUSE_L1"); return e ? atoi(e) : 1; }();
    const int fuse_l1 = (two_cta && levels >= 2 && env_fuse) ? 1 : 0;
    {
        TcPrepArgs pa{};
        pa.levels = levels; pa.B = B; pa.C = C; pa.skip_l1 = fuse_l1;
        pa.qsrc = fmap1; pa.qdst = reinterpret_cast<__nv_bfloat16*>(wsb + wl.q_split);
        pa.qplane = (long long)B * N * C; pa.qscale = 1.0f / sqrtf((float)C);
        for (int l = 0; l < levels; ++l) {
            pa.dst[l] = reinterpret_cast<__nv_bfloat16*>(wsb + wl.t_split[l]);
            pa.plane[l] = (long long)B * L.h[l] * L.w[l] * C;
            pa.h[l] = L.h[l]; pa.w[l] = L.w[l];
        }
        dim3 grid(ceil_div(W, 32), ceil_div(H, 8), 2 * B * (C / 32));
        if (grid.z > 65535) return PCFA_E_TOOLARGE;
        prep_targets_kernel<<<grid, 256, 0, s>>>(fmap2, pa);
        PCFA_TRY(after_launch());
    }

    // ---- tensor maps
    TcMaps maps;
    {
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)(2 * B)};
        cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)N * C * 2};
        cuuint32_t box[3] = {TC_BK, TC_BN, 1}, es[3] = {1, 1, 1};
        if (enc(&maps.q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, wsb + wl.q_split, dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return PCFA_E_BADARG;
    }
    TcParams P{};
    P.B = B; P.N = N; P.levels = levels; P.kchunks = C / TC_BK; P.qblocks = ceil_div(N, TC_BN);
    P.scale = 1.0f / sqrtf((float)C);
    P.fuse_l1 = fuse_l1;
    int toff = 0;
    for (int l = 0; l < levels; ++l) {
        P.lh[l] = L.h[l]; P.lw[l] = L.w[l]; P.lvl_off[l] = L.off[l];
        P.tiles_x[l] = ceil_div(L.w[l], TC_PW);
        P.tile_off[l] = toff;
        if (!(fuse_l1 && l == 1)) toff += P.tiles_x[l] * ceil_div(L.h[l], TC_PH);
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)L.w[l], (cuuint64_t)L.h[l], (cuuint64_t)(2 * B)};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)L.w[l] * C * 2, (cuuint64_t)L.h[l] * L.w[l] * C * 2};
        cuuint32_t box[4] = {TC_BK, TC_PW, TC_PH, 1}, es[4] = {1, 1, 1, 1};
        if (enc(&maps.t[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, wsb + wl.t_split[l], dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return PCFA_E_BADARG;
    }
    for (int l = levels; l < TC_MAX_LEVELS; ++l) maps.t[l] = maps.t[0];
    P.tile_off[levels] = toff;
    for (int l = levels + 1; l <= TC_MAX_LEVELS; ++l) P.tile_off[l] = toff;
    P.tiles_per_qb = toff;
    P.total_tiles = (long long)B * P.qblocks * toff;

    const int smem = two_cta ? 2 * TC_MAX_KCHUNKS * TC_TILE_BYTES + TC2_SLOTS * TC_TILE_BYTES + TC2_XBUF_BYTES + 1024 + 256
                             : 2 * TC_MAX_KCHUNKS * TC_TILE_BYTES + TC_STAGES * 2 * TC_TILE_BYTES + 1024 + 256;
    if (two_cta) {
        P.qb2blocks = ceil_div(P.qblocks, 2);
        P.pairs_per_qb2 = ceil_div(toff, 2);
        P.total_pairs = (long long)B * P.qb2blocks * P.pairs_per_qb2;
        static bool attr2_set = false;
        if (!attr2_set) {
            PCFA_CUDA_TRY(cudaFuncSetAttribute(corr_pyramid_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr2_set = true;
        }
        long long clusters = tc_num_sms() / 2;
        if (clusters > P.total_pairs) clusters = P.total_pairs;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * clusters)); cfg.blockDim = dim3(TC2_THREADS);
        cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        static const int env_pdl = [] { const char* e = getenv("PCFA_FWD_PDL"); return e ? atoi(e) : 1; }();
        cfg.attrs = at; cfg.numAttrs = env_pdl ? 1 : 0;
        PCFA_CUDA_TRY(cudaLaunchKernelEx(&cfg, corr_pyramid_tc2_kernel, maps, P, pyramid));
        return after_launch();
    }
    static bool attr_set = false;
    if (!attr_set) {
        PCFA_CUDA_TRY(cudaFuncSetAttribute(corr_pyramid_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    int grid = tc_num_sms();
    if ((long long)grid > P.total_tiles) grid = (int)P.total_tiles;
    corr_pyramid_tc_kernel<<<grid, TC_THREADS, smem, s>>>(maps, P, pyramid);
    return after_launch();
}

}  // namespace pcfa
