"""Synthetic store-pattern experiment for the all-pairs build's epilogue: write the 198 MB level-0 volume [7040 q][55][128]
with the epilogue's address pattern and NO tensor work, for several target-patch geometries, to see what the HBM write
path gives for 64..512-byte chunks scattered over 256 query rows per tile.  148 CTAs x 512 threads (16 'epilogue warps'),
each warp: 32 targets (lanes) x 64 queries (a loop of st.global.cs), like corr_pyramid_tc2_kernel."""
import json, statistics, torch
from torch.utils.cpp_extension import load_inline
src = r'''
#include <torch/extension.h>
#include <cuda_runtime.h>
template <int PH, int PW, int VEC, int ST = 0>
__global__ void __launch_bounds__(512) pattern_kernel(float* __restrict__ out, int N, int H, int W, int tiles_x, int tiles_y, int qblocks) {
    // a "tile" = 256 queries x (two patches of PH x PW targets = 256 targets: CTA pair); here one CTA does one patch x 256 queries
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ew = warp & 3, cq = warp >> 2;                 // lane quarter (32 targets), query quarter (64 queries)
    const long long total = (long long)qblocks * tiles_x * tiles_y;
    const long long lo = total * blockIdx.x / gridDim.x, hi = total * (blockIdx.x + 1) / gridDim.x;
    for (long long t = lo; t < hi; ++t) {
        const int qb = (int)(t / (tiles_x * tiles_y)), nt = (int)(t % (tiles_x * tiles_y));
        const int ty = nt / tiles_x, tx = nt % tiles_x;
        const int p = ew * 32 + lane;
        if (VEC == 1) {
            const int y = ty * PH + p / PW, x = tx * PW + p % PW;
            if (y >= H || x >= W) continue;
            float* o = out + ((long long)(qb * 256 + cq * 64)) * H * W + (long long)y * W + x;
            const int nq = N - (qb * 256 + cq * 64);
            for (int j = 0; j < 64; ++j) {
                if (j < nq) {
                    if (ST == 0) __stcs(o, (float)j);
                    else if (ST == 1) *o = (float)j;
                    else if (ST == 2) __stcg(o, (float)j);
                    else __stwt(o, (float)j);
                }
                o += (long long)H * W;
            }
        } else {
            // transposed: lane = (query sub-index, 4 consecutive targets): 128-bit stores, 8 lanes cover 32 targets of one query
            const int qs = lane >> 3, tg = (lane & 7) * 4 + ew * 32;       // 4 queries per instruction
            const int y = ty * PH + tg / PW, x = tx * PW + tg % PW;
            if (y >= H || x >= W) continue;
            float* o = out + ((long long)(qb * 256 + cq * 64 + qs)) * H * W + (long long)y * W + x;
            const int nq = N - (qb * 256 + cq * 64);
            for (int j = 0; j < 64; j += 4) { if (j + qs < nq) __stcs(reinterpret_cast<float4*>(o), make_float4(j, j, j, j)); o += 4LL * H * W; }
        }
    }
}
// reference pattern: the SAME 128-byte line stores, but a warp writes 64 CONSECUTIVE lines of one query plane (what a tile
// order with many targets per query would produce) instead of one line in each of 64 planes
__global__ void __launch_bounds__(512) plane_run_kernel(float* __restrict__ out, int N, int HW) {
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * 16 + (threadIdx.x >> 5), nw = (long long)gridDim.x * 16;
    const int lines = HW / 32, segs = (lines + 63) / 64;
    for (long long u = gw; u < (long long)N * segs; u += nw) {
        const int q = (int)(u / segs), sg = (int)(u % segs);
        float* o = out + (long long)q * HW + (long long)sg * 64 * 32 + lane;
        for (int j = 0; j < 64; ++j) { if (sg * 64 + j < lines) __stcs(o, (float)j); o += 32; }
    }
}
void run(torch::Tensor out, int N, int H, int W, int ph, int pw, int vec, int ctas) {
    const int tx = (W + pw - 1) / pw, ty = (H + ph - 1) / ph, qb = (N + 255) / 256;
    auto s = at::cuda::getCurrentCUDAStream();
    float* o = out.data_ptr<float>();
#define GO(PH, PW, V) pattern_kernel<PH, PW, V><<<ctas, 512, 0, s>>>(o, N, H, W, tx, ty, qb)
    if (vec == 99) { plane_run_kernel<<<ctas, 512, 0, s>>>(o, N, H * W); return; }
    if (ph == 8 && pw == 16 && vec == 1) GO(8, 16, 1);
    else if (ph == 4 && pw == 32 && vec == 1) GO(4, 32, 1);
    else if (ph == 2 && pw == 64 && vec == 1) GO(2, 64, 1);
    else if (ph == 1 && pw == 128 && vec == 1) GO(1, 128, 1);
    else if (ph == 8 && pw == 16 && vec == 4) GO(8, 16, 4);
    else if (ph == 4 && pw == 32 && vec == 4) GO(4, 32, 4);
    else if (ph == 1 && pw == 128 && vec == 4) GO(1, 128, 4);
    else if (ph == 4 && pw == 32 && vec == 11) pattern_kernel<4, 32, 1, 1><<<ctas, 512, 0, s>>>(o, N, H, W, tx, ty, qb);
    else if (ph == 4 && pw == 32 && vec == 12) pattern_kernel<4, 32, 1, 2><<<ctas, 512, 0, s>>>(o, N, H, W, tx, ty, qb);
    else if (ph == 4 && pw == 32 && vec == 13) pattern_kernel<4, 32, 1, 3><<<ctas, 512, 0, s>>>(o, N, H, W, tx, ty, qb);
}
'''
mod = load_inline("pattern", cpp_sources="void run(torch::Tensor out, int N, int H, int W, int ph, int pw, int vec, int ctas);",
                  cuda_sources="#include <ATen/cuda/CUDAContext.h>\n" + src, functions=["run"], extra_cuda_cflags=["-O3", "-gencode", "arch=compute_100a,code=sm_100a"], verbose=False)
N, H, W = 7040, 55, 128
out = torch.empty(N * H * W, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
# vec 11 / 12 / 13: 4-byte stores as plain st.global / st.global.cg / st.global.wt instead of st.global.cs
for (ph, pw, vec) in [(1, 1, 99), (8, 16, 1), (4, 32, 1), (4, 32, 11), (4, 32, 12), (4, 32, 13), (2, 64, 1), (1, 128, 1), (8, 16, 4), (4, 32, 4), (1, 128, 4)]:
    for ctas in (148,):
        ts = []
        for i in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mod.run(out, N, H, W, ph, pw, vec, ctas); e1.record(); torch.cuda.synchronize()
            if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
        us = statistics.median(ts)
        rows.append(dict(patch=f"{ph}x{pw}", store_bytes_per_lane=4 * vec, ctas=ctas, us=round(us, 1), GBps=round(out.numel() * 4 / us / 1e3)))
        print(rows[-1], flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush.zero_(); e0.record(); out.fill_(1.0); e1.record(); torch.cuda.synchronize()
print({"sequential_fill_us": round(e0.elapsed_time(e1) * 1e3, 1), "GBps": round(out.numel() * 4 / (e0.elapsed_time(e1) * 1e3) / 1e3)})
