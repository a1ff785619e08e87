// Shared helpers for the pcfa_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pcfa_b200.h"

namespace pcfa {

extern unsigned long long g_launch_count;   // defined in abi.cu (host side, approximate under threads)

inline cudaStream_t as_stream(pcfa_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Call after every kernel launch: counts it and converts a launch error into a status.
inline int after_launch() {
    ++g_launch_count;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return PCFA_OK;
}

#define PCFA_CUDA_TRY(expr)                                                     \
    do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cudaGetLastError(); return (int)_e; } } while (0)

#define PCFA_TRY(expr) do { int _s = (expr); if (_s != PCFA_OK) return _s; } while (0)

template <typename T> __host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;     // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fire-and-forget float add (RED.E.ADD.F32 in SASS; no return value round trip)
__device__ __forceinline__ void red_add(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// Level geometry of the correlation pyramid (models/raft/corr.py:25-27: floor halving).
struct PyramidLayout {
    int     levels;
    int     h[8], w[8];
    int64_t off[9];   // float offsets; off[levels] = total
};

__host__ __device__ inline PyramidLayout make_pyramid_layout(int B, int H, int W, int levels) {
    PyramidLayout L;
    L.levels = levels;
    int64_t rows = (int64_t)B * H * W, o = 0;
    int h = H, w = W;
    for (int l = 0; l < 8; ++l) {
        if (l < levels) {
            L.h[l] = h; L.w[l] = w; L.off[l] = o;
            o += rows * h * w;
            h /= 2; w /= 2;
        } else { L.h[l] = 0; L.w[l] = 0; L.off[l] = o; }
    }
    L.off[8] = o;
    if (levels < 8) L.off[levels] = o;
    return L;
}

// Occupancy bitmap of the gradient pyramid (sparse backward): bit (g, c) of sample b says that some element of
// G[queries 32g .. 32g+31][cells 32k .. 32k+31 of level l] may be non-zero, with column c = chunk_off[l] + k.
// Words: [B][qgroups][words] uint32.  Written by the lookup backward (_occ variants), read by the pyramid backward.
struct OccLayout {
    int qgroups, chunks, words;      // rows (= ceil(N/32)), columns over all levels, 32-bit words per row
    int chunk_off[9];
};

__host__ __device__ inline OccLayout make_occ_layout(int H, int W, int levels) {
    OccLayout o;
    o.qgroups = (H * W + 31) / 32;
    int h = H, w = W, c = 0;
    for (int l = 0; l < 8; ++l) {
        o.chunk_off[l] = c;
        if (l < levels) { c += (h * w + 31) / 32; h /= 2; w /= 2; }
    }
    o.chunk_off[8] = c;
    o.chunks = c;
    o.words = (c + 31) / 32;
    return o;
}

}  // namespace pcfa
