// Strided fp32 SIMT GEMM used as (1) the exact-fp32 variant of the all-pairs correlation and its
// backward for shapes the tcgen05 path does not cover and (2) the on-device cross-check of the
// tensor-core path.   C(m,n) (+)= alpha * sum_k A(m,k) * B(n,k)
// 128x128x16 tiles, 256 threads, 8x8 register micro-tiles split as 2x2 float4 quads so that shared
// memory reads are conflict-free.
#pragma once
#include "common.cuh"

namespace pcfa {

struct GemmArgs {
    const float* A; const float* B; float* C;
    int M, N, K;
    int64_t a_sm, a_sk, b_sn, b_sk, c_sm, c_sn;   // element strides
    int64_t a_sb, b_sb, c_sb;                     // batch strides (blockIdx.z)
    float alpha;
    int accumulate;                               // C += instead of C =
};

constexpr int GBM = 128, GBN = 128, GBK = 16, GTHREADS = 256;

template <bool A_MCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(GTHREADS) sgemm_simt_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[GBK][GBM + 4];
    __shared__ __align__(16) float Bs[GBK][GBN + 4];

    const float* A = g.A + (int64_t)blockIdx.z * g.a_sb;
    const float* B = g.B + (int64_t)blockIdx.z * g.b_sb;
    float*       C = g.C + (int64_t)blockIdx.z * g.c_sb;
    const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += GBK) {
        // stage A tile: consecutive threads run along the contiguous dimension
#pragma unroll
        for (int e = threadIdx.x; e < GBM * GBK; e += GTHREADS) {
            int m, k;
            if (A_MCONTIG) { m = e % GBM; k = e / GBM; } else { k = e % GBK; m = e / GBK; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < g.M && gk < g.K) ? __ldg(A + gm * g.a_sm + gk * g.a_sk) : 0.f;
        }
#pragma unroll
        for (int e = threadIdx.x; e < GBN * GBK; e += GTHREADS) {
            int n, k;
            if (B_NCONTIG) { n = e % GBN; k = e / GBN; } else { k = e % GBK; n = e / GBK; }
            const int gn = n0 + n, gk = k0 + k;
            Bs[k][n] = (gn < g.N && gk < g.K) ? __ldg(B + gn * g.b_sn + gk * g.b_sk) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= g.M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int gn0 = n0 + h * 64 + tx * 4;
            float* p = C + gm * g.c_sm + gn0 * g.c_sn;
            if (g.c_sn == 1 && gn0 + 3 < g.N && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
                float4 v = make_float4(g.alpha * acc[i][h * 4 + 0], g.alpha * acc[i][h * 4 + 1],
                                       g.alpha * acc[i][h * 4 + 2], g.alpha * acc[i][h * 4 + 3]);
                if (g.accumulate) {
                    const float4 o = *reinterpret_cast<float4*>(p);
                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                }
                *reinterpret_cast<float4*>(p) = v;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (gn0 + j >= g.N) continue;
                    const float v = g.alpha * acc[i][h * 4 + j];
                    float* pj = p + j * g.c_sn;
                    *pj = g.accumulate ? (*pj + v) : v;
                }
            }
        }
    }
}

inline int launch_sgemm(const GemmArgs& g, int batch, cudaStream_t s) {
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return PCFA_OK;
    dim3 grid(ceil_div(g.M, GBM), ceil_div(g.N, GBN), batch);
    const bool am = (g.a_sm == 1), bn = (g.b_sn == 1);
    if (am && bn)        sgemm_simt_kernel<true, true><<<grid, GTHREADS, 0, s>>>(g);
    else if (am && !bn)  sgemm_simt_kernel<true, false><<<grid, GTHREADS, 0, s>>>(g);
    else if (!am && bn)  sgemm_simt_kernel<false, true><<<grid, GTHREADS, 0, s>>>(g);
    else                 sgemm_simt_kernel<false, false><<<grid, GTHREADS, 0, s>>>(g);
    return after_launch();
}

}  // namespace pcfa
