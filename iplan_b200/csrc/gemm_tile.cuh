// Shared-memory tiled FP32 GEMM building block used by the learner kernels.
//   C[m][n] (+)= sum_k A(m,k) * B(k,n)      m in [m0, m0+BM), n in [n0, n0+BN), k in [k0, k1)
// A and B are read through functors so the same routine serves the three products of
// a Linear layer (y = x W^T, dx = dy W, dW = dy^T x) on strided, agent-batched buffers.
// FP32 FMA accumulation (parity bar: 1e-4 rel vs the fp32 reference).
#pragma once
#include "common.cuh"

namespace iplan {

template <int BM, int BN, int BK, int TM, int TN>
struct GemmTile {
    static constexpr int THREADS = (BM / TM) * (BN / TN);
    static constexpr int LDA = BM + 4;
    static constexpr int LDB = BN + 4;
    static constexpr int SMEM_FLOATS = BK * (LDA + LDB);

    // A_K_FAST: consecutive threads fetch consecutive k of A (A is k-contiguous in memory),
    // otherwise consecutive m.  Same for B with n.
    template <bool A_K_FAST, bool B_K_FAST, class ALoad, class BLoad>
    __device__ __forceinline__ static void run(float* smem, int M, int N, int m0, int n0, int k0, int k1,
                                               ALoad a_at, BLoad b_at, float (&acc)[TM][TN]) {
        float* As = smem;                 // [BK][LDA]
        float* Bs = smem + BK * LDA;      // [BK][LDB]
        const int tid = threadIdx.x;
        const int tx = tid % (BN / TN), ty = tid / (BN / TN);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

        for (int kt = k0; kt < k1; kt += BK) {
#pragma unroll
            for (int idx = tid; idx < BM * BK; idx += THREADS) {
                int m, k;
                if (A_K_FAST) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
                const int gm = m0 + m, gk = kt + k;
                As[k * LDA + m] = (gm < M && gk < k1) ? a_at(gm, gk) : 0.0f;
            }
#pragma unroll
            for (int idx = tid; idx < BN * BK; idx += THREADS) {
                int n, k;
                if (B_K_FAST) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
                const int gn = n0 + n, gk = kt + k;
                Bs[k * LDB + n] = (gn < N && gk < k1) ? b_at(gk, gn) : 0.0f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float av[TM], bv[TN];
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(&As[k * LDA + ty * TM + i]);
                    av[i] = t.x; av[i + 1] = t.y; av[i + 2] = t.z; av[i + 3] = t.w;
                }
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(&Bs[k * LDB + tx * TN + j]);
                    bv[j] = t.x; bv[j + 1] = t.y; bv[j + 2] = t.z; bv[j + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    __device__ __forceinline__ static int row_of(int i) { return (threadIdx.x / (BN / TN)) * TM + i; }
    __device__ __forceinline__ static int col_of(int j) { return (threadIdx.x % (BN / TN)) * TN + j; }
};

}  // namespace iplan
