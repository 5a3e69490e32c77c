// Tensor-core tile product for the 64-wide layers of the learner (replaces the FFMA GemmTile):
//   C[m][n] (+)= sum_k A(m,k) * B(k,n)      m in [m0, m0+BM), n in [n0, n0+BN), k in [k0, k1)
// A and B are fp32, read through functors (strided, agent-batched buffers, any transpose);
// tiles are staged in shared memory as fp32, fragments are split into f16 hi + lo in
// registers and multiplied with mma.sync.m16n8k16 (hi*hi + lo*hi + hi*lo, fp32 accumulate:
// ~2^-22 relative).  FLUSH keeps every tensor-core accumulation chain to one k-tile and adds
// the partials in fp32 RN (needed when K is long; see fc1_mma.cu).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace iplan {

__device__ __forceinline__ void mt_split(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void mt_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int BM, int BN, int WM, int WN>
struct MmaTile {
    static constexpr int THREADS = 32 * WM * WN;
    static constexpr int BK = 16;
    static constexpr int MT = BM / WM / 16;      // m-tiles per warp
    static constexpr int NT = BN / WN / 8;       // n-tiles per warp
    static constexpr int LDA = BM + 4;           // 2t*LDA mod 32 = 8t: conflict-free fragment reads
    static constexpr int LDB = BN + 4;
    static constexpr int SMEM_FLOATS = BK * (LDA + LDB);

    template <bool A_K_FAST, bool B_K_FAST, bool FLUSH, class ALoad, class BLoad>
    __device__ __forceinline__ static void run(float* smem, int M, int N, int m0, int n0, int k0, int k1,
                                               ALoad a_at, BLoad b_at, float (&acc)[MT][NT][4]) {
        float* As = smem;                 // [BK][LDA]  (k-major)
        float* Bs = smem + BK * LDA;      // [BK][LDB]
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int wm = warp / WN, wn = warp % WN;
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0f;

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
            uint32_t ah[MT][4], al[MT][4];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int r = wm * (BM / WM) + i * 16 + g;
                mt_split(As[(2 * t) * LDA + r], As[(2 * t + 1) * LDA + r], ah[i][0], al[i][0]);
                mt_split(As[(2 * t) * LDA + r + 8], As[(2 * t + 1) * LDA + r + 8], ah[i][1], al[i][1]);
                mt_split(As[(2 * t + 8) * LDA + r], As[(2 * t + 9) * LDA + r], ah[i][2], al[i][2]);
                mt_split(As[(2 * t + 8) * LDA + r + 8], As[(2 * t + 9) * LDA + r + 8], ah[i][3], al[i][3]);
            }
            float part[MT][NT][4];
            if (FLUSH) {
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) part[i][j][0] = part[i][j][1] = part[i][j][2] = part[i][j][3] = 0.0f;
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int c = wn * (BN / WN) + j * 8 + g;
                uint32_t bh0, bl0, bh1, bl1;
                mt_split(Bs[(2 * t) * LDB + c], Bs[(2 * t + 1) * LDB + c], bh0, bl0);
                mt_split(Bs[(2 * t + 8) * LDB + c], Bs[(2 * t + 9) * LDB + c], bh1, bl1);
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    float (&d)[4] = FLUSH ? part[i][j] : acc[i][j];
                    mt_mma(d, ah[i], bh0, bh1);
                    mt_mma(d, al[i], bh0, bh1);
                    mt_mma(d, ah[i], bl0, bl1);
                }
            }
            if (FLUSH) {
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        acc[i][j][0] += part[i][j][0]; acc[i][j][1] += part[i][j][1];
                        acc[i][j][2] += part[i][j][2]; acc[i][j][3] += part[i][j][3];
                    }
            }
            __syncthreads();
        }
    }
    // element e of acc[i][j] is C(row_of(i, e), col_of(j, e)) relative to (m0, n0)
    __device__ __forceinline__ static int row_of(int i, int e) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        return (warp / WN) * (BM / WM) + i * 16 + (lane >> 2) + ((e & 2) ? 8 : 0);
    }
    __device__ __forceinline__ static int col_of(int j, int e) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        return (warp % WN) * (BN / WN) + j * 8 + 2 * (lane & 3) + (e & 1);
    }
};

}  // namespace iplan
