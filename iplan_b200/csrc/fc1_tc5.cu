// fc1 forward of the IPPO update on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   Z1[a][r][0..128) = rstd_r * (X_a[r] . W'_a[n] - mean_r * ws[n]) + cc[n]
//
// (feature LayerNorm + fc1 of the actor, n < 64, and the critic, n >= 64, in ONE pass over the packed
// episode rows; reference utils/mappo_utils/mlp.py:50-56, called from learners/ippo_learner.py:267-288.)
// Same operands and the same arithmetic contract as fc1_fwd_mma_kernel (fc1_mma.cu): X and W' as f16
// hi + lo copies, products hi*hi + lo*hi + hi*lo accumulated in fp32, i.e. fp32-class accuracy.
//
// One CTA = one 128-row tile of one agent, 6 warps with fixed roles:
//   warp 0      TMA producer: per 64-wide k-block four 128x64 f16 boxes (Xh, Xl, W'h, W'l; 64 KB) land in a
//               128B-swizzled K-major stage, completion counted on the stage's `full` mbarrier
//   warp 1      TMEM owner + MMA issuer: one lane issues 12 tcgen05.mma (M=128, N=128, K=16) per k-block
//               into one of two 128-column TMEM accumulators, then tcgen05.commit -> `empty` (stage free)
//               and -> `tfull` (accumulator ready)
//   warps 2..5  drain: tcgen05.ld the finished accumulator (thread = one row, 128 columns) and add it to a
//               running fp32 sum in registers.  The tensor core's own fp32 accumulation truncates, so a
//               chain is kept to the 12 MMAs of one k-block and the 39 partial sums are added in fp32 RN
//               on the CUDA cores (same policy as the mma.sync kernels); then the LayerNorm fold + store.
// The drain of k-block i overlaps the MMAs of k-block i+1 (two accumulators) and the TMA of i+2, i+3.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc5.cuh"

namespace iplan {

constexpr int T5_BM = 128, T5_BN = 128, T5_BK = 64, T5_STAGES = 3;
constexpr int T5_THREADS = 192;
constexpr int T5_TILE_BYTES = T5_BM * T5_BK * 2;               // 16 KB: one 128 x 64 f16 operand tile
constexpr int T5_STAGE_BYTES = 4 * T5_TILE_BYTES;              // Xh, Xl, Wh, Wl
constexpr size_t T5_SMEM = (size_t)T5_STAGES * T5_STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int T5_TMEM_COLS = 256;                              // two 128-column fp32 accumulators

constexpr uint32_t T5_IDESC = tc5_idesc(T5_BM, T5_BN);
// as T5_IDESC with A and B MN-major (bits 15, 16)
constexpr uint32_t T5_IDESC_MN = T5_IDESC | (1u << 15) | (1u << 16);

struct Tc5Maps { CUtensorMap xh, xl, wh, wl; };

__global__ void __launch_bounds__(T5_THREADS, 1) fc1_fwd_tc5_kernel(
    const __grid_constant__ Tc5Maps maps, int rows, int ldx,
    const float* __restrict__ ws, const float* __restrict__ cc, const float* __restrict__ stat, float* __restrict__ Z1) {
    extern __shared__ unsigned char t5_raw[];
    const int a = blockIdx.y, m0 = blockIdx.x * T5_BM;
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
    const uint32_t base = (smem_u32(t5_raw) + 1023u) & ~1023u;            // swizzle atoms are 1024-byte aligned
    const uint32_t bars = base + T5_STAGES * T5_STAGE_BYTES;              // full[3] | empty[3] | tfull[2] | tempty[2] | tmem ptr
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (T5_STAGES + s); };
    auto tfull = [&](int b) { return bars + 8u * (2 * T5_STAGES + b); };
    auto tempty = [&](int b) { return bars + 8u * (2 * T5_STAGES + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * T5_STAGES + 4);
    const int nkb = (ldx + T5_BK - 1) / T5_BK;                     // a ragged last k-block is zero-filled by TMA

    if (tid == 0) {
        for (int s = 0; s < T5_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull(b), 1); mbar_init(tempty(b), 4); }    // 4 drain warps
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(T5_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % T5_STAGES, it = kb / T5_STAGES;
                if (it > 0) mbar_wait(empty(s), (it - 1) & 1);
                mbar_expect_tx(full(s), T5_STAGE_BYTES);
                const uint32_t dst = base + s * T5_STAGE_BYTES;
                tma_load_2d(dst + 0 * T5_TILE_BYTES, &maps.xh, kb * T5_BK, a * rows + m0, full(s));
                tma_load_2d(dst + 1 * T5_TILE_BYTES, &maps.xl, kb * T5_BK, a * rows + m0, full(s));
                tma_load_2d(dst + 2 * T5_TILE_BYTES, &maps.wh, kb * T5_BK, a * T5_BN, full(s));
                tma_load_2d(dst + 3 * T5_TILE_BYTES, &maps.wl, kb * T5_BK, a * T5_BN, full(s));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % T5_STAGES, buf = kb & 1, ib = kb >> 1;
                if (ib > 0) mbar_wait(tempty(buf), (ib - 1) & 1);          // the drain warps have emptied this accumulator
                mbar_wait(full(s), (kb / T5_STAGES) & 1);
                tc5_fence_after();
                const uint32_t st = base + s * T5_STAGE_BYTES;
                const uint64_t ah = tc5_smem_desc(st + 0 * T5_TILE_BYTES), al = tc5_smem_desc(st + 1 * T5_TILE_BYTES);
                const uint64_t bh = tc5_smem_desc(st + 2 * T5_TILE_BYTES), bl = tc5_smem_desc(st + 3 * T5_TILE_BYTES);
                const uint32_t d = tmem_base + buf * T5_BN;
#pragma unroll
                for (int k = 0; k < T5_BK / 16; ++k) {                       // 16 f16 = 32 B = 2 descriptor units along K
                    tc5_mma(d, ah + 2 * k, bh + 2 * k, T5_IDESC, k > 0);
                    tc5_mma(d, al + 2 * k, bh + 2 * k, T5_IDESC, 1);
                    tc5_mma(d, ah + 2 * k, bl + 2 * k, T5_IDESC, 1);
                }
                tc5_commit(empty(s));                                      // stage reusable once these MMAs have read it
                tc5_commit(tfull(buf));                                    // accumulator complete
            }
        }
    } else {
        // ===== drain warps: TMEM lane group = warp % 4 =====
        const int lg = warp & 3;
        const int r = m0 + 32 * lg + lane;                                  // this thread's row
        float acc[T5_BN];
#pragma unroll
        for (int j = 0; j < T5_BN; ++j) acc[j] = 0.0f;
        for (int kb = 0; kb < nkb; ++kb) {
            const int buf = kb & 1, ib = kb >> 1;
            mbar_wait(tfull(buf), ib & 1);
            tc5_fence_after();
            const uint32_t t = tmem_base + ((uint32_t)(32 * lg) << 16) + buf * T5_BN;
#pragma unroll
            for (int q = 0; q < T5_BN / 32; ++q) {
                float v[32];
                tc5_ld32(t + 32 * q, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[32 * q + j] += v[j];
            }
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(buf));
        }
        if (r < rows) {                                                     // epilogue: fold the LayerNorm statistics
            const float mean = stat[((int64_t)a * rows + r) * 2], rstd = stat[((int64_t)a * rows + r) * 2 + 1];
            float4* zr = reinterpret_cast<float4*>(Z1 + ((int64_t)a * rows + r) * T5_BN);
            const float4* ws4 = reinterpret_cast<const float4*>(ws + a * T5_BN);
            const float4* cc4 = reinterpret_cast<const float4*>(cc + a * T5_BN);
#pragma unroll
            for (int j = 0; j < T5_BN / 4; ++j) {
                const float4 w = __ldg(ws4 + j), c = __ldg(cc4 + j);
                float4 o;
                o.x = rstd * (acc[4 * j + 0] - mean * w.x) + c.x;
                o.y = rstd * (acc[4 * j + 1] - mean * w.y) + c.y;
                o.z = rstd * (acc[4 * j + 2] - mean * w.z) + c.z;
                o.w = rstd * (acc[4 * j + 3] - mean * w.w) + c.w;
                zr[j] = o;
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc5_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(T5_TMEM_COLS) : "memory");
    }
}

// ---- host: tensor maps ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// [n_rows][ldx] f16 row-major, box = 64 columns x 128 rows, 128B swizzle
static int make_map(CUtensorMap* m, const void* ptr, uint64_t n_rows, uint64_t ldx) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("fc1_forward_tc5: cuTensorMapEncodeTiled not available from the driver"); return -2; }
    const cuuint64_t gdim[2] = {ldx, n_rows};
    const cuuint64_t gstr[1] = {ldx * sizeof(__half)};
    const cuuint32_t box[2] = {(cuuint32_t)T5_BK, (cuuint32_t)T5_BM};
    const cuuint32_t est[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, est,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("fc1_forward_tc5: cuTensorMapEncodeTiled failed (%d)", (int)r); return -3; }
    return 0;
}

int launch_fc1_fwd_tc5(const void* Xh, const void* Xl, int ldx, int64_t rows, int n_agents, const void* Wh, const void* Wl,
                       const float* ws, const float* cc, const float* stat, float* Z1, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fc1_fwd_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T5_SMEM);
        if (e != cudaSuccess) { set_error("fc1_forward_tc5: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    // the operand buffers live for a whole train(): encode the four maps once per (pointers, shape)
    static Tc5Maps maps;
    static const void* key[4] = {nullptr, nullptr, nullptr, nullptr};
    static int64_t key_rows = -1; static int key_ldx = -1, key_agents = -1;
    if (key[0] != Xh || key[1] != Xl || key[2] != Wh || key[3] != Wl || key_rows != rows || key_ldx != ldx || key_agents != n_agents) {
        int rc;
        if ((rc = make_map(&maps.xh, Xh, (uint64_t)n_agents * rows, ldx))) return rc;
        if ((rc = make_map(&maps.xl, Xl, (uint64_t)n_agents * rows, ldx))) return rc;
        if ((rc = make_map(&maps.wh, Wh, (uint64_t)n_agents * T5_BN, ldx))) return rc;
        if ((rc = make_map(&maps.wl, Wl, (uint64_t)n_agents * T5_BN, ldx))) return rc;
        key[0] = Xh; key[1] = Xl; key[2] = Wh; key[3] = Wl; key_rows = rows; key_ldx = ldx; key_agents = n_agents;
    }
    dim3 grid((unsigned)((rows + T5_BM - 1) / T5_BM), n_agents);
    fc1_fwd_tc5_kernel<<<grid, T5_THREADS, T5_SMEM, st>>>(maps, (int)rows, ldx, ws, cc, stat, Z1);
    return 0;
}


// ---------------------------------------------------------------------------------------------------
// backward:  G[a][kk][f] += unscale_a * sum_{r in chunk} dZs[a][r][kk] * X_a[r][f]
// (fc1.weight / feature_norm gradients; the reduction runs over the episode rows, so both operands are
//  MN-major: A = dZs^T (M = kk, contiguous), B = X (N = f, contiguous).)
// One CTA = one (agent, 128-wide f tile, chunk of rows); same warp roles and the same chunked fp32 drain as the
// forward; the chunk's partial product is added to G with vector reductions (red.global.add.v4.f32).
// ---------------------------------------------------------------------------------------------------
constexpr int T5B_BOX_BYTES = 64 * 64 * 2;                     // one 64 (mn) x 64 (rows) f16 box = 8 KB
struct Tc5BwdMaps { CUtensorMap dh, dl, xh, xl; };

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(T5_THREADS, 1) fc1_bwd_tc5_kernel(
    const __grid_constant__ Tc5BwdMaps maps, int rows, int ldx, int kb_per_chunk,
    const float* __restrict__ gscale, float* __restrict__ G /* [A][128][ldx] */) {
    extern __shared__ unsigned char t5_raw[];
    const int a = blockIdx.z, f0 = blockIdx.x * T5_BN;
    const int kb0 = blockIdx.y * kb_per_chunk;
    const int nkb = min(kb_per_chunk, (rows + T5_BK - 1) / T5_BK - kb0);       // 64-row blocks of this chunk
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
    const uint32_t base = (smem_u32(t5_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + T5_STAGES * T5_STAGE_BYTES;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (T5_STAGES + s); };
    auto tfull = [&](int b) { return bars + 8u * (2 * T5_STAGES + b); };
    auto tempty = [&](int b) { return bars + 8u * (2 * T5_STAGES + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * T5_STAGES + 4);

    if (tid == 0) {
        for (int s = 0; s < T5_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull(b), 1); mbar_init(tempty(b), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(T5_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % T5_STAGES, it = kb / T5_STAGES;
                if (it > 0) mbar_wait(empty(s), (it - 1) & 1);
                mbar_expect_tx(full(s), T5_STAGE_BYTES);
                const uint32_t dst = base + s * T5_STAGE_BYTES;
                const int r = (kb0 + kb) * T5_BK;                          // rows past the agent's last are zero-filled
#pragma unroll
                for (int h = 0; h < 2; ++h) {                              // two 64-column boxes per 128-wide operand
                    tma_load_3d(dst + 0 * T5_TILE_BYTES + h * T5B_BOX_BYTES, &maps.dh, 64 * h, r, a, full(s));
                    tma_load_3d(dst + 1 * T5_TILE_BYTES + h * T5B_BOX_BYTES, &maps.dl, 64 * h, r, a, full(s));
                    tma_load_3d(dst + 2 * T5_TILE_BYTES + h * T5B_BOX_BYTES, &maps.xh, f0 + 64 * h, r, a, full(s));
                    tma_load_3d(dst + 3 * T5_TILE_BYTES + h * T5B_BOX_BYTES, &maps.xl, f0 + 64 * h, r, a, full(s));
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % T5_STAGES, buf = kb & 1, ib = kb >> 1;
                if (ib > 0) mbar_wait(tempty(buf), (ib - 1) & 1);
                mbar_wait(full(s), (kb / T5_STAGES) & 1);
                tc5_fence_after();
                const uint32_t st = base + s * T5_STAGE_BYTES;
                const uint64_t ah = tc5_smem_desc_mn(st + 0 * T5_TILE_BYTES, T5B_BOX_BYTES), al = tc5_smem_desc_mn(st + 1 * T5_TILE_BYTES, T5B_BOX_BYTES);
                const uint64_t bh = tc5_smem_desc_mn(st + 2 * T5_TILE_BYTES, T5B_BOX_BYTES), bl = tc5_smem_desc_mn(st + 3 * T5_TILE_BYTES, T5B_BOX_BYTES);
                const uint32_t d = tmem_base + buf * T5_BN;
#pragma unroll
                for (int k = 0; k < T5_BK / 16; ++k) {                       // 16 rows = two 8-row groups = 2048 B = 128 units
                    tc5_mma(d, ah + 128 * k, bh + 128 * k, T5_IDESC_MN, k > 0);
                    tc5_mma(d, al + 128 * k, bh + 128 * k, T5_IDESC_MN, 1);
                    tc5_mma(d, ah + 128 * k, bl + 128 * k, T5_IDESC_MN, 1);
                }
                tc5_commit(empty(s));
                tc5_commit(tfull(buf));
            }
        }
    } else {
        const int lg = warp & 3;
        const int kk = 32 * lg + lane;                                      // this thread's output row
        float acc[T5_BN];
#pragma unroll
        for (int j = 0; j < T5_BN; ++j) acc[j] = 0.0f;
        for (int kb = 0; kb < nkb; ++kb) {
            const int buf = kb & 1, ib = kb >> 1;
            mbar_wait(tfull(buf), ib & 1);
            tc5_fence_after();
            const uint32_t t = tmem_base + ((uint32_t)(32 * lg) << 16) + buf * T5_BN;
#pragma unroll
            for (int q = 0; q < T5_BN / 32; ++q) {
                float v[32];
                tc5_ld32(t + 32 * q, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[32 * q + j] += v[j];
            }
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(buf));
        }
        const float unscale = gscale[a];
        float* gr = G + ((int64_t)a * T5_BM + kk) * ldx + f0;
#pragma unroll
        for (int j = 0; j < T5_BN; j += 4)
            if (f0 + j < ldx)                                                // ldx % 4 == 0: a group of 4 is all in or all out
                red_add_v4(gr + j, acc[j] * unscale, acc[j + 1] * unscale, acc[j + 2] * unscale, acc[j + 3] * unscale);
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc5_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(T5_TMEM_COLS) : "memory");
    }
}

// [n_agents][n_rows][ld] f16, box = 64 columns x 64 rows x 1 agent, 128B swizzle; rows / columns past the end read as zero
static int make_map3(CUtensorMap* m, const void* ptr, uint64_t n_agents, uint64_t n_rows, uint64_t ld) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("fc1_backward_tc5: cuTensorMapEncodeTiled not available from the driver"); return -2; }
    const cuuint64_t gdim[3] = {ld, n_rows, n_agents};
    const cuuint64_t gstr[2] = {ld * sizeof(__half), n_rows * ld * sizeof(__half)};
    const cuuint32_t box[3] = {64, 64, 1};
    const cuuint32_t est[3] = {1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, est,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("fc1_backward_tc5: cuTensorMapEncodeTiled failed (%d)", (int)r); return -3; }
    return 0;
}

int launch_fc1_bwd_tc5(const void* Xh, const void* Xl, int ldx, int64_t rows, int n_agents, const void* Dh, const void* Dl,
                       const float* unscale, float* G, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fc1_bwd_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T5_SMEM);
        if (e != cudaSuccess) { set_error("fc1_backward_tc5: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    static Tc5BwdMaps maps;
    static const void* key[4] = {nullptr, nullptr, nullptr, nullptr};
    static int64_t key_rows = -1; static int key_ldx = -1, key_agents = -1;
    if (key[0] != Xh || key[1] != Xl || key[2] != Dh || key[3] != Dl || key_rows != rows || key_ldx != ldx || key_agents != n_agents) {
        int rc;
        if ((rc = make_map3(&maps.dh, Dh, n_agents, rows, T5_BM))) return rc;
        if ((rc = make_map3(&maps.dl, Dl, n_agents, rows, T5_BM))) return rc;
        if ((rc = make_map3(&maps.xh, Xh, n_agents, rows, ldx))) return rc;
        if ((rc = make_map3(&maps.xl, Xl, n_agents, rows, ldx))) return rc;
        key[0] = Xh; key[1] = Xl; key[2] = Dh; key[3] = Dl; key_rows = rows; key_ldx = ldx; key_agents = n_agents;
    }
    // split the rows so that (f tiles x chunks x agents) fills the 148 SMs in whole waves as nearly as possible
    const int ftiles = (ldx + T5_BN - 1) / T5_BN;
    const int total_kb = (int)((rows + T5_BK - 1) / T5_BK);
    int best_chunks = 1; double best_eff = 0.0;
    for (int c = 1; c <= 16 && c <= total_kb; ++c) {
        const int ctas = ftiles * c * n_agents;
        const double eff = (double)ctas / (148.0 * ((ctas + 147) / 148));
        if (eff > best_eff + 1e-9) { best_eff = eff; best_chunks = c; }
    }
    const int kb_per_chunk = (total_kb + best_chunks - 1) / best_chunks;
    const int chunks = (total_kb + kb_per_chunk - 1) / kb_per_chunk;
    dim3 grid((unsigned)ftiles, (unsigned)chunks, (unsigned)n_agents);
    fc1_bwd_tc5_kernel<<<grid, T5_THREADS, T5_SMEM, st>>>(maps, (int)rows, ldx, kb_per_chunk, unscale, G);
    return 0;
}

}  // namespace iplan
