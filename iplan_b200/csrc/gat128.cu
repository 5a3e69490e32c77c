// GAT_Net.forward at GAT_hidden_dim = attention_dim = 128 (BASELINE.json configs[4]: the synthetic GAT + GRU microbench,
// 8192 envs x 32 agent-nets x 16 slots x 128-d; the op is reference nova/GAT_Net.py:41-142 with wider layers).
//
// At H = 128 the hidden-state weights of one direction (384 x 128, as f16 hi + lo: 196 KB) no longer fit in shared
// memory next to anything else, and one chain's ego projection is 384 values.  So the product is turned around:
//
//     D^T[gate row (384 = 3 M-tiles of 128)][chain (N = 16 = the egos of one item)] = W_hh[384 x 128] . h^T[128 x 16]
//
//   A = W_hh lives in TENSOR MEMORY for the lifetime of the CTA (tcgen05.mma with the A operand in TMEM: 3 M-tiles x
//       (64 hi + 64 lo) 32-bit columns = 384 of the 512 columns), written once with tcgen05.st;
//   B = h^T of one item, a K-major SWIZZLE_128B tile [16 chains][128 hi | 128 lo] f16 (8 KB) rewritten by the gate warps every step;
//   D = 3 x 16 fp32 columns per item; two items are in flight per CTA (two groups of four warps, 48 columns each), so
//       one group's gate math overlaps the other group's product.
// A thread is one HIDDEN UNIT u (TMEM lane u of every M-tile holds the r, z and n pre-activations of unit u) for the 16
// chains of its group's item: it adds P (registers, constant over the 15 steps) and Q (shared memory), applies the gates
// (packed fp32 pairs over adjacent chains, ex2 + shared rcp as in K1), keeps h in fp32 registers and writes its f16 hi / lo
// halves into the operand tile.  The per-step hard-attention logit (a sum over the 128 units = over the 128 threads of the
// group) is a warp butterfly + a 4-warp exchange through shared memory.
//
// The GEMM-shaped parts of the op (encode, the factored input projections P | Q, q | k | v, the GRUCell projections) are
// plain row x weight products over all 4.2 M slots: the host side runs them as library GEMMs (cuBLAS fp32 through
// torch.matmul, iplan_b200/nova/gat128.py); this file holds the recurrence, the attention and the GRUCell gate kernels.
#include <stdlib.h>

#include "common.cuh"
#include "gat_common.cuh"
#include "tc5.cuh"

namespace iplan {

constexpr int HB = 128, G3B = 3 * HB, NB = 16, STEPS = NB - 1;
constexpr int G8_THREADS = 256;                      // two groups of four warps; thread = (group, hidden unit)
constexpr int G8_BT_BYTES = 4 * NB * 128;            // h^T operand tile: 4 sub-tiles (hi k<64, hi k>=64, lo, lo) of [16 rows][128 B]
constexpr int G8_Q_BYTES = NB * G3B * 4;             // one item's Q table
constexpr int G8_OFF_BT = 0;                         // 2 groups
constexpr int G8_OFF_Q = 2 * G8_BT_BYTES;            // 2 groups x 2 buffers
constexpr int G8_OFF_PL = G8_OFF_Q + 4 * G8_Q_BYTES; // logit exchange [2 groups][2 parities][4 warps][16]
constexpr int G8_OFF_BAR = G8_OFF_PL + 2 * 2 * 4 * NB * 4;
constexpr size_t G8_SMEM = G8_OFF_BAR + 64 + 1024;
constexpr int G8_D_COL = 384;                        // accumulators: group g at columns 384 + 48 g (r | z | n, 16 chains each)

struct Gat128Args {
    const float* P; const float* Q;                  // [2 dirs][A][items][16][384] gate-scaled ego / neighbour projections (+ biases in Q)
    const float* whh;                                // [A][2][384][128]
    const float* bhn;                                // [A][2][128]  b_hn (unscaled)
    const float* lw;                                 // [A][2][128]  logit-difference weights
    float* dl;                                       // [A][items][2][15][16]
    int n_agents; int64_t n_items; int items_per_cta;
};

__device__ __forceinline__ void tc5_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tc5_wait_ld16x3(float (&a)[16], float (&b)[16], float (&c)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]), "+f"(a[4]), "+f"(a[5]), "+f"(a[6]), "+f"(a[7]),
                   "+f"(a[8]), "+f"(a[9]), "+f"(a[10]), "+f"(a[11]), "+f"(a[12]), "+f"(a[13]), "+f"(a[14]), "+f"(a[15]),
                   "+f"(b[0]), "+f"(b[1]), "+f"(b[2]), "+f"(b[3]), "+f"(b[4]), "+f"(b[5]), "+f"(b[6]), "+f"(b[7]),
                   "+f"(b[8]), "+f"(b[9]), "+f"(b[10]), "+f"(b[11]), "+f"(b[12]), "+f"(b[13]), "+f"(b[14]), "+f"(b[15]),
                   "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]), "+f"(c[4]), "+f"(c[5]), "+f"(c[6]), "+f"(c[7]),
                   "+f"(c[8]), "+f"(c[9]), "+f"(c[10]), "+f"(c[11]), "+f"(c[12]), "+f"(c[13]), "+f"(c[14]), "+f"(c[15])
                 :: "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint16_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(G8_THREADS, 1) gat128_recur_kernel(Gat128Args a) {
    extern __shared__ unsigned char g8_raw[];
    // warp index through a shuffle: warp-uniform for the compiler, so the MMA operands stay in uniform registers (no per-operand
    // register-to-uniform waterfall around each of the 72 tcgen05.mma of an item-step)
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int grp = warp >> 2, u = tid & 127;                       // hidden unit = TMEM lane
    const int ag = blockIdx.y, dir = blockIdx.z;
    const uint32_t raw_u = smem_u32(g8_raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;
    unsigned char* gb = g8_raw + (base - raw_u);
    const uint32_t bars = base + G8_OFF_BAR;
    auto d_full = [&](int g) { return bars + 8u * g; };
    const uint32_t tmem_slot = bars + 32u;
    if (tid == 0) { mbar_init(d_full(0), 1); mbar_init(d_full(1), 1); mbar_init_fence(); }
    if (warp == 0) tc5_alloc<512>(tmem_slot);
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);

    // ---- W_hh -> tensor memory: M-tile m (gate rows 128 m ..) at columns 128 m: 64 columns of f16 hi pairs, 64 of lo pairs ------
    {
        const float* whh = a.whh + ((int64_t)(ag * 2 + dir) * G3B) * HB;
        for (int m = grp; m < 3; m += 2) {                          // group 0: r and n rows, group 1: z rows
            const float ks = m < 2 ? K_RZ : K_N;                    // gate-activation scale folded in (see gat_common.cuh)
            const float* wrow = whh + (int64_t)(m * HB + u) * HB;
#pragma unroll 1
            for (int kc = 0; kc < 4; ++kc) {                        // 32 k at a time
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 w = *reinterpret_cast<const float4*>(wrow + 32 * kc + 4 * q);
                    split_f16(ks * w.x, ks * w.y, hi[2 * q], lo[2 * q]);
                    split_f16(ks * w.z, ks * w.w, hi[2 * q + 1], lo[2 * q + 1]);
                }
                tc5_st16(tlane + m * 128 + 16 * kc, hi);
                tc5_st16(tlane + m * 128 + 64 + 16 * kc, lo);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    const float bn = K_N * a.bhn[(ag * 2 + dir) * HB + u];
    const float lwu = a.lw[(ag * 2 + dir) * HB + u];
    const uint32_t bt = base + G8_OFF_BT + grp * G8_BT_BYTES;       // this group's h^T operand tile
    float* qbuf = reinterpret_cast<float*>(gb + G8_OFF_Q + grp * 2 * G8_Q_BYTES);
    float* plx = reinterpret_cast<float*>(gb + G8_OFF_PL) + grp * (2 * 4 * NB);
    const uint32_t d_col = G8_D_COL + grp * 48;
    const bool issuer_warp = (warp & 3) == 0;                       // warp-uniform; one elected lane issues
    constexpr uint32_t IDESC = tc5_idesc(128, NB);
    tc5_fence_before();
    __syncthreads();                                                // W complete in TMEM before any product
    tc5_fence_after();

    const int64_t it0 = (int64_t)blockIdx.x * a.items_per_cta;
    const int64_t it1 = it0 + a.items_per_cta < a.n_items ? it0 + a.items_per_cta : a.n_items;
    const int64_t pq_stride_dir = (int64_t)a.n_agents * a.n_items * NB * G3B;
    const float* Pd = a.P + dir * pq_stride_dir + (int64_t)ag * a.n_items * NB * G3B;
    const float* Qd = a.Q + dir * pq_stride_dir + (int64_t)ag * a.n_items * NB * G3B;
    auto load_q = [&](int64_t item, int buf) {                      // one item's Q table: 24 KB, 16-byte cp.async by the group's 128 threads
        const float* src = Qd + item * NB * G3B;
        const uint32_t dst = smem_u32(qbuf) + buf * G8_Q_BYTES;
        for (int c = u; c < G8_Q_BYTES / 16; c += 128) cp_async16(dst + 16 * c, src + 4 * c);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t phase = 0;
    int buf = 0;
    if (it0 + grp < it1) load_q(it0 + grp, 0);
    for (int64_t item = it0 + grp; item < it1; item += 2, buf ^= 1) {
        // ---- per item: P into registers (pairs over adjacent chains), h = 0, operand tile = 0 ---------------------------
        f32x2 P2[3][NB / 2];
        {
            const float* pp = Pd + item * NB * G3B + u;
#pragma unroll
            for (int g3 = 0; g3 < 3; ++g3)
#pragma unroll
                for (int c2 = 0; c2 < NB / 2; ++c2)
                    P2[g3][c2] = pk2(pp[(2 * c2) * G3B + g3 * HB], pp[(2 * c2 + 1) * G3B + g3 * HB]);
        }
        f32x2 h2[NB / 2];
#pragma unroll
        for (int c2 = 0; c2 < NB / 2; ++c2) h2[c2] = pk2(0.f, 0.f);
        {
            const uint32_t z = bt + 64 * u;                         // 8 KB / 128 threads
#pragma unroll
            for (int q = 0; q < 4; ++q) asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(z + 16 * q), "r"(0u) : "memory");
        }
        if (item + 2 < it1) load_q(item + 2, buf ^ 1);               // next item's Q while this one runs
        if (item + 2 < it1) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async();
        tc5_fence_before();
        asm volatile("barrier.sync %0, 128;" ::"r"(1 + grp) : "memory");
        const float* qt = qbuf + buf * (NB * G3B);
        auto issue = [&]() {                                        // D^T = W_hh h^T for this group's item: 3 M-tiles x 24 MMAs
            tc5_fence_after();
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const uint32_t dst = tmem_base + d_col + 16 * m;
                const uint32_t wa = tmem_base + m * 128;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {              // hi*hi, lo*hi (h lo x W hi), hi*lo (h hi x W lo)
                    const uint32_t wofs = pass == 2 ? 64u : 0u;     // W lo columns
                    const uint32_t hsub = pass == 1 ? 2u : 0u;      // h lo sub-tiles
#pragma unroll
                    for (int kb = 0; kb < 8; ++kb) {
                        const uint64_t db = tc5_smem_desc(bt + (hsub + (kb >> 2)) * (NB * 128)) + 2 * (kb & 3);
                        tc5_mma_ts(dst, wa + wofs + 8 * kb, db, IDESC, (pass | kb) != 0);
                    }
                }
            }
            tc5_commit(d_full(grp));
        };
        if (issuer_warp) { if (elect_one()) issue(); }              // step 0 (h = 0)
        for (int step = 0; step < STEPS; ++step) {
            const int s = dir ? STEPS - 1 - step : step;
            mbar_wait(d_full(grp), phase & 1);
            ++phase;
            tc5_fence_after();
            float dr[NB], dz[NB], dn[NB];
            tc5_ld16_nowait(tlane + d_col, dr);
            tc5_ld16_nowait(tlane + d_col + 16, dz);
            tc5_ld16_nowait(tlane + d_col + 32, dn);
            // neighbour of ego i at position s: j = s < i ? s : s + 1 -> chains 0..s read Q row s+1, chains s+1.. read row s
            const float* qa = qt + s * G3B + u;                     // row s
            const float* qb = qa + G3B;                             // row s + 1
            const float qr_a = qa[0], qz_a = qa[HB], qn_a = qa[2 * HB];
            const float qr_b = qb[0], qz_b = qb[HB], qn_b = qb[2 * HB];
            tc5_wait_ld16x3(dr, dz, dn);
            const f32x2 one2 = pk2(1.0f, 1.0f), mtwo2 = pk2(-2.0f, -2.0f), bn2 = pk2(bn, bn), lw2 = pk2(lwu, lwu);
            f32x2 pl2[NB / 2];
            uint32_t hh[NB / 2], hl[NB / 2];
#pragma unroll
            for (int c4 = 0; c4 < NB / 4; ++c4) {                   // four chains at a time (one shared rcp per gate)
                const int c0 = 4 * c4;
                auto qsel = [&](float qa_, float qb_, int c) { return s < c ? qa_ : qb_; };
                const f32x2 qr0 = pk2(qsel(qr_a, qr_b, c0), qsel(qr_a, qr_b, c0 + 1)), qr1 = pk2(qsel(qr_a, qr_b, c0 + 2), qsel(qr_a, qr_b, c0 + 3));
                const f32x2 qz0 = pk2(qsel(qz_a, qz_b, c0), qsel(qz_a, qz_b, c0 + 1)), qz1 = pk2(qsel(qz_a, qz_b, c0 + 2), qsel(qz_a, qz_b, c0 + 3));
                const f32x2 qn0 = pk2(qsel(qn_a, qn_b, c0), qsel(qn_a, qn_b, c0 + 1)), qn1 = pk2(qsel(qn_a, qn_b, c0 + 2), qsel(qn_a, qn_b, c0 + 3));
                f32x2 r0, r1, z0, z1, i0, i1;
                sigmoid4_den(add2(add2(pk2(dr[c0], dr[c0 + 1]), P2[0][2 * c4]), qr0),
                             add2(add2(pk2(dr[c0 + 2], dr[c0 + 3]), P2[0][2 * c4 + 1]), qr1), r0, r1);
                sigmoid4_den(add2(add2(pk2(dz[c0], dz[c0 + 1]), P2[1][2 * c4]), qz0),
                             add2(add2(pk2(dz[c0 + 2], dz[c0 + 3]), P2[1][2 * c4 + 1]), qz1), z0, z1);
                sigmoid4_den(fma2(r0, add2(pk2(dn[c0], dn[c0 + 1]), bn2), add2(P2[2][2 * c4], qn0)),
                             fma2(r1, add2(pk2(dn[c0 + 2], dn[c0 + 3]), bn2), add2(P2[2][2 * c4 + 1], qn1)), i0, i1);
                const f32x2 n0 = fma2(mtwo2, i0, one2), n1 = fma2(mtwo2, i1, one2);
                h2[2 * c4] = fma2(z0, sub2(h2[2 * c4], n0), n0);
                h2[2 * c4 + 1] = fma2(z1, sub2(h2[2 * c4 + 1], n1), n1);
                pl2[2 * c4] = mul2(lw2, h2[2 * c4]);
                pl2[2 * c4 + 1] = mul2(lw2, h2[2 * c4 + 1]);
                split_f16p(h2[2 * c4], hh[2 * c4], hl[2 * c4]);     // (chain c0, c0+1) halves of unit u
                split_f16p(h2[2 * c4 + 1], hh[2 * c4 + 1], hl[2 * c4 + 1]);
            }
            // h^T tile: element (chain c, k = u): sub-tile u >> 6 (hi) / 2 + (u >> 6) (lo), row c, 16-byte chunk (u & 63) >> 3 swizzled
            {
                const uint32_t sub = bt + (uint32_t)(u >> 6) * (NB * 128), ch = (uint32_t)(u & 63) >> 3, in = (uint32_t)(u & 7) * 2;
#pragma unroll
                for (int c2 = 0; c2 < NB / 2; ++c2) {
#pragma unroll
                    for (int w2 = 0; w2 < 2; ++w2) {
                        const uint32_t c = 2 * c2 + w2;
                        const uint32_t off = (c >> 3) * 1024u + (c & 7) * 128u + ((ch ^ (c & 7)) << 4) + in;
                        sts16(sub + off, (uint16_t)(w2 ? hh[c2] >> 16 : hh[c2] & 0xffffu));
                        sts16(sub + 2 * (NB * 128) + off, (uint16_t)(w2 ? hl[c2] >> 16 : hl[c2] & 0xffffu));
                    }
                }
            }
            // hard-attention logit difference of every chain: sum over the 128 units = warp butterfly, then the 4 warps through shared memory
            float* plw = plx + (step & 1) * (4 * NB) + (warp & 3) * NB;
#pragma unroll
            for (int c2 = 0; c2 < NB / 2; ++c2) {
                float pa, pb;
                upk2(pl2[c2], pa, pb);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { pa += __shfl_xor_sync(0xffffffffu, pa, o); pb += __shfl_xor_sync(0xffffffffu, pb, o); }
                if (lane == 0) { plw[2 * c2] = pa; plw[2 * c2 + 1] = pb; }
            }
            fence_proxy_async();
            tc5_fence_before();
            asm volatile("barrier.sync %0, 128;" ::"r"(1 + grp) : "memory");
            if (issuer_warp && step + 1 < STEPS) { if (elect_one()) issue(); }
            if (u < NB) {                                           // dl[item][dir][s][i]
                const float* px = plx + (step & 1) * (4 * NB) + u;
                a.dl[((((int64_t)ag * a.n_items + item) * 2 + dir) * STEPS + s) * NB + u] = ((px[0] + px[NB]) + px[2 * NB]) + px[3 * NB];
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 0) { tc5_fence_after(); tc5_dealloc<512>(tmem_base); }
}

// ---- attention over one item (16 slots): scores, gumbel hard gate, soft-max, aggregation ---------------------------------
// qkv [rows][384] (q | k | v, v WITHOUT bias / ReLU), dl [A][items][2][15][16], gumbel NULL or [A][items][16][15][2];
// x out [rows][128].  One warp per ego, 8 warps per CTA = half an item; grid = (2 * items, A).
__global__ void __launch_bounds__(256) gat128_attend_kernel(const float* __restrict__ qkv, const float* __restrict__ v_bias /* [A][128] */,
                                                             const float* __restrict__ dl, const float* __restrict__ he_b /* [A][2] */,
                                                             const float* __restrict__ gumbel, uint64_t seed, uint64_t counter, float inv_tau,
                                                             float* __restrict__ x, int64_t n_items) {
    const int ag = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t item = blockIdx.x >> 1;
    const int i = (blockIdx.x & 1) * 8 + warp;
    const int64_t row0 = ((int64_t)ag * n_items + item) * NB;
    const float* qi = qkv + (row0 + i) * G3B;
    float q4[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) q4[c] = qi[lane + 32 * c];
    const float db = he_b[ag * 2 + 1] - he_b[ag * 2];
    float sc[NB], hd[NB], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        sc[j] = -INFINITY; hd[j] = 0.0f;
        if (j != i) {
            const float* kj = qkv + (row0 + j) * G3B + HB;
            float d = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) d = fmaf(q4[c], kj[lane + 32 * c], d);
            d = warp_sum(d);
            sc[j] = d / 11.313708498984761f;                           // sqrt(attention_dim = 128)
            const int s = j < i ? j : j - 1;
            const int64_t e0 = (((int64_t)ag * n_items + item) * 2) * STEPS * NB;
            const float dlog = (dl[e0 + s * NB + i] + dl[e0 + (STEPS + s) * NB + i]) + db;
            const int64_t edge = (row0 + i) * STEPS + s;
            float noise;
            if (gumbel) noise = gumbel[2 * edge + 1] - gumbel[2 * edge];
            else {
                const uint4 rnd = philox4x32(make_uint4((uint32_t)edge, (uint32_t)(edge >> 32), (uint32_t)counter, (uint32_t)(counter >> 32)),
                                             make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
                const float uu = u01(rnd.x);
                noise = __logf(uu) - __logf(1.0f - uu);
            }
            hd[j] = sigmoidf_acc((dlog + noise) * inv_tau);
            mx = fmaxf(mx, sc[j]);
        }
    }
    float den = 0.0f;
#pragma unroll
    for (int j = 0; j < NB; ++j) { sc[j] = expf(sc[j] - mx); den += sc[j]; }
    float xa[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        if (j != i) {
            const float w = (sc[j] / den) * hd[j];
            const float* vj = qkv + (row0 + j) * G3B + 2 * HB;
#pragma unroll
            for (int c = 0; c < 4; ++c) xa[c] = fmaf(w, fmaxf(vj[lane + 32 * c] + v_bias[ag * HB + lane + 32 * c], 0.0f), xa[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) x[(row0 + i) * HB + lane + 32 * c] = xa[c];
}

// ---- GRUCell gates (nova/GAT_Net.py:140): gi, gh [rows][384] (biases included), h_prev / out [rows][128] ----------------------
__global__ void gat128_gates_kernel(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ hprev,
                                    float* __restrict__ out, int64_t n) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx >> 7;
        const int c = (int)(idx & 127);
        const float* a = gi + r * G3B;
        const float* b = gh + r * G3B;
        const float rg = sigmoidf_acc(a[c] + b[c]);
        const float zg = sigmoidf_acc(a[HB + c] + b[HB + c]);
        const float ng = tanhf_acc(a[2 * HB + c] + rg * b[2 * HB + c]);
        out[idx] = (1.0f - zg) * ng + zg * hprev[idx];
    }
}

}  // namespace iplan

using namespace iplan;

extern "C" int iplan_gat128_recur(const float* P, const float* Q, const float* whh, const float* bhn, const float* lw, float* dl,
                                  int n_agents, int64_t n_items, void* stream) {
    IPLAN_REQUIRE(P && Q && whh && bhn && lw && dl && n_agents > 0 && n_agents <= 65535 && n_items > 0, "gat128_recur: bad arguments");
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gat128_recur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G8_SMEM);
        if (e != cudaSuccess) { set_error("gat128_recur: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    Gat128Args a;
    a.P = P; a.Q = Q; a.whh = whh; a.bhn = bhn; a.lw = lw; a.dl = dl; a.n_agents = n_agents; a.n_items = n_items;
    // (agent-net, direction) pairs x chunks: whole waves of 148 CTAs as nearly as possible, at least 2 items per CTA
    const int pairs = 2 * n_agents;
    int chunks = (148 * 3 + pairs - 1) / pairs;
    if ((int64_t)chunks * 2 > n_items) chunks = (int)((n_items + 1) / 2);
    if (chunks < 1) chunks = 1;
    a.items_per_cta = (int)((n_items + chunks - 1) / chunks);
    chunks = (int)((n_items + a.items_per_cta - 1) / a.items_per_cta);
    gat128_recur_kernel<<<dim3((unsigned)chunks, (unsigned)n_agents, 2), G8_THREADS, G8_SMEM, (cudaStream_t)stream>>>(a);
    count_launch();
    return check_launch("gat128_recur");
}

extern "C" int iplan_gat128_attend(const float* qkv, const float* v_bias, const float* dl, const float* he_b, const float* gumbel,
                                   uint64_t seed, uint64_t counter, float tau, float* x, int n_agents, int64_t n_items, void* stream) {
    IPLAN_REQUIRE(qkv && v_bias && dl && he_b && x && n_agents > 0 && n_items > 0 && tau > 0.f, "gat128_attend: bad arguments");
    IPLAN_REQUIRE(2 * n_items <= 2147483647LL, "gat128_attend: too many items for one launch");
    gat128_attend_kernel<<<dim3((unsigned)(2 * n_items), (unsigned)n_agents), 256, 0, (cudaStream_t)stream>>>(
        qkv, v_bias, dl, he_b, gumbel, seed, counter, 1.0f / tau, x, n_items);
    count_launch();
    return check_launch("gat128_attend");
}

extern "C" int iplan_gat128_gates(const float* gi, const float* gh, const float* hprev, float* out, int64_t rows, void* stream) {
    IPLAN_REQUIRE(gi && gh && hprev && out && rows > 0, "gat128_gates: bad arguments");
    gat128_gates_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(gi, gh, hprev, out, rows * HB);
    count_launch();
    return check_launch("gat128_gates");
}
