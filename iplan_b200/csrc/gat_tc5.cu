// K1 recurrence on the 5th-generation tensor cores: the hard-attention bidirectional GRU of GAT_Net.forward
// (reference nova/GAT_Net.py:57-97) with the hidden-state product on tcgen05.mma, accumulators in TMEM.
//
// One CTA = two environments x one agent-net, BOTH directions.  A direction is one M = 128 tile: TMEM lane / tile row
// r = 64 e + i is the chain of ego slot i of environment e (rows i >= N are padding and never stored).  Per step and tile
//
//     D[128 x 96] = h[128 x 32] . W_hh^T            six tcgen05.mma.kind::f16 (M 128, N 96, K 16):
//                                                   hi*hi (2 k-blocks) + lo*hi (2) + hi*lo (2)  -> fp32-class accuracy
//
// with h kept in shared memory as a K-major SWIZZLE_128B operand tile [128 rows][hi(32) | lo(32)] f16 that the gate
// warps rewrite every step (generic-proxy stores + fence.proxy.async + mbarrier), and W_hh (gate-activation scale
// folded in) as a [96 rows][hi | lo] tile written once.  Eight gate warps per tile (warp -> TMEM lane quarter and one
// half of the hidden units; thread = one chain x 16 units): tcgen05.ld the r|z|n pre-activations AND the chain's ego
// part P (which the prologue's product left in TMEM) 8 hidden units at a time, add the neighbour part Q (shared memory,
// broadcast: a step touches two Q rows; the biases are folded into the Q table), sigmoid / tanh through ex2 + a shared
// rcp, new h -> f16 hi/lo -> the operand tile, the per-step hard-attention logit difference -> dl.
// The two tiles (directions) interleave on the SM: while one tile's gates run on the MUFU / FMA pipes the other
// tile's product runs on the tensor core.  16 warps = 4 per scheduler at <= 128 registers: the step of a warp is a long
// dependent chain (tcgen05.ld -> ex2 -> rcp -> ...), and the MUFU pipe — the unit this loop is bound by — only stays
// busy with several warps per scheduler in different phases (2 warps per scheduler measured 1.00 ms per launch at
// B = 512, half the MUFU bound).  No separate issuer warps: once the eight warps of a tile have passed their named
// barrier (h tile written), lane 0 of the tile's first warp issues the six MMAs and commits them to the tile's
// mbarrier, which all eight warps then wait on.
//
// Prologue, also on the tensor core: enc = ReLU(W_e x + b) per row (thread = row, fp32 FMA, K <= 16), then
// [P | Q] = enc . [W_ih[:, :H] | W_ih[:, H:]]^T as ONE M 128 x N 192 product per direction; P lands in the TMEM lane of
// the thread that owns the chain (-> registers), Q is spilled once to a [row][96] table in shared memory.
//
// Output: dl[dir][s][i] (same scratch layout as gat_recur_kernel), consumed by gat_attend_kernel.
#include "common.cuh"
#include "gat_common.cuh"
#include "tc5.cuh"

namespace iplan {

constexpr int G5_THREADS = 512;             // warp w: tile (direction) (w >> 2) & 1, unit half w >> 3, TMEM lane quarter w & 3
constexpr int G5_TMEM_COLS = 512;
constexpr int G5_A_BYTES = 128 * 128;       // h operand tile: 128 rows x (32 hi + 32 lo) f16
constexpr int G5_BHH_BYTES = G3 * 128;      // W_hh operand tile: 96 rows
constexpr int G5_BIH_BYTES = 2 * G3 * 128;  // [W_ih ego | W_ih neighbour]: 192 rows (prologue only)
constexpr int G5_QP = 100;                  // Q table row pitch in floats: lanes 16 B apart mod 128 B -> conflict-free row writes
constexpr int G5_Q_BYTES = 128 * G5_QP * 4;
constexpr int G5_OFF_A = 0;
constexpr int G5_OFF_BHH = G5_OFF_A + 2 * G5_A_BYTES;
constexpr int G5_OFF_Q = G5_OFF_BHH + 2 * G5_BHH_BYTES;          // the W_ih tiles alias the Q tables (dead before Q is written)
constexpr int G5_OFF_SMALL = G5_OFF_Q + 2 * G5_Q_BYTES;
constexpr int G5_SMALL_FLOATS = H * IN_MAX + H + 2 * G3 + 2 * H + 2 * H + 2 * 2 * 128;   // W_e | b_e | P bias | b_hn | logit weights | logit partials
constexpr int G5_OFF_BAR = G5_OFF_SMALL + G5_SMALL_FLOATS * 4;
constexpr size_t G5_SMEM = G5_OFF_BAR + 64 + 1024;
static_assert(2 * G5_BIH_BYTES <= 2 * G5_Q_BYTES, "W_ih tiles must fit under the Q tables");
static_assert(G5_OFF_BHH % 1024 == 0 && G5_OFF_Q % 1024 == 0 && G5_BIH_BYTES % 1024 == 0 && G5_BHH_BYTES % 1024 == 0, "swizzle atoms are 1024-byte aligned");

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// tcgen05.wait::ld that also names the loaded registers, so that no use of them can be scheduled above the wait
__device__ __forceinline__ void tc5_wait_ld24(float (&a)[8], float (&b)[8], float (&c)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]), "+f"(a[4]), "+f"(a[5]), "+f"(a[6]), "+f"(a[7]),
                   "+f"(b[0]), "+f"(b[1]), "+f"(b[2]), "+f"(b[3]), "+f"(b[4]), "+f"(b[5]), "+f"(b[6]), "+f"(b[7]),
                   "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]), "+f"(c[4]), "+f"(c[5]), "+f"(c[6]), "+f"(c[7])
                 :: "memory");
}
__device__ __forceinline__ void tc5_wait_ld8(float (&a)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]), "+f"(a[4]), "+f"(a[5]), "+f"(a[6]), "+f"(a[7]) :: "memory");
}

__global__ void __launch_bounds__(G5_THREADS, 1) gat_recur_tc5_kernel(GatArgs a) {
    extern __shared__ unsigned char g5_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ag = blockIdx.y, b0 = blockIdx.x * 2;
    const int N = a.n_slots, NM1 = N - 1, in_dim = a.obs_dim + a.latent_dim;
    const float* __restrict__ W = a.params + (int64_t)ag * a.param_stride;
    const GatLayout L = gat_layout(in_dim);

    const uint32_t raw_u = smem_u32(g5_raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;                 // swizzle atoms are 1024-byte aligned
    unsigned char* gb = g5_raw + (base - raw_u);
    float* s_we = reinterpret_cast<float*>(gb + G5_OFF_SMALL);      // [H][IN_MAX] zero padded
    float* s_be = s_we + H * IN_MAX;                                // [H]
    float* s_pb = s_be + H;                                         // [2][96] gate-scaled b_ih (+ b_hh for r|z)
    float* s_bn = s_pb + 2 * G3;                                    // [2][32] K_N b_hn
    float* s_lw = s_bn + 2 * H;                                     // [2][32] logit-difference weights
    float* s_pl = s_lw + 2 * H;                                     // [2 tiles][2 step parities][128 rows] logit partial of unit half 1
    const uint32_t bars = base + G5_OFF_BAR;
    auto d_full = [&](int t) { return bars + 8u * (2 + t); };       // MMA issuer -> the tile's warps: accumulator complete
    const uint32_t pro_bar = bars + 32u, tmem_slot = bars + 40u;

    if (tid == 0) {
        for (int t = 0; t < 2; ++t) mbar_init(d_full(t), 1);
        mbar_init(pro_bar, 1);
        mbar_init_fence();
    }
    if (warp == 0) tc5_alloc<G5_TMEM_COLS>(tmem_slot);

    // ---- inputs of this thread's row (threads 0..127): x = [history | behaviour latent] ----------
    float x[IN_MAX];
#pragma unroll
    for (int k = 0; k < IN_MAX; ++k) x[k] = 0.0f;
    bool row_ok = false;
    if (tid < 128) {
        const int e = tid >> 6, i = tid & 63, b = b0 + e;
        row_ok = i < N && b < a.n_envs;
        if (row_ok) {
            const float* hist = a.hist.ptr + ag * a.hist.stride_agent + b * a.hist.stride_env + i * a.hist.stride_slot;
            const float* beh = a.beh.ptr + ag * a.beh.stride_agent + b * a.beh.stride_env + i * a.beh.stride_slot;
#pragma unroll
            for (int k = 0; k < IN_MAX; ++k) {
                if (k < a.obs_dim) x[k] = hist[k];
                else if (k < in_dim) x[k] = beh[k - a.obs_dim];
            }
        }
    }
    // ---- small per-agent-net constants ------------------------------------------------------------
    for (int idx = tid; idx < H * IN_MAX; idx += G5_THREADS) {
        const int c = idx / IN_MAX, k = idx - c * IN_MAX;
        s_we[idx] = k < in_dim ? W[L.enc_w + c * in_dim + k] : 0.0f;
    }
    if (tid < H) s_be[tid] = W[L.enc_b + tid];
    for (int idx = tid; idx < 2 * G3; idx += G5_THREADS) {
        const int d = idx / G3, c = idx - d * G3;
        const float bih = W[(d ? L.bih_r : L.bih_f) + c], bhh = W[(d ? L.bhh_r : L.bhh_f) + c];
        s_pb[idx] = c < 2 * H ? K_RZ * (bih + bhh) : K_N * bih;
    }
    if (tid < 2 * H) {
        const int d = tid >> 5, u = tid & 31;
        s_bn[tid] = K_N * W[(d ? L.bhh_r : L.bhh_f) + 2 * H + u];
        s_lw[tid] = W[L.he_w + 2 * H + d * H + u] - W[L.he_w + d * H + u];
    }
    // ---- operand tiles of the weights: f16 hi | lo, gate-activation scale folded in -----------------
    // one task = 8 consecutive k of one row: hi chunk `ch`, lo chunk `4 + ch`
    constexpr int HH_TASKS = 2 * G3 * 4, IH_TASKS = 2 * 2 * G3 * 4;
    for (int task = tid; task < HH_TASKS + IH_TASKS; task += G5_THREADS) {
        const float* src;
        uint32_t tile;
        int row, ch, gate;
        if (task < HH_TASKS) {
            const int d = task / (G3 * 4), rr = task - d * (G3 * 4);
            row = rr >> 2; ch = rr & 3; gate = row;
            src = W + (d ? L.whh_r : L.whh_f) + row * H + ch * 8;
            tile = base + G5_OFF_BHH + d * G5_BHH_BYTES;
        } else {
            const int tt = task - HH_TASKS, d = tt / (2 * G3 * 4), rr = tt - d * (2 * G3 * 4);
            row = rr >> 2; ch = rr & 3;
            const int part = row / G3;                      // 0: ego columns (-> P), 1: neighbour columns (-> Q)
            gate = row - part * G3;
            src = W + (d ? L.wih_r : L.wih_f) + gate * 2 * H + part * H + ch * 8;
            tile = base + G5_OFF_Q + d * G5_BIH_BYTES;
        }
        const float ks = gate < 2 * H ? K_RZ : K_N;
        const float4 w0 = *reinterpret_cast<const float4*>(src), w1 = *reinterpret_cast<const float4*>(src + 4);
        uint32_t hi[4], lo[4];
        split_f16(ks * w0.x, ks * w0.y, hi[0], lo[0]);
        split_f16(ks * w0.z, ks * w0.w, hi[1], lo[1]);
        split_f16(ks * w1.x, ks * w1.y, hi[2], lo[2]);
        split_f16(ks * w1.z, ks * w1.w, hi[3], lo[3]);
        sts128(tile + swz128(row, ch), hi[0], hi[1], hi[2], hi[3]);
        sts128(tile + swz128(row, 4 + ch), lo[0], lo[1], lo[2], lo[3]);
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // ---- enc = ReLU(W_e x + b_e) (:50) -> operand tile 0 (shared by both directions' [P | Q] products) ----
    if (tid < 128) {
        const uint32_t tile = base + G5_OFF_A;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float v[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = 8 * ch + 2 * p + q;
                    float acc = s_be[c];
#pragma unroll
                    for (int k = 0; k < IN_MAX; ++k) acc = fmaf(s_we[c * IN_MAX + k], x[k], acc);
                    v[q] = row_ok ? fmaxf(acc, 0.0f) : 0.0f;
                }
                split_f16(v[0], v[1], hi[p], lo[p]);
            }
            sts128(tile + swz128(tid, ch), hi[0], hi[1], hi[2], hi[3]);
            sts128(tile + swz128(tid, 4 + ch), lo[0], lo[1], lo[2], lo[3]);
        }
    }
    fence_proxy_async();
    __syncthreads();

    constexpr uint32_t IDESC_IH = tc5_idesc(128, 2 * G3), IDESC_HH = tc5_idesc(128, G3);
    constexpr int PQ_COL = 128;                               // [P | Q] fwd at TMEM columns 128..319, rev at 320..511
    if (tid == 0) {
        tc5_fence_after();
        const uint64_t da = tc5_smem_desc(base + G5_OFF_A);
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const uint64_t db = tc5_smem_desc(base + G5_OFF_Q + d * G5_BIH_BYTES);
            const uint32_t dst = tmem_base + PQ_COL + d * 2 * G3;
            tc5_mma(dst, da + 0, db + 0, IDESC_IH, 0);        // hi * hi
            tc5_mma(dst, da + 2, db + 2, IDESC_IH, 1);
            tc5_mma(dst, da + 4, db + 0, IDESC_IH, 1);        // lo * hi
            tc5_mma(dst, da + 6, db + 2, IDESC_IH, 1);
            tc5_mma(dst, da + 0, db + 4, IDESC_IH, 1);        // hi * lo
            tc5_mma(dst, da + 2, db + 6, IDESC_IH, 1);
        }
        tc5_commit(pro_bar);
    }

    {
        // ================= thread = chain `row` of tile (direction) t, hidden units 16 hh .. 16 hh + 15 =================
        const int t = (warp >> 2) & 1, hh = warp >> 3, row = (warp & 3) * 32 + lane;
        const int e = row >> 6, i = row & 63, b = b0 + e;
        const bool ok = i < N && b < a.n_envs;
        const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t a_tile = base + G5_OFF_A + t * G5_A_BYTES;
        float* q_tab = reinterpret_cast<float*>(gb + G5_OFF_Q + t * G5_Q_BYTES);
        const uint32_t p_col = tlane + PQ_COL + t * 2 * G3;            // this chain's P: stays in TMEM for the whole recurrence

        mbar_wait(pro_bar, 0);
        tc5_fence_after();
        // neighbour part: this row's Q (+ the gate biases: every step adds exactly one Q row) -> the table (W_ih tiles are dead)
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int col = g * H + 16 * hh + 8 * k;
                float v[8];
                tc5_ld8_nowait(p_col + G3 + col, v);
                tc5_wait_ld8(v);
                const float4 pb0 = *reinterpret_cast<const float4*>(s_pb + t * G3 + col);
                const float4 pb1 = *reinterpret_cast<const float4*>(s_pb + t * G3 + col + 4);
                float4* dst = reinterpret_cast<float4*>(q_tab + row * G5_QP + col);
                dst[0] = make_float4(v[0] + pb0.x, v[1] + pb0.y, v[2] + pb0.z, v[3] + pb0.w);
                dst[1] = make_float4(v[4] + pb1.x, v[5] + pb1.y, v[6] + pb1.z, v[7] + pb1.w);
            }
        const uint32_t row_base = a_tile + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
        const uint32_t rx = (uint32_t)(row & 7);
        // h0 = 0: zero this thread's part of the row (tile 0 held enc; its products are complete)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            sts128(row_base + (((uint32_t)(2 * hh + k) ^ rx) << 4), 0u, 0u, 0u, 0u);
            sts128(row_base + (((uint32_t)(4 + 2 * hh + k) ^ rx) << 4), 0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
        tc5_fence_before();
        __syncthreads();                                        // Q tables complete, TMEM Q columns free
        tc5_fence_after();
        // the tile's product h . W_hh^T: issued by one lane once the tile's eight warps have written h (named barrier)
        const bool issuer = (warp & 3) == 0 && hh == 0 && lane == 0;
        const uint64_t mma_a = tc5_smem_desc(a_tile), mma_b = tc5_smem_desc(base + G5_OFF_BHH + t * G5_BHH_BYTES);
        // accumulators: tile 0 at columns 0..95; tile 1 takes over the forward Q columns (224..319), free by now
        const uint32_t d_off = t ? (uint32_t)(PQ_COL + G3) : 0u;
        const uint32_t mma_d = tmem_base + d_off;
        auto issue = [&]() {
            tc5_fence_after();
            tc5_mma(mma_d, mma_a + 0, mma_b + 0, IDESC_HH, 0);      // hi * hi
            tc5_mma(mma_d, mma_a + 2, mma_b + 2, IDESC_HH, 1);
            tc5_mma(mma_d, mma_a + 4, mma_b + 0, IDESC_HH, 1);      // lo * hi
            tc5_mma(mma_d, mma_a + 6, mma_b + 2, IDESC_HH, 1);
            tc5_mma(mma_d, mma_a + 0, mma_b + 4, IDESC_HH, 1);      // hi * lo
            tc5_mma(mma_d, mma_a + 2, mma_b + 6, IDESC_HH, 1);
            tc5_commit(d_full(t));
        };
        if (issuer) issue();                                    // step 0 (h = 0)

        f32x2 h2[8];                                            // this thread's 16 hidden units, fp32
#pragma unroll
        for (int p = 0; p < 8; ++p) h2[p] = pk2(0.0f, 0.0f);
        const f32x2 one2 = pk2(1.0f, 1.0f), mtwo2 = pk2(-2.0f, -2.0f);
        const uint32_t d_col = tlane + d_off;
        const float* q_env = q_tab + (e * 64) * G5_QP;
        const float* bn = s_bn + t * H;
        const float* lw = s_lw + t * H;
        float* dlp = a.dl + ((((int64_t)ag * a.n_envs + (ok ? b : 0)) * 2 + t) * NM1) * DLP + i;
        float* plb = s_pl + t * 256 + row;

        for (int step = 0; step < NM1; ++step) {
            const int s = t ? NM1 - 1 - step : step;
            const float* q = q_env + (s < i ? s : s + 1) * G5_QP;       // neighbour of ego i at position s (:60-66)
            mbar_wait(d_full(t), step & 1);
            tc5_fence_after();
            f32x2 pl = pk2(0.0f, 0.0f);
#pragma unroll
            for (int k = 0; k < 2; ++k) {                               // hidden units 8c .. 8c+7
                const int c = 2 * hh + k;
                float vr[8], vz[8], vn[8], pr[8], pz[8], pn[8];
                tc5_ld8_nowait(d_col + 8 * c, vr);
                tc5_ld8_nowait(d_col + H + 8 * c, vz);
                tc5_ld8_nowait(d_col + 2 * H + 8 * c, vn);
                tc5_ld8_nowait(p_col + 8 * c, pr);
                tc5_ld8_nowait(p_col + H + 8 * c, pz);
                tc5_ld8_nowait(p_col + 2 * H + 8 * c, pn);
                const float4 qr0 = *reinterpret_cast<const float4*>(q + 8 * c), qr1 = *reinterpret_cast<const float4*>(q + 8 * c + 4);
                const float4 qz0 = *reinterpret_cast<const float4*>(q + H + 8 * c), qz1 = *reinterpret_cast<const float4*>(q + H + 8 * c + 4);
                tc5_wait_ld24(vr, vz, vn);
                tc5_wait_ld24(pr, pz, pn);
                f32x2 r[4], z[4], xx[4];
                xx[0] = add2(add2(pk2(vr[0], vr[1]), pk2(pr[0], pr[1])), pk2(qr0.x, qr0.y));
                xx[1] = add2(add2(pk2(vr[2], vr[3]), pk2(pr[2], pr[3])), pk2(qr0.z, qr0.w));
                xx[2] = add2(add2(pk2(vr[4], vr[5]), pk2(pr[4], pr[5])), pk2(qr1.x, qr1.y));
                xx[3] = add2(add2(pk2(vr[6], vr[7]), pk2(pr[6], pr[7])), pk2(qr1.z, qr1.w));
                sigmoid4_den(xx[0], xx[1], r[0], r[1]);                 // r = 1 / (1 + 2^x')
                sigmoid4_den(xx[2], xx[3], r[2], r[3]);
                xx[0] = add2(add2(pk2(vz[0], vz[1]), pk2(pz[0], pz[1])), pk2(qz0.x, qz0.y));
                xx[1] = add2(add2(pk2(vz[2], vz[3]), pk2(pz[2], pz[3])), pk2(qz0.z, qz0.w));
                xx[2] = add2(add2(pk2(vz[4], vz[5]), pk2(pz[4], pz[5])), pk2(qz1.x, qz1.y));
                xx[3] = add2(add2(pk2(vz[6], vz[7]), pk2(pz[6], pz[7])), pk2(qz1.z, qz1.w));
                sigmoid4_den(xx[0], xx[1], z[0], z[1]);
                sigmoid4_den(xx[2], xx[3], z[2], z[3]);
                const float4 qn0 = *reinterpret_cast<const float4*>(q + 2 * H + 8 * c), qn1 = *reinterpret_cast<const float4*>(q + 2 * H + 8 * c + 4);
                const float4 bn0 = *reinterpret_cast<const float4*>(bn + 8 * c), bn1 = *reinterpret_cast<const float4*>(bn + 8 * c + 4);
                // n pre-activation: (W_in x + b_in) + r (W_hn h + b_hn)   (GRU gate order r, z, n)
                xx[0] = fma2(r[0], add2(pk2(vn[0], vn[1]), pk2(bn0.x, bn0.y)), add2(pk2(pn[0], pn[1]), pk2(qn0.x, qn0.y)));
                xx[1] = fma2(r[1], add2(pk2(vn[2], vn[3]), pk2(bn0.z, bn0.w)), add2(pk2(pn[2], pn[3]), pk2(qn0.z, qn0.w)));
                xx[2] = fma2(r[2], add2(pk2(vn[4], vn[5]), pk2(bn1.x, bn1.y)), add2(pk2(pn[4], pn[5]), pk2(qn1.x, qn1.y)));
                xx[3] = fma2(r[3], add2(pk2(vn[6], vn[7]), pk2(bn1.z, bn1.w)), add2(pk2(pn[6], pn[7]), pk2(qn1.z, qn1.w)));
                f32x2 in[4];
                sigmoid4_den(xx[0], xx[1], in[0], in[1]);
                sigmoid4_den(xx[2], xx[3], in[2], in[3]);
                const float4 lw0 = *reinterpret_cast<const float4*>(lw + 8 * c), lw1 = *reinterpret_cast<const float4*>(lw + 8 * c + 4);
                const f32x2 lwp[4] = {pk2(lw0.x, lw0.y), pk2(lw0.z, lw0.w), pk2(lw1.x, lw1.y), pk2(lw1.z, lw1.w)};
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const f32x2 nn = fma2(mtwo2, in[p], one2);                      // tanh = 1 - 2 / (1 + 2^x')
                    h2[4 * k + p] = fma2(z[p], sub2(h2[4 * k + p], nn), nn);        // (1 - z) n + z h
                    pl = fma2(lwp[p], h2[4 * k + p], pl);
                    split_f16p(h2[4 * k + p], hi[p], lo[p]);
                }
                sts128(row_base + (((uint32_t)c ^ rx) << 4), hi[0], hi[1], hi[2], hi[3]);
                sts128(row_base + (((uint32_t)(4 + c) ^ rx) << 4), lo[0], lo[1], lo[2], lo[3]);
            }
            float pa, pb;
            upk2(pl, pa, pb);
            if (hh) plb[(step & 1) * 128] = pa + pb;                    // the other half's thread adds it after the barrier
            fence_proxy_async();                                        // this thread's h stores -> async proxy
            tc5_fence_before();
            asm volatile("barrier.sync %0, 256;" ::"r"(1 + t) : "memory");   // the tile's 8 warps: h tile complete, D consumed
            if (issuer && step + 1 < NM1) issue();
            if (!hh && ok) dlp[(int64_t)s * DLP] = (pa + pb) + plb[(step & 1) * 128];   // lanes = consecutive egos: coalesced
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc5_fence_after();
        tc5_dealloc<G5_TMEM_COLS>(tmem_base);
    }
}

int launch_gat_recur_tc5(const GatArgs& a, int n_agents, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gat_recur_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G5_SMEM);
        if (e != cudaSuccess) { set_error("gat_step: tcgen05 recurrence smem attr %zu: %s", G5_SMEM, cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    gat_recur_tc5_kernel<<<dim3((a.n_envs + 1) / 2, n_agents), G5_THREADS, G5_SMEM, st>>>(a);
    count_launch();
    return check_launch("gat_step(recur, tcgen05)");
}

}  // namespace iplan
