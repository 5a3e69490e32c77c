// K1 on the 5th-generation tensor cores: the whole GAT step of GAT_Net.forward (reference nova/GAT_Net.py:41-142) in ONE launch;
// its core is the hard-attention bidirectional GRU (:57-97) with the hidden-state product on tcgen05.mma, accumulators AND the
// hidden state itself in tensor memory.
//
// One CTA = two environments x one agent-net, BOTH directions.  A direction is one M = 128 tile: TMEM lane / tile row
// r = 64 e + i is the chain of ego slot i of environment e (rows i >= N are padding and never stored).  Per step and tile
//
//     D[128 x 96] = h[128 x 32] . W_hh^T            six tcgen05.mma.kind::f16 (M 128, N 96, K 16):
//                                                   hi*hi (2 k-blocks) + lo*hi (2) + hi*lo (2)  -> fp32-class accuracy
//
// The A operand h is read FROM TENSOR MEMORY (32 columns per tile: f16 pairs of the hi parts, then of the lo parts), written
// every step by the gate warps with tcgen05.st (thread = TMEM lane = chain); only W_hh (a [96 rows][hi | lo] SWIZZLE_128B tile,
// gate-activation scale folded in, written once) comes from shared memory.  An h tile in shared memory made each MMA fetch
// 4 KB + 3 KB of operands and the six of them ran at the shared-memory bandwidth: ~1000 cycles from "h written" to
// "pre-activations ready"; with A in tensor memory ~530 (tools/k1_trace.py, profiles/r2_k1_trace_*.txt).
// Eight gate warps per tile (warp -> TMEM lane quarter and one half of the hidden units; thread = one chain x 16 units):
// tcgen05.ld the r|z|n pre-activations AND the chain's ego part P (which the prologue's product left in TMEM) 8 hidden units
// at a time, add the neighbour part Q (shared memory, broadcast: a step touches two Q rows; the biases are folded into the Q
// table), sigmoid / tanh through ex2 + a shared rcp, new h -> f16 hi/lo -> tensor memory, the per-step hard-attention logit
// difference -> dl.  16 warps = 4 per scheduler at <= 128 registers.  Once the eight warps of a tile have passed their named
// barrier one ELECTED lane of the tile's first warp issues the six MMAs and commits them to the tile's mbarrier, which all eight
// warps then wait on.  The warp index comes from a shuffle so that the compiler knows it is warp-uniform: descriptors and TMEM
// addresses then stay in uniform registers and the issue is ~25 instructions (with a per-thread predicate the compiler wrapped
// every MMA operand in a register-to-uniform "waterfall" loop: ~800 cycles per step).
//
// What bounds the loop (measured, B = 512): the XU pipe (ex2 / rcp: 60 MUFU per warp-step = 1920 cycles per step of both tiles)
// under a step period of ~2950 cycles: gate phase ~2600 (both tiles drift into lock-step and share the XU pipe) + ~530 of MMA
// round trip.  Tried and measured, not kept: all P / Q / bias adds as extra MMAs with 13 of 19 issued a step ahead (0.82 ms but
// shared-memory-bandwidth bound: 19 x 7 KB of operands per tile-step); a dedicated issuer warp (17 warps cap the kernel at 96
// registers); a coupling that holds the tiles half a phase apart (the MMA then contends with the other tile's tcgen05.ld/st);
// 2^x of the z gate as an FMA-pipe polynomial (correct to 2e-6, 7 % slower: the extra instructions cost more than the MUFU slots).
//
// Prologue, also on the tensor core: enc = ReLU(W_e x + b) per row (4 threads per row, fp32 FMA, K <= 16), then
// [P | Q] = enc . [W_ih[:, :H] | W_ih[:, H:]]^T as ONE M 128 x N 192 product per direction; P stays in the TMEM lane of
// the thread that owns the chain, Q is spilled once to a [row][96] table in shared memory.
//
// FUSED: attention + GRUCell (:99-142) in the same kernel, q|k|v, h_prev W_hh^T and x W_ih^T as three more tcgen05 products over the
// dead operand tiles; the per-edge logits stay in shared memory.  FUSED = false: dl scratch (layout of gat_recur_kernel) for
// gat_attend_kernel (n_slots too large for the fused shared-memory map, and the cross-check path impl 2).
#include "common.cuh"
#include "gat_common.cuh"
#include "tc5.cuh"

#include <stdlib.h>

namespace iplan {

constexpr int G5_THREADS = 512;             // warp w: tile (direction) (w >> 2) & 1, unit half w >> 3, TMEM lane quarter w & 3
constexpr int G5_TMEM_COLS = 512;
constexpr int G5_A_BYTES = 128 * 128;       // h operand tile: 128 rows x (32 hi + 32 lo) f16
constexpr int G5_BHH_BYTES = G3 * 128;      // W_hh operand tile: 96 rows
constexpr int G5_BIH_BYTES = 2 * G3 * 128;  // [W_ih ego | W_ih neighbour]: 192 rows (prologue only)
constexpr int G5_QP = 100;                  // Q table row pitch in floats: lanes 16 B apart mod 128 B -> conflict-free row writes
constexpr int G5_KP = 36;                   // k / v table row pitch (attention phase)
// small constants: W_e | b_e | P bias | b_hn | logit weights | logit partials | (fused) v bias | GRUCell biases | soft-max exchange
constexpr int G5_SMALL_FLOATS = H * IN_MAX + H + 2 * G3 + 2 * H + 2 * H + 2 * 2 * 128 + H + 2 * G3 + 2 * 4 * 128;
constexpr int G5_ATT_BYTES = 2 * 128 * G5_KP * 4 + 3 * H * 128 * 4;     // k | v tables + three partial aggregates [3][32][128]

// Shared-memory map (byte offsets from the 1024-aligned base).  FUSED adds a copy of the enc operand tile (the h tile
// overwrites the first one) and the per-edge logit tables dl[dir][s][row]; after the recurrence the attention phase
// reuses the Q tables (k, v, partial aggregates), the W_hh tiles (W_q|k|v, GRUCell W_hh) and the h tiles (h_prev / x, GRUCell W_ih).
struct G5Layout { uint32_t a, enc, bhh, q, q_dir, dl, small, bar, total; };
__host__ __device__ inline G5Layout g5_layout(int N, bool fused) {
    G5Layout l;
    l.a = 0;
    l.enc = l.a + 2 * G5_A_BYTES;
    l.bhh = l.enc + (fused ? G5_A_BYTES : 0);
    l.q = l.bhh + 2 * G5_BHH_BYTES;                                  // the W_ih tiles alias the Q tables (dead before Q is written)
    uint32_t qd = (uint32_t)(2 * N * G5_QP * 4);
    if (qd < (uint32_t)G5_BIH_BYTES) qd = G5_BIH_BYTES;
    if (fused && 2 * qd < (uint32_t)G5_ATT_BYTES) qd = (G5_ATT_BYTES + 1) / 2;
    l.q_dir = (qd + 1023u) & ~1023u;
    l.dl = l.q + 2 * l.q_dir;
    l.small = l.dl + (fused ? (uint32_t)(2 * (N - 1) * 128 * 4) : 0u);
    l.bar = l.small + G5_SMALL_FLOATS * 4;
    l.total = l.bar + 64 + 1024;
    return l;
}
constexpr size_t G5_SMEM_MAX = 227 * 1024;

// DBG (timing experiments only, results are garbage): bit 0 = tile 1 (reverse direction) does nothing; bit 1 = the MUFU
// instructions of the gates are replaced by FMA-pipe stand-ins
// bit 2 = one CTA stamps clock64() at its phase boundaries into g5_clk (read back with iplan_gat_debug_clocks)
__device__ long long g5_clk[32];
// bit 6 (64): per-step event clocks of warps 0 / 4 / 8 / 12 (lane 0) of one CTA for steps 20..27: [warp slot 4][step 8][event 8]
__device__ long long g5_trace[4 * 8 * 8];
#define G5_TRACE(ev)                                                                                                   \
    do {                                                                                                               \
        if constexpr ((DBG & 64) != 0) {                                                                               \
            if (blockIdx.x == 3 && blockIdx.y == 0 && lane == 0 && (warp & 3) == 0 && step >= 20 && step < 28)         \
                g5_trace[((warp >> 2) * 8 + (step - 20)) * 8 + (ev)] = clock64();                                      \
        }                                                                                                              \
    } while (0)
#define G5_STAMP(k)                                                                             \
    do {                                                                                        \
        if constexpr ((DBG & 4) != 0) {                                                         \
            if (blockIdx.x == 3 && blockIdx.y == 0 && tid == 0) g5_clk[k] = clock64();         \
        }                                                                                       \
    } while (0)
template <int DBG, bool FUSED>
__global__ void __launch_bounds__(G5_THREADS, 1) gat_tc5_kernel(GatArgs a) {
    extern __shared__ unsigned char g5_raw[];
    // warp index through a shuffle: the compiler then knows it is warp-uniform, and everything derived from it (tile,
    // descriptors, TMEM columns) stays in uniform registers — the MMA issue below is a handful of instructions instead of
    // a register-to-uniform "waterfall" per operand (which cost ~800 cycles per step: tools/k1_trace.py)
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ag = blockIdx.y, b0 = blockIdx.x * 2;
    const int N = a.n_slots, NM1 = N - 1, in_dim = a.obs_dim + a.latent_dim;
    const float* __restrict__ W = a.params + (int64_t)ag * a.param_stride;
    const GatLayout L = gat_layout(in_dim);

    const uint32_t raw_u = smem_u32(g5_raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;                 // swizzle atoms are 1024-byte aligned
    unsigned char* gb = g5_raw + (base - raw_u);
    const G5Layout Y = g5_layout(N, FUSED);
    float* s_we = reinterpret_cast<float*>(gb + Y.small);           // [H][IN_MAX] zero padded
    float* s_be = s_we + H * IN_MAX;                                // [H]
    float* s_pb = s_be + H;                                         // [2][96] gate-scaled b_ih (+ b_hh for r|z)
    float* s_bn = s_pb + 2 * G3;                                    // [2][32] K_N b_hn
    float* s_lw = s_bn + 2 * H;                                     // [2][32] logit-difference weights
    float* s_pl = s_lw + 2 * H;                                     // [2 tiles][2 step parities][128 rows] logit partial of unit half 1
    float* s_vb = s_pl + 2 * 2 * 128;                               // (fused) [32] v bias | [96] GRUCell b_ih | [96] b_hh | [4][128] max | [4][128] sum
    float* s_cb = s_vb + H;
    float* s_pm = s_cb + 2 * G3;
    float* s_ps = s_pm + 4 * 128;
    float* s_dl = reinterpret_cast<float*>(gb + Y.dl);              // (fused) [2 dirs][N-1][128 rows]
    const uint32_t bars = base + Y.bar;
    auto d_full = [&](int t) { return bars + 8u * (2 + t); };       // MMA issuer -> the tile's warps: accumulator complete
    const uint32_t pro_bar = bars + 32u, tmem_slot = bars + 40u, att_bar = bars + 48u;

    if (tid == 0) {
        for (int t = 0; t < 2; ++t) mbar_init(d_full(t), 1);
        mbar_init(pro_bar, 1);
        mbar_init(att_bar, 1);
        mbar_init_fence();
    }
    G5_STAMP(0);
    if (warp == 0) tc5_alloc<G5_TMEM_COLS>(tmem_slot);

    // ---- inputs of this thread's row (row = tid & 127; the four threads of a row each encode 8 of its 32 units) ----------
    float x[IN_MAX];
#pragma unroll
    for (int k = 0; k < IN_MAX; ++k) x[k] = 0.0f;
    bool row_ok = false;
    {
        const int xr = tid & 127, e = xr >> 6, i = xr & 63, b = b0 + e;
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
    constexpr int W_ITERS = (HH_TASKS + IH_TASKS + G5_THREADS - 1) / G5_THREADS;
    {
        // all of a thread's loads are issued before its first store: one L2 round trip instead of one per task
        float4 w0[W_ITERS], w1[W_ITERS];
        uint32_t dst[W_ITERS];
        float ksc[W_ITERS];
#pragma unroll
        for (int it = 0; it < W_ITERS; ++it) {
            const int task = tid + it * G5_THREADS;
            dst[it] = 0u; ksc[it] = 0.0f;
            w0[it] = w1[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (task < HH_TASKS + IH_TASKS) {
                const float* src;
                uint32_t tile;
                int row, ch, gate;
                if (task < HH_TASKS) {
                    const int d = task / (G3 * 4), rr = task - d * (G3 * 4);
                    row = rr >> 2; ch = rr & 3; gate = row;
                    src = W + (d ? L.whh_r : L.whh_f) + row * H + ch * 8;
                    tile = base + Y.bhh + d * G5_BHH_BYTES;
                } else {
                    const int tt = task - HH_TASKS, d = tt / (2 * G3 * 4), rr = tt - d * (2 * G3 * 4);
                    row = rr >> 2; ch = rr & 3;
                    const int part = row / G3;                      // 0: ego columns (-> P), 1: neighbour columns (-> Q)
                    gate = row - part * G3;
                    src = W + (d ? L.wih_r : L.wih_f) + gate * 2 * H + part * H + ch * 8;
                    tile = base + Y.q + d * G5_BIH_BYTES;
                }
                ksc[it] = gate < 2 * H ? K_RZ : K_N;
                w0[it] = *reinterpret_cast<const float4*>(src);
                w1[it] = *reinterpret_cast<const float4*>(src + 4);
                dst[it] = tile + swz128(row, ch);               // hi chunk `ch`; the lo chunk `4 + ch` is 64 bytes further (xor of bit 2 of the chunk index)
            }
        }
#pragma unroll
        for (int it = 0; it < W_ITERS; ++it) {
            if (dst[it]) {
                const float ks = ksc[it];
                uint32_t hi[4], lo[4];
                split_f16(ks * w0[it].x, ks * w0[it].y, hi[0], lo[0]);
                split_f16(ks * w0[it].z, ks * w0[it].w, hi[1], lo[1]);
                split_f16(ks * w1[it].x, ks * w1[it].y, hi[2], lo[2]);
                split_f16(ks * w1[it].z, ks * w1[it].w, hi[3], lo[3]);
                sts128(dst[it], hi[0], hi[1], hi[2], hi[3]);
                sts128(dst[it] ^ 64u, lo[0], lo[1], lo[2], lo[3]);
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    G5_STAMP(1);                                                  // constants + weight operand tiles staged
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // ---- enc = ReLU(W_e x + b_e) (:50) -> operand tile 0 (shared by both directions' [P | Q] products) ----
    {
        const uint32_t tile = base + Y.a;
        const int xr = tid & 127, ch = tid >> 7;                   // 8 units of the row: operand chunk `ch` (hi) and `4 + ch` (lo)
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
        sts128(tile + swz128(xr, ch), hi[0], hi[1], hi[2], hi[3]);
        sts128(tile + swz128(xr, 4 + ch), lo[0], lo[1], lo[2], lo[3]);
        if constexpr (FUSED) {                                    // the attention phase needs enc again; the h tile overwrites this one
            sts128(base + Y.enc + swz128(xr, ch), hi[0], hi[1], hi[2], hi[3]);
            sts128(base + Y.enc + swz128(xr, 4 + ch), lo[0], lo[1], lo[2], lo[3]);
        }
    }
    fence_proxy_async();
    __syncthreads();
    G5_STAMP(2);                                                  // enc tile written

    constexpr uint32_t IDESC_IH = tc5_idesc(128, 2 * G3), IDESC_HH = tc5_idesc(128, G3);
    constexpr int PQ_COL = 128;                               // [P | Q] fwd at TMEM columns 128..319, rev at 320..511
    if (warp == 0 && elect_one()) {
        tc5_fence_after();
        const uint64_t da = tc5_smem_desc(base + Y.a);
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const uint64_t db = tc5_smem_desc(base + Y.q + d * G5_BIH_BYTES);
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
        const uint32_t a_tile = base + Y.a + t * G5_A_BYTES;
        float* q_tab = reinterpret_cast<float*>(gb + Y.q + t * Y.q_dir);     // [2 envs x N slots][G5_QP]
        const uint32_t p_col = tlane + PQ_COL + t * 2 * G3;            // this chain's P: stays in TMEM for the whole recurrence

        mbar_wait(pro_bar, 0);
        tc5_fence_after();
        G5_STAMP(3);                                              // [P | Q] products complete
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
                float4* dst = reinterpret_cast<float4*>(q_tab + (e * N + (i < N ? i : 0)) * G5_QP + col);
                if (i < N) {                                   // rows of padding slots have no table entry
                    dst[0] = make_float4(v[0] + pb0.x, v[1] + pb0.y, v[2] + pb0.z, v[3] + pb0.w);
                    dst[1] = make_float4(v[4] + pb1.x, v[5] + pb1.y, v[6] + pb1.z, v[7] + pb1.w);
                }
            }
        tc5_fence_before();
        __syncthreads();                                        // Q tables complete, TMEM Q columns free
        tc5_fence_after();
        // The hidden state is the A operand of the step's product and lives in TENSOR MEMORY: 32 columns per tile, column c = the
        // f16 pair of units (2c, 2c+1) for c < 16 (hi parts), their lo parts at 16 + c.  The gate warps write it with tcgen05.st
        // (thread = TMEM lane = chain): no shared-memory store, no proxy fence, and the product reads only W_hh from shared memory
        // (an A tile in shared memory made each of the six MMAs fetch 4 KB + 3 KB: they ran at the shared-memory bandwidth).
        const uint32_t h_col = t ? (uint32_t)(PQ_COL + 3 * G3) : (uint32_t)G3;   // tile 0: columns 96..127; tile 1: 416..447 (reverse Q, spilled)
        const uint32_t h_st = tlane + h_col + 8 * hh;
#pragma unroll
        for (int k = 0; k < 2; ++k) {                            // h0 = 0 (:77)
            tc5_st4(h_st + 4 * k, 0u, 0u, 0u, 0u);
            tc5_st4(h_st + 16 + 4 * k, 0u, 0u, 0u, 0u);
        }
        tc5_wait_st();
        tc5_fence_before();
        asm volatile("barrier.sync %0, 256;" ::"r"(1 + t) : "memory");
        G5_STAMP(4);
        // the tile's product h . W_hh^T: issued by one lane once the tile's eight warps have written h (named barrier)
        const bool issuer_warp = (warp & 3) == 0 && hh == 0;         // warp-uniform
        const uint32_t mma_a = tmem_base + h_col;
        const uint64_t mma_b = tc5_smem_desc(base + Y.bhh + t * G5_BHH_BYTES);
        // accumulators: tile 0 at columns 0..95; tile 1 takes over the forward Q columns (224..319), free by now
        const uint32_t d_off = t ? (uint32_t)(PQ_COL + G3) : 0u;
        const uint32_t mma_d = tmem_base + d_off;
        auto issue = [&]() {
            tc5_fence_after();
            tc5_mma_ts(mma_d, mma_a + 0, mma_b + 0, IDESC_HH, 0);   // hi * hi   (A: 8 columns per K = 16 block)
            tc5_mma_ts(mma_d, mma_a + 8, mma_b + 2, IDESC_HH, 1);
            tc5_mma_ts(mma_d, mma_a + 16, mma_b + 0, IDESC_HH, 1);  // lo * hi
            tc5_mma_ts(mma_d, mma_a + 24, mma_b + 2, IDESC_HH, 1);
            tc5_mma_ts(mma_d, mma_a + 0, mma_b + 4, IDESC_HH, 1);   // hi * lo
            tc5_mma_ts(mma_d, mma_a + 8, mma_b + 6, IDESC_HH, 1);
            tc5_commit(d_full(t));
        };
        if (issuer_warp) { if (elect_one()) issue(); }          // step 0 (h = 0)

        f32x2 h2[8];                                            // this thread's 16 hidden units, fp32
#pragma unroll
        for (int p = 0; p < 8; ++p) h2[p] = pk2(0.0f, 0.0f);
        const f32x2 one2 = pk2(1.0f, 1.0f), mtwo2 = pk2(-2.0f, -2.0f);
        const uint32_t d_col = tlane + d_off;
        const float* q_env = q_tab + (e * N) * G5_QP;
        const float* bn = s_bn + t * H;
        const float* lw = s_lw + t * H;
        float* dlp = a.dl + ((((int64_t)ag * a.n_envs + (ok ? b : 0)) * 2 + t) * NM1) * DLP + i;
        float* plb = s_pl + t * 256 + row;

        for (int step = 0; step < ((DBG & 1) && t ? 0 : NM1); ++step) {
            const int s = t ? NM1 - 1 - step : step;
            const float* q = q_env + (s < i ? s : s + 1) * G5_QP;       // neighbour of ego i at position s (:60-66)
            G5_TRACE(0);
            mbar_wait(d_full(t), step & 1);
            tc5_fence_after();
            G5_TRACE(1);
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
                if (k == 0) G5_TRACE(2);
                f32x2 r[4], z[4], xx[4];
                auto pq = [&](float v0, float v1, float p0, float p1, float q0, float q1) -> f32x2 {   // D + P + Q
                    return add2(add2(pk2(v0, v1), pk2(p0, p1)), pk2(q0, q1));
                };
                xx[0] = pq(vr[0], vr[1], pr[0], pr[1], qr0.x, qr0.y);
                xx[1] = pq(vr[2], vr[3], pr[2], pr[3], qr0.z, qr0.w);
                xx[2] = pq(vr[4], vr[5], pr[4], pr[5], qr1.x, qr1.y);
                xx[3] = pq(vr[6], vr[7], pr[6], pr[7], qr1.z, qr1.w);
                auto sig4 = [&](f32x2 x01, f32x2 x23, f32x2& ia, f32x2& ib) {
                    if constexpr ((DBG & 2) != 0) { ia = fma2(x01, x23, one2); ib = fma2(x23, x01, mtwo2); }
                    else sigmoid4_den(x01, x23, ia, ib);
                };
                sig4(xx[0], xx[1], r[0], r[1]);                         // r = 1 / (1 + 2^x')
                sig4(xx[2], xx[3], r[2], r[3]);
                xx[0] = pq(vz[0], vz[1], pz[0], pz[1], qz0.x, qz0.y);
                xx[1] = pq(vz[2], vz[3], pz[2], pz[3], qz0.z, qz0.w);
                xx[2] = pq(vz[4], vz[5], pz[4], pz[5], qz1.x, qz1.y);
                xx[3] = pq(vz[6], vz[7], pz[6], pz[7], qz1.z, qz1.w);
                sig4(xx[0], xx[1], z[0], z[1]);
                sig4(xx[2], xx[3], z[2], z[3]);
                const float4 qn0 = *reinterpret_cast<const float4*>(q + 2 * H + 8 * c), qn1 = *reinterpret_cast<const float4*>(q + 2 * H + 8 * c + 4);
                const float4 bn0 = *reinterpret_cast<const float4*>(bn + 8 * c), bn1 = *reinterpret_cast<const float4*>(bn + 8 * c + 4);
                // n pre-activation: (W_in x + b_in) + r (W_hn h + b_hn)   (GRU gate order r, z, n)
                xx[0] = fma2(r[0], add2(pk2(vn[0], vn[1]), pk2(bn0.x, bn0.y)), add2(pk2(pn[0], pn[1]), pk2(qn0.x, qn0.y)));
                xx[1] = fma2(r[1], add2(pk2(vn[2], vn[3]), pk2(bn0.z, bn0.w)), add2(pk2(pn[2], pn[3]), pk2(qn0.z, qn0.w)));
                xx[2] = fma2(r[2], add2(pk2(vn[4], vn[5]), pk2(bn1.x, bn1.y)), add2(pk2(pn[4], pn[5]), pk2(qn1.x, qn1.y)));
                xx[3] = fma2(r[3], add2(pk2(vn[6], vn[7]), pk2(bn1.z, bn1.w)), add2(pk2(pn[6], pn[7]), pk2(qn1.z, qn1.w)));
                f32x2 in[4];
                sig4(xx[0], xx[1], in[0], in[1]);
                sig4(xx[2], xx[3], in[2], in[3]);
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
                tc5_st4(h_st + 4 * k, hi[0], hi[1], hi[2], hi[3]);
                tc5_st4(h_st + 16 + 4 * k, lo[0], lo[1], lo[2], lo[3]);
                if (k == 0) G5_TRACE(3); else G5_TRACE(4);
            }
            float pa, pb;
            upk2(pl, pa, pb);
            if (hh) plb[(step & 1) * 128] = pa + pb;                    // the other half's thread adds it after the barrier
            tc5_wait_st();                                              // this thread's h stores have landed in tensor memory
            tc5_fence_before();
            asm volatile("barrier.sync %0, 256;" ::"r"(1 + t) : "memory");   // the tile's 8 warps: h tile complete, D consumed
            G5_TRACE(5);
            if (issuer_warp && step + 1 < NM1) { if (elect_one()) issue(); }
            G5_TRACE(6);
            if (!hh) {
                const float dlv = (pa + pb) + plb[(step & 1) * 128];
                if constexpr (FUSED) s_dl[(t * NM1 + s) * 128 + row] = dlv;
                else if (ok) dlp[(int64_t)s * DLP] = dlv;               // lanes = consecutive egos: coalesced
            }
        }
    }
    if constexpr (FUSED) {
        // ================= attention + GRUCell (nova/GAT_Net.py:99-142), 4 threads per row =================
        // thread = (row, part): part p owns the neighbours j = p, p + 4, ... and the hidden units 8p .. 8p + 7 of its row
        const int part = warp >> 2, row = (warp & 3) * 32 + lane;
        const int e = row >> 6, i = row & 63, b = b0 + e;
        const bool ok = i < N && b < a.n_envs;
        const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const float* hprev = a.hprev.ptr + ag * a.hprev.stride_agent + (ok ? b : 0) * a.hprev.stride_env + (ok ? i : 0) * a.hprev.stride_slot;
        float* k_s = reinterpret_cast<float*>(gb + Y.q);                       // [128 rows][G5_KP]
        float* v_s = k_s + 128 * G5_KP;
        float* x_part = v_s + 128 * G5_KP;                                     // [3 parts][32 units][128 rows]
        const uint32_t t_hx = base + Y.a, t_wih = base + Y.a + G5_A_BYTES;     // h_prev / x operand tile | GRUCell W_ih tile
        const uint32_t t_qkv = base + Y.bhh, t_whh = base + Y.bhh + G5_BHH_BYTES;
        const uint32_t row_base = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u, rx = (uint32_t)(row & 7);

        tc5_fence_before();
        G5_STAMP(5);                           // this thread's recurrence done
        __syncthreads();                       // both recurrences done: dl tables complete, operand tiles and TMEM columns free
        tc5_fence_after();
        G5_STAMP(6);
        // ---- stage W_q|k|v, GRUCell W_hh, GRUCell W_ih as operand tiles; h_prev as an operand tile; biases -----------
        constexpr int A_ITERS = (3 * G3 * 4 + G5_THREADS - 1) / G5_THREADS;
        {
            float4 w0[A_ITERS], w1[A_ITERS];
            uint32_t dst[A_ITERS];
#pragma unroll
            for (int it = 0; it < A_ITERS; ++it) {
                const int task = tid + it * G5_THREADS;
                dst[it] = 0u;
                w0[it] = w1[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (task < 3 * G3 * 4) {
                    const int sel = task / (G3 * 4), rr = task - sel * (G3 * 4), wr = rr >> 2, ch = rr & 3;
                    const float* src = W + (sel == 0 ? L.q_w : (sel == 1 ? L.c_whh : L.c_wih)) + wr * H + ch * 8;
                    w0[it] = *reinterpret_cast<const float4*>(src);
                    w1[it] = *reinterpret_cast<const float4*>(src + 4);
                    dst[it] = (sel == 0 ? t_qkv : (sel == 1 ? t_whh : t_wih)) + swz128(wr, ch);
                }
            }
#pragma unroll
            for (int it = 0; it < A_ITERS; ++it) {
                if (dst[it]) {
                    uint32_t hi[4], lo[4];
                    split_f16(w0[it].x, w0[it].y, hi[0], lo[0]);
                    split_f16(w0[it].z, w0[it].w, hi[1], lo[1]);
                    split_f16(w1[it].x, w1[it].y, hi[2], lo[2]);
                    split_f16(w1[it].z, w1[it].w, hi[3], lo[3]);
                    sts128(dst[it], hi[0], hi[1], hi[2], hi[3]);
                    sts128(dst[it] ^ 64u, lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
        const float he_db = W[L.he_b + 1] - W[L.he_b + 0];
        float hp[8];                           // h_prev of this thread's 8 units (fp32, kept for the GRUCell's last line)
        {
#pragma unroll
            for (int q = 0; q < 8; ++q) hp[q] = ok ? hprev[8 * part + q] : 0.0f;      // (rows of the packed store are only 4-byte aligned)
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) split_f16(hp[2 * q], hp[2 * q + 1], hi[q], lo[q]);
            sts128(t_hx + row_base + (((uint32_t)part ^ rx) << 4), hi[0], hi[1], hi[2], hi[3]);
            sts128(t_hx + row_base + (((uint32_t)(4 + part) ^ rx) << 4), lo[0], lo[1], lo[2], lo[3]);
        }
        if (tid < H) s_vb[tid] = W[L.v_b + tid];
        if (tid < G3) { s_cb[tid] = W[L.c_bih + tid]; s_cb[G3 + tid] = W[L.c_bhh + tid]; }
        fence_proxy_async();
        tc5_fence_before();
        __syncthreads();
        tc5_fence_after();
        G5_STAMP(7);                           // attention operand tiles staged
        constexpr int QKV_COL = 0, GH_COL = G3, GI_COL = 2 * G3;
        auto product = [&](uint32_t col, uint32_t ta, uint32_t tb) {          // D[128 x 96] = A . B^T, fp32-class (3 passes)
            const uint64_t da = tc5_smem_desc(ta), db = tc5_smem_desc(tb);
            const uint32_t dst = tmem_base + col;
            tc5_mma(dst, da + 0, db + 0, IDESC_HH, 0);
            tc5_mma(dst, da + 2, db + 2, IDESC_HH, 1);
            tc5_mma(dst, da + 4, db + 0, IDESC_HH, 1);
            tc5_mma(dst, da + 6, db + 2, IDESC_HH, 1);
            tc5_mma(dst, da + 0, db + 4, IDESC_HH, 1);
            tc5_mma(dst, da + 2, db + 6, IDESC_HH, 1);
        };
        if (warp == 0 && elect_one()) {
            product(QKV_COL, base + Y.enc, t_qkv);                             // q | k | v = enc [W_q | W_k | W_v]^T      (:99-103)
            product(GH_COL, t_hx, t_whh);                                      // GRUCell: h_prev W_hh^T                   (:140)
            tc5_commit(att_bar);
        }
        mbar_wait(att_bar, 0);
        tc5_fence_after();
        G5_STAMP(8);                           // q|k|v and gh products complete
        float qv[H];
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            float v[8];
            tc5_ld8_nowait(tlane + QKV_COL + 8 * c8, v);
            tc5_wait_ld8(v);
#pragma unroll
            for (int q = 0; q < 8; ++q) qv[8 * c8 + q] = v[q];
        }
        {
            float kk[8], vv[8];
            tc5_ld8_nowait(tlane + QKV_COL + H + 8 * part, kk);
            tc5_ld8_nowait(tlane + QKV_COL + 2 * H + 8 * part, vv);
            tc5_wait_ld8(kk);
            tc5_wait_ld8(vv);
            float4* kd = reinterpret_cast<float4*>(k_s + row * G5_KP + 8 * part);
            float4* vd = reinterpret_cast<float4*>(v_s + row * G5_KP + 8 * part);
            kd[0] = make_float4(kk[0], kk[1], kk[2], kk[3]);
            kd[1] = make_float4(kk[4], kk[5], kk[6], kk[7]);
            const float* vb = s_vb + 8 * part;                                 // v = ReLU(W_v enc + b_v)                  (:103)
            vd[0] = make_float4(fmaxf(vv[0] + vb[0], 0.f), fmaxf(vv[1] + vb[1], 0.f), fmaxf(vv[2] + vb[2], 0.f), fmaxf(vv[3] + vb[3], 0.f));
            vd[1] = make_float4(fmaxf(vv[4] + vb[4], 0.f), fmaxf(vv[5] + vb[5], 0.f), fmaxf(vv[6] + vb[6], 0.f), fmaxf(vv[7] + vb[7], 0.f));
        }
        __syncthreads();
        G5_STAMP(9);                           // k / v tables written
        // ---- scores, hard gate, soft-max over the neighbours (:107-129); this thread's neighbours j = part + 4 u --------
        // neighbours of a thread: the four groups g of four consecutive slots j = 16 g + 4 part + w (one Philox call per group)
        constexpr int JG = IPLAN_MAX_SLOTS / 16, JU = 4 * JG;
        float sc[JU], hd[JU];
        const float db = he_db;
        const int64_t ego = ((int64_t)ag * a.n_envs + (ok ? b : 0)) * N + i;
        float mx = -INFINITY;
#pragma unroll
        for (int g = 0; g < JG; ++g) {
            const int j0 = 16 * g + 4 * part;
            uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
            if (!a.gumbel && j0 < N && i < N) {                                // same Philox stream as gat_attend_kernel: key (ego, j >> 2), word j & 3
                const int64_t key = ego * 16 + (j0 >> 2);
                rnd = philox4x32(make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)a.counter, (uint32_t)(a.counter >> 32)),
                                 make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int u = 4 * g + w, j = j0 + w;
                sc[u] = -INFINITY; hd[u] = 0.0f;
                if (j < N && j != i && i < N) {
                    const float4* kr = reinterpret_cast<const float4*>(k_s + (e * 64 + j) * G5_KP);
                    float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
                    for (int c4 = 0; c4 < H / 4; c4 += 2) {
                        const float4 ka = kr[c4], kb = kr[c4 + 1];
                        d0 = fmaf(qv[4 * c4], ka.x, d0); d0 = fmaf(qv[4 * c4 + 1], ka.y, d0);
                        d0 = fmaf(qv[4 * c4 + 2], ka.z, d0); d0 = fmaf(qv[4 * c4 + 3], ka.w, d0);
                        d1 = fmaf(qv[4 * c4 + 4], kb.x, d1); d1 = fmaf(qv[4 * c4 + 5], kb.y, d1);
                        d1 = fmaf(qv[4 * c4 + 6], kb.z, d1); d1 = fmaf(qv[4 * c4 + 7], kb.w, d1);
                    }
                    sc[u] = (d0 + d1) * 0.17677669529663687f;                  // / np.sqrt(attention_dim), :126
                    const int s = j < i ? j : j - 1;                           // position of neighbour j in ego i's sequence
                    float noise;
                    if (a.gumbel) {
                        const int64_t edge = ego * NM1 + s;
                        noise = ok ? a.gumbel[2 * edge + 1] - a.gumbel[2 * edge] : 0.0f;
                    } else {
                        const float uu = u01(w == 0 ? rnd.x : (w == 1 ? rnd.y : (w == 2 ? rnd.z : rnd.w)));
                        noise = __logf(uu) - __logf(1.0f - uu);                // Gumbel - Gumbel ~ Logistic(0,1)
                    }
                    const float dlog = (s_dl[s * 128 + row] + s_dl[(NM1 + s) * 128 + row]) + db;
                    hd[u] = __fdividef(1.0f, 1.0f + expf(-(dlog + noise) * a.inv_tau));   // gumbel-softmax(tau)[..., 1]   (:93-95)
                    if (a.dbg_hard && ok) a.dbg_hard[ego * NM1 + s] = hd[u];
                    mx = fmaxf(mx, sc[u]);
                }
            }
        }
        s_pm[part * 128 + row] = mx;
        __syncthreads();
        G5_STAMP(10);                          // scores + hard gates
        mx = fmaxf(fmaxf(s_pm[row], s_pm[128 + row]), fmaxf(s_pm[256 + row], s_pm[384 + row]));
        float psum = 0.0f;
#pragma unroll
        for (int u = 0; u < JU; ++u) { sc[u] = expf(sc[u] - mx); psum += sc[u]; }     // exp(-inf) = 0 for self / padding
        s_ps[part * 128 + row] = psum;
        __syncthreads();
        G5_STAMP(11);
        const float rden = 1.0f / (((s_ps[row] + s_ps[128 + row]) + s_ps[256 + row]) + s_ps[384 + row]);
        // ---- x_i = sum_j soft_ij hard_ij v_j (no renormalisation, :132): this part's neighbours, then the four parts -----
        float xa[H];
#pragma unroll
        for (int c = 0; c < H; ++c) xa[c] = 0.0f;
#pragma unroll
        for (int u = 0; u < JU; ++u) {
            const int j = 16 * (u >> 2) + 4 * part + (u & 3);
            if (j < N && j != i && i < N) {
                const float wgt = (sc[u] * rden) * hd[u];
                const float4* vr = reinterpret_cast<const float4*>(v_s + (e * 64 + j) * G5_KP);
#pragma unroll
                for (int c4 = 0; c4 < H / 4; ++c4) {
                    const float4 vq = vr[c4];
                    xa[4 * c4] = fmaf(wgt, vq.x, xa[4 * c4]); xa[4 * c4 + 1] = fmaf(wgt, vq.y, xa[4 * c4 + 1]);
                    xa[4 * c4 + 2] = fmaf(wgt, vq.z, xa[4 * c4 + 2]); xa[4 * c4 + 3] = fmaf(wgt, vq.w, xa[4 * c4 + 3]);
                }
            }
        }
        if (part) {
#pragma unroll
            for (int c = 0; c < H; ++c) x_part[((part - 1) * H + c) * 128 + row] = xa[c];
        }
        __syncthreads();
        G5_STAMP(12);                          // aggregation partials
        if (!part) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v2[2];
#pragma unroll
                    for (int w2 = 0; w2 < 2; ++w2) {
                        const int c = 8 * ch + 2 * q + w2;
                        v2[w2] = ((xa[c] + x_part[c * 128 + row]) + x_part[(H + c) * 128 + row]) + x_part[(2 * H + c) * 128 + row];
                    }
                    split_f16(v2[0], v2[1], hi[q], lo[q]);
                }
                sts128(t_hx + row_base + (((uint32_t)ch ^ rx) << 4), hi[0], hi[1], hi[2], hi[3]);     // h_prev tile is dead: its product is complete
                sts128(t_hx + row_base + (((uint32_t)(4 + ch) ^ rx) << 4), lo[0], lo[1], lo[2], lo[3]);
            }
        }
        fence_proxy_async();
        tc5_fence_before();
        __syncthreads();
        tc5_fence_after();
        if (warp == 0 && elect_one()) {
            product(GI_COL, t_hx, t_wih);                                      // GRUCell: x W_ih^T                        (:140)
            tc5_commit(att_bar);
        }
        G5_STAMP(13);                          // x operand tile written, product issued
        mbar_wait(att_bar, 1);
        tc5_fence_after();
        G5_STAMP(14);
        {
            float gir[8], giz[8], gin[8], ghr[8], ghz[8], ghn[8];
            tc5_ld8_nowait(tlane + GI_COL + 8 * part, gir);
            tc5_ld8_nowait(tlane + GI_COL + H + 8 * part, giz);
            tc5_ld8_nowait(tlane + GI_COL + 2 * H + 8 * part, gin);
            tc5_ld8_nowait(tlane + GH_COL + 8 * part, ghr);
            tc5_ld8_nowait(tlane + GH_COL + H + 8 * part, ghz);
            tc5_ld8_nowait(tlane + GH_COL + 2 * H + 8 * part, ghn);
            tc5_wait_ld24(gir, giz, gin);
            tc5_wait_ld24(ghr, ghz, ghn);
            float* outp = a.out.ptr + ag * a.out.stride_agent + (ok ? b : 0) * a.out.stride_env + (ok ? i : 0) * a.out.stride_slot;
            const float* bi = s_cb, * bh = s_cb + G3;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int c = 8 * part + q;
                // ex2.approx / rcp.approx forms (abs error ~1e-7, as in the recurrence); the clamp keeps 2^x finite
                const float r = rcp_approx(1.0f + ex2_approx(fminf(K_RZ * ((gir[q] + bi[c]) + (ghr[q] + bh[c])), 80.0f)));
                const float z = rcp_approx(1.0f + ex2_approx(fminf(K_RZ * ((giz[q] + bi[H + c]) + (ghz[q] + bh[H + c])), 80.0f)));
                const float nn = fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(fminf(K_N * ((gin[q] + bi[2 * H + c]) + r * (ghn[q] + bh[2 * H + c])), 80.0f))), 1.0f);
                if (ok) outp[c] = (1.0f - z) * nn + z * hp[q];
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    G5_STAMP(15);
    if (warp == 0) {
        tc5_fence_after();
        tc5_dealloc<G5_TMEM_COLS>(tmem_base);
    }
}

// fused == true: the whole K1 step in ONE launch (falls back to false when the fused shared-memory map does not fit, i.e.
// n_slots close to IPLAN_MAX_SLOTS); fused == false: the recurrence only, dl scratch for gat_attend_kernel.
// Returns 0 / error; *did_fuse tells the caller whether the attention kernel is still needed.
int launch_gat_tc5(const GatArgs& a, int n_agents, bool fused, bool* did_fuse, cudaStream_t st) {
    static bool configured = false;
    static int dbg = 0;
    if (!configured) {
        const char* ev = getenv("IPLAN_GAT_DBG");
        dbg = ev ? atoi(ev) : 0;
        cudaError_t e = cudaFuncSetAttribute(gat_tc5_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G5_SMEM_MAX);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gat_tc5_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G5_SMEM_MAX);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gat_tc5_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G5_SMEM_MAX);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gat_tc5_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G5_SMEM_MAX);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gat_tc5_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G5_SMEM_MAX);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gat_tc5_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G5_SMEM_MAX);
        if (e != cudaSuccess) { set_error("gat_step: tcgen05 kernel smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    if (fused && g5_layout(a.n_slots, true).total > G5_SMEM_MAX) fused = false;
    if (dbg & ~4) fused = false;
    const G5Layout Y = g5_layout(a.n_slots, fused);
    if (Y.total > G5_SMEM_MAX) { set_error("gat_step: n_slots %d needs %u bytes of shared memory", a.n_slots, Y.total); return -1; }
    const dim3 grid((a.n_envs + 1) / 2, n_agents);
    if (fused && dbg == 4) gat_tc5_kernel<4, true><<<grid, G5_THREADS, Y.total, st>>>(a);
    else if (fused) gat_tc5_kernel<0, true><<<grid, G5_THREADS, Y.total, st>>>(a);
    else if (dbg == 1) gat_tc5_kernel<1, false><<<grid, G5_THREADS, Y.total, st>>>(a);
    else if (dbg == 2) gat_tc5_kernel<2, false><<<grid, G5_THREADS, Y.total, st>>>(a);
    else if (dbg == 64) gat_tc5_kernel<64, false><<<grid, G5_THREADS, Y.total, st>>>(a);
    else gat_tc5_kernel<0, false><<<grid, G5_THREADS, Y.total, st>>>(a);
    count_launch();
    if (did_fuse) *did_fuse = fused;
    return check_launch(fused ? "gat_step(fused, tcgen05)" : "gat_step(recur, tcgen05)");
}

}  // namespace iplan

// timing experiments (IPLAN_GAT_DBG=4): the phase-boundary clock stamps of one CTA of the last fused launch
extern "C" int iplan_gat_debug_trace(long long* out256) {
    const cudaError_t e = cudaMemcpyFromSymbol(out256, iplan::g5_trace, sizeof(long long) * 256);
    if (e != cudaSuccess) { iplan::set_error("gat_debug_trace: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
extern "C" int iplan_gat_debug_clocks(long long* out32) {
    const cudaError_t e = cudaMemcpyFromSymbol(out32, iplan::g5_clk, sizeof(long long) * 32);
    if (e != cudaSuccess) { iplan::set_error("gat_debug_clocks: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
