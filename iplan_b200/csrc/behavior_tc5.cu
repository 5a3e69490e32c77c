// K1b on the 5th-generation tensor cores: Behavior_policy.latent_update + EncoderRNN.forward (reference
// nova/stable_behavior_policy.py:83-123, nova/behavior_net.py:17-22) in the structure of K1's recurrence (gat_tc5.cu).
//
// One CTA = 256 chains (chain = one (env, slot) of one agent-net) = two M = 128 tiles; TMEM lane r of tile t is chain
// 256 blockIdx.x + 128 t + r.  Per window step w and tile
//
//     D_h[128 x 96] = h[128 x 32] . W_hh^T        D_u[128 x 96] = u_w[128 x 32] . W_ih^T        u_w = ReLU(W_l x_w + b_l)
//
// twelve tcgen05.mma.kind::f16 (M 128, N 96, K 16; f16 hi/lo split of both operands: hi*hi + lo*hi + hi*lo per k-block),
// BOTH A operands (h and u_w) read from tensor memory, where the gate warps put them with tcgen05.st (thread = TMEM lane =
// chain): per tile 96 + 96 accumulator columns, 32 + 32 operand columns; only W_hh / W_ih (SWIZZLE_128B tiles, the gate-
// activation scales folded in, shared by both tiles) come from shared memory.  Eight gate warps per tile (warp -> TMEM lane
// quarter and one half of the hidden units; thread = one chain x 16 units): tcgen05.ld the r|z|n columns of D_h and D_u,
// add the biases, sigmoid / tanh through ex2 + a shared rcp (gat_common.cuh), new h -> f16 hi/lo -> tensor memory, and
// the NEXT step's u (five FMAs per unit from the chain's window rows, staged once per CTA in shared memory) right behind
// it (computed in registers while the products are still in flight); one elected lane per tile issues the step's twelve
// MMAs after the tile's named barrier.  After the last step:
// hidden state out, latent head (8 x 32, two partial sums per chain), soft-max, soft update.
//
// The mma.sync kernel (behavior_step.cu) runs at the legacy tensor pipe's limit (ncu: tensor pipe 54 % busy with three
// passes per product, math-pipe throttle the top stall); it stays as the cross-check (iplan_behavior_set_impl(1)) and as
// the fallback for shapes this kernel does not take (obs_dim > 8, window longer than 64 values, odd strides).
#include "common.cuh"
#include "behavior_common.cuh"
#include "gat_common.cuh"
#include "tc5.cuh"

#include <stdlib.h>

namespace iplan {

constexpr int B5_THREADS = 512;
constexpr int B5_TMEM_COLS = 512;
constexpr int B5_CHAINS = 256;
constexpr int B5_W_BYTES = G3 * 128;           // a weight operand tile: 96 rows x (32 hi + 32 lo) f16
constexpr int B5_WIN_MAX = 64, B5_LAT = 8, B5_OBS = 8;
// small constants (floats): [W_l^T ; b_l] [9][32] | gate bias row [96] | K_N b_hn [32] | W_o [8][32] | b_o [8] | head partials [256][8]
constexpr int B5_SMALL = H * B5_OBS + H + G3 + H + B5_LAT * H + B5_LAT + B5_CHAINS * B5_LAT;

struct B5Layout { uint32_t whh, wih, win, small, bar, total; int pitch; };
__host__ __device__ inline B5Layout b5_layout(int wo) {
    B5Layout l;
    l.whh = 0; l.wih = B5_W_BYTES; l.win = 2 * B5_W_BYTES;
    l.pitch = wo | 1;                                           // odd row pitch: lanes (consecutive chains) fall in different banks
    l.small = (l.win + (uint32_t)(B5_CHAINS * l.pitch * 4) + 15u) & ~15u;
    l.bar = l.small + B5_SMALL * 4;
    l.total = l.bar + 64 + 1024;
    return l;
}

// DBG (IPLAN_BEH_DBG=1): thread 0 of CTA (1, 0) stamps clock64() at its phase boundaries (read back with iplan_behavior_debug_clocks)
__device__ long long b5_clk[64];
#define B5_STAMP(k)                                                                          \
    do {                                                                                     \
        if constexpr (DBG) {                                                                 \
            if (blockIdx.x == 1 && blockIdx.y == 0 && tid == 0) b5_clk[k] = clock64();       \
        }                                                                                    \
    } while (0)
template <bool DBG>
__global__ void __launch_bounds__(B5_THREADS, 1) behavior_tc5_kernel(BehArgs a) {
    extern __shared__ unsigned char b5_raw[];
    // warp index through a shuffle: warp-uniform for the compiler (see gat_tc5.cu: no register-to-uniform loops at the MMA issue)
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ag = blockIdx.y;
    const int N = a.n_slots, o = a.obs_dim, Wn = a.hist_len, Ld = a.latent_dim, wo = Wn * o;
    const int total = a.n_envs * N;
    const float* __restrict__ P = a.params + (int64_t)ag * a.param_stride;
    const BehLayout L = beh_layout(o, Ld);

    const uint32_t raw_u = smem_u32(b5_raw);
    const uint32_t base = (raw_u + 1023u) & ~1023u;             // swizzle atoms are 1024-byte aligned
    unsigned char* gb = b5_raw + (base - raw_u);
    const B5Layout Y = b5_layout(wo);
    float* s_win = reinterpret_cast<float*>(gb + Y.win);        // [256 chains][pitch]
    float* s_lw = reinterpret_cast<float*>(gb + Y.small);       // [k = 0 .. 8][32 units]: W_l transposed, zero padded, b_l at row obs_dim
    float* s_qb = s_lw + (B5_OBS + 1) * H;                                     // [96] K_RZ (b_ih + b_hh) for r|z, K_N b_in
    float* s_bn = s_qb + G3;                                    // [32] K_N b_hn
    float* s_ow = s_bn + H;                                     // [8][32] zero padded
    float* s_ob = s_ow + B5_LAT * H;
    float* s_part = s_ob + B5_LAT;                              // [256][8] head partials of the upper unit half
    const uint32_t bars = base + Y.bar;
    auto d_full = [&](int t) { return bars + 8u * t; };         // MMA issuer -> the tile's warps: accumulators complete
    const uint32_t tmem_slot = bars + 32u;

    if (tid == 0) {
        mbar_init(d_full(0), 1);
        mbar_init(d_full(1), 1);
        mbar_init_fence();
    }
    B5_STAMP(0);
    if (warp == 0) tc5_alloc<B5_TMEM_COLS>(tmem_slot);
    B5_STAMP(1);

    // thread = chain `row` of tile t, hidden units 16 hh .. 16 hh + 15 (the recurrence below); its carried hidden state is
    // requested first, so that the load is back long before the operand tiles are staged
    const int t = (warp >> 2) & 1, hh = warp >> 3, row = (warp & 3) * 32 + lane;
    const int c = t * 128 + row, q = blockIdx.x * B5_CHAINS + c;
    const bool ok = q < total;
    const int bb = ok ? q / N : 0, nn = ok ? q - bb * N : 0;
    float4 h0v[4];
    {
        const float* hp = a.hid.ptr + ag * a.hid.stride_agent + bb * a.hid.stride_env + nn * a.hid.stride_slot + 16 * hh;
#pragma unroll
        for (int p = 0; p < 4; ++p) h0v[p] = ok ? *reinterpret_cast<const float4*>(hp + 4 * p) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ---- constants: every load is issued here, the stores follow the operand-tile loads below (one trip to L2 for the lot) ------
    // W_l is kept transposed, [k][unit], with b_l as row `o`: the 16 units of a thread are four LDS.128 per input column
    float c_lw[(B5_OBS + 1) * H / B5_THREADS + 1], c_ow = 0.0f, c_qb = 0.0f, c_bn = 0.0f, c_ob = 0.0f;
#pragma unroll
    for (int it = 0; it < (B5_OBS + 1) * H / B5_THREADS + 1; ++it) {
        const int idx = tid + it * B5_THREADS, k = idx / H, un = idx - k * H;
        c_lw[it] = 0.0f;
        if (idx < (B5_OBS + 1) * H) c_lw[it] = k < o ? P[L.lin_w + un * o + k] : (k == o ? P[L.lin_b + un] : 0.0f);
    }
    if (tid < B5_LAT * H) c_ow = tid / H < Ld ? P[L.out_w + tid] : 0.0f;
    if (tid < G3) c_qb = tid < 2 * H ? K_RZ * (P[L.bih + tid] + P[L.bhh + tid]) : K_N * P[L.bih + tid];
    if (tid < H) c_bn = K_N * P[L.bhh + 2 * H + tid];
    if (tid < B5_LAT) c_ob = tid < Ld ? P[L.out_b + tid] : 0.0f;
    // ---- weight operand tiles (f16 hi | lo, gate-activation scale folded in; one task = 8 consecutive k of one row) and the
    //      chains' windows -> shared memory.  Every global load of the prologue is issued before the first dependent store: one
    //      round trip to L2 instead of one per loop iteration.  Window: thread = (chain tid >> 1, half of its rows): no
    //      integer division per element. ------------------------------------------------------------------------------------
    {
        constexpr int W_TASKS = 2 * G3 * 4, W_IT = (W_TASKS + B5_THREADS - 1) / B5_THREADS;
        constexpr int RH_MAX = 8;                                   // window rows per thread (hist_len <= 16)
        float4 w0[W_IT], w1[W_IT];
#pragma unroll
        for (int it = 0; it < W_IT; ++it) {
            const int task = tid + it * B5_THREADS;
            w0[it] = w1[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (task < W_TASKS) {
                const int m = task / (G3 * 4), rr = task - m * (G3 * 4), row = rr >> 2, ch = rr & 3;
                const float* src = P + (m ? L.wih : L.whh) + row * H + ch * 8;
                w0[it] = *reinterpret_cast<const float4*>(src);
                w1[it] = *reinterpret_cast<const float4*>(src + 4);
            }
        }
        const int wc = tid >> 1, half = tid & 1, rh = (Wn + 1) >> 1, r0 = half * rh;
        const int wq = blockIdx.x * B5_CHAINS + wc;
        float xv[RH_MAX][B5_OBS];
        {
            const bool on = wq < total;
            const int wb = on ? wq / N : 0, wn = on ? wq - wb * N : 0;
            const float* src = a.window.ptr + ag * a.window.stride_agent + wb * a.window.stride_env + wn * a.window.stride_slot;
            const int64_t rstep = a.win_step ? a.win_step : (int64_t)o;
#pragma unroll
            for (int r = 0; r < RH_MAX; ++r) {
                const int w = r0 + r;
                const bool ld = on && r < rh && w < Wn && w >= a.win_pad;
                const float* rp = src + (int64_t)(ld ? w - a.win_pad : 0) * rstep;
#pragma unroll
                for (int k = 0; k < B5_OBS; ++k) xv[r][k] = (ld && k < o) ? rp[k] : 0.0f;
            }
        }
#pragma unroll
        for (int it = 0; it < W_IT; ++it) {
            const int task = tid + it * B5_THREADS;
            if (task < W_TASKS) {
                const int m = task / (G3 * 4), rr = task - m * (G3 * 4), row = rr >> 2, ch = rr & 3;
                const float ks = row < 2 * H ? K_RZ : K_N;
                uint32_t hi[4], lo[4];
                split_f16(ks * w0[it].x, ks * w0[it].y, hi[0], lo[0]);
                split_f16(ks * w0[it].z, ks * w0[it].w, hi[1], lo[1]);
                split_f16(ks * w1[it].x, ks * w1[it].y, hi[2], lo[2]);
                split_f16(ks * w1[it].z, ks * w1[it].w, hi[3], lo[3]);
                const uint32_t dst = base + (m ? Y.wih : Y.whh) + swz128(row, ch);     // hi chunk `ch`; the lo chunk `4 + ch` is 64 bytes further
                sts128(dst, hi[0], hi[1], hi[2], hi[3]);
                sts128(dst ^ 64u, lo[0], lo[1], lo[2], lo[3]);
            }
        }
        float* wdst = s_win + wc * Y.pitch;
#pragma unroll
        for (int r = 0; r < RH_MAX; ++r) {
            const int w = r0 + r;
            if (r < rh && w < Wn) {
#pragma unroll
                for (int k = 0; k < B5_OBS; ++k)
                    if (k < o) wdst[w * o + k] = xv[r][k];
            }
        }
    }
#pragma unroll
    for (int it = 0; it < (B5_OBS + 1) * H / B5_THREADS + 1; ++it)
        if (tid + it * B5_THREADS < (B5_OBS + 1) * H) s_lw[tid + it * B5_THREADS] = c_lw[it];
    if (tid < B5_LAT * H) s_ow[tid] = c_ow;
    if (tid < G3) s_qb[tid] = c_qb;
    if (tid < H) s_bn[tid] = c_bn;
    if (tid < B5_LAT) s_ob[tid] = c_ob;
    B5_STAMP(2);                                                // operand tiles, windows and constants stored
    fence_proxy_async();                                        // the weight tiles are read by the async proxy (tcgen05.mma)
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    B5_STAMP(3);
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    // ================= thread = chain `row` of tile t, hidden units 16 hh .. 16 hh + 15 =================
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t cb = 256u * t;                               // tile t: D_h cb .. +95, D_u +96 .. +191, h operand +192 .. +223, u operand +224 .. +255
    const uint32_t dh_col = tlane + cb, du_col = tlane + cb + G3;
    const uint32_t h_st = tlane + cb + 2 * G3 + 8 * hh, u_st = h_st + 32;
    const float* win = s_win + c * Y.pitch;

    f32x2 h2[8];                                                // this thread's 16 hidden units, fp32
#pragma unroll
    for (int p = 0; p < 4; ++p) { h2[2 * p] = pk2(h0v[p].x, h0v[p].y); h2[2 * p + 1] = pk2(h0v[p].z, h0v[p].w); }
    auto store_operand = [&](uint32_t st, const f32x2 (&v)[8]) {   // 16 values -> f16 hi | lo pairs -> the operand columns of this lane
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) split_f16p(v[4 * k + p], hi[p], lo[p]);
            tc5_st4(st + 4 * k, hi[0], hi[1], hi[2], hi[3]);
            tc5_st4(st + 16 + 4 * k, lo[0], lo[1], lo[2], lo[3]);
        }
    };
    f32x2 u2[8];                                                // the NEXT step's input-layer output for this thread's 16 units
    auto input_layer = [&](int w) {                             // u_w = ReLU(W_l x_w + b_l) -> u2 (registers)
        const float4* lw4 = reinterpret_cast<const float4*>(s_lw + 16 * hh);
        float acc[16];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float4 b4 = lw4[(o * H) / 4 + p];
            acc[4 * p] = b4.x; acc[4 * p + 1] = b4.y; acc[4 * p + 2] = b4.z; acc[4 * p + 3] = b4.w;
        }
#pragma unroll
        for (int k = 0; k < B5_OBS; ++k) {
            if (k < o) {
                const float x = win[w * o + k];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float4 w4 = lw4[(k * H) / 4 + p];
                    acc[4 * p] = fmaf(w4.x, x, acc[4 * p]); acc[4 * p + 1] = fmaf(w4.y, x, acc[4 * p + 1]);
                    acc[4 * p + 2] = fmaf(w4.z, x, acc[4 * p + 2]); acc[4 * p + 3] = fmaf(w4.w, x, acc[4 * p + 3]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) u2[p] = pk2(fmaxf(acc[2 * p], 0.0f), fmaxf(acc[2 * p + 1], 0.0f));
    };
    store_operand(h_st, h2);
    input_layer(0);
    store_operand(u_st, u2);
    tc5_wait_st();
    tc5_fence_before();
    asm volatile("barrier.sync %0, 256;" ::"r"(1 + t) : "memory");

    constexpr uint32_t IDESC = tc5_idesc(128, G3);
    const bool issuer_warp = (warp & 3) == 0 && hh == 0;        // warp-uniform
    const uint32_t mma_dh = tmem_base + cb, mma_du = mma_dh + G3, mma_ah = mma_dh + 2 * G3, mma_au = mma_ah + 32;
    const uint64_t b_hh = tc5_smem_desc(base + Y.whh), b_ih = tc5_smem_desc(base + Y.wih);
    auto issue = [&]() {
        tc5_fence_after();
        tc5_mma_ts(mma_dh, mma_ah + 0, b_hh + 0, IDESC, 0);     // hi * hi   (A: 8 columns per K = 16 block)
        tc5_mma_ts(mma_dh, mma_ah + 8, b_hh + 2, IDESC, 1);
        tc5_mma_ts(mma_dh, mma_ah + 16, b_hh + 0, IDESC, 1);    // lo * hi
        tc5_mma_ts(mma_dh, mma_ah + 24, b_hh + 2, IDESC, 1);
        tc5_mma_ts(mma_dh, mma_ah + 0, b_hh + 4, IDESC, 1);     // hi * lo
        tc5_mma_ts(mma_dh, mma_ah + 8, b_hh + 6, IDESC, 1);
        tc5_mma_ts(mma_du, mma_au + 0, b_ih + 0, IDESC, 0);
        tc5_mma_ts(mma_du, mma_au + 8, b_ih + 2, IDESC, 1);
        tc5_mma_ts(mma_du, mma_au + 16, b_ih + 0, IDESC, 1);
        tc5_mma_ts(mma_du, mma_au + 24, b_ih + 2, IDESC, 1);
        tc5_mma_ts(mma_du, mma_au + 0, b_ih + 4, IDESC, 1);
        tc5_mma_ts(mma_du, mma_au + 8, b_ih + 6, IDESC, 1);
        tc5_commit(d_full(t));
    };
    if (issuer_warp) { if (elect_one()) issue(); }
    B5_STAMP(4);                                                // first operands written, first products issued

    const f32x2 one2 = pk2(1.0f, 1.0f), mtwo2 = pk2(-2.0f, -2.0f);
    for (int w = 0; w < Wn; ++w) {
        if (w < 10) B5_STAMP(8 + 4 * w);
        if (w + 1 < Wn) input_layer(w + 1);                     // under the MMA round trip: the next step's input layer, in registers
        if (w < 10) B5_STAMP(9 + 4 * w);
        mbar_wait(d_full(t), w & 1);
        tc5_fence_after();
        if (w < 10) B5_STAMP(10 + 4 * w);
#pragma unroll
        for (int k = 0; k < 2; ++k) {                           // hidden units 8 c8 .. 8 c8 + 7
            const int c8 = 2 * hh + k;
            float vr[8], vz[8], vn[8], pr[8], pz[8], pn[8];
            tc5_ld8_nowait(dh_col + 8 * c8, vr);
            tc5_ld8_nowait(dh_col + H + 8 * c8, vz);
            tc5_ld8_nowait(dh_col + 2 * H + 8 * c8, vn);
            tc5_ld8_nowait(du_col + 8 * c8, pr);
            tc5_ld8_nowait(du_col + H + 8 * c8, pz);
            tc5_ld8_nowait(du_col + 2 * H + 8 * c8, pn);
            const float4 qr0 = *reinterpret_cast<const float4*>(s_qb + 8 * c8), qr1 = *reinterpret_cast<const float4*>(s_qb + 8 * c8 + 4);
            const float4 qz0 = *reinterpret_cast<const float4*>(s_qb + H + 8 * c8), qz1 = *reinterpret_cast<const float4*>(s_qb + H + 8 * c8 + 4);
            tc5_wait_ld24(vr, vz, vn);
            tc5_wait_ld24(pr, pz, pn);
            f32x2 r[4], z[4], xx[4];
            auto pq = [&](float v0, float v1, float p0, float p1, float q0, float q1) -> f32x2 {   // D_h + D_u + bias
                return add2(add2(pk2(v0, v1), pk2(p0, p1)), pk2(q0, q1));
            };
            xx[0] = pq(vr[0], vr[1], pr[0], pr[1], qr0.x, qr0.y);
            xx[1] = pq(vr[2], vr[3], pr[2], pr[3], qr0.z, qr0.w);
            xx[2] = pq(vr[4], vr[5], pr[4], pr[5], qr1.x, qr1.y);
            xx[3] = pq(vr[6], vr[7], pr[6], pr[7], qr1.z, qr1.w);
            sigmoid4_den(xx[0], xx[1], r[0], r[1]);             // r = 1 / (1 + 2^x')
            sigmoid4_den(xx[2], xx[3], r[2], r[3]);
            xx[0] = pq(vz[0], vz[1], pz[0], pz[1], qz0.x, qz0.y);
            xx[1] = pq(vz[2], vz[3], pz[2], pz[3], qz0.z, qz0.w);
            xx[2] = pq(vz[4], vz[5], pz[4], pz[5], qz1.x, qz1.y);
            xx[3] = pq(vz[6], vz[7], pz[6], pz[7], qz1.z, qz1.w);
            sigmoid4_den(xx[0], xx[1], z[0], z[1]);
            sigmoid4_den(xx[2], xx[3], z[2], z[3]);
            const float4 qn0 = *reinterpret_cast<const float4*>(s_qb + 2 * H + 8 * c8), qn1 = *reinterpret_cast<const float4*>(s_qb + 2 * H + 8 * c8 + 4);
            const float4 bn0 = *reinterpret_cast<const float4*>(s_bn + 8 * c8), bn1 = *reinterpret_cast<const float4*>(s_bn + 8 * c8 + 4);
            // n pre-activation: (W_in u + b_in) + r (W_hn h + b_hn)   (GRU gate order r, z, n)
            xx[0] = fma2(r[0], add2(pk2(vn[0], vn[1]), pk2(bn0.x, bn0.y)), add2(pk2(pn[0], pn[1]), pk2(qn0.x, qn0.y)));
            xx[1] = fma2(r[1], add2(pk2(vn[2], vn[3]), pk2(bn0.z, bn0.w)), add2(pk2(pn[2], pn[3]), pk2(qn0.z, qn0.w)));
            xx[2] = fma2(r[2], add2(pk2(vn[4], vn[5]), pk2(bn1.x, bn1.y)), add2(pk2(pn[4], pn[5]), pk2(qn1.x, qn1.y)));
            xx[3] = fma2(r[3], add2(pk2(vn[6], vn[7]), pk2(bn1.z, bn1.w)), add2(pk2(pn[6], pn[7]), pk2(qn1.z, qn1.w)));
            f32x2 in[4];
            sigmoid4_den(xx[0], xx[1], in[0], in[1]);
            sigmoid4_den(xx[2], xx[3], in[2], in[3]);
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const f32x2 nv = fma2(mtwo2, in[p], one2);                          // tanh = 1 - 2 / (1 + 2^x')
                h2[4 * k + p] = fma2(z[p], sub2(h2[4 * k + p], nv), nv);            // (1 - z) n + z h
                split_f16p(h2[4 * k + p], hi[p], lo[p]);
            }
            tc5_st4(h_st + 4 * k, hi[0], hi[1], hi[2], hi[3]);
            tc5_st4(h_st + 16 + 4 * k, lo[0], lo[1], lo[2], lo[3]);
        }
        if (w + 1 < Wn) store_operand(u_st, u2);                // the step's products are complete: the u operand is free
        tc5_wait_st();                                          // this thread's operand stores have landed in tensor memory
        if (w < 10) B5_STAMP(11 + 4 * w);
        tc5_fence_before();
        asm volatile("barrier.sync %0, 256;" ::"r"(1 + t) : "memory");   // the tile's 8 warps: operands complete, D consumed
        if (issuer_warp && w + 1 < Wn) { if (elect_one()) issue(); }
    }

    B5_STAMP(50);
    // ---- hidden state out; latent head z = softmax(W_o h + b_o); soft update (:118) ------------------------------------------
    float hv[16];
#pragma unroll
    for (int p = 0; p < 8; ++p) upk2(h2[p], hv[2 * p], hv[2 * p + 1]);
    if (ok) {
        float* hp = a.hid.ptr + ag * a.hid.stride_agent + bb * a.hid.stride_env + nn * a.hid.stride_slot + 16 * hh;
#pragma unroll
        for (int p = 0; p < 4; ++p) *reinterpret_cast<float4*>(hp + 4 * p) = make_float4(hv[4 * p], hv[4 * p + 1], hv[4 * p + 2], hv[4 * p + 3]);
    }
    float lg[B5_LAT];
#pragma unroll
    for (int l = 0; l < B5_LAT; ++l) {
        float s = 0.0f;
        const float4* ow4 = reinterpret_cast<const float4*>(s_ow + l * H + 16 * hh);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float4 w4 = ow4[p];
            s = fmaf(w4.x, hv[4 * p], s); s = fmaf(w4.y, hv[4 * p + 1], s); s = fmaf(w4.z, hv[4 * p + 2], s); s = fmaf(w4.w, hv[4 * p + 3], s);
        }
        lg[l] = s;
    }
    if (hh) {
#pragma unroll
        for (int l = 0; l < B5_LAT; ++l) s_part[c * B5_LAT + l] = lg[l];
    }
    asm volatile("barrier.sync %0, 256;" ::"r"(1 + t) : "memory");
    if (!hh && ok) {
        float mx = -INFINITY, den = 0.0f;
#pragma unroll
        for (int l = 0; l < B5_LAT; ++l) {
            lg[l] = l < Ld ? (lg[l] + s_part[c * B5_LAT + l]) + s_ob[l] : -INFINITY;
            mx = fmaxf(mx, lg[l]);
        }
#pragma unroll
        for (int l = 0; l < B5_LAT; ++l) { lg[l] = l < Ld ? expf(lg[l] - mx) : 0.0f; den += lg[l]; }
        const float* lp = a.lat_prev.ptr + ag * a.lat_prev.stride_agent + bb * a.lat_prev.stride_env + nn * a.lat_prev.stride_slot;
        float* lo_ = a.lat_out.ptr + ag * a.lat_out.stride_agent + bb * a.lat_out.stride_env + nn * a.lat_out.stride_slot;
#pragma unroll
        for (int l = 0; l < B5_LAT; ++l)        // (1 - c) * prev + z * c, each product rounded as numpy does (:118)
            if (l < Ld) lo_[l] = __fadd_rn(__fmul_rn(1.0f - a.coef, lp[l]), __fmul_rn(lg[l] / den, a.coef));
    }
    B5_STAMP(51);
    tc5_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc5_fence_after();
        tc5_dealloc<B5_TMEM_COLS>(tmem_base);
    }
    B5_STAMP(52);
}

bool behavior_tc5_supports(const BehArgs& a) {
    auto al4 = [](const iplan_view& v) { return (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0 && v.stride_agent % 4 == 0 && v.stride_env % 4 == 0 && v.stride_slot % 4 == 0; };
    return a.obs_dim <= B5_OBS && a.latent_dim <= B5_LAT && a.hist_len * a.obs_dim <= B5_WIN_MAX && a.hist_len <= 16 && al4(a.hid) &&
           (reinterpret_cast<uintptr_t>(a.params) & 15) == 0 && a.param_stride % 4 == 0;
}

int launch_behavior_tc5(const BehArgs& a, int n_agents, cudaStream_t st) {
    const B5Layout Y = b5_layout(a.hist_len * a.obs_dim);
    static uint32_t configured = 0;
    static int dbg = 0;
    if (Y.total > configured) {
        const char* ev = getenv("IPLAN_BEH_DBG");
        dbg = ev ? atoi(ev) : 0;
        cudaError_t e = cudaFuncSetAttribute(behavior_tc5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Y.total);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(behavior_tc5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Y.total);
        if (e != cudaSuccess) { set_error("behavior_step: tcgen05 kernel smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        configured = Y.total;
    }
    const int chains = a.n_envs * a.n_slots;
    const dim3 grid((chains + B5_CHAINS - 1) / B5_CHAINS, n_agents);
    if (dbg) behavior_tc5_kernel<true><<<grid, B5_THREADS, Y.total, st>>>(a);
    else behavior_tc5_kernel<false><<<grid, B5_THREADS, Y.total, st>>>(a);
    count_launch();
    return check_launch("behavior_step(tcgen05)");
}

}  // namespace iplan

// timing experiments (IPLAN_BEH_DBG=1): the phase-boundary clock stamps of one CTA of the last launch
extern "C" int iplan_behavior_debug_clocks(long long* out64) {
    const cudaError_t e = cudaMemcpyFromSymbol(out64, iplan::b5_clk, sizeof(long long) * 64);
    if (e != cudaSuccess) { iplan::set_error("behavior_debug_clocks: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
