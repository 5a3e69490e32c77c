// K1b — behaviour-incentive encoder step.
//
// Replaces Behavior_policy.latent_update (reference nova/stable_behavior_policy.py:83-123)
// and EncoderRNN.forward (nova/behavior_net.py:17-22):
//   u_t   = ReLU(W_l w_t + b_l)                       for the W window rows w_t of a slot
//   h     = GRU(u_1..u_W ; h_carried)                 (one layer, batch_first)
//   z     = softmax(W_o h + b_o)                      latent_dim values
//   new   = (1 - c) * prev + c * z                    soft update (:118), c = soft_update_coef
// A "node" is one (env, agent-net, slot); nodes are independent.  One warp advances
// NODE_G nodes together: lane c owns hidden unit c, W_hh rows live in registers, W_ih^T
// is staged once per CTA in shared memory.
#include "common.cuh"

namespace iplan {

constexpr int E = IPLAN_HID;       // encoder_rnn_dim
constexpr int E3 = 3 * E;
constexpr int BEH_THREADS = 256;
constexpr int BEH_WARPS = BEH_THREADS / 32;
constexpr int NODE_G = 8;          // nodes advanced together by one warp
constexpr int BEH_WT_LD = 97;
constexpr int WIN_MAX = 64;        // hist_len * obs_dim upper bound per node
constexpr int LAT_MAX = 16;

struct BehArgs {
    const float* params; int64_t param_stride;
    iplan_view window, hid, lat_prev, lat_out;
    float coef;
    int n_envs, n_slots, obs_dim, latent_dim, hist_len;
};

struct BehSmem {
    float wih_t[E * BEH_WT_LD];                 // W_ih^T  [k][g]
    float win[BEH_WARPS][NODE_G][WIN_MAX];      // staged windows
    float u[BEH_WARPS][NODE_G][E];              // u_t broadcast buffer
    float hb[BEH_WARPS][NODE_G][E];             // hidden broadcast buffer
};

__global__ void __launch_bounds__(BEH_THREADS, 1) behavior_step_kernel(BehArgs a) {
    extern __shared__ __align__(16) unsigned char raw[];
    BehSmem& S = *reinterpret_cast<BehSmem*>(raw);
    const int ag = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.n_slots, o = a.obs_dim, Wn = a.hist_len, Ld = a.latent_dim;
    const float* __restrict__ P = a.params + (int64_t)ag * a.param_stride;
    const BehLayout L = beh_layout(o, Ld);
    const int total_nodes = a.n_envs * N;

    for (int idx = tid; idx < E3 * E; idx += BEH_THREADS) {
        const int g = idx >> 5, k = idx & 31;
        S.wih_t[k * BEH_WT_LD + g] = P[L.wih + g * E + k];
    }
    float w_r[E], w_z[E], w_n[E];
#pragma unroll
    for (int k = 0; k < E; ++k) {
        w_r[k] = P[L.whh + (lane) * E + k];
        w_z[k] = P[L.whh + (E + lane) * E + k];
        w_n[k] = P[L.whh + (2 * E + lane) * E + k];
    }
    const float bi_r = P[L.bih + lane], bi_z = P[L.bih + E + lane], bi_n = P[L.bih + 2 * E + lane];
    const float bh_r = P[L.bhh + lane], bh_z = P[L.bhh + E + lane], bh_n = P[L.bhh + 2 * E + lane];
    float wl[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) wl[q] = q < o ? P[L.lin_w + lane * o + q] : 0.0f;
    const float bl = P[L.lin_b + lane];
    __syncthreads();

    const int node0 = (blockIdx.x * BEH_WARPS + warp) * NODE_G;
    if (node0 >= total_nodes) return;

    float h[NODE_G];
    int64_t off_win[NODE_G], off_hid[NODE_G];
    bool valid[NODE_G];
#pragma unroll
    for (int g = 0; g < NODE_G; ++g) {
        const int node = node0 + g;
        valid[g] = node < total_nodes;
        const int nd = valid[g] ? node : node0;
        const int b = nd / N, n = nd - b * N;
        off_win[g] = ag * a.window.stride_agent + b * a.window.stride_env + n * a.window.stride_slot;
        off_hid[g] = ag * a.hid.stride_agent + b * a.hid.stride_env + n * a.hid.stride_slot;
        h[g] = a.hid.ptr[off_hid[g] + lane];
        S.hb[warp][g][lane] = h[g];
        for (int q = lane; q < Wn * o; q += 32) S.win[warp][g][q] = a.window.ptr[off_win[g] + q];
    }
    __syncwarp();

    for (int t = 0; t < Wn; ++t) {
#pragma unroll
        for (int g = 0; g < NODE_G; ++g) {
            float acc = bl;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < o) acc = fmaf(wl[q], S.win[warp][g][t * o + q], acc);
            S.u[warp][g][lane] = fmaxf(acc, 0.0f);
        }
        __syncwarp();
        float ar[NODE_G], az[NODE_G], ain[NODE_G], ahn[NODE_G];
#pragma unroll
        for (int g = 0; g < NODE_G; ++g) { ar[g] = bi_r + bh_r; az[g] = bi_z + bh_z; ain[g] = bi_n; ahn[g] = bh_n; }
#pragma unroll
        for (int kk = 0; kk < E / 4; ++kk) {
            float wi_r[4], wi_z[4], wi_n[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* row = S.wih_t + (4 * kk + q) * BEH_WT_LD;
                wi_r[q] = row[lane]; wi_z[q] = row[E + lane]; wi_n[q] = row[2 * E + lane];
            }
#pragma unroll
            for (int g = 0; g < NODE_G; ++g) {
                const float4 uv = *reinterpret_cast<const float4*>(&S.u[warp][g][4 * kk]);
                const float4 hv = *reinterpret_cast<const float4*>(&S.hb[warp][g][4 * kk]);
                const float ux[4] = {uv.x, uv.y, uv.z, uv.w};
                const float hx[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    ar[g] = fmaf(wi_r[q], ux[q], ar[g]);
                    az[g] = fmaf(wi_z[q], ux[q], az[g]);
                    ain[g] = fmaf(wi_n[q], ux[q], ain[g]);
                    ar[g] = fmaf(w_r[4 * kk + q], hx[q], ar[g]);
                    az[g] = fmaf(w_z[4 * kk + q], hx[q], az[g]);
                    ahn[g] = fmaf(w_n[4 * kk + q], hx[q], ahn[g]);
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int g = 0; g < NODE_G; ++g) {
            const float r = sigmoidf_acc(ar[g]);
            const float z = sigmoidf_acc(az[g]);
            const float n = tanhf_acc(ain[g] + r * ahn[g]);
            h[g] = (1.0f - z) * n + z * h[g];
            S.hb[warp][g][lane] = h[g];
        }
        __syncwarp();
    }

    // latent = softmax(W_o h + b_o); soft update; store hidden
#pragma unroll
    for (int g = 0; g < NODE_G; ++g) {
        if (!valid[g]) continue;      // warp-uniform
        a.hid.ptr[off_hid[g] + lane] = h[g];
        float logit = -INFINITY;
        if (lane < Ld) {
            float acc = P[L.out_b + lane];
            const float* wo = P + L.out_w + lane * E;
#pragma unroll 8
            for (int k = 0; k < E; ++k) acc = fmaf(wo[k], S.hb[warp][g][k], acc);
            logit = acc;
        }
        const float mx = warp_max(logit);
        const float ex = lane < Ld ? expf(logit - mx) : 0.0f;
        const float den = warp_sum(ex);
        if (lane < Ld) {
            const int node = node0 + g;
            const int b = node / N, n = node - b * N;
            const float prev = a.lat_prev.ptr[ag * a.lat_prev.stride_agent + b * a.lat_prev.stride_env +
                                              n * a.lat_prev.stride_slot + lane];
            // (1 - c) * prev + z * c, each product rounded as numpy does (:118)
            const float val = __fadd_rn(__fmul_rn(1.0f - a.coef, prev), __fmul_rn(ex / den, a.coef));
            a.lat_out.ptr[ag * a.lat_out.stride_agent + b * a.lat_out.stride_env + n * a.lat_out.stride_slot + lane] = val;
        }
    }
}

}  // namespace iplan

extern "C" int iplan_behavior_step(const float* beh_params, int64_t param_stride,
                                   iplan_view window, iplan_view hid_io, iplan_view lat_prev, iplan_view lat_out,
                                   float soft_coef,
                                   int n_envs, int n_agents, int n_slots, int obs_dim, int latent_dim, int hist_len,
                                   void* stream) {
    using namespace iplan;
    IPLAN_REQUIRE(obs_dim > 0 && obs_dim <= 8, "behavior_step: obs_dim %d not in [1,8]", obs_dim);
    IPLAN_REQUIRE(hist_len > 0 && hist_len * obs_dim <= WIN_MAX, "behavior_step: hist_len*obs_dim %d > %d", hist_len * obs_dim, WIN_MAX);
    IPLAN_REQUIRE(latent_dim > 0 && latent_dim <= LAT_MAX, "behavior_step: latent_dim %d not in [1,%d]", latent_dim, LAT_MAX);
    IPLAN_REQUIRE(n_envs > 0 && n_agents > 0 && n_agents <= 65535 && n_slots > 0, "behavior_step: bad sizes");
    IPLAN_REQUIRE(beh_params && window.ptr && hid_io.ptr && lat_prev.ptr && lat_out.ptr, "behavior_step: null pointer");
    BehArgs a;
    a.params = beh_params; a.param_stride = param_stride;
    a.window = window; a.hid = hid_io; a.lat_prev = lat_prev; a.lat_out = lat_out;
    a.coef = soft_coef;
    a.n_envs = n_envs; a.n_slots = n_slots; a.obs_dim = obs_dim; a.latent_dim = latent_dim; a.hist_len = hist_len;
    const size_t smem = sizeof(BehSmem);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(behavior_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("behavior_step: smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    const int nodes = n_envs * n_slots;
    dim3 grid((nodes + BEH_WARPS * NODE_G - 1) / (BEH_WARPS * NODE_G), n_agents);
    behavior_step_kernel<<<grid, BEH_THREADS, smem, (cudaStream_t)stream>>>(a);
    count_launch();
    return check_launch("behavior_step");
}
