// K1 — fused GAT step for one timestep of the rollout.
//
// Replaces Prediction_policy.GAT_latent_update (reference nova/prediction_policy.py:92-118)
// and, inside it, GAT_Net.forward (nova/GAT_Net.py:41-142).  One CTA owns one
// (env b, agent-net a) pair and keeps every intermediate in shared memory:
//
//   x[n]      = [history_t[n] | behaviour_{t-1}[n]]                       (:102-105)
//   enc[n]    = ReLU(W_e x[n] + b_e)                                       (:50)
//   hard attention: for ego i, a bidirectional GRU runs over the N-1 neighbours
//     j(i,s) = s < i ? s : s+1 with input [enc_i ; enc_j], h0 = 0          (:57-83)
//     -> factored input projection  W_ih[enc_i;enc_j] = P[i] + Q[j]  
//     -> logits(i,s) = W_he [h_fwd ; h_rev] + b_he, gumbel-softmax tau      (:85-95)
//        only softmax(.)[1] is used, i.e. sigmoid((l1-l0 + g1-g0)/tau)
//   soft attention: softmax_s(q_i . k_j / sqrt(D)), v = ReLU(W_v enc + b)   (:99-129)
//   x_i = sum_s soft * hard * v_j (no renormalisation)                      (:132)
//   out = GRUCell(x_i, h_prev_i)                                            (:140)
//
// HBM traffic per (b,a): read N*(o+L+32) floats, write N*32 floats; weights (28.8k
// floats) come from L2.  The work is the 2*N sequential GRU chains of length N-1.
#include "common.cuh"

namespace iplan {

constexpr int H = IPLAN_HID;   // 32 == GAT_hidden_dim == attention_dim
constexpr int G3 = 3 * H;      // gate rows r|z|n
constexpr int GAT_THREADS = 256;
constexpr int GAT_WARPS = GAT_THREADS / 32;
constexpr int SEQ_G = 7;       // GRU chains a warp advances together (register blocking)
constexpr int IN_MAX = 16;     // obs_dim + latent_dim upper bound
constexpr int WT_LD = 97;      // padded leading dim of the transposed GRUCell weights

struct GatArgs {
    const float* params; int64_t param_stride;
    iplan_view hist, beh, hprev, out;
    const float* gumbel; float* dbg_hard;
    uint64_t seed, counter;
    float inv_tau;
    int n_envs, n_slots, obs_dim, latent_dim;
};

__host__ __device__ inline size_t gat_smem_floats(int N) {
    return (size_t)N * IN_MAX + (size_t)N * H + 4 * (size_t)N * G3 + 2 * (size_t)N * (N - 1) +
           GAT_WARPS * SEQ_G * H + 2 * H * WT_LD + GAT_WARPS * 64;
}

__global__ void __launch_bounds__(GAT_THREADS, 1) gat_step_kernel(GatArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int b = blockIdx.x, ag = blockIdx.y;
    const int N = a.n_slots, NM1 = N - 1;
    const int in_dim = a.obs_dim + a.latent_dim;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ W = a.params + (int64_t)ag * a.param_stride;
    const GatLayout L = gat_layout(in_dim);

    float* s_x = smem;                                  // [N][IN_MAX]
    float* s_enc = s_x + N * IN_MAX;                    // [N][H]
    float* s_P = s_enc + N * H;                         // [2][N][96]  ego part + b_ih
    float* s_Q = s_P + 2 * N * G3;                      // [2][N][96]  neighbour part
    float* s_dl = s_Q + 2 * N * G3;                     // [2][N][N-1] per-direction logit diff
    float* s_hb = s_dl + 2 * N * NM1;                   // [warps][SEQ_G][H] chain states
    float* s_wt = s_hb + GAT_WARPS * SEQ_G * H;         // [2][H][WT_LD] GRUCell W^T
    float* s_w = s_wt + 2 * H * WT_LD;                  // [warps][64] attention weights
    // after the recurrence the P/Q region is reused:
    float* s_q = s_P;                                   // [N][33]
    float* s_k = s_q + N * 33;                          // [N][33]
    float* s_v = s_k + N * 33;                          // [N][H]
    float* s_xa = s_v + N * H;                          // [N][H] aggregated messages

    const float* hist = a.hist.ptr + ag * a.hist.stride_agent + b * a.hist.stride_env;
    const float* beh = a.beh.ptr + ag * a.beh.stride_agent + b * a.beh.stride_env;
    const float* hprev = a.hprev.ptr + ag * a.hprev.stride_agent + b * a.hprev.stride_env;
    float* outp = a.out.ptr + ag * a.out.stride_agent + b * a.out.stride_env;

    // ---- phase 0: gather x = [history | behaviour latent] ---------------------------
    for (int idx = tid; idx < N * in_dim; idx += GAT_THREADS) {
        const int n = idx / in_dim, c = idx - n * in_dim;
        float v = (c < a.obs_dim) ? hist[n * a.hist.stride_slot + c]
                                  : beh[n * a.beh.stride_slot + (c - a.obs_dim)];
        s_x[n * IN_MAX + c] = v;
    }
    __syncthreads();

    // ---- phase 1: enc = ReLU(W_e x + b_e) ------------------------------------------
    for (int idx = tid; idx < N * H; idx += GAT_THREADS) {
        const int n = idx >> 5, c = idx & 31;
        float acc = W[L.enc_b + c];
        const float* w = W + L.enc_w + c * in_dim;
        for (int k = 0; k < in_dim; ++k) acc = fmaf(w[k], s_x[n * IN_MAX + k], acc);
        s_enc[idx] = fmaxf(acc, 0.0f);
    }
    __syncthreads();

    // ---- phase 2: factored input projections P (ego, + b_ih) and Q (neighbour) ------
    for (int col = tid; col < 4 * G3; col += GAT_THREADS) {
        const int d = col / (2 * G3);
        const int part = (col - d * 2 * G3) / G3;
        const int g = col % G3;
        const float* wih = W + (d ? L.wih_r : L.wih_f) + g * (2 * H) + part * H;
        float w[H];
#pragma unroll
        for (int k = 0; k < H; ++k) w[k] = wih[k];
        const float bias = part == 0 ? W[(d ? L.bih_r : L.bih_f) + g] : 0.0f;
        float* dst = (part == 0 ? s_P : s_Q) + (size_t)d * N * G3 + g;
        for (int n = 0; n < N; ++n) {
            float acc = bias;
            const float4* e4 = reinterpret_cast<const float4*>(s_enc + n * H);
#pragma unroll
            for (int kk = 0; kk < H / 4; ++kk) {
                const float4 e = e4[kk];
                acc = fmaf(w[4 * kk + 0], e.x, acc);
                acc = fmaf(w[4 * kk + 1], e.y, acc);
                acc = fmaf(w[4 * kk + 2], e.z, acc);
                acc = fmaf(w[4 * kk + 3], e.w, acc);
            }
            dst[n * G3] = acc;
        }
    }
    __syncthreads();

    // ---- phase 3: the 2N GRU chains -------------------------------------------------
    {
        const int d = warp >> 2;          // warps 0-3 forward, 4-7 reverse
        const int wi = warp & 3;
        const int count = (N - wi + 3) >> 2;      // egos wi, wi+4, ...
        const float* whh = W + (d ? L.whh_r : L.whh_f);
        const float* bhh = W + (d ? L.bhh_r : L.bhh_f);
        float w_r[H], w_z[H], w_n[H];
#pragma unroll
        for (int k = 0; k < H; ++k) {
            w_r[k] = whh[(lane) * H + k];
            w_z[k] = whh[(H + lane) * H + k];
            w_n[k] = whh[(2 * H + lane) * H + k];
        }
        const float b_r = bhh[lane], b_z = bhh[H + lane], b_n = bhh[2 * H + lane];
        const float wd = W[L.he_w + 2 * H + d * H + lane] - W[L.he_w + d * H + lane];
        const float* Pd = s_P + (size_t)d * N * G3;
        const float* Qd = s_Q + (size_t)d * N * G3;
        float* dl = s_dl + (size_t)d * N * NM1;
        float* hb = s_hb + warp * SEQ_G * H;

        for (int base = 0; base < count; base += SEQ_G) {
            int ego[SEQ_G];
            float pr[SEQ_G], pz[SEQ_G], pn[SEQ_G], h[SEQ_G];
#pragma unroll
            for (int g = 0; g < SEQ_G; ++g) {
                const bool valid = base + g < count;
                ego[g] = valid ? wi + 4 * (base + g) : -1;
                const int i = valid ? ego[g] : 0;
                pr[g] = Pd[i * G3 + lane];
                pz[g] = Pd[i * G3 + H + lane];
                pn[g] = Pd[i * G3 + 2 * H + lane];
                h[g] = 0.0f;
                hb[g * H + lane] = 0.0f;
            }
            __syncwarp();
            for (int step = 0; step < NM1; ++step) {
                const int s = d ? NM1 - 1 - step : step;
                float ar[SEQ_G], az[SEQ_G], an[SEQ_G];
#pragma unroll
                for (int g = 0; g < SEQ_G; ++g) { ar[g] = b_r; az[g] = b_z; an[g] = b_n; }
#pragma unroll
                for (int kk = 0; kk < H / 4; ++kk) {
#pragma unroll
                    for (int g = 0; g < SEQ_G; ++g) {
                        const float4 hv = *reinterpret_cast<const float4*>(hb + g * H + 4 * kk);
                        ar[g] = fmaf(w_r[4 * kk + 0], hv.x, ar[g]);
                        az[g] = fmaf(w_z[4 * kk + 0], hv.x, az[g]);
                        an[g] = fmaf(w_n[4 * kk + 0], hv.x, an[g]);
                        ar[g] = fmaf(w_r[4 * kk + 1], hv.y, ar[g]);
                        az[g] = fmaf(w_z[4 * kk + 1], hv.y, az[g]);
                        an[g] = fmaf(w_n[4 * kk + 1], hv.y, an[g]);
                        ar[g] = fmaf(w_r[4 * kk + 2], hv.z, ar[g]);
                        az[g] = fmaf(w_z[4 * kk + 2], hv.z, az[g]);
                        an[g] = fmaf(w_n[4 * kk + 2], hv.z, an[g]);
                        ar[g] = fmaf(w_r[4 * kk + 3], hv.w, ar[g]);
                        az[g] = fmaf(w_z[4 * kk + 3], hv.w, az[g]);
                        an[g] = fmaf(w_n[4 * kk + 3], hv.w, an[g]);
                    }
                }
                __syncwarp();   // every lane has consumed the old states
#pragma unroll
                for (int g = 0; g < SEQ_G; ++g) {
                    if (ego[g] >= 0) {       // warp-uniform
                        const int i = ego[g];
                        const int j = s < i ? s : s + 1;
                        const float* qj = Qd + j * G3;
                        const float r = sigmoidf_acc(pr[g] + qj[lane] + ar[g]);
                        const float z = sigmoidf_acc(pz[g] + qj[H + lane] + az[g]);
                        const float n = tanhf_acc(pn[g] + qj[2 * H + lane] + r * an[g]);
                        h[g] = (1.0f - z) * n + z * h[g];
                        hb[g * H + lane] = h[g];
                        const float part = warp_sum(wd * h[g]);
                        if (lane == 0) dl[i * NM1 + s] = part;
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncthreads();

    // ---- phase 4: q, k, v and the transposed GRUCell weights -------------------------
    for (int idx = tid; idx < N * H; idx += GAT_THREADS) {
        const int n = idx >> 5, c = idx & 31;
        const float* e = s_enc + n * H;
        const float* wq = W + L.q_w + c * H;
        const float* wk = W + L.k_w + c * H;
        const float* wv = W + L.v_w + c * H;
        float aq = 0.0f, ak = 0.0f, av = W[L.v_b + c];
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
            const float ek = e[k];
            aq = fmaf(wq[k], ek, aq);
            ak = fmaf(wk[k], ek, ak);
            av = fmaf(wv[k], ek, av);
        }
        s_q[n * 33 + c] = aq;
        s_k[n * 33 + c] = ak;
        s_v[n * H + c] = fmaxf(av, 0.0f);
    }
    for (int idx = tid; idx < 2 * G3 * H; idx += GAT_THREADS) {
        const int m = idx / (G3 * H);
        const int r = idx - m * G3 * H;
        const int g = r >> 5, k = r & 31;
        s_wt[(m * H + k) * WT_LD + g] = W[(m ? L.c_whh : L.c_wih) + g * H + k];
    }
    __syncthreads();

    // ---- phase 5: soft x hard attention + GRUCell, one warp per ego ------------------
    const float db = W[L.he_b + 1] - W[L.he_b + 0];
    float* wbuf = s_w + warp * 64;
    float* hpb = s_hb + warp * SEQ_G * H;     // reuse: broadcast buffer for h_prev
    for (int i = warp; i < N; i += GAT_WARPS) {
        float sc[2], hd[2];
        int jj[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int s = lane + 32 * u;
            sc[u] = -INFINITY; hd[u] = 0.0f; jj[u] = 0;
            if (s < NM1) {
                const int j = s < i ? s : s + 1;
                jj[u] = j;
                float acc = 0.0f;
#pragma unroll 8
                for (int k = 0; k < H; ++k) acc = fmaf(s_q[i * 33 + k], s_k[j * 33 + k], acc);
                sc[u] = acc / 5.656854249492381f;            // np.sqrt(attention_dim), :126
                float noise;
                const int64_t edge = (((int64_t)ag * a.n_envs + b) * N + i) * NM1 + s;
                if (a.gumbel) {
                    noise = a.gumbel[2 * edge + 1] - a.gumbel[2 * edge];
                } else {
                    const uint4 rnd = philox4x32(
                        make_uint4((uint32_t)edge, (uint32_t)(edge >> 32), (uint32_t)a.counter,
                                   (uint32_t)(a.counter >> 32)),
                        make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
                    const float uu = u01(rnd.x);
                    noise = logf(uu) - log1pf(-uu);          // Gumbel - Gumbel ~ Logistic(0,1)
                }
                const float dlog = s_dl[i * NM1 + s] + s_dl[(size_t)N * NM1 + i * NM1 + s] + db;
                hd[u] = sigmoidf_acc((dlog + noise) * a.inv_tau);
                if (a.dbg_hard) a.dbg_hard[edge] = hd[u];
            }
        }
        const float mx = warp_max(fmaxf(sc[0], sc[1]));
        const float e0 = (lane < NM1) ? expf(sc[0] - mx) : 0.0f;
        const float e1 = (lane + 32 < NM1) ? expf(sc[1] - mx) : 0.0f;
        const float den = warp_sum(e0 + e1);
        wbuf[lane] = (e0 / den) * hd[0];
        wbuf[lane + 32] = (e1 / den) * hd[1];
        hpb[lane] = hprev[i * a.hprev.stride_slot + lane];
        __syncwarp();
        float xa = 0.0f;
        for (int s = 0; s < NM1; ++s) {
            const int j = s < i ? s : s + 1;
            xa = fmaf(wbuf[s], s_v[j * H + lane], xa);
        }
        s_xa[i * H + lane] = xa;
        __syncwarp();
        // GRUCell(x_i, h_prev_i): lane c owns gate rows c, H+c, 2H+c
        float gi0 = W[L.c_bih + lane], gi1 = W[L.c_bih + H + lane], gi2 = W[L.c_bih + 2 * H + lane];
        float gh0 = W[L.c_bhh + lane], gh1 = W[L.c_bhh + H + lane], gh2 = W[L.c_bhh + 2 * H + lane];
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
            const float xk = s_xa[i * H + k];
            const float hk = hpb[k];
            const float* wi = s_wt + k * WT_LD;
            const float* wh = s_wt + (H + k) * WT_LD;
            gi0 = fmaf(wi[lane], xk, gi0);
            gi1 = fmaf(wi[H + lane], xk, gi1);
            gi2 = fmaf(wi[2 * H + lane], xk, gi2);
            gh0 = fmaf(wh[lane], hk, gh0);
            gh1 = fmaf(wh[H + lane], hk, gh1);
            gh2 = fmaf(wh[2 * H + lane], hk, gh2);
        }
        const float r = sigmoidf_acc(gi0 + gh0);
        const float z = sigmoidf_acc(gi1 + gh1);
        const float n = tanhf_acc(gi2 + r * gh2);
        outp[i * a.out.stride_slot + lane] = (1.0f - z) * n + z * hpb[lane];
        __syncwarp();
    }
}

}  // namespace iplan

extern "C" int iplan_gat_step(const float* gat_params, int64_t param_stride,
                              iplan_view hist, iplan_view beh_prev, iplan_view h_prev, iplan_view out,
                              const float* gumbel, uint64_t seed, uint64_t counter,
                              float tau, float* dbg_hard,
                              int n_envs, int n_agents, int n_slots, int obs_dim, int latent_dim,
                              void* stream) {
    using namespace iplan;
    IPLAN_REQUIRE(n_slots >= 2 && n_slots <= IPLAN_MAX_SLOTS, "gat_step: n_slots %d not in [2,%d]", n_slots, IPLAN_MAX_SLOTS);
    IPLAN_REQUIRE(obs_dim + latent_dim <= IN_MAX && obs_dim > 0 && latent_dim >= 0, "gat_step: obs_dim+latent_dim %d > %d", obs_dim + latent_dim, IN_MAX);
    IPLAN_REQUIRE(n_envs > 0 && n_agents > 0 && n_agents <= 65535, "gat_step: bad n_envs/n_agents");
    IPLAN_REQUIRE(gat_params && hist.ptr && beh_prev.ptr && h_prev.ptr && out.ptr, "gat_step: null pointer");
    IPLAN_REQUIRE(tau > 0.f, "gat_step: tau must be > 0");
    GatArgs a;
    a.params = gat_params; a.param_stride = param_stride;
    a.hist = hist; a.beh = beh_prev; a.hprev = h_prev; a.out = out;
    a.gumbel = gumbel; a.dbg_hard = dbg_hard; a.seed = seed; a.counter = counter;
    a.inv_tau = 1.0f / tau;
    a.n_envs = n_envs; a.n_slots = n_slots; a.obs_dim = obs_dim; a.latent_dim = latent_dim;
    const size_t smem = gat_smem_floats(n_slots) * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(gat_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("gat_step: smem attr %zu: %s", smem, cudaGetErrorString(e)); return (int)e; }
        configured = smem;
    }
    dim3 grid(n_envs, n_agents);
    gat_step_kernel<<<grid, GAT_THREADS, smem, (cudaStream_t)stream>>>(a);
    count_launch();
    return check_launch("gat_step");
}
