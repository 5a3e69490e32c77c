// Host-buffer entry points of the rollout step: what a binding of the reference's numpy API calls
// (Prediction_policy.GAT_latent_update nova/prediction_policy.py:92-118, Behavior_policy.latent_update
// nova/stable_behavior_policy.py:83-123).  One synchronous call = the env dimension cut into chunks; chunk c's inputs go
// host -> device on a copy stream, its kernel runs on the caller's stream as soon as they have landed, its outputs go
// device -> host on a second copy stream while chunk c + 1 computes.  The whole pipeline is enqueued by this one native
// call (no interpreter between the chunks); it returns when the host output is complete, so the caller's
// numpy-in / numpy-out contract (and the reusability of its input buffers) is that of the reference.
// Host buffers may be pageable (cudaMemcpyAsync then stages them synchronously); page-locked ones overlap.
#include "common.cuh"

namespace iplan {
namespace {

constexpr int MAX_CHUNKS = 16;
struct Pipe {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t ev_main = nullptr, ev_in[MAX_CHUNKS] = {}, ev_k[MAX_CHUNKS] = {};
    int device = -1;
};
Pipe g_pipe;

int pipe_init() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("host_api: cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
    if (g_pipe.device == dev) return 0;
    IPLAN_REQUIRE(g_pipe.device < 0, "host_api: one device per process (pipeline built on device %d, called on %d)", g_pipe.device, dev);
    e = cudaStreamCreateWithFlags(&g_pipe.h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g_pipe.d2h, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_pipe.ev_main, cudaEventDisableTiming);
    for (int c = 0; c < MAX_CHUNKS && e == cudaSuccess; ++c) {
        e = cudaEventCreateWithFlags(&g_pipe.ev_in[c], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_pipe.ev_k[c], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { set_error("host_api: stream / event setup: %s", cudaGetErrorString(e)); return (int)e; }
    g_pipe.device = dev;
    return 0;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) { set_error("host_api: %s: %s", #call, cudaGetErrorString(e_)); return (int)e_; } \
    } while (0)

inline iplan_view view_of(const float* base, int64_t env0, int A, int N, int dim) {
    iplan_view v;
    v.ptr = const_cast<float*>(base) + env0 * (int64_t)A * N * dim;
    v.stride_agent = (int64_t)N * dim; v.stride_env = (int64_t)A * N * dim; v.stride_slot = dim;
    return v;
}

}  // namespace
}  // namespace iplan

extern "C" int iplan_gat_latent_update_host(const float* gat_params, int64_t param_stride,
                                            const float* host_hist, float* dev_hist, const float* host_h, float* dev_h,
                                            const float* host_beh, float* dev_beh, float* dev_out, float* host_out,
                                            uint64_t seed, uint64_t counter, float tau, float* scratch, int64_t scratch_floats,
                                            int n_envs, int n_agents, int n_slots, int obs_dim, int latent_dim, int n_chunks, const int32_t* chunk_end,
                                            void* stream) {
    using namespace iplan;
    IPLAN_REQUIRE(gat_params && dev_hist && dev_h && dev_beh && dev_out && host_out && scratch, "gat_latent_update_host: null pointer");
    IPLAN_REQUIRE(n_chunks >= 1 && n_chunks <= MAX_CHUNKS && n_envs > 0, "gat_latent_update_host: 1..%d chunks", MAX_CHUNKS);
    if (int rc = pipe_init()) return rc;
    cudaStream_t main = (cudaStream_t)stream;
    const int A = n_agents, N = n_slots, D = IPLAN_HID;
    CU(cudaEventRecord(g_pipe.ev_main, main));                       // the staging buffers may still be read by earlier launches
    CU(cudaStreamWaitEvent(g_pipe.h2d, g_pipe.ev_main, 0));
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t lo = chunk_end ? (c ? chunk_end[c - 1] : 0) : (int64_t)n_envs * c / n_chunks;
        const int64_t hi = chunk_end ? chunk_end[c] : (int64_t)n_envs * (c + 1) / n_chunks;
        IPLAN_REQUIRE(lo >= 0 && hi <= n_envs && (c + 1 < n_chunks || hi == n_envs), "gat_latent_update_host: piece %d = [%lld, %lld) of %d envs", c, (long long)lo, (long long)hi, n_envs);
        if (hi <= lo) continue;
        const int64_t rows = (hi - lo) * A * N, r0 = lo * A * N;
        bool copied = false;
        if (host_hist) { CU(cudaMemcpyAsync(dev_hist + r0 * obs_dim, host_hist + r0 * obs_dim, sizeof(float) * rows * obs_dim, cudaMemcpyHostToDevice, g_pipe.h2d)); copied = true; }
        if (host_h) { CU(cudaMemcpyAsync(dev_h + r0 * D, host_h + r0 * D, sizeof(float) * rows * D, cudaMemcpyHostToDevice, g_pipe.h2d)); copied = true; }
        if (host_beh) { CU(cudaMemcpyAsync(dev_beh + r0 * latent_dim, host_beh + r0 * latent_dim, sizeof(float) * rows * latent_dim, cudaMemcpyHostToDevice, g_pipe.h2d)); copied = true; }
        if (copied) {
            CU(cudaEventRecord(g_pipe.ev_in[c], g_pipe.h2d));
            CU(cudaStreamWaitEvent(main, g_pipe.ev_in[c], 0));
        }
        const int rc = iplan_gat_step_ex(gat_params, param_stride, view_of(dev_hist, lo, A, N, obs_dim), view_of(dev_beh, lo, A, N, latent_dim),
                                         view_of(dev_h, lo, A, N, D), view_of(dev_out, lo, A, N, D), nullptr, seed, counter + (uint64_t)c, tau, nullptr,
                                         scratch, scratch_floats, (int)(hi - lo), A, N, obs_dim, latent_dim, nullptr, nullptr, nullptr, stream);
        if (rc) return rc;
        CU(cudaEventRecord(g_pipe.ev_k[c], main));
        CU(cudaStreamWaitEvent(g_pipe.d2h, g_pipe.ev_k[c], 0));
        CU(cudaMemcpyAsync(host_out + r0 * D, dev_out + r0 * D, sizeof(float) * rows * D, cudaMemcpyDeviceToHost, g_pipe.d2h));
    }
    CU(cudaStreamSynchronize(g_pipe.d2h));
    return 0;
}

extern "C" int iplan_behavior_latent_update_host(const float* beh_params, int64_t param_stride,
                                                 const float* host_window, float* dev_window, const float* host_prev, float* dev_prev,
                                                 float* dev_hid_io, float* dev_new, float* host_new, float soft_coef,
                                                 int n_envs, int n_agents, int n_slots, int obs_dim, int latent_dim, int hist_len, int n_chunks, void* stream) {
    using namespace iplan;
    IPLAN_REQUIRE(beh_params && dev_window && dev_prev && dev_hid_io && dev_new && host_new, "behavior_latent_update_host: null pointer");
    IPLAN_REQUIRE(n_chunks >= 1 && n_chunks <= MAX_CHUNKS && n_envs > 0, "behavior_latent_update_host: 1..%d chunks", MAX_CHUNKS);
    if (int rc = pipe_init()) return rc;
    cudaStream_t main = (cudaStream_t)stream;
    const int A = n_agents, N = n_slots, E = IPLAN_HID, wo = hist_len * obs_dim;
    CU(cudaEventRecord(g_pipe.ev_main, main));
    CU(cudaStreamWaitEvent(g_pipe.h2d, g_pipe.ev_main, 0));
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t lo = (int64_t)n_envs * c / n_chunks, hi = (int64_t)n_envs * (c + 1) / n_chunks;
        if (hi <= lo) continue;
        const int64_t rows = (hi - lo) * A * N, r0 = lo * A * N;
        bool copied = false;
        if (host_window) { CU(cudaMemcpyAsync(dev_window + r0 * wo, host_window + r0 * wo, sizeof(float) * rows * wo, cudaMemcpyHostToDevice, g_pipe.h2d)); copied = true; }
        if (host_prev) { CU(cudaMemcpyAsync(dev_prev + r0 * latent_dim, host_prev + r0 * latent_dim, sizeof(float) * rows * latent_dim, cudaMemcpyHostToDevice, g_pipe.h2d)); copied = true; }
        if (copied) {
            CU(cudaEventRecord(g_pipe.ev_in[c], g_pipe.h2d));
            CU(cudaStreamWaitEvent(main, g_pipe.ev_in[c], 0));
        }
        const int rc = iplan_behavior_step_ex(beh_params, param_stride, view_of(dev_window, lo, A, N, wo), 0, 0, view_of(dev_hid_io, lo, A, N, E),
                                              view_of(dev_prev, lo, A, N, latent_dim), view_of(dev_new, lo, A, N, latent_dim), soft_coef,
                                              (int)(hi - lo), A, N, obs_dim, latent_dim, hist_len, stream);
        if (rc) return rc;
        CU(cudaEventRecord(g_pipe.ev_k[c], main));
        CU(cudaStreamWaitEvent(g_pipe.d2h, g_pipe.ev_k[c], 0));
        CU(cudaMemcpyAsync(host_new + r0 * latent_dim, dev_new + r0 * latent_dim, sizeof(float) * rows * latent_dim, cudaMemcpyDeviceToHost, g_pipe.d2h));
    }
    CU(cudaStreamSynchronize(g_pipe.d2h));
    return 0;
}

// n device -> host copies on `stream`, then ONE synchronize (select_actions_ippo returns four small arrays)
extern "C" int iplan_d2h_batch(void* const* dst_host, const void* const* src_dev, const int64_t* bytes, int n, void* stream) {
    using namespace iplan;
    IPLAN_REQUIRE(dst_host && src_dev && bytes && n >= 0, "d2h_batch: null pointer");
    for (int i = 0; i < n; ++i) CU(cudaMemcpyAsync(dst_host[i], src_dev[i], (size_t)bytes[i], cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}
