// Observation-history wrapper step: the per-(env, agent) bookkeeping of
// observersation_state_history_wrapper.obs_history_create / obs_history_output / obs_single_history_output
// (reference observation_wrapper.py:68-141) as one kernel, so a vectorised simulator's raw observations can feed
// K1 / K1b without a host round trip.
//
// Reference semantics restated: every agent keeps a list of the vehicle ids it has seen, in first-seen order; the
// position in that list is the vehicle's SLOT for the rest of the episode (:82-88).  Each timestep every known slot
// gets one new history row: the observation (without its id column, :90) if the vehicle is observed now, zeros
// otherwise (:92-96).  The outputs are the last `hist_len` rows of every slot, right aligned (:101-119), and the
// newest row (:124-141).  Appending a zero row to a slot that does not exist yet changes nothing (its window is
// zero), so the update is uniform over the N slots: shift the window left by one row, append the new row.
//
// One CTA per (env, agent).  Slot assignment is done by warp 0, 32 observed rows at a time: every lane looks its id up
// in the agent's slot table; rows with an unknown id find the first row of the same id among themselves
// (__match_any_sync) and the leaders take consecutive new slots in row order (ballot + popc) = first-seen order.  Of
// several rows with the same id the LAST one provides the values (a deque's last append), decided by the same match
// mask, so there is no write race.  The whole CTA then shifts the N x W x o window in place.
#include "common.cuh"

namespace iplan {

constexpr int OBS_THREADS = 128;
constexpr int OBS_MAX_ROWS = 64;          // rows of one raw observation (n_obs_vehicles)
constexpr int OBS_MAX_REG = 32;           // window elements per thread held across the in-place shift

__global__ void __launch_bounds__(OBS_THREADS) obs_history_step_kernel(
    const float* __restrict__ obs, int n_obs, int obs_dim, int32_t* __restrict__ slot_ids, int32_t* __restrict__ slot_count,
    float* __restrict__ window, float* __restrict__ single, int32_t* __restrict__ overflow, int N, int W) {
    extern __shared__ float s_new[];                   // [N][o] the rows appended this step
    const int ka = blockIdx.x, tid = threadIdx.x;
    const int o = obs_dim - 1;
    const float* ob = obs + (int64_t)ka * n_obs * obs_dim;
    int32_t* ids = slot_ids + (int64_t)ka * N;

    for (int idx = tid; idx < N * o; idx += OBS_THREADS) s_new[idx] = 0.0f;
    __syncthreads();
    if (tid < 32) {
        const int lane = tid;
        int count = slot_count[ka];                                    // warp-uniform
        for (int j0 = 0; j0 < n_obs; j0 += 32) {                       // chunks in row order: later rows see earlier chunks' slots
            const int j = j0 + lane;
            bool any = false;                                          // np.any(obs[k, i, j, :]) — the id column included (:80)
            int id = 0;
            if (j < n_obs) {
                for (int c = 0; c < obs_dim; ++c) any |= ob[j * obs_dim + c] != 0.0f;
                id = (int)ob[j * obs_dim];                             // int(...) truncation, as :81
            }
            int slot = -1;
            if (any)
                for (int q = 0; q < count; ++q) if (ids[q] == id) { slot = q; break; }
            // rows whose id is not in the table yet: group by id, the first row of a group opens the slot
            const bool fresh = any && slot < 0;
            const unsigned fresh_mask = __ballot_sync(0xffffffffu, fresh);
            unsigned same = 0u;
            if (fresh) same = __match_any_sync(fresh_mask, id);
            const bool leader = fresh && (__ffs(same) - 1) == lane;
            const unsigned leaders = __ballot_sync(0xffffffffu, leader);
            int new_slot = -1;
            if (leader) {
                new_slot = count + __popc(leaders & ((1u << lane) - 1u));
                if (new_slot < N) ids[new_slot] = id;
                else { new_slot = -1; atomicExch(overflow, 1); }       // the reference would raise IndexError at :116
            }
            const int lead_lane = fresh ? __ffs(same) - 1 : 0;
            const int got = __shfl_sync(0xffffffffu, new_slot, lead_lane);
            if (fresh) slot = got;
            count = min(N, count + __popc(leaders));
            // values: of the rows that share a slot in this chunk the last one writes (later chunks overwrite in order)
            const unsigned valid = __ballot_sync(0xffffffffu, slot >= 0);
            unsigned grp = 0u;
            if (slot >= 0) grp = __match_any_sync(valid, slot);
            if (slot >= 0 && (31 - __clz(grp)) == lane)
                for (int c = 0; c < o; ++c) s_new[slot * o + c] = ob[j * obs_dim + 1 + c];
            __syncwarp();
        }
        if (lane == 0) slot_count[ka] = count;
    }
    __syncthreads();
    // shift left by one row and append: element (slot, w, c) <- (slot, w + 1, c); read everything, barrier, write
    float* win = window + (int64_t)ka * N * W * o;
    float* sg = single + (int64_t)ka * N * o;
    const int total = N * W * o;
    float keep[OBS_MAX_REG];
#pragma unroll
    for (int q = 0; q < OBS_MAX_REG; ++q) {
        const int idx = tid + q * OBS_THREADS;
        if (idx < total) {
            const int c = idx % o, w = (idx / o) % W, slot = idx / (o * W);
            keep[q] = w + 1 < W ? win[idx + o] : s_new[slot * o + c];
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < OBS_MAX_REG; ++q) {
        const int idx = tid + q * OBS_THREADS;
        if (idx < total) win[idx] = keep[q];
    }
    for (int idx = tid; idx < N * o; idx += OBS_THREADS) sg[idx] = s_new[idx];
}

}  // namespace iplan

extern "C" int iplan_obs_history_step(const float* obs, int n_envs, int n_agents, int n_obs, int obs_dim,
                                      int32_t* slot_ids, int32_t* slot_count, float* window, float* single,
                                      int32_t* overflow, int n_slots, int hist_len, void* stream) {
    using namespace iplan;
    IPLAN_REQUIRE(obs && slot_ids && slot_count && window && single && overflow, "obs_history_step: null pointer");
    IPLAN_REQUIRE(n_envs > 0 && n_agents > 0 && obs_dim >= 2, "obs_history_step: bad sizes");
    IPLAN_REQUIRE(n_obs > 0 && n_obs <= OBS_MAX_ROWS, "obs_history_step: n_obs %d not in [1,%d]", n_obs, OBS_MAX_ROWS);
    IPLAN_REQUIRE(n_slots > 0 && hist_len > 0 && (int64_t)n_slots * hist_len * (obs_dim - 1) <= (int64_t)OBS_MAX_REG * OBS_THREADS,
                  "obs_history_step: window of %d x %d x %d exceeds %d elements", n_slots, hist_len, obs_dim - 1, OBS_MAX_REG * OBS_THREADS);
    const size_t smem = (size_t)n_slots * (obs_dim - 1) * sizeof(float);
    obs_history_step_kernel<<<n_envs * n_agents, OBS_THREADS, smem, (cudaStream_t)stream>>>(
        obs, n_obs, obs_dim, slot_ids, slot_count, window, single, overflow, n_slots, hist_len);
    count_launch();
    return check_launch("obs_history_step");
}
