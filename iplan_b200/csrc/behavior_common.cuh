// Launch arguments shared by the two K1b kernels (behavior_step.cu: mma.sync; behavior_tc5.cu: tcgen05).
#pragma once
#include "common.cuh"

namespace iplan {

struct BehArgs {
    const float* params; int64_t param_stride;
    iplan_view window, hid, lat_prev, lat_out;
    float coef;
    int n_envs, n_slots, obs_dim, latent_dim, hist_len;
    int64_t win_step;      // 0: window rows are contiguous ([hist_len][obs_dim] per node); else element stride between rows
    int win_pad;           // (win_step != 0) leading window rows that are zero padding; window.ptr = the first real row
};

// tcgen05 kernel (behavior_tc5.cu); returns 0 / error
int launch_behavior_tc5(const BehArgs& a, int n_agents, cudaStream_t st);
bool behavior_tc5_supports(const BehArgs& a);

}  // namespace iplan
