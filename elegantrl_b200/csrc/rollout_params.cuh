// Kernel parameter block shared by the two fused-rollout implementations (rollout.cu: FP32 pipe; rollout_tc.cu:
// tcgen05 / TMEM).  Mirrors b200rl_rollout_args (include/b200rl.h) with the net descriptors by value.
#pragma once
#include "common.cuh"

struct RolloutParams {
    b200rl_net actor, critic;
    int has_critic;
    int N, H, max_step;
    float reward_scale;
    float* theta; float* theta_dot; int* cur_step;
    float* states; float* actions; float* logprobs; float* rewards;
    uint8_t* undones; uint8_t* unmasks; float* values; float* last_state; float* last_value;
    const float* eps; const float* reset_noise;
    uint64_t seed, step_offset;
    int64_t env_offset;
    int deterministic;  // B200RL_ROLLOUT_DETERMINISTIC: zero policy noise (evaluation rollout)
    int rows_per_cta;   // rollout_ts.cu: envs per persistent CTA (set by its launcher)
};

// th.remainder(a, b) for b > 0 (exact: fmod then sign fix, as ATen)
__device__ __forceinline__ float remainder_pos(float a, float b) {
    float r = fmodf(a, b);
    return (r < 0.0f) ? __fadd_rn(r, b) : r;
}

int b200rl_launch_rollout_tc(const RolloutParams& P, cudaStream_t stream);  // rollout_tc.cu
int b200rl_launch_rollout_ts(const RolloutParams& P, cudaStream_t stream);  // rollout_ts.cu (A operand in tensor memory)
