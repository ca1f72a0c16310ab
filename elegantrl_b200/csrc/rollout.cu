// Fused rollout on the built-in Pendulum-v1 vec env: the whole Python loop of reference
// AgentPPO._explore_vec_env (elegantrl/agents/AgentPPO.py:87-129) -- ActorPPO.get_action (:368-376), tanh,
// env.step (:119, physics of elegantrl_b200/envs/pendulum.py), the six trajectory stores (:115-123), the
// post-processing (:125-128) -- plus the critic values pass of update_net (:141-143) and V(last_state)
// (:219-220), in ONE persistent kernel.  Env state stays in registers for all H steps; network weights stay in
// shared memory; the trajectory is written once, struct-of-arrays, time-major, with 128-bit coalesced stores
// staged through shared memory.  HBM traffic: 30 B written per env-step (26 B trajectory + 4 B value), 0 read.
//
// This file is the FP32-pipe (FFMA) implementation: one thread per env, hidden vector in registers, weights
// broadcast from shared memory as LDS.128.  The tcgen05 / TMEM implementation for the 64-wide layers lives in
// rollout_tc.cu; both must agree with the oracle to rtol 1e-4.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "rollout_params.cuh"

namespace {

DEV float4 ld4s(const float* p) { return *reinterpret_cast<const float4*>(p); }

// shared-memory image of one 3-Linear net, transposed for float4-over-outputs access
template <int S, int A, int H1, int H2>
struct NetSmem {
    static constexpr int kW1t = 0;                  // [S][H1]
    static constexpr int kB1 = kW1t + S * H1;       // [H1]
    static constexpr int kW2t = kB1 + H1;           // [H1][H2]
    static constexpr int kB2 = kW2t + H1 * H2;      // [H2]
    static constexpr int kW3 = kB2 + H2;            // [A][H2]
    static constexpr int kB3 = kW3 + A * H2;        // [A] (+pad)
    static constexpr int kAvg = kB3 + ((A + 3) / 4) * 4;  // [S] (+pad)
    static constexpr int kStd = kAvg + ((S + 3) / 4) * 4; // [S] holds std + 1e-4
    static constexpr int kFloats = kStd + ((S + 3) / 4) * 4;
};

template <int S, int A, int H1, int H2>
DEV void load_net_smem(const b200rl_net& net, float* w) {
    using L = NetSmem<S, A, H1, H2>;
    const int nt = blockDim.x, tid = threadIdx.x;
    for (int i = tid; i < S * H1; i += nt) { int k = i / H1, j = i - k * H1; w[L::kW1t + i] = net.weight[0][j * S + k]; }
    for (int i = tid; i < H1; i += nt) w[L::kB1 + i] = net.bias[0][i];
    for (int i = tid; i < H1 * H2; i += nt) { int k = i / H2, j = i - k * H2; w[L::kW2t + i] = net.weight[1][j * H1 + k]; }
    for (int i = tid; i < H2; i += nt) w[L::kB2 + i] = net.bias[1][i];
    for (int i = tid; i < A * H2; i += nt) w[L::kW3 + i] = net.weight[2][i];
    for (int i = tid; i < A; i += nt) w[L::kB3 + i] = net.bias[2][i];
    for (int i = tid; i < S; i += nt) {
        w[L::kAvg + i] = net.state_avg ? net.state_avg[i] : 0.0f;
        w[L::kStd + i] = net.state_std ? net.state_std[i] + 1e-4f : 1.0f;
    }
}

// out = Linear3(act(Linear2(act(Linear1(state_norm(obs))))))   -- per thread, weights from shared memory
template <int S, int A, int H1, int H2>
DEV void mlp3_eval(const float* w, const float (&obs)[S], bool has_norm, int act, float (&out)[A]) {
    using L = NetSmem<S, A, H1, H2>;
    float x[S];
#pragma unroll
    for (int i = 0; i < S; ++i) x[i] = has_norm ? (obs[i] - w[L::kAvg + i]) / w[L::kStd + i] : obs[i];
    float h1[H1];
#pragma unroll
    for (int j = 0; j < H1; j += 4) {
        float4 acc = ld4s(w + L::kB1 + j);
#pragma unroll
        for (int i = 0; i < S; ++i) {
            float4 wv = ld4s(w + L::kW1t + i * H1 + j);
            acc.x = fmaf(x[i], wv.x, acc.x); acc.y = fmaf(x[i], wv.y, acc.y);
            acc.z = fmaf(x[i], wv.z, acc.z); acc.w = fmaf(x[i], wv.w, acc.w);
        }
        h1[j] = act_fn_rt(acc.x, act); h1[j + 1] = act_fn_rt(acc.y, act);
        h1[j + 2] = act_fn_rt(acc.z, act); h1[j + 3] = act_fn_rt(acc.w, act);
    }
#pragma unroll
    for (int a = 0; a < A; ++a) out[a] = w[L::kB3 + a];
    constexpr int JC = 16;
#pragma unroll 1
    for (int jc = 0; jc < H2; jc += JC) {
        float acc[JC];
#pragma unroll
        for (int q = 0; q < JC; q += 4) {
            float4 b = ld4s(w + L::kB2 + jc + q);
            acc[q] = b.x; acc[q + 1] = b.y; acc[q + 2] = b.z; acc[q + 3] = b.w;
        }
        const float* w2 = w + L::kW2t + jc;
#pragma unroll
        for (int k = 0; k < H1; ++k) {
#pragma unroll
            for (int q = 0; q < JC; q += 4) {
                float4 wv = ld4s(w2 + k * H2 + q);
                acc[q] = fmaf(h1[k], wv.x, acc[q]); acc[q + 1] = fmaf(h1[k], wv.y, acc[q + 1]);
                acc[q + 2] = fmaf(h1[k], wv.z, acc[q + 2]); acc[q + 3] = fmaf(h1[k], wv.w, acc[q + 3]);
            }
        }
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            float g = act_fn_rt(acc[j], act);
#pragma unroll
            for (int a = 0; a < A; ++a) out[a] = fmaf(g, w[L::kW3 + a * H2 + jc + j], out[a]);
        }
    }
}


template <int H1, int H2>
__global__ void __launch_bounds__(512, 1) rollout_pendulum_kernel(const __grid_constant__ RolloutParams P) {
    constexpr int S = 3, A = 1;
    using L = NetSmem<S, A, H1, H2>;
    extern __shared__ float4 smem4[];
    float* w_act = reinterpret_cast<float*>(smem4);
    float* w_cri = w_act + L::kFloats;
    float* stage = w_cri + L::kFloats;  // [warps][32 * S] staging for the 128-bit state stores

    load_net_smem<S, A, H1, H2>(P.actor, w_act);
    if (P.has_critic) load_net_smem<S, A, H1, H2>(P.critic, w_cri);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_warp0 = n - lane;
    const bool live = n < P.N;
    const int N = P.N;
    // fast (vector) store path needs a full warp of envs and 16-byte aligned rows
    const bool vec_ok = (n_warp0 + 32 <= N) && ((N & 3) == 0);
    float* my_stage = stage + warp * (32 * S);

    float theta = live ? P.theta[n] : 0.0f, theta_dot = live ? P.theta_dot[n] : 0.0f;
    int cur_step = live ? P.cur_step[n] : 0;
    const float sd = expf(P.actor.action_std_log[0]);
    const float log_sd = logf(sd), var2 = __fmul_rn(2.0f, __fmul_rn(sd, sd));
    const bool a_norm = P.actor.state_avg != nullptr, c_norm = P.critic.state_avg != nullptr;
    const int a_act = P.actor.activation, c_act = P.critic.activation;

    for (int t = 0; t < P.H; ++t) {
        float sin_t, cos_t;
        sincosf(theta, &sin_t, &cos_t);
        const float obs[S] = {cos_t, sin_t, theta_dot};
        const size_t row = (size_t)t * N;

        // ---- policy: a = mu + sigma * eps, logprob                             (reference :368-376)
        float mu[A];
        mlp3_eval<S, A, H1, H2>(w_act, obs, a_norm, a_act, mu);
        float e;
        float2 reset_u;
        if (P.eps == nullptr || P.reset_noise == nullptr) {
            RolloutNoise nz = rollout_noise(P.seed, (uint64_t)(P.env_offset + n), P.step_offset + (uint64_t)t, 0u);
            e = nz.normal.x;
            reset_u = nz.uniform;
        }
        if (P.eps && live) e = P.eps[row + n];
        if (P.deterministic) e = 0.0f;
        if (P.reset_noise && live) reset_u = make_float2(P.reset_noise[(row + n) * 2], P.reset_noise[(row + n) * 2 + 1]);
        const float action = __fadd_rn(__fmul_rn(e, sd), mu[0]);
        const float diff = __fsub_rn(action, mu[0]);
        const float logprob = __fsub_rn(__fsub_rn(-__fdiv_rn(__fmul_rn(diff, diff), var2), log_sd), kLogSqrt2Pi);

        // ---- critic value of the pre-step state (what update_net :141-143 recomputes)
        float val[A] = {0.0f};
        if (P.has_critic) mlp3_eval<S, A, H1, H2>(w_cri, obs, c_norm, c_act, val);

        // ---- env.step(tanh(action))   (elegantrl_b200/envs/pendulum.py, op for op, no FMA contraction)
        const float torque = fminf(fmaxf(__fmul_rn(tanhf(action), 2.0f), -2.0f), 2.0f);
        const float th_n = __fsub_rn(remainder_pos(__fadd_rn(theta, kPi), kTwoPi), kPi);
        const float cost = __fadd_rn(__fadd_rn(__fmul_rn(th_n, th_n), __fmul_rn(0.1f, __fmul_rn(theta_dot, theta_dot))),
                                     __fmul_rn(0.001f, __fmul_rn(torque, torque)));
        const float reward = __fmul_rn(__fmul_rn(cost, -0.5f), P.reward_scale);
        const float accel = __fadd_rn(__fmul_rn(15.0f, sin_t), __fmul_rn(3.0f, torque));
        float new_theta_dot = fminf(fmaxf(__fadd_rn(theta_dot, __fmul_rn(accel, 0.05f)), -8.0f), 8.0f);
        float new_theta = __fadd_rn(theta, __fmul_rn(new_theta_dot, 0.05f));
        cur_step += 1;
        const bool truncate = cur_step >= P.max_step;
        if (truncate) {  // auto-reset: theta ~ U(-pi, pi), theta_dot ~ U(-1, 1)
            new_theta = __fmul_rn(__fsub_rn(__fmul_rn(reset_u.x, 2.0f), 1.0f), kPi);
            new_theta_dot = __fsub_rn(__fmul_rn(reset_u.y, 2.0f), 1.0f);
            cur_step = 0;
        }

        // ---- trajectory stores, SoA, time-major
        if (vec_ok) {
            my_stage[lane * S + 0] = obs[0]; my_stage[lane * S + 1] = obs[1]; my_stage[lane * S + 2] = obs[2];
            __syncwarp();
            float4* dst = reinterpret_cast<float4*>(P.states + (row + n_warp0) * S);
            for (int i = lane; i < 8 * S; i += 32) dst[i] = ld4s(my_stage + 4 * i);
            __syncwarp();
            // 32 bool bytes of the warp -> 8 x 32-bit stores
            const unsigned um_bits = __ballot_sync(0xffffffffu, !truncate);
            if (lane < 8) {
                unsigned m4 = (um_bits >> (4 * lane)) & 0xFu;
                unsigned word = (m4 & 1u) | ((m4 & 2u) << 7) | ((m4 & 4u) << 14) | ((m4 & 8u) << 21);
                reinterpret_cast<unsigned*>(P.unmasks + row + n_warp0)[lane] = word;
                reinterpret_cast<unsigned*>(P.undones + row + n_warp0)[lane] = 0x01010101u;  // never terminal
            }
        } else if (live) {
            P.states[(row + n) * S + 0] = obs[0]; P.states[(row + n) * S + 1] = obs[1]; P.states[(row + n) * S + 2] = obs[2];
            P.unmasks[row + n] = truncate ? 0 : 1;
            P.undones[row + n] = 1;
        }
        if (live) {
            P.actions[row + n] = action;
            P.logprobs[row + n] = logprob;
            P.rewards[row + n] = reward;
            if (P.values) P.values[row + n] = val[0];
        }
        theta = new_theta;
        theta_dot = new_theta_dot;
    }

    // ---- epilogue: last_state, V(last_state), env state back to HBM
    float sin_t, cos_t;
    sincosf(theta, &sin_t, &cos_t);
    const float obs[S] = {cos_t, sin_t, theta_dot};
    float val[A] = {0.0f};
    if (P.has_critic && P.last_value) mlp3_eval<S, A, H1, H2>(w_cri, obs, c_norm, c_act, val);
    if (live) {
        P.last_state[(size_t)n * S + 0] = obs[0]; P.last_state[(size_t)n * S + 1] = obs[1]; P.last_state[(size_t)n * S + 2] = obs[2];
        if (P.has_critic && P.last_value) P.last_value[n] = val[0];
        P.theta[n] = theta; P.theta_dot[n] = theta_dot; P.cur_step[n] = cur_step;
    }
}

bool is_mlp3(const b200rl_net* net, int s, int h1, int h2, int out) {
    return net->num_linear == 3 && net->dims[0] == s && net->dims[1] == h1 && net->dims[2] == h2 && net->dims[3] == out;
}

template <int H1, int H2>
int launch_rollout(const RolloutParams& P, cudaStream_t stream) {
    using L = NetSmem<3, 1, H1, H2>;
    int dev = 0, sms = 148;
    B200RL_CHECK_CUDA(cudaGetDevice(&dev));
    B200RL_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // one persistent CTA per SM when the env count allows it: threads = ceil(N / SMs) rounded up to a warp
    int threads = ((P.N + sms - 1) / sms + 31) / 32 * 32;
    threads = threads < 64 ? 64 : (threads > 512 ? 512 : threads);
    int grid = (P.N + threads - 1) / threads;
    size_t smem = (size_t)(2 * L::kFloats + (threads / 32) * 32 * 3) * sizeof(float);
    auto kern = rollout_pendulum_kernel<H1, H2>;
    B200RL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, stream>>>(P);
    B200RL_COUNT_LAUNCH(1);
    B200RL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

int b200rl_rollout_pendulum(const b200rl_rollout_args* a, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200RL_REQUIRE(a != nullptr, "rollout_pendulum: args is NULL");
    if (int rc = b200rl_validate_net(a->actor, "rollout.actor", true)) return rc;
    if (a->critic)
        if (int rc = b200rl_validate_net(a->critic, "rollout.critic", false)) return rc;
    B200RL_REQUIRE(a->num_envs >= 1 && a->horizon_len >= 1 && a->max_step >= 1, "rollout_pendulum: N=%d H=%d max_step=%d",
                   a->num_envs, a->horizon_len, a->max_step);
    B200RL_REQUIRE(a->theta && a->theta_dot && a->cur_step && a->states && a->actions && a->logprobs && a->rewards &&
                       a->undones && a->unmasks && a->last_state, "rollout_pendulum: NULL buffer");
    B200RL_REQUIRE(!a->values || a->critic, "rollout_pendulum: values requested without a critic");
    const int h1 = a->actor->dims[1], h2 = a->actor->dims[2];
    B200RL_REQUIRE(is_mlp3(a->actor, 3, h1, h2, 1), "rollout_pendulum: actor must be 3 -> h1 -> h2 -> 1 (two hidden layers)");
    B200RL_REQUIRE(!a->critic || is_mlp3(a->critic, 3, h1, h2, 1), "rollout_pendulum: critic must have the actor's hidden dims");

    RolloutParams P{};
    P.actor = *a->actor;
    if (a->critic) P.critic = *a->critic;
    P.has_critic = a->critic != nullptr;
    P.N = a->num_envs; P.H = a->horizon_len; P.max_step = a->max_step; P.reward_scale = a->reward_scale;
    P.theta = a->theta; P.theta_dot = a->theta_dot; P.cur_step = a->cur_step;
    P.states = a->states; P.actions = a->actions; P.logprobs = a->logprobs; P.rewards = a->rewards;
    P.undones = a->undones; P.unmasks = a->unmasks; P.values = a->values; P.last_state = a->last_state;
    P.last_value = a->last_value; P.eps = a->eps; P.reset_noise = a->reset_noise;
    P.deterministic = (a->flags & B200RL_ROLLOUT_DETERMINISTIC) ? 1 : 0;
    P.seed = a->seed; P.step_offset = a->step_offset; P.env_offset = a->env_offset;

    // 2x64 GELU actor + critic (BASELINE config 2): the tcgen05 kernel with the layer-2 A operand in tensor memory
    // (rollout_ts.cu).  B200RL_ROLLOUT=tc selects the earlier tcgen05 kernel with the shared-memory operand ring
    // (rollout_tc.cu), =ffma the FP32-pipe kernel below -- both kept as independent cross-checks (the tests run all three).
    const char* mode = getenv("B200RL_ROLLOUT");
    const bool want_ffma = mode && strcmp(mode, "ffma") == 0;
    if (h1 == 64 && h2 == 64 && a->critic && !want_ffma && a->actor->activation == B200RL_ACT_GELU &&
        a->critic->activation == B200RL_ACT_GELU) {
        if (mode && strcmp(mode, "tc") == 0) return b200rl_launch_rollout_tc(P, stream);
        return b200rl_launch_rollout_ts(P, stream);
    }
    if (h1 == 64 && h2 == 64) return launch_rollout<64, 64>(P, stream);
    if (h1 == 128 && h2 == 64) return launch_rollout<128, 64>(P, stream);
    b200rl_set_error("rollout_pendulum: no fused kernel for hidden dims %dx%d (built: 64x64, 128x64); "
                     "use b200rl_policy_step with the env's own step()", h1, h2);
    return 3;
}

}  // extern "C"
