// Exploration epilogues shared by the CUDA-core (forward.cu) and tcgen05 (forward_tc.cu) policy-step kernels: what happens to
// the actor's output row.  ActorPPO.get_action + convert_action_for_env (reference AgentPPO.py:368-376, 388-390) and
// ActorDiscretePPO.get_action (:407-413).  `get(a)` returns output a of this row.
#pragma once
#include "common.cuh"

enum { kPlain = 0, kGaussian = 1, kCategorical = 2 };

struct PolicyOut {
    const float* eps;  // [rows, A] or nullptr (Gaussian: N(0,1); categorical: Exp(1))
    uint64_t seed, step;
    const uint64_t* step_base;  // optional device counter added to `step` (CUDA-graph replays)
    int64_t env_offset;
    float* action;
    float* logprob;
    float* env_action;
    int32_t* action_index;  // categorical
};

// a = mu + sigma * eps; logprob = sum_a Normal(mu, sigma).log_prob(a)   (torch op order, no contraction)
template <class Get>
DEV void gaussian_epilogue(const b200rl_net& net, const PolicyOut& po, uint64_t rng_step, int64_t row, int J, Get get) {
    float logp = 0.0f;
    for (int a = 0; a < J; ++a) {
        const float mu = get(a);
        const float sd = expf(net.action_std_log[a]);
        float e;
        if (po.eps) {
            e = po.eps[row * J + a];
        } else {
            RolloutNoise nz = rollout_noise(po.seed, (uint64_t)(po.env_offset + row), rng_step, (uint32_t)(a >> 1));
            e = (a & 1) ? nz.normal.y : nz.normal.x;
        }
        const float act = __fadd_rn(__fmul_rn(e, sd), mu);
        const float diff = __fsub_rn(act, mu);
        const float var = __fmul_rn(sd, sd);
        const float lp = __fsub_rn(__fsub_rn(-__fdiv_rn(__fmul_rn(diff, diff), __fmul_rn(2.0f, var)), logf(sd)), kLogSqrt2Pi);
        logp = __fadd_rn(logp, lp);
        po.action[row * J + a] = act;
        po.env_action[row * J + a] = tanhf(act);
    }
    po.logprob[row] = logp;
}

// softmax -> Categorical.sample() -> log_prob.  torch.multinomial's one-draw path is argmax(p / q), q ~ Exp(1); the first
// maximum wins ties (argmax).
template <class Get>
DEV void categorical_epilogue(const PolicyOut& po, uint64_t rng_step, int64_t row, int J, Get get) {
    float m = -INFINITY;
    for (int a = 0; a < J; ++a) m = fmaxf(m, get(a));
    float sum = 0.0f;
    for (int a = 0; a < J; ++a) sum += expf(get(a) - m);
    const float inv_sum = 1.0f / sum;
    float best = -1.0f, best_p = 0.0f;
    int best_a = 0;
    uint4 bits = make_uint4(0u, 0u, 0u, 0u);
    for (int a = 0; a < J; ++a) {
        const float p = expf(get(a) - m) * inv_sum;
        float q;
        if (po.eps) {
            q = po.eps[row * J + a];
        } else {
            if ((a & 3) == 0) bits = rollout_bits(po.seed, (uint64_t)(po.env_offset + row), rng_step, (uint32_t)(a >> 2));
            const uint32_t w = (a & 3) == 0 ? bits.x : (a & 3) == 1 ? bits.y : (a & 3) == 2 ? bits.z : bits.w;
            q = -logf(u32_to_unit_open(w));
        }
        const float race = __fdiv_rn(p, q);
        if (race > best) { best = race; best_a = a; best_p = p; }
    }
    po.action_index[row] = best_a;
    po.logprob[row] = logf(best_p);
}
