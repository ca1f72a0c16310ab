/* libb200rl -- C ABI of the B200-native on-policy engine (rollout -> GAE -> PPO update).
 *
 * The reference (AI4Finance-Foundation/ElegantRL) has no FFI for this path: the boundary is the Python
 * duck-typed Agent API.  This header is the C-ABI a binding for that boundary calls;
 * each entry point names the reference function it replaces (paths relative to the reference checkout).
 * The Python shim that mirrors the reference's `AgentPPO` on top of it is elegantrl_b200/agents/AgentPPO.py;
 * the ctypes stub is elegantrl_b200/_lib.py (also shown in INTEGRATION.md).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless a comment says host
 *  - all work is enqueued on the caller's CUDA stream (`stream` = cudaStream_t, may be NULL = default stream);
 *    no entry point synchronises the device or allocates device memory
 *  - scratch memory is a caller-provided `workspace` (size from b200rl_workspace_bytes)
 *  - every entry point returns 0 on success, non-zero on failure; b200rl_last_error() gives the text of the
 *    last failure on the calling thread.  One host thread per engine (the reference is single-threaded per
 *    process).
 *  - float = IEEE fp32; masks are 1-byte bools exactly as torch.bool tensors; indices are int64.
 */
#ifndef B200RL_H_
#define B200RL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200RL_API __attribute__((visibility("default")))
#else
#define B200RL_API
#endif

#define B200RL_MAX_LINEAR 8 /* nn.Linear layers per net = len(net_dims) + 1 */
#define B200RL_ACT_GELU 0   /* nn.GELU(), exact erf form -- reference AgentBase.py:353-354 */
#define B200RL_ACT_RELU 1   /* helloworld nets -- reference helloworld_PPO_single_file.py:172-212 */

/* One MLP of the reference's `build_mlp` (elegantrl/agents/AgentBase.py:345-360) plus the state_norm
 * statistics of ActorPPO / CriticPPO (elegantrl/agents/AgentPPO.py:357-361, 432-441).  Pointers alias the
 * storages of the caller's nn.Parameter tensors: the engine reads and (in the update) writes them in place. */
typedef struct b200rl_net {
    int32_t num_linear;                   /* 1..B200RL_MAX_LINEAR */
    int32_t activation;                   /* B200RL_ACT_* applied after every Linear but the last */
    int32_t dims[B200RL_MAX_LINEAR + 1];  /* dims[0] = state_dim, dims[num_linear] = output dim */
    int32_t reserved;
    float* weight[B200RL_MAX_LINEAR];     /* [dims[l+1], dims[l]] row-major (nn.Linear.weight) */
    float* bias[B200RL_MAX_LINEAR];       /* [dims[l+1]] */
    const float* state_avg;               /* [dims[0]] or NULL: no state_norm at all (helloworld) */
    const float* state_std;               /* [dims[0]]; normalised input = (s - avg) / (std + 1e-4) */
    float* action_std_log;                /* Gaussian actor: [action_dim] (ActorPPO.action_std_log); critic and the
                                             categorical actor (ActorDiscretePPO, AgentPPO.py:393-425): NULL */
} b200rl_net;

/* torch.optim.Adam state of one net (reference AgentPPO.py:24-25; stepped by AgentBase.py:239-248), tensor
 * order W0, b0, W1, b1, ..., action_std_log.  Pointers alias optimizer.state[p]['exp_avg' / 'exp_avg_sq']. */
typedef struct b200rl_adam {
    float* exp_avg_w[B200RL_MAX_LINEAR];
    float* exp_avg_b[B200RL_MAX_LINEAR];
    float* exp_avg_sq_w[B200RL_MAX_LINEAR];
    float* exp_avg_sq_b[B200RL_MAX_LINEAR];
    float* exp_avg_std;                   /* actor only, else NULL */
    float* exp_avg_sq_std;
    float lr, beta1, beta2, eps;
    int64_t step;                         /* optimizer steps already taken; advanced by the engine (host field) */
} b200rl_adam;

/* Hyper-parameters of AgentPPO.update_objectives (reference AgentPPO.py:27-30, 189-204). */
typedef struct b200rl_ppo_hyper {
    float ratio_clip;                     /* AgentPPO.ratio_clip (0.25) */
    float lambda_entropy;                 /* AgentPPO.lambda_entropy (0.001) */
    float clip_grad_norm;                 /* Config.clip_grad_norm (3.0); <= 0 disables clipping */
    int32_t flags;                        /* 0 = elegantrl AgentPPO; B200RL_PPO_* bits select the helloworld variant */
} b200rl_ppo_hyper;
/* variant bits of b200rl_ppo_hyper.flags -- reference helloworld/helloworld_PPO_single_file.py:319-342, 366-370 */
#define B200RL_PPO_SMOOTH_L1 1       /* critic criterion SmoothL1Loss (:246) instead of MSELoss */
#define B200RL_PPO_MIN_CLIP 2        /* surrogate = min(adv*ratio, adv*clamp(ratio, 1-c, 1+c)) (:337-339) */
#define B200RL_PPO_ENTROPY_BONUS 4   /* objective = surrogate + lambda_entropy * entropy (:340); default subtracts */
#define B200RL_PPO_ACTOR_UNMASKED 8  /* actor terms are not multiplied by unmask (:339-340) */
#define B200RL_PPO_CRITIC_MASK_MEAN 16 /* critic loss weighted by mean(unmask) of the minibatch: the [B] x [B, 1]
                                          broadcast of (:325, 332) averages to mean(loss) * mean(unmask) */
#define B200RL_PPO_A2C 32            /* AgentA2C.update_objectives (elegantrl/agents/AgentPPO.py:306-311): obj_actor =
                                          mean over [B, A] of advantage * new_logprob -- no ratio, no clip; callers also set
                                          ACTOR_UNMASKED and lambda_entropy = 0.  Coherent for single-env buffers only */

/* The training buffer of AgentPPO.update_net after the GAE pass (reference AgentPPO.py:151):
 * (states, actions, unmasks, logprobs, advantages, reward_sums), time-major [H, N, ...]. */
typedef struct b200rl_train_buffer {
    const float* states;                  /* [H, N, S] */
    const float* actions;                 /* [H, N, A] fp32; int32 [H, N] action indices when discrete_actions != 0 */
    const uint8_t* unmasks;               /* [H, N] bool */
    const float* logprobs;                /* [H, N] */
    const float* advantages;              /* [H, N] */
    const float* reward_sums;             /* [H, N] */
    const float* adv_stats;               /* [4] = {mean, std, ...} from b200rl_adv_stats: (adv - mean)/(std + 1e-5)
                                             is applied at gather time; NULL when `advantages` is already
                                             normalised (reference :149) */
    int32_t horizon_len;                  /* H */
    int32_t num_envs;                     /* N */
    int32_t discrete_actions;             /* != 0: AgentDiscretePPO (AgentPPO.py:252-270, 103-104): `actions` holds int32
                                             indices [H, N], the actor's outputs are logits of a Categorical
                                             (log-prob / entropy of ActorDiscretePPO.get_logprob_entropy :415-421) and
                                             the actor net carries no action_std_log.  Packed records keep the index as
                                             a float in the first action slot. */
    int32_t reserved;
} b200rl_train_buffer;

/* Fused rollout on the built-in Pendulum-v1 vec env: replaces the Python loop of
 * AgentPPO._explore_vec_env (reference AgentPPO.py:87-129) including ActorPPO.get_action (:368-376), the
 * env.step call (:119) and -- when `critic` is given -- the values pass of update_net (:141-143). */
typedef struct b200rl_rollout_args {
    const b200rl_net* actor;              /* host pointer */
    const b200rl_net* critic;             /* host pointer or NULL (then `values` is not written) */
    int32_t num_envs;                     /* N (this rank's shard) */
    int32_t horizon_len;                  /* H */
    int32_t max_step;                     /* episode truncation length (200 for Pendulum-v1) */
    float reward_scale;                   /* AgentBase.reward_scale, applied to the stored rewards (:126) */
    float* theta;                         /* [N] in/out env state */
    float* theta_dot;                     /* [N] in/out */
    int32_t* cur_step;                    /* [N] in/out steps since reset */
    float* states;                        /* out [H, N, 3] pre-step observation */
    float* actions;                       /* out [H, N, 1] raw (pre-tanh) action */
    float* logprobs;                      /* out [H, N] */
    float* rewards;                       /* out [H, N] (already * reward_scale) */
    uint8_t* undones;                     /* out [H, N] = !terminal */
    uint8_t* unmasks;                     /* out [H, N] = !truncate */
    float* values;                        /* out [H, N] V(s_t) of `critic`, or NULL */
    float* last_state;                    /* out [N, 3] observation after the last step (agent.last_state) */
    float* last_value;                    /* out [N] V(last_state) (reference :219-220) or NULL */
    const float* eps;                     /* optional injected N(0,1) policy noise [H, N, 1]; NULL = Philox */
    const float* reset_noise;             /* optional injected U[0,1) reset noise [H, N, 2]; NULL = Philox */
    uint64_t seed;                        /* Philox key */
    uint64_t step_offset;                 /* global step index of t = 0 (Philox counter) */
    int64_t env_offset;                   /* global env index of local env 0 (Philox counter, multi-GPU shards) */
    int32_t flags;                        /* B200RL_ROLLOUT_*: 0 = exploration (ActorPPO.get_action) */
    int32_t reserved;
} b200rl_rollout_args;
/* Deterministic policy: the env receives tanh(mean) = ActorPPO.forward (AgentPPO.py:363-366), the policy noise is zero --
 * the evaluation rollout of the reference's Evaluator (elegantrl/train/evaluator.py:200-216: `action = actor(state)`). */
#define B200RL_ROLLOUT_DETERMINISTIC 1

B200RL_API const char* b200rl_version(void);
B200RL_API const char* b200rl_last_error(void);
/* Number of CUDA kernels this library has launched in this process (bench.py reports it as gpu_launches). */
B200RL_API int64_t b200rl_launch_count(void);

/* Bytes of scratch the update entry points need for this pair of nets (flat gradient buffer + counters). */
B200RL_API int64_t b200rl_workspace_bytes(const b200rl_net* actor, const b200rl_net* critic);
/* Offset (bytes) and length (floats) of the flat fp32 gradient buffer inside the workspace: actor tensors
 * (W0, b0, ..., action_std_log), zero padding up to a multiple of 4 floats, then critic tensors.  This is the
 * buffer a multi-GPU caller all-reduces (b200rl_grad_numel includes the padding). */
B200RL_API int64_t b200rl_workspace_grad_offset(void);
B200RL_API int64_t b200rl_grad_numel(const b200rl_net* actor, const b200rl_net* critic);

/* net(state_norm(x)) for `rows` rows: CriticPPO.forward (AgentPPO.py:435-438; out_tanh = 0) or
 * ActorPPO.forward (:363-366; out_tanh = 1).  x [rows, dims[0]] -> out [rows, dims[num_linear]]. */
B200RL_API int b200rl_mlp_forward(const b200rl_net* net, const float* x, int64_t rows, float* out, int32_t out_tanh,
                       void* stream);

/* One exploration step for an arbitrary (external) vec env: ActorPPO.get_action (AgentPPO.py:368-376) +
 * convert_action_for_env (:388-390) + optional critic value.  eps [rows, A] injected noise or NULL (Philox
 * keyed by seed / step / env_offset + row).  Outputs: action [rows, A] (pre-tanh), logprob [rows],
 * env_action [rows, A] = tanh(action), value [rows] (if critic and value are non-NULL). */
B200RL_API int b200rl_policy_step(const b200rl_net* actor, const b200rl_net* critic, const float* state, int64_t rows,
                       const float* eps, uint64_t seed, uint64_t step, int64_t env_offset, float* action,
                       float* logprob, float* env_action, float* value, void* stream);
/* Optional device-resident Philox step base for the two policy-step entry points: when set (non-NULL), the kernels use
 * step + *step_base.  A caller that captures its rollout loop in a CUDA graph keeps the counter in device memory and
 * advances it inside the graph, so that every replay draws fresh noise.  Host-side, per calling thread; NULL resets. */
B200RL_API void b200rl_set_policy_step_base(const uint64_t* step_base);

/* The same for the categorical policy: ActorDiscretePPO.get_action (AgentPPO.py:407-413) -- softmax, one draw of
 * torch.multinomial (its single-sample path is the exponential race argmax(p / q), q ~ Exp(1)), log-prob of the drawn
 * index.  expo [rows, A] injected Exp(1) noise or NULL (Philox).  Outputs: action int32 [rows] (what
 * convert_action_for_env :423-425 hands to env.step after .long()), logprob [rows], value [rows] (optional). */
B200RL_API int b200rl_policy_step_discrete(const b200rl_net* actor, const b200rl_net* critic, const float* state,
                                           int64_t rows, const float* expo, uint64_t seed, uint64_t step,
                                           int64_t env_offset, int32_t* action, float* logprob, float* value,
                                           void* stream);

B200RL_API int b200rl_rollout_pendulum(const b200rl_rollout_args* args, void* stream);

/* AgentPPO.get_advantages (AgentPPO.py:207-232) + reward_sums (:146) + the reduction inputs of the
 * normalisation (:149) in one reverse-scan kernel.  rewards / undones are updated IN PLACE for truncated steps
 * exactly as the reference does (:211-214).  stat_sums (device double[4]) receives {sum(adv), sum over the
 * [::4, ::4] lattice of adv, of adv^2, sum(adv^2)}; the lattice is taken on the GLOBAL env index
 * env_offset + n so that shards agree.  A multi-GPU caller all-reduces stat_sums before b200rl_adv_stats. */
B200RL_API int b200rl_gae(float* rewards, uint8_t* undones, const uint8_t* unmasks, const float* values,
               const float* last_value, int32_t horizon_len, int32_t num_envs, float gamma, float lambda_gae,
               int32_t if_use_v_trace, int64_t env_offset, float* advantages, float* reward_sums,
               double* stat_sums, void* stream);
/* stats_out (device float[4]) = {mean, std_unbiased(lattice), 1 / (std + 1e-5), 0} from the (all-reduced)
 * sums.  count_all = H * N_global; count_lattice = ceil(H/4) * ceil(N_global/4), or 0 for the unbiased std over
 * the whole buffer (helloworld_PPO_single_file.py:296). */
B200RL_API int b200rl_adv_stats(const double* stat_sums, int64_t count_all, int64_t count_lattice, float* stats_out,
                     void* stream);
/* advantages = (advantages - mean) / (std + 1e-5) in place (materialises reference AgentPPO.py:149). */
B200RL_API int b200rl_normalize_adv(float* advantages, int64_t count, const float* stats, void* stream);

/* update_times minibatch updates: AgentPPO.update_net loop (AgentPPO.py:158-165) over update_objectives
 * (:173-205) with AgentBase.optimizer_backward (AgentBase.py:239-248) per net: gather -> critic/actor
 * forward -> losses -> backward -> per-net grad-norm clip -> Adam, all on device.
 * ids: device int64 [update_times, batch_size] in [0, H*N) (t = id % H, n = id / H as reference :178-180) or
 * NULL to draw them on device (Philox keyed by seed / draw_offset + update index).
 * out_scalars (device float[3]) = mean over the updates of (obj_critic, obj_surrogate, obj_entropy)
 * (reference :168-171).  opt->step fields are advanced on the host. */
B200RL_API int b200rl_ppo_update(const b200rl_net* actor, const b200rl_net* critic, b200rl_adam* actor_opt,
                      b200rl_adam* critic_opt, const b200rl_train_buffer* buffer, const b200rl_ppo_hyper* hyper,
                      int32_t batch_size, int32_t update_times, const int64_t* ids, uint64_t seed,
                      uint64_t draw_offset, float* out_scalars, void* workspace, int64_t workspace_bytes,
                      void* stream);

/* Env-sharded (multi-GPU) update, preferred form: every rank packs ITS share of all the minibatches of one
 * update_net -- update_times x local_batch sampled transitions -- into records of R = roundup4(state_dim + action_dim)
 * + 4 floats {state, action, pad, unmask, logprob, normalised advantage, reward_sum}; the caller all-gathers them (ONE
 * collective per cycle, ~30 KB) and then calls b200rl_ppo_update on every rank with a PACKED buffer:
 * b200rl_train_buffer{states = records, horizon_len = 0, num_envs = number of records}, ids[u][i] = record index.
 * ids: [update_times, local_batch] local indices into this rank's [H, N] buffer, or NULL (Philox). */
B200RL_API int b200rl_pack_minibatches(const b200rl_train_buffer* buffer, int32_t state_dim, int32_t action_dim,
                                       int32_t local_batch, int32_t update_times, const int64_t* ids, uint64_t seed,
                                       uint64_t draw_offset, float* out_records, void* stream);

/* Multi-GPU split of one minibatch update (gradient all-reduce form).  b200rl_ppo_grads leaves this rank's gradient SUM over its
 * `local_batch` samples, already divided by `global_batch`, in the workspace's flat buffer (zeroing it first)
 * and accumulates the three loss sums into loss_sums (device double[3], caller zeroes once per update_net);
 * the caller all-reduces (sum) the flat buffer; b200rl_ppo_apply does clip + Adam from it on every rank. */
B200RL_API int b200rl_ppo_grads(const b200rl_net* actor, const b200rl_net* critic, const b200rl_train_buffer* buffer,
                     const b200rl_ppo_hyper* hyper, int32_t local_batch, int32_t global_batch,
                     const int64_t* ids, uint64_t seed, uint64_t draw_offset, double* loss_sums,
                     void* workspace, int64_t workspace_bytes, void* stream);
B200RL_API int b200rl_ppo_apply(const b200rl_net* actor, const b200rl_net* critic, b200rl_adam* actor_opt,
                     b200rl_adam* critic_opt, const b200rl_ppo_hyper* hyper, void* workspace,
                     int64_t workspace_bytes, void* stream);
/* Env-sharded update with the gradient all-reduce INSIDE the update kernel, over peer-mapped memory (NVLink): no
 * host-issued collective at all.  Every rank calls it with ITS shard of the buffer; per minibatch every rank computes the
 * gradient of its batch_size / world samples on tcgen05 tiles (update_tc.cu), stores it in its own exchange buffer,
 * raises a flag in every peer's flag array, waits for the peers' flags and sums the `world` buffers in RANK ORDER with
 * loads from the peers' memory -- the same order on every rank, so clip + Adam produce bit-identical replicas and no
 * parameter broadcast is needed.  The advantage statistics (reference AgentPPO.py:149; b200rl_gae's stat_sums of every
 * shard) ride on the same exchange before the first minibatch.  This is the "NCCL all-reduce of MLP gradients" of the
 * env-sharded design fused into the compute kernel; it replaces the reference's host-pipe trajectory
 * gather (elegantrl/train/run.py:305-320).  Nets must have the shape b200rl_update_tc_supported accepts.
 *   data[r] / flags[r]: rank r's exchange buffer (b200rl_peer_exchange_floats floats) / flag array (B200RL_PX_FLAGS
 *   uint32, zero-initialised once) as mapped into THIS process (symmetric allocation; data[rank] is the own one).
 *   epoch: flags written by this call are epoch + 1 ... epoch + update_times; the caller advances it by update_times.
 * exchange_mode 0 is the above; exchange_mode 1 (batch_size <= 128) exchanges DATA instead of gradients, once per call: the
 * minibatch indices do not depend on the parameters, so every rank packs its share of ALL minibatches of this update_net
 * (update_times x batch_size / world records, the layout of b200rl_pack_minibatches, advantages already normalised with the
 * exchanged statistics) into its exchange buffer, one flag round makes them visible, and every rank runs the identical
 * full-minibatch update, gathering its 128 samples from all ranks' buffers with peer loads -- two flag rounds per cycle
 * instead of one per minibatch, no reduction, replicas bit-identical by construction. */
#define B200RL_MAX_PEERS 8
#define B200RL_PX_FLAGS 64
typedef struct b200rl_peer_exchange {
    int32_t rank, world;
    float* data[B200RL_MAX_PEERS];
    uint32_t* flags[B200RL_MAX_PEERS];
    uint32_t epoch;                       /* minibatches exchanged so far (gradient mode); the caller advances it */
    uint32_t reserved;                    /* update_net calls so far (record mode: flag value and buffer parity); caller advances */
} b200rl_peer_exchange;
B200RL_API int64_t b200rl_workspace_error_offset(void);
B200RL_API int32_t b200rl_update_tc_supported(const b200rl_net* actor, const b200rl_net* critic, const b200rl_ppo_hyper* hyper);
B200RL_API int64_t b200rl_peer_exchange_floats(const b200rl_net* actor, const b200rl_net* critic, int32_t local_batch,
                                               int32_t update_times);
/* batch_size is the GLOBAL minibatch (a multiple of world, <= 128 * world); ids: [update_times, batch_size / world] local
 * indices or NULL.  stat_sums: this shard's device double[4] from b200rl_gae; count_all / count_lattice as b200rl_adv_stats
 * (global counts); adv_stats_out: device float[4], receives {mean, std, 1 / (std + 1e-5), 0}.  buffer->adv_stats is ignored
 * (the kernel normalises with the statistics it has just reduced).  A peer that does not answer within ~2 s makes the kernel
 * give up WITHOUT hanging the GPU and set the uint32 at b200rl_workspace_error_offset() of the workspace to 1 (the caller
 * checks it when it reads out_scalars; results are invalid then). */
B200RL_API int b200rl_ppo_update_sharded(const b200rl_net* actor, const b200rl_net* critic, b200rl_adam* actor_opt,
                                         b200rl_adam* critic_opt, const b200rl_train_buffer* buffer,
                                         const b200rl_ppo_hyper* hyper, int32_t batch_size, int32_t update_times,
                                         const int64_t* ids, uint64_t seed, uint64_t draw_offset, const double* stat_sums,
                                         int64_t count_all, int64_t count_lattice, float* adv_stats_out, float* out_scalars,
                                         void* workspace, int64_t workspace_bytes, const b200rl_peer_exchange* px,
                                         int32_t exchange_mode, void* stream);

/* out_scalars[i] = loss_sums[i] / update_times (after the caller all-reduced loss_sums if sharded). */
B200RL_API int b200rl_loss_means(const double* loss_sums, int32_t update_times, float* out_scalars, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Off-policy path (BASELINE configs[3]): ReplayBuffer + AgentSAC.
 * ------------------------------------------------------------------------------------------------------------------ */
#define B200RL_SAC_MAX_ENSEMBLES 8
#define B200RL_MAX_GROUP_TENSORS 64

/* ActorSAC (elegantrl/agents/AgentSAC.py:167-198): state encoder + one Linear producing [mean | log_std]. */
typedef struct b200rl_sac_actor {
    b200rl_net net_s;                     /* build_mlp([S, *net_dims], if_raw_out=False): activation after EVERY Linear */
    b200rl_net net_a;                     /* one Linear net_dims[-1] -> 2 * action_dim */
} b200rl_sac_actor;

/* CriticEnsemble (AgentSAC.py:244-259): shared raw Linear encoder of (state, action), E decoder MLPs -> Q [B, E]. */
typedef struct b200rl_sac_critic {
    b200rl_net encoder;                   /* one Linear (S + A) -> net_dims[0], no activation */
    int32_t num_ensembles;                /* 1..B200RL_SAC_MAX_ENSEMBLES */
    int32_t reserved;
    b200rl_net decoder[B200RL_SAC_MAX_ENSEMBLES];   /* build_mlp([*net_dims, 1]) each */
} b200rl_sac_critic;

/* One torch.optim.Adam over a flat parameter list (optimizer order), stepped by AgentBase.optimizer_backward
 * (elegantrl/agents/AgentBase.py:239-248: clip_grad_norm_ over the whole list, then Adam.step). */
typedef struct b200rl_param_group {
    int32_t num_tensors;
    int32_t reserved;
    float* param[B200RL_MAX_GROUP_TENSORS];
    float* exp_avg[B200RL_MAX_GROUP_TENSORS];
    float* exp_avg_sq[B200RL_MAX_GROUP_TENSORS];
    int32_t numel[B200RL_MAX_GROUP_TENSORS];
    float lr, beta1, beta2, eps;
    int64_t step;                         /* steps already taken; advanced by the engine (host field) */
} b200rl_param_group;

/* ReplayBuffer storage (elegantrl/train/replay_buffer.py:55-59): time-major rings [max_size, num_seqs, ...], all fp32
 * (undones / unmasks are float32 in the reference).  `max_size` is the TIME length. */
typedef struct b200rl_replay_buffer {
    float* states;                        /* [max_size, num_seqs, state_dim] */
    float* actions;                       /* [max_size, num_seqs, action_dim] (tanh'ed actions, AgentSAC.py:176-182) */
    float* rewards;                       /* [max_size, num_seqs] */
    float* undones;                       /* [max_size, num_seqs] */
    float* unmasks;                       /* [max_size, num_seqs] */
    int32_t max_size, num_seqs, state_dim, action_dim;
} b200rl_replay_buffer;

typedef struct b200rl_sac_hyper {
    float gamma;                          /* Config.gamma */
    float soft_update_tau;                /* Config.soft_update_tau (AgentBase.py:270-278) */
    float clip_grad_norm;                 /* Config.clip_grad_norm, applied per optimizer */
    float target_entropy;                 /* AgentSAC: +log(action_dim) (:31) */
} b200rl_sac_hyper;

/* ReplayBuffer.update (replay_buffer.py:78-118): rows [p, p + rows) (mod max_size) of the five rings <- one rollout
 * (states [rows, N, S], actions [rows, N, A], rewards [rows, N] fp32; undones / unmasks [rows, N] torch.bool, converted to
 * the buffer's float32).  The pointer arithmetic (p, cur_size, if_full) stays with the caller, as in the reference. */
B200RL_API int b200rl_replay_append(const b200rl_replay_buffer* buffer, int32_t p, int32_t rows, const float* states,
                                    const float* actions, const float* rewards, const uint8_t* undones,
                                    const uint8_t* unmasks, void* stream);
/* ActorSAC.get_action (AgentSAC.py:176-182): action = tanh(mean + clamp(log_std, -16, 2).exp() * eps) for `rows` states.
 * eps [rows, A] injected N(0,1) or NULL (Philox keyed by seed / step / env_offset + row). */
B200RL_API int b200rl_sac_policy_step(const b200rl_sac_actor* actor, const float* state, int64_t rows, const float* eps,
                                      uint64_t seed, uint64_t step, int64_t env_offset, float* action, void* stream);
B200RL_API int64_t b200rl_sac_workspace_bytes(const b200rl_sac_actor* actor, const b200rl_sac_critic* critic, int32_t batch_size);
/* update_times x AgentSAC.update_objectives (AgentSAC.py:42-86; no PER, lambda_fit_cum_r = 0): ReplayBuffer.sample
 * (:120-134, next state = next time row), q_label from the actor and the target ensemble, critic step, soft target update,
 * temperature step, actor step through the target ensemble -- including the reference's quirks (log-prob at the mean,
 * 1.000001 - tanh^2, alpha taken after its Adam step and before the clamp).  cur_size: valid time rows of the buffer.
 * ids: device int64 [update_times, batch_size] in [0, (cur_size - 1) * num_seqs) or NULL (Philox); eps_next / eps_pg: the two
 * rsample draws per update, device fp32 [update_times, batch_size, A], or NULL (Philox).  alpha_log: device fp32 [1] with
 * its one-tensor Adam group.  out_scalars (device float[2]) = means of (obj_critic, obj_actor) (AgentBase.py:172-189). */
B200RL_API int b200rl_sac_update(const b200rl_sac_actor* actor, const b200rl_sac_critic* critic,
                                 const b200rl_sac_critic* critic_target, b200rl_param_group* actor_group,
                                 b200rl_param_group* critic_group, b200rl_param_group* alpha_group,
                                 const b200rl_replay_buffer* buffer, int32_t cur_size, const b200rl_sac_hyper* hyper,
                                 int32_t batch_size, int32_t update_times, const int64_t* ids, const float* eps_next,
                                 const float* eps_pg, uint64_t seed, uint64_t draw_offset, float* out_scalars,
                                 void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200RL_H_ */
