// PPO minibatch update (any depth / widths): gather -> critic & actor forward -> losses -> backward ->
// per-net grad-norm clip -> Adam, one kernel launch per minibatch, no host synchronisation.
//
// Replaces reference AgentPPO.update_objectives (elegantrl/agents/AgentPPO.py:173-205) and
// AgentBase.optimizer_backward (elegantrl/agents/AgentBase.py:239-248: zero_grad, backward,
// clip_grad_norm_, Adam.step -- per net, hence two independent norms).
//
// Grid: one CTA per tile of 32 sampled transitions.  Each CTA accumulates its weight-gradient tile products
// into the flat fp32 gradient buffer of the workspace with RED.ADD; the CTA that takes the last ticket
// ("last block done") computes the two grad norms, clips, applies Adam in place on the caller's parameter /
// moment tensors, and re-zeroes the buffer for the next minibatch.  Sharded (multi-GPU) callers stop after the
// gradient phase (b200rl_ppo_grads), all-reduce the flat buffer, then run b200rl_ppo_apply.
#include <cooperative_groups.h>
#include <stddef.h>
#include <stdlib.h>
#include <algorithm>
#include <string.h>

#include "mlp_grad.cuh"
#include "update_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kUpdThreads = 256;
constexpr int UTB = 32;  // samples per tile
constexpr int kMaxGradCtas = 296;  // per net: 2 x 148 SMs; larger minibatches walk several tiles per CTA
constexpr int kMultiTileMin = 148; // from this many tiles on, the 128-register instantiation (2 CTAs per SM) is used: measured
                                   // 0.42 vs 0.51 ms per update_net at 4 x 8 192 samples, 1.64 vs 2.69 ms at 4 x 65 536
using UT = SmemTile<UTB>;

// Gradient phase of one (sample tile, net): gather -> forward -> loss -> backward, RED.ADD into `grads` (flat buffer
// of BOTH nets), loss sums into A.loss_sums.  WM: where the parameters are read from (mlp_tile.cuh); with W_SMEM the
// net's parameters are first staged into shared memory with coalesced L2-coherent loads (small nets).
#ifdef B200RL_PROFILE_PHASES
#define PHASE_MARK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(A.hdr)[8 + (i)] = clock64(); } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif

template <int WM>
DEV void grads_phase(const UpdateArgs& A, int tile, int ni, const int64_t* ids, uint64_t draw, float* grads, float* smem,
                     int64_t* s_tn, bool stage = true, float* gacc = nullptr) {
    float* s_um_mean = reinterpret_cast<float*>(s_tn + UTB);  // one extra scalar behind the index array
    const int H = A.buf.horizon_len, N = A.buf.num_envs;
    // packed mode (horizon_len == 0): `states` points at records {state[S], action[A], unmask, logprob, advantage
    // (already normalised), reward_sum} written by b200rl_pack_minibatches (and all-gathered across ranks); id = record
    const bool packed = (H == 0);
    const int rec_act = A.net[0].dims[0], rec_tail = (A.net[0].dims[0] + A.net[0].dims[A.net[0].num_linear] + 3) & ~3;
    const int rec = rec_tail + 4;
    const int slot0 = tile * UTB;
    float* sc = smem + A.smem_scalar_off;  // per-sample scalars
    float* s_unmask = sc, *s_logp = sc + UTB, *s_adv = sc + 2 * UTB, *s_rsum = sc + 3 * UTB, *s_act = sc + 4 * UTB;

    // ---- parameters first: their staging loads overlap the dependent load chain of the index gather below
    const b200rl_net& net = A.net[ni];
    const int L = net.num_linear, S = net.dims[0], OUT = net.dims[L];
    const float* Wl[B200RL_MAX_LINEAR];
    const float* bl[B200RL_MAX_LINEAR];
    const float* std_log = net.action_std_log;
    if (WM == W_SMEM) {
        // warps 1..7 stage; warp 0 goes straight to the index gather (a warp issues in order: it would otherwise sit
        // on its own staging loads before starting the gather's dependent chain)
        float* w = smem + A.smem_weight_off;
        const int stid = (int)threadIdx.x - 32, snt = kUpdThreads - 32;
        for (int l = 0; l < L; ++l) {
            const int J = net.dims[l + 1], K = net.dims[l];
            if (stid >= 0 && stage) stage_weight(net.weight[l], J, K, w, stid, snt);
            Wl[l] = w; w += (J * K + 3) & ~3;
            if (stid >= 0 && stage) for (int i = stid; i < J; i += snt) w[i] = __ldcg(net.bias[l] + i);
            bl[l] = w; w += (J + 3) & ~3;
        }
        if (net.action_std_log) {
            if (stid >= 0 && stage) for (int i = stid; i < OUT; i += snt) w[i] = __ldcg(net.action_std_log + i);
            std_log = w;
        }
    } else {
        for (int l = 0; l < L; ++l) { Wl[l] = net.weight[l]; bl[l] = net.bias[l]; }
    }
    PHASE_MARK(0);
    // ---- gather (reference :178-187): ids -> (t = id % H, n = id / H)
    if (threadIdx.x < UTB) {
        const int b = threadIdx.x, slot = slot0 + b;
        int64_t tn = -1;
        float um = 0.f, lp = 0.f, adv = 0.f, rs = 0.f;
        if (slot < A.local_batch) {
            if (packed) {
                tn = ids ? ids[slot] : (int64_t)draw * A.local_batch + slot;  // default: minibatch u = records [u*B, (u+1)*B)
                const float4 tail = *reinterpret_cast<const float4*>(A.buf.states + tn * rec + rec_tail);
                um = tail.x; lp = tail.y; adv = tail.z; rs = tail.w;
            } else {
                int64_t id = ids ? ids[slot] : sample_index(A.seed, draw, (uint32_t)slot, (uint64_t)H * (uint64_t)N);
                int64_t t = id % H, n = id / H;
                tn = t * N + n;
                um = A.buf.unmasks[tn] ? 1.0f : 0.0f;
                lp = A.buf.logprobs[tn];
                adv = A.buf.advantages[tn];
                if (A.buf.adv_stats) adv = (adv - A.buf.adv_stats[0]) / (A.buf.adv_stats[1] + 1e-5f);
                rs = A.buf.reward_sums[tn];
            }
        }
        s_tn[b] = tn; s_unmask[b] = um; s_logp[b] = lp; s_adv[b] = adv; s_rsum[b] = rs;
        if (A.hp.flags & B200RL_PPO_CRITIC_MASK_MEAN) {
            // helloworld's critic weight: the MEAN of unmask over the whole minibatch (its [B] x [B, 1] broadcast)
            float s = 0.0f;
            for (int sl = b; sl < A.local_batch; sl += UTB) {
                if (packed) {
                    const int64_t r = ids ? ids[sl] : (int64_t)draw * A.local_batch + sl;
                    s += A.buf.states[r * rec + rec_tail];
                } else {
                    const int64_t id = ids ? ids[sl] : sample_index(A.seed, draw, (uint32_t)sl, (uint64_t)H * (uint64_t)N);
                    s += A.buf.unmasks[(id % H) * N + id / H] ? 1.0f : 0.0f;
                }
            }
            s = warp_sum(s);
            if (b == 0) *s_um_mean = s / (float)A.local_batch;
        }
    }
    __syncthreads();
    const bool discrete = A.buf.discrete_actions != 0;
    {
        // discrete (AgentDiscretePPO): ONE int32 action index per transition, kept as a float in action slot 0
        const int Adim = discrete ? 1 : A.net[0].dims[A.net[0].num_linear];
        for (int idx = threadIdx.x; idx < UTB * Adim; idx += kUpdThreads) {
            int b = idx / Adim, a = idx - b * Adim;
            float v = 0.0f;
            if (s_tn[b] >= 0) {
                if (packed) v = A.buf.states[s_tn[b] * rec + rec_act + a];
                else if (discrete) v = (float)reinterpret_cast<const int32_t*>(A.buf.actions)[s_tn[b]];
                else v = A.buf.actions[s_tn[b] * Adim + a];
            }
            s_act[a * UTB + b] = v;
        }
    }
    const float inv_bsz = 1.0f / (float)A.global_batch;
    float loss_c = 0.f, loss_s = 0.f, loss_e = 0.f;  // valid in threads < UTB

    PHASE_MARK(1);
    // smem map: X[0..L-1] (inputs of each Linear), G[1..L-1] (act' at each hidden layer), dzA, dzB
    int xoff[B200RL_MAX_LINEAR + 1], goff[B200RL_MAX_LINEAR + 1];
    int off = 0;
    for (int l = 0; l < L; ++l) { xoff[l] = off; off += net.dims[l] * UTB; }
    for (int l = 1; l < L; ++l) { goff[l] = off; off += net.dims[l] * UTB; }
    float* dzA = smem + off;
    float* dzB = dzA + A.maxdim * UTB;

    // gather + state_norm into X[0]
    for (int idx = threadIdx.x; idx < UTB * S; idx += kUpdThreads) {
        int b = idx / S, k = idx - b * S;
        float v = 0.0f;
        if (s_tn[b] >= 0) {
            v = packed ? A.buf.states[s_tn[b] * rec + k] : A.buf.states[s_tn[b] * S + k];
            if (net.state_avg) v = (v - net.state_avg[k]) / (net.state_std[k] + 1e-4f);
        }
        smem[xoff[0] + UT::elem(k, b)] = v;
    }
    __syncthreads();
    PHASE_MARK(2);
    // forward, keeping every layer input and act'
    for (int l = 0; l < L; ++l) {
        const bool hidden = l < L - 1;
        linear_forward<UTB, kUpdThreads, WM>(Wl[l], bl[l], net.dims[l], net.dims[l + 1], smem + xoff[l],
                                                   hidden ? smem + xoff[l + 1] : dzA, hidden ? smem + goff[l + 1] : nullptr,
                                                   net.activation, hidden);
        __syncthreads();
        PHASE_MARK(3 + l);
    }
    // loss and d loss / d output, in place in dzA
    float* g = gacc ? gacc : grads + A.grad_off[ni];
    const bool atomic = gacc == nullptr;
    if (threadIdx.x < UTB) {
        const int b = threadIdx.x;
        const bool valid = s_tn[b] >= 0;
        const float um = s_unmask[b];
        const int flags = A.hp.flags;
        if (ni == 1) {
            // obj_critic = mean(criterion(V(s), reward_sum) * unmask); criterion = MSE (:189-190) or, with
            // B200RL_PPO_SMOOTH_L1, SmoothL1Loss(beta = 1) (helloworld_PPO_single_file.py:246, 332)
            float err = dzA[UT::elem(0, b)] - s_rsum[b];
            float l, dl;
            if (flags & B200RL_PPO_SMOOTH_L1) {
                const float ae = fabsf(err);
                l = ae < 1.0f ? 0.5f * err * err : ae - 0.5f;
                dl = ae < 1.0f ? err : copysignf(1.0f, err);
            } else {
                l = err * err;
                dl = 2.0f * err;
            }
            const float um_c = (flags & B200RL_PPO_CRITIC_MASK_MEAN) ? *s_um_mean : um;
            loss_c = valid ? l * um_c : 0.0f;
            dzA[UT::elem(0, b)] = valid ? dl * um_c * inv_bsz : 0.0f;
        } else {
            // new_logprob, ratio, clip, entropy                           (:193-203; helloworld :335-340)
            float logp = 0.0f, ent = 0.0f, lse = 0.0f;
            const int act_idx = discrete ? min(max((int)s_act[b], 0), OUT - 1) : 0;
            if (discrete) {
                // Categorical(softmax(logits)): logp = log p[a], entropy = -sum p log p   (reference :415-421)
                float m = -INFINITY, sum = 0.0f;
                for (int a = 0; a < OUT; ++a) m = fmaxf(m, dzA[UT::elem(a, b)]);
                for (int a = 0; a < OUT; ++a) sum += expf(dzA[UT::elem(a, b)] - m);
                lse = m + logf(sum);
                for (int a = 0; a < OUT; ++a) {
                    const float lp = dzA[UT::elem(a, b)] - lse;
                    ent -= expf(lp) * lp;
                }
                logp = dzA[UT::elem(act_idx, b)] - lse;
            } else {
                for (int a = 0; a < OUT; ++a) {
                    float sd = expf(ld_param<WM>(std_log + a));
                    float diff = s_act[a * UTB + b] - dzA[UT::elem(a, b)];
                    float lsd = logf(sd);
                    logp += -(diff * diff) / (2.0f * (sd * sd)) - lsd - kLogSqrt2Pi;
                    ent += 0.5f + kLogSqrt2Pi + lsd;  // 0.5 + 0.5 log(2 pi) + log(scale)
                }
            }
            const float ratio = expf(logp - s_logp[b]);
            const float adv = s_adv[b];
            const float um_a = (flags & B200RL_PPO_ACTOR_UNMASKED) ? 1.0f : um;  // helloworld does not mask the actor terms
            float surr, dsurr_dratio;  // surrogate and its derivative w.r.t. ratio
            if (flags & B200RL_PPO_MIN_CLIP) {
                // min(adv * ratio, adv * clamp(ratio, 1 - c, 1 + c))   (helloworld :337-339)
                const float rc = fminf(fmaxf(ratio, 1.0f - A.hp.ratio_clip), 1.0f + A.hp.ratio_clip);
                const float s1 = adv * ratio, s2 = adv * rc;
                const bool take1 = s1 <= s2;  // torch.min passes the gradient to the smaller operand (first on ties)
                surr = take1 ? s1 : s2;
                dsurr_dratio = take1 ? adv : ((rc == ratio) ? adv : 0.0f);
            } else {
                // the reference's constant-factor "clip": adv * ratio * where(adv > 0, 1 - c, 1 + c)   (:196-199)
                const float kappa = adv > 0.0f ? 1.0f - A.hp.ratio_clip : 1.0f + A.hp.ratio_clip;
                surr = adv * ratio * kappa;
                dsurr_dratio = adv * kappa;
            }
            float dlogp_scale = ratio;  // d ratio / d new_logprob
            if (flags & B200RL_PPO_A2C) {
                // AgentA2C (reference AgentPPO.py:309): (advantage [B,1] * new_logprob [B,A]).mean() = mean_b(adv * logp) / A
                surr = adv * logp / (float)OUT;
                dsurr_dratio = adv / (float)OUT;
                dlogp_scale = 1.0f;
            }
            loss_s = valid ? surr * um_a : 0.0f;
            loss_e = valid ? ent * um_a : 0.0f;
            // loss = -(obj_surrogate -/+ lambda_entropy * obj_entropy): the reference subtracts the entropy term
            // (:203-204), helloworld adds it (:340)
            const float ent_sign = (flags & B200RL_PPO_ENTROPY_BONUS) ? -1.0f : 1.0f;  // sign of d loss / d entropy
            const float gl = valid ? -(dsurr_dratio * dlogp_scale * um_a) * inv_bsz : 0.0f;  // d loss / d new_logprob
            const float ge = valid ? ent_sign * A.hp.lambda_entropy * um_a * inv_bsz : 0.0f;  // d loss / d entropy
            if (discrete) {
                // d logp[a*] / d z_j = [j == a*] - p_j;   d entropy / d z_j = -p_j (log p_j + entropy)
                for (int a = 0; a < OUT; ++a) {
                    const float lp = dzA[UT::elem(a, b)] - lse, pa = expf(lp);
                    dzA[UT::elem(a, b)] = gl * ((a == act_idx ? 1.0f : 0.0f) - pa) - ge * pa * (lp + ent);
                }
            } else {
                for (int a = 0; a < OUT; ++a) {
                    float sd = expf(ld_param<WM>(std_log + a));
                    float var = sd * sd;
                    float diff = s_act[a * UTB + b] - dzA[UT::elem(a, b)];
                    dzA[UT::elem(a, b)] = gl * diff / var;
                    float dstd = gl * (diff * diff / var - 1.0f) + ge;
                    dstd = warp_sum(dstd);  // UTB == 32: exactly warp 0
                    if (b == 0) { if (atomic) atomicAdd(g + A.grad_numel[0] - OUT + a, dstd); else g[A.grad_numel[0] - OUT + a] += dstd; }
                }
            }
        }
    }
    __syncthreads();
    PHASE_MARK(6);
    // backward
    float* dz = dzA;
    float* dzn = dzB;
    int goff_w = 0;  // float offset of W_l inside this net's flat gradient
    int woff[B200RL_MAX_LINEAR];
    for (int l = 0; l < L; ++l) { woff[l] = goff_w; goff_w += net.dims[l + 1] * net.dims[l] + net.dims[l + 1]; }
    for (int l = L - 1; l >= 0; --l) {
        const int J = net.dims[l + 1], K = net.dims[l];
        weight_grad<UTB, kUpdThreads>(dz, smem + xoff[l], J, K, g + woff[l], g + woff[l] + J * K, atomic);
        if (l > 0) data_grad<UTB, kUpdThreads, WM>(Wl[l], dz, smem + goff[l], dzn, J, K);
        __syncthreads();
        PHASE_MARK(7 + (L - 1 - l));
        float* t = dz; dz = dzn; dzn = t;
    }

    // ---- loss sums (means over the global batch) -> double accumulators
    if (threadIdx.x < 32) {
        float c = warp_sum(loss_c), s = warp_sum(loss_s), e = warp_sum(loss_e);
        if (threadIdx.x == 0) {
            if (ni == 1) atomicAdd(A.loss_sums + 0, (double)(c * inv_bsz));
            else {
                atomicAdd(A.loss_sums + 1, (double)(s * inv_bsz));
                atomicAdd(A.loss_sums + 2, (double)(e * inv_bsz));
            }
        }
    }
}

// one launch per minibatch: grid = (sample tiles, 2 nets); last-block-done apply per net
template <bool MULTI>
__global__ void __launch_bounds__(kUpdThreads) ppo_grads_kernel(const __grid_constant__ UpdateArgs A) {
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    __shared__ float red[32];
    __shared__ int64_t s_tn[UTB + 1];  // t * N + n of each sample (+ one scalar slot)
    __shared__ int s_last;
    const int ni = blockIdx.y;  // the critic and actor updates are disjoint (reference :189-204): concurrent CTAs
    // Large minibatches: the grid is capped and every CTA walks several 32-sample tiles, summing its weight gradients
    // in shared memory (plain adds) before ONE RED.ADD per element -- instead of one per tile and element.
    // (MULTI is a separate instantiation so that the one-tile-per-CTA kernel keeps its register allocation)
    if (MULTI) {
        const int tiles = (A.local_batch + UTB - 1) / UTB;
        float* gacc = smem + A.smem_gacc_off;
        for (int i = threadIdx.x; i < A.grad_numel[ni]; i += kUpdThreads) gacc[i] = 0.0f;
        __syncthreads();
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            const bool first = tile == (int)blockIdx.x;
            if (A.stage_weights) grads_phase<W_SMEM>(A, tile, ni, A.ids, A.draw, A.grads, smem, s_tn, first, gacc);
            else grads_phase<W_LDG>(A, tile, ni, A.ids, A.draw, A.grads, smem, s_tn, first, gacc);
        }
        __syncthreads();
        float* g = A.grads + A.grad_off[ni];
        for (int i = threadIdx.x; i < A.grad_numel[ni]; i += kUpdThreads) atomicAdd(g + i, gacc[i]);
    } else {
        if (A.stage_weights) grads_phase<W_SMEM>(A, blockIdx.x, ni, A.ids, A.draw, A.grads, smem, s_tn);
        else grads_phase<W_LDG>(A, blockIdx.x, ni, A.ids, A.draw, A.grads, smem, s_tn);
    }
    if (!A.fused_apply) return;

    // ---- last block done (per net): clip + Adam for this net, then re-zero its slice of the gradient buffer
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&A.hdr->ticket[ni], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    apply_net<kUpdThreads>(A.net[ni], A.opt[ni], A.adam[ni], A.grads + A.grad_off[ni], A.grad_numel[ni], A.hp.clip_grad_norm, red);
    __syncthreads();
    for (int i = threadIdx.x; i < A.grad_numel[ni]; i += kUpdThreads) A.grads[A.grad_off[ni] + i] = 0.0f;
    if (threadIdx.x == 0) A.hdr->ticket[ni] = 0u;
}

// ALL minibatches of one update_net in ONE launch: the grid is a single thread-block cluster (2 nets x sample tiles
// <= 8 CTAs), minibatches are separated by hardware cluster barriers instead of kernel boundaries, the clip + Adam of
// each net is shared by that net's CTAs, gradient buffers alternate so that re-zeroing overlaps the next minibatch.
__global__ void __launch_bounds__(kUpdThreads) ppo_update_cluster_kernel(const __grid_constant__ UpdateArgs A) {
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    __shared__ float red[32];
    __shared__ int64_t s_tn[UTB + 1];
    __shared__ AdamScalars s_adam;
    cg::cluster_group cluster = cg::this_cluster();
    const int tiles = gridDim.x >> 1;
    const int ni = blockIdx.x / tiles, tile = blockIdx.x - ni * tiles;
    const int gtotal = A.grad_off[1] + A.grad_numel[1];

    for (int u = 0; u < A.update_times; ++u) {
        float* gcur = A.grads + (u & 1) * A.grad_stride;
        float* gnext = A.grads + ((u + 1) & 1) * A.grad_stride;
        if (threadIdx.x == 0) {  // torch.optim.Adam bias corrections of this step, in double like torch
            const double step = (double)(A.opt[ni].step + u + 1);
            s_adam.step_size = (float)((double)A.opt[ni].lr / (1.0 - pow((double)A.opt[ni].beta1, step)));
            s_adam.bc2_sqrt = (float)sqrt(1.0 - pow((double)A.opt[ni].beta2, step));
        }
        const int64_t* ids_u = A.ids ? A.ids + (size_t)u * A.local_batch : nullptr;
        if (A.stage_weights) grads_phase<W_SMEM>(A, tile, ni, ids_u, A.draw + (uint64_t)u, gcur, smem, s_tn);
        else grads_phase<W_LDCG>(A, tile, ni, ids_u, A.draw + (uint64_t)u, gcur, smem, s_tn);
        PHASE_MARK(10);
        __threadfence();
        cluster.sync();
        PHASE_MARK(11);
        // this net's CTAs share its clip + Adam; they also re-zero the other gradient buffer for minibatch u + 1
        apply_net<kUpdThreads>(A.net[ni], A.opt[ni], s_adam, gcur + A.grad_off[ni], A.grad_numel[ni], A.hp.clip_grad_norm, red, tile, tiles);
        PHASE_MARK(12);
        for (int i = blockIdx.x * kUpdThreads + threadIdx.x; i < gtotal; i += gridDim.x * kUpdThreads) gnext[i] = 0.0f;
        __threadfence();
        cluster.sync();
        PHASE_MARK(13);
    }
    if (blockIdx.x == 0 && threadIdx.x < 3)
        A.out_scalars[threadIdx.x] = (float)(__ldcg(A.loss_sums + threadIdx.x) / (double)A.update_times);
}

__global__ void __launch_bounds__(kUpdThreads) ppo_apply_kernel(const __grid_constant__ UpdateArgs A) {
    __shared__ float red[32];
    const int ni = blockIdx.x;  // one CTA per net
    apply_net<kUpdThreads>(A.net[ni], A.opt[ni], A.adam[ni], A.grads + A.grad_off[ni], A.grad_numel[ni], A.hp.clip_grad_norm, red);
}

// One thread per sampled transition: draw / read its index, gather the six fields, normalise the advantage, and write a
// 16-byte aligned record {state[S], action[A], pad to a multiple of 4, unmask, logprob, advantage, reward_sum}.
__global__ void pack_minibatches_kernel(const b200rl_train_buffer buf, int S, int Adim, int local_batch, int update_times,
                                        const int64_t* ids, uint64_t seed, uint64_t draw_offset, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= local_batch * update_times) return;
    const int u = i / local_batch, slot = i - u * local_batch;
    const int H = buf.horizon_len, N = buf.num_envs;
    const int64_t id = ids ? ids[i] : sample_index(seed, draw_offset + (uint64_t)u, (uint32_t)slot, (uint64_t)H * (uint64_t)N);
    const int64_t t = id % H, n = id / H, tn = t * N + n;
    const int tail = (S + Adim + 3) & ~3;
    float* r = out + (size_t)i * (tail + 4);
    for (int k = 0; k < S; ++k) r[k] = buf.states[tn * S + k];
    if (buf.discrete_actions) {  // one int32 index per transition -> float in slot 0, the other action slots stay zero
        r[S] = (float)reinterpret_cast<const int32_t*>(buf.actions)[tn];
        for (int k = S + 1; k < tail; ++k) r[k] = 0.0f;
    } else {
        for (int a = 0; a < Adim; ++a) r[S + a] = buf.actions[tn * Adim + a];
        for (int k = S + Adim; k < tail; ++k) r[k] = 0.0f;
    }
    float adv = buf.advantages[tn];
    if (buf.adv_stats) adv = (adv - buf.adv_stats[0]) / (buf.adv_stats[1] + 1e-5f);
    r[tail + 0] = buf.unmasks[tn] ? 1.0f : 0.0f;
    r[tail + 1] = buf.logprobs[tn];
    r[tail + 2] = adv;
    r[tail + 3] = buf.reward_sums[tn];
}

__global__ void loss_means_kernel(const double* loss_sums, double inv_updates, float* out) {
    if (threadIdx.x < 3) out[threadIdx.x] = (float)(loss_sums[threadIdx.x] * inv_updates);
}

AdamScalars adam_scalars(const b200rl_adam* opt, int64_t step) {
    // torch.optim.Adam (_single_tensor_adam): step_size = lr / (1 - beta1^t); sqrt(1 - beta2^t) -- in double
    double bc1 = 1.0 - pow((double)opt->beta1, (double)step);
    double bc2 = 1.0 - pow((double)opt->beta2, (double)step);
    AdamScalars s;
    s.step_size = (float)((double)opt->lr / bc1);
    s.bc2_sqrt = (float)sqrt(bc2);
    return s;
}

int fill_args(UpdateArgs& A, const b200rl_net* actor, const b200rl_net* critic, const b200rl_adam* actor_opt,
              const b200rl_adam* critic_opt, const b200rl_train_buffer* buffer, const b200rl_ppo_hyper* hyper,
              void* workspace, int64_t workspace_bytes, size_t* smem_bytes) {
    // (b200rl_ppo_apply passes no buffer: there the actor's own layout decides)
    const bool discrete = buffer ? buffer->discrete_actions != 0 : (actor && actor->action_std_log == nullptr);
    if (int rc = b200rl_validate_net(actor, "ppo.actor", !discrete)) return rc;
    B200RL_REQUIRE(!discrete || actor->action_std_log == nullptr,
                   "ppo: discrete_actions needs a categorical actor (action_std_log must be NULL)");
    if (int rc = b200rl_validate_net(critic, "ppo.critic", false)) return rc;
    B200RL_REQUIRE(hyper && workspace, "ppo: hyper/workspace is NULL");
    B200RL_REQUIRE(critic->dims[critic->num_linear] == 1, "ppo: critic output dim must be 1");
    B200RL_REQUIRE(actor->dims[0] == critic->dims[0], "ppo: actor/critic state_dim differ");
    B200RL_REQUIRE(workspace_bytes >= b200rl_workspace_bytes(actor, critic), "ppo: workspace too small (%lld < %lld)",
                   (long long)workspace_bytes, (long long)b200rl_workspace_bytes(actor, critic));
    B200RL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "ppo: workspace must be 256-byte aligned");
    A.net[0] = *actor; A.net[1] = *critic;
    if (actor_opt) A.opt[0] = *actor_opt;
    if (critic_opt) A.opt[1] = *critic_opt;
    if (buffer) A.buf = *buffer;
    A.hp = *hyper;
    A.hdr = reinterpret_cast<WorkspaceHeader*>(workspace);
    A.grads = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + B200RL_WS_HEADER_BYTES);
    A.grad_numel[0] = (int)b200rl_net_numel(actor);
    A.grad_numel[1] = (int)b200rl_net_numel(critic);
    A.grad_off[0] = 0;
    A.grad_off[1] = (A.grad_numel[0] + 3) & ~3;  // 16-byte aligned start of the critic segment
    // dynamic smem: max over nets of (sum of Linear input dims + sum of hidden dims) + 2 * maxdim, + scalars
    int maxdim = 0, rows = 0;
    for (int ni = 0; ni < 2; ++ni) {
        const b200rl_net& n = A.net[ni];
        int r = 0;
        for (int l = 0; l < n.num_linear; ++l) r += n.dims[l];
        for (int l = 1; l < n.num_linear; ++l) r += n.dims[l];
        rows = r > rows ? r : rows;
        int md = b200rl_net_maxdim(&n);
        maxdim = md > maxdim ? md : maxdim;
    }
    A.maxdim = maxdim;
    A.smem_scalar_off = (rows + 2 * maxdim) * UTB;
    int scalars = (4 + actor->dims[actor->num_linear]) * UTB;
    // small nets: + a shared-memory copy of the net's parameters (rows padded to 16 bytes)
    int wfloats = 0;
    for (int ni = 0; ni < 2; ++ni) {
        const b200rl_net& n = A.net[ni];
        int w = 0;
        for (int l = 0; l < n.num_linear; ++l) w += ((n.dims[l + 1] * n.dims[l] + 3) & ~3) + ((n.dims[l + 1] + 3) & ~3);
        w += (n.dims[n.num_linear] + 3) & ~3;
        wfloats = w > wfloats ? w : wfloats;
    }
    A.smem_weight_off = (A.smem_scalar_off + scalars + 3) & ~3;
    A.stage_weights = (wfloats <= 16 * 1024 && (size_t)(A.smem_weight_off + wfloats) * sizeof(float) <= 200 * 1024) ? 1 : 0;
    int total = A.stage_weights ? A.smem_weight_off + wfloats : A.smem_scalar_off + scalars;
    // per-CTA gradient accumulator for multi-tile CTAs (only when it leaves room for >= 2 CTAs per SM)
    const int gmax = (int)(b200rl_net_numel(actor) > b200rl_net_numel(critic) ? b200rl_net_numel(actor) : b200rl_net_numel(critic));
    total = (total + 3) & ~3;
    if ((size_t)(total + gmax) * sizeof(float) <= 100 * 1024) { A.smem_gacc_off = total; total += gmax; }
    else A.smem_gacc_off = -1;
    *smem_bytes = (size_t)total * sizeof(float);
    B200RL_REQUIRE(*smem_bytes <= 227 * 1024, "ppo: nets too wide for the update kernel (%zu B of shared memory needed)",
                   *smem_bytes);
    return 0;
}

int check_buffer(const b200rl_train_buffer* b) {
    B200RL_REQUIRE(b && b->states, "ppo: NULL training buffer");
    if (b->horizon_len == 0) {  // packed records (b200rl_pack_minibatches)
        B200RL_REQUIRE(b->num_envs >= 1, "ppo: packed buffer with %d records", b->num_envs);
        return 0;
    }
    B200RL_REQUIRE(b->actions && b->unmasks && b->logprobs && b->advantages && b->reward_sums, "ppo: NULL training buffer field");
    B200RL_REQUIRE(b->horizon_len >= 1 && b->num_envs >= 1, "ppo: horizon_len=%d num_envs=%d", b->horizon_len, b->num_envs);
    return 0;
}

}  // namespace

extern "C" {

int b200rl_ppo_update(const b200rl_net* actor, const b200rl_net* critic, b200rl_adam* actor_opt, b200rl_adam* critic_opt,
                      const b200rl_train_buffer* buffer, const b200rl_ppo_hyper* hyper, int32_t batch_size,
                      int32_t update_times, const int64_t* ids, uint64_t seed, uint64_t draw_offset, float* out_scalars,
                      void* workspace, int64_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200RL_REQUIRE(actor_opt && critic_opt && out_scalars, "ppo_update: NULL optimizer / output");
    B200RL_REQUIRE(batch_size >= 1 && update_times >= 1, "ppo_update: batch_size=%d update_times=%d", batch_size, update_times);
    if (int rc = check_buffer(buffer)) return rc;
    UpdateArgs A{};
    size_t smem = 0;
    if (int rc = fill_args(A, actor, critic, actor_opt, critic_opt, buffer, hyper, workspace, workspace_bytes, &smem)) return rc;
    A.loss_sums = A.hdr->loss_sums;
    A.fused_apply = 1;
    A.local_batch = A.global_batch = batch_size;
    A.seed = seed;
    B200RL_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)b200rl_workspace_bytes(actor, critic), stream));
    const int tiles = (batch_size + UTB - 1) / UTB;
    // B200RL_UPDATE: unset / "tc" -> the tcgen05 kernels (update_tc.cu) when the nets have their shape, else the generic ones;
    // "cluster" / "multilaunch" -> the generic FP32-pipe kernels below (kept as the cross-check; the tests run all three)
    const char* mode = getenv("B200RL_UPDATE");
    const bool want_multi = mode && strcmp(mode, "multilaunch") == 0;
    const bool want_generic = want_multi || (mode && strcmp(mode, "cluster") == 0);
    if (!want_generic && b200rl_update_tc_eligible(actor, critic, hyper)) {
        const int tiles128 = (batch_size + 127) / 128;
        A.ids = ids;
        A.draw = draw_offset;
        if (tiles128 == 1) {   // the whole update_net loop as ONE persistent launch (two CTAs that never meet)
            A.update_times = update_times;
            A.out_scalars = out_scalars;
            if (int rc = b200rl_launch_update_tc(A, 1, stream)) return rc;
            B200RL_COUNT_LAUNCH(1);
        } else {
            for (int u = 0; u < update_times; ++u) {
                A.ids = ids ? ids + (size_t)u * batch_size : nullptr;
                A.draw = draw_offset + (uint64_t)u;
                A.adam[0] = adam_scalars(actor_opt, actor_opt->step + u + 1);
                A.adam[1] = adam_scalars(critic_opt, critic_opt->step + u + 1);
                if (int rc = b200rl_launch_update_tc(A, tiles128, stream)) return rc;
            }
            loss_means_kernel<<<1, 32, 0, stream>>>(A.loss_sums, 1.0 / (double)update_times, out_scalars);
            B200RL_COUNT_LAUNCH(update_times + 1);
            B200RL_CHECK_CUDA(cudaGetLastError());
        }
        actor_opt->step += update_times;
        critic_opt->step += update_times;
        return 0;
    }
    if (2 * tiles <= 8 && !want_multi) {
        // small minibatches (the Config default 128): the whole update_net loop as ONE persistent cluster launch
        A.update_times = update_times;
        A.grad_stride = (int)((b200rl_grad_numel(actor, critic) + 63) & ~(int64_t)63);
        A.out_scalars = out_scalars;
        A.ids = ids;
        A.draw = draw_offset;
        B200RL_CHECK_CUDA(cudaFuncSetAttribute(ppo_update_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * tiles);
        cfg.blockDim = dim3(kUpdThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2 * tiles;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        B200RL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ppo_update_cluster_kernel, A));
        B200RL_COUNT_LAUNCH(1);
    } else {
        const bool multi = A.smem_gacc_off >= 0 && tiles >= kMultiTileMin;
        auto kern = multi ? ppo_grads_kernel<true> : ppo_grads_kernel<false>;
        B200RL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const dim3 grid((unsigned)(multi ? std::min(tiles, kMaxGradCtas) : tiles), 2);
        for (int u = 0; u < update_times; ++u) {
            A.ids = ids ? ids + (size_t)u * batch_size : nullptr;
            A.draw = draw_offset + (uint64_t)u;
            A.adam[0] = adam_scalars(actor_opt, actor_opt->step + u + 1);
            A.adam[1] = adam_scalars(critic_opt, critic_opt->step + u + 1);
            kern<<<grid, kUpdThreads, smem, stream>>>(A);
        }
        B200RL_COUNT_LAUNCH(update_times + 1);
        B200RL_CHECK_CUDA(cudaGetLastError());
        loss_means_kernel<<<1, 32, 0, stream>>>(A.loss_sums, 1.0 / (double)update_times, out_scalars);
    }
    B200RL_CHECK_CUDA(cudaGetLastError());
    actor_opt->step += update_times;
    critic_opt->step += update_times;
    return 0;
}

int b200rl_ppo_update_sharded(const b200rl_net* actor, const b200rl_net* critic, b200rl_adam* actor_opt, b200rl_adam* critic_opt,
                              const b200rl_train_buffer* buffer, const b200rl_ppo_hyper* hyper, int32_t batch_size,
                              int32_t update_times, const int64_t* ids, uint64_t seed, uint64_t draw_offset,
                              const double* stat_sums, int64_t count_all, int64_t count_lattice, float* adv_stats_out,
                              float* out_scalars, void* workspace, int64_t workspace_bytes, const b200rl_peer_exchange* px,
                              int32_t exchange_mode, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200RL_REQUIRE(actor_opt && critic_opt && out_scalars && px && stat_sums, "ppo_update_sharded: NULL argument");
    B200RL_REQUIRE(px->world >= 1 && px->world <= B200RL_MAX_PEERS && px->rank >= 0 && px->rank < px->world,
                   "ppo_update_sharded: rank %d of %d", px->rank, px->world);
    B200RL_REQUIRE(batch_size >= px->world && batch_size % px->world == 0 && batch_size / px->world <= 128 && update_times >= 1,
                   "ppo_update_sharded: batch_size=%d must be a multiple of world=%d with at most 128 samples per rank", batch_size,
                   px->world);
    for (int r = 0; r < px->world; ++r)
        B200RL_REQUIRE(px->data[r] && px->flags[r], "ppo_update_sharded: peer %d is not mapped", r);
    if (int rc = check_buffer(buffer)) return rc;
    B200RL_REQUIRE(buffer->horizon_len >= 1, "ppo_update_sharded: needs the [H, N] shard, not packed records");
    UpdateArgs A{};
    size_t smem = 0;
    if (int rc = fill_args(A, actor, critic, actor_opt, critic_opt, buffer, hyper, workspace, workspace_bytes, &smem)) return rc;
    B200RL_REQUIRE(b200rl_update_tc_eligible(actor, critic, hyper),
                   "ppo_update_sharded: nets must be S -> 64 -> 64 -> OUT GELU (b200rl_update_tc_supported)");
    A.loss_sums = A.hdr->loss_sums;
    A.fused_apply = 1;
    B200RL_REQUIRE(exchange_mode == 0 || exchange_mode == 1, "ppo_update_sharded: exchange_mode=%d", exchange_mode);
    B200RL_REQUIRE(exchange_mode == 0 || batch_size <= 128, "ppo_update_sharded: the record gather needs batch_size <= 128 (one tile)");
    // gradient all-reduce: each rank's tile holds its own batch_size / world samples; record gather: every rank's tile holds the
    // whole minibatch, gathered from all ranks' exchange buffers
    A.local_batch = exchange_mode == 1 ? batch_size : batch_size / px->world;
    A.px_local_batch = batch_size / px->world;
    A.global_batch = batch_size;
    A.seed = seed;
    A.ids = ids;
    A.draw = draw_offset;
    A.update_times = update_times;
    A.out_scalars = out_scalars;
    A.px_on = exchange_mode == 1 ? 2 : 1;
    A.px = *px;
    A.stat_sums = stat_sums;
    A.count_all = (double)count_all;
    A.count_lat = (double)count_lattice;
    A.stats_out = adv_stats_out;
    B200RL_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)b200rl_workspace_bytes(actor, critic), stream));
    if (int rc = b200rl_launch_update_tc(A, 1, stream)) return rc;
    B200RL_COUNT_LAUNCH(1);
    actor_opt->step += update_times;
    critic_opt->step += update_times;
    return 0;
}

int64_t b200rl_workspace_error_offset(void) { return (int64_t)offsetof(WorkspaceHeader, pad); }

int b200rl_ppo_grads(const b200rl_net* actor, const b200rl_net* critic, const b200rl_train_buffer* buffer,
                     const b200rl_ppo_hyper* hyper, int32_t local_batch, int32_t global_batch, const int64_t* ids,
                     uint64_t seed, uint64_t draw_offset, double* loss_sums, void* workspace, int64_t workspace_bytes,
                     void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200RL_REQUIRE(loss_sums, "ppo_grads: loss_sums is NULL");
    B200RL_REQUIRE(local_batch >= 1 && global_batch >= local_batch, "ppo_grads: local_batch=%d global_batch=%d", local_batch,
                   global_batch);
    if (int rc = check_buffer(buffer)) return rc;
    UpdateArgs A{};
    size_t smem = 0;
    if (int rc = fill_args(A, actor, critic, nullptr, nullptr, buffer, hyper, workspace, workspace_bytes, &smem)) return rc;
    A.loss_sums = loss_sums;
    A.fused_apply = 0;
    A.local_batch = local_batch;
    A.global_batch = global_batch;
    A.ids = ids;
    A.seed = seed;
    A.draw = draw_offset;
    const int tiles = (local_batch + UTB - 1) / UTB;
    B200RL_CHECK_CUDA(cudaMemsetAsync(A.grads, 0, (size_t)(A.grad_off[1] + A.grad_numel[1]) * sizeof(float), stream));
    {
        const char* mode = getenv("B200RL_UPDATE");
        const bool want_generic = mode && (strcmp(mode, "multilaunch") == 0 || strcmp(mode, "cluster") == 0);
        if (!want_generic && b200rl_update_tc_eligible(actor, critic, hyper)) {
            if (int rc = b200rl_launch_update_tc(A, (local_batch + 127) / 128, stream)) return rc;
            B200RL_COUNT_LAUNCH(1);
            return 0;
        }
    }
    const bool multi = A.smem_gacc_off >= 0 && tiles >= kMultiTileMin;
    auto kern = multi ? ppo_grads_kernel<true> : ppo_grads_kernel<false>;
    B200RL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)(multi ? std::min(tiles, kMaxGradCtas) : tiles), 2), kUpdThreads, smem, stream>>>(A);
    B200RL_COUNT_LAUNCH(1);
    B200RL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int b200rl_ppo_apply(const b200rl_net* actor, const b200rl_net* critic, b200rl_adam* actor_opt, b200rl_adam* critic_opt,
                     const b200rl_ppo_hyper* hyper, void* workspace, int64_t workspace_bytes, void* stream_) {
    B200RL_REQUIRE(actor_opt && critic_opt, "ppo_apply: NULL optimizer");
    UpdateArgs A{};
    size_t smem = 0;
    if (int rc = fill_args(A, actor, critic, actor_opt, critic_opt, nullptr, hyper, workspace, workspace_bytes, &smem)) return rc;
    A.adam[0] = adam_scalars(actor_opt, actor_opt->step + 1);
    A.adam[1] = adam_scalars(critic_opt, critic_opt->step + 1);
    ppo_apply_kernel<<<2, kUpdThreads, 0, (cudaStream_t)stream_>>>(A);
    B200RL_COUNT_LAUNCH(1);
    B200RL_CHECK_CUDA(cudaGetLastError());
    actor_opt->step += 1;
    critic_opt->step += 1;
    return 0;
}

int b200rl_pack_minibatches(const b200rl_train_buffer* buffer, int32_t state_dim, int32_t action_dim, int32_t local_batch,
                            int32_t update_times, const int64_t* ids, uint64_t seed, uint64_t draw_offset, float* out_records,
                            void* stream) {
    if (int rc = check_buffer(buffer)) return rc;
    B200RL_REQUIRE(buffer->horizon_len >= 1, "pack_minibatches: source buffer must not be packed itself");
    B200RL_REQUIRE(out_records && local_batch >= 1 && update_times >= 1, "pack_minibatches: bad arguments");
    const int total = local_batch * update_times;
    pack_minibatches_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*buffer, state_dim, action_dim, local_batch,
                                                                                  update_times, ids, seed, draw_offset, out_records);
    B200RL_COUNT_LAUNCH(1);
    B200RL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int b200rl_loss_means(const double* loss_sums, int32_t update_times, float* out_scalars, void* stream) {
    B200RL_REQUIRE(loss_sums && out_scalars && update_times >= 1, "loss_means: bad arguments");
    loss_means_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss_sums, 1.0 / (double)update_times, out_scalars);
    B200RL_COUNT_LAUNCH(1);
    B200RL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
