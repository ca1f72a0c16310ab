// Off-policy path (BASELINE configs[3]): ReplayBuffer ring writes and AgentSAC.update_objectives.
//
// Replaces reference elegantrl/train/replay_buffer.py:78-134 (update / sample), elegantrl/agents/AgentSAC.py:42-86
// (update_objectives with ActorSAC :167-198 and CriticEnsemble :244-259), AgentBase.optimizer_backward / soft_update
// (elegantrl/agents/AgentBase.py:239-248, 270-278).  Arithmetic restated in oracle/sac_oracle.py and pinned to goldens minted
// from the reference (tests/golden/sac_*.npz); every quirk listed there is kept.
//
// One minibatch = four launches over tiles of 32 sampled transitions (generic depth / widths, FP32 pipe, the tile routines of
// mlp_tile.cuh / mlp_grad.cuh; parameters are read L2-coherently because the last CTA of a launch rewrites them):
//   label   s' -> actor (rsample eps_next) -> target ensemble -> q_label = r + undone * gamma * (min_e Q - logp' * alpha)
//   critic  (s, a) -> ensemble forward / backward, one decoder per CTA (grid = tiles x E: the decoders are independent given
//           the shared encoder output, and the encoder's gradient is linear in their contributions) -> gradients;
//           last CTA: clip_grad_norm_ + Adam over the critic's parameter list, then the soft update of the target
//   pgrad   s -> actor (rsample eps_pg) -> target ensemble forward and DATA gradient -> d obj / d tanh(action), sums of
//           logp and Q; last CTA: temperature step (alpha taken after it, before the clamp), obj_actor
//   actor   s -> actor forward with caches (same eps_pg) -> backward -> gradients; last CTA: clip + Adam
// The reference issues ~600 tiny launches and 3 host synchronisations per update_objectives; here nothing returns to the
// host until update_net reads the two means.
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "mlp_grad.cuh"
#include "update_common.cuh"

namespace {

constexpr int kNT = 256;
constexpr int TB = 32;
using T = SmemTile<TB>;
constexpr int WM = W_LDCG;

struct SacHeader {
    double obj_critic_sum, obj_actor_sum;   // over the updates of this call
    double logp_sum, q_sum, td_sum;         // over the tiles of the running update
    unsigned int ticket[4];
    float alpha_used;                       // exp(alpha_log) after the temperature step, before the clamp
    float pad[3];
};

struct MlpShape {        // host + device view of one build_mlp
    int L;
    int dims[B200RL_MAX_LINEAR + 1];
};

struct SacArgs {
    b200rl_sac_actor actor;
    b200rl_sac_critic critic, target;
    b200rl_param_group g_actor, g_critic;
    float* target_param[B200RL_MAX_GROUP_TENSORS];   // the target ensemble's tensors in the critic group's order
    float *alpha_log, *alpha_m, *alpha_v;
    float alpha_b1, alpha_b2, alpha_eps;
    AdamScalars adam_actor, adam_critic, adam_alpha;
    b200rl_replay_buffer rb;
    int cur_size;
    b200rl_sac_hyper hp;
    int batch;
    const int64_t* ids;        // [batch] of this update or nullptr
    const float* eps_next;     // [batch, A] or nullptr
    const float* eps_pg;
    uint64_t seed, draw;
    SacHeader* hdr;
    int64_t* rows;             // [batch] sampled ring row index (time * num_seqs + seq)
    float* q_label;            // [batch]
    float* logp;               // [batch]
    float* d_tanh;             // [batch, A]
    float* grads_c;            // flat, critic group order
    float* grads_a;            // flat, actor group order
    int numel_c, numel_a;
    int maxw;                  // widest layer of any net (rows of a ping-pong buffer)
    float inv_updates;
    float* out_scalars;
    int last_update;
};

// ---- Philox streams of this path (counter.w tags; common.cuh uses 0x0 and 0x1D5)
constexpr uint32_t kStreamSacIds = 0x5AC1u, kStreamSacNext = 0x5AC2u, kStreamSacPg = 0x5AC3u, kStreamSacExplore = 0x5AC4u;
DEV uint4 sac_bits(uint64_t seed, uint64_t draw, uint32_t slot, uint32_t chunk, uint32_t stream) {
    uint4 ctr = make_uint4(slot, chunk, (uint32_t)draw, (uint32_t)(draw >> 32) ^ (stream << 16));
    return philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
DEV float sac_normal(uint64_t seed, uint64_t draw, uint32_t slot, int a, uint32_t stream) {
    const uint4 r = sac_bits(seed, draw, slot, (uint32_t)(a >> 1), stream);
    const float2 n = box_muller(r.x, r.y);
    return (a & 1) ? n.y : n.x;
}

// ---- MLP over a tile (activations feature-major in shared memory)
// forward without caches: ping-pong between bufA / bufB; returns the buffer that holds the output
DEV float* mlp_forward_nocache(const b200rl_net& net, bool act_last, const float* in, float* bufA, float* bufB) {
    const float* x = in;
    float* y = bufA;
    for (int l = 0; l < net.num_linear; ++l) {
        const bool act = l < net.num_linear - 1 || act_last;
        linear_forward<TB, kNT, WM>(net.weight[l], net.bias[l], net.dims[l], net.dims[l + 1], x, y, nullptr, net.activation, act);
        __syncthreads();
        x = y;
        y = (y == bufA) ? bufB : bufA;
    }
    return const_cast<float*>(x);
}
// forward with caches: X[l] = input of layer l (X[0] = `in`, caller-owned), G[l] = act' at the output of layer l - 1.
// cache holds, for l = 1 .. L-1 (and L when act_last): rows dims[l] of X then rows dims[l] of G.  The raw output of the last
// layer (no activation) goes to `out`.  xs / gs receive the pointers.
DEV void mlp_forward_cached(const b200rl_net& net, bool act_last, const float* in, float* cache, float* out, const float** xs,
                            const float** gs) {
    xs[0] = in;
    gs[0] = nullptr;
    float* c = cache;
    for (int l = 0; l < net.num_linear; ++l) {
        const bool act = l < net.num_linear - 1 || act_last;
        const int J = net.dims[l + 1];
        float* y = act ? c : out;
        float* gbuf = act ? c + J * TB : nullptr;
        linear_forward<TB, kNT, WM>(net.weight[l], net.bias[l], net.dims[l], J, xs[l], y, gbuf, net.activation, act);
        __syncthreads();
        xs[l + 1] = y;
        gs[l + 1] = gbuf;
        if (act) c += 2 * J * TB;
    }
}
// backward of mlp_forward_cached: dz = d loss / d (output of the last layer, AFTER its activation derivative was applied by
// the caller when act_last).  Weight gradients are RED.ADDed at g (tensor order W0, b0, W1, b1, ...) unless g == nullptr
// (data gradient only).  d_in (may be nullptr): receives / accumulates d loss / d input, without an activation derivative.
DEV void mlp_backward(const b200rl_net& net, const float** xs, const float** gs, float* dz, float* dz_other, float* g,
                      float* d_in, bool accumulate_in) {
    int woff[B200RL_MAX_LINEAR];
    int off = 0;
    for (int l = 0; l < net.num_linear; ++l) { woff[l] = off; off += net.dims[l + 1] * net.dims[l] + net.dims[l + 1]; }
    for (int l = net.num_linear - 1; l >= 0; --l) {
        const int J = net.dims[l + 1], K = net.dims[l];
        if (g) weight_grad<TB, kNT>(dz, xs[l], J, K, g + woff[l], g + woff[l] + J * K, true);
        if (l > 0) {
            data_grad<TB, kNT, WM>(net.weight[l], dz, gs[l], dz_other, J, K);
            __syncthreads();
            float* t = dz; dz = dz_other; dz_other = t;
        } else {
            if (d_in) data_grad<TB, kNT, WM>(net.weight[l], dz, nullptr, d_in, J, K, accumulate_in);
            __syncthreads();
        }
    }
}
DEV int mlp_numel(const b200rl_net& net) {
    int n = 0;
    for (int l = 0; l < net.num_linear; ++l) n += net.dims[l + 1] * net.dims[l] + net.dims[l + 1];
    return n;
}

// ---- sampling (ReplayBuffer.sample, replay_buffer.py:120-134): ids -> (time = ids % sample_len, seq = ids / sample_len)
DEV int64_t sample_row(const SacArgs& A, int slot) {
    const int64_t sample_len = A.cur_size - 1;
    int64_t id;
    if (A.ids) id = A.ids[slot];
    else {
        const uint4 r = sac_bits(A.seed, A.draw, (uint32_t)slot, 0u, kStreamSacIds);
        id = (int64_t)__umul64hi(((uint64_t)r.x << 32) | r.y, (uint64_t)(sample_len * A.rb.num_seqs));
    }
    return (id % sample_len) * A.rb.num_seqs + id / sample_len;
}
// rows [row_first, +rows) of a [*, dim] ring tensor -> feature rows [f0, f0 + dim) of a tile
DEV void gather_rows(const float* src, int dim, const int64_t* s_row, int64_t row_shift, int f0, float* Xs) {
    for (int idx = threadIdx.x; idx < TB * dim; idx += kNT) {
        const int b = idx / dim, k = idx - b * dim;
        Xs[T::elem(f0 + k, b)] = s_row[b] >= 0 ? src[(s_row[b] + row_shift) * dim + k] : 0.0f;
    }
}

// ActorSAC head on the raw [mean | log_std] rows (AgentSAC.py:184-196): one thread per sample.
// Writes tanh(action) into feature rows [f0, f0 + A) of `dst` and returns logprob (evaluated at the MEAN, tanh-corrected).
DEV float actor_head(const float* out, int A_dim, int b, const float* eps, uint64_t seed, uint64_t draw, uint32_t slot,
                     uint32_t stream, float* dst, int f0) {
    float logp = 0.0f;
    for (int a = 0; a < A_dim; ++a) {
        const float avg = out[T::elem(a, b)];
        const float lsd = fminf(fmaxf(out[T::elem(A_dim + a, b)], -16.0f), 2.0f);
        const float sd = expf(lsd);
        const float e = eps ? eps[(size_t)slot * A_dim + a] : sac_normal(seed, draw, slot, a, stream);
        const float t = tanhf(__fadd_rn(avg, __fmul_rn(sd, e)));            // Normal.rsample(): loc + eps * scale
        dst[T::elem(f0 + a, b)] = t;
        logp += (-logf(sd) - kLogSqrt2Pi) - logf(-t * t + 1.000001f);      // log_prob at the mean (:193), tanh fix (:194)
    }
    return logp;
}

// ------------------------------------------------------------------------------------------------ generic group apply
DEV void apply_group(const b200rl_param_group& grp, const AdamScalars& as, const float* g, int numel, float clip_grad_norm, float* red) {
    float ss = 0.0f;
    for (int i = threadIdx.x; i < numel; i += kNT) { const float v = __ldcg(g + i); ss = fmaf(v, v, ss); }
    const float total_norm = sqrtf(block_sum<kNT>(ss, red));
    float coef = 1.0f;
    if (clip_grad_norm > 0.0f) coef = fminf(clip_grad_norm / (total_norm + 1e-6f), 1.0f);
    int off = 0;
    for (int t = 0; t < grp.num_tensors; ++t) {
        float *P = grp.param[t], *M = grp.exp_avg[t], *V = grp.exp_avg_sq[t];
        for (int i = threadIdx.x; i < grp.numel[t]; i += kNT) {
            float p = __ldcg(P + i), m = __ldcg(M + i), v = __ldcg(V + i);
            adam_one(p, m, v, __ldcg(g + off + i) * coef, grp.beta1, grp.beta2, grp.eps, as);
            P[i] = p; M[i] = m; V[i] = v;
        }
        off += grp.numel[t];
    }
}
DEV bool last_block(unsigned int* ticket, int* s_flag) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *s_flag = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
    __syncthreads();
    if (*s_flag) __threadfence();
    return *s_flag != 0;
}

// ------------------------------------------------------------------------------------------------------------ kernels
// smem map (rows of TB floats): sa [S + A] | enc [d0] | bufA [maxw] | bufB [maxw] | scalars
__global__ void __launch_bounds__(kNT) sac_label_kernel(const __grid_constant__ SacArgs A) {
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    __shared__ int64_t s_row[TB];
    const int S = A.rb.state_dim, Ad = A.rb.action_dim, d0 = A.target.encoder.dims[1];
    float* sa = smem;
    float* enc = sa + (S + Ad) * TB;
    float* bufA = enc + d0 * TB;
    float* bufB = bufA + A.maxw * TB;
    float* s_logp = bufB + A.maxw * TB;
    float* s_minq = s_logp + TB;
    const int slot0 = blockIdx.x * TB;
    if (threadIdx.x < TB) {
        const int slot = slot0 + threadIdx.x;
        const int64_t r = slot < A.batch ? sample_row(A, slot) : -1;
        s_row[threadIdx.x] = r;
        if (slot < A.batch) A.rows[slot] = r;
    }
    __syncthreads();
    gather_rows(A.rb.states, S, s_row, A.rb.num_seqs, 0, sa);   // next state = the NEXT time row of the same sequence
    __syncthreads();
    float* h = mlp_forward_nocache(A.actor.net_s, true, sa, bufA, bufB);
    float* out = (h == bufA) ? bufB : bufA;
    linear_forward<TB, kNT, WM>(A.actor.net_a.weight[0], A.actor.net_a.bias[0], A.actor.net_a.dims[0], 2 * Ad, h, out, nullptr, 0, false);
    __syncthreads();
    if (threadIdx.x < TB) {
        const int b = threadIdx.x;
        s_logp[b] = actor_head(out, Ad, b, A.eps_next, A.seed, A.draw, (uint32_t)(slot0 + b), kStreamSacNext, sa, S);
        s_minq[b] = INFINITY;
    }
    __syncthreads();
    linear_forward<TB, kNT, WM>(A.target.encoder.weight[0], A.target.encoder.bias[0], S + Ad, d0, sa, enc, nullptr, 0, false);
    __syncthreads();
    for (int e = 0; e < A.target.num_ensembles; ++e) {
        const float* q = mlp_forward_nocache(A.target.decoder[e], false, enc, bufA, bufB);
        if (threadIdx.x < TB) s_minq[threadIdx.x] = fminf(s_minq[threadIdx.x], q[T::elem(0, threadIdx.x)]);
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < TB * Ad; idx += kNT)
        if (slot0 + idx / Ad < A.batch) A.d_tanh[(size_t)slot0 * Ad + idx] = 0.0f;
    if (threadIdx.x < TB && s_row[threadIdx.x] >= 0) {
        const int b = threadIdx.x;
        const int64_t r = s_row[b];
        const float alpha = expf(__ldcg(A.alpha_log));
        A.q_label[slot0 + b] = A.rb.rewards[r] + A.rb.undones[r] * A.hp.gamma * (s_minq[b] - s_logp[b] * alpha);
    }
}

// smem map: sa [S + A] | enc [d0] | d_enc [d0] | cache [2 * sum hidden dims of one decoder] | out [1] | dzA [maxw] | dzB [maxw]
template <bool TRAIN>
DEV void ensemble_pass(const SacArgs& A, const b200rl_sac_critic& cri, float* smem, const int64_t* s_row, int slot0, float* red,
                       float& q_mean_out, float& td_out) {
    const int S = A.rb.state_dim, Ad = A.rb.action_dim, d0 = cri.encoder.dims[1], E = cri.num_ensembles;
    float* sa = smem;
    float* enc = sa + (S + Ad) * TB;
    float* d_enc = enc + d0 * TB;
    float* cache = d_enc + d0 * TB;
    int hidden = 0;
    for (int l = 1; l < cri.decoder[0].num_linear; ++l) hidden += cri.decoder[0].dims[l];
    float* out = cache + 2 * hidden * TB;
    float* dzA = out + TB;
    float* dzB = dzA + A.maxw * TB;
    linear_forward<TB, kNT, WM>(cri.encoder.weight[0], cri.encoder.bias[0], S + Ad, d0, sa, enc, nullptr, 0, false);
    for (int i = threadIdx.x; i < d0 * TB; i += kNT) d_enc[i] = 0.0f;
    __syncthreads();
    const float inv_be = 1.0f / ((float)A.batch * (float)E);
    float q_sum = 0.0f, td = 0.0f;   // valid in threads < TB
    const int enc_numel = mlp_numel(cri.encoder), dec_numel = mlp_numel(cri.decoder[0]);
    {
        const int e = blockIdx.y;   // one decoder per CTA
        const float* xs[B200RL_MAX_LINEAR + 1];
        const float* gs[B200RL_MAX_LINEAR + 1];
        mlp_forward_cached(cri.decoder[e], false, enc, cache, out, xs, gs);
        if (threadIdx.x < TB) {
            const int b = threadIdx.x;
            const bool valid = s_row[b] >= 0;
            const float q = out[T::elem(0, b)];
            float dq;
            if (TRAIN) {   // td_error = mean_e (q - label)^2 * unmask; obj = mean_b                      (:57-63)
                const float um = valid ? A.rb.unmasks[s_row[b]] : 0.0f;
                const float err = valid ? q - A.q_label[slot0 + b] : 0.0f;
                td += err * err * um / (float)E;
                dq = 2.0f * err * um * inv_be;
            } else {       // d(-obj_actor) / d q of  mean_b mean_e Q_target(s, a_pg)                      (:81-83)
                q_sum += valid ? q / (float)E : 0.0f;
                dq = valid ? -inv_be : 0.0f;
            }
            dzA[T::elem(0, b)] = dq;
        }
        __syncthreads();
        mlp_backward(cri.decoder[e], xs, gs, dzA, dzB, TRAIN ? A.grads_c + enc_numel + e * dec_numel : nullptr, d_enc, true);
    }
    if (TRAIN) {
        weight_grad<TB, kNT>(d_enc, sa, d0, S + Ad, A.grads_c, A.grads_c + d0 * (S + Ad), true);
        __syncthreads();
    } else {
        // d / d (state, action) through the raw encoder; only the action rows are needed
        data_grad<TB, kNT, WM>(cri.encoder.weight[0], d_enc, nullptr, dzA, d0, S + Ad);
        __syncthreads();
        for (int idx = threadIdx.x; idx < TB * Ad; idx += kNT) {   // the E decoders' CTAs add up (zeroed by the label kernel)
            const int b = idx / Ad, a = idx - b * Ad;
            if (s_row[b] >= 0) atomicAdd(&A.d_tanh[(size_t)(slot0 + b) * Ad + a], dzA[T::elem(S + a, b)]);
        }
    }
    q_mean_out = q_sum;
    td_out = td;
    (void)red;
}

__global__ void __launch_bounds__(kNT) sac_critic_kernel(const __grid_constant__ SacArgs A) {
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    __shared__ int64_t s_row[TB];
    __shared__ float red[32];
    __shared__ int s_flag;
    const int S = A.rb.state_dim, Ad = A.rb.action_dim;
    const int slot0 = blockIdx.x * TB;
    if (threadIdx.x < TB) s_row[threadIdx.x] = slot0 + threadIdx.x < A.batch ? A.rows[slot0 + threadIdx.x] : -1;
    __syncthreads();
    gather_rows(A.rb.states, S, s_row, 0, 0, smem);
    gather_rows(A.rb.actions, Ad, s_row, 0, S, smem);
    __syncthreads();
    float q_unused, td;
    ensemble_pass<true>(A, A.critic, smem, s_row, slot0, red, q_unused, td);
    if (threadIdx.x < 32) {
        const float s = warp_sum(threadIdx.x < TB ? td : 0.0f);
        if (threadIdx.x == 0) atomicAdd(&A.hdr->td_sum, (double)s);
    }
    if (!last_block(&A.hdr->ticket[0], &s_flag)) return;
    // ---- clip_grad_norm_ + Adam over the critic's parameters, soft update of the target, bookkeeping
    apply_group(A.g_critic, A.adam_critic, A.grads_c, A.numel_c, A.hp.clip_grad_norm, red);
    __syncthreads();
    const float tau = A.hp.soft_update_tau;
    for (int t = 0; t < A.g_critic.num_tensors; ++t)
        for (int i = threadIdx.x; i < A.g_critic.numel[t]; i += kNT) {
            float* tar = A.target_param[t] + i;
            *tar = __fadd_rn(__fmul_rn(A.g_critic.param[t][i], tau), __fmul_rn(__ldcg(tar), 1.0f - tau));   // cur * tau + tar * (1 - tau)
        }
    for (int i = threadIdx.x; i < A.numel_c; i += kNT) A.grads_c[i] = 0.0f;
    if (threadIdx.x == 0) {
        A.hdr->obj_critic_sum += __ldcg(&A.hdr->td_sum) / (double)A.batch;
        A.hdr->td_sum = 0.0;
        A.hdr->ticket[0] = 0u;
    }
}

__global__ void __launch_bounds__(kNT) sac_pgrad_kernel(const __grid_constant__ SacArgs A) {
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    __shared__ int64_t s_row[TB];
    __shared__ float red[32];
    __shared__ int s_flag;
    const int S = A.rb.state_dim, Ad = A.rb.action_dim, d0 = A.target.encoder.dims[1];
    const int slot0 = blockIdx.x * TB;
    if (threadIdx.x < TB) s_row[threadIdx.x] = slot0 + threadIdx.x < A.batch ? A.rows[slot0 + threadIdx.x] : -1;
    __syncthreads();
    float* sa = smem;
    gather_rows(A.rb.states, S, s_row, 0, 0, sa);
    __syncthreads();
    // the actor's forward pass borrows the (still unused) dz buffers at the end of the ensemble's shared-memory map
    int hidden = 0;
    for (int l = 1; l < A.target.decoder[0].num_linear; ++l) hidden += A.target.decoder[0].dims[l];
    float* bufA = sa + ((S + Ad) + 2 * d0 + 2 * hidden + 1) * TB;
    float* bufB = bufA + A.maxw * TB;
    float* h = mlp_forward_nocache(A.actor.net_s, true, sa, bufA, bufB);
    float* out = (h == bufA) ? bufB : bufA;
    linear_forward<TB, kNT, WM>(A.actor.net_a.weight[0], A.actor.net_a.bias[0], A.actor.net_a.dims[0], 2 * Ad, h, out, nullptr, 0, false);
    __syncthreads();
    float logp = 0.0f;
    if (threadIdx.x < TB) {
        const int b = threadIdx.x;
        logp = actor_head(out, Ad, b, A.eps_pg, A.seed, A.draw, (uint32_t)(slot0 + b), kStreamSacPg, sa, S);
        if (s_row[b] >= 0 && blockIdx.y == 0) A.logp[slot0 + b] = logp; else logp = 0.0f;   // counted once over the decoders' CTAs
    }
    __syncthreads();
    float q_mean, td_unused;
    ensemble_pass<false>(A, A.target, smem, s_row, slot0, red, q_mean, td_unused);
    if (threadIdx.x < 32) {
        const float sl = warp_sum(threadIdx.x < TB ? logp : 0.0f), sq = warp_sum(threadIdx.x < TB ? q_mean : 0.0f);
        if (threadIdx.x == 0) { atomicAdd(&A.hdr->logp_sum, (double)sl); atomicAdd(&A.hdr->q_sum, (double)sq); }
    }
    if (!last_block(&A.hdr->ticket[1], &s_flag)) return;
    if (threadIdx.x == 0) {
        // ---- temperature: obj_alpha = mean(alpha_log * (target_entropy - logprob))                           (:71-74)
        const float mean_logp = (float)(__ldcg(&A.hdr->logp_sum) / (double)A.batch);
        float g = A.hp.target_entropy - mean_logp;
        if (A.hp.clip_grad_norm > 0.0f) g *= fminf(A.hp.clip_grad_norm / (fabsf(g) + 1e-6f), 1.0f);
        float p = *A.alpha_log, m = *A.alpha_m, v = *A.alpha_v;
        adam_one(p, m, v, g, A.alpha_b1, A.alpha_b2, A.alpha_eps, A.adam_alpha);
        *A.alpha_m = m; *A.alpha_v = v;
        const float alpha = expf(p);                               // after its step, before the clamp           (:77-79)
        A.hdr->alpha_used = alpha;
        *A.alpha_log = fminf(fmaxf(p, -16.0f), 2.0f);
        const float mean_q = (float)(__ldcg(&A.hdr->q_sum) / (double)A.batch);
        A.hdr->obj_actor_sum += (double)(mean_q - mean_logp * alpha);   // obj_actor = (q_value_pg - logprob * alpha).mean()
        A.hdr->logp_sum = 0.0; A.hdr->q_sum = 0.0;
        A.hdr->ticket[1] = 0u;
    }
}

// smem map: X0 [S] | cache [2 * sum net_s dims[1..]] | out [2A] | dzA [maxw] | dzB [maxw]
__global__ void __launch_bounds__(kNT) sac_actor_kernel(const __grid_constant__ SacArgs A) {
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    __shared__ int64_t s_row[TB];
    __shared__ float red[32];
    __shared__ int s_flag;
    const int S = A.rb.state_dim, Ad = A.rb.action_dim;
    const b200rl_net& ns = A.actor.net_s;
    const b200rl_net& na = A.actor.net_a;
    const int slot0 = blockIdx.x * TB;
    if (threadIdx.x < TB) s_row[threadIdx.x] = slot0 + threadIdx.x < A.batch ? A.rows[slot0 + threadIdx.x] : -1;
    __syncthreads();
    float* x0 = smem;
    float* cache = x0 + S * TB;
    int hidden = 0;
    for (int l = 1; l <= ns.num_linear; ++l) hidden += ns.dims[l];
    float* out = cache + 2 * hidden * TB;
    float* dzA = out + 2 * Ad * TB;
    float* dzB = dzA + A.maxw * TB;
    gather_rows(A.rb.states, S, s_row, 0, 0, x0);
    __syncthreads();
    const float* xs[B200RL_MAX_LINEAR + 1];
    const float* gs[B200RL_MAX_LINEAR + 1];
    mlp_forward_cached(ns, true, x0, cache, nullptr, xs, gs);
    const float* h = xs[ns.num_linear];
    linear_forward<TB, kNT, WM>(na.weight[0], na.bias[0], na.dims[0], 2 * Ad, h, out, nullptr, 0, false);
    __syncthreads();
    if (threadIdx.x < TB) {
        // oracle/sac_oracle.py::actor_backward: d_tanh from the target ensemble, d_logprob = alpha / B                 (:81-83)
        const int b = threadIdx.x;
        const bool valid = s_row[b] >= 0;
        const float d_lp = valid ? __ldcg(&A.hdr->alpha_used) / (float)A.batch : 0.0f;
        for (int a = 0; a < Ad; ++a) {
            const float avg = out[T::elem(a, b)], raw = out[T::elem(Ad + a, b)];
            const float lsd = fminf(fmaxf(raw, -16.0f), 2.0f);
            const float sd = expf(lsd);
            const float e = A.eps_pg ? A.eps_pg[(size_t)(slot0 + b) * Ad + a] : sac_normal(A.seed, A.draw, (uint32_t)(slot0 + b), a, kStreamSacPg);
            const float t = tanhf(__fadd_rn(avg, __fmul_rn(sd, e)));
            const float one_m_t2 = 1.0f - t * t;
            const float dt_ = valid ? A.d_tanh[(size_t)(slot0 + b) * Ad + a] : 0.0f;
            const float d_action = dt_ * one_m_t2 + d_lp * (2.0f * t * one_m_t2 / (1.000001f - t * t));
            const float inside = (raw >= -16.0f && raw <= 2.0f) ? 1.0f : 0.0f;
            dzA[T::elem(a, b)] = d_action;
            dzA[T::elem(Ad + a, b)] = (d_action * sd * e - d_lp) * inside;
        }
    }
    __syncthreads();
    // net_a (one raw Linear), then net_s (every layer activated: its last activation derivative is applied on the way down)
    const int ns_numel = mlp_numel(ns);
    weight_grad<TB, kNT>(dzA, h, 2 * Ad, na.dims[0], A.grads_a + ns_numel, A.grads_a + ns_numel + 2 * Ad * na.dims[0], true);
    data_grad<TB, kNT, WM>(na.weight[0], dzA, gs[ns.num_linear], dzB, 2 * Ad, na.dims[0]);
    __syncthreads();
    mlp_backward(ns, xs, gs, dzB, dzA, A.grads_a, nullptr, false);
    if (!last_block(&A.hdr->ticket[2], &s_flag)) return;
    apply_group(A.g_actor, A.adam_actor, A.grads_a, A.numel_a, A.hp.clip_grad_norm, red);
    __syncthreads();
    for (int i = threadIdx.x; i < A.numel_a; i += kNT) A.grads_a[i] = 0.0f;
    if (threadIdx.x == 0) {
        A.hdr->ticket[2] = 0u;
        if (A.last_update) {
            A.out_scalars[0] = (float)(A.hdr->obj_critic_sum * (double)A.inv_updates);
            A.out_scalars[1] = (float)(A.hdr->obj_actor_sum * (double)A.inv_updates);
        }
    }
}

// ---- exploration: ActorSAC.get_action
struct StepArgs {
    b200rl_sac_actor actor;
    const float* state;
    int64_t rows;
    const float* eps;
    uint64_t seed, step;
    int64_t env_offset;
    float* action;
    int maxw;
};
__global__ void __launch_bounds__(kNT) sac_policy_step_kernel(const __grid_constant__ StepArgs P) {
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    const int S = P.actor.net_s.dims[0], Ad = P.actor.net_a.dims[1] / 2;
    float* x0 = smem;
    float* bufA = x0 + S * TB;
    float* bufB = bufA + P.maxw * TB;
    float* act = bufB + P.maxw * TB;
    const int64_t row0 = (int64_t)blockIdx.x * TB;
    b200rl_net plain = P.actor.net_s;
    load_state_tile<TB, kNT>(plain, P.state, P.rows, row0, x0);
    __syncthreads();
    float* h = mlp_forward_nocache(P.actor.net_s, true, x0, bufA, bufB);
    float* out = (h == bufA) ? bufB : bufA;
    linear_forward<TB, kNT, WM>(P.actor.net_a.weight[0], P.actor.net_a.bias[0], P.actor.net_a.dims[0], 2 * Ad, h, out, nullptr, 0, false);
    __syncthreads();
    if (threadIdx.x < TB && row0 + threadIdx.x < P.rows) {
        const int b = threadIdx.x;
        const int64_t row = row0 + b;
        for (int a = 0; a < Ad; ++a) {
            const float avg = out[T::elem(a, b)];
            const float sd = expf(fminf(fmaxf(out[T::elem(Ad + a, b)], -16.0f), 2.0f));
            float e;
            if (P.eps) e = P.eps[row * Ad + a];
            else {
                const uint64_t env = (uint64_t)(P.env_offset + row);
                const uint4 r = philox4x32_10(make_uint4((uint32_t)env, (uint32_t)(env >> 32) ^ ((uint32_t)(a >> 1) << 8), (uint32_t)P.step,
                                                         (uint32_t)(P.step >> 32) ^ (kStreamSacExplore << 16)),
                                              make_uint2((uint32_t)P.seed, (uint32_t)(P.seed >> 32)));
                const float2 n = box_muller(r.x, r.y);
                e = (a & 1) ? n.y : n.x;
            }
            P.action[row * Ad + a] = tanhf(__fadd_rn(avg, __fmul_rn(sd, e)));
        }
    }
    (void)act;
}

// ---- ReplayBuffer.update: ring write of `rows` time rows
__global__ void replay_append_kernel(const b200rl_replay_buffer rb, int p, int rows, const float* states, const float* actions,
                                     const float* rewards, const uint8_t* undones, const uint8_t* unmasks) {
    const int64_t n = rb.num_seqs, per_row = n * (rb.state_dim + rb.action_dim + 3);
    const int64_t total = (int64_t)rows * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i / per_row;
        int64_t j = i - t * per_row;
        const int64_t dst_t = (p + t) % rb.max_size;
        const int64_t ns = n * rb.state_dim, na = n * rb.action_dim;
        if (j < ns) { rb.states[dst_t * ns + j] = states[t * ns + j]; continue; }
        j -= ns;
        if (j < na) { rb.actions[dst_t * na + j] = actions[t * na + j]; continue; }
        j -= na;
        if (j < n) { rb.rewards[dst_t * n + j] = rewards[t * n + j]; continue; }
        j -= n;
        if (j < n) { rb.undones[dst_t * n + j] = undones[t * n + j] ? 1.0f : 0.0f; continue; }
        j -= n;
        rb.unmasks[dst_t * n + j] = unmasks[t * n + j] ? 1.0f : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------------------- host helpers
int check_mlp(const b200rl_net* n, const char* name) {
    B200RL_REQUIRE(n->num_linear >= 1 && n->num_linear <= B200RL_MAX_LINEAR, "%s: num_linear=%d", name, n->num_linear);
    for (int l = 0; l < n->num_linear; ++l) {
        B200RL_REQUIRE(n->weight[l] && n->bias[l], "%s: NULL parameter of layer %d", name, l);
        B200RL_REQUIRE(n->dims[l] >= 1 && n->dims[l + 1] >= 1 && n->dims[l] <= 1024 && n->dims[l + 1] <= 1024, "%s: dims", name);
    }
    B200RL_REQUIRE(n->state_avg == nullptr, "%s: SAC nets have no state_norm", name);
    return 0;
}
int host_numel(const b200rl_net& n) {
    int s = 0;
    for (int l = 0; l < n.num_linear; ++l) s += n.dims[l + 1] * n.dims[l] + n.dims[l + 1];
    return s;
}
int max_width(const b200rl_net& n, int m) {
    for (int l = 0; l <= n.num_linear; ++l) m = n.dims[l] > m ? n.dims[l] : m;
    return m;
}
int check_actor(const b200rl_sac_actor* a) {
    B200RL_REQUIRE(a, "sac: NULL actor");
    if (int rc = check_mlp(&a->net_s, "sac.actor.net_s")) return rc;
    if (int rc = check_mlp(&a->net_a, "sac.actor.net_a")) return rc;
    B200RL_REQUIRE(a->net_a.num_linear == 1 && a->net_a.dims[0] == a->net_s.dims[a->net_s.num_linear] && a->net_a.dims[1] % 2 == 0,
                   "sac.actor: net_a must be one Linear net_dims[-1] -> 2 * action_dim");
    return 0;
}
int check_critic(const b200rl_sac_critic* c, const char* name) {
    B200RL_REQUIRE(c, "sac: NULL critic");
    B200RL_REQUIRE(c->num_ensembles >= 1 && c->num_ensembles <= B200RL_SAC_MAX_ENSEMBLES, "%s: num_ensembles=%d", name, c->num_ensembles);
    if (int rc = check_mlp(&c->encoder, name)) return rc;
    B200RL_REQUIRE(c->encoder.num_linear == 1, "%s: the encoder is one Linear", name);
    for (int e = 0; e < c->num_ensembles; ++e) {
        if (int rc = check_mlp(&c->decoder[e], name)) return rc;
        B200RL_REQUIRE(c->decoder[e].dims[0] == c->encoder.dims[1] && c->decoder[e].dims[c->decoder[e].num_linear] == 1 &&
                           c->decoder[e].num_linear == c->decoder[0].num_linear,
                       "%s: decoder %d shape", name, e);
        for (int l = 0; l <= c->decoder[0].num_linear; ++l)
            B200RL_REQUIRE(c->decoder[e].dims[l] == c->decoder[0].dims[l], "%s: decoders must share their dims", name);
    }
    return 0;
}
struct Layout {
    int numel_c, numel_a, maxw;
    size_t off_rows, off_label, off_logp, off_dtanh, off_gc, off_ga, bytes;
    size_t smem_ensemble, smem_label, smem_actor;
};
Layout layout_of(const b200rl_sac_actor* a, const b200rl_sac_critic* c, int batch) {
    Layout L{};
    const int S = a->net_s.dims[0], Ad = a->net_a.dims[1] / 2, d0 = c->encoder.dims[1];
    L.numel_a = host_numel(a->net_s) + host_numel(a->net_a);
    L.numel_c = host_numel(c->encoder) + c->num_ensembles * host_numel(c->decoder[0]);
    int m = max_width(a->net_s, 2 * Ad);
    m = max_width(c->encoder, m);
    m = max_width(c->decoder[0], m);
    L.maxw = m;
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = align(sizeof(SacHeader));
    L.off_rows = off; off = align(off + (size_t)batch * 8);
    L.off_label = off; off = align(off + (size_t)batch * 4);
    L.off_logp = off; off = align(off + (size_t)batch * 4);
    L.off_dtanh = off; off = align(off + (size_t)batch * Ad * 4);
    L.off_gc = off; off = align(off + (size_t)L.numel_c * 4);
    L.off_ga = off; off = align(off + (size_t)L.numel_a * 4);
    L.bytes = off;
    int hidden_d = 0, hidden_s = 0;
    for (int l = 1; l < c->decoder[0].num_linear; ++l) hidden_d += c->decoder[0].dims[l];
    for (int l = 1; l <= a->net_s.num_linear; ++l) hidden_s += a->net_s.dims[l];
    L.smem_ensemble = (size_t)((S + Ad) + 2 * d0 + 2 * hidden_d + 1 + 2 * m) * TB * 4;
    L.smem_label = (size_t)((S + Ad) + d0 + 2 * m + 2) * TB * 4;
    L.smem_actor = (size_t)(S + 2 * hidden_s + 2 * Ad + 2 * m) * TB * 4;
    return L;
}
int check_group(const b200rl_param_group* g, int numel, const char* name) {
    B200RL_REQUIRE(g && g->num_tensors >= 1 && g->num_tensors <= B200RL_MAX_GROUP_TENSORS, "%s: bad parameter group", name);
    int s = 0;
    for (int t = 0; t < g->num_tensors; ++t) {
        B200RL_REQUIRE(g->param[t] && g->exp_avg[t] && g->exp_avg_sq[t] && g->numel[t] >= 1, "%s: tensor %d", name, t);
        s += g->numel[t];
    }
    B200RL_REQUIRE(s == numel, "%s: the group holds %d elements, the nets %d", name, s, numel);
    return 0;
}
AdamScalars group_scalars(const b200rl_param_group* g, int64_t step) {
    AdamScalars s;
    s.step_size = (float)((double)g->lr / (1.0 - pow((double)g->beta1, (double)step)));
    s.bc2_sqrt = (float)sqrt(1.0 - pow((double)g->beta2, (double)step));
    return s;
}

}  // namespace

extern "C" {

int b200rl_replay_append(const b200rl_replay_buffer* buffer, int32_t p, int32_t rows, const float* states, const float* actions,
                         const float* rewards, const uint8_t* undones, const uint8_t* unmasks, void* stream) {
    B200RL_REQUIRE(buffer && buffer->states && buffer->actions && buffer->rewards && buffer->undones && buffer->unmasks,
                   "replay_append: NULL buffer");
    B200RL_REQUIRE(states && actions && rewards && undones && unmasks, "replay_append: NULL rollout tensor");
    B200RL_REQUIRE(rows >= 1 && rows <= buffer->max_size && p >= 0 && p <= buffer->max_size, "replay_append: p=%d rows=%d max_size=%d", p,
                   rows, buffer->max_size);
    const int64_t total = (int64_t)rows * buffer->num_seqs * (buffer->state_dim + buffer->action_dim + 3);
    const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    replay_append_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*buffer, p, rows, states, actions, rewards, undones, unmasks);
    B200RL_COUNT_LAUNCH(1);
    B200RL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int b200rl_sac_policy_step(const b200rl_sac_actor* actor, const float* state, int64_t rows, const float* eps, uint64_t seed,
                           uint64_t step, int64_t env_offset, float* action, void* stream) {
    if (int rc = check_actor(actor)) return rc;
    B200RL_REQUIRE(state && action && rows >= 1, "sac_policy_step: bad arguments");
    StepArgs P{};
    P.actor = *actor; P.state = state; P.rows = rows; P.eps = eps; P.seed = seed; P.step = step; P.env_offset = env_offset;
    P.action = action;
    P.maxw = max_width(actor->net_s, actor->net_a.dims[1]);
    const size_t smem = (size_t)(actor->net_s.dims[0] + 2 * P.maxw) * TB * 4;
    B200RL_REQUIRE(smem <= 227 * 1024, "sac_policy_step: nets too wide");
    B200RL_CHECK_CUDA(cudaFuncSetAttribute(sac_policy_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sac_policy_step_kernel<<<(unsigned)((rows + TB - 1) / TB), kNT, smem, (cudaStream_t)stream>>>(P);
    B200RL_COUNT_LAUNCH(1);
    B200RL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int64_t b200rl_sac_workspace_bytes(const b200rl_sac_actor* actor, const b200rl_sac_critic* critic, int32_t batch_size) {
    if (!actor || !critic || batch_size < 1) return 0;
    return (int64_t)layout_of(actor, critic, batch_size).bytes;
}

int b200rl_sac_update(const b200rl_sac_actor* actor, const b200rl_sac_critic* critic, const b200rl_sac_critic* critic_target,
                      b200rl_param_group* actor_group, b200rl_param_group* critic_group, b200rl_param_group* alpha_group,
                      const b200rl_replay_buffer* buffer, int32_t cur_size, const b200rl_sac_hyper* hyper, int32_t batch_size,
                      int32_t update_times, const int64_t* ids, const float* eps_next, const float* eps_pg, uint64_t seed,
                      uint64_t draw_offset, float* out_scalars, void* workspace, int64_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int rc = check_actor(actor)) return rc;
    if (int rc = check_critic(critic, "sac.critic")) return rc;
    if (int rc = check_critic(critic_target, "sac.critic_target")) return rc;
    B200RL_REQUIRE(buffer && hyper && out_scalars && workspace && alpha_group, "sac_update: NULL argument");
    B200RL_REQUIRE(batch_size >= 1 && update_times >= 1, "sac_update: batch_size=%d update_times=%d", batch_size, update_times);
    B200RL_REQUIRE(cur_size >= 2 && cur_size <= buffer->max_size, "sac_update: cur_size=%d (needs >= 2 time rows)", cur_size);
    const int S = actor->net_s.dims[0], Ad = actor->net_a.dims[1] / 2;
    B200RL_REQUIRE(buffer->state_dim == S && buffer->action_dim == Ad && critic->encoder.dims[0] == S + Ad &&
                       critic_target->encoder.dims[0] == S + Ad,
                   "sac_update: state / action dims of the nets and the buffer disagree");
    B200RL_REQUIRE(critic_target->num_ensembles == critic->num_ensembles && critic_target->encoder.dims[1] == critic->encoder.dims[1],
                   "sac_update: critic and target shapes differ");
    const Layout L = layout_of(actor, critic, batch_size);
    B200RL_REQUIRE(workspace_bytes >= (int64_t)L.bytes, "sac_update: workspace too small (%lld < %zu)", (long long)workspace_bytes, L.bytes);
    B200RL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "sac_update: workspace must be 256-byte aligned");
    B200RL_REQUIRE(L.smem_ensemble <= 227 * 1024 && L.smem_label <= 227 * 1024 && L.smem_actor <= 227 * 1024,
                   "sac_update: nets too wide for a 32-sample tile in shared memory");
    if (int rc = check_group(actor_group, L.numel_a, "sac.actor_group")) return rc;
    if (int rc = check_group(critic_group, L.numel_c, "sac.critic_group")) return rc;
    B200RL_REQUIRE(alpha_group->num_tensors == 1 && alpha_group->numel[0] == 1 && alpha_group->param[0], "sac.alpha_group: one scalar");

    SacArgs A{};
    A.actor = *actor; A.critic = *critic; A.target = *critic_target;
    A.g_actor = *actor_group; A.g_critic = *critic_group;
    {   // the target's tensors in the critic group's order: encoder W, b, then every decoder's layers
        int t = 0;
        A.target_param[t++] = critic_target->encoder.weight[0];
        A.target_param[t++] = critic_target->encoder.bias[0];
        for (int e = 0; e < critic_target->num_ensembles; ++e)
            for (int l = 0; l < critic_target->decoder[e].num_linear; ++l) {
                A.target_param[t++] = critic_target->decoder[e].weight[l];
                A.target_param[t++] = critic_target->decoder[e].bias[l];
            }
        B200RL_REQUIRE(t == critic_group->num_tensors, "sac.critic_group: %d tensors, the ensemble has %d", critic_group->num_tensors, t);
    }
    A.alpha_log = alpha_group->param[0]; A.alpha_m = alpha_group->exp_avg[0]; A.alpha_v = alpha_group->exp_avg_sq[0];
    A.alpha_b1 = alpha_group->beta1; A.alpha_b2 = alpha_group->beta2; A.alpha_eps = alpha_group->eps;
    A.rb = *buffer; A.cur_size = cur_size; A.hp = *hyper; A.batch = batch_size; A.seed = seed;
    char* ws = reinterpret_cast<char*>(workspace);
    A.hdr = reinterpret_cast<SacHeader*>(ws);
    A.rows = reinterpret_cast<int64_t*>(ws + L.off_rows);
    A.q_label = reinterpret_cast<float*>(ws + L.off_label);
    A.logp = reinterpret_cast<float*>(ws + L.off_logp);
    A.d_tanh = reinterpret_cast<float*>(ws + L.off_dtanh);
    A.grads_c = reinterpret_cast<float*>(ws + L.off_gc);
    A.grads_a = reinterpret_cast<float*>(ws + L.off_ga);
    A.numel_c = L.numel_c; A.numel_a = L.numel_a; A.maxw = L.maxw;
    A.inv_updates = 1.0f / (float)update_times;
    A.out_scalars = out_scalars;
    B200RL_CHECK_CUDA(cudaMemsetAsync(workspace, 0, L.bytes, stream));
    B200RL_CHECK_CUDA(cudaFuncSetAttribute(sac_label_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_label));
    B200RL_CHECK_CUDA(cudaFuncSetAttribute(sac_critic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_ensemble));
    B200RL_CHECK_CUDA(cudaFuncSetAttribute(sac_pgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_ensemble));
    B200RL_CHECK_CUDA(cudaFuncSetAttribute(sac_actor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem_actor));
    const unsigned tiles = (unsigned)((batch_size + TB - 1) / TB);
    for (int u = 0; u < update_times; ++u) {
        A.ids = ids ? ids + (size_t)u * batch_size : nullptr;
        A.eps_next = eps_next ? eps_next + (size_t)u * batch_size * Ad : nullptr;
        A.eps_pg = eps_pg ? eps_pg + (size_t)u * batch_size * Ad : nullptr;
        A.draw = draw_offset + (uint64_t)u;
        A.adam_actor = group_scalars(actor_group, actor_group->step + u + 1);
        A.adam_critic = group_scalars(critic_group, critic_group->step + u + 1);
        A.adam_alpha = group_scalars(alpha_group, alpha_group->step + u + 1);
        A.last_update = u == update_times - 1;
        sac_label_kernel<<<tiles, kNT, L.smem_label, stream>>>(A);
        sac_critic_kernel<<<dim3(tiles, (unsigned)critic->num_ensembles), kNT, L.smem_ensemble, stream>>>(A);
        sac_pgrad_kernel<<<dim3(tiles, (unsigned)critic->num_ensembles), kNT, L.smem_ensemble, stream>>>(A);
        sac_actor_kernel<<<tiles, kNT, L.smem_actor, stream>>>(A);
    }
    B200RL_COUNT_LAUNCH(4 * update_times);
    B200RL_CHECK_CUDA(cudaGetLastError());
    actor_group->step += update_times;
    critic_group->step += update_times;
    alpha_group->step += update_times;
    return 0;
}

}  // extern "C"
