// Pieces shared by the two PPO-update implementations (update.cu: generic depth / width on the FP32 pipe; update_tc.cu:
// S -> 64 -> 64 -> OUT GELU nets on tcgen05): the kernel argument block, torch.optim.Adam's arithmetic and the per-net
// clip_grad_norm_ + Adam.step (reference elegantrl/agents/AgentBase.py:239-248).
#pragma once
#include "common.cuh"

struct AdamScalars {
    float step_size;  // lr / (1 - beta1^t)
    float bc2_sqrt;   // sqrt(1 - beta2^t)
};

struct UpdateArgs {
    b200rl_net net[2];  // 0 = actor, 1 = critic
    b200rl_adam opt[2];
    AdamScalars adam[2];
    b200rl_train_buffer buf;
    b200rl_ppo_hyper hp;
    const int64_t* ids;  // [local_batch] or nullptr
    uint64_t seed, draw;
    int local_batch, global_batch;
    float* grads;              // flat: actor tensors then critic tensors
    int grad_off[2];           // float offset of each net's first tensor
    int grad_numel[2];
    WorkspaceHeader* hdr;
    double* loss_sums;         // [3] obj_critic, obj_surrogate, obj_entropy (sums over updates)
    int fused_apply;
    int smem_scalar_off;       // float offset of the per-sample scalar block in dynamic smem
    int maxdim;
    int stage_weights;         // parameters of the net fit in shared memory: stage them per minibatch
    int smem_weight_off;       // float offset of the staged parameters in dynamic smem
    int smem_gacc_off;         // float offset of the per-CTA gradient accumulator (large minibatches), or -1
    int update_times;          // persistent (cluster) kernel: minibatches per launch
    int grad_stride;           // floats between the two gradient buffers of the persistent kernel
    float* out_scalars;        // persistent kernel: means of the three logged scalars
    // env-sharded update with the in-kernel gradient exchange over peer memory (update_tc.cu, b200rl_ppo_update_sharded)
    int px_on;                 // 0: single GPU; 1: gradient all-reduce per minibatch; 2: record gather once per cycle
    int px_local_batch;        // record gather: samples per rank and minibatch
    b200rl_peer_exchange px;
    const double* stat_sums;   // this shard's advantage sums (b200rl_gae)
    double count_all, count_lat;
    float* stats_out;          // [4] reduced statistics
    int profile;               // B200RL_PROFILE=1: clock64 phase marks of the actor CTA into the workspace header (+64 B)
};

// update_tc.cu: the tcgen05 implementation for S -> 64 -> 64 -> OUT GELU nets
bool b200rl_update_tc_eligible(const b200rl_net* actor, const b200rl_net* critic, const b200rl_ppo_hyper* hp);
int b200rl_launch_update_tc(const UpdateArgs& A, int tiles, cudaStream_t stream);

template <int NT>
DEV float block_sum(float v, float* red /*[32]*/) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += red[w];
    return t;
}

// torch.optim.Adam single-tensor update, op order of torch (_single_tensor_adam), no FMA contraction
DEV void adam_one(float& p, float& m, float& v, float g, float b1, float b2, float eps, const AdamScalars& as) {
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), 1.0f - b1));                 // exp_avg.lerp_(grad, 1 - beta1)
    v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(1.0f - b2, g), g));  // mul_(beta2).addcmul_(g, g, 1 - beta2)
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), as.bc2_sqrt), eps);
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(-as.step_size, m), denom));         // addcdiv_(exp_avg, denom, -step_size)
}

// The same update with MUFU reciprocal / reciprocal-square-root instead of the IEEE division and square root (which expand to
// ~190 instructions per element and were half of the tcgen05 update kernel's instruction stream, executed by only four warps):
// sqrt(v) = v * rsqrt(v), / bc2_sqrt -> * (1 / bc2_sqrt), the final quotient with one Newton step on the approximate
// reciprocal.  Deviates from torch by <= 2 ulp per step (rtol 1e-4 parity is unaffected; replicas on several GPUs still agree
// bit for bit because they all run this arithmetic).
DEV void adam_one_fast(float& p, float& m, float& v, float g, float b1, float b2, float eps, const AdamScalars& as, float inv_bc2_sqrt) {
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), 1.0f - b1));
    v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(1.0f - b2, g), g));
    float rs, r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(fmaxf(v, 1e-37f)));   // one MUFU each (the _rn intrinsics expand to ~50 instructions)
    const float sq = v * rs;
    const float denom = fmaf(sq, inv_bc2_sqrt, eps);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(denom));
    r = fmaf(fmaf(-denom, r, 1.0f), r, r);                                   // one Newton step: relative error ~1e-7
    p = fmaf(-as.step_size * m, r, p);
}
template <bool FAST>
DEV void adam_step(float& p, float& m, float& v, float g, float b1, float b2, float eps, const AdamScalars& as, float inv_bc2_sqrt) {
    if (FAST) adam_one_fast(p, m, v, g, b1, b2, eps, as, inv_bc2_sqrt);
    else adam_one(p, m, v, g, b1, b2, eps, as);
}

// clip_grad_norm_ + Adam.step for one net.  The whole CTA computes the norm of the net's gradient; it then updates
// the part `part` of `nparts` of every tensor (nparts = 1: the whole net; > 1: the CTAs of a cluster share the net).
// GSMEM: the gradient lives in shared memory (plain loads) instead of the L2-resident workspace (ld.global.cg)
template <bool GSMEM>
DEV float ld_grad(const float* p) { return GSMEM ? *p : __ldcg(p); }
template <bool GSMEM>
DEV float4 ld_grad4(const float4* p) { return GSMEM ? *p : __ldcg(p); }

// three sums at the price of one (same per-value arithmetic as block_sum: warp shuffle tree, then the warps in order)
template <int NT>
DEV void block_sum3(float& a, float& b, float& c, float* red /*[32]*/) {
    static_assert(3 * (NT / 32) <= 32, "red[] holds three partial sums per warp");
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        red[w] = a; red[NT / 32 + w] = b; red[2 * (NT / 32) + w] = c;
    }
    __syncthreads();
    a = b = c = 0.0f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) { a += red[w]; b += red[NT / 32 + w]; c += red[2 * (NT / 32) + w]; }
}

// clip_grad_norm_'s scale factor for one net's gradient (the whole CTA takes part)
template <int NT, bool GSMEM>
DEV float clip_coef(const float* g, int numel, float clip_grad_norm, float* red) {
    // squared norm of the whole gradient: batches of 8 independent L2 loads per thread (a plain loop would wait for
    // one ~700-cycle load per iteration)
    float ss = 0.0f;
    for (int i0 = threadIdx.x; i0 < numel; i0 += 8 * NT) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (i0 + q * NT < numel) ? ld_grad<GSMEM>(g + i0 + q * NT) : 0.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q) ss = fmaf(v[q], v[q], ss);
    }
    const float total_norm = sqrtf(block_sum<NT>(ss, red));
    return clip_grad_norm > 0.0f ? fminf(clip_grad_norm / (total_norm + 1e-6f), 1.0f) : 1.0f;
}

template <int NT, bool GSMEM = false, bool FAST = false>
DEV void apply_net(const b200rl_net& net, const b200rl_adam& opt, const AdamScalars& as, const float* g, int numel,
                   float clip_grad_norm, float* red, int part = 0, int nparts = 1) {
    const float coef = clip_coef<NT, GSMEM>(g, numel, clip_grad_norm, red);
    const float b1 = opt.beta1, b2 = opt.beta2, eps = opt.eps, inv_bc2 = 1.0f / as.bc2_sqrt;
    const int n_tensors = 2 * net.num_linear + (net.action_std_log ? 1 : 0);
    const int first = part * NT + threadIdx.x, stride = nparts * NT;
    // Every tensor is visited in rounds of up to kBatch tensors: all loads of a round are issued before any
    // arithmetic / store, so their (L2) latencies overlap instead of adding up tensor after tensor.
    constexpr int kBatch = 4;
    int off = 0;
    for (int t0 = 0; t0 < n_tensors; t0 += kBatch) {
        float* P[kBatch]; float* M[kBatch]; float* V[kBatch];
        int cnt[kBatch], goff[kBatch];
        bool vecs[kBatch];
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
            const int ti = t0 + q;
            cnt[q] = 0; goff[q] = off; vecs[q] = false; P[q] = M[q] = V[q] = nullptr;
            if (ti < n_tensors) {
                const int l = ti >> 1;
                if (ti == 2 * net.num_linear) {
                    P[q] = net.action_std_log; M[q] = opt.exp_avg_std; V[q] = opt.exp_avg_sq_std; cnt[q] = net.dims[net.num_linear];
                } else if ((ti & 1) == 0) {
                    P[q] = net.weight[l]; M[q] = opt.exp_avg_w[l]; V[q] = opt.exp_avg_sq_w[l]; cnt[q] = net.dims[l + 1] * net.dims[l];
                } else {
                    P[q] = net.bias[l]; M[q] = opt.exp_avg_b[l]; V[q] = opt.exp_avg_sq_b[l]; cnt[q] = net.dims[l + 1];
                }
                vecs[q] = ((cnt[q] & 3) == 0) &&
                          (((reinterpret_cast<uintptr_t>(g + off) | reinterpret_cast<uintptr_t>(P[q]) | reinterpret_cast<uintptr_t>(M[q]) |
                             reinterpret_cast<uintptr_t>(V[q])) & 15) == 0);
                off += cnt[q];
            }
        }
        // first item of every tensor of the round (covers whole tensors up to 4 * stride floats): batched loads
        float4 gg[kBatch], pp[kBatch], mm[kBatch], vv[kBatch];
        bool have[kBatch];
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
            have[q] = vecs[q] && first < (cnt[q] >> 2);
            if (have[q]) {
                gg[q] = ld_grad4<GSMEM>(reinterpret_cast<const float4*>(g + goff[q]) + first);
                pp[q] = __ldcg(reinterpret_cast<const float4*>(P[q]) + first);
                mm[q] = __ldcg(reinterpret_cast<const float4*>(M[q]) + first);
                vv[q] = __ldcg(reinterpret_cast<const float4*>(V[q]) + first);
            }
        }
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
            if (have[q]) {
                adam_step<FAST>(pp[q].x, mm[q].x, vv[q].x, gg[q].x * coef, b1, b2, eps, as, inv_bc2);
                adam_step<FAST>(pp[q].y, mm[q].y, vv[q].y, gg[q].y * coef, b1, b2, eps, as, inv_bc2);
                adam_step<FAST>(pp[q].z, mm[q].z, vv[q].z, gg[q].z * coef, b1, b2, eps, as, inv_bc2);
                adam_step<FAST>(pp[q].w, mm[q].w, vv[q].w, gg[q].w * coef, b1, b2, eps, as, inv_bc2);
                reinterpret_cast<float4*>(P[q])[first] = pp[q];
                reinterpret_cast<float4*>(M[q])[first] = mm[q];
                reinterpret_cast<float4*>(V[q])[first] = vv[q];
            }
        }
        // remaining items (large tensors) and tensors that cannot be accessed as float4
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
            if (vecs[q]) {
                const float4* g4 = reinterpret_cast<const float4*>(g + goff[q]);
                float4 *p4 = reinterpret_cast<float4*>(P[q]), *m4 = reinterpret_cast<float4*>(M[q]), *v4 = reinterpret_cast<float4*>(V[q]);
                // up to 4 items per trip, all their loads issued before any arithmetic (one L2 latency per trip, not per item)
                const int n4 = cnt[q] >> 2;
                for (int i = first + stride; i < n4; i += 4 * stride) {
                    float4 a[4], b[4], c[4], d[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int idx = i + u * stride;
                        if (idx < n4) { a[u] = ld_grad4<GSMEM>(g4 + idx); b[u] = __ldcg(p4 + idx); c[u] = __ldcg(m4 + idx); d[u] = __ldcg(v4 + idx); }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int idx = i + u * stride;
                        if (idx < n4) {
                            adam_step<FAST>(b[u].x, c[u].x, d[u].x, a[u].x * coef, b1, b2, eps, as, inv_bc2);
                            adam_step<FAST>(b[u].y, c[u].y, d[u].y, a[u].y * coef, b1, b2, eps, as, inv_bc2);
                            adam_step<FAST>(b[u].z, c[u].z, d[u].z, a[u].z * coef, b1, b2, eps, as, inv_bc2);
                            adam_step<FAST>(b[u].w, c[u].w, d[u].w, a[u].w * coef, b1, b2, eps, as, inv_bc2);
                            p4[idx] = b[u]; m4[idx] = c[u]; v4[idx] = d[u];
                        }
                    }
                }
            } else {
                for (int i = first; i < cnt[q]; i += stride) {
                    float pi = __ldcg(P[q] + i), mi = __ldcg(M[q] + i), vi = __ldcg(V[q] + i);
                    adam_step<FAST>(pi, mi, vi, ld_grad<GSMEM>(g + goff[q] + i) * coef, b1, b2, eps, as, inv_bc2);
                    P[q][i] = pi; M[q][i] = mi; V[q][i] = vi;
                }
            }
        }
    }
}

