// PPO minibatch update on tcgen05 for the reference's default net shape: S -> 64 -> 64 -> OUT GELU actor and critic
// (S <= 11, OUT <= 8; BASELINE configs[1]: 3 -> 64 -> 64 -> 1).  Same contract, arguments and results as update.cu
// (reference AgentPPO.update_objectives, elegantrl/agents/AgentPPO.py:173-205; AgentBase.optimizer_backward,
// elegantrl/agents/AgentBase.py:239-248); every other shape keeps the generic FP32-pipe kernels of update.cu.
//
// A CTA works on tiles of 128 sampled transitions of ONE net (grid = CTAs x 2 nets); 256 threads: warps w and w + 4 share the
// TMEM lane quarter w % 4 (sample = lane) and split the 64 feature columns, so two threads work on every sample.  Every dense
// contraction of the forward AND backward pass is a tcgen05 tile (3xTF32: hi / lo planes, fp32 accumulate in TMEM):
//   L1   Z1 [128 x 64] = x~ B1^T            x~ = [x_hi, 1, x_lo, 0] (bias folded), K = round_up(2 S + 1, 8)      SS
//   L2   Z2 [128 x 64] = H1 W2^T            H1 = GELU(Z1) as hi / lo planes in tensor memory (b2 added on read)   TS
//   dH1  [128 x 64]    = dZ2 W2             dZ2 planes in tensor memory; W2's backward (MN-major) image          TS
//   G2   [ 64 x 64]    = dZ2^T H1           = dW2, in two passes of 32 columns: contraction over the SAMPLES;
//   G1   [ 64 x N1]    = dZ1^T [X  | 1]     = [dW1 | db1]  both operands are row-written MN-major images          SS
// GELU / GELU' (packed pairs), the head (64 -> OUT), the PPO loss and its derivative run on CUDA cores, thread = sample; the
// head's weight gradient is reduced with shuffles.
//   * batch_size <= 128 (the Config default): the launch is PERSISTENT -- all minibatches of update_net in one launch of two
//     CTAs that never synchronise with each other; parameters and Adam moments stay resident in tensor memory between the
//     minibatches (opt_load / opt_apply_resident), the gradient lives in shared memory, the operand images of minibatch u + 1
//     are written from the registers of minibatch u's Adam step, and the sampled record of minibatch u + 1 is requested
//     during the backward pass of minibatch u.  Env shards: the exchange with the peer GPUs happens in here too (px_*).
//   * larger minibatches: one launch per minibatch, at most 74 CTAs per net, each walking several tiles with the weight-gradient
//     UMMAs accumulating in tensor memory; one RED.ADD pass per CTA into the flat buffer, last CTA of a net applies clip + Adam
//     (apply_net, update_common.cuh).
#include <stdlib.h>
#include <string.h>

#include "tc_train.cuh"
#include "update_common.cuh"

namespace {

using namespace tctrain;

constexpr int kT = 128;    // samples per tile = TMEM lanes
constexpr int kNT = 256;   // threads: TWO per sample (warp w and w + 4 share a lane quarter and split the 64 columns)
constexpr int kMaxS = 11, kMaxOut = 8, kK1Max = 24;
constexpr int kAdamTab = 64;
constexpr int kMaxCtasPerNet = 74;   // 148 SMs / 2 nets

// ---- dynamic shared memory map (bytes)
constexpr int kOffW2 = 0;                                      // W2 hi / lo K-major images (forward)
constexpr int kOffWB = kOffW2 + 2 * kWPlaneBytes;              // W2 hi / lo backward images (MN-major)
constexpr int kOffGA = kOffWB + 2 * kWPlaneBytes;              // dZ rows (hi / lo): A operand of G2, then of G1
constexpr int kOffGB = kOffGA + 2 * kGAPlaneBytes;             // ONE 32-column group (hi / lo): H1[:, 0:32], H1[:, 32:64], [X | 1]
constexpr int kA1Bytes = kT * kK1Max * 4;
constexpr int kOffA1 = kOffGB + 2 * kGroupPlaneBytes;          // x~ rows, K-major
constexpr int kB1PlaneBytes = kHid * kK1Max * 4;
constexpr int kOffB1 = kOffA1 + kA1Bytes;                      // layer-1 B planes
constexpr int kOffSmall = kOffB1 + 2 * kB1PlaneBytes;
constexpr int kSmW3 = 0, kSmB3 = 512, kSmStd = 520, kSmAvg = 528, kSmSd = 544, kSmGW3 = 560, kSmGB3 = 1072, kSmGStd = 1080,
              kSmB2 = 1088, kSmGB2 = 1152, kSmW0 = 1216, kSmB0 = 1216 + 64 * kMaxS, kSmallFloats = kSmB0 + 64;
constexpr int kOffPart = kOffSmall + kSmallFloats * 4;   // head partial sums of the two column halves: [2][128][OUTC <= 8]
constexpr int kOffBar = kOffPart + 2 * kT * kMaxOut * 4;
constexpr int kSmemBytes = kOffBar + 16;
static_assert(kSmemBytes <= 226 * 1024, "shared memory budget");
static_assert(kOffGA % 1024 == 0 && kOffGB % 1024 == 0 && kOffWB % 1024 == 0, "MN-major images: swizzle-atom alignment");

// ---- tensor memory columns
constexpr uint32_t cZ1 = 0, cPhi = 64, cPlo = 128, cZ2 = 192, cG2 = 256, cG1 = 320;
// Persistent launch: the net's parameters and Adam moments stay RESIDENT in tensor memory between the minibatches (the 176
// columns behind the accumulators).  Thread (row, hf) owns W2 items 4 (tid + 256 q) + e (q, e < 4) and the small-tensor items
// tid + 256 q (q < 6; index space W0 | b0 | b1 | W3 | b3 | action_std_log); TMEM lane = row, columns
// cOpt + 48 kind + 24 hf + [0, 16) (W2) / + [16, 22) (small), kind = 0 parameter, 1 exp_avg, 2 exp_avg_sq.
constexpr uint32_t cOpt = 336, kOptKind = 48, kOptHalf = 24, kOptSmall = 16;
static_assert(cOpt + 3 * kOptKind <= 512, "tensor memory columns");

// ---- peer-memory exchange (env-sharded update): system-scope release / acquire flags, relaxed system-scope data loads
DEV void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
DEV uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
DEV float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
DEV float4 ld_relaxed_sys_v4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
DEV uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
constexpr int kPxStatFloats = 8;   // 4 doubles at the head of every exchange buffer
DEV int px_segment(int numel) { return (numel + 4 + 3) & ~3; }   // gradient + 3 loss partial sums + padding
// One exchange round of channel `chan` (0 = statistics, 1 = actor gradient, 2 = critic gradient): everything this CTA stored
// into its own buffer becomes visible to the peers, and theirs to this CTA.  Flags only ever grow (epochs), so there is no
// reset and no ABA; a peer that does not answer within ~2 s raises the error word of the workspace header instead of
// hanging the GPU.
DEV void px_round(const b200rl_peer_exchange& px, int chan, uint32_t value, WorkspaceHeader* hdr) {
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < px.world) st_release_sys(px.flags[threadIdx.x] + chan * B200RL_MAX_PEERS + px.rank, value);
    if ((int)threadIdx.x < px.world) {
        const uint32_t* f = px.flags[px.rank] + chan * B200RL_MAX_PEERS + threadIdx.x;
        const uint64_t t0 = global_timer_ns();
        while ((int32_t)(ld_acquire_sys(f) - value) < 0) {
            if (global_timer_ns() - t0 > 2000000000ull) { atomicExch(&hdr->pad[0], 1u); break; }
        }
    }
    __syncthreads();
}

// Parameters of one net -> operand images and the small shared-memory tables.  Every load is issued before anything is stored
// (ONE L2 latency per minibatch, not one per element).  Not inlined: its 43 staging registers would otherwise be live across
// the minibatch loop's already full register file.
__device__ __noinline__ void stage_params(const b200rl_net& net, float* small, uint32_t w2_hi, uint32_t w2_lo, uint32_t wb_hi,
                                          uint32_t wb_lo, int tid, int S, int OUT, bool gaussian) {
    {
            float4 w[4];   // W2: 4096 floats = 4 float4 per thread
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = __ldcg(reinterpret_cast<const float4*>(net.weight[1]) + tid + kNT * q);
            // the small tensors as one index space: W0 [64 x S] | b0 | b1 | W3 [OUT x 64] | b3 | action_std_log
            const int nW0 = kHid * S, nW3 = OUT * kHid;
            const int total = nW0 + 2 * kHid + nW3 + OUT + (gaussian ? OUT : 0);
            float sv[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                int idx = tid + kNT * q;
                const float* ptr = nullptr;
                if (idx < total) {
                    if (idx < nW0) ptr = net.weight[0] + idx;
                    else if ((idx -= nW0) < kHid) ptr = net.bias[0] + idx;
                    else if ((idx -= kHid) < kHid) ptr = net.bias[1] + idx;
                    else if ((idx -= kHid) < nW3) ptr = net.weight[2] + idx;
                    else if ((idx -= nW3) < OUT) ptr = net.bias[2] + idx;
                    else ptr = net.action_std_log + (idx - OUT);
                }
                sv[q] = ptr ? __ldcg(ptr) : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {   // forward (K-major) and backward (MN-major) images of W2, hi / lo planes
                const int i = 4 * (tid + kNT * q), n = i >> 6, k = i & 63;
                const float v4[4] = {w[q].x, w[q].y, w[q].z, w[q].w};
                float h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) { h[e] = tc05::tf32_hi(v4[e]); l[e] = v4[e] - h[e]; }
                const uint32_t offf = tc05::operand_offset(n, k, kHid);   // 4 consecutive k: one 16-byte chunk
                st_shared_v4(w2_hi + offf, h[0], h[1], h[2], h[3]);
                st_shared_v4(w2_lo + offf, l[0], l[1], l[2], l[3]);
                // backward image: MN = k (contiguous), K = n: 4 consecutive k = one 16-byte unit of row n
                const uint32_t offb = mn_swizzle((uint32_t)(k >> 5) * kWbLBO + (uint32_t)(n >> 2) * kMnSBO + (uint32_t)(n & 3) * 128 + (uint32_t)(k & 31) * 4);
                st_shared_v4(wb_hi + offb, h[0], h[1], h[2], h[3]);
                st_shared_v4(wb_lo + offb, l[0], l[1], l[2], l[3]);
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                int idx = tid + kNT * q;
                if (idx < total) {
                    if (idx < nW0) small[kSmW0 + idx] = sv[q];
                    else if ((idx -= nW0) < kHid) small[kSmB0 + idx] = sv[q];
                    else if ((idx -= kHid) < kHid) { small[kSmB2 + idx] = sv[q]; small[kSmGB2 + idx] = 0.0f; }
                    else if ((idx -= kHid) < nW3) { small[kSmW3 + idx] = sv[q]; small[kSmGW3 + idx] = 0.0f; }
                    else if ((idx -= nW3) < OUT) { small[kSmB3 + idx] = sv[q]; small[kSmGB3 + idx] = 0.0f; }
                    else small[kSmStd + (idx - OUT)] = sv[q];
                }
            }
            if (tid < OUT) small[kSmGStd + tid] = 0.0f;
            if (tid < S) {
                small[kSmAvg + tid] = net.state_avg ? net.state_avg[tid] : 0.0f;
                small[kSmSd + tid] = net.state_std ? net.state_std[tid] + 1e-4f : 1.0f;
            }
        }
}

// Ordered sum of the ranks' gradient buffers into shared memory.  All loads of one peer's buffer are in flight together (one
// NVLink latency per peer instead of one per 16 bytes); the ranks are added in rank order, the same on every GPU.
__device__ __noinline__ void px_reduce(const b200rl_peer_exchange& px, int px_off, int numel, float* g_smem) {
    constexpr int kIt = 6;   // 6 x 256 threads x 4 floats = 6 144 >= the largest eligible net (5 456 parameters)
    const int tid = threadIdx.x, n4 = numel >> 2;
    float4 acc[kIt];
    for (int r = 0; r < px.world; ++r) {
        const float* src = px.data[r] + px_off;
        float4 v[kIt];
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
            const int i4 = tid + kNT * it;
            v[it] = i4 < n4 ? ld_relaxed_sys_v4(src + 4 * i4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < kIt; ++it) {
            if (r == 0) acc[it] = v[it];
            else { acc[it].x += v[it].x; acc[it].y += v[it].y; acc[it].z += v[it].z; acc[it].w += v[it].w; }
        }
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int i4 = tid + kNT * it;
        if (i4 < n4) reinterpret_cast<float4*>(g_smem)[i4] = acc[it];
    }
    for (int i = (numel & ~3) + tid; i < numel; i += kNT) {
        float a = ld_relaxed_sys(px.data[0] + px_off + i);
        for (int r = 1; r < px.world; ++r) a += ld_relaxed_sys(px.data[r] + px_off + i);
        g_smem[i] = a;
    }
}

// clip + Adam of one net, out of line for the same reason (its batched loads use 64 + registers of their own)
__device__ __noinline__ void apply_from_global(const b200rl_net& net, const b200rl_adam& opt, const AdamScalars& as, const float* g, int numel,
                                               float clip_grad_norm, float* red) {
    apply_net<kNT, false, true>(net, opt, as, g, numel, clip_grad_norm, red);
}

// ---- resident optimizer state (persistent launch)
struct SmallItem { float* p; float* m; float* v; int g, sm, gz; };   // global pointers; flat-gradient, `small` and accumulator offsets
DEV bool small_item(const b200rl_net& net, const b200rl_adam& opt, int idx, int S, int OUT, bool gaussian, SmallItem& it) {
    const int nW0 = kHid * S, nW3 = OUT * kHid;
    const int oB0 = nW0, oW1 = oB0 + kHid, oB1 = oW1 + kHid * kHid, oW2 = oB1 + kHid, oB2 = oW2 + nW3, oStd = oB2 + OUT;
    it.gz = -1;
    if (idx < nW0) { it.p = net.weight[0] + idx; it.m = opt.exp_avg_w[0] + idx; it.v = opt.exp_avg_sq_w[0] + idx; it.g = idx; it.sm = kSmW0 + idx; return true; }
    if ((idx -= nW0) < kHid) { it.p = net.bias[0] + idx; it.m = opt.exp_avg_b[0] + idx; it.v = opt.exp_avg_sq_b[0] + idx; it.g = oB0 + idx; it.sm = kSmB0 + idx; return true; }
    if ((idx -= kHid) < kHid) { it.p = net.bias[1] + idx; it.m = opt.exp_avg_b[1] + idx; it.v = opt.exp_avg_sq_b[1] + idx; it.g = oB1 + idx; it.sm = kSmB2 + idx; it.gz = kSmGB2 + idx; return true; }
    if ((idx -= kHid) < nW3) { it.p = net.weight[2] + idx; it.m = opt.exp_avg_w[2] + idx; it.v = opt.exp_avg_sq_w[2] + idx; it.g = oW2 + idx; it.sm = kSmW3 + idx; it.gz = kSmGW3 + idx; return true; }
    if ((idx -= nW3) < OUT) { it.p = net.bias[2] + idx; it.m = opt.exp_avg_b[2] + idx; it.v = opt.exp_avg_sq_b[2] + idx; it.g = oB2 + idx; it.sm = kSmB3 + idx; it.gz = kSmGB3 + idx; return true; }
    if (gaussian && (idx -= OUT) < OUT) { it.p = net.action_std_log + idx; it.m = opt.exp_avg_std + idx; it.v = opt.exp_avg_sq_std + idx; it.g = oStd + idx; it.sm = kSmStd + idx; it.gz = kSmGStd + idx; return true; }
    return false;
}
// parameters + moments: global -> tensor memory, once per launch
__device__ __noinline__ void opt_load(const b200rl_net& net, const b200rl_adam& opt, uint32_t tl, int tid, int S, int OUT, bool gaussian) {
    const uint32_t base = tl + cOpt + (uint32_t)(tid >> 7) * kOptHalf;
    const float* w2[3] = {net.weight[1], opt.exp_avg_w[1], opt.exp_avg_sq_w[1]};
#pragma unroll
    for (int kind = 0; kind < 3; ++kind) {
        uint32_t r[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 w = __ldcg(reinterpret_cast<const float4*>(w2[kind]) + tid + kNT * q);
            r[4 * q] = __float_as_uint(w.x); r[4 * q + 1] = __float_as_uint(w.y); r[4 * q + 2] = __float_as_uint(w.z); r[4 * q + 3] = __float_as_uint(w.w);
        }
        tc05::tmem_st_32x32b_x16(base + kind * kOptKind, r);
        uint32_t sr[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            SmallItem it;
            sr[q] = 0u;
            if (q < 6 && small_item(net, opt, tid + kNT * q, S, OUT, gaussian, it))
                sr[q] = __float_as_uint(__ldcg(kind == 0 ? it.p : (kind == 1 ? it.m : it.v)));
        }
        tc05::tmem_st_32x32b_x8(base + kind * kOptKind + kOptSmall, sr);
    }
    tc05::tmem_st_wait();
}
// clip_grad_norm_ + Adam.step on the resident state, and the next minibatch's operand images straight from the registers that
// hold the new parameters: no global round trip per minibatch (the L2 latencies of p / m / v loads, their stores and the
// re-read by the staging were 30 k of the 78 k cycles of a minibatch, profiles/r02_v8_update_tc_phases.log).  `last`: write
// everything back to the caller's tensors.
__device__ __noinline__ void opt_apply_resident(const b200rl_net& net, const b200rl_adam& opt, const AdamScalars& as, const float* g, int numel,
                                                float clip_grad_norm, float* red, float* small, uint32_t w2_hi, uint32_t w2_lo, uint32_t wb_hi,
                                                uint32_t wb_lo, uint32_t tl, int tid, int S, int OUT, bool gaussian, bool last) {
    const float coef = clip_coef<kNT, true>(g, numel, clip_grad_norm, red);
    const float b1 = opt.beta1, b2 = opt.beta2, eps = opt.eps, inv_bc2 = 1.0f / as.bc2_sqrt;
    const uint32_t base = tl + cOpt + (uint32_t)(tid >> 7) * kOptHalf;
    const int oW1 = kHid * S + kHid;
    {   // ---- W2
        float p[16], m[16], v[16];
        tc05::tmem_ld_32x32b_x16(base, p);
        tc05::tmem_ld_32x32b_x16(base + kOptKind, m);
        tc05::tmem_ld_32x32b_x16(base + 2 * kOptKind, v);
        tc05::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 gg = *reinterpret_cast<const float4*>(g + oW1 + 4 * (tid + kNT * q));
            adam_step<true>(p[4 * q], m[4 * q], v[4 * q], gg.x * coef, b1, b2, eps, as, inv_bc2);
            adam_step<true>(p[4 * q + 1], m[4 * q + 1], v[4 * q + 1], gg.y * coef, b1, b2, eps, as, inv_bc2);
            adam_step<true>(p[4 * q + 2], m[4 * q + 2], v[4 * q + 2], gg.z * coef, b1, b2, eps, as, inv_bc2);
            adam_step<true>(p[4 * q + 3], m[4 * q + 3], v[4 * q + 3], gg.w * coef, b1, b2, eps, as, inv_bc2);
        }
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(p[i]);
        tc05::tmem_st_32x32b_x16(base, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(m[i]);
        tc05::tmem_st_32x32b_x16(base + kOptKind, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
        tc05::tmem_st_32x32b_x16(base + 2 * kOptKind, r);
#pragma unroll
        for (int q = 0; q < 4; ++q) {   // forward (K-major) and backward (MN-major) images of the new W2, hi / lo planes
            const int i = 4 * (tid + kNT * q), n = i >> 6, k = i & 63;
            float h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { h[e] = tc05::tf32_hi(p[4 * q + e]); l[e] = p[4 * q + e] - h[e]; }
            const uint32_t offf = tc05::operand_offset(n, k, kHid);
            st_shared_v4(w2_hi + offf, h[0], h[1], h[2], h[3]);
            st_shared_v4(w2_lo + offf, l[0], l[1], l[2], l[3]);
            const uint32_t offb = mn_swizzle((uint32_t)(k >> 5) * kWbLBO + (uint32_t)(n >> 2) * kMnSBO + (uint32_t)(n & 3) * 128 + (uint32_t)(k & 31) * 4);
            st_shared_v4(wb_hi + offb, h[0], h[1], h[2], h[3]);
            st_shared_v4(wb_lo + offb, l[0], l[1], l[2], l[3]);
            if (last) {
                reinterpret_cast<float4*>(net.weight[1])[tid + kNT * q] = make_float4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
                reinterpret_cast<float4*>(opt.exp_avg_w[1])[tid + kNT * q] = make_float4(m[4 * q], m[4 * q + 1], m[4 * q + 2], m[4 * q + 3]);
                reinterpret_cast<float4*>(opt.exp_avg_sq_w[1])[tid + kNT * q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
        }
    }
    {   // ---- the small tensors
        float p[8], m[8], v[8];
        tc05::tmem_ld_32x32b_x8(base + kOptSmall, p);
        tc05::tmem_ld_32x32b_x8(base + kOptKind + kOptSmall, m);
        tc05::tmem_ld_32x32b_x8(base + 2 * kOptKind + kOptSmall, v);
        tc05::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            SmallItem it;
            if (small_item(net, opt, tid + kNT * q, S, OUT, gaussian, it)) {
                adam_step<true>(p[q], m[q], v[q], g[it.g] * coef, b1, b2, eps, as, inv_bc2);
                small[it.sm] = p[q];
                if (it.gz >= 0) small[it.gz] = 0.0f;   // the head's gradient accumulators of the next minibatch
                if (last) { *it.p = p[q]; *it.m = m[q]; *it.v = v[q]; }
            }
        }
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(p[i]);
        tc05::tmem_st_32x32b_x8(base + kOptSmall, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(m[i]);
        tc05::tmem_st_32x32b_x8(base + kOptKind + kOptSmall, r);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(v[i]);
        tc05::tmem_st_32x32b_x8(base + 2 * kOptKind + kOptSmall, r);
    }
    tc05::tmem_st_wait();
}

#define TC_MARK(i) do { if (A.profile && ni == 0 && tid == 0 && u == U - 1) reinterpret_cast<long long*>(A.hdr)[8 + (i)] = clock64(); } while (0)

// OUTC / SC: compile-time capacities (>= the nets' OUT / S) of the per-thread head / state arrays.  The body is straight-line
// code executed once per minibatch by four warps: with the capacities at their maxima (8, 11) it was 15 k instructions (240 KB)
// and instruction fetch was the top stall (profiles/r02_v7_update_tc_phases_before_code_size_fix.log); sized to the net it fits
// the instruction cache.
template <int OUTC, int SC>
__global__ void __launch_bounds__(kNT, 1) ppo_update_tc_kernel(const __grid_constant__ UpdateArgs A) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ float red[32];
    __shared__ AdamScalars s_adam;
    __shared__ AdamScalars s_adam_tab[kAdamTab];   // bias corrections of the first kAdamTab minibatches, computed in parallel up front
    __shared__ int s_last;
    __shared__ uint32_t tmem_slot;
    float* small = reinterpret_cast<float*>(smem + kOffSmall);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & (kT - 1), hf = tid >> 7, quarter = warp & 3;   // sample row, column half, TMEM lane quarter
    const int ni = blockIdx.y, tile = blockIdx.x;
    const int n_tiles = (A.local_batch + kT - 1) / kT;   // tiles of one minibatch (persistent launch: 1)
    const b200rl_net& net = A.net[ni];
    const int S = net.dims[0], OUT = net.dims[3];
    const int K1 = (2 * S + 1 + 7) & ~7, N1 = (S + 1 + 7) & ~7;
    const bool persistent = A.update_times > 0;
    const int U = persistent ? A.update_times : 1;
    const bool discrete = A.buf.discrete_actions != 0;
    const bool gaussian = ni == 0 && !discrete;
    const int H = A.buf.horizon_len, N = A.buf.num_envs;
    const bool packed = (H == 0);
    const int rec_act = A.net[0].dims[0], rec_tail = (A.net[0].dims[0] + A.net[0].dims[3] + 3) & ~3, rec = rec_tail + 4;
    const int flags = A.hp.flags;
    const float inv_bsz = 1.0f / (float)A.global_batch;

    // ---- one-time setup
    if (warp == 0) tc05::tmem_alloc<512>(&tmem_slot);
    if (tid == 32) { tc05::mbar_init(bar, 1); tc05::mbar_fence_init(); }
    for (int i = tid; i < kA1Bytes / 4; i += kNT) reinterpret_cast<float*>(smem + kOffA1)[i] = 0.0f;
    if (A.update_times > 0 && tid < kAdamTab && tid < A.update_times) {
        // torch.optim.Adam bias corrections, in double like torch: one thread per minibatch instead of a serial pow() in front of
        // every apply
        const double step = (double)(A.opt[ni].step + tid + 1);
        s_adam_tab[tid].step_size = (float)((double)A.opt[ni].lr / (1.0 - pow((double)A.opt[ni].beta1, step)));
        s_adam_tab[tid].bc2_sqrt = (float)sqrt(1.0 - pow((double)A.opt[ni].beta2, step));
    }
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);   // this warp's lane quarter
    const uint32_t w2_hi = tc05::smem_u32(smem + kOffW2), w2_lo = w2_hi + kWPlaneBytes;
    const uint32_t wb_hi = tc05::smem_u32(smem + kOffWB), wb_lo = wb_hi + kWPlaneBytes;
    const uint32_t ga_addr = tc05::smem_u32(smem + kOffGA), gb_addr = tc05::smem_u32(smem + kOffGB);
    uint32_t phase = 0;
    double acc_c = 0.0, acc_s = 0.0, acc_e = 0.0;   // thread 0: loss sums over the minibatches (persistent mode)
    if (persistent) opt_load(net, A.opt[ni], tl, tid, S, OUT, gaussian);

    // advantage normalisation (reference :149): from the caller's statistics, or -- env-sharded -- reduced here over the shards
    __shared__ float s_stats[2];
    // env shards: px_on == 1: gradient all-reduce per minibatch; px_on == 2: the sampled RECORDS of all minibatches are exchanged
    // once per cycle (minibatch indices do not depend on the parameters), every rank then runs the same full-batch update
    const bool sharded = A.px_on == 1, gathered = A.px_on == 2;
    bool normalise = A.buf.adv_stats != nullptr;
    if ((sharded || gathered) && ni == 0) {
        if (tid < 4) reinterpret_cast<double*>(A.px.data[A.px.rank])[tid] = A.stat_sums[tid];
        px_round(A.px, 0, A.px.epoch + 1, A.hdr);
        if (tid == 0) {
            double sums[4] = {0.0, 0.0, 0.0, 0.0};
            for (int r = 0; r < A.px.world; ++r)
                for (int k = 0; k < 4; ++k) {
                    const float lo = ld_relaxed_sys(A.px.data[r] + 2 * k), hi = ld_relaxed_sys(A.px.data[r] + 2 * k + 1);
                    sums[k] += __hiloint2double(__float_as_int(hi), __float_as_int(lo));
                }
            // b200rl_adv_stats' arithmetic (gae.cu)
            const double mean = sums[0] / A.count_all;
            const bool full = A.count_lat == 0.0;
            const double cnt = full ? A.count_all : A.count_lat;
            const double m = (full ? sums[0] : sums[1]) / cnt, sq = full ? sums[3] : sums[2];
            const double var = (sq - cnt * m * m) / (cnt - 1.0);
            float sd = (float)sqrt(var > 0.0 ? var : 0.0);
            if (cnt < 2.0) sd = nanf("");
            s_stats[0] = (float)mean; s_stats[1] = sd;
            if (A.stats_out) { A.stats_out[0] = (float)mean; A.stats_out[1] = sd; A.stats_out[2] = 1.0f / (sd + 1e-5f); A.stats_out[3] = 0.0f; }
        }
        normalise = true;
        __syncthreads();
    } else if (normalise) {
        if (tid < 2) s_stats[tid] = A.buf.adv_stats[tid];
        __syncthreads();
    }
    // record gather: this CTA packs ITS rank's share of every minibatch of the cycle -- {state, action, pad, unmask, logprob,
    // normalised advantage, reward_sum}, the layout of b200rl_pack_minibatches -- into the own exchange buffer (one region per
    // net and per cycle parity: a rank can only be one cycle ahead of a peer, cf. the flag argument in px_round), raises its flag
    // and waits for the peers; the minibatch loop then gathers its 128 samples from all ranks' buffers with peer loads.
    const int lb = A.px_local_batch;
    const int px_rec_off = kPxStatFloats + ((int)(A.px.reserved & 1u) * 2 + ni) * (U * lb * rec);
    if (gathered) {
        float* own = A.px.data[A.px.rank] + px_rec_off;
        const int Adim = discrete ? 1 : A.net[0].dims[3];
        for (int i = tid; i < U * lb; i += kNT) {
            const int uu = i / lb, sl = i - uu * lb;
            const int64_t id = A.ids ? A.ids[i] : sample_index(A.seed, A.draw + (uint64_t)uu, (uint32_t)sl, (uint64_t)H * (uint64_t)N);
            const int64_t tn = (id % H) * N + id / H;
            float* r = own + (size_t)i * rec;
            for (int k = 0; k < S; ++k) r[k] = A.buf.states[tn * S + k];
            if (discrete) r[rec_act] = (float)reinterpret_cast<const int32_t*>(A.buf.actions)[tn];
            else for (int a = 0; a < Adim; ++a) r[rec_act + a] = A.buf.actions[tn * Adim + a];
            for (int k = rec_act + Adim; k < rec_tail; ++k) r[k] = 0.0f;
            float adv = A.buf.advantages[tn];
            if (normalise) adv = (adv - s_stats[0]) / (s_stats[1] + 1e-5f);
            r[rec_tail + 0] = A.buf.unmasks[tn] ? 1.0f : 0.0f;
            r[rec_tail + 1] = A.buf.logprobs[tn];
            r[rec_tail + 2] = adv;
            r[rec_tail + 3] = A.buf.reward_sums[tn];
        }
        px_round(A.px, 1 + ni, A.px.reserved + 1u, A.hdr);
    }

    // flat gradient layout of this net: W0 [64 x S], b0, W1 [64 x 64], b1, W2 [OUT x 64], b2, (action_std_log)
    const int oB0 = kHid * S, oW1 = oB0 + kHid, oB1 = oW1 + kHid * kHid, oW2 = oB1 + kHid, oB2 = oW2 + OUT * kHid, oStd = oB2 + OUT;
    float* const g_local = A.grads + A.grad_off[ni];
    const int numel = A.grad_numel[ni];
    const int px_off0 = kPxStatFloats + (ni ? px_segment(A.grad_numel[0]) : 0);
    const int px_parity_stride = px_segment(A.grad_numel[0]) + px_segment(A.grad_numel[1]);

    // one sampled transition of minibatch uu for this thread's slot (both threads of a sample load it)
    auto load_record = [&](int uu, int tt, float& um, float& lp_old, float& adv, float& rs, float (&act)[OUTC], float (&x)[SC]) {
        um = 0.f; lp_old = 0.f; adv = 0.f; rs = 0.f;
#pragma unroll
        for (int a = 0; a < OUTC; ++a) act[a] = 0.0f;
#pragma unroll
        for (int k = 0; k < SC; ++k) x[k] = 0.0f;
        const int slot = tt * kT + row;
        if (slot >= A.local_batch) return;
        const int Adim = discrete ? 1 : A.net[0].dims[3];
        if (gathered) {   // sample `slot` of the global minibatch = record (uu, slot % lb) of rank slot / lb
            const float* recp = A.px.data[slot / lb] + px_rec_off + (size_t)(uu * lb + slot % lb) * rec;
            const float4 tail = ld_relaxed_sys_v4(recp + rec_tail);
            um = tail.x; lp_old = tail.y; adv = tail.z; rs = tail.w;
#pragma unroll
            for (int k = 0; k < SC; ++k) if (k < S) x[k] = ld_relaxed_sys(recp + k);
#pragma unroll
            for (int a = 0; a < OUTC; ++a) if (a < Adim) act[a] = ld_relaxed_sys(recp + rec_act + a);
            return;
        }
        const int64_t* ids_u = A.ids ? A.ids + (persistent ? (size_t)uu * A.local_batch : 0) : nullptr;
        const uint64_t draw = A.draw + (uint64_t)uu;
        if (packed) {
            const int64_t sampled = ids_u ? ids_u[slot] : (int64_t)draw * A.local_batch + slot;
            const float* recp = A.buf.states + sampled * rec;
            const float4 tail = *reinterpret_cast<const float4*>(recp + rec_tail);
            um = tail.x; lp_old = tail.y; adv = tail.z; rs = tail.w;
#pragma unroll
            for (int k = 0; k < SC; ++k) if (k < S) x[k] = recp[k];
#pragma unroll
            for (int a = 0; a < OUTC; ++a) if (a < Adim) act[a] = recp[rec_act + a];
            return;
        }
        const uint64_t range = (uint64_t)H * (uint64_t)N;
        const int64_t sampled = ids_u ? ids_u[slot] : sample_index(A.seed, draw, (uint32_t)slot, range);
        int64_t tn;
        if (range <= 0xFFFFFFFFull) {   // 32-bit division: the 64-bit one is a ~150-instruction dependent chain in front of the loads
            const uint32_t s32 = (uint32_t)sampled, q = s32 / (uint32_t)H;
            tn = (int64_t)(s32 - q * (uint32_t)H) * N + q;
        } else {
            tn = (sampled % H) * N + sampled / H;
        }
        um = A.buf.unmasks[tn] ? 1.0f : 0.0f;
        lp_old = A.buf.logprobs[tn];
        adv = A.buf.advantages[tn];
        if (normalise) adv = (adv - s_stats[0]) / (s_stats[1] + 1e-5f);
        rs = A.buf.reward_sums[tn];
#pragma unroll
        for (int k = 0; k < SC; ++k) if (k < S) x[k] = A.buf.states[tn * S + k];
        if (discrete) act[0] = (float)reinterpret_cast<const int32_t*>(A.buf.actions)[tn];
        else {
#pragma unroll
            for (int a = 0; a < OUTC; ++a) if (a < Adim) act[a] = A.buf.actions[tn * Adim + a];
        }
    };
    float n_um = 0.f, n_lp = 0.f, n_adv = 0.f, n_rs = 0.f;   // the prefetched record of the next minibatch
    float n_act[OUTC];
    float n_x[SC];
#pragma unroll
    for (int a = 0; a < OUTC; ++a) n_act[a] = 0.0f;
#pragma unroll
    for (int k = 0; k < SC; ++k) n_x[k] = 0.0f;

    for (int u = 0; u < U; ++u) {
        // env-sharded: this minibatch's gradient goes to the own exchange buffer (two alternate so that a rank that is one
        // minibatch ahead never overwrites what a peer still reads); the reduced gradient lands in the workspace as usual
        // (parity of the GLOBAL minibatch count, so that consecutive minibatches alternate across calls as well: writing buffer p
        // at minibatch g needs the wait of g - 1 behind it, which every peer passes only after it has read g - 2 = the last use of p)
        const int px_off = px_off0 + (int)((A.px.epoch + (uint32_t)u) & 1u) * px_parity_stride;
        // persistent mode: the flat gradient of this minibatch lives in SHARED memory (the dZ row buffer is free once G1 is done);
        // env-sharded: the own exchange buffer, the reduced sum then goes to shared memory; several tiles: the global workspace
        float* const g_smem = reinterpret_cast<float*>(smem + kOffGA);
        float* const g = sharded ? A.px.data[A.px.rank] + px_off : (persistent ? g_smem : g_local);
        // ------------------------------------------------------------ parameters -> operand images (Adam rewrote them)
        // every load of the minibatch's parameters is issued before anything is stored: ONE L2 latency, not one per element
        TC_MARK(0);
        // (persistent launch: minibatch u > 0 finds the images of its parameters written by the previous minibatch's Adam step)
        if (!persistent || u == 0) stage_params(net, small, w2_hi, w2_lo, wb_hi, wb_lo, tid, S, OUT, gaussian);
        __syncthreads();   // staged small parameters (state_norm statistics, W0 / b0) are visible
        for (int i = tid; i < kHid * K1; i += kNT) {   // layer-1 B planes from the staged W0 / b0
            const int n = i / K1, k = i - n * K1;
            float full = 0.0f;
            if (k < S) full = small[kSmW0 + n * S + k];
            else if (k == S) full = small[kSmB0 + n];
            else if (k <= 2 * S) full = small[kSmW0 + n * S + (k - S - 1)];
            const float hi = tc05::tf32_hi(full);
            const uint32_t off = tc05::operand_offset(n, k, K1);
            *reinterpret_cast<float*>(smem + kOffB1 + off) = hi;                                            // [W_hi, b_hi, W_hi, 0]
            *reinterpret_cast<float*>(smem + kOffB1 + kB1PlaneBytes + off) = k <= S ? full - hi : 0.0f;     // [W_lo, b_lo, 0, 0]
        }
        // A minibatch of several tiles (one launch per minibatch): this CTA walks tiles tile0, tile0 + gridDim.x, ... with the
        // operand images staged ONCE; the weight-gradient accumulators in tensor memory (and the head's in shared memory) keep
        // accumulating over its tiles, so the flat gradient is added to the global buffer once per CTA, not once per tile.
        float tot_c = 0.0f, tot_s = 0.0f, tot_e = 0.0f;
        float x[SC];
        bool valid = false;
        for (int tt = tile; tt < n_tiles; tt += (int)gridDim.x) {
        const bool first_tile = tt == tile;
        const int slot = tt * kT + row;   // both threads of a sample gather its scalars (the loss is evaluated redundantly)
        valid = slot < A.local_batch;
        // ------------------------------------------------------------ gather (reference :178-187): ids -> (t, n).  The record of
        // every tile / minibatch but the first was requested during the backward pass of the previous one (the indices do not
        // depend on the parameters), so its HBM / NVLink latency and the index arithmetic are off the critical path
        float um, lp_old, adv, rs;
        float act[OUTC];
        if (u == 0 && first_tile) load_record(0, tt, um, lp_old, adv, rs, act, x);
        else {
            um = n_um; lp_old = n_lp; adv = n_adv; rs = n_rs;
#pragma unroll
            for (int a = 0; a < OUTC; ++a) act[a] = n_act[a];
#pragma unroll
            for (int k = 0; k < SC; ++k) x[k] = n_x[k];
        }
        TC_MARK(1);
        if (valid && net.state_avg) {
#pragma unroll
            for (int k = 0; k < SC; ++k) if (k < S) x[k] = (x[k] - small[kSmAvg + k]) / small[kSmSd + k];
        }
        if (hf == 0) {   // x~ row (K-major A operand of layer 1)
#pragma unroll
            for (int k = 0; k < SC; ++k) {
                if (k < S) {
                    const float hi = tc05::tf32_hi(x[k]);
                    *reinterpret_cast<float*>(smem + kOffA1 + tc05::operand_offset(row, k, K1)) = hi;
                    *reinterpret_cast<float*>(smem + kOffA1 + tc05::operand_offset(row, S + 1 + k, K1)) = x[k] - hi;
                }
            }
            *reinterpret_cast<float*>(smem + kOffA1 + tc05::operand_offset(row, S, K1)) = 1.0f;
        }
        tc05::fence_proxy_async_smem();
        tc05::fence_before_thread_sync();
        __syncthreads();

        TC_MARK(2);
        // ------------------------------------------------------------ layer 1 on the tensor core
        if (tid == 0) {
            tc05::fence_after_thread_sync();
            const uint32_t idesc = tc05::make_idesc_tf32(kT, kHid);
            const uint32_t sbo = (uint32_t)(K1 / 4) * 128;
            const uint32_t a1 = tc05::smem_u32(smem + kOffA1), b1 = tc05::smem_u32(smem + kOffB1);
            for (int p = 0; p < 2; ++p)
                for (int ks = 0; ks < K1 / 8; ++ks)
                    tc05::mma_tf32(tmem_base + cZ1, tc05::make_smem_desc_ex(a1 + ks * 256, 128, sbo),
                                   tc05::make_smem_desc_ex(b1 + p * kB1PlaneBytes + ks * 256, 128, sbo), idesc, p > 0 || ks > 0);
            tc05::mma_commit(bar);
        }
        tc05::mbar_wait(bar, phase & 1); ++phase;
        tc05::fence_after_thread_sync();

        TC_MARK(3);
        // ------------------------------------------------------------ H1 = GELU(Z1): TMEM planes (layer-2 A); columns 0..31 also as
        // rows of the group buffer (B operand of the first weight-gradient pass)
#pragma unroll 1
        for (int c = 2 * hf; c < 2 * hf + 2; ++c) {   // this thread's half of the row: columns [32 hf, 32 hf + 32)
            float z[16];
            tc05::tmem_ld_32x32b_x16(tl + cZ1 + 16 * c, z);
            tc05::tmem_ld_wait();
            gelu_only16(z);
            store_hi_lo_tmem(tl + cPhi + 16 * c, tl + cPlo + 16 * c, z);
            if (c < 2) store_hi_lo_rows(smem + kOffGB, smem + kOffGB + kGroupPlaneBytes, row, 16 * c, z);
        }
        tc05::tmem_st_wait();
        tc05::fence_proxy_async_smem();
        tc05::fence_before_thread_sync();
        __syncthreads();
        if (tid == 0) {
            tc05::fence_after_thread_sync();
            issue_linear_ts(tmem_base + cZ2, tmem_base + cPhi, tmem_base + cPlo, w2_hi, w2_lo, false);
            tc05::mma_commit(bar);
        }
        tc05::mbar_wait(bar, phase & 1); ++phase;
        tc05::fence_after_thread_sync();

        TC_MARK(4);
        // ------------------------------------------------------------ head (64 -> OUT) on CUDA cores
        float out[OUTC];
#pragma unroll
        for (int a = 0; a < OUTC; ++a) out[a] = 0.0f;
#pragma unroll 1
        for (int c = 2 * hf; c < 2 * hf + 2; ++c) {
            float z[16];
            tc05::tmem_ld_32x32b_x16(tl + cZ2 + 16 * c, z);
            tc05::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] += small[kSmB2 + 16 * c + j];
            gelu_only16(z);
#pragma unroll
            for (int a = 0; a < OUTC; ++a) {
                if (a < OUT) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) out[a] = fmaf(z[j], small[kSmW3 + a * kHid + 16 * c + j], out[a]);
                }
            }
        }
        {   // the two halves of a row meet through shared memory; both
            // threads then hold the full head output and evaluate the loss redundantly (fixed order: half 0 + half 1 + bias)
            float* part = reinterpret_cast<float*>(smem + kOffPart);
#pragma unroll
            for (int a = 0; a < OUTC; ++a) if (a < OUT) part[(hf * kT + row) * OUTC + a] = out[a];
            __syncthreads();
#pragma unroll
            for (int a = 0; a < OUTC; ++a) out[a] = a < OUT ? (part[row * OUTC + a] + part[(kT + row) * OUTC + a]) + small[kSmB3 + a] : 0.0f;
        }

        // ------------------------------------------------------------ loss and d loss / d output (update.cu grads_phase, same
        // arithmetic; reference :189-204, helloworld_PPO_single_file.py:332-340 behind the variant flags)
        float dout[OUTC];
#pragma unroll
        for (int a = 0; a < OUTC; ++a) dout[a] = 0.0f;
        float loss_c = 0.f, loss_s = 0.f, loss_e = 0.f;
        if (ni == 1) {
            const float err = out[0] - rs;
            float l, dl;
            if (flags & B200RL_PPO_SMOOTH_L1) {
                const float ae = fabsf(err);
                l = ae < 1.0f ? 0.5f * err * err : ae - 0.5f;
                dl = ae < 1.0f ? err : copysignf(1.0f, err);
            } else {
                l = err * err;
                dl = 2.0f * err;
            }
            loss_c = valid ? l * um : 0.0f;
            dout[0] = valid ? dl * um * inv_bsz : 0.0f;
        } else {
            float logp = 0.0f, ent = 0.0f, lse = 0.0f;
            const int act_idx = discrete ? min(max((int)act[0], 0), OUT - 1) : 0;
            if (discrete) {
                float m = -INFINITY, sum = 0.0f;
#pragma unroll
                for (int a = 0; a < OUTC; ++a) if (a < OUT) m = fmaxf(m, out[a]);
#pragma unroll
                for (int a = 0; a < OUTC; ++a) if (a < OUT) sum += expf(out[a] - m);
                lse = m + logf(sum);
#pragma unroll
                for (int a = 0; a < OUTC; ++a) {
                    if (a < OUT) {
                        const float lp = out[a] - lse;
                        ent -= expf(lp) * lp;
                        if (a == act_idx) logp = lp;
                    }
                }
            } else {
#pragma unroll
                for (int a = 0; a < OUTC; ++a) {
                    if (a < OUT) {
                        const float sd = expf(small[kSmStd + a]);
                        const float diff = act[a] - out[a];
                        const float lsd = logf(sd);
                        logp += -(diff * diff) / (2.0f * (sd * sd)) - lsd - kLogSqrt2Pi;
                        ent += 0.5f + kLogSqrt2Pi + lsd;
                    }
                }
            }
            const float ratio = expf(logp - lp_old);
            const float um_a = (flags & B200RL_PPO_ACTOR_UNMASKED) ? 1.0f : um;
            float surr, dsurr_dratio;
            if (flags & B200RL_PPO_MIN_CLIP) {
                const float rc = fminf(fmaxf(ratio, 1.0f - A.hp.ratio_clip), 1.0f + A.hp.ratio_clip);
                const float s1 = adv * ratio, s2 = adv * rc;
                const bool take1 = s1 <= s2;
                surr = take1 ? s1 : s2;
                dsurr_dratio = take1 ? adv : ((rc == ratio) ? adv : 0.0f);
            } else {
                const float kappa = adv > 0.0f ? 1.0f - A.hp.ratio_clip : 1.0f + A.hp.ratio_clip;
                surr = adv * ratio * kappa;
                dsurr_dratio = adv * kappa;
            }
            float dlogp_scale = ratio;
            if (flags & B200RL_PPO_A2C) {
                surr = adv * logp / (float)OUT;
                dsurr_dratio = adv / (float)OUT;
                dlogp_scale = 1.0f;
            }
            loss_s = valid ? surr * um_a : 0.0f;
            loss_e = valid ? ent * um_a : 0.0f;
            const float ent_sign = (flags & B200RL_PPO_ENTROPY_BONUS) ? -1.0f : 1.0f;
            const float gl = valid ? -(dsurr_dratio * dlogp_scale * um_a) * inv_bsz : 0.0f;
            const float ge = valid ? ent_sign * A.hp.lambda_entropy * um_a * inv_bsz : 0.0f;
            if (discrete) {
#pragma unroll
                for (int a = 0; a < OUTC; ++a) {
                    if (a < OUT) {
                        const float lp = out[a] - lse, pa = expf(lp);
                        dout[a] = gl * ((a == act_idx ? 1.0f : 0.0f) - pa) - ge * pa * (lp + ent);
                    }
                }
            } else {
#pragma unroll
                for (int a = 0; a < OUTC; ++a) {
                    if (a < OUT) {   // OUT is uniform over the CTA: the warp-wide reduction below is convergent
                        const float sd = expf(small[kSmStd + a]);
                        const float var = sd * sd;
                        const float diff = act[a] - out[a];
                        dout[a] = gl * diff / var;
                        const float dstd = warp_sum(gl * (diff * diff / var - 1.0f) + ge);
                        if (lane == 0 && hf == 0) atomicAdd(&small[kSmGStd + a], dstd);
                    }
                }
            }
        }
#pragma unroll
        for (int a = 0; a < OUTC; ++a) {
            if (a < OUT) {
                const float s = warp_sum(dout[a]);
                if (lane == 0 && hf == 0) atomicAdd(&small[kSmGB3 + a], s);
            }
        }

        TC_MARK(5);
        // the scalars of this minibatch's record are dead from here on: request the next one (consumed at the top of the loop)
        if (tt + (int)gridDim.x < n_tiles) load_record(u, tt + (int)gridDim.x, n_um, n_lp, n_adv, n_rs, n_act, n_x);
        else if (u + 1 < U) load_record(u + 1, tile, n_um, n_lp, n_adv, n_rs, n_act, n_x);
        // ------------------------------------------------------------ dZ2 = (dOut W3) * GELU'(Z2): TMEM planes + rows; dW3 by shuffles
        if (hf == 0) { tot_c += loss_c; tot_s += loss_s; tot_e += loss_e; }   // the loss sums count every sample once
#pragma unroll 1
        for (int c = 2 * hf; c < 2 * hf + 2; ++c) {
            float z[16], gz[16], dz[16];
            tc05::tmem_ld_32x32b_x16(tl + cZ2 + 16 * c, z);
            tc05::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float2 g2, dg2;
                gelu_and_grad2(make_float2(z[j] + small[kSmB2 + 16 * c + j], z[j + 1] + small[kSmB2 + 16 * c + j + 1]), g2, dg2);
                gz[j] = g2.x; gz[j + 1] = g2.y;
                float dh0 = 0.0f, dh1 = 0.0f;
#pragma unroll
                for (int a = 0; a < OUTC; ++a) {
                    if (a < OUT) {
                        dh0 = fmaf(dout[a], small[kSmW3 + a * kHid + 16 * c + j], dh0);
                        dh1 = fmaf(dout[a], small[kSmW3 + a * kHid + 16 * c + j + 1], dh1);
                    }
                }
                dz[j] = dh0 * dg2.x; dz[j + 1] = dh1 * dg2.y;
            }
            store_hi_lo_tmem(tl + cPhi + 16 * c, tl + cPlo + 16 * c, dz);
            store_hi_lo_rows(smem + kOffGA, smem + kOffGA + kGAPlaneBytes, row, 16 * c, dz);
            {   // db2 = column sums of dZ2
                float w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = dz[j];
                const float tot = warp_reduce16(w, lane);
                if (!(lane & 1)) atomicAdd(&small[kSmGB2 + 16 * c + ((lane >> 1) & 15)], tot);
            }
#pragma unroll
            for (int a = 0; a < OUTC; ++a) {
                if (a < OUT) {
                    float w[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[j] = dout[a] * gz[j];
                    const float tot = warp_reduce16(w, lane);
                    if (!(lane & 1)) atomicAdd(&small[kSmGW3 + a * kHid + 16 * c + ((lane >> 1) & 15)], tot);
                }
            }
        }
        tc05::tmem_st_wait();
        tc05::fence_proxy_async_smem();
        tc05::fence_before_thread_sync();
        __syncthreads();
        if (tid == 0) {
            tc05::fence_after_thread_sync();
            issue_linear_ts_backward(tmem_base + cZ2, tmem_base + cPhi, tmem_base + cPlo, wb_hi, wb_lo);   // dH1 over Z2
            issue_weight_grad(tmem_base + cG2, ga_addr, kGAPlaneBytes, gb_addr, kGroupPlaneBytes, 32, !first_tile);   // dW2[:, 0:32]
            tc05::mma_commit(bar);
        }
        tc05::mbar_wait(bar, phase & 1); ++phase;
        tc05::fence_after_thread_sync();
        TC_MARK(6);
        // second pass: H1[:, 32:64] (recomputed from Z1, still in tensor memory) into the same group buffer -> dW2[:, 32:64]
        if (hf == 1) {   // the threads that own columns 32..63
#pragma unroll 1
            for (int c = 2; c < 4; ++c) {
                float z[16];
                tc05::tmem_ld_32x32b_x16(tl + cZ1 + 16 * c, z);
                tc05::tmem_ld_wait();
                gelu_only16(z);
                store_hi_lo_rows(smem + kOffGB, smem + kOffGB + kGroupPlaneBytes, row, 16 * (c - 2), z);
            }
        }
        tc05::fence_proxy_async_smem();
        tc05::fence_before_thread_sync();
        __syncthreads();
        if (tid == 0) {
            tc05::fence_after_thread_sync();
            issue_weight_grad(tmem_base + cG2 + 32, ga_addr, kGAPlaneBytes, gb_addr, kGroupPlaneBytes, 32, !first_tile);
            tc05::mma_commit(bar);
        }
        tc05::mbar_wait(bar, phase & 1); ++phase;
        tc05::fence_after_thread_sync();

        TC_MARK(7);
        // ------------------------------------------------------------ dZ1 = dH1 * GELU'(Z1) -> rows (A operand of G1)
#pragma unroll 1
        for (int c = 2 * hf; c < 2 * hf + 2; ++c) {
            float dh[16], z[16];
            tc05::tmem_ld_32x32b_x16(tl + cZ2 + 16 * c, dh);
            tc05::tmem_ld_32x32b_x16(tl + cZ1 + 16 * c, z);
            tc05::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float2 g2, dg2;
                gelu_and_grad2(make_float2(z[j], z[j + 1]), g2, dg2);
                dh[j] *= dg2.x; dh[j + 1] *= dg2.y;
            }
            store_hi_lo_rows(smem + kOffGA, smem + kOffGA + kGAPlaneBytes, row, 16 * c, dh);
        }
        if (hf == 0) {   // [X | 1 | 0] rows: the B operand of G1 (the group buffer is free: both passes of G2 are complete)
            float xrow[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) xrow[k] = (k < SC && k < S) ? x[k < SC ? k : 0] : ((k == S && valid) ? 1.0f : 0.0f);
            if (N1 == 8) {
                float x8[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x8[k] = xrow[k];
                store_hi_lo_rows8(smem + kOffGB, smem + kOffGB + kGroupPlaneBytes, row, 0, x8);
            } else {
                store_hi_lo_rows(smem + kOffGB, smem + kOffGB + kGroupPlaneBytes, row, 0, xrow);
            }
        }
        tc05::fence_proxy_async_smem();
        tc05::fence_before_thread_sync();
        __syncthreads();
        if (tid == 0) {
            tc05::fence_after_thread_sync();
            issue_weight_grad(tmem_base + cG1, ga_addr, kGAPlaneBytes, gb_addr, kGroupPlaneBytes, N1, !first_tile);
            tc05::mma_commit(bar);
        }
        tc05::mbar_wait(bar, phase & 1); ++phase;
        tc05::fence_after_thread_sync();
        }   // tiles of this CTA

        TC_MARK(8);
        // ------------------------------------------------------------ gradients -> flat buffer (rows j = 16 quarter + lane, lane < 16;
        // the two warps of a lane quarter take 32 columns each)
        const bool atomic = !persistent;   // several CTAs (tiles / the sharded path) add into a zeroed buffer
        {
            const int j = 16 * quarter + (lane & 15);
            const bool row_owner = lane < 16;
#pragma unroll 1
            for (int c = 4 * hf; c < 4 * hf + 4; ++c) {
                float v[8];
                tc05::tmem_ld_32x32b_x8(tl + cG2 + 8 * c, v);
                tc05::tmem_ld_wait();
                if (row_owner) {
                    float* dst = g + oW1 + j * kHid + 8 * c;
                    if (atomic) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) atomicAdd(dst + i, v[i]);
                    } else {
                        // 16 lanes store 16 rows of 64 floats: the same column would be a 16-way bank conflict, so lane l
                        // stores its 8 values rotated by l (a three-stage barrel shifter of selects, no indexed registers)
                        const int r = lane & 7;
                        float t[8], w8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) t[i] = (r & 1) ? v[(i + 1) & 7] : v[i];
#pragma unroll
                        for (int i = 0; i < 8; ++i) w8[i] = (r & 2) ? t[(i + 2) & 7] : t[i];
#pragma unroll
                        for (int i = 0; i < 8; ++i) t[i] = (r & 4) ? w8[(i + 4) & 7] : w8[i];
#pragma unroll
                        for (int i = 0; i < 8; ++i) dst[(i + r) & 7] = t[i];   // t[i] = v[(i + r) & 7]
                    }
                }
            }
#pragma unroll 1
            for (int c = 0; c < (hf == 0 ? N1 / 8 : 0); ++c) {
                float v[8];
                tc05::tmem_ld_32x32b_x8(tl + cG1 + 8 * c, v);
                tc05::tmem_ld_wait();
                if (row_owner) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int k = 8 * c + i;
                        if (k < S) { if (atomic) atomicAdd(g + j * S + k, v[i]); else g[j * S + k] = v[i]; }
                        else if (k == S) { if (atomic) atomicAdd(g + oB0 + j, v[i]); else g[oB0 + j] = v[i]; }
                    }
                }
            }
        }
        __syncthreads();   // the shared-memory accumulators of the head are complete
        for (int i = tid; i < OUT * kHid; i += kNT) { if (atomic) atomicAdd(g + oW2 + i, small[kSmGW3 + i]); else g[oW2 + i] = small[kSmGW3 + i]; }
        if (tid < kHid) { if (atomic) atomicAdd(g + oB1 + tid, small[kSmGB2 + tid]); else g[oB1 + tid] = small[kSmGB2 + tid]; }
        if (tid < OUT) {
            if (atomic) atomicAdd(g + oB2 + tid, small[kSmGB3 + tid]); else g[oB2 + tid] = small[kSmGB3 + tid];
            if (gaussian) { if (atomic) atomicAdd(g + oStd + tid, small[kSmGStd + tid]); else g[oStd + tid] = small[kSmGStd + tid]; }
        }

        // ------------------------------------------------------------ loss sums
        {
            float c = tot_c, s = tot_s, e = tot_e;
            block_sum3<kNT>(c, s, e, red);
            if (tid == 0) {
                if (sharded) { g[numel] = c * inv_bsz; g[numel + 1] = s * inv_bsz; g[numel + 2] = e * inv_bsz; }
                else if (persistent) { acc_c += (double)(c * inv_bsz); acc_s += (double)(s * inv_bsz); acc_e += (double)(e * inv_bsz); }
                else if (ni == 1) atomicAdd(A.loss_sums + 0, (double)(c * inv_bsz));
                else { atomicAdd(A.loss_sums + 1, (double)(s * inv_bsz)); atomicAdd(A.loss_sums + 2, (double)(e * inv_bsz)); }
            }
        }

        TC_MARK(9);
        // ------------------------------------------------------------ clip_grad_norm_ + Adam.step
        if (sharded) {
            // ---- the gradient all-reduce, in the kernel: flags over NVLink, then ordered sums of the peers' buffers
            px_round(A.px, 1 + ni, A.px.epoch + 1 + (uint32_t)u, A.hdr);
            const int world = A.px.world;
            px_reduce(A.px, px_off, numel, g_smem);
            if (tid == 0) {
                float c = 0.f, s = 0.f, e = 0.f;
                for (int r = 0; r < world; ++r) {
                    c += ld_relaxed_sys(A.px.data[r] + px_off + numel);
                    s += ld_relaxed_sys(A.px.data[r] + px_off + numel + 1);
                    e += ld_relaxed_sys(A.px.data[r] + px_off + numel + 2);
                }
                acc_c += (double)c; acc_s += (double)s; acc_e += (double)e;
            }
            __syncthreads();
        }
        if (persistent) {
            if (tid == 0) {
                if (u < kAdamTab) s_adam = s_adam_tab[u];
                else {   // long schedules: computed on the spot
                    const double step = (double)(A.opt[ni].step + u + 1);
                    s_adam.step_size = (float)((double)A.opt[ni].lr / (1.0 - pow((double)A.opt[ni].beta1, step)));
                    s_adam.bc2_sqrt = (float)sqrt(1.0 - pow((double)A.opt[ni].beta2, step));
                }
            }
            __syncthreads();
            opt_apply_resident(net, A.opt[ni], s_adam, g_smem, numel, A.hp.clip_grad_norm, red, small, w2_hi, w2_lo, wb_hi, wb_lo, tl, tid,
                               S, OUT, gaussian, u == U - 1);
            __syncthreads();
            TC_MARK(10);
        } else if (A.fused_apply) {
            __threadfence();
            __syncthreads();
            if (tid == 0) s_last = (atomicAdd(&A.hdr->ticket[ni], 1u) == gridDim.x - 1) ? 1 : 0;
            __syncthreads();
            if (s_last) {   // last CTA of this net: the whole gradient is in the buffer
                __threadfence();
                apply_from_global(net, A.opt[ni], A.adam[ni], g_local, numel, A.hp.clip_grad_norm, red);
                __syncthreads();
                for (int i = tid; i < numel; i += kNT) g_local[i] = 0.0f;
                if (tid == 0) A.hdr->ticket[ni] = 0u;
            }
        }
    }
    if (persistent && tid == 0) {
        if ((sharded || gathered) && atomicAdd(&A.hdr->pad[0], 0u) != 0u) acc_c = acc_s = acc_e = (double)nanf("");   // a peer never answered
        const double inv = 1.0 / (double)A.update_times;
        if (ni == 1) A.out_scalars[0] = (float)(acc_c * inv);
        else { A.out_scalars[1] = (float)(acc_s * inv); A.out_scalars[2] = (float)(acc_e * inv); }
    }
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<512>(tmem_base);
}

}  // namespace

bool b200rl_update_tc_eligible(const b200rl_net* actor, const b200rl_net* critic, const b200rl_ppo_hyper* hp) {
    for (const b200rl_net* n : {actor, critic}) {
        if (n->num_linear != 3 || n->dims[1] != kHid || n->dims[2] != kHid || n->activation != B200RL_ACT_GELU) return false;
        if (n->dims[0] < 1 || n->dims[0] > kMaxS || n->dims[3] < 1 || n->dims[3] > kMaxOut) return false;
    }
    if (critic->dims[3] != 1 || actor->dims[0] != critic->dims[0]) return false;
    if (hp->flags & B200RL_PPO_CRITIC_MASK_MEAN) return false;   // needs the minibatch-wide mean of unmask (tutorial variant: ReLU nets anyway)
    return true;
}

extern "C" int32_t b200rl_update_tc_supported(const b200rl_net* actor, const b200rl_net* critic, const b200rl_ppo_hyper* hyper) {
    return (actor && critic && hyper && b200rl_update_tc_eligible(actor, critic, hyper)) ? 1 : 0;
}
extern "C" int64_t b200rl_peer_exchange_floats(const b200rl_net* actor, const b200rl_net* critic, int32_t local_batch, int32_t update_times) {
    // gradient all-reduce: two alternating (gradient + 3 loss sums) segments per net;  record gather: two cycle parities x two
    // nets x update_times x local_batch records of roundup4(S + A) + 4 floats
    const int seg = ((int)b200rl_net_numel(actor) + 4 + 3 & ~3) + ((int)b200rl_net_numel(critic) + 4 + 3 & ~3);
    const int rec = ((actor->dims[0] + actor->dims[actor->num_linear] + 3) & ~3) + 4;
    const int64_t grads = 2 * (int64_t)seg, records = 4 * (int64_t)update_times * local_batch * rec;
    return kPxStatFloats + (grads > records ? grads : records);
}

// grid = (tiles, 2 nets).  A.update_times > 0: persistent (tiles must be 1); otherwise one minibatch.
int b200rl_launch_update_tc(const UpdateArgs& A_in, int tiles, cudaStream_t stream) {
    UpdateArgs A = A_in;
    const char* prof = getenv("B200RL_PROFILE");   // debug: clock64 phase marks (tools/profile_update_phases.py)
    A.profile = (prof && prof[0] == '1') ? 1 : 0;
    const int S = A.net[0].dims[0], OUT = A.net[0].dims[3];
    auto launch = [&](auto kern) -> int {
        B200RL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        // one CTA per SM and net at most (196 KB of shared memory each): a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
        const int ctas = tiles < kMaxCtasPerNet ? tiles : kMaxCtasPerNet;
        kern<<<dim3((unsigned)ctas, 2), kNT, kSmemBytes, stream>>>(A);
        return 0;
    };
    int rc;
    if (OUT <= 1 && S <= 4) rc = launch(ppo_update_tc_kernel<1, 4>);          // Pendulum (BASELINE configs[1])
    else if (OUT <= 2 && S <= 8) rc = launch(ppo_update_tc_kernel<2, 8>);     // LunarLanderContinuous / CartPole dims
    else if (OUT <= 4 && S <= kMaxS) rc = launch(ppo_update_tc_kernel<4, kMaxS>);   // Hopper dims
    else rc = launch(ppo_update_tc_kernel<kMaxOut, kMaxS>);
    if (rc) return rc;
    B200RL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
