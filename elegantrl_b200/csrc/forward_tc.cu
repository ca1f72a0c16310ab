// tcgen05 forward / exploration step for the reference's default net shape  S -> 64 -> 64 -> OUT  (GELU; S <= 16, OUT <= 8):
//   b200rl_mlp_forward   CriticPPO.forward / ActorPPO.forward       (reference AgentPPO.py:435-438, 363-366)
//   b200rl_policy_step   ActorPPO.get_action + convert_action_for_env (reference AgentPPO.py:368-376, 388-390)
//   b200rl_policy_step_discrete   ActorDiscretePPO.get_action        (reference AgentPPO.py:407-413)
// -- the per-step engine path of external vec envs (loop body of _explore_vec_env :113-119), the values pass over a
// [H * N, S] buffer that did not come from a fused rollout (:141-143) and the evaluator's deterministic policy.  Any other
// shape keeps the CUDA-core kernel of forward.cu (same entry points; B200RL_FORWARD=ffma forces it).
//
// Arithmetic = the forward half of update_tc.cu (3xTF32: hi / lo planes of both operands, fp32 accumulate in tensor memory):
//   L1  Z1 = x~ B1^T    x~ = [x_hi, 1, x_lo, 0..] K-major in shared memory (bias folded), K1 = roundup8(2 S + 1)        SS
//   L2  Z2 = H1 W2^T    H1 = GELU(Z1) as hi / lo planes in TENSOR MEMORY; W2 exactly as nn.Linear stores it (K-major)  TS
//   head (64 -> OUT), b2, GELU and the policy epilogue on CUDA cores.
// Mapping: persistent CTAs (2 per SM: 256 TMEM columns and ~75 KB of shared memory each) walk tiles of 128 rows; the operand
// images of the net are staged ONCE per CTA.  256 threads = two per row (warps w and w + 4 share a TMEM lane quarter and
// split the 64 columns), thread 0 issues the UMMAs; the two CTAs of an SM run out of phase, so one tile's tensor-core round
// trips are covered by the other's GELUs.  Z2 reuses Z1's columns (Z1 is dead once H1 is written).
#include <stdlib.h>

#include "policy_epilogue.cuh"
#include "tc_train.cuh"

namespace {

using namespace tctrain;

constexpr int kT = 128, kNT = 256;
constexpr int kMaxS = 16, kMaxOut = 8, kK1Max = 40;   // roundup8(2 * 16 + 1)
constexpr int cZ = 0, cPhi = 64, cPlo = 128;          // TMEM columns: Z1 / Z2, H1 hi plane, H1 lo plane (256 allocated)

constexpr int kOffW2 = 0;                                      // W2 hi / lo K-major images
constexpr int kA1Bytes = kT * kK1Max * 4;
constexpr int kOffA1 = kOffW2 + 2 * kWPlaneBytes;              // x~ rows, K-major
constexpr int kB1PlaneBytes = kHid * kK1Max * 4;
constexpr int kOffB1 = kOffA1 + kA1Bytes;                      // layer-1 B planes [W_hi, b_hi, W_hi, 0] / [W_lo, b_lo, 0, 0]
constexpr int kSmW3 = 0, kSmB2 = 512, kSmB3 = 576, kSmAvg = 584, kSmSd = 600, kSmallFloats = 616;
constexpr int kOffSmall = kOffB1 + 2 * kB1PlaneBytes;
constexpr int kOffPart = kOffSmall + kSmallFloats * 4;         // head partial sums of the upper column half: [128][OUTC]
constexpr int kOffBar = kOffPart + kT * kMaxOut * 4;
constexpr int kSmemBytes = kOffBar + 16;
static_assert(2 * kSmemBytes <= 226 * 1024, "two CTAs per SM");
static_assert(kOffA1 % 128 == 0 && kOffB1 % 128 == 0, "operand alignment");

template <int OUTC, int POLICY>
__global__ void __launch_bounds__(kNT, 2)
mlp64_forward_tc_kernel(const __grid_constant__ b200rl_net net, const float* __restrict__ x, int64_t rows, float* out, int out_tanh,
                        const __grid_constant__ PolicyOut po) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    float* small = reinterpret_cast<float*>(smem + kOffSmall);
    float* part = reinterpret_cast<float*>(smem + kOffPart);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (kT - 1), hf = tid >> 7, quarter = warp & 3;
    const int S = net.dims[0], OUT = net.dims[3];
    const int K1 = (2 * S + 1 + 7) & ~7;

    // ---- one-time setup: TMEM, barrier, operand images of the net
    if (warp == 0) tc05::tmem_alloc<256>(&tmem_slot);
    if (tid == 32) { tc05::mbar_init(bar, 1); tc05::mbar_fence_init(); }
    for (int i = tid; i < kA1Bytes / 4; i += kNT) reinterpret_cast<float*>(smem + kOffA1)[i] = 0.0f;
    stage_w_planes(net.weight[1], smem + kOffW2, smem + kOffW2 + kWPlaneBytes, tid, kNT);
    for (int i = tid; i < kHid * K1; i += kNT) {
        const int n = i / K1, k = i - n * K1;
        float full = 0.0f;
        if (k < S) full = net.weight[0][n * S + k];
        else if (k == S) full = net.bias[0][n];
        else if (k <= 2 * S) full = net.weight[0][n * S + (k - S - 1)];
        const float hi = tc05::tf32_hi(full);
        const uint32_t off = tc05::operand_offset(n, k, K1);
        *reinterpret_cast<float*>(smem + kOffB1 + off) = hi;
        *reinterpret_cast<float*>(smem + kOffB1 + kB1PlaneBytes + off) = k <= S ? full - hi : 0.0f;
    }
    for (int i = tid; i < OUT * kHid; i += kNT) small[kSmW3 + i] = net.weight[2][i];
    if (tid < kHid) small[kSmB2 + tid] = net.bias[1][tid];
    if (tid < OUT) small[kSmB3 + tid] = net.bias[2][tid];
    if (tid < S) {
        small[kSmAvg + tid] = net.state_avg ? net.state_avg[tid] : 0.0f;
        small[kSmSd + tid] = net.state_std ? net.state_std[tid] + 1e-4f : 1.0f;
    }
    tc05::fence_proxy_async_smem();
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t tl = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t w2_hi = tc05::smem_u32(smem + kOffW2), w2_lo = w2_hi + kWPlaneBytes;
    const bool norm = net.state_avg != nullptr;
    const uint64_t rng_step = (POLICY != kPlain && po.step_base) ? po.step + *po.step_base : po.step;
    uint32_t phase = 0;

    const int64_t tiles = (rows + kT - 1) / kT;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t r = tile * kT + row;
        const bool valid = r < rows;
        // ---- x~ row (lower-half threads): [x_hi, 1, x_lo, 0..]; rows past the end keep their previous (finite) content
        if (hf == 0 && valid) {
            const float* xr = x + r * S;
            for (int k = 0; k < S; ++k) {
                float v = xr[k];
                if (norm) v = (v - small[kSmAvg + k]) / small[kSmSd + k];
                const float hi = tc05::tf32_hi(v);
                *reinterpret_cast<float*>(smem + kOffA1 + tc05::operand_offset(row, k, K1)) = hi;
                *reinterpret_cast<float*>(smem + kOffA1 + tc05::operand_offset(row, S + 1 + k, K1)) = v - hi;
            }
            *reinterpret_cast<float*>(smem + kOffA1 + tc05::operand_offset(row, S, K1)) = 1.0f;
        }
        tc05::fence_proxy_async_smem();
        tc05::fence_before_thread_sync();   // ... and this thread's Z2 reads of the previous tile precede the UMMAs below
        __syncthreads();
        if (tid == 0) {   // layer 1
            tc05::fence_after_thread_sync();
            const uint32_t idesc = tc05::make_idesc_tf32(kT, kHid);
            const uint32_t sbo = (uint32_t)(K1 / 4) * 128;
            const uint32_t a1 = tc05::smem_u32(smem + kOffA1), b1 = tc05::smem_u32(smem + kOffB1);
            for (int p = 0; p < 2; ++p)
                for (int ks = 0; ks < K1 / 8; ++ks)
                    tc05::mma_tf32(tmem_base + cZ, tc05::make_smem_desc_ex(a1 + ks * 256, 128, sbo),
                                   tc05::make_smem_desc_ex(b1 + p * kB1PlaneBytes + ks * 256, 128, sbo), idesc, p > 0 || ks > 0);
            tc05::mma_commit(bar);
        }
        tc05::mbar_wait(bar, phase & 1); ++phase;
        tc05::fence_after_thread_sync();
        // ---- H1 = GELU(Z1) -> hi / lo planes in tensor memory (this thread's 32 columns)
#pragma unroll 1
        for (int c = 2 * hf; c < 2 * hf + 2; ++c) {
            float z[16];
            tc05::tmem_ld_32x32b_x16(tl + cZ + 16 * c, z);
            tc05::tmem_ld_wait();
            gelu_only16(z);
            store_hi_lo_tmem(tl + cPhi + 16 * c, tl + cPlo + 16 * c, z);
        }
        tc05::tmem_st_wait();
        tc05::fence_before_thread_sync();
        __syncthreads();
        if (tid == 0) {   // layer 2 (Z2 overwrites Z1: every thread has read its Z1 columns)
            tc05::fence_after_thread_sync();
            issue_linear_ts(tmem_base + cZ, tmem_base + cPhi, tmem_base + cPlo, w2_hi, w2_lo, false);
            tc05::mma_commit(bar);
        }
        tc05::mbar_wait(bar, phase & 1); ++phase;
        tc05::fence_after_thread_sync();
        // ---- head: b2, GELU, 64 -> OUT (fixed order: lower half + upper half + bias, as update_tc.cu)
        float o[OUTC];
#pragma unroll
        for (int a = 0; a < OUTC; ++a) o[a] = 0.0f;
#pragma unroll 1
        for (int c = 2 * hf; c < 2 * hf + 2; ++c) {
            float z[16];
            tc05::tmem_ld_32x32b_x16(tl + cZ + 16 * c, z);
            tc05::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] += small[kSmB2 + 16 * c + j];
            gelu_only16(z);
#pragma unroll
            for (int a = 0; a < OUTC; ++a) {
                if (a < OUT) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[a] = fmaf(z[j], small[kSmW3 + a * kHid + 16 * c + j], o[a]);
                }
            }
        }
        if (hf == 1) {
#pragma unroll
            for (int a = 0; a < OUTC; ++a) part[row * OUTC + a] = o[a];
        }
        __syncthreads();
        if (hf == 0 && valid) {
#pragma unroll
            for (int a = 0; a < OUTC; ++a) o[a] = a < OUT ? (o[a] + part[row * OUTC + a]) + small[kSmB3 + a] : 0.0f;
            if (POLICY == kPlain) {
#pragma unroll
                for (int a = 0; a < OUTC; ++a) if (a < OUT) out[r * OUT + a] = out_tanh ? tanhf(o[a]) : o[a];
            } else {
                auto get = [&](int a) {
                    float v = o[0];
#pragma unroll
                    for (int q = 1; q < OUTC; ++q) v = a == q ? o[q] : v;
                    return v;
                };
                if (POLICY == kCategorical) categorical_epilogue(po, rng_step, r, OUT, get);
                else gaussian_epilogue(net, po, rng_step, r, OUT, get);
            }
        }
        // `part` is rewritten only after the two barriers of the next tile; x~ rows are free since layer 1 has completed
    }

    tc05::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<256>(tmem_base);
}

template <int OUTC, int POLICY>
int launch(const b200rl_net* net, const float* x, int64_t rows, float* out, int out_tanh, const PolicyOut& po, cudaStream_t stream) {
    auto kern = mlp64_forward_tc_kernel<OUTC, POLICY>;
    static thread_local int configured_device = -1;
    int dev = 0;
    B200RL_CHECK_CUDA(cudaGetDevice(&dev));
    if (configured_device != dev) {
        B200RL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        configured_device = dev;
    }
    const int64_t tiles = (rows + kT - 1) / kT;
    const unsigned grid = (unsigned)(tiles < 2 * 148 ? tiles : 2 * 148);
    kern<<<grid, kNT, kSmemBytes, stream>>>(*net, x, rows, out, out_tanh, po);
    B200RL_COUNT_LAUNCH(1);
    B200RL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

bool b200rl_forward_tc_eligible(const b200rl_net* n) {
    const char* f = getenv("B200RL_FORWARD");
    if (f && f[0] == 'f') return false;   // "ffma": the CUDA-core kernel for every shape
    return n->num_linear == 3 && n->dims[1] == kHid && n->dims[2] == kHid && n->activation == B200RL_ACT_GELU &&
           n->dims[0] >= 1 && n->dims[0] <= kMaxS && n->dims[3] >= 1 && n->dims[3] <= kMaxOut;
}

// policy: kPlain / kGaussian / kCategorical
int b200rl_launch_forward_tc(int policy, const b200rl_net* net, const float* x, int64_t rows, float* out, int out_tanh,
                             const PolicyOut& po, cudaStream_t stream) {
    if (rows <= 0) return 0;
    const int OUT = net->dims[3];
    if (policy == kPlain) {
        if (OUT == 1) return launch<1, kPlain>(net, x, rows, out, out_tanh, po, stream);
        if (OUT <= 4) return launch<4, kPlain>(net, x, rows, out, out_tanh, po, stream);
        return launch<8, kPlain>(net, x, rows, out, out_tanh, po, stream);
    }
    if (policy == kGaussian) {
        if (OUT == 1) return launch<1, kGaussian>(net, x, rows, out, out_tanh, po, stream);
        if (OUT <= 4) return launch<4, kGaussian>(net, x, rows, out, out_tanh, po, stream);
        return launch<8, kGaussian>(net, x, rows, out, out_tanh, po, stream);
    }
    if (OUT <= 4) return launch<4, kCategorical>(net, x, rows, out, out_tanh, po, stream);
    return launch<8, kCategorical>(net, x, rows, out, out_tanh, po, stream);
}
