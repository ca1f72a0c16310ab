// Shared device helpers of libb200rl (sm_100a).  See include/b200rl.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/b200rl.h"

// ------------------------------------------------------------------------------------------ errors
void b200rl_set_error(const char* fmt, ...);

#define B200RL_CHECK_CUDA(expr)                                                                    \
    do {                                                                                           \
        cudaError_t err__ = (expr);                                                                \
        if (err__ != cudaSuccess) {                                                                \
            b200rl_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(err__)); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

#define B200RL_REQUIRE(cond, ...)                                                                  \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            b200rl_set_error(__VA_ARGS__);                                                         \
            return 2;                                                                              \
        }                                                                                          \
    } while (0)

// number of parameter tensors / scalars of a net, in the flat order W0, b0, W1, b1, ..., action_std_log
static inline int64_t b200rl_net_numel(const b200rl_net* net) {
    int64_t n = 0;
    for (int l = 0; l < net->num_linear; ++l) n += (int64_t)net->dims[l + 1] * net->dims[l] + net->dims[l + 1];
    if (net->action_std_log) n += net->dims[net->num_linear];
    return n;
}
static inline int b200rl_net_maxdim(const b200rl_net* net) {
    int m = 0;
    for (int l = 0; l <= net->num_linear; ++l) m = net->dims[l] > m ? net->dims[l] : m;
    return m;
}
int b200rl_validate_net(const b200rl_net* net, const char* name, bool is_actor);
extern long long g_b200rl_launches;  // kernels launched by this library in this process (b200rl_launch_count)
#define B200RL_COUNT_LAUNCH(n) (g_b200rl_launches += (n))

// workspace layout (bytes): [0, 256) header {double loss_sums[4]; unsigned ticket[2]; ...}; [256, ...) flat grads
#define B200RL_WS_HEADER_BYTES 256
struct WorkspaceHeader {
    double loss_sums[4];
    unsigned int ticket[2];  // per net: CTAs of the running minibatch that finished their gradient phase
    unsigned int pad[2];
};

// ------------------------------------------------------------------------------------------- math
#ifdef __CUDACC__
#define DEV __device__ __forceinline__

constexpr float kSqrtHalf = 0.70710678118654752440f;
constexpr float kInvSqrt2Pi = 0.39894228040143267794f;
constexpr float kLogSqrt2Pi = 0.91893853320467274178f;  // log(sqrt(2 pi))
constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;

// nn.GELU() exact form: x * 0.5 * (1 + erf(x / sqrt(2)))  (torch CPU kernel op order)
DEV float gelu_erf(float x) { return x * 0.5f * (1.0f + erff(x * kSqrtHalf)); }
// d/dx: Phi(x) + x * phi(x)
DEV float gelu_erf_grad(float x) {
    float cdf = 0.5f * (1.0f + erff(x * kSqrtHalf));
    float pdf = expf(-0.5f * x * x) * kInvSqrt2Pi;
    return cdf + x * pdf;
}
template <int ACT>
DEV float act_fn(float z) {
    if (ACT == B200RL_ACT_GELU) return gelu_erf(z);
    return fmaxf(z, 0.0f);
}
DEV float act_fn_rt(float z, int act) { return act == B200RL_ACT_GELU ? gelu_erf(z) : fmaxf(z, 0.0f); }
DEV float act_grad_rt(float z, int act) {
    return act == B200RL_ACT_GELU ? gelu_erf_grad(z) : (z > 0.0f ? 1.0f : 0.0f);
}

// ------------------------------------------------------------------------------------------ Philox
// Philox4x32-10 counter-based RNG (Salmon et al. 2011): stateless, keyed by (seed), counter = (env, chunk, step).
DEV uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
DEV float u32_to_unit_open(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0, 1)
DEV float u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }                // [0, 1)
// two N(0,1) from two uint32 (Box-Muller)
DEV float2 box_muller(uint32_t a, uint32_t b) {
    float r = sqrtf(-2.0f * logf(u32_to_unit_open(a)));
    float s, c;
    sincosf(kTwoPi * u32_to_unit(b), &s, &c);
    return make_float2(r * c, r * s);
}
constexpr uint32_t kStreamRollout = 0x0u;  // counter.w tags
constexpr uint32_t kStreamIds = 0x1D5u;

struct RolloutNoise {
    float2 normal;   // two N(0,1)
    float2 uniform;  // two U[0,1) (env reset)
};
DEV RolloutNoise rollout_noise(uint64_t seed, uint64_t env, uint64_t step, uint32_t chunk) {
    uint4 ctr = make_uint4((uint32_t)env, (uint32_t)(env >> 32) ^ (chunk << 8), (uint32_t)step,
                           (uint32_t)(step >> 32) ^ kStreamRollout);
    uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    RolloutNoise n;
    n.normal = box_muller(r.x, r.y);
    n.uniform = make_float2(u32_to_unit(r.z), u32_to_unit(r.w));
    return n;
}
// raw Philox words of the rollout stream (categorical sampling draws its Exp(1) noise from them, 4 actions per call;
// chunk numbers >= 0x800000 keep it apart from the Gaussian path's chunks)
DEV uint4 rollout_bits(uint64_t seed, uint64_t env, uint64_t step, uint32_t chunk) {
    uint4 ctr = make_uint4((uint32_t)env, (uint32_t)(env >> 32) ^ ((chunk | 0x800000u) << 8), (uint32_t)step,
                           (uint32_t)(step >> 32) ^ kStreamRollout);
    return philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// uniform integer in [0, range) for minibatch sampling (replaces th.randint of reference AgentPPO.py:178)
DEV int64_t sample_index(uint64_t seed, uint64_t draw, uint32_t slot, uint64_t range) {
    uint4 ctr = make_uint4(slot, 0u, (uint32_t)draw, (uint32_t)(draw >> 32) ^ (kStreamIds << 16));
    uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    uint64_t x = ((uint64_t)r.x << 32) | r.y;
    return (int64_t)__umul64hi(x, range);
}

// ----------------------------------------------------------------------------------- reductions
DEV float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif  // __CUDACC__
