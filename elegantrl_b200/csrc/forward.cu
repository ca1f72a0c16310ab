// Generic MLP forward + exploration step (any depth / widths), CUDA cores.
//   b200rl_mlp_forward  : CriticPPO.forward / ActorPPO.forward   (reference AgentPPO.py:435-438, 363-366)
//   b200rl_policy_step  : ActorPPO.get_action + convert_action_for_env (+ critic value)
//                         (reference AgentPPO.py:368-376, 388-390; loop body of _explore_vec_env :113-119)
#include "mlp_tile.cuh"
#include "policy_epilogue.cuh"

// forward_tc.cu: the tcgen05 kernel for S -> 64 -> 64 -> OUT GELU nets
bool b200rl_forward_tc_eligible(const b200rl_net* net);
int b200rl_launch_forward_tc(int policy, const b200rl_net* net, const float* x, int64_t rows, float* out, int out_tanh,
                             const PolicyOut& po, cudaStream_t stream);

namespace {

constexpr int kFwdThreads = 256;

template <int TB, int POLICY>
__global__ void __launch_bounds__(kFwdThreads)
mlp_forward_kernel(const __grid_constant__ b200rl_net net, const float* __restrict__ x, int64_t rows, float* out,
                   int out_tanh, int buf_floats, const __grid_constant__ PolicyOut po) {
    extern __shared__ float4 smem4[];
    float* cur = reinterpret_cast<float*>(smem4);
    float* nxt = cur + buf_floats;
    using T = SmemTile<TB>;
    const int64_t row0 = (int64_t)blockIdx.x * TB;
    const int L = net.num_linear;

    load_state_tile<TB, kFwdThreads>(net, x, rows, row0, cur);
    __syncthreads();
    for (int l = 0; l < L; ++l) {
        linear_forward<TB, kFwdThreads>(net.weight[l], net.bias[l], net.dims[l], net.dims[l + 1], cur, nxt, nullptr,
                                        net.activation, l < L - 1);
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
    }
    const int J = net.dims[L];
    const uint64_t rng_step = (POLICY != kPlain && po.step_base) ? po.step + *po.step_base : po.step;
    if (POLICY == kPlain) {
        for (int idx = threadIdx.x; idx < TB * J; idx += kFwdThreads) {
            int b = idx / J, j = idx - b * J;
            int64_t row = row0 + b;
            if (row < rows) {
                float v = cur[T::elem(j, b)];
                out[row * J + j] = out_tanh ? tanhf(v) : v;
            }
        }
    } else {
        const int b = threadIdx.x;
        const int64_t row = row0 + b;
        if (b < TB && row < rows) {
            auto get = [&](int a) { return cur[T::elem(a, b)]; };
            if (POLICY == kCategorical) categorical_epilogue(po, rng_step, row, J, get);
            else gaussian_epilogue(net, po, rng_step, row, J, get);
        }
    }
}

template <int POLICY>
int launch_forward(const b200rl_net* net, const float* x, int64_t rows, float* out, int out_tanh, const PolicyOut& po,
                   cudaStream_t stream) {
    if (rows <= 0) return 0;
    if (b200rl_forward_tc_eligible(net)) return b200rl_launch_forward_tc(POLICY, net, x, rows, out, out_tanh, po, stream);
    const int maxdim = b200rl_net_maxdim(net);
    // 64-sample tiles when two ping-pong buffers fit comfortably, else 32
    if (maxdim <= 256) {
        constexpr int TB = 64;
        int buf_floats = maxdim * TB;
        size_t smem = 2 * (size_t)buf_floats * sizeof(float);
        auto kern = mlp_forward_kernel<TB, POLICY>;
        if (smem > 48 * 1024) B200RL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t grid = (rows + TB - 1) / TB;
        kern<<<(unsigned)grid, kFwdThreads, smem, stream>>>(*net, x, rows, out, out_tanh, buf_floats, po);
        B200RL_COUNT_LAUNCH(1);
    } else {
        constexpr int TB = 32;
        int buf_floats = maxdim * TB;
        size_t smem = 2 * (size_t)buf_floats * sizeof(float);
        B200RL_REQUIRE(smem <= 227 * 1024, "mlp_forward: widest layer %d too large for shared memory", maxdim);
        auto kern = mlp_forward_kernel<TB, POLICY>;
        if (smem > 48 * 1024) B200RL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t grid = (rows + TB - 1) / TB;
        kern<<<(unsigned)grid, kFwdThreads, smem, stream>>>(*net, x, rows, out, out_tanh, buf_floats, po);
        B200RL_COUNT_LAUNCH(1);
    }
    B200RL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

thread_local const uint64_t* t_step_base = nullptr;

extern "C" {

void b200rl_set_policy_step_base(const uint64_t* step_base) { t_step_base = step_base; }

int b200rl_mlp_forward(const b200rl_net* net, const float* x, int64_t rows, float* out, int32_t out_tanh, void* stream) {
    if (int rc = b200rl_validate_net(net, "mlp_forward", false)) return rc;
    B200RL_REQUIRE(x && out, "mlp_forward: x/out is NULL");
    PolicyOut po{};
    return launch_forward<kPlain>(net, x, rows, out, out_tanh, po, (cudaStream_t)stream);
}

int b200rl_policy_step(const b200rl_net* actor, const b200rl_net* critic, const float* state, int64_t rows,
                       const float* eps, uint64_t seed, uint64_t step, int64_t env_offset, float* action,
                       float* logprob, float* env_action, float* value, void* stream) {
    if (int rc = b200rl_validate_net(actor, "policy_step.actor", true)) return rc;
    B200RL_REQUIRE(state && action && logprob && env_action, "policy_step: NULL buffer");
    PolicyOut po{eps, seed, step, t_step_base, env_offset, action, logprob, env_action, nullptr};
    if (int rc = launch_forward<kGaussian>(actor, state, rows, nullptr, 0, po, (cudaStream_t)stream)) return rc;
    if (critic && value) {
        if (int rc = b200rl_validate_net(critic, "policy_step.critic", false)) return rc;
        B200RL_REQUIRE(critic->dims[critic->num_linear] == 1, "policy_step: critic output dim must be 1");
        PolicyOut none{};
        return launch_forward<kPlain>(critic, state, rows, value, 0, none, (cudaStream_t)stream);
    }
    return 0;
}

int b200rl_policy_step_discrete(const b200rl_net* actor, const b200rl_net* critic, const float* state, int64_t rows,
                                const float* expo, uint64_t seed, uint64_t step, int64_t env_offset, int32_t* action,
                                float* logprob, float* value, void* stream) {
    if (int rc = b200rl_validate_net(actor, "policy_step_discrete.actor", false)) return rc;
    B200RL_REQUIRE(actor->action_std_log == nullptr, "policy_step_discrete: a categorical actor has no action_std_log");
    B200RL_REQUIRE(state && action && logprob, "policy_step_discrete: NULL buffer");
    PolicyOut po{expo, seed, step, t_step_base, env_offset, nullptr, logprob, nullptr, action};
    if (int rc = launch_forward<kCategorical>(actor, state, rows, nullptr, 0, po, (cudaStream_t)stream)) return rc;
    if (critic && value) {
        if (int rc = b200rl_validate_net(critic, "policy_step_discrete.critic", false)) return rc;
        B200RL_REQUIRE(critic->dims[critic->num_linear] == 1, "policy_step_discrete: critic output dim must be 1");
        PolicyOut none{};
        return launch_forward<kPlain>(critic, state, rows, value, 0, none, (cudaStream_t)stream);
    }
    return 0;
}

}  // extern "C"
