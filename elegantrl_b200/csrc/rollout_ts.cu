// Fused Pendulum rollout, second tcgen05 implementation ("TS": the layer-2 A operand lives in tensor memory) for the
// 3 -> 64 -> 64 -> 1 GELU actor + critic (BASELINE config 2).  Same contract and outputs as rollout.cu / rollout_tc.cu
// (reference AgentPPO._explore_vec_env, elegantrl/agents/AgentPPO.py:87-129, + values pass :141-143 + V(last_state)
// :219-220); parity vs the oracle rtol 1e-4.
//
// Why a second design (profiles/r02_*): rollout_tc.cu keeps the GELU outputs of layer 1 in a shared-memory ring (32
// STS.128 per thread and evaluation, 144 KB of UMMA operand reads per tile and evaluation) and evaluates layer 1 on
// CUDA cores out of broadcast LDS.128 -- its shared-memory data pipe is the saturated unit.  Here (ts_mlp.cuh):
//   * layer 1 runs on the tensor core too (x~ = [x_hi, 1, x_lo, 0], bias folded; 2 tiny UMMAs);
//   * its fp32 output is converted IN PLACE in tensor memory to the fp16 {hi, lo} A operand of layer 2 (tcgen05.ld ->
//     GELU -> split -> tcgen05.st to the same columns): no shared-memory traffic for activations, no proxy fence;
//   * layer 2's bias enters through a constant-A UMMA, so the epilogue is GELU + dot only.
// Per env-step ~2 500 issued lane-instructions instead of ~3 560, ~30 KB of shared-memory operand reads per tile and
// evaluation instead of ~300 KB.
//
// Mapping.  One persistent CTA per SM owns 448 envs = 3.5 tiles of 128 rows for all H steps (65 536 envs -> 147 CTAs).
// 14 worker warps (thread = env; warp_id % 4 = TMEM lane quarter) evaluate actor AND critic for their env, one after
// the other, sharing the tile's two 64-column TMEM regions X (layer-1 output / layer-2 A operand) and D (layer-2
// accumulator); 4 issuer warps (one per tile) wait on the workers' mbarriers and issue the UMMAs.  Per step and tile:
//   workers: x~(t) in shared memory -> arrive bar_x
//   issuer : layer 1 (actor) -> commit d1;   per 16 hidden units that the 4 worker warps have converted (bar_a2[c]):
//            4 UMMAs into D;  after the last: commit d2;  then immediately layer 1 of the critic (same x~ row image of the
//            critic's normalisation) into X -- the tensor pipe executes in issue order, so X is free by then
//   workers: wait d1 -> 4 chunks in place -> Philox noise (covers the tail of the UMMAs) -> wait d2 -> head -> action;
//            critic chunks -> env step, trajectory stores, x~(t+1), arrive bar_x (the next step's layer 1 overlaps the
//            critic head) -> wait d2 -> head -> value.
// Every barrier completes exactly once per evaluation, so all parities are (evaluation counter & 1).
#include <stdlib.h>

#include "rollout_params.cuh"
#include "ts_mlp.cuh"

namespace {

using namespace tsmlp;

constexpr int kTiles = 4, kRowsPerCta = 448;
constexpr int kWorkerWarps = 14, kIssuerWarps = kTiles;
// Variant bits (B200RL_TS_VARIANT, measured in profiles/r02_*_ts_variants.log):
//   1  software-pipelined tensor-memory loads (the load of chunk c + 1 is in flight while chunk c is evaluated)
//   2  issuer warps on the two schedulers that host only 3 worker warps (warp ids 14, 15, 18, 19; 16 and 17 stay idle)
//   4  one arrival per 32 hidden units instead of per 16 (half the arrivals / issuer wake-ups; 8 UMMAs per wake-up)
//   8  the state row of step t is stored while the actor's last UMMAs run (it is known at the start of the step)
constexpr int kDefaultVariant = 2;   // profiles/r02_v3_rollout_ts_variants.log
//  16  experiment: the critic is NOT evaluated inside the time loop (values / last_value are left unwritten) -- measures what an
//      actor-only loop + a separate throughput-bound values pass over the stored states could gain
constexpr int kVarPrefetch = 1, kVarIssuerPlace = 2, kVarArrive2 = 4, kVarEarlyStore = 8, kVarNoCritic = 16;
template <int V> constexpr int warps_of() { return (V & kVarIssuerPlace) ? 20 : kWorkerWarps + kIssuerWarps; }
constexpr int kTileCols = 128;   // TMEM columns per tile: X [0, 64) + D [64, 128)

// ---- dynamic shared memory map (bytes)
constexpr int kOffB2 = 0;                                       // [net][plane] 16 KB
constexpr int kOffB1 = kOffB2 + 4 * kB2PlaneBytes;              // [net][2] 2 KB
constexpr int kOffBb = kOffB1 + 4 * kB1Bytes;                   // [net] 2 KB
constexpr int kOffAc = kOffBb + 2 * kB1Bytes;                   // constant A, 4 KB
constexpr int kOffA1 = kOffAc + kA1Bytes;                       // [tile][net] 4 KB
constexpr int kSmallFloats = 96;                                // per net: w3[64], b3, avg[3], std[3] (+ padding)
constexpr int kOffSmall = kOffA1 + kTiles * 2 * kA1Bytes;
constexpr int kOffStage = kOffSmall + 2 * kSmallFloats * 4;     // worker warps: 14 x 96 fp32 (state rows -> 128-bit stores)
constexpr int kBarsPerTile = 7;                                 // bar_x, d1, d2, a2[4]
constexpr int kOffBars = kOffStage + kWorkerWarps * 96 * 4;
constexpr int kOffTmemSlot = kOffBars + kTiles * kBarsPerTile * 8;
constexpr int kSmemBytes = kOffTmemSlot + 16;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(kOffB1 % 1024 == 0 && kOffA1 % 128 == 0, "operand alignment");

constexpr int kW3 = 0, kB3 = 64, kAvg = 68, kStd = 72;         // small-parameter block (float offsets)

DEV NetImage net_image(int net) {
    NetImage im;
    im.b2[0] = kOffB2 + (net * 2 + 0) * kB2PlaneBytes;
    im.b2[1] = kOffB2 + (net * 2 + 1) * kB2PlaneBytes;
    im.b1[0] = kOffB1 + (net * 2 + 0) * kB1Bytes;
    im.b1[1] = kOffB1 + (net * 2 + 1) * kB1Bytes;
    im.bb = kOffBb + net * kB1Bytes;
    return im;
}

DEV void load_small(const b200rl_net& net, float* sm) {
    for (int i = threadIdx.x; i < kHid; i += blockDim.x) sm[kW3 + i] = net.weight[2][i];
    if (threadIdx.x == 0) sm[kB3] = net.bias[2][0];
    if (threadIdx.x < 3) {
        sm[kAvg + threadIdx.x] = net.state_avg ? net.state_avg[threadIdx.x] : 0.0f;
        sm[kStd + threadIdx.x] = net.state_std ? net.state_std[threadIdx.x] + 1e-4f : 1.0f;
    }
}

template <int V>
__global__ void __launch_bounds__(warps_of<V>() * 32, 1) rollout_pendulum_ts_kernel(const __grid_constant__ RolloutParams P) {
    constexpr bool kPrefetch = (V & kVarPrefetch) != 0, kArrive2 = (V & kVarArrive2) != 0, kEarly = (V & kVarEarlyStore) != 0;
    constexpr bool kCritic = (V & kVarNoCritic) == 0;
    constexpr int kThreads = warps_of<V>() * 32;
    extern __shared__ __align__(1024) unsigned char smem[];
    float* small = reinterpret_cast<float*>(smem + kOffSmall);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmemSlot);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = P.N, H = P.H;
    const int rows_cta = P.rows_per_cta;          // envs of this CTA: a multiple of 32 (whole warps), <= 448
    const int warps_cta = rows_cta >> 5;

    // ---- one-time setup: TMEM, mbarriers, operand images
    if (warp == 0) tc05::tmem_alloc<512>(tmem_slot);
    if (threadIdx.x == 32) {
        for (int t = 0; t < kTiles; ++t) {
            const int w = warps_cta - 4 * t;              // worker warps of the tile
            const uint32_t arrivals = w >= 4 ? 4u : (w > 0 ? (uint32_t)w : 1u);   // one elected lane per worker warp of the tile
            uint64_t* b = bars + t * kBarsPerTile;
            tc05::mbar_init(&b[0], arrivals);              // bar_x
            tc05::mbar_init(&b[1], 1);                     // d1 (tcgen05.commit)
            tc05::mbar_init(&b[2], 1);                     // d2
            for (int c = 0; c < 4; ++c) tc05::mbar_init(&b[3 + c], arrivals);   // variant 4 uses the first two
        }
        tc05::mbar_fence_init();
    }
    for (int which = 0; which < 2; ++which) {
        const b200rl_net& nn = which ? P.critic : P.actor;
        stage_net(nn, smem, net_image(which), threadIdx.x, kThreads);
        load_small(nn, small + which * kSmallFloats);
    }
    stage_const_a(smem + kOffAc, threadIdx.x, kThreads);
    for (int i = threadIdx.x; i < kTiles * 2 * kA1Bytes / 4; i += kThreads) reinterpret_cast<float*>(smem + kOffA1)[i] = 0.0f;
    tc05::fence_proxy_async_smem();
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    // issuer warp of tile i: warp 14 + i, or (variant 2) warps 14, 15, 18, 19 -- ids 2, 3 (mod 4), the schedulers with 3 workers
    int issuer_tile = -1;
    if (warp >= kWorkerWarps) {
        if (V & kVarIssuerPlace) issuer_tile = (warp & 2) ? ((warp - kWorkerWarps) >> 2) * 2 + (warp & 1) : -1;
        else issuer_tile = warp - kWorkerWarps;
    }
    if (warp >= kWorkerWarps) {
        // ======================================================================== issuer warp of one tile
        const int tile = issuer_tile;
        const bool tile_used = tile >= 0 && tile * kTileRows < rows_cta && blockIdx.x * rows_cta + tile * kTileRows < N;
        if (tile_used) {
            uint64_t* b = bars + tile * kBarsPerTile;
            const uint32_t tX = tmem_base + (uint32_t)(tile * kTileCols), tD = tX + kHid;
            const NetDescs nd_a = make_descs(smem, net_image(0)), nd_c = make_descs(smem, net_image(1));
            const uint64_t ac_desc = tc05::make_smem_desc(tc05::smem_u32(smem + kOffAc), kSboK8);
            const uint64_t a1_a = tc05::make_smem_desc(tc05::smem_u32(smem + kOffA1 + (tile * 2 + 0) * kA1Bytes), kSboK8);
            const uint64_t a1_c = tc05::make_smem_desc(tc05::smem_u32(smem + kOffA1 + (tile * 2 + 1) * kA1Bytes), kSboK8);
            uint32_t ph = 0;
            auto evaluate = [&](const NetDescs& nd, uint64_t a1_desc) {
                tc05::fence_after_thread_sync();
                if (tc05::elect_one()) {
                    issue_layer1(tX, a1_desc, nd);
                    tc05::mma_commit(&b[1]);
                }
                __syncwarp();
                constexpr int kArrivals = kArrive2 ? 2 : 4;
#pragma unroll 1
                for (int c = 0; c < kArrivals; ++c) {
                    tc05::mbar_wait(&b[3 + c], ph & 1);
                    tc05::fence_after_thread_sync();
                    if (tc05::elect_one()) {
                        if (c == 0) issue_bias(tD, ac_desc, nd);   // D is free: every worker read it before arriving
                        if (kArrive2) { issue_layer2_chunk(tD, tX, nd, 2 * c); issue_layer2_chunk(tD, tX, nd, 2 * c + 1); }
                        else issue_layer2_chunk(tD, tX, nd, c);
                        if (c == kArrivals - 1) tc05::mma_commit(&b[2]);
                    }
                    __syncwarp();
                }
                ph += 1;
            };
            for (int t = 0; t <= H; ++t) {
                tc05::mbar_wait(&b[0], t & 1);             // x~(t) of both nets is in shared memory
                if (t < H) evaluate(nd_a, a1_a);
                if (kCritic) evaluate(nd_c, a1_c);
            }
        }
    } else {
        // =================================================================== worker warp: thread = one env
        const int tile = warp >> 2, quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int n = blockIdx.x * rows_cta + tile * kTileRows + row;
        const int n_warp0 = n - lane;
        const bool live = n < N;
        const bool warp_used = n_warp0 < N;    // a warp without envs must still not deadlock its tile: see below
        const bool vec_ok = (n_warp0 + 32 <= N) && ((N & 3) == 0);
        uint64_t* b = bars + tile * kBarsPerTile;
        const uint32_t tX = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tile * kTileCols), tD = tX + kHid;
        unsigned char* a1_act = smem + kOffA1 + (tile * 2 + 0) * kA1Bytes;
        unsigned char* a1_cri = smem + kOffA1 + (tile * 2 + 1) * kA1Bytes;
        const float* sm_a = small;
        const float* sm_c = small + kSmallFloats;
        const bool norm_a = P.actor.state_avg != nullptr, norm_c = P.critic.state_avg != nullptr;
        float* my_stage = reinterpret_cast<float*>(smem + kOffStage) + warp * 96;
        // the tile's barriers expect one arrival per worker warp of the tile that holds envs of the CTA's LAST tile too: a
        // partially filled tile (ragged N) keeps all its warps running on dead rows (finite garbage, never stored)
        (void)warp_used;
        const bool tile_used = warp < warps_cta && blockIdx.x * rows_cta + tile * kTileRows < N;
        if (tile_used) {
            float theta = live ? P.theta[n] : 0.0f, theta_dot = live ? P.theta_dot[n] : 0.0f;
            int cur_step = live ? P.cur_step[n] : 0;
            const float sd = expf(P.actor.action_std_log[0]);
            const float log_sd = logf(sd), var2 = __fmul_rn(2.0f, __fmul_rn(sd, sd));
            float sin_t, cos_t;
            sincosf(theta, &sin_t, &cos_t);

            auto publish_obs = [&](float o0, float o1, float o2) {   // x~ of both nets -> shared memory -> bar_x
                float xa[3] = {o0, o1, o2}, xc[3] = {o0, o1, o2};
                if (norm_a) { xa[0] = (o0 - sm_a[kAvg]) / sm_a[kStd]; xa[1] = (o1 - sm_a[kAvg + 1]) / sm_a[kStd + 1]; xa[2] = (o2 - sm_a[kAvg + 2]) / sm_a[kStd + 2]; }
                if (norm_c) { xc[0] = (o0 - sm_c[kAvg]) / sm_c[kStd]; xc[1] = (o1 - sm_c[kAvg + 1]) / sm_c[kStd + 1]; xc[2] = (o2 - sm_c[kAvg + 2]) / sm_c[kStd + 2]; }
                write_x_row(a1_act, row, xa);
                if (kCritic) write_x_row(a1_cri, row, xc);
                tc05::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) tc05::mbar_arrive(&b[0]);
            };
            auto chunk_done = [&](int c) {   // the converted columns of chunk c (and, variant 4, c - 1) may be consumed
                if (kArrive2 && !(c & 1)) return;
                tc05::tmem_st_wait();
                tc05::fence_before_thread_sync();
                __syncwarp();
                if (lane == 0) tc05::mbar_arrive(&b[3 + (kArrive2 ? (c >> 1) : c)]);
            };
            auto hidden = [&]() {   // this row's 64 hidden units: fp32 -> GELU -> fp16 pairs in place, 16 at a time
                if (kPrefetch) {
                    float va[16], vb[16];
                    tc05::tmem_ld_32x32b_x16(tX, va);
                    tmem_ld_wait_tied(va);
                    tc05::tmem_ld_32x32b_x16(tX + 16, vb);
                    hidden_chunk_from_regs<true>(tX, va);
                    chunk_done(0);
                    tmem_ld_wait_tied(vb);
                    tc05::tmem_ld_32x32b_x16(tX + 32, va);
                    hidden_chunk_from_regs<true>(tX + 16, vb);
                    chunk_done(1);
                    tmem_ld_wait_tied(va);
                    tc05::tmem_ld_32x32b_x16(tX + 48, vb);
                    hidden_chunk_from_regs<true>(tX + 32, va);
                    chunk_done(2);
                    tmem_ld_wait_tied(vb);
                    hidden_chunk_from_regs<true>(tX + 48, vb);
                    chunk_done(3);
                } else {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        hidden_chunk_inplace<true>(tX + 16 * c);
                        chunk_done(c);
                    }
                }
            };
            auto head = [&](const float* sm) {
                return kPrefetch ? head_dot_pipelined<true>(tD, sm + kW3, sm[kB3]) : head_dot<true>(tD, sm + kW3, sm[kB3]);
            };
            auto store_state_row = [&](float* dst_states, float o0, float o1, float o2) {   // staged -> 128-bit stores
                if (vec_ok) {
                    my_stage[lane * 3 + 0] = o0; my_stage[lane * 3 + 1] = o1; my_stage[lane * 3 + 2] = o2;
                    __syncwarp();
                    if (lane < 24) reinterpret_cast<float4*>(dst_states + (size_t)n_warp0 * 3)[lane] = *reinterpret_cast<const float4*>(my_stage + 4 * lane);
                    __syncwarp();
                } else if (live) {
                    dst_states[(size_t)n * 3 + 0] = o0; dst_states[(size_t)n * 3 + 1] = o1; dst_states[(size_t)n * 3 + 2] = o2;
                }
            };

            publish_obs(cos_t, sin_t, theta_dot);
            uint32_t ph = 0;
            for (int t = 0; t <= H; ++t) {
                const bool last = t == H;
                const size_t rowbase = (size_t)t * N;
                const float obs0 = cos_t, obs1 = sin_t, obs2 = theta_dot;
                float action = 0.0f, logprob = 0.0f;
                float2 reset_u = make_float2(0.0f, 0.0f);
                if (!last) {
                    // ---------------------------------------------------------------- actor
                    tc05::mbar_wait(&b[1], ph & 1);
                    tc05::fence_after_thread_sync();
                    hidden();
                    // noise while the last UMMAs of the actor run (it does not depend on the policy output)
                    float e = 0.0f;
                    if (P.eps == nullptr || P.reset_noise == nullptr) {
                        RolloutNoise nz = rollout_noise(P.seed, (uint64_t)(P.env_offset + n), P.step_offset + (uint64_t)t, 0u);
                        e = nz.normal.x;
                        reset_u = nz.uniform;
                    }
                    if (P.eps && live) e = P.eps[rowbase + n];
                    if (P.deterministic) e = 0.0f;
                    if (P.reset_noise && live) reset_u = make_float2(P.reset_noise[(rowbase + n) * 2], P.reset_noise[(rowbase + n) * 2 + 1]);
                    if (kEarly) store_state_row(P.states + (size_t)t * N * 3, obs0, obs1, obs2);
                    tc05::mbar_wait(&b[2], ph & 1);
                    tc05::fence_after_thread_sync();
                    const float mu = head(sm_a);
                    tc05::fence_before_thread_sync();   // D reads are ordered before this thread's next arrival (critic chunk 0)
                    ph += 1;
                    action = __fadd_rn(__fmul_rn(e, sd), mu);
                    const float diff = __fsub_rn(action, mu);
                    logprob = __fsub_rn(__fsub_rn(-__fdiv_rn(__fmul_rn(diff, diff), var2), log_sd), kLogSqrt2Pi);
                }
                // -------------------------------------------------------------------- critic: V(s_t)
                if (kCritic) {
                    tc05::mbar_wait(&b[1], ph & 1);
                    tc05::fence_after_thread_sync();
                    hidden();
                }
                // state row of step t (or last_state)
                if (!kEarly || last) store_state_row(last ? P.last_state : P.states + (size_t)t * N * 3, obs0, obs1, obs2);
                if (!last) {
                    // env.step(tanh(action))  -- same op sequence as rollout.cu / envs/pendulum.py
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
                    if (truncate) {
                        new_theta = __fmul_rn(__fsub_rn(__fmul_rn(reset_u.x, 2.0f), 1.0f), kPi);
                        new_theta_dot = __fsub_rn(__fmul_rn(reset_u.y, 2.0f), 1.0f);
                        cur_step = 0;
                    }
                    theta = new_theta;
                    theta_dot = new_theta_dot;
                    sincosf(theta, &sin_t, &cos_t);
                    publish_obs(cos_t, sin_t, theta_dot);   // the next step's layer 1 overlaps the critic head below
                    if (live) {
                        P.actions[rowbase + n] = action;
                        P.logprobs[rowbase + n] = logprob;
                        P.rewards[rowbase + n] = reward;
                    }
                    if (vec_ok) {
                        const unsigned um_bits = __ballot_sync(0xffffffffu, !truncate);
                        if (lane < 8) {
                            unsigned m4 = (um_bits >> (4 * lane)) & 0xFu;
                            unsigned word = (m4 & 1u) | ((m4 & 2u) << 7) | ((m4 & 4u) << 14) | ((m4 & 8u) << 21);
                            reinterpret_cast<unsigned*>(P.unmasks + rowbase + n_warp0)[lane] = word;
                            reinterpret_cast<unsigned*>(P.undones + rowbase + n_warp0)[lane] = 0x01010101u;
                        }
                    } else if (live) {
                        P.unmasks[rowbase + n] = truncate ? 0 : 1;
                        P.undones[rowbase + n] = 1;
                    }
                }
                if (kCritic) {
                    tc05::mbar_wait(&b[2], ph & 1);
                    tc05::fence_after_thread_sync();
                    const float val = head(sm_c);
                    tc05::fence_before_thread_sync();
                    ph += 1;
                    if (live) {
                        if (!last) { if (P.values) P.values[rowbase + n] = val; }
                        else if (P.last_value) P.last_value[n] = val;
                    }
                }
            }
            if (live) { P.theta[n] = theta; P.theta_dot[n] = theta_dot; P.cur_step[n] = cur_step; }
        }
    }

    tc05::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<512>(tmem_base);
}

}  // namespace

template <int V>
static int launch_variant(const RolloutParams& P, cudaStream_t stream) {
    B200RL_CHECK_CUDA(cudaFuncSetAttribute(rollout_pendulum_ts_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const int grid = (P.N + P.rows_per_cta - 1) / P.rows_per_cta;
    rollout_pendulum_ts_kernel<V><<<grid, warps_of<V>() * 32, kSmemBytes, stream>>>(P);
    B200RL_COUNT_LAUNCH(1);
    B200RL_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Envs per CTA: the fewest whole warps that still cover N with one CTA per SM (148 SMs; 448 = 3.5 tiles at 65 536 envs).
// Small shards (an 8-GPU split of 65 536 envs, BASELINE configs[2]'s 2 048 envs per GPU) get short CTAs on MANY SMs
// instead of a few full ones: the kernel's time is the per-CTA step latency times H, which shrinks with the rows.
static int rows_per_cta_for(int N) {
    const char* o = getenv("B200RL_TS_ROWS");
    int rows = o ? atoi(o) : ((N + 147) / 148 + 31) / 32 * 32;
    rows = rows < 32 ? 32 : (rows > kRowsPerCta ? kRowsPerCta : rows);
    return rows / 32 * 32;
}

int b200rl_launch_rollout_ts(const RolloutParams& P_in, cudaStream_t stream) {
    RolloutParams P = P_in;
    P.rows_per_cta = rows_per_cta_for(P.N);
    const char* v = getenv("B200RL_TS_VARIANT");
    // built: the default, its baseline and the actor-only experiment; the other scheduling variants measured in
    // profiles/r02_v3_rollout_ts_variants.log (bits 1, 4, 8) are template parameters away
    switch (v ? atoi(v) : kDefaultVariant) {
        case 0: return launch_variant<0>(P, stream);
        case 2: return launch_variant<2>(P, stream);
        case 18: return launch_variant<18>(P, stream);
        default: break;
    }
    b200rl_set_error("rollout_ts: B200RL_TS_VARIANT=%s is not built", v);
    return 3;
}
