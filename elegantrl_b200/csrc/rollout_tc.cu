// Fused Pendulum rollout, tcgen05 / TMEM implementation for the 3 -> 64 -> 64 -> 1 GELU actor + critic
// (BASELINE config 2).  Same contract and outputs as rollout.cu (reference AgentPPO._explore_vec_env,
// elegantrl/agents/AgentPPO.py:87-129, + values pass :141-143 + V(last_state) :219-220); parity vs the oracle
// rtol 1e-4.
//
// Mapping.  One persistent CTA per SM owns 448 envs = 3.5 tiles of 128 rows for all H steps (65 536 envs ->
// 147 CTAs on 148 SMs; env state lives in registers).  896 threads = 28 warps = 7 per SM sub-partition.  A warp
// serves ONE net for 32 env rows: 8 warp-groups g = (net, tile); the half tile has 2 warps per net, the
// actor's on TMEM lane quarters 0-1 and the critic's on quarters 2-3 (a warp may only touch TMEM lanes
// 32 * (warp_id % 4) .. +31, and both nets keep separate A operands / accumulators, so each net can place the 64
// live rows of the half tile where its warp ids allow).  Actor warps also step the env and own the action /
// logprob / reward / mask stores, critic warps own the state / value stores; the next observation goes from
// actor to critic through shared memory guarded by full / empty mbarriers.
//
// Per step and group (the only dense contraction, Linear 64x64, goes to the tensor core):
//   layer 1 (K = 3) + GELU on CUDA cores, 8 hidden units at a time -> split hi/lo (3xTF32) -> K-major A chunk
//   [128 rows x 8] in a 2-slot shared-memory ring -> fence.proxy.async + warp-aggregated acq_rel counter; the
//   LAST warp of the group to arrive issues 3 tcgen05.mma (Ahi*Bhi, Alo*Bhi, Ahi*Blo; M=128, N=64, K=8) into the
//   group's 64 TMEM columns and commits to the slot's mbarrier (slot reuse) -- no warp ever blocks on its peers;
//   after the 8th chunk the accumulator is complete: tcgen05.ld 32x32b (thread = row) -> bias + GELU + dot with
//   the output layer on CUDA cores.
// W2 (both nets, hi/lo planes, K-major exactly as nn.Linear stores it) stays resident in shared memory (64 KB).
//
// GELU: exact-erf GELU(x) = max(x,0) - 0.5|x| erfc(|x|/sqrt2), erfc(z) = exp2(-z P(z)) with a degree-5 minimax P
// (tools/fit_gelu.py; max abs error of GELU 5.8e-7 in fp32), evaluated two values at a time with packed
// FFMA2 -- 6.5 issue slots per GELU instead of ~31 for erff.
#include "rollout_params.cuh"
#include "tc05.cuh"

namespace {

constexpr int kHid = 64;
constexpr int kTileRows = 128, kTiles = 4, kRowsPerCta = 448, kWarps = 28, kThreads = kWarps * 32, kGroups = 8;
constexpr int kChunks = kHid / 8;  // A chunks (one UMMA K-step of 8 tf32 each)

// ---- dynamic shared memory map (bytes)
constexpr int kPlaneB = kHid * kHid * 4;                   // one 64x64 fp32 plane of W2: 16 KB
constexpr int kOffB = 0;                                   // [net][hi/lo] planes
constexpr int kSlotBytes = 2 * kTileRows * 8 * 4;          // hi plane 4 KB + lo plane 4 KB
constexpr int kOffRing = kOffB + 4 * kPlaneB;              // [group][slot]
constexpr int kSmallFloats = 512;                          // per net: W1t[3][64], b1, b2, w3, b3, avg, std
constexpr int kOffSmall = kOffRing + kGroups * 2 * kSlotBytes;
constexpr int kOffObs = kOffSmall + 2 * kSmallFloats * 4;  // [tile][slot][3][128] fp32
constexpr int kOffStage = kOffObs + kTiles * 2 * 3 * kTileRows * 4;   // critic warps: 14 x 96 fp32
constexpr int kOffBars = kOffStage + 14 * 96 * 4;          // mbarriers
constexpr int kNumBars = kGroups * 2 + kGroups + kTiles * 2 + kTiles * 2;
constexpr int kOffFill = kOffBars + kNumBars * 8;          // [group][2] uint32 arrival counters
constexpr int kOffTmemSlot = kOffFill + kGroups * 2 * 4;
constexpr int kSmemBytes = kOffTmemSlot + 16;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// small-parameter block (float offsets)
constexpr int kW1t = 0, kB1 = 192, kB2 = 256, kW3 = 320, kB3 = 384, kAvg = 388, kStd = 392;

// Packed GELU on a pair of values (FFMA2: one issue slot per two fp32 FMAs; tools/fit_gelu.py packed_form):
//   zn = -min(|x|, L);  t = zn * Pt(zn) - 1;  GELU(x) = max(x, 0) + zn * exp2(t)       [exp2(t) = 0.5 erfc(|x|/sqrt2)]
// 4 FMNMX + 7 FFMA2 + 2 MUFU.EX2 per pair; max abs error 5.8e-7 (fp32 rounding of the final FMA dominates).
DEV float2 splat(float v) { return make_float2(v, v); }
DEV float2 gelu_fast2(float2 x) {
    constexpr float L = 6.2225397f;
    const float2 zn = make_float2(fmaxf(-fabsf(x.x), -L), fmaxf(-fabsf(x.y), -L));
    const float2 r = make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f));
    float2 p = __ffma2_rn(splat(1.775934289e-05f), zn, splat(6.477866232e-04f));
    p = __ffma2_rn(p, zn, splat(7.724114180e-03f));
    p = __ffma2_rn(p, zn, splat(5.292681266e-02f));
    p = __ffma2_rn(p, zn, splat(-4.590827042e-01f));
    p = __ffma2_rn(p, zn, splat(1.151116861e+00f));
    const float2 t = __ffma2_rn(p, zn, splat(-1.0f));
    const float2 e = make_float2(tc05::ex2_approx(t.x), tc05::ex2_approx(t.y));
    return __ffma2_rn(zn, e, r);
}

DEV uint32_t atom_add_acq_rel_smem(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(tc05::smem_u32(p)), "r"(v) : "memory");
    return old;
}

struct GroupCtx {
    uint32_t ring_addr;     // shared address of this group's 2-slot A ring
    uint32_t a_desc_lo;     // low words of the shared-memory matrix descriptors: A slot 0 hi plane, B hi / lo plane chunk 0
    uint32_t bhi_desc_lo, blo_desc_lo;
    uint64_t* slot_free;    // [2]
    uint64_t* d_ready;
    uint32_t* fill;         // [2] arrival counters of the two ring slots (monotonic)
    uint32_t tmem_d;        // TMEM address (lane field = this warp's quarter, column = 64 * group)
    uint32_t group_warps;   // 4, or 2 for the half tile
    uint32_t evals;         // completed evaluations (parity source)
    const float* small;     // this net's small-parameter block
    uint32_t row_off;       // byte offset of this thread's row inside a chunk plane
};

// one MLP evaluation of this group's net for this thread's row.  x = normalised observation.
DEV float eval_net(GroupCtx& c, const float (&x)[3], int lane) {
    const float* sm = c.small;
    constexpr uint32_t idesc = tc05::make_idesc_tf32(kTileRows, kHid);
    const float2 x2[3] = {splat(x[0]), splat(x[1]), splat(x[2])};
    // 4 rounds of 16 hidden units: two 8-column chunks are produced back to back into the two ring slots, then ONE
    // proxy fence + arrival covers both and the last-arriving warp issues the 6 MMAs of the two K-steps
#pragma unroll 1
    for (int cp = 0; cp < kChunks / 2; ++cp) {
        const uint32_t use = c.evals * (kChunks / 2) + cp;  // how often the slot pair has been filled before
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int ch = cp * 2 + half;
            float2 hi[4], lo[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = ch * 8 + h * 4;
                const float4 b = *reinterpret_cast<const float4*>(sm + kB1 + j);
                const float4 w0 = *reinterpret_cast<const float4*>(sm + kW1t + j);
                const float4 w1 = *reinterpret_cast<const float4*>(sm + kW1t + 64 + j);
                const float4 w2 = *reinterpret_cast<const float4*>(sm + kW1t + 128 + j);
                float2 a01 = __ffma2_rn(x2[0], make_float2(w0.x, w0.y), make_float2(b.x, b.y));
                float2 a23 = __ffma2_rn(x2[0], make_float2(w0.z, w0.w), make_float2(b.z, b.w));
                a01 = __ffma2_rn(x2[1], make_float2(w1.x, w1.y), a01);
                a23 = __ffma2_rn(x2[1], make_float2(w1.z, w1.w), a23);
                a01 = __ffma2_rn(x2[2], make_float2(w2.x, w2.y), a01);
                a23 = __ffma2_rn(x2[2], make_float2(w2.z, w2.w), a23);
                const float2 g01 = gelu_fast2(a01), g23 = gelu_fast2(a23);
                hi[h * 2 + 0] = make_float2(tc05::tf32_hi(g01.x), tc05::tf32_hi(g01.y));
                hi[h * 2 + 1] = make_float2(tc05::tf32_hi(g23.x), tc05::tf32_hi(g23.y));
                lo[h * 2 + 0] = __ffma2_rn(hi[h * 2 + 0], splat(-1.0f), g01);
                lo[h * 2 + 1] = __ffma2_rn(hi[h * 2 + 1], splat(-1.0f), g23);
            }
            // the MMAs that read the previous content of the slot pair must be done before it is overwritten
            if (half == 0 && use > 0) tc05::mbar_wait(&c.slot_free[0], (use - 1) & 1);
            const uint32_t slot = c.ring_addr + half * kSlotBytes + c.row_off;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "f"(hi[0].x), "f"(hi[0].y), "f"(hi[1].x), "f"(hi[1].y) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot + 128), "f"(hi[2].x), "f"(hi[2].y), "f"(hi[3].x), "f"(hi[3].y) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot + 4096), "f"(lo[0].x), "f"(lo[0].y), "f"(lo[1].x), "f"(lo[1].y) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot + 4096 + 128), "f"(lo[2].x), "f"(lo[2].y), "f"(lo[3].x), "f"(lo[3].y) : "memory");
        }
        tc05::fence_proxy_async_smem();
        __syncwarp();
        // warp-aggregated arrival; the last warp of the group issues the MMAs of both chunks.  The outcome is voted so that
        // the issue code runs in warp-UNIFORM control flow (descriptors in uniform registers, one elected lane issues)
        uint32_t old = 0u;
        if (lane == 0) old = atom_add_acq_rel_smem(&c.fill[0], 1u);
        const bool last = __ballot_sync(0xffffffffu, lane == 0 && old == c.group_warps * (use + 1) - 1) != 0u;
        if (last) {
            tc05::fence_after_thread_sync();
            const uint32_t d = c.tmem_d & 0x0000FFFFu;  // lane field 0: the MMA addresses the whole 128-lane tile
            if (tc05::elect_one()) {
                // descriptors = precomputed low word + (byte offset >> 4); high words: SBO >> 4 | version 1 (bit 46)
                constexpr uint64_t kHiA = (uint64_t)(0x4000u | (256u >> 4)) << 32, kHiB = (uint64_t)(0x4000u | (2048u >> 4)) << 32;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int ch = cp * 2 + half;
                    const uint64_t a_hi = kHiA | (uint64_t)(c.a_desc_lo + half * (kSlotBytes >> 4));
                    const uint64_t a_lo = kHiA | (uint64_t)(c.a_desc_lo + half * (kSlotBytes >> 4) + (4096 >> 4));
                    const uint64_t b_hi = kHiB | (uint64_t)(c.bhi_desc_lo + ch * (256 >> 4));
                    const uint64_t b_lo = kHiB | (uint64_t)(c.blo_desc_lo + ch * (256 >> 4));
                    tc05::mma_tf32(d, a_hi, b_hi, idesc, ch > 0);
                    tc05::mma_tf32(d, a_lo, b_hi, idesc, true);
                    tc05::mma_tf32(d, a_hi, b_lo, idesc, true);
                }
                tc05::mma_commit(&c.slot_free[0]);
                if (cp == kChunks / 2 - 1) tc05::mma_commit(c.d_ready);
            }
            __syncwarp();
        }
    }
    tc05::mbar_wait(c.d_ready, c.evals & 1);
    c.evals += 1;
    tc05::fence_after_thread_sync();
    float2 out2 = make_float2(sm[kB3], 0.0f);
#pragma unroll 1
    for (int cc = 0; cc < kHid / 16; ++cc) {
        float v[16];
        tc05::tmem_ld_32x32b_x16(c.tmem_d + cc * 16, v);
        tc05::tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            const float4 b = *reinterpret_cast<const float4*>(sm + kB2 + cc * 16 + q4 * 4);
            const float4 w = *reinterpret_cast<const float4*>(sm + kW3 + cc * 16 + q4 * 4);
            const float2 g01 = gelu_fast2(__fadd2_rn(make_float2(v[q4 * 4 + 0], v[q4 * 4 + 1]), make_float2(b.x, b.y)));
            const float2 g23 = gelu_fast2(__fadd2_rn(make_float2(v[q4 * 4 + 2], v[q4 * 4 + 3]), make_float2(b.z, b.w)));
            out2 = __ffma2_rn(g01, make_float2(w.x, w.y), out2);
            out2 = __ffma2_rn(g23, make_float2(w.z, w.w), out2);
        }
    }
    const float out = out2.x + out2.y;
    // TMEM reads of this evaluation are ordered before the (release) arrival that precedes the next evaluation's MMAs
    tc05::fence_before_thread_sync();
    return out;
}

DEV void load_small(const b200rl_net& net, float* sm) {
    for (int i = threadIdx.x; i < 3 * kHid; i += kThreads) { int k = i / kHid, j = i - k * kHid; sm[kW1t + i] = net.weight[0][j * 3 + k]; }
    for (int i = threadIdx.x; i < kHid; i += kThreads) {
        sm[kB1 + i] = net.bias[0][i];
        sm[kB2 + i] = net.bias[1][i];
        sm[kW3 + i] = net.weight[2][i];
    }
    if (threadIdx.x == 0) sm[kB3] = net.bias[2][0];
    if (threadIdx.x < 3) {
        sm[kAvg + threadIdx.x] = net.state_avg ? net.state_avg[threadIdx.x] : 0.0f;
        sm[kStd + threadIdx.x] = net.state_std ? net.state_std[threadIdx.x] + 1e-4f : 1.0f;
    }
}

__global__ void __launch_bounds__(kThreads, 1) rollout_pendulum_tc_kernel(const __grid_constant__ RolloutParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* small = reinterpret_cast<float*>(smem + kOffSmall);
    float* obs_sm = reinterpret_cast<float*>(smem + kOffObs);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
    uint64_t* slot_free = bars;                          // [group][2]
    uint64_t* d_ready = bars + kGroups * 2;              // [group]
    uint64_t* obs_full = d_ready + kGroups;              // [tile][2]
    uint64_t* obs_empty = obs_full + kTiles * 2;         // [tile][2]
    uint32_t* fill = reinterpret_cast<uint32_t*>(smem + kOffFill);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmemSlot);

    // ---- warp roles (see file header): warp_id % 4 is always the TMEM lane quarter the warp works on
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int net, tile, critic_warp;
    if (warp < 24) { net = warp / 12; tile = (warp % 12) >> 2; critic_warp = warp - 12; }
    else { net = (warp - 24) >> 1; tile = 3; critic_warp = 12 + (warp - 26); }
    const int quarter = warp & 3;
    const int group = net * kTiles + tile;
    const int row = quarter * 32 + lane;                              // A / accumulator row inside the tile
    const int env_in_tile = (tile == 3 && net == 1) ? row - 64 : row; // half tile: critic rows 64..127 <-> envs 0..63
    const int tile_envs = tile == 3 ? 64 : 128;
    const int N = P.N;
    const int n = blockIdx.x * kRowsPerCta + tile * kTileRows + env_in_tile;
    const int n_warp0 = n - lane;
    const bool live = n < N;
    const bool vec_ok = (n_warp0 + 32 <= N) && ((N & 3) == 0);

    // ---- one-time setup: TMEM, mbarriers, weights
    if (warp == 0) tc05::tmem_alloc<512>(tmem_slot);
    if (threadIdx.x == 32) {
        for (int i = 0; i < kGroups * 2 + kGroups; ++i) tc05::mbar_init(&bars[i], 1);
        for (int t = 0; t < kTiles; ++t)
            for (int sl = 0; sl < 2; ++sl) {
                tc05::mbar_init(&obs_full[t * 2 + sl], t == 3 ? 64 : 128);
                tc05::mbar_init(&obs_empty[t * 2 + sl], t == 3 ? 64 : 128);
            }
        tc05::mbar_fence_init();
    }
    if (threadIdx.x < kGroups * 2) fill[threadIdx.x] = 0u;
    for (int which = 0; which < 2; ++which) {
        const b200rl_net& nn = which ? P.critic : P.actor;
        float* bhi = reinterpret_cast<float*>(smem + kOffB + (which * 2 + 0) * kPlaneB);
        float* blo = reinterpret_cast<float*>(smem + kOffB + (which * 2 + 1) * kPlaneB);
        for (int i = threadIdx.x; i < kHid * kHid; i += kThreads) {
            const int r = i >> 6, k = i & 63;
            const float w = nn.weight[1][i];
            const float h = tc05::tf32_hi(w);
            const uint32_t off = tc05::operand_offset(r, k, kHid) >> 2;
            bhi[off] = h;
            blo[off] = w - h;
        }
        load_small(nn, small + which * kSmallFloats);
    }
    // rows of the half tile's A operands that no thread ever writes: keep them finite (they feed unused TMEM lanes)
    for (int i = threadIdx.x; i < 2 * 2 * kSlotBytes / 4; i += kThreads) {
        const int g = (i < 2 * kSlotBytes / 4) ? 3 : 7;
        reinterpret_cast<float*>(smem + kOffRing + g * 2 * kSlotBytes)[i % (2 * kSlotBytes / 4)] = 0.0f;
    }
    tc05::fence_proxy_async_smem();
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = *tmem_slot;

    GroupCtx ctx;
    ctx.ring_addr = tc05::smem_u32(smem + kOffRing + group * 2 * kSlotBytes);
    ctx.a_desc_lo = (uint32_t)tc05::make_smem_desc(ctx.ring_addr, 256);
    ctx.bhi_desc_lo = (uint32_t)tc05::make_smem_desc(tc05::smem_u32(smem + kOffB + (net * 2 + 0) * kPlaneB), 2048);
    ctx.blo_desc_lo = (uint32_t)tc05::make_smem_desc(tc05::smem_u32(smem + kOffB + (net * 2 + 1) * kPlaneB), 2048);
    ctx.slot_free = slot_free + group * 2;
    ctx.d_ready = d_ready + group;
    ctx.fill = fill + group * 2;
    ctx.tmem_d = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(group * kHid);
    ctx.group_warps = tile == 3 ? 2u : 4u;
    ctx.evals = 0;
    ctx.small = small + net * kSmallFloats;
    ctx.row_off = (uint32_t)((row >> 3) * 256 + (row & 7) * 16);
    const float* sm = ctx.small;
    const float avg0 = sm[kAvg], avg1 = sm[kAvg + 1], avg2 = sm[kAvg + 2];
    const float std0 = sm[kStd], std1 = sm[kStd + 1], std2 = sm[kStd + 2];
    const bool has_norm = net ? (P.critic.state_avg != nullptr) : (P.actor.state_avg != nullptr);
    float* obs_tile = obs_sm + tile * (2 * 3 * kTileRows);  // [slot][3][128], indexed by env_in_tile
    (void)tile_envs;

    if (net == 0) {
        // =================================================================== actor warps: policy + env
        float theta = live ? P.theta[n] : 0.0f, theta_dot = live ? P.theta_dot[n] : 0.0f;
        int cur_step = live ? P.cur_step[n] : 0;
        const float sd = expf(P.actor.action_std_log[0]);
        const float log_sd = logf(sd), var2 = __fmul_rn(2.0f, __fmul_rn(sd, sd));
        float sin_t, cos_t;
        sincosf(theta, &sin_t, &cos_t);
        obs_tile[0 * kTileRows + env_in_tile] = cos_t; obs_tile[1 * kTileRows + env_in_tile] = sin_t; obs_tile[2 * kTileRows + env_in_tile] = theta_dot;
        tc05::mbar_arrive(&obs_full[tile * 2 + 0]);

        for (int t = 0; t < P.H; ++t) {
            const size_t rowbase = (size_t)t * N;
            float x[3] = {cos_t, sin_t, theta_dot};
            if (has_norm) { x[0] = (x[0] - avg0) / std0; x[1] = (x[1] - avg1) / std1; x[2] = (x[2] - avg2) / std2; }
            // noise first: it does not depend on the policy output, so it is off the critical path to the next observation
            float e = 0.0f;
            float2 reset_u = make_float2(0.0f, 0.0f);
            if (P.eps == nullptr || P.reset_noise == nullptr) {
                RolloutNoise nz = rollout_noise(P.seed, (uint64_t)(P.env_offset + n), P.step_offset + (uint64_t)t, 0u);
                e = nz.normal.x;
                reset_u = nz.uniform;
            }
            if (P.eps && live) e = P.eps[rowbase + n];
            if (P.deterministic) e = 0.0f;
            if (P.reset_noise && live) reset_u = make_float2(P.reset_noise[(rowbase + n) * 2], P.reset_noise[(rowbase + n) * 2 + 1]);
            const float mu = eval_net(ctx, x, lane);
            const float action = __fadd_rn(__fmul_rn(e, sd), mu);
            const float diff = __fsub_rn(action, mu);
            const float logprob = __fsub_rn(__fsub_rn(-__fdiv_rn(__fmul_rn(diff, diff), var2), log_sd), kLogSqrt2Pi);

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

            // hand the next observation to the critic warps of this tile (slot (t+1) & 1)
            {
                const int t1 = t + 1, slot = t1 & 1;
                if (t1 >= 2) tc05::mbar_wait(&obs_empty[tile * 2 + slot], ((t1 >> 1) - 1) & 1);
                float* o = obs_tile + slot * (3 * kTileRows);
                o[0 * kTileRows + env_in_tile] = cos_t; o[1 * kTileRows + env_in_tile] = sin_t; o[2 * kTileRows + env_in_tile] = theta_dot;
                tc05::mbar_arrive(&obs_full[tile * 2 + slot]);
            }
            // trajectory stores owned by the actor warps: action, logprob, reward, masks
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
        if (live) { P.theta[n] = theta; P.theta_dot[n] = theta_dot; P.cur_step[n] = cur_step; }
    } else {
        // ============================================== critic warps: V(s_t), state stores, V(last_state)
        float* my_stage = reinterpret_cast<float*>(smem + kOffStage) + critic_warp * 96;
        for (int t = 0; t <= P.H; ++t) {
            const int slot = t & 1;
            tc05::mbar_wait(&obs_full[tile * 2 + slot], (t >> 1) & 1);
            const float* o = obs_tile + slot * (3 * kTileRows);
            const float obs0 = o[0 * kTileRows + env_in_tile], obs1 = o[1 * kTileRows + env_in_tile], obs2 = o[2 * kTileRows + env_in_tile];
            tc05::mbar_arrive(&obs_empty[tile * 2 + slot]);
            float x[3] = {obs0, obs1, obs2};
            if (has_norm) { x[0] = (x[0] - avg0) / std0; x[1] = (x[1] - avg1) / std1; x[2] = (x[2] - avg2) / std2; }
            const float val = eval_net(ctx, x, lane);
            const bool last = (t == P.H);
            float* dst_states = last ? P.last_state : P.states + (size_t)t * N * 3;
            if (vec_ok) {
                my_stage[lane * 3 + 0] = obs0; my_stage[lane * 3 + 1] = obs1; my_stage[lane * 3 + 2] = obs2;
                __syncwarp();
                if (lane < 24) reinterpret_cast<float4*>(dst_states + (size_t)n_warp0 * 3)[lane] = *reinterpret_cast<const float4*>(my_stage + 4 * lane);
                __syncwarp();
            } else if (live) {
                dst_states[(size_t)n * 3 + 0] = obs0; dst_states[(size_t)n * 3 + 1] = obs1; dst_states[(size_t)n * 3 + 2] = obs2;
            }
            if (live) {
                if (!last) { if (P.values) P.values[(size_t)t * N + n] = val; }
                else if (P.last_value) P.last_value[n] = val;
            }
        }
    }

    tc05::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<512>(tmem_base);
}

}  // namespace

int b200rl_launch_rollout_tc(const RolloutParams& P, cudaStream_t stream) {
    // per-device attribute: set on every launch (a process may drive more than one GPU)
    B200RL_CHECK_CUDA(cudaFuncSetAttribute(rollout_pendulum_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const int grid = (P.N + kRowsPerCta - 1) / kRowsPerCta;
    rollout_pendulum_tc_kernel<<<grid, kThreads, kSmemBytes, stream>>>(P);
    B200RL_COUNT_LAUNCH(1);
    B200RL_CHECK_CUDA(cudaGetLastError());
    return 0;
}
