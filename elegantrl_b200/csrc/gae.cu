// GAE reverse scan over [horizon, num_envs] -- one kernel for reference AgentPPO.get_advantages
// (elegantrl/agents/AgentPPO.py:207-232), reward_sums (:146) and the reduction inputs of the advantage
// normalisation (:149).  HBM-bound: 18 B per env-step (read r 4, V 4, undone 1, unmask 1; write adv 4, rsum 4).
//
// Layout: everything is time-major [H, N]; lane <-> env, so every load/store of a warp is one contiguous
// 128 B (fp32) / 32 B (bool) segment.  The recurrence  y_t = d_t + c_t * y_{t+1}  is linear, so the time axis
// is cut into `chunks` segments scanned by different warps of the CTA (blockDim.y): pass 1 reduces each
// segment to its affine map (a, P) : y_in -> a + P * y_in, the maps are staged in shared memory and folded
// (suffix composition), pass 2 re-scans each segment from its true carry-in and writes the outputs.
// With chunks == 1 (what large num_envs uses: the env axis alone fills the GPU) pass 1 is skipped and the
// scan is the reference's sequential order op for op (bit-exact; no FMA contraction).
#include "common.cuh"

namespace {

constexpr int kUnroll = 8;  // time steps whose loads are in flight together (memory-level parallelism)

struct StepIn {
    float r, v;
    bool undone, unmask;
};

// one step of the recurrence.  y = carried value, vnext = V_{t+1}.  Returns the advantage A_t.
template <bool VTRACE>
DEV float gae_step(const StepIn& s, float gamma, float lam, float& y, float& vnext, float& r_fixed, bool& undone_fixed) {
    const bool trunc = !s.unmask;
    // rewards[trunc] += V(s_trunc); undones[trunc] = False        (reference :211-214)
    r_fixed = trunc ? __fadd_rn(s.r, s.v) : s.r;
    undone_fixed = s.undone && !trunc;
    const float m = undone_fixed ? gamma : 0.0f;  // masks = undones * gamma  (:216)
    float adv;
    if (VTRACE) {  // :223-227
        float nv = __fadd_rn(r_fixed, __fmul_rn(m, vnext));
        adv = __fadd_rn(__fsub_rn(nv, s.v), __fmul_rn(__fmul_rn(m, lam), y));
        y = adv;
        vnext = s.v;
    } else {  // :228-231
        adv = __fadd_rn(__fsub_rn(r_fixed, s.v), __fmul_rn(m, y));
        y = __fadd_rn(s.v, __fmul_rn(lam, adv));
    }
    return adv;
}

template <bool VTRACE>
__global__ void __launch_bounds__(512) gae_kernel(float* __restrict__ rewards, uint8_t* __restrict__ undones,
                           const uint8_t* __restrict__ unmasks, const float* __restrict__ values,
                           const float* __restrict__ last_value, int H, int N, float gamma, float lam,
                           int64_t env_offset, float* __restrict__ adv_out, float* __restrict__ rsum_out,
                           double* stat_sums) {
    extern __shared__ float2 agg[];  // [chunks][blockDim.x] affine maps (a, P)
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunks = blockDim.y, cy = threadIdx.y;
    const int len = (H + chunks - 1) / chunks;
    const int t_lo = cy * len, t_hi = min(H, t_lo + len);
    const bool live = n < N && t_lo < t_hi;

    float y_in = 0.0f;
    if (chunks > 1) {
        // pass 1: affine map of this time segment
        float a = 0.0f, p = 1.0f;
        if (live) {
            float vnext = (t_hi == H) ? last_value[n] : values[(size_t)t_hi * N + n];
            float y = 0.0f;
            for (int t = t_hi - 1; t >= t_lo; --t) {
                size_t i = (size_t)t * N + n;
                StepIn s{rewards[i], values[i], undones[i] != 0, unmasks[i] != 0};
                float rf; bool uf;
                gae_step<VTRACE>(s, gamma, lam, y, vnext, rf, uf);
                p *= (uf ? gamma : 0.0f) * lam;
            }
            a = y;
        }
        agg[cy * blockDim.x + threadIdx.x] = make_float2(a, p);
        __syncthreads();
        for (int c = chunks - 1; c > cy; --c) {  // fold the later segments: y at the end of mine
            float2 m = agg[c * blockDim.x + threadIdx.x];
            y_in = m.x + m.y * y_in;
        }
    }

    double s_all = 0.0, s_all2 = 0.0, s_lat = 0.0, s_lat2 = 0.0;
    if (live) {
        // pass 2: scan with the true carry-in, write outputs, fix rewards/undones in place
        float vnext = (t_hi == H) ? last_value[n] : values[(size_t)t_hi * N + n];
        float y = y_in;
        const bool lat_env = ((env_offset + n) & 3) == 0;
        // software pipeline: the loads of the next kUnroll time steps are in flight while this batch is scanned
        StepIn s[kUnroll], nx[kUnroll];
        auto load_batch = [&](StepIn (&dst)[kUnroll], int t1) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                int t = t1 - 1 - u;
                if (t >= t_lo) {
                    size_t i = (size_t)t * N + n;
                    dst[u] = StepIn{rewards[i], values[i], undones[i] != 0, unmasks[i] != 0};
                }
            }
        };
        load_batch(s, t_hi);
        for (int t1 = t_hi; t1 > t_lo; t1 -= kUnroll) {
            if (t1 - kUnroll > t_lo) load_batch(nx, t1 - kUnroll);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                int t = t1 - 1 - u;
                if (t >= t_lo) {
                    size_t i = (size_t)t * N + n;
                    float rf; bool uf;
                    float adv = gae_step<VTRACE>(s[u], gamma, lam, y, vnext, rf, uf);
                    adv_out[i] = adv;
                    rsum_out[i] = __fadd_rn(adv, s[u].v);  // reward_sums = advantages + values  (:146)
                    if (!s[u].unmask) { rewards[i] = rf; undones[i] = 0; }
                    s_all += (double)adv;
                    s_all2 += (double)adv * (double)adv;
                    if (lat_env && (t & 3) == 0) { s_lat += (double)adv; s_lat2 += (double)adv * (double)adv; }
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) s[u] = nx[u];
        }
    }
    // block reduction of the three sums -> one atomicAdd(double) each per CTA
    __shared__ double red[4][32];
    s_all = warp_sum(s_all); s_lat = warp_sum(s_lat); s_lat2 = warp_sum(s_lat2); s_all2 = warp_sum(s_all2);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nwarps = (blockDim.x * blockDim.y + 31) >> 5;
    if ((tid & 31) == 0) { red[0][tid >> 5] = s_all; red[1][tid >> 5] = s_lat; red[2][tid >> 5] = s_lat2; red[3][tid >> 5] = s_all2; }
    __syncthreads();
    if (tid < 32) {
        double a = tid < nwarps ? red[0][tid] : 0.0, b = tid < nwarps ? red[1][tid] : 0.0, c = tid < nwarps ? red[2][tid] : 0.0;
        double d = tid < nwarps ? red[3][tid] : 0.0;
        a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); d = warp_sum(d);
        if (tid == 0) { atomicAdd(stat_sums + 0, a); atomicAdd(stat_sums + 1, b); atomicAdd(stat_sums + 2, c); atomicAdd(stat_sums + 3, d); }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Large-N variant of the sequential scan: same thread <-> env mapping and the same op order (bit-exact with the
// reference), but the four input streams are staged through shared memory by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx): one elected thread keeps a whole batch of kStage time steps (20 KB per CTA)
// in flight ahead of the scan, which the register-prefetch kernel above cannot afford (it is latency-bound at ~14 %
// of DRAM throughput, profiles/r01_gae_*).  Requires 128 envs per CTA and N % 128 == 0 (16-byte aligned row segments).
constexpr int kGaeEnvs = 128;
constexpr int kStage = 16;
struct __align__(128) GaeStage {
    float r[kStage][kGaeEnvs];
    float v[kStage][kGaeEnvs];
    uint8_t ud[kStage][kGaeEnvs];
    uint8_t um[kStage][kGaeEnvs];
};

DEV uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

template <bool VTRACE>
__global__ void __launch_bounds__(kGaeEnvs) gae_tma_kernel(float* __restrict__ rewards, uint8_t* __restrict__ undones,
                                                           const uint8_t* __restrict__ unmasks, const float* __restrict__ values,
                                                           const float* __restrict__ last_value, int H, int N, float gamma,
                                                           float lam, int64_t env_offset, float* __restrict__ adv_out,
                                                           float* __restrict__ rsum_out, double* stat_sums) {
    __shared__ GaeStage stage[2];
    __shared__ uint64_t full[2];
    const int n0 = blockIdx.x * kGaeEnvs, n = n0 + threadIdx.x;
    const int nb = (H + kStage - 1) / kStage;

    auto issue = [&](int k) {  // batch k covers t in [t_lo, t_hi), scanning downwards from H
        const int t_hi = H - k * kStage, t_lo = max(0, t_hi - kStage), rows = t_hi - t_lo;
        GaeStage& st = stage[k & 1];
        uint64_t* bar = &full[k & 1];
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"((uint32_t)(rows * 1280)) : "memory");
        for (int j = 0; j < rows; ++j) {
            const size_t g = (size_t)(t_lo + j) * N + n0;
            bulk_g2s(st.r[j], rewards + g, 512, bar);
            bulk_g2s(st.v[j], values + g, 512, bar);
            bulk_g2s(st.ud[j], undones + g, 128, bar);
            bulk_g2s(st.um[j], unmasks + g, 128, bar);
        }
    };
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        issue(0);
        if (nb > 1) issue(1);
    }
    __syncthreads();

    float vnext = last_value[n];
    float y = 0.0f;
    double s_all = 0.0, s_all2 = 0.0, s_lat = 0.0, s_lat2 = 0.0;
    const bool lat_env = ((env_offset + n) & 3) == 0;
    for (int k = 0; k < nb; ++k) {
        const int t_hi = H - k * kStage, t_lo = max(0, t_hi - kStage);
        const uint32_t parity = (k >> 1) & 1;
        {   // wait for the batch
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 10000;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_addr(&full[k & 1])), "r"(parity) : "memory");
            }
        }
        const GaeStage& st = stage[k & 1];
#pragma unroll 4
        for (int t = t_hi - 1; t >= t_lo; --t) {
            const int j = t - t_lo;
            const size_t i = (size_t)t * N + n;
            StepIn in{st.r[j][threadIdx.x], st.v[j][threadIdx.x], st.ud[j][threadIdx.x] != 0, st.um[j][threadIdx.x] != 0};
            float rf; bool uf;
            const float adv = gae_step<VTRACE>(in, gamma, lam, y, vnext, rf, uf);
            adv_out[i] = adv;
            rsum_out[i] = __fadd_rn(adv, in.v);
            if (!in.unmask) { rewards[i] = rf; undones[i] = 0; }
            s_all += (double)adv;
            s_all2 += (double)adv * (double)adv;
            if (lat_env && (t & 3) == 0) { s_lat += (double)adv; s_lat2 += (double)adv * (double)adv; }
        }
        __syncthreads();  // everyone is done with stage[k & 1]
        if (threadIdx.x == 0 && k + 2 < nb) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads before the async-proxy overwrite
            issue(k + 2);
        }
    }
    __shared__ double red[4][4];
    s_all = warp_sum(s_all); s_lat = warp_sum(s_lat); s_lat2 = warp_sum(s_lat2); s_all2 = warp_sum(s_all2);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[0][w] = s_all; red[1][w] = s_lat; red[2][w] = s_lat2; red[3][w] = s_all2; }
    __syncthreads();
    if (threadIdx.x < 4) {
        const double tot = red[threadIdx.x][0] + red[threadIdx.x][1] + red[threadIdx.x][2] + red[threadIdx.x][3];
        atomicAdd(stat_sums + threadIdx.x, tot);
    }
}

__global__ void adv_stats_kernel(const double* stat_sums, double count_all, double count_lat, float* stats_out) {
    // mean over everything; unbiased std over the [::4, ::4] lattice   (reference AgentPPO.py:149)
    double mean = stat_sums[0] / count_all;
    // count_lat == 0 selects the unbiased std over EVERYTHING (helloworld_PPO_single_file.py:296, adv.std(dim=0))
    const bool full = count_lat == 0.0;
    const double cnt = full ? count_all : count_lat;
    const double m = (full ? stat_sums[0] : stat_sums[1]) / cnt, sq = full ? stat_sums[3] : stat_sums[2];
    double var = (sq - cnt * m * m) / (cnt - 1.0);
    float sd = (float)sqrt(var > 0.0 ? var : 0.0);
    if (cnt < 2.0) sd = nanf("");  // torch .std() of one element is NaN
    stats_out[0] = (float)mean;
    stats_out[1] = sd;
    stats_out[2] = 1.0f / (sd + 1e-5f);
    stats_out[3] = 0.0f;
}

__global__ void normalize_adv_kernel(float* __restrict__ adv, int64_t count, const float* __restrict__ stats) {
    const float mean = stats[0], sd = stats[1];
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < count) {
        float4 v = *reinterpret_cast<float4*>(adv + i);
        v.x = (v.x - mean) / (sd + 1e-5f); v.y = (v.y - mean) / (sd + 1e-5f);
        v.z = (v.z - mean) / (sd + 1e-5f); v.w = (v.w - mean) / (sd + 1e-5f);
        *reinterpret_cast<float4*>(adv + i) = v;
    } else {
        for (; i < count; ++i) adv[i] = (adv[i] - mean) / (sd + 1e-5f);
    }
}

}  // namespace

extern "C" {

int b200rl_gae(float* rewards, uint8_t* undones, const uint8_t* unmasks, const float* values, const float* last_value,
               int32_t horizon_len, int32_t num_envs, float gamma, float lambda_gae, int32_t if_use_v_trace,
               int64_t env_offset, float* advantages, float* reward_sums, double* stat_sums, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    B200RL_REQUIRE(rewards && undones && unmasks && values && last_value && advantages && reward_sums && stat_sums,
                   "gae: NULL buffer");
    B200RL_REQUIRE(horizon_len >= 1 && num_envs >= 1, "gae: horizon_len=%d num_envs=%d", horizon_len, num_envs);
    B200RL_CHECK_CUDA(cudaMemsetAsync(stat_sums, 0, 4 * sizeof(double), stream));
    // env axis fills the GPU when N is large; otherwise cut the time axis so that >= ~2 warps/SM-quadrant exist
    int bx = num_envs >= 128 ? 128 : 32;
    int chunks = 1;
    if (num_envs < 148 * 256) {
        int64_t want = (148LL * 512) / (num_envs > 0 ? num_envs : 1);  // threads we would like / env
        while (chunks < 32 && chunks * 2 <= want && horizon_len / (chunks * 2) >= 8 && bx * chunks * 2 <= 512) chunks *= 2;
    }
    if (chunks == 1 && (num_envs % kGaeEnvs) == 0 && num_envs >= 148 * 64 && horizon_len >= 2 * kStage) {
        // TMA-staged sequential scan (bit-exact, bandwidth-oriented)
        const unsigned g = num_envs / kGaeEnvs;
        if (if_use_v_trace)
            gae_tma_kernel<true><<<g, kGaeEnvs, 0, stream>>>(rewards, undones, unmasks, values, last_value, horizon_len, num_envs,
                                                            gamma, lambda_gae, env_offset, advantages, reward_sums, stat_sums);
        else
            gae_tma_kernel<false><<<g, kGaeEnvs, 0, stream>>>(rewards, undones, unmasks, values, last_value, horizon_len, num_envs,
                                                             gamma, lambda_gae, env_offset, advantages, reward_sums, stat_sums);
        B200RL_CHECK_CUDA(cudaGetLastError());
        B200RL_COUNT_LAUNCH(1);
        return 0;
    }
    dim3 block(bx, chunks), grid((num_envs + bx - 1) / bx);
    size_t smem = chunks > 1 ? (size_t)chunks * bx * sizeof(float2) : 0;
    if (if_use_v_trace)
        gae_kernel<true><<<grid, block, smem, stream>>>(rewards, undones, unmasks, values, last_value, horizon_len, num_envs,
                                                       gamma, lambda_gae, env_offset, advantages, reward_sums, stat_sums);
    else
        gae_kernel<false><<<grid, block, smem, stream>>>(rewards, undones, unmasks, values, last_value, horizon_len, num_envs,
                                                        gamma, lambda_gae, env_offset, advantages, reward_sums, stat_sums);
    B200RL_CHECK_CUDA(cudaGetLastError());
    B200RL_COUNT_LAUNCH(1);
    return 0;
}

int b200rl_adv_stats(const double* stat_sums, int64_t count_all, int64_t count_lattice, float* stats_out, void* stream) {
    B200RL_REQUIRE(stat_sums && stats_out, "adv_stats: NULL buffer");
    adv_stats_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(stat_sums, (double)count_all, (double)count_lattice, stats_out);
    B200RL_CHECK_CUDA(cudaGetLastError());
    B200RL_COUNT_LAUNCH(1);
    return 0;
}

int b200rl_normalize_adv(float* advantages, int64_t count, const float* stats, void* stream) {
    B200RL_REQUIRE(advantages && stats, "normalize_adv: NULL buffer");
    if (count <= 0) return 0;
    int64_t threads = (count + 3) / 4;
    normalize_adv_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(advantages, count, stats);
    B200RL_CHECK_CUDA(cudaGetLastError());
    B200RL_COUNT_LAUNCH(1);
    return 0;
}

}  // extern "C"
