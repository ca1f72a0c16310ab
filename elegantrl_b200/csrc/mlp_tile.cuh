// Generic (runtime-dims) MLP tile routines on CUDA cores, shared by forward.cu and update.cu.
//
// A CTA owns a tile of TB samples.  Activations live in shared memory feature-major, X[k][b], so that the 32
// lanes of a warp read consecutive samples of one feature row (conflict-free LDS.128) in the forward and
// data-gradient GEMMs.  The weight-gradient GEMM (dW[j][k] = sum_b dZ[j][b] * X[k][b]) instead has lanes on
// different feature rows at the same sample chunk; to keep that conflict-free too, the 16-byte sample chunk c
// of row r is stored at chunk (c ^ (r/4)) -- an XOR swizzle keyed by the row's 4-row group.
#pragma once
#include "common.cuh"

template <int TB>
struct SmemTile {
    static constexpr int kChunks = TB / 4;
    // float offset of the 4-sample chunk c of feature row r
    static DEV int chunk(int r, int c) { return r * TB + (((c ^ (r >> 2)) & (kChunks - 1)) << 2); }
    static DEV int elem(int r, int b) { return chunk(r, b >> 2) + (b & 3); }
};

DEV float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
DEV void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Load TB rows of x [rows, S] (row-major) starting at row0 into Xs[S][TB], applying state_norm
// (reference AgentPPO.py:360-361): (s - avg) / (std + 1e-4).  Rows past the end are zero.
template <int TB, int NT>
DEV void load_state_tile(const b200rl_net& net, const float* __restrict__ x, int64_t rows, int64_t row0, float* Xs) {
    const int S = net.dims[0];
    for (int idx = threadIdx.x; idx < TB * S; idx += NT) {
        int b = idx / S, k = idx - b * S;
        int64_t row = row0 + b;
        float v = 0.0f;
        if (row < rows) {
            v = x[row * S + k];
            if (net.state_avg) v = (v - net.state_avg[k]) / (net.state_std[k] + 1e-4f);
        }
        Xs[SmemTile<TB>::elem(k, b)] = v;
    }
}

// ---- where a Linear's parameters are read from
//  W_LDG   global memory, read-only (non-coherent) path: kernels that only read the parameters
//  W_LDCG  global memory, L2-coherent loads: parameters rewritten by other CTAs during the kernel
//  W_SMEM  a copy staged in shared memory by stage_weight(): rows of K floats whose 16-byte chunks are XOR-swizzled
//          by the row's 4-row group (when K % 32 == 0), so that both "4 consecutive rows, same k" (forward) and
//          "same row, 4 consecutive k-chunks per lane group" (data gradient) are bank-conflict free
enum WeightMode { W_LDG = 0, W_LDCG = 1, W_SMEM = 2 };

template <int WM>
struct WeightView {
    const float* base;
    int K;
    bool vec;   // float4 access legal (K % 4 == 0 and aligned)
    bool swz;   // W_SMEM only: swizzled rows
    DEV WeightView(const float* b, int k) : base(b), K(k) {
        vec = ((k & 3) == 0) && ((reinterpret_cast<uintptr_t>(b) & 15) == 0);
        swz = (WM == W_SMEM) && ((k & 31) == 0);
    }
    DEV int off4(int j, int k) const {  // float offset of the 16-byte chunk holding (j, k..k+3), k % 4 == 0
        return swz ? j * K + ((((k >> 2) ^ (j >> 2)) & 7) << 2) + ((k >> 5) << 5) : j * K + k;
    }
    DEV float4 ld4(int j, int k) const {
        const float* p = base + off4(j, k);
        if (WM == W_LDG) return __ldg(reinterpret_cast<const float4*>(p));
        if (WM == W_LDCG) return __ldcg(reinterpret_cast<const float4*>(p));
        return *reinterpret_cast<const float4*>(p);
    }
    DEV float ld1(int j, int k) const {
        const float* p = base + (swz ? off4(j, k & ~3) + (k & 3) : j * K + k);
        if (WM == W_LDG) return __ldg(p);
        if (WM == W_LDCG) return __ldcg(p);
        return *p;
    }
};
template <int WM>
DEV float ld_param(const float* p) { return WM == W_LDG ? __ldg(p) : (WM == W_LDCG ? __ldcg(p) : *p); }

// copy W [J][K] (global, coherent loads) into the shared-memory layout WeightView<W_SMEM> reads
// (executed by the `nthreads` threads tid = 0 .. nthreads-1 of the caller's choice)
DEV void stage_weight(const float* __restrict__ W, int J, int K, float* dst, int tid, int nthreads) {
    const bool swz = (K & 31) == 0;
    if (((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0)) {
        const int chunks = K >> 2;
        for (int i = tid; i < J * chunks; i += nthreads) {
            const int j = i / chunks, c = i - j * chunks;
            const float4 v = __ldcg(reinterpret_cast<const float4*>(W) + i);
            const int cs = swz ? ((c & ~7) | ((c ^ (j >> 2)) & 7)) : c;
            *reinterpret_cast<float4*>(dst + j * K + (cs << 2)) = v;
        }
    } else {
        for (int i = tid; i < J * K; i += nthreads) dst[i] = __ldcg(W + i);
    }
}

// activation and its derivative from ONE erf evaluation (GELU) -- used by the update's forward pass
DEV void act_and_grad(float z, int act, float& h, float& g) {
    if (act == B200RL_ACT_GELU) {
        const float cdf = 0.5f * (1.0f + erff(z * kSqrtHalf));
        h = z * cdf;  // == z * 0.5 * (1 + erf(z / sqrt 2)) up to one rounding
        g = cdf + z * (expf(-0.5f * z * z) * kInvSqrt2Pi);
    } else {
        h = fmaxf(z, 0.0f);
        g = z > 0.0f ? 1.0f : 0.0f;
    }
}

// Ys[j][b] = act(sum_k W[j][k] * Xs[k][b] + bias[j]); optionally Gs[j][b] = act'(pre-activation).
// Thread tile: 4 samples x JT outputs (JT = 4, or 2 for narrow layers so that every thread of the CTA has work).
template <int TB, int NT, int WM, int JT>
DEV void linear_forward_tile(const float* Wp, const float* bias, int K, int J, const float* Xs, float* Ys, float* Gs, int act,
                             bool apply_act) {
    constexpr int NSG = TB / 4, NOL = NT / NSG;
    using T = SmemTile<TB>;
    const int sg = threadIdx.x % NSG, ol = threadIdx.x / NSG;
    const WeightView<WM> W(Wp, K);
    for (int j0 = ol * JT; j0 < J; j0 += NOL * JT) {
        float acc[JT][4];
#pragma unroll
        for (int jj = 0; jj < JT; ++jj) {
            float bv = (j0 + jj < J) ? ld_param<WM>(bias + j0 + jj) : 0.0f;
#pragma unroll
            for (int s = 0; s < 4; ++s) acc[jj][s] = bv;
        }
        if (W.vec) {
#pragma unroll 2
            for (int k = 0; k < K; k += 4) {
                float4 xv[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) xv[kk] = ld4(Xs + T::chunk(k + kk, sg));
#pragma unroll
                for (int jj = 0; jj < JT; ++jj) {
                    if (j0 + jj < J) {
                        float4 w = W.ld4(j0 + jj, k);
                        acc[jj][0] = fmaf(w.x, xv[0].x, acc[jj][0]); acc[jj][1] = fmaf(w.x, xv[0].y, acc[jj][1]);
                        acc[jj][2] = fmaf(w.x, xv[0].z, acc[jj][2]); acc[jj][3] = fmaf(w.x, xv[0].w, acc[jj][3]);
                        acc[jj][0] = fmaf(w.y, xv[1].x, acc[jj][0]); acc[jj][1] = fmaf(w.y, xv[1].y, acc[jj][1]);
                        acc[jj][2] = fmaf(w.y, xv[1].z, acc[jj][2]); acc[jj][3] = fmaf(w.y, xv[1].w, acc[jj][3]);
                        acc[jj][0] = fmaf(w.z, xv[2].x, acc[jj][0]); acc[jj][1] = fmaf(w.z, xv[2].y, acc[jj][1]);
                        acc[jj][2] = fmaf(w.z, xv[2].z, acc[jj][2]); acc[jj][3] = fmaf(w.z, xv[2].w, acc[jj][3]);
                        acc[jj][0] = fmaf(w.w, xv[3].x, acc[jj][0]); acc[jj][1] = fmaf(w.w, xv[3].y, acc[jj][1]);
                        acc[jj][2] = fmaf(w.w, xv[3].z, acc[jj][2]); acc[jj][3] = fmaf(w.w, xv[3].w, acc[jj][3]);
                    }
                }
            }
        } else {
            for (int k = 0; k < K; ++k) {
                float4 xv = ld4(Xs + T::chunk(k, sg));
#pragma unroll
                for (int jj = 0; jj < JT; ++jj) {
                    if (j0 + jj < J) {
                        float w = W.ld1(j0 + jj, k);
                        acc[jj][0] = fmaf(w, xv.x, acc[jj][0]); acc[jj][1] = fmaf(w, xv.y, acc[jj][1]);
                        acc[jj][2] = fmaf(w, xv.z, acc[jj][2]); acc[jj][3] = fmaf(w, xv.w, acc[jj][3]);
                    }
                }
            }
        }
#pragma unroll
        for (int jj = 0; jj < JT; ++jj) {
            if (j0 + jj < J) {
                float4 y = make_float4(acc[jj][0], acc[jj][1], acc[jj][2], acc[jj][3]);
                if (apply_act) {
                    if (Gs) {
                        float4 g;
                        act_and_grad(y.x, act, y.x, g.x); act_and_grad(y.y, act, y.y, g.y);
                        act_and_grad(y.z, act, y.z, g.z); act_and_grad(y.w, act, y.w, g.w);
                        st4(Gs + T::chunk(j0 + jj, sg), g);
                    } else {
                        y = make_float4(act_fn_rt(y.x, act), act_fn_rt(y.y, act), act_fn_rt(y.z, act), act_fn_rt(y.w, act));
                    }
                }
                st4(Ys + T::chunk(j0 + jj, sg), y);
            }
        }
    }
}

template <int TB, int NT, int WM = W_LDG>
DEV void linear_forward(const float* Wp, const float* bias, int K, int J, const float* Xs, float* Ys, float* Gs, int act,
                        bool apply_act) {
    constexpr int NOL = NT / (TB / 4);
    if (J <= NOL * 2) linear_forward_tile<TB, NT, WM, 2>(Wp, bias, K, J, Xs, Ys, Gs, act, apply_act);
    else linear_forward_tile<TB, NT, WM, 4>(Wp, bias, K, J, Xs, Ys, Gs, act, apply_act);
}
