// Backward-pass tile routines of the generic (runtime-dims) MLP kernels: weight gradient and data gradient of one Linear
// over a tile of TB samples held feature-major in shared memory (layout: mlp_tile.cuh).  Shared by update.cu (PPO) and
// sac.cu (SAC).
#pragma once
#include "mlp_tile.cuh"

// dW[j][k] += sum_b dZ[j][b] * X[k][b];  db[j] += sum_b dZ[j][b]      (RED.ADD into the flat buffer)
// `atomic` = false: gW / gb point at this CTA's private shared-memory accumulator (every element is owned by exactly one
// thread, so plain adds are race-free); the CTA flushes it with one RED.ADD per element after its last sample tile.
template <int TB, int NT>
DEV void weight_grad(const float* dZ, const float* X, int J, int K, float* gW, float* gb, bool atomic) {
    const int JT = (J + 3) >> 2, KT = (K + 3) >> 2;
    for (int tile = threadIdx.x; tile < JT * KT; tile += NT) {
        const int jt = tile / KT, kt = tile - jt * KT;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
#pragma unroll 2
        for (int c = 0; c < SmemTile<TB>::kChunks; ++c) {
            float4 dz[4], xv[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
                dz[jj] = (4 * jt + jj < J) ? ld4(dZ + SmemTile<TB>::chunk(4 * jt + jj, c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                xv[kk] = (4 * kt + kk < K) ? ld4(X + SmemTile<TB>::chunk(4 * kt + kk, c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    acc[jj][kk] = fmaf(dz[jj].x, xv[kk].x, acc[jj][kk]);
                    acc[jj][kk] = fmaf(dz[jj].y, xv[kk].y, acc[jj][kk]);
                    acc[jj][kk] = fmaf(dz[jj].z, xv[kk].z, acc[jj][kk]);
                    acc[jj][kk] = fmaf(dz[jj].w, xv[kk].w, acc[jj][kk]);
                }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                if (4 * jt + jj < J && 4 * kt + kk < K) {
                    float* dst = gW + (size_t)(4 * jt + jj) * K + 4 * kt + kk;
                    if (atomic) atomicAdd(dst, acc[jj][kk]);
                    else *dst += acc[jj][kk];
                }
    }
    for (int j = threadIdx.x; j < J; j += NT) {
        float s = 0.0f;
        for (int c = 0; c < SmemTile<TB>::kChunks; ++c) {
            float4 v = ld4(dZ + SmemTile<TB>::chunk(j, c));
            s += (v.x + v.y) + (v.z + v.w);
        }
        if (atomic) atomicAdd(gb + j, s);
        else gb[j] += s;
    }
}

// dZprev[k][b] = (sum_j W[j][k] * dZ[j][b]) * G[k][b]     thread tile: 4 samples x KT inputs (KT = 4, or 2 when K is
// small enough that 4-wide tiles would leave half of the CTA idle)
template <int TB, int NT, int WM, int KT>
DEV void data_grad_tile(const float* Wp, const float* dZ, const float* G, float* dZprev, int J, int K, bool accumulate) {
    constexpr int NSG = TB / 4, NOL = NT / NSG;
    const int sg = threadIdx.x % NSG, ol = threadIdx.x / NSG;
    const WeightView<WM> W(Wp, K);
    const bool vec = W.vec;
    for (int k0 = ol * KT; k0 < K; k0 += NOL * KT) {
        float acc[KT][4];
#pragma unroll
        for (int a = 0; a < KT; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
#pragma unroll 4
        for (int j = 0; j < J; ++j) {
            float4 dz = ld4(dZ + SmemTile<TB>::chunk(j, sg));
            float w[KT];
            if (vec) {
                float4 w4 = W.ld4(j, k0 & ~3);
                if (KT == 4) { w[0] = w4.x; w[1] = w4.y; w[KT - 2] = w4.z; w[KT - 1] = w4.w; }
                else if (k0 & 2) { w[0] = w4.z; w[1] = w4.w; }
                else { w[0] = w4.x; w[1] = w4.y; }
            } else {
#pragma unroll
                for (int kk = 0; kk < KT; ++kk) w[kk] = (k0 + kk < K) ? W.ld1(j, k0 + kk) : 0.0f;
            }
#pragma unroll
            for (int kk = 0; kk < KT; ++kk) {
                acc[kk][0] = fmaf(w[kk], dz.x, acc[kk][0]); acc[kk][1] = fmaf(w[kk], dz.y, acc[kk][1]);
                acc[kk][2] = fmaf(w[kk], dz.z, acc[kk][2]); acc[kk][3] = fmaf(w[kk], dz.w, acc[kk][3]);
            }
        }
#pragma unroll
        for (int kk = 0; kk < KT; ++kk) {
            if (k0 + kk < K) {
                // G == nullptr: the layer below has no activation (a raw Linear output); accumulate: several heads feed it
                const float4 g = G ? ld4(G + SmemTile<TB>::chunk(k0 + kk, sg)) : make_float4(1.f, 1.f, 1.f, 1.f);
                float4 o = make_float4(acc[kk][0] * g.x, acc[kk][1] * g.y, acc[kk][2] * g.z, acc[kk][3] * g.w);
                float* dst = dZprev + SmemTile<TB>::chunk(k0 + kk, sg);
                if (accumulate) { const float4 p = ld4(dst); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                st4(dst, o);
            }
        }
    }
}
template <int TB, int NT, int WM>
DEV void data_grad(const float* Wp, const float* dZ, const float* G, float* dZprev, int J, int K, bool accumulate = false) {
    constexpr int NOL = NT / (TB / 4);
    if (K <= NOL * 2) data_grad_tile<TB, NT, WM, 2>(Wp, dZ, G, dZprev, J, K, accumulate);
    else data_grad_tile<TB, NT, WM, 4>(Wp, dZ, G, dZprev, J, K, accumulate);
}

