// Building blocks of the tcgen05 PPO update (update_tc.cu) and its unit test (tools/tc_train_test.cu): forward AND backward
// of a  S -> 64 -> 64 -> OUT  GELU MLP on a tile of 128 sampled transitions, thread = sample = TMEM lane.
// Reference arithmetic: AgentPPO.update_objectives (elegantrl/agents/AgentPPO.py:173-205) through build_mlp's
// Linear-GELU-Linear-GELU-Linear (elegantrl/agents/AgentBase.py:345-365); fp32 parity via 3xTF32 (hi/lo planes of both
// operands, three UMMAs per K step: hi*hi + lo*hi + hi*lo).
//
// Operand images (no swizzle; tc05::make_smem_desc_ex):
//   * "K-major image" of a weight W [R][K] exactly as nn.Linear stores it: offset(r, k) = (r/8)*(K/4)*128 + (k/4)*128 + (r%8)*16
//     + (k%4)*4.  Forward reads it K-major (B = W, D = A W^T); the backward data gradient reads THE SAME bytes as the MN-major
//     operand W^T (D = dZ W) by swapping the two strides of the descriptor -- no transposed copy.
//   * "row-written image" of a per-sample matrix Q [128 samples][C]: sample b stores its own row with 128-bit stores; the
//     tensor core reads it as an MN-major operand with MN = column, K = sample -- this is how the weight gradients
//     dW = dZ^T X (a contraction over the SAMPLE axis) reach the tensor core without a transposition through shared memory.
//     MN-major kind::tf32 operands exist in ONE layout only, SWIZZLE_128B_BASE32B (cutlass sm100_common.inl: "for mn-major
//     tf32 operands, SW128_32B is the only available smem layout"; measured: with any other layout type the UMMA writes
//     nothing, profiles/r02_v4_mn_major_probe.log).  The layout was decoded on the B200 by filling the operand region with
//     its own word index against a one-hot second operand (tools/tc_mn_probe3.cu, profiles/r02_v4_mn_major_decode.log):
//     atom = 32 columns (128 B) x 4 samples at 128 B, and the 32-BYTE chunk index of a row is XORed with the row index:
//       offset(b, c) = (c/32)*LBO + (b/4)*SBO + (b%4)*128 + (((c%32)*4) ^ ((b%4) << 5))
//     (Swizzle<2,5,2> on byte addresses: bits [5,7) ^= bits [7,9)); descriptor LBO = stride between 32-column groups, SBO =
//     stride between 4-sample groups, for A and for B.  A sample's 16-byte stores simply go to unit u ^ (2 * (b % 4)); the
//     four rows of an atom thereby hit four different bank groups.  LBO = 128 samples / 4 * 512 = 16 KB.
//   * the backward image of W2 (for dH1 = dZ2 W2: B [n = input][K = output]) is the same layout with MN = input feature,
//     K = output feature: offset(j, k) = (k/32)*8192 + (j/4)*512 + (j%4)*128 + (k%32)*4, swizzled.
#pragma once
#include "common.cuh"
#include "tc05.cuh"

namespace tctrain {

constexpr int kHid = 64, kTile = 128;
constexpr int kWPlaneBytes = kHid * kHid * 4;        // one tf32 plane of a 64 x 64 weight, K-major image: 16 KB
constexpr uint32_t kMnSBO = 512;                      // bytes between 4-deep K groups of an MN-major (SW128_32B) image
constexpr uint32_t kRowLBO = kTile / 4 * kMnSBO;      // bytes between 32-column groups of a row-written image: 16 KB
constexpr uint32_t kWbLBO = kHid / 4 * kMnSBO;        // ... of the backward weight image (K = 64 output features): 8 KB
constexpr int kGroupPlaneBytes = (int)kRowLBO;        // one 32-column group of a row-written image
constexpr int kGAPlaneBytes = 2 * kGroupPlaneBytes;   // row-written image [128][64]: 32 KB

// ---- exact-erf GELU and its derivative from ONE evaluation of q = Phi(-|x|) (tools/fit_gelu.py; max abs error 5.8e-7 / 1e-6)
//   GELU(x) = max(x, 0) - |x| q,   GELU'(x) = Phi(x) + x phi(x),  Phi(x) = x >= 0 ? 1 - q : q,  phi(x) = exp(-x^2/2) / sqrt(2 pi)
DEV void gelu_and_grad(float x, float& g, float& dg) {
    constexpr float L = 6.2225397f;
    const float zn = fmaxf(-fabsf(x), -L);
    float p = fmaf(1.775934289e-05f, zn, 6.477866232e-04f);
    p = fmaf(p, zn, 7.724114180e-03f);
    p = fmaf(p, zn, 5.292681266e-02f);
    p = fmaf(p, zn, -4.590827042e-01f);
    p = fmaf(p, zn, 1.151116861e+00f);
    const float q = tc05::ex2_approx(fmaf(p, zn, -1.0f));
    g = fmaf(zn, q, fmaxf(x, 0.0f));
    const float pdf = tc05::ex2_approx(zn * zn * -0.72134752044f) * kInvSqrt2Pi;   // exp(-x^2/2); the clamp only matters where it is 0
    dg = fmaf(x, pdf, x >= 0.0f ? 1.0f - q : q);
}

DEV float gelu_only(float x) {
    constexpr float L = 6.2225397f;
    const float zn = fmaxf(-fabsf(x), -L);
    float p = fmaf(1.775934289e-05f, zn, 6.477866232e-04f);
    p = fmaf(p, zn, 7.724114180e-03f);
    p = fmaf(p, zn, 5.292681266e-02f);
    p = fmaf(p, zn, -4.590827042e-01f);
    p = fmaf(p, zn, 1.151116861e+00f);
    return fmaf(zn, tc05::ex2_approx(fmaf(p, zn, -1.0f)), fmaxf(x, 0.0f));
}
// The same two functions on PAIRS with the packed fp32 instructions of sm_100 (FFMA2 / FMUL2: one issue slot for two lanes'
// worth of fma; every operation rounds exactly like its scalar counterpart above, so results are bit-identical): the GELUs
// are most of this kernel's CUDA-core work AND of its code size (it is executed by a handful of warps, instruction fetch is
// its second largest stall).
DEV float2 splat2(float v) { return make_float2(v, v); }
DEV float2 gelu_only2(float2 x) {
    constexpr float L = 6.2225397f;
    const float2 zn = make_float2(fmaxf(-fabsf(x.x), -L), fmaxf(-fabsf(x.y), -L));
    float2 p = __ffma2_rn(splat2(1.775934289e-05f), zn, splat2(6.477866232e-04f));
    p = __ffma2_rn(p, zn, splat2(7.724114180e-03f));
    p = __ffma2_rn(p, zn, splat2(5.292681266e-02f));
    p = __ffma2_rn(p, zn, splat2(-4.590827042e-01f));
    p = __ffma2_rn(p, zn, splat2(1.151116861e+00f));
    const float2 t = __ffma2_rn(p, zn, splat2(-1.0f));
    const float2 q = make_float2(tc05::ex2_approx(t.x), tc05::ex2_approx(t.y));
    return __ffma2_rn(zn, q, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}
DEV void gelu_and_grad2(float2 x, float2& g, float2& dg) {
    constexpr float L = 6.2225397f;
    const float2 zn = make_float2(fmaxf(-fabsf(x.x), -L), fmaxf(-fabsf(x.y), -L));
    float2 p = __ffma2_rn(splat2(1.775934289e-05f), zn, splat2(6.477866232e-04f));
    p = __ffma2_rn(p, zn, splat2(7.724114180e-03f));
    p = __ffma2_rn(p, zn, splat2(5.292681266e-02f));
    p = __ffma2_rn(p, zn, splat2(-4.590827042e-01f));
    p = __ffma2_rn(p, zn, splat2(1.151116861e+00f));
    const float2 t = __ffma2_rn(p, zn, splat2(-1.0f));
    const float2 q = make_float2(tc05::ex2_approx(t.x), tc05::ex2_approx(t.y));
    g = __ffma2_rn(zn, q, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
    const float2 e = __fmul2_rn(__fmul2_rn(zn, zn), splat2(-0.72134752044f));
    const float2 pdf = __fmul2_rn(make_float2(tc05::ex2_approx(e.x), tc05::ex2_approx(e.y)), splat2(kInvSqrt2Pi));
    dg = __ffma2_rn(x, pdf, make_float2(x.x >= 0.0f ? 1.0f - q.x : q.x, x.y >= 0.0f ? 1.0f - q.y : q.y));
}
// 16 values in place
DEV void gelu_only16(float (&z)[16]) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        const float2 g = gelu_only2(make_float2(z[j], z[j + 1]));
        z[j] = g.x; z[j + 1] = g.y;
    }
}

// Sum over the 32 lanes of a warp of 16 per-lane values with 16 shuffles (recursive halving: every step exchanges half of
// the values a lane still holds): afterwards lane l holds the total of value index (l >> 1) & 15 (lanes l and l ^ 1 the same).
DEV float warp_reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int s = 8, mask = 16; s >= 1; s >>= 1, mask >>= 1) {
        const bool upper = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = upper ? v[i] : v[i + s];
            const float keep = upper ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// ---- staging
// nn.Linear weight [64][64] (read L2-coherently: Adam rewrites it between minibatches) -> hi / lo K-major images
DEV void stage_w_planes(const float* W, unsigned char* hi, unsigned char* lo, int tid, int nthreads) {
    for (int i = tid; i < kHid * kHid; i += nthreads) {
        const int n = i >> 6, k = i & 63;
        const float w = __ldcg(W + i), h = tc05::tf32_hi(w);
        const uint32_t off = tc05::operand_offset(n, k, kHid);
        *reinterpret_cast<float*>(hi + off) = h;
        *reinterpret_cast<float*>(lo + off) = w - h;
    }
}
DEV uint32_t mn_swizzle(uint32_t byte_off) { return byte_off ^ (((byte_off >> 7) & 3u) << 5); }
// ... and its backward image (MN = input feature k, K = output feature j), hi / lo planes
DEV void stage_w_planes_backward(const float* W, unsigned char* hi, unsigned char* lo, int tid, int nthreads) {
    for (int i = tid; i < kHid * kHid; i += nthreads) {
        const int j = i >> 6, k = i & 63;
        const float w = __ldcg(W + i), h = tc05::tf32_hi(w);
        const uint32_t off = mn_swizzle((uint32_t)(k >> 5) * kWbLBO + (uint32_t)(j >> 2) * kMnSBO + (uint32_t)(j & 3) * 128 + (uint32_t)(k & 31) * 4);
        *reinterpret_cast<float*>(hi + off) = h;
        *reinterpret_cast<float*>(lo + off) = w - h;
    }
}
// 16 consecutive columns of this thread's row -> hi / lo planes in tensor memory (two tcgen05.st; caller waits)
DEV void store_hi_lo_tmem(uint32_t taddr_hi, uint32_t taddr_lo, const float (&v)[16]) {
    uint32_t h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float hi = tc05::tf32_hi(v[i]);
        h[i] = __float_as_uint(hi);
        l[i] = __float_as_uint(v[i] - hi);
    }
    tc05::tmem_st_32x32b_x16(taddr_hi, h);
    tc05::tmem_st_32x32b_x16(taddr_lo, l);
}
DEV void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// NV consecutive columns [col0, col0 + NV) of sample b's row -> hi / lo row-written images (col0 % 4 == 0; the NV columns stay
// inside one 32-column group).  16-byte unit u of the row is stored at unit u ^ (2 * (b % 4)) (the layout's 32-byte swizzle).
template <int NV>
DEV void store_hi_lo_rows_n(unsigned char* hi_plane, unsigned char* lo_plane, int b, int col0, const float (&v)[NV]) {
    const uint32_t row = (uint32_t)(col0 >> 5) * kRowLBO + (uint32_t)(b >> 2) * kMnSBO + (uint32_t)(b & 3) * 128;
    const uint32_t u0 = (uint32_t)(col0 & 31) >> 2, x = (uint32_t)(b & 3) << 1;
    const uint32_t hi0 = tc05::smem_u32(hi_plane) + row, lo0 = tc05::smem_u32(lo_plane) + row;
#pragma unroll
    for (int q = 0; q < NV / 4; ++q) {
        const uint32_t unit = ((u0 + q) ^ x) * 16;
        const float h0 = tc05::tf32_hi(v[4 * q]), h1 = tc05::tf32_hi(v[4 * q + 1]), h2 = tc05::tf32_hi(v[4 * q + 2]), h3 = tc05::tf32_hi(v[4 * q + 3]);
        st_shared_v4(hi0 + unit, h0, h1, h2, h3);
        st_shared_v4(lo0 + unit, v[4 * q] - h0, v[4 * q + 1] - h1, v[4 * q + 2] - h2, v[4 * q + 3] - h3);
    }
}
DEV void store_hi_lo_rows(unsigned char* hi_plane, unsigned char* lo_plane, int b, int col0, const float (&v)[16]) {
    store_hi_lo_rows_n<16>(hi_plane, lo_plane, b, col0, v);
}
DEV void store_hi_lo_rows8(unsigned char* hi_plane, unsigned char* lo_plane, int b, int col0, const float (&v)[8]) {
    store_hi_lo_rows_n<8>(hi_plane, lo_plane, b, col0, v);
}

// MN-major tf32 descriptor: layout type SWIZZLE_128B_BASE32B; LBO = stride between 32-element MN groups, SBO = stride
// between 4-deep K groups (cute make_umma_desc<Major::MN>, swizzled branch)
DEV uint64_t make_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    return tc05::make_smem_desc_ex(smem_addr, lbo_bytes, kMnSBO) | ((uint64_t)1 << 61);
}

// ---- issuer side (ONE thread)
// The issuing thread is ONE thread of a warp that has its scheduler to itself: every instruction it spends per UMMA is exposed
// latency (measured: ~110 cycles per UMMA with descriptors rebuilt in a loop, 170 UMMAs per minibatch).  The loops below are
// fully unrolled and the descriptors advance by adding the K step to the start-address field (units of 16 bytes).
// D[128 x 64] (+)= A * W^T: A = hi / lo planes in tensor memory (64 columns each, lane = sample), W = hi / lo K-major
// images of a 64 x 64 weight.  24 UMMAs 128 x 64 x 8.
DEV void issue_linear_ts(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t tmem_a_lo, uint32_t w_hi, uint32_t w_lo, bool accumulate) {
    constexpr uint32_t idesc = tc05::make_idesc_tf32_ex(kTile, kHid, false, false);
    constexpr uint32_t sbo_k = (kHid / 4) * 128;   // 2048: stride between 8-row groups of the K-major image
    const uint64_t b_hi0 = tc05::make_smem_desc_ex(w_hi, 128, sbo_k), b_lo0 = tc05::make_smem_desc_ex(w_lo, 128, sbo_k);
#pragma unroll
    for (int ks = 0; ks < kHid / 8; ++ks) {
        const uint64_t o = (uint64_t)(ks * (256 >> 4));   // K = input feature: two 16-byte chunks per step
        tc05::mma_tf32_ts(tmem_d, tmem_a_hi + 8 * ks, b_hi0 + o, idesc, accumulate || ks > 0);
        tc05::mma_tf32_ts(tmem_d, tmem_a_lo + 8 * ks, b_hi0 + o, idesc, true);
        tc05::mma_tf32_ts(tmem_d, tmem_a_hi + 8 * ks, b_lo0 + o, idesc, true);
    }
}
// D[128 x 64] = A * W (the data gradient): the B operand is the backward image of W (MN-major: MN = input feature)
DEV void issue_linear_ts_backward(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t tmem_a_lo, uint32_t wb_hi, uint32_t wb_lo) {
    constexpr uint32_t idesc = tc05::make_idesc_tf32_ex(kTile, kHid, false, true);
    const uint64_t b_hi0 = make_mn_desc(wb_hi, kWbLBO), b_lo0 = make_mn_desc(wb_lo, kWbLBO);
#pragma unroll
    for (int ks = 0; ks < kHid / 8; ++ks) {      // K = output feature: two 4-deep K groups per step
        const uint64_t o = (uint64_t)(ks * ((2 * kMnSBO) >> 4));
        tc05::mma_tf32_ts(tmem_d, tmem_a_hi + 8 * ks, b_hi0 + o, idesc, ks > 0);
        tc05::mma_tf32_ts(tmem_d, tmem_a_lo + 8 * ks, b_hi0 + o, idesc, true);
        tc05::mma_tf32_ts(tmem_d, tmem_a_hi + 8 * ks, b_lo0 + o, idesc, true);
    }
}
// D[64 x N] = G^T * Q over the 128 samples: G [128][64] and Q [128][N <= 32] as row-written hi / lo images (lo plane = hi
// plane address + plane bytes).  48 UMMAs 64 x N x 8; accumulator rows at TMEM lanes (m % 16) + 32 * (m / 16).  `accumulate`:
// add to what the accumulator holds (a CTA walking several tiles of a minibatch).
DEV void issue_weight_grad(uint32_t tmem_d, uint32_t ga, uint32_t ga_plane_bytes, uint32_t gb, uint32_t gb_plane_bytes, int N,
                           bool accumulate = false) {
    const uint32_t idesc = tc05::make_idesc_tf32_ex(64, N, true, true);
    const uint64_t a_hi0 = make_mn_desc(ga, kRowLBO), a_lo0 = make_mn_desc(ga + ga_plane_bytes, kRowLBO);
    const uint64_t b_hi0 = make_mn_desc(gb, kRowLBO), b_lo0 = make_mn_desc(gb + gb_plane_bytes, kRowLBO);
#pragma unroll
    for (int ks = 0; ks < kTile / 8; ++ks) {
        const uint64_t o = (uint64_t)(ks * ((2 * kMnSBO) >> 4));
        tc05::mma_tf32(tmem_d, a_hi0 + o, b_hi0 + o, idesc, ks > 0 || accumulate);
        tc05::mma_tf32(tmem_d, a_lo0 + o, b_hi0 + o, idesc, true);
        tc05::mma_tf32(tmem_d, a_hi0 + o, b_lo0 + o, idesc, true);
    }
}

}  // namespace tctrain
