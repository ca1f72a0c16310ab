// Building blocks of the "TS" MLP evaluation on tcgen05 (A operand in tensor memory), shared by rollout_ts.cu and the
// unit test tools/tc_ts_test.cu.  One tile = 128 rows (TMEM lanes); a thread owns one row.  For a 3 -> 64 -> 64 -> 1
// GELU MLP (reference ActorPPO / CriticPPO with net_dims (64, 64), elegantrl/agents/AgentPPO.py:348-441):
//
//   layer 1   x~ = [x_hi(3), 1, x_lo(3), 0] (tf32, one 32-byte row of a K-major shared-memory operand)
//             Z1 = x~ * [W1_hi; b1_hi; W1_hi; 0]^T + x~ * [W1_lo; b1_lo; 0; 0]^T          2 x UMMA 128x64x8 kind::tf32 (SS)
//             -> 64 TMEM columns "X" (fp32 pre-activations, bias included; 3xTF32 accuracy)
//   in place  every thread reads 16 columns of its row, applies GELU, splits h = hi + lo with hi = h & 0xFFFFE000 (11
//             significant bits: exact in fp16) and lo = h - hi (exact in fp32, rounded to fp16), packs {hi, lo} as one
//             f16x2 word and writes it back to the SAME column: the 64 columns now hold the K-major fp16 A operand of
//             layer 2 with K = 128 (hidden unit j at K positions 2j, 2j+1).
//   layer 2   Z2 = b2 (UMMA 128x64x8 kind::tf32 with a constant A = [1, 1, 0...] and B = [b2_hi, b2_lo, 0...])
//                  + A * B_hi^T + A * B_lo^T   (8 K-steps x 2 UMMA 128x64x16 kind::f16, A from TMEM, B from shared memory)
//             where B_p[n][2j] = B_p[n][2j+1] = fp16 piece p of W2[n][j], i.e. (h_hi + h_lo) * (w_hi + w_lo): ~22 bits.
//   head      GELU(Z2) . w3 + b3 on CUDA cores (thread = row).
// Nothing of the activations ever touches shared memory; the only shared-memory operands are the weights.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc05.cuh"

namespace tsmlp {

constexpr int kHid = 64;
constexpr int kTileRows = 128;

// ---- shared-memory operand images (bytes)
constexpr int kB2PlaneBytes = kHid * 2 * kHid * 2;   // [64 out][128 K] fp16 = 16 KB
constexpr int kB1Bytes = kHid * 8 * 4;               // [64 out][8 K] tf32 = 2 KB
constexpr int kA1Bytes = kTileRows * 8 * 4;          // [128 rows][8 K] tf32 = 4 KB
constexpr uint32_t kSboK8 = 256;                     // stride between 8-row groups of a K = 8 tf32 operand (2 K-chunks x 128 B)
constexpr uint32_t kSboB2 = 2048;                    // ... of the K = 128 fp16 operand (16 K-chunks x 128 B)

struct NetImage {            // byte offsets (relative to a 1024-aligned base) of one net's operands
    int b2[2];               // fp16 planes hi / lo
    int b1[2];               // tf32: {W1_hi, b1_hi, W1_hi, 0}, {W1_lo, b1_lo, 0, 0}
    int bb;                  // tf32: {b2_hi, b2_lo, 0...}
};

// byte offset of element (row, k) of a K = 8 tf32 operand (rows x 32 bytes, canonical no-swizzle K-major layout)
DEV uint32_t k8_offset(int row, int k) { return (uint32_t)((row >> 3) * 256 + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4); }

// stage one net's weights into its operand images (all threads of the CTA; caller fences + syncs afterwards)
DEV void stage_net(const b200rl_net& net, unsigned char* base, const NetImage& im, int tid, int nthreads) {
    for (int i = tid; i < kHid * kHid; i += nthreads) {
        const int n = i >> 6, j = i & 63;
        const float w = net.weight[1][i];
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const uint32_t off = tc05::operand_offset(n, j, kHid);   // the {2j, 2j+1} fp16 pair sits where tf32 element j would
        *reinterpret_cast<__half2*>(base + im.b2[0] + off) = __half2(hi, hi);
        *reinterpret_cast<__half2*>(base + im.b2[1] + off) = __half2(lo, lo);
    }
    for (int i = tid; i < kHid * 8; i += nthreads) {
        const int n = i >> 3, k = i & 7, kk = k & 3;
        const float full = kk < 3 ? net.weight[0][n * 3 + kk] : net.bias[0][n];
        const float hi = tc05::tf32_hi(full), lo = full - hi;
        const float b2 = net.bias[1][n], b2hi = tc05::tf32_hi(b2);
        const uint32_t off = k8_offset(n, k);
        *reinterpret_cast<float*>(base + im.b1[0] + off) = k == 7 ? 0.0f : hi;      // [W_hi(3), b_hi, W_hi(3), 0]
        *reinterpret_cast<float*>(base + im.b1[1] + off) = k < 4 ? lo : 0.0f;       // [W_lo(3), b_lo, 0, 0, 0, 0]
        *reinterpret_cast<float*>(base + im.bb + off) = k == 0 ? b2hi : (k == 1 ? b2 - b2hi : 0.0f);
    }
}
// constant A operand of the bias UMMA: [128 rows][8] tf32 with columns 0 and 1 equal to one
DEV void stage_const_a(unsigned char* dst, int tid, int nthreads) {
    for (int i = tid; i < kTileRows * 8; i += nthreads) {
        const int r = i >> 3, k = i & 7;
        *reinterpret_cast<float*>(dst + k8_offset(r, k)) = k < 2 ? 1.0f : 0.0f;
    }
}

// a thread writes its row of the layer-1 A operand: x~ = [x_hi(3), 1, x_lo(3), 0]
DEV void write_x_row(unsigned char* a1, int row, const float (&x)[3]) {
    const float h0 = tc05::tf32_hi(x[0]), h1 = tc05::tf32_hi(x[1]), h2 = tc05::tf32_hi(x[2]);
    const uint32_t addr = tc05::smem_u32(a1) + (uint32_t)((row >> 3) * 256 + (row & 7) * 16);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(h0), "f"(h1), "f"(h2), "f"(1.0f) : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr + 128), "f"(x[0] - h0), "f"(x[1] - h1), "f"(x[2] - h2), "f"(0.0f) : "memory");
}

// Packed exact-erf GELU on a pair (tools/fit_gelu.py packed_form): zn = -min(|x|, L); t = zn * P(zn) - 1;
// GELU(x) = max(x, 0) + zn * exp2(t); 4 FMNMX + 7 FFMA2 + 2 MUFU.EX2 per pair, max abs error 5.8e-7.
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

// columns [col, col + 16) of this thread's row: fp32 pre-activations -> GELU -> {hi, lo} fp16 pairs, in place
template <bool GELU>
DEV void hidden_chunk_inplace(uint32_t taddr) {
    float v[16];
    tc05::tmem_ld_32x32b_x16(taddr, v);
    tc05::tmem_ld_wait();
    uint32_t w[16];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        float2 g = make_float2(v[2 * p], v[2 * p + 1]);
        if (GELU) g = gelu_fast2(g);
        const float2 hi = make_float2(tc05::tf32_hi(g.x), tc05::tf32_hi(g.y));
        const float2 lo = __ffma2_rn(hi, splat(-1.0f), g);
        w[2 * p] = tc05::pack_f16x2(lo.x, hi.x);
        w[2 * p + 1] = tc05::pack_f16x2(lo.y, hi.y);
    }
    tc05::tmem_st_32x32b_x16(taddr, w);
}

// tcgen05.wait::ld that is ALSO a data dependence on the 16 destination registers of an earlier tcgen05.ld: the loads are
// asynchronous, and a software-pipelined sequence (issue the next chunk's load, then work on the current chunk) must keep
// the compiler from scheduling a use of the registers above the wait.
DEV void tmem_ld_wait_tied(float (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                   "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :: "memory");
}
// GELU -> {hi, lo} fp16 pairs of 16 pre-activations already in registers, written back to columns [taddr, taddr + 16)
template <bool GELU>
DEV void hidden_chunk_from_regs(uint32_t taddr, const float (&v)[16]) {
    uint32_t w[16];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        float2 g = make_float2(v[2 * p], v[2 * p + 1]);
        if (GELU) g = gelu_fast2(g);
        const float2 hi = make_float2(tc05::tf32_hi(g.x), tc05::tf32_hi(g.y));
        const float2 lo = __ffma2_rn(hi, splat(-1.0f), g);
        w[2 * p] = tc05::pack_f16x2(lo.x, hi.x);
        w[2 * p + 1] = tc05::pack_f16x2(lo.y, hi.y);
    }
    tc05::tmem_st_32x32b_x16(taddr, w);
}
// 16 terms of the head's dot product from registers
template <bool GELU>
DEV float2 head_chunk_from_regs(const float (&v)[16], const float* w3c, float2 out2) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        const float4 w = *reinterpret_cast<const float4*>(w3c + q4 * 4);
        float2 g01 = make_float2(v[q4 * 4 + 0], v[q4 * 4 + 1]), g23 = make_float2(v[q4 * 4 + 2], v[q4 * 4 + 3]);
        if (GELU) { g01 = gelu_fast2(g01); g23 = gelu_fast2(g23); }
        out2 = __ffma2_rn(g01, make_float2(w.x, w.y), out2);
        out2 = __ffma2_rn(g23, make_float2(w.z, w.w), out2);
    }
    return out2;
}
// head with the tensor-memory loads software-pipelined: the load of chunk c + 1 is in flight while chunk c is evaluated
template <bool GELU>
DEV float head_dot_pipelined(uint32_t taddr_d, const float* w3, float b3) {
    float2 out2 = make_float2(b3, 0.0f);
    float a[16], b[16];
    tc05::tmem_ld_32x32b_x16(taddr_d, a);
    tmem_ld_wait_tied(a);
    tc05::tmem_ld_32x32b_x16(taddr_d + 16, b);
    out2 = head_chunk_from_regs<GELU>(a, w3, out2);
    tmem_ld_wait_tied(b);
    tc05::tmem_ld_32x32b_x16(taddr_d + 32, a);
    out2 = head_chunk_from_regs<GELU>(b, w3 + 16, out2);
    tmem_ld_wait_tied(a);
    tc05::tmem_ld_32x32b_x16(taddr_d + 48, b);
    out2 = head_chunk_from_regs<GELU>(a, w3 + 32, out2);
    tmem_ld_wait_tied(b);
    out2 = head_chunk_from_regs<GELU>(b, w3 + 48, out2);
    return out2.x + out2.y;
}

// head: sum_j GELU(Z2[j]) * w3[j] + b3 for this thread's row; w3 = 64 floats in shared memory (16-byte aligned)
template <bool GELU>
DEV float head_dot(uint32_t taddr_d, const float* w3, float b3) {
    float2 out2 = make_float2(b3, 0.0f);
#pragma unroll 1
    for (int cc = 0; cc < kHid / 16; ++cc) {
        float v[16];
        tc05::tmem_ld_32x32b_x16(taddr_d + cc * 16, v);
        tc05::tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            const float4 w = *reinterpret_cast<const float4*>(w3 + cc * 16 + q4 * 4);
            float2 g01 = make_float2(v[q4 * 4 + 0], v[q4 * 4 + 1]), g23 = make_float2(v[q4 * 4 + 2], v[q4 * 4 + 3]);
            if (GELU) { g01 = gelu_fast2(g01); g23 = gelu_fast2(g23); }
            out2 = __ffma2_rn(g01, make_float2(w.x, w.y), out2);
            out2 = __ffma2_rn(g23, make_float2(w.z, w.w), out2);
        }
    }
    return out2.x + out2.y;
}

// ---- issuer side (ONE elected thread): descriptors of one net
struct NetDescs {
    uint64_t b1[2], bb, b2[2];   // b2: K-step 0 (add k * (256 >> 4) to the low word for K-step k)
};
DEV NetDescs make_descs(unsigned char* base, const NetImage& im) {
    NetDescs d;
    d.b1[0] = tc05::make_smem_desc(tc05::smem_u32(base + im.b1[0]), kSboK8);
    d.b1[1] = tc05::make_smem_desc(tc05::smem_u32(base + im.b1[1]), kSboK8);
    d.bb = tc05::make_smem_desc(tc05::smem_u32(base + im.bb), kSboK8);
    d.b2[0] = tc05::make_smem_desc(tc05::smem_u32(base + im.b2[0]), kSboB2);
    d.b2[1] = tc05::make_smem_desc(tc05::smem_u32(base + im.b2[1]), kSboB2);
    return d;
}
// layer 1 of one tile: X[128 x 64] = x~ * B1   (tmem_x: lane field 0)
DEV void issue_layer1(uint32_t tmem_x, uint64_t a1_desc, const NetDescs& d) {
    constexpr uint32_t idesc = tc05::make_idesc_tf32(kTileRows, kHid);
    tc05::mma_tf32(tmem_x, a1_desc, d.b1[0], idesc, false);
    tc05::mma_tf32(tmem_x, a1_desc, d.b1[1], idesc, true);
}
DEV void issue_bias(uint32_t tmem_d, uint64_t aconst_desc, const NetDescs& d) {
    constexpr uint32_t idesc = tc05::make_idesc_tf32(kTileRows, kHid);
    tc05::mma_tf32(tmem_d, aconst_desc, d.bb, idesc, false);
}
// layer 2, hidden units [16 c, 16 c + 16): two K-steps of 16 fp16, hi and lo plane each
DEV void issue_layer2_chunk(uint32_t tmem_d, uint32_t tmem_x, const NetDescs& d, int c) {
    constexpr uint32_t idesc = tc05::make_idesc_f16(kTileRows, kHid);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const int k = 2 * c + ks;
        tc05::mma_f16_ts(tmem_d, tmem_x + 8 * k, d.b2[0] + (uint64_t)(k * (256 >> 4)), idesc, true);
        tc05::mma_f16_ts(tmem_d, tmem_x + 8 * k, d.b2[1] + (uint64_t)(k * (256 >> 4)), idesc, true);
    }
}

}  // namespace tsmlp
