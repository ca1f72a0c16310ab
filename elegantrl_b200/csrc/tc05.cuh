// Minimal hand-written tcgen05 / TMEM / mbarrier layer for sm_100a (inline PTX, no CUTLASS).
//
// Used by rollout_tc.cu for the 64-wide hidden layers of the actor / critic MLPs: the activations of a tile
// of 128 envs are the A operand (K-major, written to shared memory by the threads that produced them), the
// nn.Linear weight [out, in] is the B operand as it lies (K-major), the fp32 accumulator lives in TMEM
// (lane = env row, column = output feature) and is read back with tcgen05.ld 32x32b (one thread = one row).
//
// fp32 parity (rtol 1e-4) is kept with the 3xTF32 split: x = hi + lo, hi = tf32(x) (13 low mantissa bits
// cleared), lo = x - hi (exact);  A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo, all accumulated in fp32 in TMEM.
//
// Shared-memory operand layout ("no-swizzle, K-major" canonical UMMA layout, cf. the SmemDescriptor notes in
// cute/arch/mma_sm100_desc.hpp): core matrix = 8 rows x 16 bytes (4 tf32), stored as 128 contiguous bytes;
//   byte offset of element (row, k) = (row / 8) * SBO + (k / 4) * LBO + (row % 8) * 16 + (k % 4) * 4
// with LBO = 128 (the K-chunks of one 8-row group are contiguous) and SBO = (K / 4) * 128.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// suspend-time hint: the waiting thread is parked by the hardware until the phase completes (or this many ns pass)
// instead of spinning through try_wait / YIELD / BRA and stealing issue slots from the warps that do the work
constexpr uint32_t kSuspendHintNs = 20000u;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(kSuspendHintNs) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// ------------------------------------------------------------------------------------------- fences
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core reads A/B through it)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// --------------------------------------------------------------------------------------------- TMEM
// One full warp allocates `ncols` (power of two >= 32) columns; the base address lands in *smem_slot.
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives columns [c, c+32) of TMEM lane (base_lane + i).
// `taddr` = (lane << 16) | column; the lane field must be the warp's quarter: 32 * (warp_id % 4).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- MMA
constexpr uint32_t kLBO = 128;  // bytes between consecutive 16-byte K-chunks of one 8-row group

// K-major, no-swizzle shared-memory matrix descriptor (see file header).  sbo_bytes = (K / 4) * 128.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address  bits [0,14)
    d |= (uint64_t)((kLBO >> 4) & 0x3FFF) << 16;         // leading byte offset bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride byte offset  bits [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version 1 (Blackwell) bits [46,48)
    return d;                                            // base_offset 0, lbo_mode 0, layout_type 0 = no swizzle
}
// byte offset of element (row, k) of a K-major operand with K columns (k in tf32 elements)
__device__ __forceinline__ uint32_t operand_offset(int row, int k, int K) {
    return (uint32_t)((row >> 3) * (K >> 2) * 128 + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4);
}

// instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major, dense
__device__ __host__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4)                    // c_format = F32
           | (2u << 7)                  // a_format = TF32
           | (2u << 10)                 // b_format = TF32
           | ((uint32_t)(N >> 3) << 17)  // n_dim
           | ((uint32_t)(M >> 4) << 24);  // m_dim
}

// D[tmem] (+)= A[smem] * B[smem]^T   one UMMA (M x N x 8 for tf32); issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// one lane of a fully converged warp (elect.sync): the canonical issuer of tcgen05.mma / commit
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred) :: "memory");
    return pred != 0;
}

// 3xTF32 split
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

}  // namespace tc05

// ------------------------------------------------------------------ additions for the TS (A-in-TMEM) rollout kernel
namespace tc05 {

// registers -> TMEM, 32 lanes x 16 columns of 32-bit: thread i of the warp writes columns [c, c+16) of TMEM lane
// (base_lane + i).  Completion: tmem_st_wait() before the data may be consumed by another agent.
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// instruction descriptor, kind::f16 with fp16 A and B, fp32 accumulate, both K-major, dense (cf. cute UMMA::InstrDescriptor)
__device__ __host__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4)                     // c_format = F32
           | (0u << 7) | (0u << 10)      // a_format = b_format = F16
           | ((uint32_t)(N >> 3) << 17)  // n_dim
           | ((uint32_t)(M >> 4) << 24);  // m_dim
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16 (M x N x 16): A = 8 TMEM columns (two fp16 K-elements per 32-bit cell,
// lane = row) starting at tmem_a; B = K-major shared-memory descriptor.  Issued by ONE thread.
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// {hi half: fp16(a), lo half: fp16(b)}, round to nearest even
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b));
    return r;
}

}  // namespace tc05

// ------------------------------------------------- additions for the tcgen05 PPO update kernel (update_tc.cu)
namespace tc05 {

// general shared-memory matrix descriptor, no swizzle: `lbo_bytes` / `sbo_bytes` as the canonical layouts define them
//   K-major  ((8,m),(T,2k)):((1T,SBO),(1,LBO))   SBO = stride between 8-row groups, LBO = stride between 16-byte K chunks
//   MN-major ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)) SBO = stride between 16-byte MN groups, LBO = stride between 8-deep K groups
// (cute/atom/mma_traits_sm100.hpp, make_umma_desc).  An image written K-major for X [rows][K] is at the same time the MN-major
// image of X^T with the two strides swapped -- that is how the backward pass reads nn.Linear weights transposed in place.
__device__ __forceinline__ uint64_t make_smem_desc_ex(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// kind::tf32 instruction descriptor with the operand majors: bit 15 = A is MN-major, bit 16 = B is MN-major
__device__ __host__ constexpr uint32_t make_idesc_tf32_ex(int M, int N, bool a_mn_major, bool b_mn_major) {
    return make_idesc_tf32(M, N) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16);
}
// D[tmem] (+)= A[tmem] * B[smem], kind::tf32 (M x N x 8): A = 8 TMEM columns of fp32 (lane = row) starting at tmem_a
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace tc05
