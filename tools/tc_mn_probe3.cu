// Third probe: DECODE the shared-memory layout tcgen05.mma kind::tf32 reads for an MN-major operand with layout type
// SWIZZLE_128B_BASE32B.  The operand region is filled with its own word index; the other operand is one-hot, so that the
// accumulator reveals which shared-memory word the tensor core fetched for every logical element (mn, k).
//   B probe (TS):  A[m][k] = (k == m % 8) from tensor memory (trusted);   D[m][n] = B(n, k = m % 8)      m < 8
//   A probe (SS):  B[n][k] = (k == n % 8), K-major no-swizzle (trusted);  D[m][n] = A(m, k = n % 8)      n < 8
// word index = lo + 1024 * hi with two UMMAs (tf32 keeps integers < 2048 exact).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/bin/tc_mn_probe3 tools/tc_mn_probe3.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../elegantrl_b200/csrc/tc_train.cuh"

void b200rl_set_error(const char*, ...) {}
long long g_b200rl_launches = 0;

constexpr int kRegion = 65536;                 // bytes of the probed operand region (16 K words)
constexpr int kOffHot = kRegion, kSmem = kRegion + 8192;

__global__ void __launch_bounds__(128) probe_kernel(float* D, int which, int part, uint32_t lbo, uint32_t sbo, int layout_type, int M) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc05::tmem_alloc<512>(&tmem_slot);
    if (tid == 32) { tc05::mbar_init(&bar, 1); tc05::mbar_fence_init(); }
    for (int i = tid; i < kRegion / 4; i += 128) reinterpret_cast<float*>(smem)[i] = (float)(part ? (i >> 10) : (i & 1023));
    for (int i = tid; i < 64 * 8; i += 128) {   // one-hot B, K-major no-swizzle [64][8]
        const int n = i >> 3, k = i & 7;
        *reinterpret_cast<float*>(smem + kOffHot + tc05::operand_offset(n, k, 8)) = (k == (n & 7)) ? 1.0f : 0.0f;
    }
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t cA = 0, cD = 128;
    {
        uint32_t a[16], z[16];
        for (int j = 0; j < 16; ++j) { a[j] = (j < 8 && j == (tid & 7)) ? __float_as_uint(1.0f) : 0u; z[j] = 0u; }
        tc05::tmem_st_32x32b_x16(tmem_base + lane_base + cA, a);
        for (int c = 0; c < 8; ++c) tc05::tmem_st_32x32b_x16(tmem_base + lane_base + cD + 16 * c, z);
    }
    tc05::tmem_st_wait();
    tc05::fence_proxy_async_smem();
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
        tc05::fence_after_thread_sync();
        uint64_t d = tc05::make_smem_desc_ex(tc05::smem_u32(smem), lbo, sbo) | ((uint64_t)layout_type << 61);
        if (which == 0) {   // B probe
            tc05::mma_tf32_ts(tmem_base + cD, tmem_base + cA, d, tc05::make_idesc_tf32_ex(128, 64, false, true), false);
        } else {            // A probe
            const uint64_t hot = tc05::make_smem_desc_ex(tc05::smem_u32(smem + kOffHot), 128, 256);
            tc05::mma_tf32(tmem_base + cD, d, hot, tc05::make_idesc_tf32_ex(M, 64, true, false), false);
        }
        tc05::mma_commit(&bar);
    }
    tc05::mbar_wait(&bar, 0);
    tc05::fence_after_thread_sync();
    for (int c = 0; c < 4; ++c) {
        float v[16];
        tc05::tmem_ld_32x32b_x16(tmem_base + lane_base + cD + 16 * c, v);
        tc05::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[tid * 64 + 16 * c + j] = v[j];
    }
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<512>(tmem_base);
}

int main() {
    float* dD;
    cudaMalloc(&dD, 128 * 64 * 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    std::vector<float> lo(128 * 64), hi(128 * 64);
    struct Cfg { int which; uint32_t lbo, sbo; int type; int M; };
    const Cfg cfgs[] = {{0, 8192, 512, 1, 128}, {0, 4096, 1024, 1, 128}, {0, 8192, 512, 2, 128}, {1, 16384, 512, 1, 64}, {1, 16384, 512, 1, 128}, {1, 4096, 1024, 1, 64}};
    for (const Cfg& c : cfgs) {
        for (int part = 0; part < 2; ++part) {
            cudaMemset(dD, 0, 128 * 64 * 4);
            probe_kernel<<<1, 128, kSmem>>>(dD, c.which, part, c.lbo, c.sbo, c.type, c.M);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy((part ? hi : lo).data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        }
        printf("== %s probe: LBO %u SBO %u layout_type %d M %d   (byte offset of logical element; rows: k = 0..7)\n", c.which ? "A" : "B", c.lbo, c.sbo, c.type, c.M);
        for (int k = 0; k < 8; ++k) {
            printf("k=%d:", k);
            const int count = c.which ? c.M : 64;
            for (int mn = 0; mn < count; ++mn) {
                // B probe: D[m = k][n = mn];  A probe: D[lane(m = mn)][n = k]
                int row, col;
                if (c.which == 0) { row = k; col = mn; }
                else { row = (c.M == 128) ? mn : (mn % 16) + 32 * (mn / 16); col = k; }
                const int word = (int)lo[row * 64 + col] + 1024 * (int)hi[row * 64 + col];
                printf(" %d", word * 4);
            }
            printf("\n");
        }
    }
    return 0;
}
