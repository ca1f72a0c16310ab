// Second probe: MN-major tf32 operands need the SWIZZLE_128B_BASE32B layout type ("for mn-major tf32 operands, SW128_32B is
// the only available smem layout", cutlass sm100_common.inl).  Atom (cute Layout_MN_SW128_32B_Atom): 32 MN elements (128 B)
// contiguous x 4 K rows at 128 B, Swizzle<2,5,2> on BYTE addresses: bits [5,7) ^= bits [7,9) (decoded by tc_mn_probe3.cu).
//   element (mn, k) -> (mn / 32) * LBO + (k / 4) * SBO + (k % 4) * 128 + (mn % 32) * 4, then swizzled
// Tries the writer with / without the swizzle and both assignments of the descriptor's LBO / SBO fields.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/bin/tc_mn_probe2 tools/tc_mn_probe2.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../elegantrl_b200/csrc/tc_train.cuh"

void b200rl_set_error(const char*, ...) {}
long long g_b200rl_launches = 0;

constexpr int kOffGA = 0, kOffGB = 32768, kOffWB = 65536, kSmem = 65536 + 16384;
constexpr uint32_t kSBO = 512;   // K group (4 rows) stride

__device__ __forceinline__ uint64_t desc_sw32b(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = tc05::make_smem_desc_ex(addr, lbo, sbo);
    d |= (uint64_t)1 << 61;   // layout_type = SWIZZLE_128B_BASE32B
    return d;
}
__device__ __forceinline__ uint32_t sw_addr(uint32_t mn, uint32_t k, uint32_t lbo, int swz) {
    uint32_t a = (mn / 32) * lbo + (k / 4) * kSBO + (k % 4) * 128 + (mn % 32) * 4;
    if (swz) a ^= ((a >> 7) & 3) << 5;   // decoded by tools/tc_mn_probe3.cu: 32-byte chunk index ^= row index
    return a;
}

__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* W, const float* G, const float* H, float* D, int test, int swz, int swapf) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tc05::tmem_alloc<512>(&tmem_slot);
    if (tid == 32) { tc05::mbar_init(&bar, 1); tc05::mbar_fence_init(); }
    // images (hi plane only: single tf32 product is enough to tell right from wrong)
    const uint32_t lbo_g = 128 / 4 * kSBO;   // 16 KB between MN groups of the [128 samples][64] images
    const uint32_t lbo_w = 64 / 4 * kSBO;    // 8 KB for the 64 x 64 weight
    for (int c = 0; c < 64; ++c) {
        *reinterpret_cast<float*>(smem + kOffGA + sw_addr(c, tid, lbo_g, swz)) = tc05::tf32_hi(G[tid * 64 + c]);
        *reinterpret_cast<float*>(smem + kOffGB + sw_addr(c, tid, lbo_g, swz)) = tc05::tf32_hi(H[tid * 64 + c]);
    }
    for (int i = tid; i < 64 * 64; i += 128) {   // backward image of W: mn = k_in, K = j_out
        const int j = i >> 6, k = i & 63;
        *reinterpret_cast<float*>(smem + kOffWB + sw_addr(k, j, lbo_w, swz)) = tc05::tf32_hi(W[i]);
    }
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t cPhi = 0, cPlo = 64, cD = 128;
    for (int c = 0; c < 4; ++c) {
        float v[16];
        for (int j = 0; j < 16; ++j) v[j] = A[tid * 64 + 16 * c + j];
        tctrain::store_hi_lo_tmem(tmem_base + lane_base + cPhi + 16 * c, tmem_base + lane_base + cPlo + 16 * c, v);
    }
    {
        uint32_t z[16];
        for (int j = 0; j < 16; ++j) z[j] = 0u;
        for (int c = 0; c < 8; ++c) tc05::tmem_st_32x32b_x16(tmem_base + lane_base + cD + 16 * c, z);
    }
    tc05::tmem_st_wait();
    tc05::fence_proxy_async_smem();
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
        tc05::fence_after_thread_sync();
        const uint32_t ga = tc05::smem_u32(smem + kOffGA), gb = tc05::smem_u32(smem + kOffGB), wb = tc05::smem_u32(smem + kOffWB);
        auto mk = [&](uint32_t addr, uint32_t lbo) { return swapf ? desc_sw32b(addr, kSBO, lbo) : desc_sw32b(addr, lbo, kSBO); };
        if (test == 2) {          // D[128 x 64] = A * W  (TS, B = backward image of W, MN-major)
            const uint32_t idesc = tc05::make_idesc_tf32_ex(128, 64, false, true);
            for (int ks = 0; ks < 8; ++ks) tc05::mma_tf32_ts(tmem_base + cD, tmem_base + cPhi + 8 * ks, mk(wb + ks * 2 * kSBO, lbo_w), idesc, ks > 0);
        } else {                  // D[64 x 64] = G^T H over 128 samples (SS, both MN-major)
            const uint32_t idesc = tc05::make_idesc_tf32_ex(64, 64, true, true);
            for (int ks = 0; ks < 16; ++ks)
                tc05::mma_tf32(tmem_base + cD, mk(ga + ks * 2 * kSBO, lbo_g), mk(gb + ks * 2 * kSBO, lbo_g), idesc, ks > 0);
        }
        tc05::mma_commit(&bar);
    }
    tc05::mbar_wait(&bar, 0);
    tc05::fence_after_thread_sync();
    for (int c = 0; c < 8; ++c) {
        float v[16];
        tc05::tmem_ld_32x32b_x16(tmem_base + lane_base + cD + 16 * c, v);
        tc05::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[tid * 128 + 16 * c + j] = v[j];
    }
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<512>(tmem_base);
}

int main() {
    srand(5);
    auto rnd = [](float s) { return ((rand() / (float)RAND_MAX) * 2.f - 1.f) * s; };
    std::vector<float> A(128 * 64), W(64 * 64), G(128 * 64), H(128 * 64);
    for (auto& v : A) v = rnd(2.0f);
    for (auto& v : W) v = rnd(0.4f);
    for (auto& v : G) v = rnd(1.0f);
    for (auto& v : H) v = rnd(1.5f);
    float *dA, *dW, *dG, *dH, *dD;
    auto up = [](float** d, const std::vector<float>& h) { cudaMalloc(d, h.size() * 4); cudaMemcpy(*d, h.data(), h.size() * 4, cudaMemcpyHostToDevice); };
    up(&dA, A); up(&dW, W); up(&dG, G); up(&dH, H);
    cudaMalloc(&dD, 128 * 128 * 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    std::vector<float> D(128 * 128);
    for (int test = 2; test <= 3; ++test)
        for (int swz = 0; swz < 2; ++swz)
            for (int swapf = 0; swapf < 2; ++swapf) {
                cudaMemset(dD, 0, 128 * 128 * 4);
                probe_kernel<<<1, 128, kSmem>>>(dA, dW, dG, dH, dD, test, swz, swapf);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("test %d swz %d swapf %d: CUDA error %s\n", test, swz, swapf, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
                double err = 0, mx = 0; int nz = 0;
                const int M = test == 2 ? 128 : 64;
                for (int m = 0; m < M; ++m)
                    for (int n = 0; n < 64; ++n) {
                        double s = 0;
                        if (test == 2) for (int k = 0; k < 64; ++k) s += (double)A[m * 64 + k] * W[k * 64 + n];
                        else for (int b = 0; b < 128; ++b) s += (double)G[b * 64 + m] * H[b * 64 + n];
                        const int lane = test == 2 ? m : (m % 16) + 32 * (m / 16);
                        const double got = D[lane * 128 + n];
                        err = fmax(err, fabs(got - s)); mx = fmax(mx, fabs(s)); nz += got != 0.0;
                    }
                printf("test %d writer-swizzle %d swap-fields %d: max|err| %.3e (max|ref| %.3f) nonzero %d\n", test, swz, swapf, err, mx, nz);
            }
    return 0;
}
