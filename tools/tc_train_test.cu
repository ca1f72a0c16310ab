// Unit test of the operand forms the tcgen05 PPO update kernel (elegantrl_b200/csrc/update_tc.cu) relies on, on ONE
// tile of 128 samples, against fp64 on the CPU (3xTF32: hi/lo planes of both operands, three UMMAs per K step):
//   T1  D1[128 x 64] = A  * W^T   A from tensor memory (kind::tf32 TS), W = nn.Linear weight [64 out][64 in] as a K-major image
//   T2  D2[128 x 64] = A  * W     W's backward image (MN-major, SWIZZLE_128B_BASE32B) as the B operand (data gradient)
//   T3  D3[ 64 x 64] = G^T * H    both operands MN-major "row-written" images (thread = sample writes its row), two passes
//                                  of 32 columns through ONE group buffer, M = 64 accumulator (rows at lanes (m % 16) + 32 * (m / 16));
//       D4[ 64 x 8]  = G^T * [x | 1 | 0]   the narrow pass (8-column rows)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o /tmp/tt tools/tc_train_test.cu && /tmp/tt
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../elegantrl_b200/csrc/tc_train.cuh"

void b200rl_set_error(const char*, ...) {}
long long g_b200rl_launches = 0;

using namespace tctrain;

constexpr int kOffW = 0;                       // 2 planes x 16 KB (forward, K-major)
constexpr int kOffWB = 2 * kWPlaneBytes;        // 2 planes x 16 KB (backward, MN-major)
constexpr int kOffGA = kOffWB + 2 * kWPlaneBytes;    // 2 planes x 32 KB
constexpr int kOffGB = kOffGA + 2 * kGAPlaneBytes;   // 2 planes x 16 KB: one 32-column group
constexpr int kSmem = kOffGB + 2 * kGroupPlaneBytes;

__global__ void __launch_bounds__(128) train_test_kernel(const float* A, const float* W, const float* G, const float* H,
                                                          float* D1, float* D2, float* D3) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc05::tmem_alloc<512>(&tmem_slot);
    if (tid == 32) { tc05::mbar_init(&bar, 1); tc05::mbar_fence_init(); }
    stage_w_planes(W, smem + kOffW, smem + kOffW + kWPlaneBytes, tid, 128);
    stage_w_planes_backward(W, smem + kOffWB, smem + kOffWB + kWPlaneBytes, tid, 128);
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t cPhi = 0, cPlo = 64, cD1 = 128, cD2 = 192, cD3 = 256, cD4 = 320;
    const uint32_t ga = tc05::smem_u32(smem + kOffGA), gb = tc05::smem_u32(smem + kOffGB);
    // this thread's row of A -> hi / lo planes in tensor memory;  rows of G and [H | 1] -> row-written images
    for (int c = 0; c < 4; ++c) {
        float v[16];
        for (int j = 0; j < 16; ++j) v[j] = A[tid * 64 + 16 * c + j];
        store_hi_lo_tmem(tmem_base + lane_base + cPhi + 16 * c, tmem_base + lane_base + cPlo + 16 * c, v);
        float g[16], h[16];
        for (int j = 0; j < 16; ++j) { g[j] = G[tid * 64 + 16 * c + j]; h[j] = H[tid * 64 + 16 * c + j]; }
        store_hi_lo_rows(smem + kOffGA, smem + kOffGA + kGAPlaneBytes, tid, 16 * c, g);
        if (c < 2) store_hi_lo_rows(smem + kOffGB, smem + kOffGB + kGroupPlaneBytes, tid, 16 * c, h);
    }
    tc05::tmem_st_wait();
    tc05::fence_proxy_async_smem();
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
        tc05::fence_after_thread_sync();
        const uint32_t w_hi = tc05::smem_u32(smem + kOffW), w_lo = w_hi + kWPlaneBytes;
        const uint32_t wb_hi = tc05::smem_u32(smem + kOffWB), wb_lo = wb_hi + kWPlaneBytes;
        issue_linear_ts(tmem_base + cD1, tmem_base + cPhi, tmem_base + cPlo, w_hi, w_lo, /*accumulate=*/false);
        issue_linear_ts_backward(tmem_base + cD2, tmem_base + cPhi, tmem_base + cPlo, wb_hi, wb_lo);
        issue_weight_grad(tmem_base + cD3, ga, kGAPlaneBytes, gb, kGroupPlaneBytes, 32);
        tc05::mma_commit(&bar);
    }
    tc05::mbar_wait(&bar, 0);
    tc05::fence_after_thread_sync();
    for (int c = 2; c < 4; ++c) {   // second pass: columns 32..63 of H through the same group buffer
        float h[16];
        for (int j = 0; j < 16; ++j) h[j] = H[tid * 64 + 16 * c + j];
        store_hi_lo_rows(smem + kOffGB, smem + kOffGB + kGroupPlaneBytes, tid, 16 * (c - 2), h);
    }
    tc05::fence_proxy_async_smem();
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
        tc05::fence_after_thread_sync();
        issue_weight_grad(tmem_base + cD3 + 32, ga, kGAPlaneBytes, gb, kGroupPlaneBytes, 32);
        tc05::mma_commit(&bar);
    }
    tc05::mbar_wait(&bar, 1);
    tc05::fence_after_thread_sync();
    {   // narrow pass: [x0 x1 x2 1 0 0 0 0] rows
        const float x8[8] = {H[tid * 64], H[tid * 64 + 1], H[tid * 64 + 2], 1.0f, 0.f, 0.f, 0.f, 0.f};
        store_hi_lo_rows8(smem + kOffGB, smem + kOffGB + kGroupPlaneBytes, tid, 0, x8);
    }
    tc05::fence_proxy_async_smem();
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (tid == 0) {
        tc05::fence_after_thread_sync();
        issue_weight_grad(tmem_base + cD4, ga, kGAPlaneBytes, gb, kGroupPlaneBytes, 8);
        tc05::mma_commit(&bar);
    }
    tc05::mbar_wait(&bar, 0);
    tc05::fence_after_thread_sync();
    for (int c = 0; c < 4; ++c) {
        float v[16];
        tc05::tmem_ld_32x32b_x16(tmem_base + lane_base + cD1 + 16 * c, v);
        tc05::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D1[tid * 64 + 16 * c + j] = v[j];
        tc05::tmem_ld_32x32b_x16(tmem_base + lane_base + cD2 + 16 * c, v);
        tc05::tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D2[tid * 64 + 16 * c + j] = v[j];
    }
    for (int c = 0; c < 9; ++c) {   // 64 + 8 columns (D3 | D4), rows 16 w + lane for lane < 16
        float v[8];
        tc05::tmem_ld_32x32b_x8(tmem_base + lane_base + cD3 + 8 * c, v);
        tc05::tmem_ld_wait();
        if (lane < 16) for (int j = 0; j < 8; ++j) D3[(16 * warp + lane) * 72 + 8 * c + j] = v[j];
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
    for (auto& v : G) v = rnd(0.01f) * (rand() % 7 == 0 ? 30.f : 1.f);
    for (auto& v : H) v = rnd(1.5f);
    float *dA, *dW, *dG, *dH, *dD1, *dD2, *dD3;
    auto up = [](float** d, const std::vector<float>& h) { cudaMalloc(d, h.size() * 4); cudaMemcpy(*d, h.data(), h.size() * 4, cudaMemcpyHostToDevice); };
    up(&dA, A); up(&dW, W); up(&dG, G); up(&dH, H);
    cudaMalloc(&dD1, 128 * 64 * 4); cudaMalloc(&dD2, 128 * 64 * 4); cudaMalloc(&dD3, 64 * 72 * 4);
    cudaMemset(dD1, 0, 128 * 64 * 4); cudaMemset(dD2, 0, 128 * 64 * 4); cudaMemset(dD3, 0, 64 * 72 * 4);
    cudaFuncSetAttribute(train_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    train_test_kernel<<<1, 128, kSmem>>>(dA, dW, dG, dH, dD1, dD2, dD3);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> D1(128 * 64), D2(128 * 64), D3(64 * 72);
    cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, m1 = 0, e2 = 0, m2 = 0, e3 = 0, m3 = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 64; ++n) {
            double s1 = 0, s2 = 0;
            for (int k = 0; k < 64; ++k) { s1 += (double)A[r * 64 + k] * W[n * 64 + k]; s2 += (double)A[r * 64 + k] * W[k * 64 + n]; }
            e1 = fmax(e1, fabs(D1[r * 64 + n] - s1)); m1 = fmax(m1, fabs(s1));
            e2 = fmax(e2, fabs(D2[r * 64 + n] - s2)); m2 = fmax(m2, fabs(s2));
        }
    for (int m = 0; m < 64; ++m)
        for (int n = 0; n < 72; ++n) {
            double s = 0;
            for (int b = 0; b < 128; ++b) {
                const double q = n < 64 ? H[b * 64 + n] : (n < 67 ? H[b * 64 + (n - 64)] : (n == 67 ? 1.0 : 0.0));
                s += (double)G[b * 64 + m] * q;
            }
            e3 = fmax(e3, fabs(D3[m * 72 + n] - s)); m3 = fmax(m3, fabs(s));
        }
    printf("T1 A*W^T (TS, K-major B)      max|err| %.3e (max|ref| %.3f)\n", e1, m1);
    printf("T2 A*W   (TS, MN-major image)  max|err| %.3e (max|ref| %.3f)\n", e2, m2);
    printf("T3 G^T*[H | x,1] (M=64, MN/MN)  max|err| %.3e (max|ref| %.3f)\n", e3, m3);
    const bool ok = e1 < 2e-5 * fmax(1.0, m1) && e2 < 2e-5 * fmax(1.0, m2) && e3 < 2e-5 * fmax(1.0, m3);
    printf(ok ? "TC TRAIN TEST OK\n" : "TC TRAIN TEST FAILED\n");
    return ok ? 0 : 1;
}
