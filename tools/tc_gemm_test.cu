// Unit test of the hand-written tcgen05 path (elegantrl_b200/csrc/tc05.cuh) on a single 128x64x64 tile:
//   D[128][64] = A[128][64] * B[64][64]^T   with 1xTF32 and 3xTF32, against an fp64 CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tc tools/tc_gemm_test.cu && /tmp/tc
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../elegantrl_b200/csrc/tc05.cuh"

constexpr int M = 128, N = 64, K = 64;

__global__ void __launch_bounds__(128) gemm_test_kernel(const float* A, const float* B, float* D, int split) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* a_hi = reinterpret_cast<float*>(smem);              // M*K floats
    float* a_lo = a_hi + M * K;
    float* b_hi = a_lo + M * K;                                // N*K floats
    float* b_lo = b_hi + N * K;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) tc05::tmem_alloc<64>(&tmem_base_slot);
    if (tid == 0) { tc05::mbar_init(&bar, 1); tc05::mbar_fence_init(); }
    // operands -> canonical K-major no-swizzle layout, split hi/lo
    for (int i = tid; i < M * K; i += 128) {
        int r = i / K, k = i % K;
        float x = A[i], hi = tc05::tf32_hi(x);
        uint32_t off = tc05::operand_offset(r, k, K) / 4;
        a_hi[off] = hi; a_lo[off] = x - hi;
    }
    for (int i = tid; i < N * K; i += 128) {
        int r = i / K, k = i % K;
        float x = B[i], hi = tc05::tf32_hi(x);
        uint32_t off = tc05::operand_offset(r, k, K) / 4;
        b_hi[off] = hi; b_lo[off] = x - hi;
    }
    tc05::fence_proxy_async_smem();
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = tmem_base_slot;

    if (tid == 0) {
        constexpr uint32_t idesc = tc05::make_idesc_tf32(M, N);
        constexpr uint32_t sbo = (K / 4) * 128;
        const float* a_ops[3] = {a_hi, a_lo, a_hi};
        const float* b_ops[3] = {b_hi, b_hi, b_lo};
        const int terms = split ? 3 : 1;
        bool acc = false;
        for (int t = 0; t < terms; ++t) {
            for (int ks = 0; ks < K / 8; ++ks) {
                uint64_t ad = tc05::make_smem_desc(tc05::smem_u32(a_ops[t]) + ks * 2 * tc05::kLBO, sbo);
                uint64_t bd = tc05::make_smem_desc(tc05::smem_u32(b_ops[t]) + ks * 2 * tc05::kLBO, sbo);
                tc05::mma_tf32(tmem_base, ad, bd, idesc, acc);
                acc = true;
            }
        }
        tc05::mma_commit(&bar);
    }
    tc05::mbar_wait(&bar, 0);
    tc05::fence_after_thread_sync();
    const int row = tid;  // TMEM lane = row; warp w reads lanes 32w..32w+31
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tc05::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tc05::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[row * N + c0 + j] = v[j];
    }
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<64>(tmem_base);
}

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N);
    srand(1);
    for (auto& x : A) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (auto& x : B) x = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof(float) * M * K); cudaMalloc(&dB, sizeof(float) * N * K); cudaMalloc(&dD, sizeof(float) * M * N);
    cudaMemcpy(dA, A.data(), sizeof(float) * M * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), sizeof(float) * N * K, cudaMemcpyHostToDevice);
    size_t smem = sizeof(float) * (2 * M * K + 2 * N * K);
    cudaFuncSetAttribute(gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int fails = 0;
    for (int split = 0; split < 2; ++split) {
        cudaMemset(dD, 0, sizeof(float) * M * N);
        gemm_test_kernel<<<1, 128, smem>>>(dA, dB, dD, split);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("split=%d CUDA error: %s\n", split, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
        double max_err = 0, max_ref = 0; int bad_at = -1;
        for (int i = 0; i < M * N; ++i) { double err = fabs(D[i] - ref[i]); if (err > max_err) { max_err = err; bad_at = i; } max_ref = fmax(max_ref, fabs(ref[i])); }
        printf("%s: max |err| = %.3e (max |ref| = %.3f) at (%d,%d): got %.7f want %.7f\n", split ? "3xTF32" : "1xTF32", max_err, max_ref,
               bad_at / N, bad_at % N, D[bad_at], ref[bad_at]);
        double tol = split ? 2e-5 : 2e-2;
        if (!(max_err < tol)) { ++fails; printf("  FAIL (tol %.1e)\n", tol); }
    }
    printf(fails ? "TC GEMM TEST FAILED\n" : "TC GEMM TEST OK\n");
    return fails;
}
