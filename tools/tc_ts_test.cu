// Unit test of the TS building blocks (elegantrl_b200/csrc/ts_mlp.cuh) on ONE tile of 128 rows:
//   layer 1 on the tensor core (x~ with folded bias, 3xTF32) -> in-place fp16 {hi, lo} conversion in tensor memory ->
//   layer 2 (bias UMMA + A-from-TMEM kind::f16 UMMAs) -> head, against an fp64 CPU evaluation of the same MLP.
//   mode 0: no activation (tests the UMMA chain alone, tight tolerance);  mode 1: GELU (the real net).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o /tmp/ts tools/tc_ts_test.cu && /tmp/ts
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../elegantrl_b200/csrc/ts_mlp.cuh"

void b200rl_set_error(const char*, ...) {}
long long g_b200rl_launches = 0;

using namespace tsmlp;

constexpr int kOffB2 = 0, kOffB1 = 2 * kB2PlaneBytes, kOffBb = kOffB1 + 2 * kB1Bytes, kOffAc = kOffBb + kB1Bytes,
              kOffA1 = kOffAc + kA1Bytes, kOffW3 = kOffA1 + kA1Bytes, kSmem = kOffW3 + 64 * 4 + 64;

template <bool GELU>
__global__ void __launch_bounds__(160) ts_test_kernel(b200rl_net net, const float* x, float* z2_out, float* head_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[7];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    NetImage im;
    im.b2[0] = kOffB2; im.b2[1] = kOffB2 + kB2PlaneBytes; im.b1[0] = kOffB1; im.b1[1] = kOffB1 + kB1Bytes; im.bb = kOffBb;
    if (warp == 0) tc05::tmem_alloc<128>(&tmem_slot);
    if (tid == 32) {
        tc05::mbar_init(&bars[0], 4); tc05::mbar_init(&bars[1], 1); tc05::mbar_init(&bars[2], 1);
        for (int c = 0; c < 4; ++c) tc05::mbar_init(&bars[3 + c], 4);
        tc05::mbar_fence_init();
    }
    stage_net(net, smem, im, tid, 160);
    stage_const_a(smem + kOffAc, tid, 160);
    float* w3 = reinterpret_cast<float*>(smem + kOffW3);
    if (tid < 64) w3[tid] = net.weight[2][tid];
    tc05::fence_proxy_async_smem();
    tc05::fence_before_thread_sync();
    __syncthreads();
    tc05::fence_after_thread_sync();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 4) {
        const uint32_t tX = tmem_base, tD = tmem_base + 64;
        const NetDescs nd = make_descs(smem, im);
        const uint64_t ac = tc05::make_smem_desc(tc05::smem_u32(smem + kOffAc), kSboK8);
        const uint64_t a1 = tc05::make_smem_desc(tc05::smem_u32(smem + kOffA1), kSboK8);
        tc05::mbar_wait(&bars[0], 0);
        tc05::fence_after_thread_sync();
        if (tc05::elect_one()) { issue_layer1(tX, a1, nd); tc05::mma_commit(&bars[1]); }
        __syncwarp();
        for (int c = 0; c < 4; ++c) {
            tc05::mbar_wait(&bars[3 + c], 0);
            tc05::fence_after_thread_sync();
            if (tc05::elect_one()) {
                if (c == 0) issue_bias(tD, ac, nd);
                issue_layer2_chunk(tD, tX, nd, c);
                if (c == 3) tc05::mma_commit(&bars[2]);
            }
            __syncwarp();
        }
    } else {
        const int row = tid;
        const uint32_t tX = tmem_base + ((uint32_t)(warp * 32) << 16), tD = tX + 64;
        const float xr[3] = {x[row * 3], x[row * 3 + 1], x[row * 3 + 2]};
        write_x_row(smem + kOffA1, row, xr);
        tc05::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc05::mbar_arrive(&bars[0]);
        tc05::mbar_wait(&bars[1], 0);
        tc05::fence_after_thread_sync();
        for (int c = 0; c < 4; ++c) {
            hidden_chunk_inplace<GELU>(tX + 16 * c);
            tc05::tmem_st_wait();
            tc05::fence_before_thread_sync();
            __syncwarp();
            if (lane == 0) tc05::mbar_arrive(&bars[3 + c]);
        }
        tc05::mbar_wait(&bars[2], 0);
        tc05::fence_after_thread_sync();
        for (int c = 0; c < 4; ++c) {
            float v[16];
            tc05::tmem_ld_32x32b_x16(tD + 16 * c, v);
            tc05::tmem_ld_wait();
            for (int j = 0; j < 16; ++j) z2_out[row * 64 + 16 * c + j] = v[j];
        }
        head_out[row] = head_dot<GELU>(tD, w3, net.bias[2][0]);
    }
    tc05::fence_before_thread_sync();
    __syncthreads();
    if (warp == 0) tc05::tmem_dealloc<128>(tmem_base);
}

static double gelu(double v) { return 0.5 * v * (1.0 + erf(v / sqrt(2.0))); }

int main() {
    srand(3);
    auto rnd = [](float s) { return ((rand() / (float)RAND_MAX) * 2.f - 1.f) * s; };
    std::vector<float> W1(64 * 3), b1(64), W2(64 * 64), b2(64), W3(64), b3(1), X(128 * 3);
    for (auto& v : W1) v = rnd(1.0f);
    for (auto& v : b1) v = rnd(0.5f);
    for (auto& v : W2) v = rnd(0.3f);
    for (auto& v : b2) v = rnd(0.5f);
    for (auto& v : W3) v = rnd(0.3f);
    b3[0] = 0.1f;
    for (int r = 0; r < 128; ++r) { X[r * 3] = rnd(1.f); X[r * 3 + 1] = rnd(1.f); X[r * 3 + 2] = rnd(8.f); }
    float *dW1, *db1, *dW2, *db2, *dW3, *db3, *dX, *dZ, *dH;
    auto up = [](float** d, const std::vector<float>& h) { cudaMalloc(d, h.size() * 4); cudaMemcpy(*d, h.data(), h.size() * 4, cudaMemcpyHostToDevice); };
    up(&dW1, W1); up(&db1, b1); up(&dW2, W2); up(&db2, b2); up(&dW3, W3); up(&db3, b3); up(&dX, X);
    cudaMalloc(&dZ, 128 * 64 * 4); cudaMalloc(&dH, 128 * 4);
    b200rl_net net{};
    net.num_linear = 3; net.dims[0] = 3; net.dims[1] = 64; net.dims[2] = 64; net.dims[3] = 1;
    net.weight[0] = dW1; net.bias[0] = db1; net.weight[1] = dW2; net.bias[1] = db2; net.weight[2] = dW3; net.bias[2] = db3;
    int fails = 0;
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dZ, 0, 128 * 64 * 4); cudaMemset(dH, 0, 128 * 4);
        if (mode == 0) {
            cudaFuncSetAttribute(ts_test_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
            ts_test_kernel<false><<<1, 160, kSmem>>>(net, dX, dZ, dH);
        } else {
            cudaFuncSetAttribute(ts_test_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
            ts_test_kernel<true><<<1, 160, kSmem>>>(net, dX, dZ, dH);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d CUDA error: %s\n", mode, cudaGetErrorString(e)); return 1; }
        std::vector<float> Z(128 * 64), Hd(128);
        cudaMemcpy(Z.data(), dZ, Z.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(Hd.data(), dH, Hd.size() * 4, cudaMemcpyDeviceToHost);
        double ez = 0, eh = 0, mz = 0, mh = 0;
        for (int r = 0; r < 128; ++r) {
            double h1[64], z2[64], out = b3[0];
            for (int j = 0; j < 64; ++j) {
                double s = b1[j];
                for (int k = 0; k < 3; ++k) s += (double)W1[j * 3 + k] * X[r * 3 + k];
                h1[j] = mode ? gelu(s) : s;
            }
            for (int n = 0; n < 64; ++n) {
                double s = b2[n];
                for (int j = 0; j < 64; ++j) s += (double)W2[n * 64 + j] * h1[j];
                z2[n] = s;
                out += (mode ? gelu(s) : s) * W3[n];
                ez = fmax(ez, fabs(Z[r * 64 + n] - s)); mz = fmax(mz, fabs(s));
            }
            eh = fmax(eh, fabs(Hd[r] - out)); mh = fmax(mh, fabs(out));
        }
        printf("mode %d (%s): Z2 max|err| %.3e (max|ref| %.3f)   head max|err| %.3e (max|ref| %.3f)\n", mode, mode ? "GELU" : "linear", ez, mz, eh, mh);
        if (!(ez < 2e-5 * fmax(1.0, mz)) || !(eh < 2e-5 * fmax(1.0, mh))) { ++fails; printf("  FAIL\n"); }
    }
    printf(fails ? "TS MLP TEST FAILED\n" : "TS MLP TEST OK\n");
    return fails;
}
