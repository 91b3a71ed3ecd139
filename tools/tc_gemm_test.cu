// tc_gemm_test.cu -- correctness + throughput of the 3xTF32 tcgen05 GEMM building block (posendf_b200/csrc/pndf_tc_gemm.cuh) on DFNet
// layer shapes.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/tc_gemm_test tools/tc_gemm_test.cu
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../posendf_b200/csrc/pndf_tc_gemm.cuh"

using namespace pndf_tc;

struct StoreEpi {
    float* D;
    int ldd;
    __device__ void operator()(int row, int col0, float (&v)[32]) const {
        float4* dst = reinterpret_cast<float4*>(D + (size_t)row * ldd + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
};

static float to_tf32(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x1000u; u &= 0xFFFFE000u;
    float y; memcpy(&y, &u, 4); return y;
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int NT>
static void run_case(int M, int N, int K) {
    std::vector<float> A((size_t)M * K), B((size_t)N * K), Ahi(A.size()), Alo(A.size()), Bhi(B.size()), Blo(B.size());
    srand(7);
    for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (auto& v : B) v = ((rand() / (float)RAND_MAX) * 2.f - 1.f) * 0.07f;
    for (size_t i = 0; i < A.size(); ++i) { Ahi[i] = to_tf32(A[i]); Alo[i] = to_tf32(A[i] - Ahi[i]); }
    for (size_t i = 0; i < B.size(); ++i) { Bhi[i] = to_tf32(B[i]); Blo[i] = to_tf32(B[i] - Bhi[i]); }
    float *dAhi, *dAlo, *dBhi, *dBlo, *dD;
    CK(cudaMalloc(&dAhi, A.size() * 4)); CK(cudaMalloc(&dAlo, A.size() * 4)); CK(cudaMalloc(&dBhi, B.size() * 4)); CK(cudaMalloc(&dBlo, B.size() * 4));
    CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dAhi, Ahi.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dAlo, Alo.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBhi, Bhi.data(), B.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBlo, Blo.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, (size_t)M * N * 4));
    GemmMaps maps;
    if (!make_map(&maps.a_hi, dAhi, M, K, K, kTM) || !make_map(&maps.a_lo, dAlo, M, K, K, kTM) || !make_map(&maps.b_hi, dBhi, N, K, K, NT) ||
        !make_map(&maps.b_lo, dBlo, N, K, K, NT)) { printf("tensor map creation failed\n"); exit(1); }
    auto kern = tc_gemm_kernel<NT, StoreEpi>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<NT>()));
    StoreEpi epi{dD, N};
    dim3 grid(N / NT, M / kTM);
    kern<<<grid, kThreads, smem_bytes<NT>()>>>(maps, K, epi);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)M * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double emax = 0, scale = 0, e32 = 0;
    for (int m = 0; m < M; m += std::max(1, M / 97)) {
        for (int n = 0; n < N; ++n) {
            double s = 0; float f = 0.f;
            for (int k = 0; k < K; ++k) { s += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k]; f = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], f); }
            emax = std::max(emax, std::fabs((double)D[(size_t)m * N + n] - s));
            e32 = std::max(e32, std::fabs((double)f - s));
            scale = std::max(scale, std::fabs(s));
        }
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 20;
    for (int i = 0; i < 3; ++i) kern<<<grid, kThreads, smem_bytes<NT>()>>>(maps, K, epi);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) kern<<<grid, kThreads, smem_bytes<NT>()>>>(maps, K, epi);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double flop = 2.0 * M * N * K;
    printf("M %6d N %5d K %5d NT %3d: max err %.3e of scale %.3f (= %.2e rel; fp32 FMA chain %.2e)   %.3f ms  %.1f TFLOP/s fp32-equivalent (%.0f tf32 TFLOP/s)\n",
           M, N, K, NT, emax, scale, emax / scale, e32 / scale, ms, flop / ms / 1e9, 3 * flop / ms / 1e9);
    cudaFree(dAhi); cudaFree(dAlo); cudaFree(dBhi); cudaFree(dBlo); cudaFree(dD);
}

int main() {
    run_case<128>(1024, 256, 128);
    run_case<128>(8192, 512, 1024);
    run_case<64>(8192, 64, 256);
    run_case<128>(65536, 1024, 512);
    run_case<128>(65536, 512, 1024);
    run_case<128>(65536, 256, 512);
    return 0;
}
