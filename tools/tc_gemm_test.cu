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
    static constexpr int kOutputs = 1;
    float* D;
    int ldd;
    __device__ float* out(int) const { return D; }
    __device__ int ld() const { return ldd; }
    __device__ void init4(int, float* r) const { r[0] = r[1] = r[2] = r[3] = 0.0f; }
    __device__ void operator()(int, int, const float (&v)[32], float (&o0)[32], float (&)[32]) const {
#pragma unroll
        for (int j = 0; j < 32; ++j) o0[j] = v[j];
    }
};

// the product's forward epilogue shape: bias + leaky relu + tf32 hi / lo split, TWO output arrays
struct SplitEpi {
    static constexpr int kOutputs = 2;
    const float* bias;
    float* hi_;
    float* lo_;
    int ldd;
    __device__ float* out(int w) const { return w == 0 ? hi_ : lo_; }
    __device__ int ld() const { return ldd; }
    __device__ void init4(int, float* r) const { r[0] = r[1] = r[2] = r[3] = 0.0f; }
    __device__ void operator()(int, int col0, const float (&v)[32], float (&hi)[32], float (&lo)[32]) const {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float x = v[j] + __ldg(bias + col0 + j);
            const float z = x > 0.0f ? x : 0.01f * x;
            uint32_t u;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(z));
            hi[j] = __uint_as_float(u);
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(z - hi[j]));
            lo[j] = __uint_as_float(u);
        }
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
    // operands in the kernel's TILED layout (pndf_tc_gemm.cuh::tiled_offset)
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
        const size_t t = tiled_offset(m, k, K, kTM); const float x = A[(size_t)m * K + k];
        Ahi[t] = to_tf32(x); Alo[t] = to_tf32(x - Ahi[t]);
    }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
        const size_t t = tiled_offset(n, k, K, NT); const float x = B[(size_t)n * K + k];
        Bhi[t] = to_tf32(x); Blo[t] = to_tf32(x - Bhi[t]);
    }
    float *dAhi, *dAlo, *dBhi, *dBlo, *dD;
    CK(cudaMalloc(&dAhi, A.size() * 4)); CK(cudaMalloc(&dAlo, A.size() * 4)); CK(cudaMalloc(&dBhi, B.size() * 4)); CK(cudaMalloc(&dBlo, B.size() * 4));
    CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dAhi, Ahi.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dAlo, Alo.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBhi, Bhi.data(), B.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBlo, Blo.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, (size_t)M * N * 4));
    GemmMaps maps;
    if (!make_map(&maps.a_hi, dAhi, M, K, kTM) || !make_map(&maps.a_lo, dAlo, M, K, kTM) || !make_map(&maps.b_hi, dBhi, N, K, NT) ||
        !make_map(&maps.b_lo, dBlo, N, K, NT) || !make_out_map(&maps.out[0], dD, M, N)) { printf("tensor map creation failed\n"); exit(1); }
    auto kern = tc_gemm_kernel<NT, StoreEpi>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<NT>()));
    StoreEpi epi{dD, N};
    const int m_tiles = M / kTM, n_tiles = N / NT;
    const int grid = std::min(m_tiles * n_tiles, 148);
    kern<<<grid, kThreads, smem_bytes<NT>()>>>(maps, K, m_tiles, n_tiles, epi);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)M * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double emax = 0, scale = 0, e32 = 0;
    for (int m = 0; m < M; m += std::max(1, M / 97)) {
        for (int n = 0; n < N; ++n) {
            double s = 0; float f = 0.f;
            for (int k = 0; k < K; ++k) { s += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k]; f = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], f); }
            emax = std::max(emax, std::fabs((double)D[tiled_offset(m, n, N, kTM)] - s));
            e32 = std::max(e32, std::fabs((double)f - s));
            scale = std::max(scale, std::fabs(s));
        }
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 20;
    for (int i = 0; i < 3; ++i) kern<<<grid, kThreads, smem_bytes<NT>()>>>(maps, K, m_tiles, n_tiles, epi);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) kern<<<grid, kThreads, smem_bytes<NT>()>>>(maps, K, m_tiles, n_tiles, epi);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double flop = 2.0 * M * N * K;
    {   // same GEMM with the two-output epilogue (timing only)
        float *dHi, *dLo, *dBias;
        CK(cudaMalloc(&dHi, (size_t)M * N * 4)); CK(cudaMalloc(&dLo, (size_t)M * N * 4)); CK(cudaMalloc(&dBias, N * 4)); CK(cudaMemset(dBias, 0, N * 4));
        auto k2 = tc_gemm_kernel<NT, SplitEpi>;
        CK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<NT>()));
        SplitEpi e2{dBias, dHi, dLo, N};
        GemmMaps maps2 = maps;
        if (!make_out_map(&maps2.out[0], dHi, M, N) || !make_out_map(&maps2.out[1], dLo, M, N)) { printf("tensor map creation failed\n"); exit(1); }
        for (int i = 0; i < 3; ++i) k2<<<grid, kThreads, smem_bytes<NT>()>>>(maps2, K, m_tiles, n_tiles, e2);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) k2<<<grid, kThreads, smem_bytes<NT>()>>>(maps2, K, m_tiles, n_tiles, e2);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms2; cudaEventElapsedTime(&ms2, e0, e1); ms2 /= reps;
        printf("   two-output epilogue (bias + lrelu + hi/lo): %.3f ms  %.1f TFLOP/s fp32-equivalent\n", ms2, 2.0 * M * N * K / ms2 / 1e9);
        cudaFree(dHi); cudaFree(dLo); cudaFree(dBias);
    }
    if (NT == 128 && M % 256 == 0) {      // the same GEMM on CTA pairs (cta_group::2)
        GemmMaps mp = maps;
        if (!make_map(&mp.b_x, dBhi, N, K, 64)) { printf("tensor map creation failed\n"); exit(1); }
        auto kp = tc_gemm_kernel<128, StoreEpi, true>;
        CK(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, pair_smem_bytes()));
        CK(cudaMemset(dD, 0, (size_t)M * N * 4));
        const int gp = std::min((m_tiles / 2) * n_tiles * 2, 148);
        CK(launch_pair(kp, gp, pair_smem_bytes(), 0, mp, K, m_tiles, n_tiles, epi));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double ep = 0;
        for (int m = 0; m < M; m += std::max(1, M / 97))
            for (int n = 0; n < N; ++n) {
                double sd = 0;
                for (int k = 0; k < K; ++k) sd += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
                ep = std::max(ep, std::fabs((double)D[tiled_offset(m, n, N, kTM)] - sd));
            }
        for (int i = 0; i < 3; ++i) launch_pair(kp, gp, pair_smem_bytes(), 0, mp, K, m_tiles, n_tiles, epi);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) launch_pair(kp, gp, pair_smem_bytes(), 0, mp, K, m_tiles, n_tiles, epi);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float msp; cudaEventElapsedTime(&msp, e0, e1); msp /= reps;
        printf("   CTA pairs (cta_group::2, 256 x 128 tiles): max err %.3e (= %.2e rel)   %.3f ms  %.1f TFLOP/s fp32-equivalent\n", ep, ep / scale, msp,
               flop / msp / 1e9);
    }
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
