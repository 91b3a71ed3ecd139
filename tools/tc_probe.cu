// tc_probe.cu -- feasibility probe for a tcgen05 3xTF32 GEMM on sm_100a (NOT part of the product path).
//
// Question it answers on a real B200: do hand-built UMMA shared-memory / instruction descriptors, TMEM allocation,
// tcgen05.mma.kind::tf32 and tcgen05.ld work the way this repo assumes, and how accurate is the 3xTF32 split
// (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, fp32 accumulation in TMEM) against an fp64 reference -- i.e. could the DFNet
// GEMM chain move to the 5th-gen tensor cores without breaking the 1e-5 parity bar?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tc_probe tools/tc_probe.cu && gpurun_out/tc_probe
//
// D[128 x 256] = A[128 x K] * B[256 x K]^T, one CTA, operands staged into shared memory by plain loads in the
// canonical K-major SWIZZLE_128B layout (no TMA, to keep the number of unknowns small).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 256, KB = 32;   // K block = 32 tf32 = 128 bytes = one swizzle row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B canonical layout: 8-row x 128-byte atoms (1024 B), 16-byte chunk c of row r stored at c ^ (r & 7)
__device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
    const int atom = row >> 3, r = row & 7, chunk = k >> 2;
    return atom * 1024 + r * 128 + (((chunk ^ r) & 7) << 4) + ((k & 3) << 2);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)1 << 16;                               // leading byte offset (ignored for swizzled K-major), 16 B
    d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
    return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ Ahi, const float* __restrict__ Alo,
                                                       const float* __restrict__ Bhi, const float* __restrict__ Blo, float* __restrict__ D,
                                                       int K, int terms) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sAhi = smem;                       // 128 rows * 128 B = 16 KB
    uint8_t* sAlo = smem + 16384;
    uint8_t* sBhi = smem + 32768;               // 256 rows * 128 B = 32 KB
    uint8_t* sBlo = smem + 65536;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;

    // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

    uint32_t phase = 0;
    for (int kb = 0; kb < K / KB; ++kb) {
        for (int idx = tid; idx < M * KB; idx += 128) {
            const int r = idx / KB, k = idx % KB;
            *reinterpret_cast<float*>(sAhi + sw128_offset(r, k)) = Ahi[(size_t)r * K + kb * KB + k];
            *reinterpret_cast<float*>(sAlo + sw128_offset(r, k)) = Alo[(size_t)r * K + kb * KB + k];
        }
        for (int idx = tid; idx < N * KB; idx += 128) {
            const int r = idx / KB, k = idx % KB;
            *reinterpret_cast<float*>(sBhi + sw128_offset(r, k)) = Bhi[(size_t)r * K + kb * KB + k];
            *reinterpret_cast<float*>(sBlo + sw128_offset(r, k)) = Blo[(size_t)r * K + kb * KB + k];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (int t = 0; t < terms; ++t) {                   // 0: hi*hi   1: lo*hi   2: hi*lo
                const uint8_t* a = (t == 1) ? sAlo : sAhi;
                const uint8_t* b = (t == 2) ? sBlo : sBhi;
                for (int k = 0; k < KB / 8; ++k) {              // UMMA_K = 8 for tf32 (32 bytes)
                    const uint64_t adesc = make_desc(smem_u32(a) + k * 32);
                    const uint64_t bdesc = make_desc(smem_u32(b) + k * 32);
                    const uint32_t accum = (kb > 0 || t > 0 || k > 0) ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
                        : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        // everyone waits until the MMAs of this K block have consumed the shared-memory operands
        {
            uint32_t ok = 0, spins = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
                if (++spins > (1u << 24)) __trap();
            }
            phase ^= 1;
        }
        __syncthreads();
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // epilogue: warp w owns TMEM lanes [32w, 32w+32) = rows of D; 32 columns per tcgen05.ld
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
              "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
              "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = warp * 32 + (tid & 31);
        for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

static float to_tf32(float x) {   // round to nearest (ties away), keep 10 mantissa bits
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x1000u; u &= 0xFFFFE000u;
    float y; memcpy(&y, &u, 4); return y;
}

int main() {
    const int K = 512;
    std::vector<float> A(M * K), B(N * K), Ahi(M * K), Alo(M * K), Bhi(N * K), Blo(N * K);
    srand(1);
    for (auto& v : A) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (auto& v : B) v = ((rand() / (float)RAND_MAX) * 2.f - 1.f) * 0.07f;
    for (int i = 0; i < M * K; ++i) { Ahi[i] = to_tf32(A[i]); Alo[i] = to_tf32(A[i] - Ahi[i]); }
    for (int i = 0; i < N * K; ++i) { Bhi[i] = to_tf32(B[i]); Blo[i] = to_tf32(B[i] - Bhi[i]); }
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * (double)B[n * K + k];
            ref[m * N + n] = s;
        }
    float *dAhi, *dAlo, *dBhi, *dBlo, *dD;
    cudaMalloc(&dAhi, M * K * 4); cudaMalloc(&dAlo, M * K * 4); cudaMalloc(&dBhi, N * K * 4); cudaMalloc(&dBlo, N * K * 4); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dAhi, Ahi.data(), M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dAlo, Alo.data(), M * K * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dBhi, Bhi.data(), N * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dBlo, Blo.data(), N * K * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024);
    double scale = 0;
    for (auto v : ref) scale = std::max(scale, std::fabs(v));
    // fp32 FMA reference error for context
    double e32 = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0;
            for (int k = 0; k < K; ++k) s = fmaf(A[m * K + k], B[n * K + k], s);
            e32 = std::max(e32, std::fabs((double)s - ref[m * N + n]));
        }
    printf("max|ref| %.4f   plain fp32 FMA chain: max abs err %.3e (%.3e of scale)\n", scale, e32, e32 / scale);
    for (int terms = 1; terms <= 3; terms += 2) {
        cudaMemset(dD, 0, M * N * 4);
        probe_kernel<<<1, 128, 98304 + 1024>>>(dAhi, dAlo, dBhi, dBlo, dD, K, terms);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("terms=%d: kernel failed: %s\n", terms, cudaGetErrorString(e)); return 1; }
        std::vector<float> D(M * N);
        cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        double emax = 0; int bad = 0;
        for (int i = 0; i < M * N; ++i) { double er = std::fabs((double)D[i] - ref[i]); emax = std::max(emax, er); if (er > 1e-2 * scale) ++bad; }
        printf("tcgen05 %dxTF32: max abs err %.3e (%.3e of scale), grossly wrong elements %d, D[0]=%.6f ref %.6f, D[last]=%.6f ref %.6f\n",
               terms, emax, emax / scale, bad, D[0], ref[0], D[M * N - 1], ref[M * N - 1]);
    }
    return 0;
}
