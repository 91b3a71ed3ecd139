// tc_chain_probe.cu -- measured answer to VERDICT r1 item 10 (NOT part of the product path): would the DFNet GEMM chain hold the
// 1e-5 parity bar on the 5th-gen tensor cores with a 3xTF32 split?
//
// Runs the WHOLE DFNet of a synthetic amass.yaml network (126 -> 256 -> 512 -> 1024 -> 512 -> 256 -> 64 -> 1) forward and the
// input-gradient chain backward for 128 poses, every layer as tcgen05.mma.kind::tf32 GEMMs (hand-built K-major SWIZZLE_128B
// descriptors, TMEM accumulators, same machinery as tools/tc_probe.cu), with bias / activation / hi-lo splitting between the
// layers done on the host in fp32 -- a numerical experiment, not a fast kernel.  Variants:
//   1xTF32                       A_hi B_hi
//   3xTF32                       A_hi B_hi + A_lo B_hi + A_hi B_lo            one TMEM accumulator
//   3xTF32 split accumulators    (A_hi B_hi) in one accumulator, (A_lo B_hi + A_hi B_lo) in a second, added in fp32 (RN)
//   3xTF32 split + K-chunks      as above, and a fresh accumulator pair per 128-deep K chunk, chunks added in fp32 (RN)
// and reports, against an fp64 evaluation of the same network: relative error of d and norm-wise relative error of dd/dz0, next
// to the same numbers for a plain fp32 FMA chain (what the product kernel does).
//
//   python tools/tc_chain_export.py lrelu /tmp/chain_lrelu.bin
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tc_chain_probe tools/tc_chain_probe.cu
//   gpurun_out/tc_chain_probe /tmp/chain_lrelu.bin lrelu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, KB = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
    const int atom = row >> 3, r = row & 7, chunk = k >> 2;
    return atom * 1024 + r * 128 + (((chunk ^ r) & 7) << 4) + ((k & 3) << 2);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// D[128 x N] (+)= A[128 x K] * B[N x K]^T for one N tile (N = 64 or 256).  variant: 0 = 1xTF32, 1 = 3xTF32 one accumulator,
// 2 = split accumulators, 3 = split accumulators + a fresh pair per `kchunk`-deep K chunk.
__global__ void __launch_bounds__(128, 1) gemm_kernel(const float* __restrict__ Ahi, const float* __restrict__ Alo,
                                                      const float* __restrict__ Bhi, const float* __restrict__ Blo, float* __restrict__ D,
                                                      int K, int N, int ldd, int variant, int kchunk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sAhi = smem;
    uint8_t* sAlo = smem + 16384;
    uint8_t* sBhi = smem + 32768;
    uint8_t* sBlo = smem + 65536;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const float* Bh = Bhi + (size_t)blockIdx.x * N * K;
    const float* Bl = Blo + (size_t)blockIdx.x * N * K;
    float* Dt = D + (size_t)blockIdx.x * N;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const int nkb = K / KB, kb_per_chunk = (variant == 3) ? kchunk / KB : nkb;
    // fp32 running sums of finished accumulator pairs (variants 2, 3): thread t owns row t, columns in registers are too many ->
    // accumulate straight into global D (zero-initialised by the host)
    uint32_t phase = 0;
    for (int kb0 = 0; kb0 < nkb; kb0 += kb_per_chunk) {
        for (int kb = kb0; kb < min(nkb, kb0 + kb_per_chunk); ++kb) {
            for (int idx = tid; idx < M * KB; idx += 128) {
                const int r = idx / KB, k = idx % KB;
                *reinterpret_cast<float*>(sAhi + sw128_offset(r, k)) = Ahi[(size_t)r * K + kb * KB + k];
                *reinterpret_cast<float*>(sAlo + sw128_offset(r, k)) = Alo[(size_t)r * K + kb * KB + k];
            }
            for (int idx = tid; idx < N * KB; idx += 128) {
                const int r = idx / KB, k = idx % KB;
                *reinterpret_cast<float*>(sBhi + sw128_offset(r, k)) = Bh[(size_t)r * K + kb * KB + k];
                *reinterpret_cast<float*>(sBlo + sw128_offset(r, k)) = Bl[(size_t)r * K + kb * KB + k];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
                const int terms = (variant == 0) ? 1 : 3;
                for (int t = 0; t < terms; ++t) {
                    const uint8_t* a = (t == 1) ? sAlo : sAhi;
                    const uint8_t* b = (t == 2) ? sBlo : sBhi;
                    const uint32_t acc_col = (variant >= 2 && t > 0) ? 256u : 0u;      // second accumulator for the cross terms
                    for (int k = 0; k < KB / 8; ++k) {
                        const uint64_t adesc = make_desc(smem_u32(a) + k * 32);
                        const uint64_t bdesc = make_desc(smem_u32(b) + k * 32);
                        const bool first = (kb == kb0) && (k == 0) && (variant >= 2 ? (t == 0 || t == 1) : (t == 0));
                        const uint32_t accum = first ? 0u : 1u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(tmem + acc_col), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
                            : "memory");
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
            }
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
        // drain this chunk's accumulator(s) into D (fp32 adds, round to nearest)
        asm volatile("tcgen05.fence::after_thread_sync;");
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32], w[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
#define TLD(arr, addr)                                                                                                            \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                       \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                      \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                     \
                 : "=r"(arr[0]), "=r"(arr[1]), "=r"(arr[2]), "=r"(arr[3]), "=r"(arr[4]), "=r"(arr[5]), "=r"(arr[6]), "=r"(arr[7]),   \
                   "=r"(arr[8]), "=r"(arr[9]), "=r"(arr[10]), "=r"(arr[11]), "=r"(arr[12]), "=r"(arr[13]), "=r"(arr[14]),            \
                   "=r"(arr[15]), "=r"(arr[16]), "=r"(arr[17]), "=r"(arr[18]), "=r"(arr[19]), "=r"(arr[20]), "=r"(arr[21]),          \
                   "=r"(arr[22]), "=r"(arr[23]), "=r"(arr[24]), "=r"(arr[25]), "=r"(arr[26]), "=r"(arr[27]), "=r"(arr[28]),          \
                   "=r"(arr[29]), "=r"(arr[30]), "=r"(arr[31])                                                                       \
                 : "r"(addr))
            TLD(v, taddr);
            if (variant >= 2) TLD(w, taddr + 256u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = warp * 32 + (tid & 31);
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(v[j]);
                if (variant >= 2) x += __uint_as_float(w[j]);
                float* dst = Dt + (size_t)row * ldd + c0 + j;
                *dst = (kb0 == 0) ? x : (*dst + x);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static float to_tf32(float x) {
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x1000u; u &= 0xFFFFE000u;
    float y; memcpy(&y, &u, 4); return y;
}
static void split(const std::vector<float>& x, std::vector<float>& hi, std::vector<float>& lo) {
    hi.resize(x.size()); lo.resize(x.size());
    for (size_t i = 0; i < x.size(); ++i) { hi[i] = to_tf32(x[i]); lo[i] = to_tf32(x[i] - hi[i]); }
}

struct Dev {
    float *Ahi, *Alo, *Bhi, *Blo, *D;
};
// C[128 x N] = A[128 x K] * B[N x K]^T on the tensor cores (N a multiple of 64, K of 32)
static void tc_gemm(Dev& d, const std::vector<float>& A, const std::vector<float>& B, int N, int K, std::vector<float>& C, int variant, int kchunk) {
    std::vector<float> Ahi, Alo, Bhi, Blo;
    split(A, Ahi, Alo); split(B, Bhi, Blo);
    cudaMemcpy(d.Ahi, Ahi.data(), Ahi.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(d.Alo, Alo.data(), Alo.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d.Bhi, Bhi.data(), Bhi.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(d.Blo, Blo.data(), Blo.size() * 4, cudaMemcpyHostToDevice);
    const int NT = (N % 256 == 0) ? 256 : 64;
    gemm_kernel<<<N / NT, 128, 98304 + 1024>>>(d.Ahi, d.Alo, d.Bhi, d.Blo, d.D, K, NT, N, variant, kchunk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); exit(1); }
    C.resize((size_t)M * N);
    cudaMemcpy(C.data(), d.D, C.size() * 4, cudaMemcpyDeviceToHost);
}
static void ffma_gemm(const std::vector<float>& A, const std::vector<float>& B, int N, int K, std::vector<float>& C) {
    C.assign((size_t)M * N, 0.f);
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float s = 0.f;
            for (int k = 0; k < K; ++k) s = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], s);
            C[(size_t)m * N + n] = s;
        }
}

int main(int argc, char** argv) {
    if (argc < 3) { printf("usage: tc_chain_probe <file.bin> <lrelu|softplus>\n"); return 1; }
    const bool soft = std::string(argv[2]) == "softplus";
    const float beta = 100.f;
    const int W[8] = {126, 256, 512, 1024, 512, 256, 64, 1};
    int KP[7];
    std::vector<std::vector<float>> Wt(7), bias(7);
    FILE* f = fopen(argv[1], "rb");
    if (!f) { printf("cannot open %s\n", argv[1]); return 1; }
    for (int l = 0; l < 7; ++l) {
        KP[l] = (W[l] + 31) / 32 * 32;
        Wt[l].resize((size_t)W[l + 1] * KP[l]); bias[l].resize(W[l + 1]);
        if (fread(Wt[l].data(), 4, Wt[l].size(), f) != Wt[l].size() || fread(bias[l].data(), 4, bias[l].size(), f) != bias[l].size()) return 2;
    }
    std::vector<float> z0((size_t)M * 128);
    if (fread(z0.data(), 4, z0.size(), f) != z0.size()) return 2;
    fclose(f);

    auto actf = [&](double x, double& dv) {
        if (soft) { double bx = beta * x; if (bx > 20) { dv = 1; return x; } double e = exp(bx); dv = e / (1 + e); return log1p(e) / beta; }
        dv = x > 0 ? 1.0 : 0.01; return x > 0 ? x : 0.01 * x;
    };
    // ---- fp64 reference: d and dd/dz0
    std::vector<double> dref(M), gref((size_t)M * 126);
    std::vector<std::vector<double>> dphi64(6, std::vector<double>());
    {
        std::vector<double> z(z0.begin(), z0.end());
        int kp = 128;
        std::vector<std::vector<double>> dph(6);
        for (int l = 0; l < 6; ++l) {
            std::vector<double> zn((size_t)M * KP[l + 1 < 7 ? l + 1 : 6] , 0.0);
            const int np = (l + 1 < 7) ? ((W[l + 1] + 31) / 32 * 32) : 0;
            zn.assign((size_t)M * np, 0.0); dph[l].assign((size_t)M * W[l + 1], 0.0);
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < W[l + 1]; ++n) {
                    double s = bias[l][n];
                    for (int k = 0; k < W[l]; ++k) s += z[(size_t)m * kp + k] * (double)Wt[l][(size_t)n * KP[l] + k];
                    double dv; zn[(size_t)m * np + n] = actf(s, dv); dph[l][(size_t)m * W[l + 1] + n] = dv;
                }
            z.swap(zn); kp = np;
        }
        std::vector<double> gs(M);
        for (int m = 0; m < M; ++m) {
            double s = bias[6][0];
            for (int k = 0; k < 64; ++k) s += z[(size_t)m * kp + k] * (double)Wt[6][k];
            double dv; if (soft) dref[m] = actf(s, dv); else { dref[m] = s > 0 ? s : 0; dv = s > 0 ? 1 : 0; }
            gs[m] = dv;
        }
        std::vector<double> g((size_t)M * 64);
        for (int m = 0; m < M; ++m) for (int k = 0; k < 64; ++k) g[(size_t)m * 64 + k] = gs[m] * (double)Wt[6][k];
        for (int l = 5; l >= 0; --l) {
            std::vector<double> gn((size_t)M * W[l], 0.0);
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < W[l + 1]; ++n) {
                    const double t = g[(size_t)m * W[l + 1] + n] * dph[l][(size_t)m * W[l + 1] + n];
                    for (int k = 0; k < W[l]; ++k) gn[(size_t)m * W[l] + k] += t * (double)Wt[l][(size_t)n * KP[l] + k];
                }
            g.swap(gn);
        }
        gref = g;
    }
    Dev d;
    cudaMalloc(&d.Ahi, (size_t)M * 1024 * 4); cudaMalloc(&d.Alo, (size_t)M * 1024 * 4);
    cudaMalloc(&d.Bhi, (size_t)1024 * 1024 * 4); cudaMalloc(&d.Blo, (size_t)1024 * 1024 * 4); cudaMalloc(&d.D, (size_t)M * 1024 * 4);
    cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 1024);

    printf("network: amass.yaml DFNet, act %s, 128 poses; errors against fp64\n", argv[2]);
    printf("%-44s %12s %12s %14s %14s\n", "GEMM arithmetic", "d max rel", "d median rel", "grad max rel", "grad median rel");
    const char* names[5] = {"fp32 FMA chain (the product kernel's)", "tcgen05 1xTF32", "tcgen05 3xTF32, one accumulator",
                            "tcgen05 3xTF32, split accumulators", "tcgen05 3xTF32, split acc + 128-deep K chunks"};
    for (int mode = 0; mode < 5; ++mode) {
        auto gemm = [&](const std::vector<float>& A, const std::vector<float>& B, int N, int K, std::vector<float>& C) {
            if (mode == 0) ffma_gemm(A, B, N, K, C); else tc_gemm(d, A, B, N, K, C, mode - 1, 128);
        };
        std::vector<float> z = z0; int kp = 128;
        std::vector<std::vector<float>> dph(6);
        for (int l = 0; l < 6; ++l) {
            std::vector<float> pre; gemm(z, Wt[l], W[l + 1], KP[l], pre);
            const int np = (W[l + 1] + 31) / 32 * 32;
            std::vector<float> zn((size_t)M * np, 0.f); dph[l].assign((size_t)M * W[l + 1], 0.f);
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < W[l + 1]; ++n) {
                    const float s = pre[(size_t)m * W[l + 1] + n] + bias[l][n];
                    double dv; zn[(size_t)m * np + n] = (float)actf((double)s, dv); dph[l][(size_t)m * W[l + 1] + n] = (float)dv;
                }
            z.swap(zn); kp = np;
        }
        std::vector<float> dd(M), gs(M);
        for (int m = 0; m < M; ++m) {
            float s = bias[6][0];
            for (int k = 0; k < 64; ++k) s = fmaf(z[(size_t)m * kp + k], Wt[6][k], s);
            double dv; if (soft) dd[m] = (float)actf((double)s, dv); else { dd[m] = s > 0 ? s : 0; dv = s > 0 ? 1 : 0; }
            gs[m] = (float)dv;
        }
        std::vector<float> g((size_t)M * 64);
        for (int m = 0; m < M; ++m) for (int k = 0; k < 64; ++k) g[(size_t)m * 64 + k] = gs[m] * Wt[6][k];
        for (int l = 5; l >= 0; --l) {
            const int nout = W[l + 1], nin = KP[l];          // g_l[128 x nin] = t[128 x nout] * (W_l^T)[nin x nout]^T
            const int kpad = (nout + 31) / 32 * 32;
            std::vector<float> t((size_t)M * kpad, 0.f), WT((size_t)nin * kpad, 0.f);
            for (int m = 0; m < M; ++m) for (int n = 0; n < nout; ++n) t[(size_t)m * kpad + n] = g[(size_t)m * nout + n] * dph[l][(size_t)m * nout + n];
            for (int n = 0; n < nout; ++n) for (int k = 0; k < nin; ++k) WT[(size_t)k * kpad + n] = Wt[l][(size_t)n * KP[l] + k];
            std::vector<float> gn; gemm(t, WT, nin, kpad, gn);
            if (l == 0) { g.assign((size_t)M * 126, 0.f); for (int m = 0; m < M; ++m) for (int k = 0; k < 126; ++k) g[(size_t)m * 126 + k] = gn[(size_t)m * nin + k]; }
            else { g.assign((size_t)M * W[l], 0.f); for (int m = 0; m < M; ++m) for (int k = 0; k < W[l]; ++k) g[(size_t)m * W[l] + k] = gn[(size_t)m * nin + k]; }
        }
        std::vector<double> ed(M), eg(M);
        for (int m = 0; m < M; ++m) {
            ed[m] = fabs((double)dd[m] - dref[m]) / fmax(fabs(dref[m]), 1e-30);
            double num = 0, den = 0;
            for (int k = 0; k < 126; ++k) { const double a = (double)g[(size_t)m * 126 + k] - gref[(size_t)m * 126 + k]; num += a * a; den += gref[(size_t)m * 126 + k] * gref[(size_t)m * 126 + k]; }
            eg[m] = sqrt(num) / fmax(sqrt(den), 1e-300);
        }
        auto med = [](std::vector<double> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; };
        auto mx = [](const std::vector<double>& v) { double m = 0; for (double x : v) m = fmax(m, x); return m; };
        printf("%-44s %12.3e %12.3e %14.3e %14.3e\n", names[mode], mx(ed), med(ed), mx(eg), med(eg));
    }
    return 0;
}
