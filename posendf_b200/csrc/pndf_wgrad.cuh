// pndf_wgrad.cuh -- weight gradients of the DFNet and the fused optimizer step (training, config C5).
//
// Reference: loss.backward() through the graph of model/posendf.py:62-99 (model/train_posendf.py:98) and
// torch.optim.Adam(lr, weight_decay=1e-4).step() (model/train_posendf.py:30,99).
//
// The fused kernel exports, per pose b and DFNet layer l, the layer input z_l[b], the pre-activation adjoint
// a_l[b] = dd/dpre_l (unit upstream) and -- from the tangent launch -- the forward-mode tangent zdot_l[b], pose-major:
// dump[b][kDumpRows] (column map in DESIGN.md).  The parameter gradients are batch reductions of outer products
//
//     dW_l = sum_b a_l[b] (x) r_l[b],      r_l[b] = c_b z_l[b] + w_e zdot_l[b],      db_l = sum_b c_b a_l[b]
//
// with c_b = (upstream weight) * dLoss/dd_b and w_e the upstream weight of the Eikonal loss (device scalars: nothing here
// synchronises).  Both GEMM operands are K-major rows of the dump (K = pose), so the kernel is a split-K outer-product
// SGEMM on the packed FFMA2 pipe: 128 x 128 output tile per CTA, 8 x 8 micro-tile per thread, 32-pose stages landed by
// cp.async.bulk (TMA 1-D copies, one 512-byte row segment each) on mbarriers, r formed in shared memory on the fly.
// Split-K partials go to a workspace slot per K-split (plain stores, no atomics); wgrad_reduce_kernel sums the slots in a
// fixed order into the flat gradient vector, so the result is deterministic.  Work items are launched K-split-major
// (blockIdx.y = split): all tiles that read the same 1 024 poses (22 MB of exports) run together and share them in L2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pndf_kernel.cuh"

namespace pndf {

constexpr int kWgTile = 128;          // output tile is kWgTile x kWgTile
constexpr int kWgBK = 32;             // poses per pipeline stage
constexpr int kWgStages = 2;
constexpr int kWgThreads = 256;
// poses per work item (one K-split): chosen per call by the host (multiple of kWgBK, ~1 000) so that the CTA count fills whole
// waves of 2 CTAs per SM
constexpr int kWgStageFloats = 3 * kWgBK * kWgTile;     // A | Z | Zdot
constexpr int kWgSmem = kWgStages * kWgStageFloats * 4 + 64;
constexpr int kWgMaxProblems = 6;

struct WgProblem {
    int a_col, z_col;        // first dump column of the adjoint (M side) and of the layer input (N side)
    int n_out, n_in;         // real extents of W_l (row-major n_out x n_in in the flat parameter vector)
    int m_tiles, n_tiles;
    int tile0;               // first work item (blockIdx.x) of this problem
    long long w_off, b_off;  // offsets of W_l and b_l in the flat parameter vector
};

struct WgParams {
    const float* dump;       // [B][kDumpRows] launch-1 exports
    const float* dump_t;     // tangent-launch exports (columns [0, 2752) = zdot) or nullptr
    const float* coef;       // [B] dLoss/dd per pose for unit upstream, or nullptr: `uniform` for every pose
    const float* up;         // device scalar: upstream gradient of the loss behind coef
    const float* w_eik;      // device scalar: upstream gradient of the Eikonal loss (used with dump_t) or nullptr
    float uniform;
    float* ws;               // workspace [slot][n_params]
    long long B, ws_stride;  // ws_stride = n_params rounded up to a multiple of 4 floats (16-byte aligned slots)
    int kc;                  // poses per K-split (multiple of kWgBK)
    int slot0;               // this batch's first workspace slot; slot = slot0 + blockIdx.y
    int nprob;
    WgProblem prob[kWgMaxProblems];
};

__device__ __forceinline__ void wg_ffma2(float (&acc)[8][8], const float (&a)[8], const float (&r)[8]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        unsigned long long bb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(r[2 * j]), "f"(r[2 * j + 1]));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            unsigned long long aa, cc;
            asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a[i]));
            asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(acc[i][2 * j]), "f"(acc[i][2 * j + 1]));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(aa), "l"(bb));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i][2 * j]), "=f"(acc[i][2 * j + 1]) : "l"(cc));
        }
    }
}

__global__ void __launch_bounds__(kWgThreads, 2) wgrad_kernel(const WgParams p) {
    extern __shared__ __align__(128) uint8_t wg_smem[];
    float* stage_base = reinterpret_cast<float*>(wg_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(wg_smem + kWgStages * kWgStageFloats * 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- which tile
    int pi = 0;
#pragma unroll
    for (int i = 1; i < kWgMaxProblems; ++i)
        if (i < p.nprob && (int)blockIdx.x >= p.prob[i].tile0) pi = i;
    const WgProblem pr = p.prob[pi];
    const int t = blockIdx.x - pr.tile0;
    const int mt = t / pr.n_tiles, nt = t - mt * pr.n_tiles;
    const long long k0 = (long long)blockIdx.y * p.kc;
    const long long kend = min(p.B, k0 + p.kc);
    const int nchunks = (int)((kend - k0 + kWgBK - 1) / kWgBK);
    const bool has_t = (p.dump_t != nullptr) && (p.w_eik != nullptr);
    const bool per_pose = (p.coef != nullptr);
    const float up = __ldg(p.up);
    const float we = has_t ? __ldg(p.w_eik) : 0.0f;

    if (tid == 0) {
        for (int s = 0; s < kWgStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // producer: warp 0, lane = pose row of the stage; one 512-byte bulk copy per operand and row
    auto issue = [&](int chunk) {
        const int s = chunk & (kWgStages - 1);
        const long long kb = k0 + (long long)chunk * kWgBK;
        const int rows = (int)min((long long)kWgBK, kend - kb);
        float* sa = stage_base + s * kWgStageFloats;
        const uint32_t bar = smem_u32(&full[s]);
        if (lane == 0) mbar_expect_tx_s(bar, (uint32_t)rows * (has_t ? 3u : 2u) * (kWgTile * 4));
        __syncwarp();
        if (lane < rows) {
            const float* ga = p.dump + (kb + lane) * kDumpRows + pr.a_col + mt * kWgTile;
            const float* gz = p.dump + (kb + lane) * kDumpRows + pr.z_col + nt * kWgTile;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(sa + lane * kWgTile)), "l"(ga), "r"(kWgTile * 4), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(sa + (kWgBK + lane) * kWgTile)), "l"(gz), "r"(kWgTile * 4), "r"(bar) : "memory");
            if (has_t) {
                const float* gt = p.dump_t + (kb + lane) * kDumpRows + pr.z_col + nt * kWgTile;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(sa + (2 * kWgBK + lane) * kWgTile)), "l"(gt), "r"(kWgTile * 4), "r"(bar) : "memory");
            }
        }
    };
    if (warp == 0) {
        issue(0);
        if (nchunks > 1) issue(1);
    }

    // thread tile: rows m = tm*4 + {0..3} and 64 + tm*4 + {0..3}; columns n = tn*4 + {0..3} and 64 + tn*4 + {0..3}
    const int tm = (warp >> 1) * 4 + (lane >> 3), tn = (warp & 1) * 8 + (lane & 7);
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    float bsum = 0.0f;     // bias gradient of output row (mt*128 + tid), threads 0..127 of the nt == 0 tiles

    for (int c = 0; c < nchunks; ++c) {
        const int s = c & (kWgStages - 1);
        float* sa = stage_base + s * kWgStageFloats;
        float* sz = sa + kWgBK * kWgTile;
        float* st = sz + kWgBK * kWgTile;
        mbar_wait_s(smem_u32(&full[s]), (uint32_t)(c >> 1) & 1u);
        const long long kb = k0 + (long long)c * kWgBK;
        const int rows = (int)min((long long)kWgBK, kend - kb);
        // r = c_b z + w_e zdot, in place over z; rows past the end of the batch contribute nothing
        if (per_pose || has_t || rows < kWgBK) {
#pragma unroll
            for (int i = 0; i < (kWgBK * kWgTile / 4) / kWgThreads; ++i) {
                const int f = tid + i * kWgThreads;
                const int row = f >> 5;
                float4* zp = reinterpret_cast<float4*>(sz) + f;
                if (row >= rows) {
                    *zp = make_float4(0.f, 0.f, 0.f, 0.f);
                    reinterpret_cast<float4*>(sa)[f] = make_float4(0.f, 0.f, 0.f, 0.f);
                    continue;
                }
                if (per_pose || has_t) {
                    const float cb = per_pose ? up * __ldg(p.coef + kb + row) : 1.0f;
                    float4 z = *zp;
                    z.x *= cb; z.y *= cb; z.z *= cb; z.w *= cb;
                    if (has_t) {
                        const float4 d = reinterpret_cast<const float4*>(st)[f];
                        z.x = fmaf(we, d.x, z.x); z.y = fmaf(we, d.y, z.y); z.z = fmaf(we, d.z, z.z); z.w = fmaf(we, d.w, z.w);
                    }
                    *zp = z;
                }
            }
            __syncthreads();
        }
        if (nt == 0 && tid < kWgTile) {
#pragma unroll 8
            for (int kk = 0; kk < kWgBK; ++kk) {
                const float cb = per_pose ? ((kk < rows) ? up * __ldg(p.coef + kb + kk) : 0.0f) : 1.0f;
                bsum = fmaf(cb, sa[kk * kWgTile + tid], bsum);
            }
        }
#pragma unroll 8
        for (int kk = 0; kk < kWgBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(sa + kk * kWgTile + tm * 4);
            const float4 a1 = *reinterpret_cast<const float4*>(sa + kk * kWgTile + 64 + tm * 4);
            const float4 r0 = *reinterpret_cast<const float4*>(sz + kk * kWgTile + tn * 4);
            const float4 r1 = *reinterpret_cast<const float4*>(sz + kk * kWgTile + 64 + tn * 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            wg_ffma2(acc, a, r);
        }
        __syncthreads();      // everyone is done with stage s
        if (warp == 0 && c + kWgStages < nchunks) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes of the r pass before the async refill
            issue(c + kWgStages);
        }
    }

    // ---- partial tile -> workspace slot (flat parameter layout)
    const float scale = per_pose ? 1.0f : up * p.uniform;
    float* ws = p.ws + (size_t)(p.slot0 + blockIdx.y) * (size_t)p.ws_stride;
    float* W = ws + pr.w_off;
    const bool vec = (pr.n_in & 3) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = mt * kWgTile + ((i < 4) ? (tm * 4 + i) : (64 + tm * 4 + (i - 4)));
        if (m >= pr.n_out) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = nt * kWgTile + h * 64 + tn * 4;
            float* dst = W + (size_t)m * pr.n_in + n;
            if (vec && n + 3 < pr.n_in) {
                *reinterpret_cast<float4*>(dst) = make_float4(acc[i][h * 4] * scale, acc[i][h * 4 + 1] * scale,
                                                              acc[i][h * 4 + 2] * scale, acc[i][h * 4 + 3] * scale);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < pr.n_in) dst[j] = acc[i][h * 4 + j] * scale;
            }
        }
    }
    if (nt == 0 && tid < kWgTile) {
        const int m = mt * kWgTile + tid;
        if (m < pr.n_out) ws[pr.b_off + m] = bsum * scale;
    }
}

// Last layer (W_6: 1 x 64, b_6) and nothing else: a_6 = phi_out'(s) is recovered from the distance (ReLU: d > 0, softplus:
// 1 - exp(-beta d)).  One CTA per K-split; 16 pose lanes x 16 column quads, the pose-lane partials are summed in a fixed order.
struct WgLastParams {
    const float* dump;
    const float* dump_t;
    const float* coef;
    const float* up;
    const float* w_eik;
    const float* dist;       // [B]
    float uniform;
    float* ws;
    long long B, ws_stride, w6_off, b6_off;
    int kc, slot0, z6_col, softplus;
    float beta;
};

__global__ void __launch_bounds__(256) wgrad_last_kernel(const WgLastParams p) {
    // thread (kq, n4): pose lane kq of 16, columns 4 n4 .. 4 n4 + 3 of z6 -- 16 poses x 256 bytes per iteration, 16-byte loads
    __shared__ float4 part[16][16];
    __shared__ float bpart[16];
    const int n4 = threadIdx.x & 15, kq = threadIdx.x >> 4;
    const long long k0 = (long long)blockIdx.x * p.kc, kend = min(p.B, k0 + p.kc);
    const bool has_t = (p.dump_t != nullptr) && (p.w_eik != nullptr);
    const float up = __ldg(p.up), we = has_t ? __ldg(p.w_eik) : 0.0f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float bacc = 0.0f;
#pragma unroll 4
    for (long long k = k0 + kq; k < kend; k += 16) {
        const float d = __ldg(p.dist + k);
        const float gs = p.softplus ? -expm1f(-p.beta * d) : (d > 0.0f ? 1.0f : 0.0f);
        const float cb = (p.coef != nullptr) ? up * __ldg(p.coef + k) : up * p.uniform;
        const float4 z = __ldg(reinterpret_cast<const float4*>(p.dump + k * kDumpRows + p.z6_col) + n4);
        float4 r = make_float4(cb * z.x, cb * z.y, cb * z.z, cb * z.w);
        if (has_t) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.dump_t + k * kDumpRows + p.z6_col) + n4);
            r.x = fmaf(we, t.x, r.x); r.y = fmaf(we, t.y, r.y); r.z = fmaf(we, t.z, r.z); r.w = fmaf(we, t.w, r.w);
        }
        acc.x = fmaf(gs, r.x, acc.x); acc.y = fmaf(gs, r.y, acc.y); acc.z = fmaf(gs, r.z, acc.z); acc.w = fmaf(gs, r.w, acc.w);
        if (n4 == 0) bacc = fmaf(gs, cb, bacc);
    }
    part[kq][n4] = acc;
    if (n4 == 0) bpart[kq] = bacc;
    __syncthreads();
    float* ws = p.ws + (size_t)(p.slot0 + blockIdx.x) * (size_t)p.ws_stride;
    if (threadIdx.x < 64) {      // column threadIdx.x: the 16 pose-lane partials in a fixed order
        float s = 0.0f;
        for (int q = 0; q < 16; ++q) s += reinterpret_cast<const float*>(&part[q][threadIdx.x >> 2])[threadIdx.x & 3];
        ws[p.w6_off + threadIdx.x] = s;
    } else if (threadIdx.x == 64) {
        float s = 0.0f;
        for (int q = 0; q < 16; ++q) s += bpart[q];
        ws[p.b6_off] = s;
    }
}

// grad[i] += sum over slots of ws[slot][i] (fixed order) for the DFNet part i >= enc_floats, and the encoder kernel's
// (n_enc_rows x enc_floats) block accumulators for i < enc_floats.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, int nslots, long long ws_stride, long long n_params,
                                                           int enc_floats,
                                                           const float* __restrict__ enc_rows, int n_enc_rows,
                                                           float* __restrict__ grad, int overwrite) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_params) return;
    float s = 0.0f;
    if (i < enc_floats) {
        for (int r = 0; r < n_enc_rows; ++r) s += enc_rows[(size_t)r * enc_floats + i];
    } else {
        for (int k = 0; k < nslots; ++k) s += ws[(size_t)k * ws_stride + i];
    }
    grad[i] = overwrite ? s : grad[i] + s;
}

// ------------------------------------------------------------------------------------------------ losses
// Per-pose part of model/posendf.py:85-96 on the outputs of launch 1, one thread per pose:
//   MODE_POSE : diff = d - d_gt;  coef = sign(diff)/N (l1) | 2 diff/N (l2);  sum |diff| or diff^2;
//               Eikonal: n_j = |g_{b,j}|, sum (n_j - 1)^2, v_{b,j} = 2 (n_j - 1)/(21 N) g_{b,j}/n_j (0 where n_j == 0, as torch's norm)
//   MODE_MAN  : sum |d|
// Block partials are summed by the last block to finish (fixed order -> deterministic), added to running totals (several
// chunks per step) and the means are published as device scalars.
struct LossParams {
    const float* dist;       // [B]
    const float* dist_gt;    // [B] (pose mode)
    const float* grad;       // [B][84] dd/dpose (pose mode with Eikonal) or nullptr
    float* coef;             // [B] out (pose mode)
    float* v;                // [B][84] out or nullptr
    float* partial;          // [gridDim.x][2] scratch
    unsigned int* counter;   // zero-initialised, self-resetting
    double* totals;          // [3] running sums: dist loss, Eikonal, manifold
    float* losses;           // [3] published means
    long long B;             // poses in this launch
    double inv_n;            // 1 / (poses of the whole batch)
    int mode, l2, reset;     // reset: first chunk of a step zeroes the running totals it owns
};

__global__ void __launch_bounds__(256) train_loss_kernel(const LossParams p) {
    __shared__ float red[2][8];
    __shared__ bool is_last;
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float s0 = 0.0f, s1 = 0.0f;
    if (b < p.B) {
        const float d = p.dist[b];
        if (p.mode == 0) {
            const float diff = d - p.dist_gt[b];
            const float invn = (float)p.inv_n;
            if (p.l2) {
                s0 = diff * diff;
                p.coef[b] = 2.0f * diff * invn;
            } else {
                s0 = fabsf(diff);
                p.coef[b] = (diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f)) * invn;
            }
            if (p.grad != nullptr) {
                const float4* g = reinterpret_cast<const float4*>(p.grad + b * 84);
                float4* vo = reinterpret_cast<float4*>(p.v + b * 84);
                const float cs = 2.0f * invn / 21.0f;
#pragma unroll 7
                for (int j = 0; j < 21; ++j) {
                    const float4 q = g[j];
                    const float nrm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
                    const float e = nrm - 1.0f;
                    s1 = fmaf(e, e, s1);
                    const float k = (nrm > 0.0f) ? cs * e / nrm : 0.0f;
                    vo[j] = make_float4(k * q.x, k * q.y, k * q.z, k * q.w);
                }
            }
        } else {
            s0 = fabsf(d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t0 = 0.0f, t1 = 0.0f;
        for (int w = 0; w < 8; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
        p.partial[blockIdx.x * 2] = t0;
        p.partial[blockIdx.x * 2 + 1] = t1;
        __threadfence();
        is_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last || threadIdx.x != 0) return;
    __threadfence();
    double t0 = 0.0, t1 = 0.0;
    for (unsigned i = 0; i < gridDim.x; ++i) {
        t0 += (double)__ldcg(p.partial + i * 2);
        t1 += (double)__ldcg(p.partial + i * 2 + 1);
    }
    *p.counter = 0u;
    if (p.mode == 0) {
        if (p.reset) { p.totals[0] = 0.0; p.totals[1] = 0.0; }
        p.totals[0] += t0; p.totals[1] += t1;
        p.losses[0] = (float)(p.totals[0] * p.inv_n);
        p.losses[1] = (float)(p.totals[1] * p.inv_n / 21.0);
    } else {
        if (p.reset) p.totals[2] = 0.0;
        p.totals[2] += t0;
        p.losses[2] = (float)(p.totals[2] * p.inv_n);
    }
}

// ------------------------------------------------------------------------------------------------ optimizer
// torch.optim.Adam(params, lr, betas, eps, weight_decay) single step, same operation order as torch's _single_tensor_adam
// (grad += wd * p; m.lerp_(g, 1 - b1); v = v * b2 + (1 - b2) g g; denom = sqrt(v) / sqrt(1 - b2^t) + eps;
// p += -(lr / (1 - b1^t)) * m / denom), followed in the same thread by the re-packing of the new value into the engine's
// slab stream (each DFNet weight lives there twice: forward and reverse layout) and small-parameter buffer.
struct AdamStepParams {
    float* param;
    const float* grad;
    float* m;
    float* v;
    long long n;
    // scalars are formed in double on the host exactly as torch forms them (1 - beta, lr / (1 - beta1^t), sqrt(1 - beta2^t))
    float lr_over_bias1, bias2_sqrt, one_minus_beta1, beta2, one_minus_beta2, eps, weight_decay;
    float grad_scale;                   // 1 / world size after a summing all-reduce (data parallel), else 1
    const int32_t* pos_a;               // [n] slab-stream position (forward layout) or -1
    const int32_t* pos_b;               // [n] slab-stream position (reverse layout) or -1
    const int32_t* pos_s;               // [n] small-buffer position or -1
    float* wstream;
    float* small;
};

__global__ void __launch_bounds__(256) adam_step_kernel(const AdamStepParams p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    float w = p.param[i];
    float g = p.grad[i] * p.grad_scale;
    if (p.weight_decay != 0.0f) g = fmaf(p.weight_decay, w, g);
    float m = p.m[i], v = p.v[i];
    m = fmaf(g - m, p.one_minus_beta1, m);
    v = fmaf(p.one_minus_beta2 * g, g, v * p.beta2);
    const float denom = sqrtf(v) / p.bias2_sqrt + p.eps;
    w = fmaf(-p.lr_over_bias1, m / denom, w);
    p.m[i] = m; p.v[i] = v; p.param[i] = w;
    const int32_t a = p.pos_a[i], b = p.pos_b[i], s = p.pos_s[i];
    if (a >= 0) p.wstream[a] = w;
    if (b >= 0) p.wstream[b] = w;
    if (s >= 0) p.small[s] = w;
}

}  // namespace pndf
