// pndf_knn.cuh -- distance-label rerank: for every noisy query pose, the k = 5 nearest of its K candidate
// manifold poses under the reference's quaternion metrics (data/dist_utils.py:19-30 `euc`, :41-50 `geo`), followed by
// torch.topk(k=5, largest=False) (data/prepare_traindata.py:156).  The candidates are given as indices into the pose
// database (what faiss returns), so the (Q,K,21,4) gather the reference materialises never exists.
//
// HBM-bound by construction: 336 B of candidate pose are read per (query, candidate) pair for 21*(4 FMA + abs) flops
// -> one warp per query, rows are read coalesced (see the kernel comment).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pndf {

constexpr int kKnnK = 5;

struct KnnParams {
    const float* query;      // Q x 84
    const float* database;   // N x 84
    const int32_t* cand;     // Q x K indices into the database
    float* out_val;          // Q x 5 ascending
    int32_t* out_pos;        // Q x 5 positions inside the candidate list (torch.topk indices)
    long long Q;
    int K, metric, weighted; // metric 0 = geo, 1 = euc
};

// F.normalize(joint_rank, dim=0) with joint_rank = [7,7,7,6,6,6,5,5,5,4,4,4,4,4,3,3,3,2,2,1,1] (data/dist_utils.py:15-16)
__constant__ float c_joint_w[21] = {0.33108863f, 0.33108863f, 0.33108863f, 0.28379026f, 0.28379026f, 0.28379026f, 0.23649189f, 0.23649189f, 0.23649189f, 0.18919352f, 0.18919352f, 0.18919352f, 0.18919352f, 0.18919352f, 0.14189513f, 0.14189513f, 0.14189513f, 0.09459676f, 0.09459676f, 0.04729838f, 0.04729838f};

// One warp per query.  A candidate row (21 joints x 16 B = 336 B, contiguous) is loaded by lanes 0..20 with ONE coalesced
// LDG.128 per lane (3-4 cache lines per warp instruction instead of 32 scattered ones), lane j scores joint j, a shuffle
// tree adds the 21 terms; 8 candidates are in flight per warp so the random-row HBM latency is covered.
__global__ void __launch_bounds__(128) knn_rerank_kernel(const KnnParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long q = (long long)blockIdx.x * 4 + warp;
    if (q >= p.Q) return;
    const bool act = lane < 21;
    float4 qq = make_float4(0.f, 0.f, 0.f, 0.f);
    if (act) qq = __ldg(reinterpret_cast<const float4*>(p.query + q * 84) + lane);
    const float wj = act ? (p.weighted ? c_joint_w[lane] : (1.0f / 21.0f)) : 0.0f;
    float bv[kKnnK];
    int bp[kKnnK];
#pragma unroll
    for (int i = 0; i < kKnnK; ++i) { bv[i] = 3.0e38f; bp[i] = 0x7fffffff; }
    const int32_t* cand = p.cand + q * p.K;
    constexpr int U = 8;
    for (int c0 = 0; c0 < p.K; c0 += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = min(c0 + u, p.K - 1);
            const int32_t row = __ldg(cand + c);                                   // uniform across the warp
            v[u] = act ? __ldg(reinterpret_cast<const float4*>(p.database + (size_t)row * 84) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float t;
            if (p.metric == 0) {
                t = 1.0f - fabsf(v[u].x * qq.x + v[u].y * qq.y + v[u].z * qq.z + v[u].w * qq.w);
            } else {
                const float a = qq.x - v[u].x, b = qq.y - v[u].y, e = qq.z - v[u].z, d = qq.w - v[u].w;
                t = sqrtf(a * a + b * b + e * e + d * d);
            }
            float acc = wj * t;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            const int c = c0 + u;
            // every lane keeps the same sorted top-5 (ties: lower candidate position first, like a stable sort)
            if (c < p.K && acc < bv[kKnnK - 1]) {
                bv[kKnnK - 1] = acc; bp[kKnnK - 1] = c;
#pragma unroll
                for (int i = kKnnK - 1; i > 0; --i) {
                    if (bv[i] < bv[i - 1]) {
                        const float tv = bv[i]; bv[i] = bv[i - 1]; bv[i - 1] = tv;
                        const int tp = bp[i]; bp[i] = bp[i - 1]; bp[i - 1] = tp;
                    }
                }
            }
        }
    }
    if (lane < kKnnK) {
        float ov = bv[0]; int op = bp[0];
#pragma unroll
        for (int i = 1; i < kKnnK; ++i) if (lane == i) { ov = bv[i]; op = bp[i]; }
        p.out_val[q * kKnnK + lane] = ov;
        p.out_pos[q * kKnnK + lane] = op;
    }
}

}  // namespace pndf
