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

#include "pndf_kernel.cuh"

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

// ------------------------------------------------------------------------------------------------ exact search
// Exact k = 5 nearest database poses of every query under the same metrics -- the labels data/prepare_traindata.py:138-170
// approximates with a faiss candidate list (500 nearest by SMPL joint positions) before its rerank.  Brute force is
// fp32-FMA-bound, not HBM-bound: every database row staged in shared memory is scored against 64 queries.
//
//   grid  (ceil(Q/64), nsplit): a CTA owns 64 queries and one contiguous slice of the database
//   CTA   256 threads = 16 query groups x 16 row groups; a thread scores 4 queries x 8 rows per 128-row tile
//         (rows td, td+16, ...: consecutive lanes read consecutive 336-byte rows, conflict-free LDS.128)
//   tiles 128 rows = 43 008 contiguous bytes, TMA bulk copy into a 2-stage shared-memory ring (mbarrier complete_tx)
//   top-5 per thread and query in registers, merged over the 16 row groups in shared memory, then over the database
//         slices by knn_merge_kernel.  Ties: lower database index first.
constexpr int kExQ = 64, kExD = 128, kExThreads = 256;
constexpr int kExSmem = kExQ * 84 * 4 + 2 * kExD * 84 * 4 + 2 * 8 + 16;

struct KnnExactParams {
    const float* query;      // Q x 84
    const float* database;   // N x 84
    float* part_val;         // Q x nsplit x 5
    int32_t* part_idx;
    long long Q, N, rows_per_split;
    int nsplit, weighted;
};

__device__ __forceinline__ void top5_insert(float (&bv)[kKnnK], int (&bi)[kKnnK], float d, int idx) {
    bv[kKnnK - 1] = d; bi[kKnnK - 1] = idx;
#pragma unroll
    for (int i = kKnnK - 1; i > 0; --i) {
        const bool sw = (bv[i] < bv[i - 1]) || (bv[i] == bv[i - 1] && bi[i] < bi[i - 1]);
        if (sw) {
            const float tv = bv[i]; bv[i] = bv[i - 1]; bv[i - 1] = tv;
            const int ti = bi[i]; bi[i] = bi[i - 1]; bi[i - 1] = ti;
        }
    }
}

template <int METRIC>
__global__ void __launch_bounds__(kExThreads, 2) knn_exact_kernel(const KnnExactParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* qs = reinterpret_cast<float*>(smem);
    float* dbs = qs + kExQ * 84;
    uint64_t* full = reinterpret_cast<uint64_t*>(dbs + 2 * kExD * 84);
    const int tid = threadIdx.x, tq = tid >> 4, td = tid & 15;
    const long long q0 = (long long)blockIdx.x * kExQ;
    const long long s0 = (long long)blockIdx.y * p.rows_per_split;
    const long long s1 = min(p.N, s0 + p.rows_per_split);
    const int ntiles = (int)((s1 - s0 + kExD - 1) / kExD);
    const uint32_t bar_s = smem_u32(full), dbs_s = smem_u32(dbs);

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int idx = tid; idx < kExQ * 84; idx += kExThreads) {
        const long long q = q0 + idx / 84;
        qs[idx] = (q < p.Q) ? __ldg(p.query + q0 * 84 + idx) : 0.0f;
    }
    __syncthreads();
    auto issue = [&](int t) {     // thread 0: bulk copy of tile t into stage t & 1
        const long long r0 = s0 + (long long)t * kExD;
        const uint32_t bytes = (uint32_t)(min((long long)kExD, s1 - r0) * 336);
        const uint32_t bar = bar_s + (t & 1) * 8;
        mbar_expect_tx_s(bar, bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         dbs_s + (t & 1) * (kExD * 336)),
                     "l"(p.database + r0 * 84), "r"(bytes), "r"(bar)
                     : "memory");
    };
    if (tid == 0) {
        if (ntiles > 0) issue(0);
        if (ntiles > 1) issue(1);
    }
    float wsum = 0.0f;
#pragma unroll
    for (int j = 0; j < 21; ++j) wsum += p.weighted ? c_joint_w[j] : (1.0f / 21.0f);

    float bv[4][kKnnK];
    int bi[4][kKnnK];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < kKnnK; ++k) { bv[i][k] = 3.0e38f; bi[i][k] = 0x7fffffff; }

    for (int t = 0; t < ntiles; ++t) {
        mbar_wait_s(bar_s + (t & 1) * 8, (uint32_t)((t >> 1) & 1));
        const float* tile = dbs + (t & 1) * (kExD * 84);
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[i][r] = 0.0f;
#pragma unroll 3
        for (int j = 0; j < 21; ++j) {
            const float w = p.weighted ? c_joint_w[j] : (1.0f / 21.0f);
            float4 qv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(qs + (tq * 4 + i) * 84 + j * 4);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 d = *reinterpret_cast<const float4*>(tile + (td + 16 * r) * 84 + j * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (METRIC == 0) {
                        const float dot = fmaf(qv[i].w, d.w, fmaf(qv[i].z, d.z, fmaf(qv[i].y, d.y, qv[i].x * d.x)));
                        acc[i][r] = fmaf(w, fabsf(dot), acc[i][r]);
                    } else {
                        const float a = qv[i].x - d.x, b = qv[i].y - d.y, c = qv[i].z - d.z, e = qv[i].w - d.w;
                        // |q - q'| = s * rsqrt(s): one MUFU + one FMUL instead of the IEEE sqrt sequence (2-3 ulp, far inside
                        // the 1e-5 label tolerance); rsqrt(0) = inf is masked
                        const float s2 = fmaf(e, e, fmaf(c, c, fmaf(b, b, a * a)));
                        const float nrm = (s2 > 0.0f) ? s2 * rsqrtf(s2) : 0.0f;
                        acc[i][r] = fmaf(w, nrm, acc[i][r]);
                    }
                }
            }
        }
        const long long row_base = s0 + (long long)t * kExD + td;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const long long row = row_base + 16 * r;
            if (row < s1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float dist = (METRIC == 0) ? (wsum - acc[i][r]) : acc[i][r];
                    if (dist < bv[i][kKnnK - 1]) top5_insert(bv[i], bi[i], dist, (int)row);
                }
            }
        }
        __syncthreads();                              // everyone is done with stage t & 1
        if (tid == 0 && t + 2 < ntiles) issue(t + 2);
    }
    // ---- merge the 16 row groups of every query (stage 0 of the ring is free now)
    float* mv = dbs;
    int* mi = reinterpret_cast<int*>(dbs + kExQ * 16 * kKnnK);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < kKnnK; ++k) {
            mv[((tq * 4 + i) * 16 + td) * kKnnK + k] = bv[i][k];
            mi[((tq * 4 + i) * 16 + td) * kKnnK + k] = bi[i][k];
        }
    __syncthreads();
    if (tid < kExQ && q0 + tid < p.Q) {
        float fv[kKnnK];
        int fi[kKnnK];
#pragma unroll
        for (int k = 0; k < kKnnK; ++k) { fv[k] = 3.0e38f; fi[k] = 0x7fffffff; }
        for (int c = 0; c < 16 * kKnnK; ++c) {
            const float d = mv[tid * 16 * kKnnK + c];
            const int ix = mi[tid * 16 * kKnnK + c];
            if (d < fv[kKnnK - 1] || (d == fv[kKnnK - 1] && ix < fi[kKnnK - 1])) top5_insert(fv, fi, d, ix);
        }
        const long long o = ((q0 + tid) * p.nsplit + blockIdx.y) * kKnnK;
#pragma unroll
        for (int k = 0; k < kKnnK; ++k) { p.part_val[o + k] = fv[k]; p.part_idx[o + k] = fi[k]; }
    }
}

// one thread per query: top-5 of the nsplit x 5 partial results
__global__ void knn_merge_kernel(const float* part_val, const int32_t* part_idx, long long Q, int nsplit, float* out_val,
                                 int32_t* out_idx) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    float fv[kKnnK];
    int fi[kKnnK];
#pragma unroll
    for (int k = 0; k < kKnnK; ++k) { fv[k] = 3.0e38f; fi[k] = 0x7fffffff; }
    for (int c = 0; c < nsplit * kKnnK; ++c) {
        const float d = part_val[q * nsplit * kKnnK + c];
        const int ix = part_idx[q * nsplit * kKnnK + c];
        if (d < fv[kKnnK - 1] || (d == fv[kKnnK - 1] && ix < fi[kKnnK - 1])) top5_insert(fv, fi, d, ix);
    }
#pragma unroll
    for (int k = 0; k < kKnnK; ++k) { out_val[q * kKnnK + k] = fv[k]; out_idx[q * kKnnK + k] = fi[k]; }
}

}  // namespace pndf
