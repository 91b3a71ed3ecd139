// pndf_encoder_train.cuh -- training-side kernels for the structure encoder (3 516 parameters, 0.2 % of the flops).
//
// The fused kernel leaves three per-pose vectors on the encoder output z0 (126 features):
//     up1  = delta * dd/dz0              first-order term (dist / manifold loss)            -> gradient set 0
//     upt  = dd/dz0                      adjoint of the TANGENT of z0 (Eikonal term)        -> gradient set 1
//     upz  = second-order adjoint of z0  (softplus DFNet only, else null)                   -> gradient set 1
// and the Eikonal term needs the tangent of z0 along the pose tangent v = dE/dg as the input of the tangent launch.
// Both are one-thread-per-pose walks of the kinematic tree (reference: model/network/net_modules.py:140-170):
//     enc_tangent_kernel : q = x/n, qdot = J_normalise v, forward + forward-mode tangent, writes zdot0 as [tile][128][32]
//     enc_grad_kernel    : the same forward sweep, then the reverse sweep of BOTH objectives
//                              O0 = <up1, z0>      O1 = <upt, zdot0> + <upz, z0>
//                          including the phi'' terms of a softplus encoder, reducing the parameter gradients over the
//                          warp with shuffles into one global row per warp; the rows are summed in a fixed order
//                          (enc_rows_reduce_kernel / wgrad_reduce_kernel), so the result is deterministic.
#pragma once
#include <cuda_runtime.h>

#include "pndf_kernel.cuh"

namespace pndf {

struct EncTrainParams {
    const float* x;        // B x 84 poses
    const float* v;        // B x 84 pose tangent (or nullptr: no tangent)
    const float* encw;     // 3516 encoder parameters, reference order
    const float* up1;      // B x 126 or nullptr
    const float* upt;      // B x 126 or nullptr
    const float* upz;      // B x 126 or nullptr
    // the same three vectors formed on the fly from the launch-1 export (no intermediate tensors): g0 = dd/dz0 per pose
    // (row stride g0_ld), up1 = up * coef_b * g0 (coef == nullptr: `uniform` for every pose), upt = w_eik * g0
    const float* g0;       // or nullptr: use up1 / upt above
    long long g0_ld;
    const float* coef;     // [B] or nullptr
    float uniform;
    const float* up;       // device scalar
    const float* weik;     // device scalar or nullptr (no Eikonal objective)
    int in_dim;            // 126 (row stride of upz)
    float* zdot_tiles;     // [tile][128][32] (tangent kernel)
    float* grads;          // grad kernel: per-warp rows [2][warps of the grid][3516] (set 1 rows only if an Eikonal objective is present)
    long long B;
    int normalise, act, use_enc;
    float beta;
};

// value, first and second derivative of the encoder activation
__device__ __forceinline__ void enc_act3(float s, int act, float beta, float& z, float& d1, float& d2) {
    if (act == ACT_SOFTPLUS) {
        z = softplus_eval(s, beta, 1.0f / beta, d1);
        d2 = (s * beta > 20.0f) ? 0.0f : beta * d1 * (1.0f - d1);
    } else {
        const float slope = (act == ACT_RELU) ? 0.0f : 0.01f;
        const bool pos = s > 0.0f;
        z = pos ? s : s * slope;
        d1 = pos ? 1.0f : slope;
        d2 = 0.0f;
    }
}

struct BoneState {
    float u[10], ud[10];       // input and its tangent
    float h[10], hd[10];       // hidden activation and tangent
    float d1h[10], d2h[10];    // phi', phi'' at pre1
    float p1d[10];             // tangent of pre1
    float d1f[6], d2f[6];      // phi', phi'' at pre2
    float p2d[6];              // tangent of pre2
    float f[6], fd[6];
};

// forward + tangent of joint i for one pose; q/qd = normalised pose and tangent (84 each), feat/featd = features so far
__device__ __forceinline__ void bone_fwd_tan(const float* __restrict__ w, int i, int par, const float* q, const float* qd,
                                             const float (*feat)[6], const float (*featd)[6], int act, float beta, BoneState& s) {
    const bool root = par < 0;
    const int fin = root ? 4 : 10;
#pragma unroll
    for (int c = 0; c < 4; ++c) { s.u[c] = q[i * 4 + c]; s.ud[c] = qd[i * 4 + c]; }
#pragma unroll
    for (int r = 0; r < 6; ++r) { s.u[4 + r] = root ? 0.0f : feat[par][r]; s.ud[4 + r] = root ? 0.0f : featd[par][r]; }
    const float* w1 = w; const float* b1 = w + 10 * fin; const float* w2 = b1 + 10; const float* b2 = w2 + 60;
#pragma unroll
    for (int o = 0; o < 10; ++o) {
        float a = __ldg(b1 + o), ad = 0.0f;
        for (int k = 0; k < fin; ++k) { const float ww = __ldg(w1 + o * fin + k); a = fmaf(ww, s.u[k], a); ad = fmaf(ww, s.ud[k], ad); }
        enc_act3(a, act, beta, s.h[o], s.d1h[o], s.d2h[o]);
        s.p1d[o] = ad;
        s.hd[o] = s.d1h[o] * ad;
    }
#pragma unroll
    for (int o = 0; o < 6; ++o) {
        float a = __ldg(b2 + o), ad = 0.0f;
#pragma unroll
        for (int k = 0; k < 10; ++k) { const float ww = __ldg(w2 + o * 10 + k); a = fmaf(ww, s.h[k], a); ad = fmaf(ww, s.hd[k], ad); }
        enc_act3(a, act, beta, s.f[o], s.d1f[o], s.d2f[o]);
        s.p2d[o] = ad;
        s.fd[o] = s.d1f[o] * ad;
    }
}

__device__ __forceinline__ void load_q_qd(const EncTrainParams& p, long long b, float* q, float* qd) {
    float n[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float sq = 0.0f;
        for (int j = 0; j < 21; ++j) { const float xv = p.x[b * 84 + j * 4 + c]; sq = fmaf(xv, xv, sq); }
        n[c] = p.normalise ? fmaxf(sqrtf(sq), 1e-12f) : 1.0f;
    }
    for (int e = 0; e < 84; ++e) q[e] = p.x[b * 84 + e] / n[e & 3];
    if (p.v == nullptr) {
        for (int e = 0; e < 84; ++e) qd[e] = 0.0f;
        return;
    }
    if (!p.normalise) {
        for (int e = 0; e < 84; ++e) qd[e] = p.v[b * 84 + e];
        return;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float dot = 0.0f;
        for (int j = 0; j < 21; ++j) dot = fmaf(q[j * 4 + c], p.v[b * 84 + j * 4 + c], dot);
        for (int j = 0; j < 21; ++j) qd[j * 4 + c] = (p.v[b * 84 + j * 4 + c] - q[j * 4 + c] * dot) / n[c];
    }
}

__global__ void __launch_bounds__(128) enc_tangent_kernel(const EncTrainParams p) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    float q[84], qd[84], feat[21][6], featd[21][6];
    load_q_qd(p, b, q, qd);
    float* out = p.zdot_tiles + (b >> 5) * (128 * 32) + (b & 31);
    if (!p.use_enc) {   // DFNet eats the normalised pose directly (in_dim 84): its tangent is qdot
        for (int e = 0; e < 128; ++e) out[e * 32] = (e < 84) ? qd[e] : 0.0f;
        return;
    }
    for (int i = 0; i < 21; ++i) {
        BoneState s;
        bone_fwd_tan(p.encw + enc_off(i), i, c_parent[i], q, qd, feat, featd, p.act, p.beta, s);
#pragma unroll
        for (int r = 0; r < 6; ++r) { feat[i][r] = s.f[r]; featd[i][r] = s.fd[r]; out[(i * 6 + r) * 32] = s.fd[r]; }
    }
    out[126 * 32] = 0.0f;
    out[127 * 32] = 0.0f;
}

// Transposing warp reduction: every lane holds 32 values v[0..31]; afterwards lane L holds sum over lanes of v[L].
// 31 shuffles per 32 values (a plain butterfly per value would need 160), all of a stage independent of each other.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
            const float keep = up ? v[j + half] : v[j];
            const float send = up ? v[j] : v[j + half];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

// per-pose adjoints of one joint (set 0: first-order objective; set 1: Eikonal objective, value + tangent adjoints)
struct BoneAdj {
    float p2b0[6], p2b1[6], p2db1[6];
    float p1b0[10], p1b1[10], p1db1[10];
};

// e-th parameter of a joint in the reference's own order (W1 [10][fin], b1 [10], W2 [6][10], b2 [6]): this pose's term
template <bool ROOT, int SET>
__device__ __forceinline__ float bone_grad_entry(int e, const BoneState& s, const BoneAdj& a) {
    constexpr int fin = ROOT ? 4 : 10;
    if (e < 10 * fin) {
        const int o = e / fin, k = e % fin;
        return SET == 0 ? a.p1b0[o] * s.u[k] : fmaf(a.p1b1[o], s.u[k], a.p1db1[o] * s.ud[k]);
    }
    e -= 10 * fin;
    if (e < 10) return SET == 0 ? a.p1b0[e] : a.p1b1[e];
    e -= 10;
    if (e < 60) {
        const int o = e / 10, k = e % 10;
        return SET == 0 ? a.p2b0[o] * s.h[k] : fmaf(a.p2b1[o], s.h[k], a.p2db1[o] * s.hd[k]);
    }
    e -= 60;
    if (e < 6) return SET == 0 ? a.p2b0[e] : a.p2b1[e];
    return 0.0f;
}

// reduce this joint's parameter-gradient terms over the warp and store them into this warp's row `acc` (global): every
// parameter is written exactly once per warp, rows are summed in a fixed order afterwards -> deterministic, no atomics
template <bool ROOT, int SET>
__device__ __forceinline__ void bone_grad_reduce(float* acc, const BoneState& s, const BoneAdj& a, int lane) {
    constexpr int n = ROOT ? 116 : 176;
#pragma unroll
    for (int g = 0; g < (n + 31) / 32; ++g) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = bone_grad_entry<ROOT, SET>(g * 32 + j, s, a);
        const float tot = warp_transpose_sum(v, lane);
        if (g * 32 + lane < n) acc[g * 32 + lane] = tot;
    }
}

__global__ void __launch_bounds__(128) enc_grad_kernel(const EncTrainParams p) {
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long wglob = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float* acc = p.grads + wglob * kEncFloats;                         // set 0 row of this warp; set 1 rows follow all set 0 rows
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = b < p.B;
    const long long bb = live ? b : 0;
    const int lane = threadIdx.x & 31;
    const bool from_dump = (p.g0 != nullptr);
    const bool set1 = from_dump ? ((p.weik != nullptr) || (p.upz != nullptr))
                                : ((p.upt != nullptr) || (p.upz != nullptr));     // uniform: the Eikonal objective is present
    float c1 = 0.0f, ce = 0.0f;
    if (from_dump && live) {
        c1 = __ldg(p.up) * (p.coef ? __ldg(p.coef + bb) : p.uniform);
        ce = p.weik ? __ldg(p.weik) : 0.0f;
    }
    float q[84], qd[84], feat[21][6], featd[21][6];
    load_q_qd(p, bb, q, qd);
    for (int i = 0; i < 21; ++i) {
        BoneState s;
        bone_fwd_tan(p.encw + enc_off(i), i, c_parent[i], q, qd, feat, featd, p.act, p.beta, s);
#pragma unroll
        for (int r = 0; r < 6; ++r) { feat[i][r] = s.f[r]; featd[i][r] = s.fd[r]; }
    }
    // adjoints: set 0 = objective O0 (adjoint of f only); set 1 = objective O1 (adjoint of f and of fd)
    float fb0[21][6], fb1[21][6], fdb1[21][6];
    for (int i = 0; i < 21; ++i)
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const long long o = bb * 126 + i * 6 + r;
            if (from_dump) {
                const float g = live ? __ldg(p.g0 + bb * p.g0_ld + i * 6 + r) : 0.0f;
                fb0[i][r] = c1 * g;
                fdb1[i][r] = ce * g;
            } else {
                fb0[i][r] = (live && p.up1) ? p.up1[o] : 0.0f;
                fdb1[i][r] = (live && p.upt) ? p.upt[o] : 0.0f;
            }
            fb1[i][r] = (live && p.upz) ? p.upz[o] : 0.0f;
        }
    for (int i = 20; i >= 0; --i) {
        const int par = c_parent[i];
        const bool root = par < 0;
        const int fin = root ? 4 : 10;
        const int off = enc_off(i);
        const float* w1 = p.encw + off;
        const float* w2 = w1 + 10 * fin + 10;
        BoneState s;
        bone_fwd_tan(p.encw + off, i, par, q, qd, feat, featd, p.act, p.beta, s);
        BoneAdj a;
        // layer 2
#pragma unroll
        for (int o = 0; o < 6; ++o) {
            a.p2b0[o] = fb0[i][o] * s.d1f[o];
            a.p2db1[o] = fdb1[i][o] * s.d1f[o];
            a.p2b1[o] = fb1[i][o] * s.d1f[o] + fdb1[i][o] * s.d2f[o] * s.p2d[o];
        }
        float hb0[10], hb1[10], hdb1[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) { hb0[k] = 0.0f; hb1[k] = 0.0f; hdb1[k] = 0.0f; }
#pragma unroll
        for (int o = 0; o < 6; ++o)
#pragma unroll
            for (int k = 0; k < 10; ++k) {
                const float ww = __ldg(w2 + o * 10 + k);
                hb0[k] = fmaf(ww, a.p2b0[o], hb0[k]); hb1[k] = fmaf(ww, a.p2b1[o], hb1[k]); hdb1[k] = fmaf(ww, a.p2db1[o], hdb1[k]);
            }
        // layer 1
#pragma unroll
        for (int o = 0; o < 10; ++o) {
            a.p1b0[o] = hb0[o] * s.d1h[o];
            a.p1db1[o] = hdb1[o] * s.d1h[o];
            a.p1b1[o] = hb1[o] * s.d1h[o] + hdb1[o] * s.d2h[o] * s.p1d[o];
        }
        if (!root) {
            float ub0[6], ub1[6], udb1[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) { ub0[r] = 0.0f; ub1[r] = 0.0f; udb1[r] = 0.0f; }
#pragma unroll
            for (int o = 0; o < 10; ++o)
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const float ww = __ldg(w1 + o * 10 + 4 + r);
                    ub0[r] = fmaf(ww, a.p1b0[o], ub0[r]); ub1[r] = fmaf(ww, a.p1b1[o], ub1[r]); udb1[r] = fmaf(ww, a.p1db1[o], udb1[r]);
                }
#pragma unroll
            for (int r = 0; r < 6; ++r) { fb0[par][r] += ub0[r]; fb1[par][r] += ub1[r]; fdb1[par][r] += udb1[r]; }
        }
        // parameter gradients of this joint, reduced over the 32 poses of the warp
        if (root) {
            bone_grad_reduce<true, 0>(acc + off, s, a, lane);
            if (set1) bone_grad_reduce<true, 1>(acc + nwarps * kEncFloats + off, s, a, lane);
        } else {
            bone_grad_reduce<false, 0>(acc + off, s, a, lane);
            if (set1) bone_grad_reduce<false, 1>(acc + nwarps * kEncFloats + off, s, a, lane);
        }
    }
}

// first level of the fixed-order reduction: part[set][chunk][i] = sum of rows[set][w][i] over the warps w of chunk
// (chunk = blockIdx.y of kEncChunks); the second level (kEncChunks x sets rows) is folded into wgrad_reduce_kernel
constexpr int kEncChunks = 16;
__global__ void __launch_bounds__(256) enc_rows_partial_kernel(const float* __restrict__ rows, long long nwarps, int nsets,
                                                              float* __restrict__ part) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsets * kEncFloats) return;
    const int set = i / kEncFloats, e = i - set * kEncFloats;
    const long long per = (nwarps + kEncChunks - 1) / kEncChunks;
    const long long w0 = (long long)blockIdx.y * per, w1 = min(nwarps, w0 + per);
    const float* r = rows + (size_t)set * nwarps * kEncFloats + e;
    float s = 0.0f;
    for (long long w = w0; w < w1; ++w) s += r[(size_t)w * kEncFloats];
    part[((size_t)set * kEncChunks + blockIdx.y) * kEncFloats + e] = s;
}

// out[set][i] = sum over warps of rows[set][warp][i], fixed order
__global__ void __launch_bounds__(256) enc_rows_reduce_kernel(const float* __restrict__ rows, long long nwarps, int nsets,
                                                             float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsets * kEncFloats) return;
    const int set = i / kEncFloats, e = i - set * kEncFloats;
    const float* r = rows + (size_t)set * nwarps * kEncFloats + e;
    float s = 0.0f;
    for (long long w = 0; w < nwarps; ++w) s += r[(size_t)w * kEncFloats];
    out[i] = s;
}

}  // namespace pndf
