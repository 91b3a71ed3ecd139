// pndf_kernel.cuh -- the fused PoseNDF kernel for sm_100a (B200).
//
// One persistent CTA per SM.  A CTA owns a tile of 32 poses and walks the WHOLE network for it without
// touching HBM in between:
//
//   load 32 poses -> column-normalise (model/posendf.py:71) -> structure encoder along the kinematic tree
//   (model/network/net_modules.py:140-170) -> DFNet 126->256->512->1024->512->256->64->1
//   (net_modules.py:46-72) -> analytic reverse pass (replaces torch.autograd.grad, model/posendf.py:18-27)
//   -> normalise Jacobian -> x <- x - d * dd/dx (experiments/sample_poses.py:74), optionally K times.
//
// The DFNet layers (99.8 % of the flops) run as register-tiled fp32 GEMMs on the packed FFMA2 pipe: 256 threads,
// each owning an 8-pose x 8-feature micro-tile; activations live in shared memory as [feature][pose] with a 16-byte
// XOR swizzle; every warp streams ONLY the 64 feature columns it multiplies through its own private 2-stage ring of
// 4 KB slabs (cp.async.bulk / TMA 1-D copies completing on per-warp mbarriers, refilled by the warp itself).  The
// host pre-packs all weights into eight interleaved per-warp streams in exactly the order a tile consumes them
// (forward layout W^T[k][n] for the forward ops, native W[out][in] for the reverse ops).  The 1024-wide layer never
// exists in full: L2's output is produced in two 512-feature chunks that are consumed immediately as K-chunks of L3
// (accumulators stay in registers); the reverse pass mirrors this; the 256-wide ops run split-K.  Activation
// derivatives are 1 bit per unit in shared memory for relu / lrelu and fp32 in an L2-resident per-CTA scratch for
// softplus.  MODE 2 replays the forward ops as a forward-mode tangent pass (training, Eikonal term).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pndf {

constexpr int kTileM = 32;            // poses per CTA tile
constexpr int kGemmThreads = 256;     // 8 warps, all of them compute (there is no producer warp)
constexpr int kThreads = 256;         // (256 threads -> 255 registers/thread for the 2x64 fused accumulators)
// Weight streaming: every warp owns a PRIVATE 2-stage ring of 4 KB slabs holding only the 64 (16, 8) feature columns
// that warp multiplies -- R = 1024 / columns rows per slab.  The warp that consumed a slab refills it itself
// (elected lane, cp.async.bulk), so there is no cross-warp "empty" handshake and no producer rotation at all;
// warps only meet at the per-op __syncthreads.
constexpr int kSlabFloats = 1024;     // per-warp slab: R rows x FW feature columns, R*FW == 1024 (4 KB)
constexpr int kSlabBytes = kSlabFloats * 4;
constexpr int kStages = 2;            // per-warp stages
constexpr int kWarps = 8;
// per-warp slab count of an op with K reduction rows split over KG K-groups and FW columns per warp
__host__ __device__ constexpr int slabs_of(int K, int KG, int FW) { return (K / KG) * FW / kSlabFloats; }
constexpr int kXS = 84;               // row stride of the pose tile: the tile is a plain image of 32 x 336 bytes of the batch (bulk copies)
constexpr int kMaskStride = 2656;     // bytes per pose-group plane of the derivative bit masks
constexpr int kUnits = 2624;          // hidden units of the DFNet (256+512+1024+512+256+64)
constexpr int kEncFloats = 3516;
constexpr int kMaxPeers = 7;          // other GPUs of one NVSwitch node
constexpr int kEncStageRow = 256;     // encoder weights are parked in X rows [256, 366) while the encoder runs

// unit offsets (mask / scratch index) of each hidden layer output z1..z6
constexpr int kU1 = 0, kU2 = 256, kU3 = 768, kU4 = 1792, kU5 = 2304, kU6 = 2560;

// shared memory carve-up (bytes)
constexpr int kSmX = 0;
constexpr int kSmY = kSmX + 512 * 32 * 4;
constexpr int kSmRing = kSmY + 512 * 32 * 4;
constexpr int kSmMask = kSmRing + kWarps * kStages * kSlabBytes;
constexpr int kSmXs = kSmMask + 4 * kMaskStride;
constexpr int kSmQs = kSmXs + kTileM * kXS * 4;      // column-normalised poses q = x / n (what the encoder eats)
constexpr int kSmNrm = kSmQs + kTileM * kXS * 4;
constexpr int kSmDv = kSmNrm + 4 * 32 * 4;
constexpr int kSmBar = kSmDv + 2 * 32 * 4;
constexpr int kSmTotal = kSmBar + kWarps * kStages * 8 + 16;

// debug dump row offsets ([row][32 poses] floats)
constexpr int kDumpRows = 5504;

enum { ACT_RELU = 0, ACT_LRELU = 1, ACT_SOFTPLUS = 2 };
enum { IN_QUAT = 0, IN_AXIS_ANGLE = 1 };

// ---- motion-denoise loop (experiments/motion_denoise.py:70-99), fused into the prior launch
struct AdamParams {
    float lr, beta1, beta2, eps;
    float bias1, bias2;      // 1 - beta^t for this step
    float weight;            // 1e7 / (1 + it)
};
// torch.optim.Adam's update of one axis-angle component with gradient g = scale * (raw d dist / d aa), same op order as
// seq_adam_kernel has always used (pndf_denoise.cuh)
__device__ __forceinline__ void dn_adam_update(float& a, float& m, float& v, float graw, float scale, const AdamParams& ap) {
    const float g = scale * graw;
    m = ap.beta1 * m + (1.0f - ap.beta1) * g;
    v = ap.beta2 * v + (1.0f - ap.beta2) * g * g;
    const float denom = sqrtf(v) * (1.0f / sqrtf(ap.bias2)) + ap.eps;
    a -= (ap.lr / ap.bias1) * (m / denom);
}
// One launch per optimisation step: the launch of step t FIRST applies the Adam update that step t-1's gradient asked for
// (needs mean_t dist of whole sequences -> only known once launch t-1 has finished everywhere; read from dist_prev), then
// evaluates prior + gradient at the updated poses.  The last update is applied by seq_adam_kernel.
struct DenoiseFuse {
    float* m;                 // Adam moments, [B][63]
    float* v;
    const float* graw;        // raw gradient of the previous step, [B][63] (this launch overwrites it through KParams::grad)
    const float* dist_prev;   // distances of the previous step, [S][T]
    float* loss_out;          // [S] loss of the previous step (weight * mean(dist)^2) or nullptr
    float* pose_rw;           // the axis-angle poses, updated in place
    int T;                    // frames per sequence
    int pending;              // 0: first step, nothing to apply
    AdamParams ap;            // of the pending update
};

struct KParams {
    const float* wstream;     // slab stream (forward ops then reverse ops)
    const float* bias[7];     // dfnet.lin{l}.bias
    const float* w6;          // dfnet.lin6.weight (64)
    const float* encw;        // encoder params, reference order (3516 floats) or nullptr
    const float* pose_in;     // B x 84 (or B x 63 axis-angle in prior mode)
    float* pose_out;          // B x 84 or nullptr
    float* dist;              // B or nullptr
    float* grad;              // B x 84 (or B x 63 in prior mode) or nullptr
    const float* g_up;        // B or nullptr
    float* dscratch;          // per-CTA fp32 derivative scratch (softplus) or nullptr
    float* z0scratch;         // per-CTA stash of the encoder features (128 x 32 floats), L2 resident
    float* dbg;               // dump of every intermediate tile ([row][32 poses], kDumpRows rows per tile) or nullptr
    uint8_t* act_masks;       // activation-derivative handoff, MODE 1 writes / MODE 2 reads (and skips its primal pass):
                              // relu, lrelu: [tile][4][kMaskStride] bit masks; softplus: [tile][kUnits][32] fp32
    int dump_all;             // 0: first tile only (debug hook)   1: every tile (training: exports for the weight gradients)
    const float* tan_in;      // MODE 2: tangent of the DFNet input, [tile][128][32] floats
    // fused gather (multi-GPU projection runs): the write-back also stores every projected tile -- and its distances --
    // into the gathered buffers of up to kMaxPeers other GPUs through NVLink-mapped (cudaIpc) pointers, each already
    // offset to this rank's slice.  The transfer rides under the FMA work of the following tiles.
    DenoiseFuse dn;
    float* peer_pose[kMaxPeers];
    float* peer_dist[kMaxPeers];
    int n_peers;
    long long B;
    int ntiles;
    int steps;                // projection steps fused in this launch (>=1)
    int do_step;              // apply x <- x - d*g
    int renorm;               // per-quaternion renormalise after the step
    int normalise;            // F.normalize(dim=1) on input
    int input_kind;           // IN_QUAT / IN_AXIS_ANGLE
    int use_enc, enc_act, df_act;
    float enc_beta, df_beta;
    int f0_slabs;             // slabs of the first forward op (z0 rows / 16)
    int z0_rows;              // 128 (encoder) or 96 (raw 84 + pad)
    int in_dim;               // 126 or 84
};

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void gemm_bar() { __syncthreads(); }

// ------------------------------------------------------------------------------------------------ helpers
// activations are stored [row = feature][32 poses]; the 16-byte chunk (4 poses) index is XOR-ed with
// (row>>2)&7 so that the 8 lanes of a warp that own different feature groups hit different banks.
__device__ __forceinline__ int swz(int row, int m) { return row * 32 + (((((m >> 2) ^ (row >> 2)) & 7) << 2) | (m & 3)); }

// nn.Softplus(beta) with torch's threshold 20, branch-free (the GEMM epilogues evaluate it 84 000 times per tile):
// value log1p(exp(beta x))/beta and derivative exp/(1+exp) = torch's softplus_backward z/(z+1); above the threshold
// both are selected to x and 1.  exp is clamped so the unused lane of the select never overflows.
__device__ __forceinline__ float softplus_eval(float v, float beta, float inv_beta, float& deriv) {
    const float bx = v * beta;
    const bool lin = bx > 20.0f;
    const float e = expf(fminf(bx, 20.0f));
    const float r = __frcp_rn(e + 1.0f);
    deriv = lin ? 1.0f : e * r;
    return lin ? v : log1pf(e) * inv_beta;
}
// The same on the SFU (the GEMM epilogues evaluate it 84 000 times per 32-pose tile): e = 2^(bx log2 e) (ex2.approx, 2 ulp),
// sigma = e / (1 + e) through rcp.approx (1 ulp), log1p(e) = ln2 * lg2(1 + e) (lg2.approx: absolute error < 2^-22 near 1, i.e.
// < 2e-7 ln2 / beta on the activation; a short series below e = 2^-6).  Validated against the reference's fp64 goldens at the 1e-5 bar for beta = 5, 30,
// 100 (tests/test_gpu_parity.py); the exact version above stays in use for the scalar output unit and the encoder.
__device__ __forceinline__ float softplus_fast(float v, float beta, float inv_beta, float& deriv) {
    const float bx = v * beta;
    const bool lin = bx > 20.0f;
    float e, r, l;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(bx, 20.0f) * 1.4426950408889634f));
    const float u = 1.0f + e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(u));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u));
    deriv = lin ? 1.0f : e * r;
    // small e: lg2(1 + e) would lose e to the rounding of 1 + e (and return 0 below 2^-24, breaking z > 0 and the relation
    // sigma = 1 - exp(-beta z) the second-order training chain relies on): three series terms, relative error < e^3 / 4
    const float ln1p = (e < 0.015625f) ? e * fmaf(e, fmaf(e, 0.33333334f, -0.5f), 1.0f) : l * 0.6931471805599453f;
    return lin ? v : ln1p * inv_beta;
}

__device__ __forceinline__ float act_eval(float v, int kind, float beta, float& deriv) {
    if (kind == ACT_SOFTPLUS) return softplus_eval(v, beta, 1.0f / beta, deriv);
    float slope = (kind == ACT_RELU) ? 0.0f : 0.01f;
    bool pos = v > 0.0f;
    deriv = pos ? 1.0f : slope;
    return pos ? v : v * slope;
}

// Per-warp pipeline state (identical in all lanes of the warp).
// compile-time variant for the encoder: SOFT = softplus(beta = par), else piecewise-linear with slope = par
template <bool SOFT>
__device__ __forceinline__ float act_t(float v, float par, float& deriv) {
    if (SOFT) return act_eval(v, ACT_SOFTPLUS, par, deriv);
    const bool pos = v > 0.0f;
    deriv = pos ? 1.0f : par;
    return pos ? v : v * par;
}

struct Pipe {
    uint32_t stage;       // ring slot of the slab being consumed
    uint32_t phase;       // its mbarrier phase parity
    uint32_t left;        // slabs of this warp not yet issued
    uint32_t pos;         // position (in per-warp slabs) of the next slab to issue inside the per-step stream
    uint32_t step_slabs;  // per-warp slabs per network pass
    const char* wsrc;     // this warp's slab stream base: slab i lives at wsrc + i * kWarps * kSlabBytes
    __device__ __forceinline__ void advance() {
        stage ^= 1u;
        phase ^= (stage == 0) ? 1u : 0u;
    }
};

struct Ctx {
    float* X;
    float* Y;
    const float* ring;   // THIS WARP's private ring (2 x 4 KB)
    uint32_t ring_s;
    uint8_t* mask;
    uint32_t full_s;   // shared-space address of this warp's two 'slab landed' mbarriers
    float* dscr;   // this CTA's derivative scratch (softplus) or nullptr
    int tid, lane, mg, ng;
    int ng2, kg;   // split-K ops (N = 256): feature group within a 4-warp K-group, and the K-group (0/1)
    int kq;        // small-tile kernels (KS): this lane group's residue of the reduction rows, lane >> 3 (then mg == 0)
    float slope;                       // relu 0 / lrelu 0.01 (piecewise-linear DFNet activation)
    float df_beta, df_inv_beta;        // softplus DFNet: beta, 1 / beta
};

// KG = number of K-groups an op is split into: KG == 1, all 8 warps tile N = 64*TN features; KG == 2 (split-K, used for
// the 256-wide ops) two groups of 4 warps each tile the SAME N = 32*TN features over one half of every slab's rows.
template <int KG>
__device__ __forceinline__ int ngv(const Ctx& c) { return KG == 1 ? c.ng : c.ng2; }

template <int TN, int KG = 1>
__device__ __forceinline__ int feat_of(int ng, int j) {
    if (TN == 8) return (j < 4) ? (ng * 4 + j) : (32 * TN / KG + ng * 4 + (j - 4));
    if (TN == 4) return ng * 4 + j;
    if (TN == 2) return ng * 2 + j;
    return ng;
}

template <int TN, int KG = 1>
__device__ __forceinline__ void acc_init_bias(float (&acc)[8][TN], const float* __restrict__ bias, int ng, int nreal) {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        int f = feat_of<TN, KG>(ng, j);
        float b = (f < nreal) ? __ldg(bias + f) : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][j] = b;
    }
}
template <int TN>
__device__ __forceinline__ void acc_zero(float (&acc)[8][TN]) {
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][j] = 0.0f;
}

__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_s(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
// 1-D bulk copies (TMA) global -> shared completing on an mbarrier, shared -> global in a bulk group
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
// The warp has finished reading ring slot `stage` (caller did __syncwarp): refill it with the warp's next slab.
__device__ __forceinline__ void refill(Pipe& pipe, const Ctx& c, uint32_t stage) {
    if (pipe.left != 0) {
        if (c.lane == 0) {
            const uint32_t bar = c.full_s + stage * 8;
            mbar_expect_tx_s(bar, kSlabBytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             c.ring_s + stage * kSlabBytes),
                         "l"(pipe.wsrc + (size_t)pipe.pos * (kWarps * kSlabBytes)), "r"((uint32_t)kSlabBytes), "r"(bar)
                         : "memory");
        }
        --pipe.left;
        if (++pipe.pos == pipe.step_slabs) pipe.pos = 0;
    }
}

template <int TN>
struct Operands {
    float a[8];
    float b[TN];
};

// operands of one k-step: 8 pose values of this thread's pose group (row k of the activation tile), TN weights of
// its feature group (row `r` of the warp's slab, FW columns: [32 low features | 32 high features] for TN == 8)
template <int TN>
__device__ __forceinline__ void load_operands(Operands<TN>& o, const float* __restrict__ in, int k, const float* __restrict__ w,
                                              int r, const Ctx& c) {
    constexpr int FW = 8 * TN;
    const int ngl = c.lane & 7;
    const int key = (k >> 2) & 7;
    const float* row = in + k * 32;
    const float4 a0 = *reinterpret_cast<const float4*>(row + (((c.mg * 2) ^ key) << 2));
    const float4 a1 = *reinterpret_cast<const float4*>(row + (((c.mg * 2 + 1) ^ key) << 2));
    o.a[0] = a0.x; o.a[1] = a0.y; o.a[2] = a0.z; o.a[3] = a0.w;
    o.a[4] = a1.x; o.a[5] = a1.y; o.a[6] = a1.z; o.a[7] = a1.w;
    if (TN == 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(w + r * FW + ngl * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(w + r * FW + 32 + ngl * 4);
        o.b[0] = b0.x; o.b[1 % TN] = b0.y; o.b[2 % TN] = b0.z; o.b[3 % TN] = b0.w;
        o.b[4 % TN] = b1.x; o.b[5 % TN] = b1.y; o.b[6 % TN] = b1.z; o.b[7 % TN] = b1.w;
    } else if (TN == 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(w + r * FW + ngl * 4);
        o.b[0] = b0.x; o.b[1 % TN] = b0.y; o.b[2 % TN] = b0.z; o.b[3 % TN] = b0.w;
    } else if (TN == 2) {
        const float2 b0 = *reinterpret_cast<const float2*>(w + r * FW + ngl * 2);
        o.b[0] = b0.x; o.b[1 % TN] = b0.y;
    } else {
        o.b[0] = w[r * FW + ngl];
    }
}

template <int TN>
__device__ __forceinline__ void fma_step(float (&acc)[8][TN], const Operands<TN>& o) {
    if (TN >= 2) {
        // packed fp32x2 FMA (Blackwell FFMA2): the pose value is the scalar-broadcast operand, two adjacent
        // features ride in one 64-bit register pair -> half the issue slots of scalar FFMA, same rounding.
        // Loop order matters: with the feature PAIR outermost the 64-bit b operand is the one ptxas keeps in the operand
        // reuse cache across 8 consecutive FFMA2s (measured 68.7 vs 63.4 TFLOP/s for the pose scalar outermost).
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) {
            unsigned long long bb;
            asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(o.b[2 * j]), "f"(o.b[(2 * j + 1) % TN]));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                unsigned long long aa, cc;
                asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(o.a[i]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(acc[i][2 * j]), "f"(acc[i][(2 * j + 1) % TN]));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(aa), "l"(bb));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i][2 * j]), "=f"(acc[i][(2 * j + 1) % TN]) : "l"(cc));
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(o.a[i], o.b[j], acc[i][j]);
    }
}

// acc[8 poses][TN feats] += in[k][pose] * w[k][feat] over `nslabs` slabs of this warp's private weight stream
// (R = 1024/(8*TN) reduction rows each; with KG == 2 the warp's K-group covers rows [kg*K/2, (kg+1)*K/2)).
// The slab body is straight-line code; the next slab's barrier is probed a few rows in, so the ~90-cycle mbarrier
// round trip hides under the FMAs; the refill of the slab just consumed is issued by the warp itself.
// KS (small-tile kernels, 8 poses per tile): the four 8-lane groups of a warp no longer own different poses -- all of them
// multiply the SAME 8 poses (c.mg == 0) and split the reduction rows instead, lane group kq takes rows kq, kq + 4, ... of every
// slab; kq_reduce() then sums the four partial tiles with two shuffles per accumulator.
template <int TN, int KG = 1, bool KS = false>
__device__ __forceinline__ void gemm_op(float (&acc)[8][TN], const float* __restrict__ in, int nslabs, Pipe& pipe,
                                        const Ctx& c) {
    constexpr int R = kSlabFloats / (8 * TN);
    constexpr int RT = KS ? R / 4 : R;      // rows of a slab this thread multiplies
    constexpr int RS = KS ? 4 : 1;          // their stride
    const int kbase = (KG == 1) ? 0 : c.kg * (nslabs * R);
    bool ready = mbar_try_wait_s(c.full_s + pipe.stage * 8, pipe.phase);
    for (int s = 0; s < nslabs; ++s) {
        if (!ready) mbar_wait_s(c.full_s + pipe.stage * 8, pipe.phase);
        const float* __restrict__ w = c.ring + pipe.stage * kSlabFloats;
        const uint32_t cur = pipe.stage;
        pipe.advance();
        if (TN == 8) {
            // 16 rows starting at a multiple of 16: the swizzle key (row>>2)&7 is kb, kb+1, kb+2, kb+3 with kb in {0,4},
            // so the thread's two 16-byte pose chunks sit at ((2mg ^ kb) ^ j) and that ^ 1 for row quad j: four offsets per
            // slab, every row is then base + immediate -- no address arithmetic between the FFMA2s.
            const int k0 = kbase + s * R;
            const float* __restrict__ rows = in + (k0 + (KS ? c.kq : 0)) * 32;
            const float* __restrict__ wk = w + (KS ? c.kq * 64 : 0);
            const int cb = (c.mg * 2) ^ ((k0 >> 2) & 4);
            const int ngl4 = (c.lane & 7) * 4;
#pragma unroll
            for (int rr = 0; rr < RT; ++rr) {
                constexpr int kProbe = KS ? 1 : 4;
                if (rr == kProbe) ready = mbar_try_wait_s(c.full_s + pipe.stage * 8, pipe.phase);
                const int r = rr * RS;      // (+ kq, folded into the base pointers: the swizzle key only depends on r >> 2)
                const int ca = (cb ^ (r >> 2)) << 2;
                const float4 a0 = *reinterpret_cast<const float4*>(rows + r * 32 + ca);
                const float4 a1 = *reinterpret_cast<const float4*>(rows + r * 32 + (ca ^ 4));
                const float4 b0 = *reinterpret_cast<const float4*>(wk + r * 64 + ngl4);
                const float4 b1 = *reinterpret_cast<const float4*>(wk + r * 64 + 32 + ngl4);
                Operands<TN> o;
                o.a[0] = a0.x; o.a[1] = a0.y; o.a[2] = a0.z; o.a[3] = a0.w;
                o.a[4] = a1.x; o.a[5] = a1.y; o.a[6] = a1.z; o.a[7] = a1.w;
                o.b[0] = b0.x; o.b[1 % TN] = b0.y; o.b[2 % TN] = b0.z; o.b[3 % TN] = b0.w;
                o.b[4 % TN] = b1.x; o.b[5 % TN] = b1.y; o.b[6 % TN] = b1.z; o.b[7 % TN] = b1.w;
                fma_step<TN>(acc, o);
            }
        } else {
#pragma unroll (RT > 32 ? 32 : RT)
            for (int rr = 0; rr < RT; ++rr) {
                if (rr == 4) ready = mbar_try_wait_s(c.full_s + pipe.stage * 8, pipe.phase);
                const int r = rr * RS + (KS ? c.kq : 0);
                Operands<TN> o;
                load_operands<TN>(o, in, kbase + s * R + r, w, r, c);
                fma_step<TN>(acc, o);
            }
        }
        __syncwarp();
        refill(pipe, c, cur);
    }
}

// KS: sum the four lane groups' partial tiles (every lane ends up with the full sums of its 8 poses x TN features)
template <int TN, bool KS>
__device__ __forceinline__ void kq_reduce(float (&acc)[8][TN]) {
    if (!KS) return;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float v = acc[i][j];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            acc[i][j] = v;
        }
}
// KS: after kq_reduce the four lane groups hold identical tiles; feature j of the thread's TN is finished (epilogue, stores)
// by exactly one of them
template <int TN, bool KS>
__device__ __forceinline__ bool owns(int j, int kq) {
    if (!KS) return true;
    if (TN == 8) return (j >> 1) == kq;
    if (TN == 2) return j == kq;
    return kq == 0;
}

// write one feature row segment (8 poses of this thread) into a [feature][pose] buffer
__device__ __forceinline__ void store_row8(float* buf, int f, int mg, const float (&v)[8]) {
    const int key = (f >> 2) & 7;
    float* row = buf + f * 32;
    *reinterpret_cast<float4*>(row + (((mg * 2) ^ key) << 2)) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(row + (((mg * 2 + 1) ^ key) << 2)) = make_float4(v[4], v[5], v[6], v[7]);
}

template <int TN, int KG = 1>
__device__ __forceinline__ void mask_store(uint8_t* mask, int mg, int unit0, const uint32_t (&bits)[TN]) {
    // units of one thread are consecutive in groups of min(TN,4)
    uint8_t* base = mask + mg * kMaskStride;
    if (TN >= 4) {
#pragma unroll
        for (int h = 0; h < TN / 4; ++h) {
            uint32_t wv = bits[h * 4] | (bits[h * 4 + 1] << 8) | (bits[h * 4 + 2] << 16) | (bits[h * 4 + 3] << 24);
            *reinterpret_cast<uint32_t*>(base + unit0 + h * (32 * TN / KG)) = wv;
        }
    } else if (TN == 2) {
        *reinterpret_cast<uint16_t*>(base + unit0) = (uint16_t)(bits[0] | (bits[1 % TN] << 8));
    } else {
        base[unit0] = (uint8_t)bits[0];
    }
}
template <int TN, int KG = 1>
__device__ __forceinline__ void mask_load(const uint8_t* mask, int mg, int unit0, uint32_t (&bits)[TN]) {
    const uint8_t* base = mask + mg * kMaskStride;
    if (TN >= 4) {
#pragma unroll
        for (int h = 0; h < TN / 4; ++h) {
            uint32_t wv = *reinterpret_cast<const uint32_t*>(base + unit0 + h * (32 * TN / KG));
            bits[h * 4] = wv & 0xff; bits[h * 4 + 1] = (wv >> 8) & 0xff; bits[h * 4 + 2] = (wv >> 16) & 0xff; bits[h * 4 + 3] = wv >> 24;
        }
    } else if (TN == 2) {
        uint32_t wv = *reinterpret_cast<const uint16_t*>(base + unit0);
        bits[0] = wv & 0xff; bits[1 % TN] = wv >> 8;
    } else {
        bits[0] = base[unit0];
    }
}

// forward epilogue: z = act(acc) (bias already in acc), remember the derivative, store z as next input.
// unit_base: index of feature 0 of this op in the mask / scratch unit space.
template <bool SOFT, int TN, int KG = 1, bool KS = false>
__device__ __forceinline__ void epilogue_fwd(const float (&acc)[8][TN], float* out, int unit_base, const Ctx& c, bool keep_deriv) {
    const int ng = ngv<KG>(c);
    const int unit0 = unit_base + feat_of<TN, KG>(ng, 0);
    if (SOFT) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (!owns<TN, KS>(j, c.kq)) continue;
            const int f = feat_of<TN, KG>(ng, j);
            float z[8], dv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = softplus_fast(acc[i][j], c.df_beta, c.df_inv_beta, dv[i]);
            store_row8(out, f, c.mg, z);
            if (keep_deriv) {
                float* p = c.dscr + (size_t)(unit_base + f) * 32 + c.mg * 8;
                *reinterpret_cast<float4*>(p) = make_float4(dv[0], dv[1], dv[2], dv[3]);
                *reinterpret_cast<float4*>(p + 4) = make_float4(dv[4], dv[5], dv[6], dv[7]);
            }
        }
    } else {
        const float slope = c.slope;
        uint32_t bits[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (!owns<TN, KS>(j, c.kq)) continue;
            const int f = feat_of<TN, KG>(ng, j);
            float z[8];
            uint32_t bm = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float v = acc[i][j];
                const bool pos = v > 0.0f;
                bm |= (pos ? 1u : 0u) << i;
                z[i] = pos ? v : v * slope;
            }
            bits[j] = bm;
            store_row8(out, f, c.mg, z);
            if (KS && keep_deriv) c.mask[unit_base + f] = (uint8_t)bm;      // plane 0, one byte per unit
        }
        if (!KS && keep_deriv) mask_store<TN, KG>(c.mask, c.mg, unit0, bits);
    }
}

// reverse epilogue: g = acc * act'(pre) of the layer whose input-gradient this op produced; store as the
// next reverse op's input.  unit_base < 0: no derivative (the encoder features, handled by the encoder).
template <bool SOFT, int TN, int KG = 1, bool KS = false>
__device__ __forceinline__ void epilogue_bwd(const float (&acc)[8][TN], float* out, int unit_base, const Ctx& c) {
    const int ng = ngv<KG>(c);
    if (unit_base < 0) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (!owns<TN, KS>(j, c.kq)) continue;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = acc[i][j];
            store_row8(out, feat_of<TN, KG>(ng, j), c.mg, v);
        }
        return;
    }
    if (SOFT) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (!owns<TN, KS>(j, c.kq)) continue;
            const int f = feat_of<TN, KG>(ng, j);
            const float* p = c.dscr + (size_t)(unit_base + f) * 32 + c.mg * 8;
            const float4 d0 = *reinterpret_cast<const float4*>(p);
            const float4 d1 = *reinterpret_cast<const float4*>(p + 4);
            const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = acc[i][j] * dv[i];
            store_row8(out, f, c.mg, v);
        }
    } else {
        const float slope = c.slope;
        uint32_t bits[TN];
        if (!KS) mask_load<TN, KG>(c.mask, c.mg, unit_base + feat_of<TN, KG>(ng, 0), bits);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (!owns<TN, KS>(j, c.kq)) continue;
            if (KS) bits[j] = c.mask[unit_base + feat_of<TN, KG>(ng, j)];
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = ((bits[j] >> i) & 1u) ? acc[i][j] : acc[i][j] * slope;
            store_row8(out, feat_of<TN, KG>(ng, j), c.mg, v);
        }
    }
}

// split-K ops: K-group 1 parks its partial sums in the op's output tile, K-group 0 adds them to its own and then
// runs the epilogue over the same elements (same thread <-> element mapping in both groups).
template <int TN, bool KS = false>
__device__ __forceinline__ void splitk_combine(float (&acc)[8][TN], float* out, const Ctx& c) {
    if (c.kg == 1) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (!owns<TN, KS>(j, c.kq)) continue;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = acc[i][j];
            store_row8(out, feat_of<TN, 2>(c.ng2, j), c.mg, v);
        }
    }
    gemm_bar();
    if (c.kg == 0) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (!owns<TN, KS>(j, c.kq)) continue;
            const int f = feat_of<TN, 2>(c.ng2, j);
            const int key = (f >> 2) & 7;
            const float* row = out + f * 32;
            const float4 p0 = *reinterpret_cast<const float4*>(row + (((c.mg * 2) ^ key) << 2));
            const float4 p1 = *reinterpret_cast<const float4*>(row + (((c.mg * 2 + 1) ^ key) << 2));
            acc[0][j] += p0.x; acc[1][j] += p0.y; acc[2][j] += p0.z; acc[3][j] += p0.w;
            acc[4][j] += p1.x; acc[5][j] += p1.y; acc[6][j] += p1.z; acc[7][j] += p1.w;
        }
    }
}

// Export a [feature][pose] tile as rows of the pose-major dump: dump[pose][kDumpRows] (so that the host sees plain
// strided (B x width) matrices, no repacking before the weight-gradient GEMMs).  `rows` is a multiple of 32 here, a warp
// writes 32 consecutive features of one pose (coalesced); the transposed shared-memory read is 4-way bank conflicted,
// which does not matter next to the GEMMs.
// (rows is a multiple of 32.)  One warp iteration moves a 32-feature x 4-pose block with 8 lanes per pose row: lane
// (pq, fl) reads features 4 fl .. 4 fl + 3 of pose 4 mq + pq (four scalar LDS, conflict-free under the swizzle) and writes
// them with one 16-byte store, so a warp store covers four full 128-byte lines.  The stores are streaming (evict-first):
// 22 KB per pose would otherwise push the L2-resident weight stream out of the cache -- measured on B200, 32 768 poses:
// 3.89 ms with plain stores, 3.26 ms with st.global.cs, 3.21 ms without any export.
__device__ __forceinline__ void dump_rows(float* dbg, int row0, const float* buf, int rows, int tid) {
    if (dbg == nullptr) return;
    const int warp = tid >> 5, lane = tid & 31;
    const int pq = lane >> 3, fl = lane & 7;
    const int nblk = (rows >> 5) * 8;
    for (int c = warp; c < nblk; c += kWarps) {
        const int r = ((c >> 3) << 5) + 4 * fl, m = (c & 7) * 4 + pq;
        const float* src = buf + r * 32 + (((((m >> 2) ^ (r >> 2)) & 7) << 2) | (m & 3));
        float4 v;
        v.x = src[0]; v.y = src[32]; v.z = src[64]; v.w = src[96];
        __stcs(reinterpret_cast<float4*>(dbg + (size_t)m * kDumpRows + row0 + r), v);
    }
}

// ------------------------------------------------------------------------------------------------ encoder
static __constant__ int c_parent[21] = {-1, -1, -1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19};
// non-root joint pairs of the encoder schedule (two independent sub-trees per step), after (1,0) and (3,2)
static __constant__ int c_pair_a[7] = {5, 7, 9, 12, 14, 16, 17};
static __constant__ int c_pair_b[7] = {4, 6, 8, 10, 11, 13, 15};

__device__ __forceinline__ int enc_off(int i) { return (i < 3) ? i * 116 : 348 + (i - 3) * 176; }

// The encoder runs with 8 lanes per pose (4 poses per warp, 32 poses per CTA over the 8 warps): lane l of a
// pose owns hidden units {l, 8+l} of the bone MLP's first layer and output feature l of its second layer
// (net_modules.py:86-111); vectors are exchanged with warp shuffles, features go through the [feature][pose]
// shared-memory buffer so that children find their parent's feature (net_modules.py:162-168).
struct EncLane {
    int m;        // pose column in the tile
    int l;        // lane within the 8-lane pose group
    int base;     // first lane of the group inside the warp
};

__device__ __forceinline__ float grp_get(float v, int src_l, const EncLane& e) { return __shfl_sync(0xffffffffu, v, e.base + src_l); }

// first part shared by forward and reverse: u, the two hidden units of this lane, all 10 hidden values, and this
// lane's output feature (o = min(l,5)).  d1a/d1b/d2 = activation derivatives at this lane's units.
// qs = column-normalised pose tile [pose][85]; apar = slope (relu 0 / lrelu 0.01) or softplus beta.
// ROOT is a compile-time flag so that the body is branch-free and two independent joints can be interleaved.
template <bool SOFT, bool ROOT>
__device__ __forceinline__ void bone_forward(const float* __restrict__ w, const float* __restrict__ qs, const float* feat, int i,
                                             int par, const EncLane& e, float apar, float (&u)[10], float (&h)[10], float& f,
                                             float& d1a, float& d1b, float& d2) {
    constexpr int fin = ROOT ? 4 : 10;
#pragma unroll
    for (int cpt = 0; cpt < 4; ++cpt) u[cpt] = qs[e.m * kXS + i * 4 + cpt];
#pragma unroll
    for (int r = 0; r < 6; ++r) u[4 + r] = ROOT ? 0.0f : feat[swz(par * 6 + r, e.m)];
    const int oa = e.l, ob = min(8 + e.l, 9);
    const float* w1 = w;
    const float* b1 = w + 10 * fin;
    const float* w2 = b1 + 10;
    const float* b2 = w2 + 60;
    float sa = b1[oa], sb = b1[ob];
#pragma unroll
    for (int k = 0; k < fin; ++k) {
        sa = fmaf(w1[oa * fin + k], u[k], sa);
        sb = fmaf(w1[ob * fin + k], u[k], sb);
    }
    const float ha = act_t<SOFT>(sa, apar, d1a);
    const float hb = act_t<SOFT>(sb, apar, d1b);
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = grp_get(ha, k, e);
    h[8] = grp_get(hb, 0, e);
    h[9] = grp_get(hb, 1, e);
    const int o = min(e.l, 5);
    float s2 = b2[o];
#pragma unroll
    for (int k = 0; k < 10; ++k) s2 = fmaf(w2[o * 10 + k], h[k], s2);
    f = act_t<SOFT>(s2, apar, d2);
}

// Joints are walked in a fixed schedule that pairs two joints of independent sub-trees per step (the legs / the arms),
// so that their dependent FMA / shuffle chains interleave: 12 steps instead of 21.  Every joint appears after its parent.
// (pairs: (1,0) (3,2) (5,4) (7,6) (9,8) (12,10) (14,11) (16,13) (17,15), then 18, 19, 20 alone.)
template <bool SOFT, bool RA, bool RB>
__device__ __forceinline__ void enc_fwd_pair(const float* encw, const float* qs, float* feat, float* stash, const EncLane& e,
                                             float apar, int ia, int ib) {
    float ua[10], ha[10], fa, xa, ya, za, ub[10], hb[10], fb, xb, yb, zb;
    bone_forward<SOFT, RA>(encw + enc_off(ia), qs, feat, ia, c_parent[ia], e, apar, ua, ha, fa, xa, ya, za);
    bone_forward<SOFT, RB>(encw + enc_off(ib), qs, feat, ib, c_parent[ib], e, apar, ub, hb, fb, xb, yb, zb);
    if (e.l < 6) {
        feat[swz(ia * 6 + e.l, e.m)] = fa;
        feat[swz(ib * 6 + e.l, e.m)] = fb;
        if (stash != nullptr) {
            stash[swz(ia * 6 + e.l, e.m)] = fa;     // same (swizzled) image as rows [0, 126) of the activation buffer:
            stash[swz(ib * 6 + e.l, e.m)] = fb;     // the reverse pass brings it back with ONE bulk copy
        }
    }
    __syncwarp();
}
template <bool SOFT>
__device__ __forceinline__ void enc_fwd_one(const float* encw, const float* qs, float* feat, float* stash, const EncLane& e,
                                            float apar, int i) {
    float u[10], h[10], f, x, y, z;
    bone_forward<SOFT, false>(encw + enc_off(i), qs, feat, i, c_parent[i], e, apar, u, h, f, x, y, z);
    if (e.l < 6) {
        feat[swz(i * 6 + e.l, e.m)] = f;
        if (stash != nullptr) stash[swz(i * 6 + e.l, e.m)] = f;
    }
    __syncwarp();
}

// forward encoder for this lane's pose; features are written as rows [i*6+o][m] of `feat` (and, if stash != nullptr,
// to the CTA's L2-resident stash so that the reverse pass does not have to recompute them).
template <bool SOFT>
__device__ __forceinline__ void encoder_forward(const float* encw, const float* qs, float* feat, float* stash, const EncLane& e,
                                                float apar) {
    enc_fwd_pair<SOFT, true, true>(encw, qs, feat, stash, e, apar, 1, 0);
    enc_fwd_pair<SOFT, false, true>(encw, qs, feat, stash, e, apar, 3, 2);
    // one copy of the pair body, walked 7 times (fully unrolled the encoder was 20 000 SASS instructions and stalled on
    // instruction fetch for a third of its cycles)
#pragma unroll 1
    for (int s = 0; s < 7; ++s) enc_fwd_pair<SOFT, false, false>(encw, qs, feat, stash, e, apar, c_pair_a[s], c_pair_b[s]);
#pragma unroll 1
    for (int i = 18; i <= 20; ++i) enc_fwd_one<SOFT>(encw, qs, feat, stash, e, apar, i);
}

// reverse step of one joint: everything up to (not including) the writes.  ua/ub = this lane's gradient w.r.t. its input
// elements j = l (and 8+l for l < 2, non-root joints).
template <bool SOFT, bool ROOT>
__device__ __forceinline__ void bone_backward(const float* encw, const float* qs, const float* feat, const float* gbuf, int i,
                                              const EncLane& e, float apar, float& ua, float& ub) {
    constexpr int fin = ROOT ? 4 : 10;
    const int par = c_parent[i];
    const float* w = encw + enc_off(i);
    float u[10], h[10], f, d1a, d1b, d2;
    bone_forward<SOFT, ROOT>(w, qs, feat, i, par, e, apar, u, h, f, d1a, d1b, d2);
    const float* w1 = w;
    const float* w2 = w + 10 * fin + 10;
    // t[o] = fbar[o] * act'(pre2[o]) on lane o (< 6)
    const float tl = (e.l < 6) ? gbuf[swz(i * 6 + e.l, e.m)] * d2 : 0.0f;
    float t[6];
#pragma unroll
    for (int o = 0; o < 6; ++o) t[o] = grp_get(tl, o, e);
    // s1[k] = (sum_o t[o] W2[o][k]) * act'(pre1[k]) for this lane's hidden units k = l, 8+l
    const int ka = e.l, kb = min(8 + e.l, 9);
    float ga = 0.0f, gb = 0.0f;
#pragma unroll
    for (int o = 0; o < 6; ++o) {
        ga = fmaf(t[o], w2[o * 10 + ka], ga);
        gb = fmaf(t[o], w2[o * 10 + kb], gb);
    }
    ga *= d1a;
    gb *= d1b;
    float s1[10];
#pragma unroll
    for (int k = 0; k < 8; ++k) s1[k] = grp_get(ga, k, e);
    s1[8] = grp_get(gb, 0, e);
    s1[9] = grp_get(gb, 1, e);
    const int ja = ROOT ? min(e.l, 3) : e.l, jb = min(8 + e.l, 9);
    ua = 0.0f; ub = 0.0f;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        ua = fmaf(s1[k], w1[k * fin + ja], ua);
        if (!ROOT) ub = fmaf(s1[k], w1[k * fin + jb], ub);
    }
}
template <bool ROOT>
__device__ __forceinline__ void bone_backward_write(float* gbuf, int i, const EncLane& e, float ua, float ub) {
    const int par = c_parent[i];
    if (e.l < 4) {
        gbuf[swz(128 + i * 4 + e.l, e.m)] = ua;
    } else if (!ROOT) {
        gbuf[swz(par * 6 + (e.l - 4), e.m)] += ua;
    }
    if (!ROOT && e.l < 2) gbuf[swz(par * 6 + 4 + e.l, e.m)] += ub;
}
template <bool SOFT, bool RA, bool RB>
__device__ __forceinline__ void enc_bwd_pair(const float* encw, const float* qs, const float* feat, float* gbuf, const EncLane& e,
                                             float apar, int ia, int ib) {
    float uaa, uab, uba, ubb;
    bone_backward<SOFT, RA>(encw, qs, feat, gbuf, ia, e, apar, uaa, uab);
    bone_backward<SOFT, RB>(encw, qs, feat, gbuf, ib, e, apar, uba, ubb);
    bone_backward_write<RA>(gbuf, ia, e, uaa, uab);     // the two joints of a pair never share a parent
    bone_backward_write<RB>(gbuf, ib, e, uba, ubb);
    __syncwarp();
}

// reverse encoder: features in `feat`, feature gradients in rows [0,126) of `gbuf` (accumulated in place; the schedule is
// the forward one reversed, a reverse topological order), quaternion gradients -> rows [128+e] of `gbuf`.
template <bool SOFT>
__device__ __forceinline__ void encoder_backward(const float* encw, const float* qs, const float* feat, float* gbuf,
                                                 const EncLane& e, float apar) {
#pragma unroll 1
    for (int i = 20; i >= 18; --i) {
        float ua, ub;
        bone_backward<SOFT, false>(encw, qs, feat, gbuf, i, e, apar, ua, ub);
        bone_backward_write<false>(gbuf, i, e, ua, ub);
        __syncwarp();
    }
#pragma unroll 1
    for (int s = 6; s >= 0; --s) enc_bwd_pair<SOFT, false, false>(encw, qs, feat, gbuf, e, apar, c_pair_a[s], c_pair_b[s]);
    enc_bwd_pair<SOFT, false, true>(encw, qs, feat, gbuf, e, apar, 3, 2);
    enc_bwd_pair<SOFT, true, true>(encw, qs, feat, gbuf, e, apar, 1, 0);
}

// pytorch3d 0.7.2 axis_angle_to_quaternion (formula restated; source not in the reference tree)
__device__ __forceinline__ void aa_to_quat(const float (&a)[3], float (&q)[4]) {
    const float ang = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    const float half = 0.5f * ang;
    const float k = (fabsf(ang) < 1e-6f) ? (0.5f - ang * ang / 48.0f) : (sinf(half) / ang);
    q[0] = cosf(half);
    q[1] = a[0] * k; q[2] = a[1] * k; q[3] = a[2] * k;
}
__device__ __forceinline__ void aa_to_quat_vjp(const float (&a)[3], const float (&qb)[4], float (&ab)[3]) {
    const float ang2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    const float ang = sqrtf(ang2);
    const float half = 0.5f * ang;
    const bool small = ang < 1e-6f;
    const float sa = small ? 1.0f : ang;
    const float s = sinf(half), co = cosf(half);
    const float k = small ? (0.5f - ang2 / 48.0f) : (s / sa);
    const float dk = small ? (-1.0f / 24.0f) : ((0.5f * co * sa - s) / (sa * sa * sa));
    const float dot = qb[1] * a[0] + qb[2] * a[1] + qb[3] * a[2];
    const float cw = qb[0] * (-0.5f * k) + dk * dot;
#pragma unroll
    for (int j = 0; j < 3; ++j) ab[j] = cw * a[j] + k * qb[1 + j];
}

// ------------------------------------------------------------------------------------------------ kernel
// MODE 0: forward only.  MODE 1: forward + reverse (+ step).  MODE 2: forward, then the forward-mode tangent of the
// DFNet along a given input tangent (same weights, activation replaced by a multiply with the stored derivative) --
// the second launch of a training step (Eikonal term), see posendf_b200/train.py.
// DSOFT / ESOFT: softplus DFNet / encoder (else piecewise-linear, slope from the config) -- compile-time, so every activation
// combination is its own kernel without the other variant's code in its epilogues.
// KS: small-tile variant (8 poses per tile, the lane groups split the reduction rows; see gemm_op) for batches that cannot
// fill the SMs with 32-pose tiles -- the reference's real call sites run B = 10 (experiments/sample_poses.py:96) and one motion
// sequence (experiments/motion_denoise.py:133-137).  Same buffers and layouts, only pose columns [0, 8) of a tile are live.
template <int MODE, bool DSOFT, bool ESOFT, bool KS = false>
__global__ void __launch_bounds__(kThreads, 1) pndf_fused_kernel(const KParams p) {
    static_assert(!(KS && MODE == 2), "the training launches always use 32-pose tiles");
    constexpr int kTilePoses = KS ? 8 : kTileM;
    constexpr bool kGrad = (MODE == 1);
    extern __shared__ __align__(1024) uint8_t smem[];
    float* X = reinterpret_cast<float*>(smem + kSmX);
    float* Y = reinterpret_cast<float*>(smem + kSmY);
    float* ring = reinterpret_cast<float*>(smem + kSmRing);
    uint8_t* mask = smem + kSmMask;
    // encoder weights (14 KB) are staged into the idle upper half of X around the encoder phases only
    float* encw = X + kEncStageRow * 32;
    float* xs = reinterpret_cast<float*>(smem + kSmXs);
    float* qs = reinterpret_cast<float*>(smem + kSmQs);
    float* nrm = reinterpret_cast<float*>(smem + kSmNrm);
    float* dval = reinterpret_cast<float*>(smem + kSmDv);   // [32] distance, [32] upstream*out_act'
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSmBar);   // [warp][stage]: "slab landed"

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < kWarps * kStages; ++i) mbar_init(&full[i], 1);
        mbar_init(&full[kWarps * kStages], 1);     // aux: pose tile / encoder weights / feature stash bulk copies
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // forward ops: F0(f0_slabs) F1(32) [F2a(64) F3a(64) F2b(64) F3b(64)] F4(32) F5(4)
    // reverse ops: B5(4) B4(32) [B3a(64) B2a(64) B3b(64) B2b(64)] B1(32) B0(8)
    // per-warp slab counts of the ops (4 KB each): forward F0 F1 [F2a F3a F2b F3b] F4 F5, reverse B5 B4 [B3a B2a B3b B2b] B1 B0
    constexpr int kS1 = slabs_of(256, 1, 64), kS23 = slabs_of(512, 1, 64), kS4 = slabs_of(512, 2, 64), kS5 = slabs_of(256, 1, 8);
    constexpr int kSB5 = slabs_of(64, 2, 64), kSB0 = slabs_of(256, 1, 16);
    const int fwd_slabs = p.f0_slabs + kS1 + 4 * kS23 + kS4 + kS5;
    constexpr int bwd_slabs = kSB5 + kS1 + 4 * kS23 + kS4 + kSB0;
    const int step_slabs = fwd_slabs + (kGrad ? bwd_slabs : 0);
    constexpr int kPasses = (MODE == 2) ? 2 : 1;   // MODE 2 replays the forward slab stream for the tangent pass
    // ... unless launch 1 handed over its activation-derivative masks (relu / lrelu DFNet): then only the tangent pass runs
    const bool tan_only = (MODE == 2) && (p.act_masks != nullptr);
    const int pass0 = tan_only ? 1 : 0;

    // ------------------------------------------------------------------ compute warps
    Ctx c;
    c.X = X; c.Y = Y; c.ring = ring + warp * (kStages * kSlabFloats); c.mask = mask;
    c.ring_s = smem_u32(c.ring); c.full_s = smem_u32(full + warp * kStages);
    c.tid = tid; c.lane = lane; c.mg = KS ? 0 : (lane >> 3); c.kq = lane >> 3; c.ng = warp * 8 + (lane & 7);
    c.ng2 = (warp & 3) * 8 + (lane & 7); c.kg = warp >> 2;
    c.slope = (p.df_act == ACT_RELU) ? 0.0f : 0.01f; c.df_beta = p.df_beta; c.df_inv_beta = 1.0f / p.df_beta;
    c.dscr = p.dscratch ? p.dscratch + (size_t)blockIdx.x * kUnits * 32 : nullptr;
    const uint32_t aux_s = smem_u32(full + kWarps * kStages);
    uint32_t aux_phase = 0;
    constexpr uint32_t kEncBytes = kEncFloats * 4;     // 14 064, a multiple of 16
    const bool keep = (MODE >= 1);
    const bool bias_lane = !KS || c.kq == 0;      // KS: the four lane groups' partial sums are added up -- the bias goes into one of them
    EncLane enc;
    enc.l = lane & 7; enc.base = lane & 24; enc.m = warp * 4 + (lane >> 3);
    Pipe pipe;
    pipe.stage = 0; pipe.phase = 0;
    pipe.step_slabs = (uint32_t)step_slabs;
    pipe.wsrc = reinterpret_cast<const char*>(p.wstream) + (size_t)warp * kSlabBytes;
    {
        const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        pipe.left = (uint32_t)my_tiles * (uint32_t)p.steps * (uint32_t)step_slabs * (uint32_t)(kPasses - pass0);
        pipe.pos = 0;
        // prologue: fill both stages of this warp's ring
        refill(pipe, c, 0);
        refill(pipe, c, 1);
    }

    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const long long pose0 = (long long)tile * kTilePoses;
        const int nvalid = (int)min((long long)kTilePoses, p.B - pose0);
        float* dbg = (p.dbg == nullptr) ? nullptr
                     : (p.dump_all ? p.dbg + (size_t)tile * kDumpRows * 32 : (tile == 0 ? p.dbg : nullptr));

        // ---- load the pose tile (coalesced), zero-fill the tail
        if (tan_only) {
            if (DSOFT) {
                c.dscr = reinterpret_cast<float*>(p.act_masks) + (size_t)tile * kUnits * 32;
            } else {
                const uint4* src = reinterpret_cast<const uint4*>(p.act_masks + (size_t)tile * (4 * kMaskStride));
                for (int i = tid; i < 4 * kMaskStride / 16; i += kGemmThreads) reinterpret_cast<uint4*>(mask)[i] = __ldg(src + i);
            }
        } else if (p.input_kind == IN_QUAT) {
            // ONE bulk copy: the tile is nvalid x 336 contiguous bytes of the batch; the encoder weights ride on the same barrier
            if (tid == 0) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous tile's write-back has read xs
                mbar_expect_tx_s(aux_s, (uint32_t)nvalid * 336u + (p.use_enc ? kEncBytes : 0u));
                bulk_g2s(xs, p.pose_in + pose0 * 84, (uint32_t)nvalid * 336u, aux_s);
                if (p.use_enc) bulk_g2s(encw, p.encw, kEncBytes, aux_s);
            }
            for (int idx = nvalid * 84 + tid; idx < kTileM * 84; idx += kGemmThreads) xs[idx] = 0.0f;      // ragged last tile
            mbar_wait_s(aux_s, aux_phase);
            aux_phase ^= 1u;
        } else {
            const float* src = p.pose_in + pose0 * 63;
            const bool dn = (p.dn.pending != 0);
            if (dn) {
                // per-sequence loss scale of the PREVIOUS step for every sequence that has a frame in this tile:
                // c = mean_t dist_prev[s][t] (same summation order as seq_adam_kernel), scale = weight * 2 c / T
                const int T = p.dn.T;
                const long long s_lo = pose0 / T, s_hi = (pose0 + nvalid - 1) / T;
                for (long long sq = s_lo; sq <= s_hi; ++sq) {
                    const float* dp = p.dn.dist_prev + sq * T;
                    float acc = 0.0f;
                    for (int t = tid; t < T; t += kGemmThreads) acc += dp[t];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (lane == 0) nrm[warp] = acc;
                    __syncthreads();
                    if (tid == 0) {
                        float tot = 0.0f;
                        for (int w = 0; w < kWarps; ++w) tot += nrm[w];
                        const float cmean = tot / (float)T;
                        dval[sq - s_lo] = p.dn.ap.weight * 2.0f * cmean / (float)T;
                        if (p.dn.loss_out != nullptr && sq * T >= pose0) p.dn.loss_out[sq] = p.dn.ap.weight * cmean * cmean;
                    }
                    __syncthreads();
                }
            }
            for (int idx = tid; idx < kTileM * 21; idx += kGemmThreads) {
                const int m = idx / 21, j = idx - m * 21;
                float a[3] = {0.f, 0.f, 0.f}, q[4];
                if (m < nvalid) {
                    if (dn) {
                        const long long e0 = (pose0 * 21 + idx) * 3;
                        const float scale = dval[(pose0 + m) / p.dn.T - pose0 / p.dn.T];
#pragma unroll
                        for (int k3 = 0; k3 < 3; ++k3) {
                            float av = p.dn.pose_rw[e0 + k3], mv = p.dn.m[e0 + k3], vv = p.dn.v[e0 + k3];
                            dn_adam_update(av, mv, vv, p.dn.graw[e0 + k3], scale, p.dn.ap);
                            p.dn.pose_rw[e0 + k3] = av; p.dn.m[e0 + k3] = mv; p.dn.v[e0 + k3] = vv;
                            a[k3] = av;
                        }
                    } else {
                        a[0] = __ldg(src + idx * 3); a[1] = __ldg(src + idx * 3 + 1); a[2] = __ldg(src + idx * 3 + 2);
                    }
                }
                aa_to_quat(a, q);
#pragma unroll
                for (int cpt = 0; cpt < 4; ++cpt) xs[m * kXS + j * 4 + cpt] = (m < nvalid) ? q[cpt] : 0.0f;
            }
        }
        if (p.use_enc && !tan_only && p.input_kind != IN_QUAT) {
            for (int i = tid; i < kEncFloats; i += kGemmThreads) encw[i] = __ldg(p.encw + i);
        }
        gemm_bar();

        for (int st = 0; st < p.steps; ++st) {
            float* dbg_s = (st == 0) ? dbg : nullptr;
            // ---- column norms, q = x / n, encoder: 8 lanes per pose, 4 poses per warp
            if (!tan_only) {
                const int cpt = enc.l & 3, hf = enc.l >> 2;
                if (p.normalise) {
                    float sq = 0.0f;
                    for (int j = hf; j < 21; j += 2) {
                        const float x = xs[enc.m * kXS + j * 4 + cpt];
                        sq = fmaf(x, x, sq);
                    }
                    sq += __shfl_xor_sync(0xffffffffu, sq, 4);
                    const float n = fmaxf(sqrtf(sq), 1e-12f);
                    if (hf == 0) nrm[cpt * 32 + enc.m] = n;
                    for (int j = hf; j < 21; j += 2) qs[enc.m * kXS + j * 4 + cpt] = xs[enc.m * kXS + j * 4 + cpt] / n;
                } else {
                    for (int j = hf; j < 21; j += 2) qs[enc.m * kXS + j * 4 + cpt] = xs[enc.m * kXS + j * 4 + cpt];
                }
                __syncwarp();
                if (p.use_enc) {
                    float* stash = (kGrad && p.z0scratch != nullptr) ? p.z0scratch + (size_t)blockIdx.x * 128 * 32 : nullptr;
                    encoder_forward<ESOFT>(encw, qs, X, stash, enc, ESOFT ? p.enc_beta : ((p.enc_act == ACT_RELU) ? 0.0f : 0.01f));
                    if (enc.l < 2) X[swz(126 + enc.l, enc.m)] = 0.0f;
                    // the feature stash (global, written above by this thread) comes back through a bulk copy = async proxy
                    if (kGrad) asm volatile("fence.proxy.async;" ::: "memory");
                } else {
                    for (int e = enc.l; e < 96; e += 8) X[swz(e, enc.m)] = (e < 84) ? qs[enc.m * kXS + e] : 0.0f;
                }
            }
            gemm_bar();
            for (int pass = pass0; pass < kPasses; ++pass) {
            const bool tangent = (MODE == 2) && (pass == 1);
            float* dbg_p = (MODE == 2) ? (tangent ? dbg_s : nullptr) : dbg_s;
            if (tangent) {   // the DFNet input tangent replaces z0; every op below becomes linear (no bias, act -> act')
                const float* src = p.tan_in + (size_t)tile * 128 * 32;
                for (int idx = tid; idx < p.z0_rows * 32; idx += kGemmThreads) X[swz(idx >> 5, idx & 31)] = __ldg(src + idx);
                gemm_bar();
            }
            dump_rows(dbg_p, 0, X, p.z0_rows, tid);

            // ================================================================= forward
            {   // F0: z0 (X) -> z1 (Y), 256 wide
                float acc[8][8];
                if (c.kg == 0 && !tangent && bias_lane) acc_init_bias<8, 2>(acc, p.bias[0], c.ng2, 256); else acc_zero<8>(acc);
                gemm_op<8, 2, KS>(acc, X, p.f0_slabs, pipe, c); kq_reduce<8, KS>(acc);
                splitk_combine<8, KS>(acc, Y, c);
                if (c.kg == 0) { if (!tangent) epilogue_fwd<DSOFT, 8, 2, KS>(acc, Y, kU1, c, keep); else epilogue_bwd<DSOFT, 8, 2, KS>(acc, Y, kU1, c); }
            }
            gemm_bar();
            dump_rows(dbg_p, 128, Y, 256, tid);
            {   // F1: z1 (Y) -> z2 (X), 512 wide
                float acc[8][8];
                if (!tangent && bias_lane) acc_init_bias<8>(acc, p.bias[1], c.ng, 512); else acc_zero<8>(acc);
                gemm_op<8, 1, KS>(acc, Y, kS1, pipe, c); kq_reduce<8, KS>(acc);
                if (!tangent) epilogue_fwd<DSOFT, 8, 1, KS>(acc, X, kU2, c, keep); else epilogue_bwd<DSOFT, 8, 1, KS>(acc, X, kU2, c);
            }
            gemm_bar();
            dump_rows(dbg_p, 384, X, 512, tid);
            {   // F2/F3 fused: z3 chunk (Y) is consumed at once as a K-chunk of layer 3; z4 -> X
                float acc3[8][8];
                if (!tangent && bias_lane) acc_init_bias<8>(acc3, p.bias[3], c.ng, 512); else acc_zero<8>(acc3);
                for (int ch = 0; ch < 2; ++ch) {
                    float acc2[8][8];
                    if (!tangent && bias_lane) acc_init_bias<8>(acc2, p.bias[2] + ch * 512, c.ng, 512); else acc_zero<8>(acc2);
                    gemm_op<8, 1, KS>(acc2, X, kS23, pipe, c); kq_reduce<8, KS>(acc2);
                    if (!tangent) epilogue_fwd<DSOFT, 8, 1, KS>(acc2, Y, kU3 + ch * 512, c, keep); else epilogue_bwd<DSOFT, 8, 1, KS>(acc2, Y, kU3 + ch * 512, c);
                    gemm_bar();
                    dump_rows(dbg_p, 896 + ch * 512, Y, 512, tid);
                    gemm_op<8, 1, KS>(acc3, Y, kS23, pipe, c);
                    gemm_bar();
                }
                kq_reduce<8, KS>(acc3);
                if (!tangent) epilogue_fwd<DSOFT, 8, 1, KS>(acc3, X, kU4, c, keep); else epilogue_bwd<DSOFT, 8, 1, KS>(acc3, X, kU4, c);
            }
            gemm_bar();
            dump_rows(dbg_p, 1920, X, 512, tid);
            {   // F4: z4 (X) -> z5 (Y), 256 wide
                float acc[8][8];
                if (c.kg == 0 && !tangent && bias_lane) acc_init_bias<8, 2>(acc, p.bias[4], c.ng2, 256); else acc_zero<8>(acc);
                gemm_op<8, 2, KS>(acc, X, kS4, pipe, c); kq_reduce<8, KS>(acc);
                splitk_combine<8, KS>(acc, Y, c);
                if (c.kg == 0) { if (!tangent) epilogue_fwd<DSOFT, 8, 2, KS>(acc, Y, kU5, c, keep); else epilogue_bwd<DSOFT, 8, 2, KS>(acc, Y, kU5, c); }
            }
            gemm_bar();
            dump_rows(dbg_p, 2432, Y, 256, tid);
            {   // F5: z5 (Y) -> z6 (X), 64 wide
                float acc[8][1];
                if (!tangent && bias_lane) acc_init_bias<1>(acc, p.bias[5], c.ng, 64); else acc_zero<1>(acc);
                gemm_op<1, 1, KS>(acc, Y, kS5, pipe, c); kq_reduce<1, KS>(acc);
                if (!tangent) epilogue_fwd<DSOFT, 1, 1, KS>(acc, X, kU6, c, keep); else epilogue_bwd<DSOFT, 1, 1, KS>(acc, X, kU6, c);
            }
            gemm_bar();
            dump_rows(dbg_p, 2688, X, 64, tid);
            // L6 (64 -> 1) + output activation: 8 lanes per pose, shuffle reduction
            if (!tangent) {
                const int m = enc.m;
                float s = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k) s = fmaf(__ldg(p.w6 + enc.l + 8 * k), X[swz(enc.l + 8 * k, m)], s);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __ldg(p.bias[6]);
                if (enc.l == 0) {
                    float dv;
                    const float d = act_eval(s, DSOFT ? ACT_SOFTPLUS : ACT_RELU, p.df_beta, dv);
                    dval[m] = d;
                    float gu = 1.0f;
                    if (p.g_up != nullptr && m < nvalid) gu = __ldg(p.g_up + pose0 + m);
                    dval[32 + m] = gu * dv;
                    if (m < nvalid && st == p.steps - 1) {
                        if (p.dist != nullptr) p.dist[pose0 + m] = d;
                        for (int r = 0; r < p.n_peers; ++r)
                            if (p.peer_dist[r] != nullptr) p.peer_dist[r][pose0 + m] = d;
                    }
                }
            }
            if (MODE == 2) gemm_bar();   // X (z6 / its tangent) is reloaded by the next pass
            }  // passes
            if (MODE == 1 && p.act_masks != nullptr && st == 0) {   // hand the activation derivatives to the tangent launch
                if (DSOFT) {   // fp32 derivatives: L2-resident per-CTA scratch -> per-tile buffer, streaming stores
                    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.act_masks) + (size_t)tile * kUnits * 32);
                    const float4* src = reinterpret_cast<const float4*>(c.dscr);
                    for (int i = tid; i < kUnits * 32 / 4; i += kGemmThreads) __stcs(dst + i, src[i]);
                } else {
                    uint4* dst = reinterpret_cast<uint4*>(p.act_masks + (size_t)tile * (4 * kMaskStride));
                    for (int i = tid; i < 4 * kMaskStride / 16; i += kGemmThreads) dst[i] = reinterpret_cast<const uint4*>(mask)[i];
                }
            }
            if (!kGrad) {
                gemm_bar();
                continue;
            }
            gemm_bar();

            // ================================================================= reverse
            // g6 = gs * W6 * act'(pre5)  -> Y rows [0,64)
            for (int idx = tid; idx < 64 * 32; idx += kGemmThreads) {
                const int k = idx >> 5, m = idx & 31;
                float dv;
                if (DSOFT) {
                    dv = c.dscr[(size_t)(kU6 + k) * 32 + m];
                } else {
                    const uint32_t b = mask[(m >> 3) * kMaskStride + kU6 + k];
                    dv = ((b >> (m & 7)) & 1u) ? 1.0f : c.slope;
                }
                Y[swz(k, m)] = dval[32 + m] * __ldg(p.w6 + k) * dv;
            }
            gemm_bar();
            dump_rows(dbg_s, 2752, Y, 64, tid);
            {   // B5: g6 (Y,64) -> g5 (X,256)
                float acc[8][8];
                acc_zero<8>(acc);
                gemm_op<8, 2, KS>(acc, Y, kSB5, pipe, c); kq_reduce<8, KS>(acc);
                splitk_combine<8, KS>(acc, X, c);
                if (c.kg == 0) epilogue_bwd<DSOFT, 8, 2, KS>(acc, X, kU5, c);
            }
            gemm_bar();
            dump_rows(dbg_s, 2816, X, 256, tid);
            {   // B4: g5 (X,256) -> g4 (Y,512)
                float acc[8][8];
                acc_zero<8>(acc);
                gemm_op<8, 1, KS>(acc, X, kS1, pipe, c); kq_reduce<8, KS>(acc);
                epilogue_bwd<DSOFT, 8, 1, KS>(acc, Y, kU4, c);
            }
            gemm_bar();
            dump_rows(dbg_s, 3072, Y, 512, tid);
            {   // B3/B2 fused: g3 chunk (X) is consumed at once as a K-chunk of B2; g2 -> Y
                float accb2[8][8];
                acc_zero<8>(accb2);
                for (int ch = 0; ch < 2; ++ch) {
                    float accb3[8][8];
                    acc_zero<8>(accb3);
                    gemm_op<8, 1, KS>(accb3, Y, kS23, pipe, c); kq_reduce<8, KS>(accb3);
                    epilogue_bwd<DSOFT, 8, 1, KS>(accb3, X, kU3 + ch * 512, c);
                    gemm_bar();
                    dump_rows(dbg_s, 3584 + ch * 512, X, 512, tid);
                    gemm_op<8, 1, KS>(accb2, X, kS23, pipe, c);
                    gemm_bar();
                }
                kq_reduce<8, KS>(accb2);
                epilogue_bwd<DSOFT, 8, 1, KS>(accb2, Y, kU2, c);
            }
            gemm_bar();
            // X is dead until B1 writes its rows [0, 256): bring the encoder weights back into its upper half now, under B1 / B0
            if (p.use_enc && tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx_s(aux_s, kEncBytes);
                bulk_g2s(encw, p.encw, kEncBytes, aux_s);
            }
            dump_rows(dbg_s, 4608, Y, 512, tid);
            {   // B1: g2 (Y,512) -> g1 (X,256)
                float acc[8][8];
                acc_zero<8>(acc);
                gemm_op<8, 2, KS>(acc, Y, kS4, pipe, c); kq_reduce<8, KS>(acc);
                splitk_combine<8, KS>(acc, X, c);
                if (c.kg == 0) epilogue_bwd<DSOFT, 8, 2, KS>(acc, X, kU1, c);
            }
            gemm_bar();
            dump_rows(dbg_s, 5120, X, 256, tid);
            {   // B0: g1 (X,256) -> g0 (Y,128)
                float acc[8][2];
                acc_zero<2>(acc);
                gemm_op<2, 1, KS>(acc, X, kSB0, pipe, c); kq_reduce<2, KS>(acc);
                epilogue_bwd<DSOFT, 2, 1, KS>(acc, Y, -1, c);
            }
            gemm_bar();
            dump_rows(dbg_s, 5376, Y, 128, tid);

            // ---- encoder reverse + normalise Jacobian + step: 8 lanes per pose
            if (p.use_enc) {   // X rows [0, 256) are free again: the encoder weights have landed, bring the features back (one bulk copy)
                mbar_wait_s(aux_s, aux_phase);
                aux_phase ^= 1u;
                if (tid == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx_s(aux_s, 128u * 32u * 4u);
                    bulk_g2s(X, p.z0scratch + (size_t)blockIdx.x * 128 * 32, 128u * 32u * 4u, aux_s);
                }
                mbar_wait_s(aux_s, aux_phase);
                aux_phase ^= 1u;
            }
            {
                const int m = enc.m;
                if (p.use_enc) {
                    encoder_backward<ESOFT>(encw, qs, X, Y, enc, ESOFT ? p.enc_beta : ((p.enc_act == ACT_RELU) ? 0.0f : 0.01f));   // qbar -> Y rows [128, 212)
                } else {
                    for (int e = enc.l; e < 84; e += 8) Y[swz(128 + e, m)] = Y[swz(e, m)];
                    __syncwarp();
                }
                const float d = dval[m];
                // xbar = (qbar - q <q,qbar>) / n   per component column  (Jacobian of F.normalize(dim=1))
                const int cpt = enc.l & 3, hf = enc.l >> 2;
                float n = 1.0f, dot = 0.0f;
                if (p.normalise) {
                    n = nrm[cpt * 32 + m];
                    for (int j = hf; j < 21; j += 2) dot = fmaf(qs[m * kXS + j * 4 + cpt], Y[swz(128 + j * 4 + cpt, m)], dot);
                    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
                    if (n <= 1e-12f) dot = 0.0f;   // clamp active: map is x/eps, Jacobian is I/eps
                }
                for (int j = hf; j < 21; j += 2) {
                    const int e = j * 4 + cpt;
                    const float x = xs[m * kXS + e];
                    float g = Y[swz(128 + e, m)];
                    if (p.normalise) g = (g - qs[m * kXS + e] * dot) / n;
                    Y[swz(128 + e, m)] = g;
                    if (p.do_step) xs[m * kXS + e] = __fsub_rn(x, __fmul_rn(d, g));   // two roundings, as torch's x - d*g
                }
                if (p.do_step && p.renorm) {
                    __syncwarp();
                    for (int j = enc.l; j < 21; j += 8) {
                        float sq = 0.0f;
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) sq = fmaf(xs[m * kXS + j * 4 + c4], xs[m * kXS + j * 4 + c4], sq);
                        const float inv = 1.0f / sqrtf(sq);
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) xs[m * kXS + j * 4 + c4] *= inv;
                    }
                }
            }
            // the updated tile leaves through the async proxy (bulk store below): make this thread's writes to xs visible to it
            if (p.pose_out != nullptr) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            gemm_bar();
        }  // steps

        // ---- write back (coalesced)
        if (kGrad) {
            if (p.grad != nullptr) {
                if (p.input_kind == IN_QUAT) {
                    float* dst = p.grad + pose0 * 84;
                    for (int idx = tid; idx < nvalid * 84; idx += kGemmThreads) {
                        const int m = idx / 84, e = idx - m * 84;
                        dst[idx] = Y[swz(128 + e, m)];
                    }
                } else {
                    const float* src = p.pose_in + pose0 * 63;
                    float* dst = p.grad + pose0 * 63;
                    for (int idx = tid; idx < nvalid * 21; idx += kGemmThreads) {
                        const int m = idx / 21, j = idx - m * 21;
                        // plain loads: in the fused denoise step this CTA has just rewritten these values (prologue)
                        const float a[3] = {src[idx * 3], src[idx * 3 + 1], src[idx * 3 + 2]};
                        const float qb[4] = {Y[swz(128 + j * 4, m)], Y[swz(128 + j * 4 + 1, m)], Y[swz(128 + j * 4 + 2, m)],
                                             Y[swz(128 + j * 4 + 3, m)]};
                        float ab[3];
                        aa_to_quat_vjp(a, qb, ab);
                        dst[idx * 3] = ab[0]; dst[idx * 3 + 1] = ab[1]; dst[idx * 3 + 2] = ab[2];
                    }
                }
            }
            if (p.pose_out != nullptr && tid == 0) {
                // the projected tile is a contiguous image of nvalid x 336 bytes: ONE bulk store, and one more per peer GPU
                // (fused gather: NVLink-mapped gathered buffers); they drain while the next tile computes
                bulk_s2g(p.pose_out + pose0 * 84, xs, (uint32_t)nvalid * 336u);
                for (int r = 0; r < p.n_peers; ++r) bulk_s2g(p.peer_pose[r] + pose0 * 84, xs, (uint32_t)nvalid * 336u);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        gemm_bar();   // xs / Y are reused by the next tile
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // write-backs complete before the CTA retires
}

}  // namespace pndf
