// pndf_tc_gemm.cuh -- 3xTF32 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), the building block of the tensor-core
// DFNet path.
//
//     D[M x N] = epilogue( A[M x K] * B[N x K]^T )        A, B given as tf32 hi / lo pairs (x = hi + lo, both exactly
//                                                          representable in tf32), row-major, K contiguous (K-major)
//
// Arithmetic (measured to hold the 1e-5 parity bar of the whole DFNet chain, tools/tc_chain_probe.cu):
//     A*B ~= A_hi B_hi  +  (A_lo B_hi + A_hi B_lo)     three tcgen05.mma.kind::tf32 per K step, the big term and the cross
// terms in SEPARATE TMEM accumulators, a fresh accumulator pair per 128-deep K chunk; pairs are added in fp32 (round to nearest)
// on the CUDA cores -- the tensor core's accumulator truncates, so the long sum must not live in it.
//
// Persistent CTAs walk 128 x NT output tiles (NT = 128 or 64); 320 threads, warp-specialised:
//   warp 0      TMA producer: per 32-deep K stage four tile loads (A_hi, A_lo 128 x 32, B_hi, B_lo NT x 32, SWIZZLE_128B) on a
//               `full` mbarrier; 3 stages in flight
//   warp 1      MMA issuer (one lane): 12 MMAs per stage; tcgen05.commit releases the stage (`empty`) and, at the end of a K chunk,
//               publishes the accumulator pair (`acc_full`)
//   warps 2-9   drain, two warps per TMEM lane quarter: tcgen05.ld both accumulators of the finished pair, add into NT / 2 running
//               fp32 sums per thread (thread = one output row x half of the columns), hand the pair back (`acc_empty`) so that the next-but-one chunk can overwrite it --
//               the drain of chunk c runs under the MMAs of chunk c+1 (two pairs = 4 x NT TMEM columns); finally the epilogue
//               functor turns the 128-float row segment into whatever the layer needs (stored coalesced through shared memory).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pndf_tc {

constexpr int kTM = 128;                 // rows (poses) per tile = TMEM lanes
constexpr int kKB = 32;                  // K per stage = one 128-byte swizzle row of tf32
constexpr int kStages = 3;
#ifndef PNDF_TC_CHUNK_K
#define PNDF_TC_CHUNK_K 128
#endif
constexpr int kChunkK = PNDF_TC_CHUNK_K;  // K depth per accumulator pair
constexpr int kThreads = 320;            // warp 0 TMA, warp 1 MMA, warps 2-9 drain / epilogue
constexpr int kDrainWarps = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (++spins > (1u << 26)) __trap();      // a lost arrival is a launch error, never a hung GPU
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// TMA tensor store of one box from (swizzled) shared memory; bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, const void* src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row x 128-byte atoms, 1024 bytes apart)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// ---- CTA-pair (cta_group::2) variants: the two CTAs of a cluster run ONE M = 256 MMA stream issued by the leader (cluster rank 0);
// each CTA feeds its own 128 rows of A and its half of the B operand from its own shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {      // shared::cta address -> shared::cluster address in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on a barrier of the pair's leader
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(bar_cluster_addr)
                 : "memory");
}
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair once all prior MMAs of this thread have retired
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// 256-bit global store / load (sm_100: STG.256 / LDG.256), 32-byte aligned
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const float* p, float* v) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

// TILED matrix layout of every GEMM operand / result: a [R][C] matrix (C a multiple of 32) is stored as
// [R / T][C / 32][T][32] floats -- T-row tile, 32-column (128-byte) K block -- so that a TMA box of T rows x 32 columns is one
// contiguous T * 128-byte block and a warp's 32 rows x 32 columns of an output tile are 4 KB contiguous.  T = 128 for
// activations, the N tile (128 or 64) for weights.
__host__ __device__ __forceinline__ size_t tiled_offset(long long r, int c, int C, int T = kTM) {
    return ((size_t)((r / T) * (C / 32) + c / 32) * T + (size_t)(r % T)) * 32 + (size_t)(c % 32);
}

template <int NT>
__host__ __device__ constexpr int stage_bytes() { return 2 * kTM * 128 + 2 * NT * 128; }
constexpr int kOutStage = 32 * 128;      // per drain warp: one 32-row x 32-column output block, 128-byte swizzled, the source of a TMA store
template <int NT>
__host__ __device__ constexpr int smem_bytes() { return kStages * stage_bytes<NT>() + kDrainWarps * kOutStage + 1024 + 256; }
// CTA pair: per stage and CTA A_hi, A_lo (its 128 rows), Bw = this CTA's half of [B_hi; B_lo] (rank 0: B_hi, rank 1: B_lo) and
// Bx = its half of B_hi for the cross-term MMA (rank 0: rows [0, 64), rank 1: rows [64, 128))
__host__ __device__ constexpr int pair_stage_bytes() { return 2 * kTM * 128 + 128 * 128 + 64 * 128; }
__host__ __device__ constexpr int pair_smem_bytes() { return kStages * pair_stage_bytes() + kDrainWarps * kOutStage + 1024 + 256; }

struct GemmMaps {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    CUtensorMap out[2];      // the output arrays (tiled layout), box = 32 rows x 32 columns (one drain warp's block)
    CUtensorMap b_x;         // CTA-pair kernels only: B_hi with a 64-row box
};

// Epilogue functor interface:
//   static constexpr int kOutputs            1 or 2 output arrays
//   void init4(int col, float* r) const      the values the running sums of columns col .. col + 3 start from (a bias, or 0)
//   void operator()(int row, int col0, const float (&v)[32], float (&o0)[32], float (&o1)[32]) const
//                                            called by the thread that owns `row` for every 32-column group of its tile
//   float* out(int which) const, int ld()    the output arrays (row-major, row length ld)
// The kernel stores the outputs itself: a drain warp writes its 32 rows x 32 columns into a 128-byte-swizzled 4 KB staging block
// (conflict-free STS.128) and ONE TMA tensor store moves the block -- 4 contiguous KB of the tiled layout -- to global memory.
// (Per-thread STG.256 of a row's 128 bytes touches 32 lines per instruction: the LSU queue (`stall_lg`) held 38 % of the drain
// warps' samples and the epilogue outlasted the two accumulator pairs of runway the MMA warp has.)
//
// Persistent: gridDim.x CTAs walk the (m_tile, n_tile) list (n fastest: the CTAs that share an A tile run together).  The running
// sums live in the drain warps' registers, so the TMEM accumulator pairs are free as soon as a tile's last chunk is drained: the
// MMA warp starts the next tile while the drain warps are still busy with the previous tile's epilogue (activation, hi / lo split,
// stores).
// PAIR = true (NT = 128 only, launched as clusters of 2 CTAs): the pair owns 256 x 128 output tiles; the leader issues
// cta_group::2 MMAs -- A_hi x [B_hi; B_lo] as M = 256, N = 256 with B_hi in the leader's and B_lo in the peer's shared memory, and
// A_lo x B_hi as M = 256, N = 128 with each CTA holding 64 rows of B_hi -- so every CTA's shared memory feeds 14 KB per 8-deep K
// slice instead of 20 KB; accumulators, drain and epilogue stay per CTA (its 128 rows).  MEASURED (tools/tc_gemm_test.cu,
// profiles/tc_experiments_r02.txt): bit-identical results, 0.299 vs 0.292 ms on the 65 536 x 1 024 x 512 GEMM -- no gain, so the
// product launches PAIR = false; the variant stays as the tested starting point for 256 x 256 pair tiles.
template <int NT, class Epilogue, bool PAIR = false>
__global__ void __launch_bounds__(kThreads, 1) tc_gemm_kernel(const __grid_constant__ GemmMaps maps, int K, int m_tiles, int n_tiles,
                                                              Epilogue epi) {
    static_assert(!PAIR || NT == 128, "CTA pairs: 128-column tiles only");
    constexpr int kStageBytes = PAIR ? pair_stage_bytes() : stage_bytes<NT>();
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    // work list: single CTAs walk (m tile, n tile); pairs walk (pair of m tiles, n tile) and each CTA takes the m tile of its rank
    const int walker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, walkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto m_tile_of = [&](int tile) { return PAIR ? (tile / n_tiles) * 2 + (int)rank : tile / n_tiles; };
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* out_stage = smem + kStages * kStageBytes;             // [kDrainWarps][kOutStage], 1024-byte aligned blocks
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + kDrainWarps * kOutStage);
    uint64_t* full = bars;                   // [kStages]
    uint64_t* empty = bars + kStages;        // [kStages]
    uint64_t* acc_full = bars + 2 * kStages; // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nks = K / kKB;
    constexpr int kStagesPerChunk = kChunkK / kKB;
    const int nchunks = (nks + kStagesPerChunk - 1) / kStagesPerChunk;
    const int ntiles = (PAIR ? m_tiles / 2 : m_tiles) * n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        // (pair: the leader's acc_empty collects the drain warps of both CTAs)
        for (int p = 0; p < 2; ++p) { mbar_init(&acc_full[p], 1); mbar_init(&acc_empty[p], PAIR ? 2 * kDrainWarps : kDrainWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (PAIR) {      // warp 0 of both CTAs, collectively
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(4 * NT));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(4 * NT < 32 ? 32 : 4 * NT));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (PAIR) cluster_sync_all(); else __syncthreads();      // pair: the peer's barriers exist before anything is signalled on them
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t it = 0;      // running stage counter over all tiles of this CTA
            for (int tile = walker; tile < ntiles; tile += walkers) {
                const int mt = m_tile_of(tile), nt = tile % n_tiles;
                for (int ks = 0; ks < nks; ++ks, ++it) {
                    const uint32_t s = it % kStages;
                    mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);      // fresh barrier: parity 1 passes
                    uint8_t* st = smem + s * kStageBytes;
                    // operands live in TILED layout: [row tile][K block][rows of the tile][32 floats] -- every TMA box (rows x 128 bytes) is
                    // one contiguous 16 KB (8 KB) block of memory instead of 128 row segments 2-4 KB apart
                    const int ra = (mt * nks + ks) * kTM, rb = (nt * nks + ks) * NT;
                    if (PAIR) {
                        // every byte of the pair's stage is counted on the LEADER's full barrier (the MMA issuer waits there)
                        const uint32_t bar = mapa_u32(smem_u32(&full[s]), 0);
                        if (rank == 0) mbar_expect_tx(&full[s], 2 * pair_stage_bytes());
                        const uint32_t d = smem_u32(st);
                        tma_load_2d_pair(d, &maps.a_hi, 0, ra, bar);
                        tma_load_2d_pair(d + kTM * 128, &maps.a_lo, 0, ra, bar);
                        tma_load_2d_pair(d + 2 * kTM * 128, rank == 0 ? &maps.b_hi : &maps.b_lo, 0, rb, bar);
                        tma_load_2d_pair(d + 2 * kTM * 128 + 128 * 128, &maps.b_x, 0, rb + 64 * (int)rank, bar);
                        continue;
                    }
#ifdef PNDF_TC_EXP_HALF_FEED      // bottleneck experiment (tools/tc_gemm.sh): only the hi operands travel, the MMAs run on stale lo tiles
                    mbar_expect_tx(&full[s], stage_bytes<NT>() / 2);
                    tma_load_2d(st, &maps.a_hi, 0, ra, &full[s]);
                    tma_load_2d(st + 2 * kTM * 128, &maps.b_hi, 0, rb, &full[s]);
                    continue;
#endif
                    mbar_expect_tx(&full[s], stage_bytes<NT>());
                    tma_load_2d(st, &maps.a_hi, 0, ra, &full[s]);
                    tma_load_2d(st + kTM * 128, &maps.a_lo, 0, ra, &full[s]);
                    tma_load_2d(st + 2 * kTM * 128, &maps.b_hi, 0, rb, &full[s]);
                    tma_load_2d(st + 2 * kTM * 128 + NT * 128, &maps.b_lo, 0, rb, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0 && rank == 0) {
            // D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24 (pair: M = 256 over the two CTAs)
            constexpr uint32_t kM = PAIR ? 2 * kTM : kTM;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * NT >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
            uint32_t it = 0, cc = 0;      // running stage / chunk counters
            for (int tile = walker; tile < ntiles; tile += walkers) {
                for (int ks = 0; ks < nks; ++ks, ++it) {
                    const uint32_t s = it % kStages, p = cc & 1;
                    const bool chunk_start = (ks % kStagesPerChunk) == 0;
                    if (chunk_start) mbar_wait(&acc_empty[p], ((cc >> 1) & 1) ^ 1);      // the drain of chunk cc-2 has emptied pair p
                    mbar_wait(&full[s], (it / kStages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t st = smem_u32(smem + s * kStageBytes);
                    const uint32_t a_hi = st, a_lo = st + kTM * 128, b_hi = st + 2 * kTM * 128, b_x = b_hi + 128 * 128;
                    const uint32_t acc_hh = tmem + (uint32_t)(p * 2 * NT), acc_x = acc_hh + NT;
#pragma unroll
                    for (int k = 0; k < kKB / 8; ++k) {
                        const uint32_t first = (chunk_start && k == 0) ? 0u : 1u;
                        // B_hi and B_lo are adjacent in the stage and acc_x follows acc_hh in TMEM: ONE N = 2 NT instruction computes
                        // A_hi B_hi -> acc_hh and A_hi B_lo -> acc_x (two instructions per 8-deep K slice instead of three, A_hi
                        // fetched once: -10 % GEMM time).  The CTA-pair variant cuts the per-CTA operand fetch further (20 -> 14 KB
                        // per slice) and measured the same time as this one: operand fetch is not what paces the stream.
                        if (PAIR) {
                            mma_tf32_pair(acc_hh, make_desc(a_hi + k * 32), make_desc(b_hi + k * 32), idesc2, first);
                            mma_tf32_pair(acc_x, make_desc(a_lo + k * 32), make_desc(b_x + k * 32), idesc, 1u);
                        } else {
                            mma_tf32(acc_hh, make_desc(a_hi + k * 32), make_desc(b_hi + k * 32), idesc2, first);
                            mma_tf32(acc_x, make_desc(a_lo + k * 32), make_desc(b_hi + k * 32), idesc, 1u);
                        }
                    }
                    if (PAIR) mma_commit_pair(&empty[s]); else mma_commit(&empty[s]);      // stage s may be refilled once these MMAs retire
                    if ((ks % kStagesPerChunk) == kStagesPerChunk - 1 || ks == nks - 1) {
                        if (PAIR) mma_commit_pair(&acc_full[p]); else mma_commit(&acc_full[p]);
                        ++cc;
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ drain + epilogue: thread = output row x half of the columns
        // (two warps per TMEM lane quarter, each owning NT / 2 columns: with four fat drain warps the accumulator pairs -- and with
        // them the MMA warp -- waited for the drain: 3.46 -> 2.86 ms per 65 536-pose step; sixteen thin ones with direct stores
        // were slower again, 3.15 ms)
        const int quarter = warp & 3;                       // TMEM lanes [32 q, 32 q + 32) are this warp's
        const int half = (warp - 2) >> 2;                   // columns [half * NT / 2, (half + 1) * NT / 2)
        constexpr int NH = NT / 2;
        const uint32_t stg = smem_u32(out_stage + (warp - 2) * kOutStage);
        uint32_t cc = 0;
        const uint32_t acc_empty_leader[2] = {PAIR ? mapa_u32(smem_u32(&acc_empty[0]), 0) : 0u, PAIR ? mapa_u32(smem_u32(&acc_empty[1]), 0) : 0u};
        for (int tile = walker; tile < ntiles; tile += walkers) {
            const int m0 = m_tile_of(tile) * kTM, n0 = (tile % n_tiles) * NT;
            const int row = m0 + quarter * 32 + lane;
            float run[NH];
#pragma unroll
            for (int j = 0; j < NH; j += 4) epi.init4(n0 + half * NH + j, &run[j]);
            for (int c = 0; c < nchunks; ++c, ++cc) {
                const uint32_t p = cc & 1;
                mbar_wait(&acc_full[p], (cc >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t base = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(p * 2 * NT + half * NH);
#ifndef PNDF_TC_EXP_NO_DRAIN     // bottleneck experiment: accumulators are handed straight back
#pragma unroll
                for (int g = 0; g < NH / 32; ++g) {
                    uint32_t hh[32], xx[32];
                    tmem_ld32(base + g * 32, hh);
                    tmem_ld32(base + NT + g * 32, xx);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) run[g * 32 + j] += __uint_as_float(hh[j]) + __uint_as_float(xx[j]);
                }
#endif
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_cluster(acc_empty_leader[p]); else mbar_arrive(&acc_empty[p]);
                }
            }
            // epilogue of this tile (the MMA warp is already on the next one)
#ifdef PNDF_TC_EXP_NO_EPI       // bottleneck experiment: nothing is computed or stored after the drain
            if (run[0] != 123.456f) continue;
#endif
#pragma unroll
            for (int g = 0; g < NH / 32; ++g) {
                float v[32], o0[32], o1[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = run[g * 32 + j];
                const int col0 = n0 + half * NH + g * 32;
                epi(row, col0, v, o0, o1);
#pragma unroll
                for (int w = 0; w < Epilogue::kOutputs; ++w) {
                    // tiled output (the next GEMM's A operand): this warp's 32 rows x 32 columns are ONE contiguous 4 KB block = one TMA
                    // box.  Row `lane` of the staging block, 16-byte chunk c at position c ^ (lane & 7): the 128-byte swizzle the tensor
                    // map undoes on the way out, and conflict-free for the warp's STS.128.
#ifndef PNDF_TC_STAGED_STORE      // (staging read back with LDS.128 + 512-byte-per-instruction STG.128 measured the same: 2.17 vs 2.16 ms)
                    tma_store_wait_read();              // the previous store of this warp has finished reading the block
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float* o = (w == 0) ? &o0[4 * c] : &o1[4 * c];
                        st_shared_v4(stg + lane * 128 + ((c ^ (lane & 7)) << 4), o[0], o[1], o[2], o[3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const int c1 = ((m0 / kTM) * (epi.ld() / 32) + col0 / 32) * kTM + quarter * 32;
                        tma_store_2d(&maps.out[w], 0, c1, reinterpret_cast<const void*>(out_stage + (warp - 2) * kOutStage));
                        tma_store_commit();
                    }
#else
                    __syncwarp();                       // the previous block has been read out
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float* o = (w == 0) ? &o0[4 * c] : &o1[4 * c];
                        st_shared_v4(stg + lane * 128 + ((c ^ (lane & 7)) << 4), o[0], o[1], o[2], o[3]);
                    }
                    __syncwarp();
                    float* out = epi.out(w) + tiled_offset(m0 + quarter * 32, col0, epi.ld());
#pragma unroll
                    for (int i = 0; i < 8; ++i) {      // 4 rows x 128 bytes = 512 contiguous bytes per instruction
                        const int r = i * 4 + (lane >> 3), c = lane & 7;
                        float4 t;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                     : "r"(stg + r * 128 + ((c ^ (r & 7)) << 4)));
                        *reinterpret_cast<float4*>(out + r * 32 + c * 4) = t;
                    }
#endif
                }
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // this warp's last stores have landed before the CTA retires
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    if (PAIR) cluster_sync_all(); else __syncthreads();      // pair: no CTA retires while the other may still signal its barriers
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(4 * NT));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(4 * NT < 32 ? 32 : 4 * NT));
    }
}

// ------------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiledFn)p;
    }
    return fn;
}
// tiled fp32 matrix [rows][K] (see tiled_offset, tile = box_rows): as a 2-D array of rows * K / 32 rows of 32 floats; a box =
// box_rows consecutive rows = one contiguous block; 128B swizzle on the way into shared memory
inline bool make_map(CUtensorMap* tm, const float* base, uint64_t rows, uint64_t K, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {32, rows * (K / 32)};
    const cuuint64_t strides[1] = {32 * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kKB, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// launch of a CTA-pair kernel: clusters of 2 CTAs (two SMs of one TPC), an even grid
template <class Kern, class... Args>
inline cudaError_t launch_pair(Kern kern, int grid, int smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(grid & ~1));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
// output map of a GEMM: the tiled [rows][ld] array, one drain warp's 32 x 32 block per store
inline bool make_out_map(CUtensorMap* tm, const float* base, uint64_t rows, uint64_t ld) { return make_map(tm, base, rows, ld, 32); }

}  // namespace pndf_tc
