// pndf_feed.cuh -- training-data feed: one launch assembles a whole trainer batch from tables resident in HBM.
//
// Reference: PoseData.__getitem__ (model/load_data.py:43-71) + the DataLoader's stacking of batch_size items (:73-77).  Per
// item the reference np.load()s one file, draws num_pts random rows WITH replacement (np.random.randint, :49), optionally
// flips quaternions with a negative real part (quat_flip, :12-16), averages the 5 stored neighbour distances (:53), and pairs
// them with num_pts random rows of ONE randomly chosen AMASS file (:57-61); with flip=True it then overwrites the manifold
// poses with the flipped NOISY poses (:63, `quat_flip(poses)` -- a reference bug a drop-in keeps unless told otherwise).
//
// Here all files live concatenated in device memory (pose table (N,84), dist table (N,5), AMASS table (M,84)) with per-file
// offsets; thread group (item, point) draws its two row indices from a counter-based generator (splitmix64 of seed, item,
// point: reproducible, no state) or takes them from override arrays (tests: the indices the reference drew), and one warp
// copies a 336-byte row with three coalesced 16-byte stores per lane group.  HBM-bound: 336 + 20 + 336 B read,
// 336 + 4 + 336 B written per point.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pndf {

struct FeedParams {
    const float* pose_table;      // [N][84]
    const float* dist_table;      // [N][5]
    const long long* file_off;    // [n_files + 1] row offsets of the data files inside the tables
    const float* amass_table;     // [M][84]
    const long long* amass_off;   // [n_amass + 1]
    const int* item_file;         // [b] data file of every item of the batch
    const int* item_amass;        // [b] AMASS file of every item
    const long long* rows;        // [b][num_pts] override (row inside the file) or nullptr
    const long long* amass_rows;  // [b][num_pts] override or nullptr
    float* pose_out;              // [b][num_pts][84]
    float* dist_out;              // [b][num_pts]
    float* man_out;               // [b][num_pts][84]
    int b, num_pts, flip, fix_flip_bug;
    unsigned long long seed;
};

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// uniform integer in [0, n): multiply-shift of 53 random bits (bias < n / 2^53)
__device__ __forceinline__ long long rand_below(unsigned long long key, long long n) {
    return (long long)(((splitmix64(key) >> 11) * (unsigned long long)n) >> 53);
}

// one warp per (item, point): lanes 0..20 move one joint quaternion (16 bytes) each
__global__ void __launch_bounds__(256) feed_batch_kernel(const FeedParams p) {
    const long long total = (long long)p.b * p.num_pts;
    const int lane = threadIdx.x & 31;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += ((long long)gridDim.x * blockDim.x) >> 5) {
        const int item = (int)(w / p.num_pts);
        const int f = p.item_file[item], a = p.item_amass[item];
        const long long f0 = p.file_off[f], fn = p.file_off[f + 1] - f0;
        const long long a0 = p.amass_off[a], an = p.amass_off[a + 1] - a0;
        const long long r = p.rows ? p.rows[w] : rand_below(p.seed ^ (unsigned long long)(2 * w), fn);
        const long long ar = p.amass_rows ? p.amass_rows[w] : rand_below(p.seed ^ (unsigned long long)(2 * w + 1), an);
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f), m = q;
        if (lane < 21) {
            q = __ldg(reinterpret_cast<const float4*>(p.pose_table + (f0 + r) * 84) + lane);
            if (p.flip && q.x < 0.0f) q = make_float4(-q.x, -q.y, -q.z, -q.w);
            if (p.flip && !p.fix_flip_bug) {
                m = q;                                          // model/load_data.py:63: man_poses <- quat_flip(poses)
            } else {
                m = __ldg(reinterpret_cast<const float4*>(p.amass_table + (a0 + ar) * 84) + lane);
                if (p.flip && m.x < 0.0f) m = make_float4(-m.x, -m.y, -m.z, -m.w);
            }
            reinterpret_cast<float4*>(p.pose_out + w * 84)[lane] = q;
            reinterpret_cast<float4*>(p.man_out + w * 84)[lane] = m;
        } else if (lane == 21) {
            const float* d = p.dist_table + (f0 + r) * 5;
            // np.mean(axis=1) of 5 float32: sequential float32 adds, then the division
            p.dist_out[w] = ((((__ldg(d) + __ldg(d + 1)) + __ldg(d + 2)) + __ldg(d + 3)) + __ldg(d + 4)) / 5.0f;
        }
    }
}

}  // namespace pndf
