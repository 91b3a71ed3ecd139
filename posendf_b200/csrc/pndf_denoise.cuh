// pndf_denoise.cuh -- the Adam half of the motion-denoise inner loop (prior term only).
//
// Reference: experiments/motion_denoise.py:70 (Adam(lr .02) on the axis-angle body pose), :81-83 (prior term
// mean(dist) over the frames of ONE sequence), :29-35 (weight 1e7 * c^2 / (1 + it)), :97-99 (backward + step).
// The fused kernel has already produced, per frame, dist and d(dist)/d(axis-angle) with unit upstream gradient;
// this kernel finishes the step for one sequence per CTA:
//     c      = mean_t dist[s,t]                       (the reference's loss_dict['pose_pr'])
//     scale  = weight(it) * 2 c / T                   (d(weight * c^2)/d dist[s,t])
//     g      = scale * d(dist)/d(aa)                  then torch.optim.Adam's update, same op order
// It is elementwise / HBM-bound: 5 reads + 3 writes of 4 B per parameter.
#pragma once
#include <cuda_runtime.h>

#include "pndf_kernel.cuh"      // AdamParams, dn_adam_update (shared with the update fused into the prior launch)

namespace pndf {

__global__ void __launch_bounds__(256) seq_adam_kernel(float* __restrict__ aa, const float* __restrict__ graw,
                                                       const float* __restrict__ dist, float* __restrict__ m,
                                                       float* __restrict__ v, float* __restrict__ loss_out, int T,
                                                       AdamParams ap) {
    __shared__ float red[8];
    __shared__ float s_scale;
    const int s = blockIdx.x;
    const float* d = dist + (size_t)s * T;
    float acc = 0.0f;
    for (int t = threadIdx.x; t < T; t += blockDim.x) acc += d[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.0f;
        for (int w = 0; w < 8; ++w) tot += red[w];
        const float c = tot / (float)T;
        s_scale = ap.weight * 2.0f * c / (float)T;
        if (loss_out) loss_out[s] = ap.weight * c * c;
    }
    __syncthreads();
    const float scale = s_scale;
    const size_t base = (size_t)s * T * 63;
    for (int i = threadIdx.x; i < T * 63; i += blockDim.x) {
        float av = aa[base + i], mi = m[base + i], vi = v[base + i];
        dn_adam_update(av, mi, vi, graw[base + i], scale, ap);
        aa[base + i] = av; m[base + i] = mi; v[base + i] = vi;
    }
}

}  // namespace pndf
