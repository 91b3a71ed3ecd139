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

namespace pndf {

struct AdamParams {
    float lr, beta1, beta2, eps;
    float bias1, bias2;      // 1 - beta^t for this step
    float weight;            // 1e7 / (1 + it)
};

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
    const float step = ap.lr / ap.bias1;
    const float inv_sqrt_b2 = 1.0f / sqrtf(ap.bias2);
    for (int i = threadIdx.x; i < T * 63; i += blockDim.x) {
        const float g = scale * graw[base + i];
        const float mi = ap.beta1 * m[base + i] + (1.0f - ap.beta1) * g;
        const float vi = ap.beta2 * v[base + i] + (1.0f - ap.beta2) * g * g;
        m[base + i] = mi;
        v[base + i] = vi;
        const float denom = sqrtf(vi) * inv_sqrt_b2 + ap.eps;
        aa[base + i] -= step * (mi / denom);
    }
}

}  // namespace pndf
