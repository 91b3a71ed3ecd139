// pndf_train_ops.cuh -- element-wise step of the softplus second-order adjoint chain of the Eikonal term
// (reference: loss.backward() through torch.autograd.grad(..., create_graph=True), model/posendf.py:89-96 and
// model/train_posendf.py:98; math in posendf_b200/train.py).  For hidden layer l with z' = softplus(pre_l):
//     phi'  = sigma(beta pre) = 1 - exp(-beta z')                       (from the exported activation z')
//     pdot  = zdot' / phi'                                              (tangent of pre_l from the exported tangent of z')
//     pbar  = zbar' * phi' + w * beta (1 - phi') * a_l * pdot           (a_l = exported first-order adjoint of pre_l)
// One pass over four (B x n) operands instead of eight torch element-wise kernels; z', zdot', a_l are column slices of
// the pose-major exports (row stride ld), zbar' and pbar are dense.
#pragma once
#include <cuda_runtime.h>

#include "pndf_kernel.cuh"

namespace pndf {

struct SoftplusAdjParams {
    const float* z_next;
    const float* zdot_next;
    const float* adj;
    const float* zbar;
    const float* w;        // device scalar (upstream weight of the Eikonal loss) or nullptr (= 1)
    float* pbar;
    long long ld, B;
    int n;                 // multiple of 4
    float beta;
};

__global__ void __launch_bounds__(256) softplus_adjoint_kernel(const SoftplusAdjParams p) {
    const int n4 = p.n >> 2;
    const long long total = p.B * n4;
    const float w = p.w ? __ldg(p.w) : 1.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / n4;
        const int j = (int)(i - b * n4) << 2;
        const float4 z = __ldcs(reinterpret_cast<const float4*>(p.z_next + b * p.ld + j));
        const float4 zd = __ldcs(reinterpret_cast<const float4*>(p.zdot_next + b * p.ld + j));
        const float4 a = __ldcs(reinterpret_cast<const float4*>(p.adj + b * p.ld + j));
        const float4 zb = __ldcs(reinterpret_cast<const float4*>(p.zbar + b * p.n + j));
        const float zz[4] = {z.x, z.y, z.z, z.w}, zzd[4] = {zd.x, zd.y, zd.z, zd.w}, aa[4] = {a.x, a.y, a.z, a.w},
                    zzb[4] = {zb.x, zb.y, zb.z, zb.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float d1 = -expm1f(-p.beta * zz[k]);
            const float pdot = zzd[k] / fmaxf(d1, 1e-30f);
            o[k] = zzb[k] * d1 + (w * p.beta * (1.0f - d1)) * aa[k] * pdot;
        }
        *reinterpret_cast<float4*>(p.pbar + b * p.n + j) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ---- rotation-format conversions on either side of the path (pytorch3d 0.7.2 transforms, formulas restated; the callers
// experiments/sample_poses.py:60,80 and experiments/motion_denoise.py:81 apply them around the projection / prior).
// One thread per joint rotation; n = number of rotations (B * 21).
__global__ void aa_to_quat_kernel(const float* __restrict__ aa, float* __restrict__ quat, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a[3] = {aa[i * 3], aa[i * 3 + 1], aa[i * 3 + 2]};
    float q[4];
    aa_to_quat(a, q);
    *reinterpret_cast<float4*>(quat + i * 4) = make_float4(q[0], q[1], q[2], q[3]);
}
// quaternion_to_axis_angle: norms = |q[1:]|, half = atan2(norms, q[0]), angle = 2 half,
// aa = q[1:] / (sin(half)/angle)   with the small-angle series 1/2 - angle^2/48 for |angle| < 1e-6
__global__ void quat_to_aa_kernel(const float* __restrict__ quat, float* __restrict__ aa, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 q = *reinterpret_cast<const float4*>(quat + i * 4);
    const float nrm = sqrtf(q.y * q.y + q.z * q.z + q.w * q.w);
    const float half = atan2f(nrm, q.x);
    const float ang = 2.0f * half;
    const float k = (fabsf(ang) < 1e-6f) ? (0.5f - ang * ang / 48.0f) : (sinf(half) / ang);
    aa[i * 3] = q.y / k; aa[i * 3 + 1] = q.z / k; aa[i * 3 + 2] = q.w / k;
}

}  // namespace pndf
