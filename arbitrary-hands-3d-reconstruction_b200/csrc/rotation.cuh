// Device-side rotation chain, float32, op-for-op the reference's order of evaluation so that the
// same branches are taken on the same inputs.
//   rodrigues()     : batch_rodrigues + quat2mat      /root/reference/mano/manolayer.py:423-434, 396-421
//   rot6d_to_aa()   : rot6d_to_rotmat -> rotation_matrix_to_quaternion -> quaternion_to_angle_axis
//                     -> NaN->0                         /root/reference/acr/utils.py:362-376, 826-906,
//                                                       773-823, 334-360
#pragma once
#include <cuda_runtime.h>

namespace acr {

// separately rounded products/sums (no FMA contraction), like the reference's element-wise ATen
// kernels: keeps the Gram-Schmidt residual of near-degenerate 6D inputs identical to the reference's
__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

__device__ __forceinline__ void rodrigues(float ax, float ay, float az, float* R) {
  // angle = || aa + 1e-8 ||, axis = aa / angle, q = [cos(a/2), sin(a/2) axis], q /= ||q||
  const float bx = ax + 1e-8f, by = ay + 1e-8f, bz = az + 1e-8f;
  const float ang = sqrtf(bx * bx + by * by + bz * bz);
  const float nx = ax / ang, ny = ay / ang, nz = az / ang;
  float s, c;
  sincosf(ang * 0.5f, &s, &c);
  float w = c, x = s * nx, y = s * ny, z = s * nz;
  const float qn = sqrtf(w * w + x * x + y * y + z * z);
  w /= qn; x /= qn; y /= qn; z /= qn;
  const float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  const float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;   R[2] = 2 * wy + 2 * xz;
  R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
  R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;   R[8] = w2 - x2 - y2 + z2;
}

// rotation_matrix_to_angle_axis (acr/utils.py:334-360) of a row-major 3x3 (not necessarily orthonormal)
// matrix: 4-case quaternion on the TRANSPOSED matrix (:862-906), atan2 form (:803-823), NaN -> 0.
__device__ __forceinline__ void rotmat_to_aa(const float* __restrict__ R, float* __restrict__ aa) {
  // t = R^T, i.e. t(i,j) = R(j,i)
  const float t00 = R[0], t01 = R[3], t02 = R[6], t10 = R[1], t11 = R[4], t12 = R[7], t20 = R[2], t21 = R[5], t22 = R[8];
  float qw, qx, qy, qz, tt;
  if (t22 < 1e-6f) {
    if (t00 > t11) {
      tt = 1 + t00 - t11 - t22;
      qw = t12 - t21; qx = tt; qy = t01 + t10; qz = t20 + t02;
    } else {
      tt = 1 - t00 + t11 - t22;
      qw = t20 - t02; qx = t01 + t10; qy = tt; qz = t12 + t21;
    }
  } else {
    if (t00 < -t11) {
      tt = 1 - t00 - t11 + t22;
      qw = t01 - t10; qx = t20 + t02; qy = t12 + t21; qz = tt;
    } else {
      tt = 1 + t00 + t11 + t22;
      qw = tt; qx = t12 - t21; qy = t20 - t02; qz = t01 - t10;
    }
  }
  const float sc = sqrtf(tt);
  qw = qw / sc * 0.5f; qx = qx / sc * 0.5f; qy = qy / sc * 0.5f; qz = qz / sc * 0.5f;
  const float s2 = qx * qx + qy * qy + qz * qz;
  const float sn = sqrtf(s2);
  const float two_theta = 2.0f * ((qw < 0.0f) ? atan2f(-sn, -qw) : atan2f(sn, qw));
  const float k = (s2 > 0.0f) ? two_theta / sn : 2.0f;
  const float ox = qx * k, oy = qy * k, oz = qz * k;
  aa[0] = isnan(ox) ? 0.f : ox;
  aa[1] = isnan(oy) ? 0.f : oy;
  aa[2] = isnan(oz) ? 0.f : oz;
}

__device__ __forceinline__ void rot6d_to_aa(const float* __restrict__ r6, float* __restrict__ aa) {
  // the 6 numbers are a row-major (3,2) matrix: a1 = column 0, a2 = column 1
  const float a1x = r6[0], a1y = r6[2], a1z = r6[4];
  const float a2x = r6[1], a2y = r6[3], a2z = r6[5];
  // b1 = a1 / max(||a1||, 1e-6)
  float n1 = fmaxf(sqrtf(dot3_rn(a1x, a1y, a1z, a1x, a1y, a1z)), 1e-6f);
  const float b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
  const float d = dot3_rn(b1x, b1y, b1z, a2x, a2y, a2z);
  const float ux = __fsub_rn(a2x, __fmul_rn(d, b1x)), uy = __fsub_rn(a2y, __fmul_rn(d, b1y)),
              uz = __fsub_rn(a2z, __fmul_rn(d, b1z));
  float n2 = fmaxf(sqrtf(dot3_rn(ux, uy, uz, ux, uy, uz)), 1e-6f);
  const float b2x = ux / n2, b2y = uy / n2, b2z = uz / n2;
  const float b3x = __fsub_rn(__fmul_rn(b1y, b2z), __fmul_rn(b1z, b2y)),
              b3y = __fsub_rn(__fmul_rn(b1z, b2x), __fmul_rn(b1x, b2z)),
              b3z = __fsub_rn(__fmul_rn(b1x, b2y), __fmul_rn(b1y, b2x));
  // R = [b1 b2 b3] (columns), row-major
  const float R[9] = {b1x, b2x, b3x, b1y, b2y, b3y, b1z, b2z, b3z};
  rotmat_to_aa(R, aa);
}

}  // namespace acr
