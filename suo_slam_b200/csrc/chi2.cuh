// The chi2 test shared by the hypothesis-scoring kernels (frames.cu: suo_chi2_inlier_counts; slam.cu: camera-pose vote, re-initialisation
// test): does pose P explain keypoint x of a detection?  Reference lib/object_slam.py:1040-1066 / :655-680:
//   p_C = R p_O + t (utils.transform_pts), uvw = K p_C, keep w > 0, res = uv - uvw.xy / w,
//   cov diagonal floored at 1e-4 (:1054, :669) then inverted — or 1 / manual_kp_std^2 when there is no network covariance
//   (:1059-1061) — chi2 = res^T inf res <= gate.
// The reference inverts the float32 covariance in float32 (np.linalg.inv); here the float32 inputs are widened and the 2x2 inverse is
// closed-form FP64: identical decisions unless a chi2 sits within ~1e-6 relative of the gate.
#pragma once

__device__ __forceinline__ bool chi2_inlier(const double* P /* [R|t] 3x4 row-major */, const double* Kd /* 3x3 */, const double* xk, float u, float v,
                                            const float* cov /* 2x2 or nullptr */, double inv_manual_var, double gate) {
  const double x = xk[0], y = xk[1], z = xk[2];
  const double pc0 = P[0] * x + P[1] * y + P[2] * z + P[3];
  const double pc1 = P[4] * x + P[5] * y + P[6] * z + P[7];
  const double pc2 = P[8] * x + P[9] * y + P[10] * z + P[11];
  const double w = Kd[6] * pc0 + Kd[7] * pc1 + Kd[8] * pc2;
  if (!(w > 0)) return false;
  const double r0 = (double)u - (Kd[0] * pc0 + Kd[1] * pc1 + Kd[2] * pc2) / w;
  const double r1 = (double)v - (Kd[3] * pc0 + Kd[4] * pc1 + Kd[5] * pc2) / w;
  double chi2;
  if (cov) {
    const double a = fmax((double)cov[0], 1e-4), b = (double)cov[1], c = (double)cov[2], e = fmax((double)cov[3], 1e-4);
    const double det = a * e - b * c;
    chi2 = (r0 * (e * r0 - b * r1) + r1 * (-c * r0 + a * r1)) / det;
  } else {
    chi2 = (r0 * r0 + r1 * r1) * inv_manual_var;
  }
  return chi2 <= gate;
}
