// Batched PnP on the GPU: Lambda-Twist P3P -> P4P disambiguation ->
// RANSAC -> Levenberg-Marquardt refine, FP64, one CTA per object.
//
// Replaces lambdatwist.pnp(xs, ys, threshold) — thirdparty/lambdatwist/pnp_python_binding.cpp:32-62
// -> PNP::compute / PNP::refine (pnp_ransac.cpp:188-326), p4p (p4p.cpp:11-60),
// p3p_lambdatwist (lambdatwist/lambdatwist.p3p.h:33-339) and the external Ceres solve the
// reference calls at pnp_ransac.cpp:263-278,304-322 — which the reference runs serially,
// one object at a time, through pybind.
//
// Parallel structure: RANSAC hypotheses are independent, so each thread draws a 4-subset
// from a counter-based generator keyed by (seed, object key, iteration), solves P4P and
// counts inliers; thread 0 then replays the reference's *sequential* accept rule over the
// batch (first hypothesis with a strictly larger inlier count wins, and the adaptive
// iteration budget of parameters.h:76-102 shrinks as the best count grows) so the winner
// and the number of iterations are exactly those of the serial loop.  The refine is the
// trust-region LM Ceres runs (6 local parameters), parallelised over points by one warp.
// Latency / FP64-issue bound: no bandwidth claim, no tensor cores.
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int PNP_THREADS = 128;
constexpr int MAX_RANSAC_ITERS = 1000;
constexpr int PNP_SMEM_BYTES_PER_POINT = 5 * 8 + 2;   // xs[3], ys[2] in FP64 + two selection bytes, all in dynamic shared memory

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double dot3(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 cross3(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ D3 unit3(D3 a) { const double n = sqrt(dot3(a, a)); return {a.x / n, a.y / n, a.z / n}; }
struct Mat3 { double m[9]; };
__device__ __forceinline__ D3 mv(const Mat3& A, D3 v) {
  return {A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z, A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
__device__ __forceinline__ Mat3 mm(const Mat3& A, const Mat3& B) {
  Mat3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
__device__ Mat3 inv3(const Mat3& a) {   // adjugate / determinant (utils/cvl/matrix.h:632-651)
  Mat3 M;
  M.m[0] = a.m[4] * a.m[8] - a.m[5] * a.m[7]; M.m[1] = a.m[2] * a.m[7] - a.m[1] * a.m[8]; M.m[2] = a.m[1] * a.m[5] - a.m[2] * a.m[4];
  M.m[3] = a.m[5] * a.m[6] - a.m[3] * a.m[8]; M.m[4] = a.m[0] * a.m[8] - a.m[2] * a.m[6]; M.m[5] = a.m[2] * a.m[3] - a.m[0] * a.m[5];
  M.m[6] = a.m[3] * a.m[7] - a.m[4] * a.m[6]; M.m[7] = a.m[1] * a.m[6] - a.m[0] * a.m[7]; M.m[8] = a.m[0] * a.m[4] - a.m[1] * a.m[3];
  const double idet = 1.0 / (a.m[0] * M.m[0] + a.m[1] * M.m[3] + a.m[2] * M.m[6]);
#pragma unroll
  for (int i = 0; i < 9; ++i) M.m[i] *= idet;
  return M;
}

// x^2 + b x + c (lambdatwist/solve_cubic.h:13-33)
__device__ __forceinline__ bool root2real(double b, double c, double& r1, double& r2) {
  const double v = b * b - 4.0 * c;
  if (v < 0) { r1 = r2 = 0.5 * b; return false; }
  const double y = sqrt(v);
  if (b < 0) { r1 = 0.5 * (-b + y); r2 = 0.5 * (-b - y); }
  else { r1 = 2.0 * c / (-b + y); r2 = 2.0 * c / (-b - y); }
  return true;
}
// sharpest real root of r^3 + b r^2 + c r + d, Newton from a chosen start (solve_cubic.h:134-209)
__device__ double cubick(double b, double c, double d) {
  double r0;
  if (b * b >= 3.0 * c) {
    const double v = sqrt(b * b - 3.0 * c);
    const double t1 = (-b - v) / 3.0;
    double k = ((t1 + b) * t1 + c) * t1 + d;
    if (k > 0.0) {
      r0 = t1 - sqrt(-k / (3.0 * t1 + b));
    } else {
      const double t2 = (-b + v) / 3.0;
      k = ((t2 + b) * t2 + c) * t2 + d;
      r0 = t2 + sqrt(-k / (3.0 * t2 + b));
    }
  } else {
    r0 = -b / 3.0;
    if (fabs((3.0 * r0 + 2.0 * b) * r0 + c) < 1e-4) r0 += 1;
  }
  for (int cnt = 0; cnt < 50; ++cnt) {
    const double fx = ((r0 + b) * r0 + c) * r0 + d;
    if (cnt < 7 || fabs(fx) > 1e-13) {
      const double fpx = (3.0 * r0 + 2.0 * b) * r0 + c;
      r0 -= fx / fpx;
    } else break;
  }
  return r0;
}
// symmetric 3x3 with a known zero eigenvalue (solve_eig0.h:11-84); E columns = eigenvectors
__device__ void eig_known0(const Mat3& x, Mat3& E, double& L0, double& L1) {
  D3 v3 = {x.m[3] * x.m[7] - x.m[6] * x.m[4], x.m[6] * x.m[1] - x.m[7] * x.m[0], x.m[4] * x.m[0] - x.m[3] * x.m[1]};
  v3 = unit3(v3);
  const double x01s = x.m[1] * x.m[1];
  const double b = -x.m[0] - x.m[4] - x.m[8];
  const double c = -x01s - x.m[2] * x.m[2] - x.m[5] * x.m[5] + x.m[0] * (x.m[4] + x.m[8]) + x.m[4] * x.m[8];
  double e1, e2;
  root2real(b, c, e1, e2);
  if (fabs(e1) < fabs(e2)) { const double t = e1; e1 = e2; e2 = t; }
  L0 = e1; L1 = e2;
  const double mx0011 = -x.m[0] * x.m[4];
  const double prec0 = x.m[1] * x.m[5] - x.m[2] * x.m[4];
  const double prec1 = x.m[1] * x.m[2] - x.m[0] * x.m[5];
  D3 vv[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double e = i == 0 ? e1 : e2;
    const double tmp = 1.0 / (e * (x.m[0] + x.m[4]) + mx0011 - e * e + x01s);
    double a1 = -(e * x.m[2] + prec0) * tmp;
    double a2 = -(e * x.m[5] + prec1) * tmp;
    const double rn = 1.0 / sqrt(a1 * a1 + a2 * a2 + 1.0);
    vv[i] = {a1 * rn, a2 * rn, rn};
  }
  E.m[0] = vv[0].x; E.m[1] = vv[1].x; E.m[2] = v3.x;
  E.m[3] = vv[0].y; E.m[4] = vv[1].y; E.m[5] = v3.y;
  E.m[6] = vv[0].z; E.m[7] = vv[1].z; E.m[8] = v3.z;
}
// refine_lambda.h:21-102
__device__ void refine_lambdas(double (&L)[3], double a12, double a13, double a23, double b12, double b13, double b23) {
  for (int it = 0; it < 5; ++it) {
    const double l1 = L[0], l2 = L[1], l3 = L[2];
    const double r1 = l1 * l1 + l2 * l2 + b12 * l1 * l2 - a12;
    const double r2 = l1 * l1 + l3 * l3 + b13 * l1 * l3 - a13;
    const double r3 = l2 * l2 + l3 * l3 + b23 * l2 * l3 - a23;
    if (fabs(r1) + fabs(r2) + fabs(r3) < 1e-10) break;
    const double v0 = 2.0 * l1 + b12 * l2, v1 = 2.0 * l2 + b12 * l1, v3 = 2.0 * l1 + b13 * l3;
    const double v5 = 2.0 * l3 + b13 * l1, v7 = 2.0 * l2 + b23 * l3, v8 = 2.0 * l3 + b23 * l2;
    const double det = 1.0 / (-v0 * v5 * v7 - v1 * v3 * v8);
    const double n1 = l1 - det * ((-v5 * v7) * r1 + (-v1 * v8) * r2 + (v1 * v5) * r3);
    const double n2 = l2 - det * ((-v3 * v8) * r1 + (v0 * v8) * r2 + (-v0 * v5) * r3);
    const double n3 = l3 - det * ((v3 * v7) * r1 + (-v0 * v7) * r2 + (-v1 * v3) * r3);
    const double q1 = n1 * n1 + n2 * n2 + b12 * n1 * n2 - a12;
    const double q2 = n1 * n1 + n3 * n3 + b13 * n1 * n3 - a13;
    const double q3 = n2 * n2 + n3 * n3 + b23 * n2 * n3 - a23;
    if (fabs(q1) + fabs(q2) + fabs(q3) > fabs(r1) + fabs(r2) + fabs(r3)) break;
    L[0] = n1; L[1] = n2; L[2] = n3;
  }
}

struct QPose { double q[4]; double t[3]; };   // q = (w,x,y,z), x' = R(q) x + t, q not renormalised by R()

__device__ __forceinline__ Mat3 quat_R(const double* q) {   // utils/cvl/rotation_helpers.h:213-240
  const double aa = q[0] * q[0], ab = q[0] * q[1], ac = q[0] * q[2], ad = q[0] * q[3];
  const double bb = q[1] * q[1], bc = q[1] * q[2], bd = q[1] * q[3], cc = q[2] * q[2], cd = q[2] * q[3], dd = q[3] * q[3];
  Mat3 R;
  R.m[0] = aa + bb - cc - dd; R.m[1] = 2.0 * (bc - ad); R.m[2] = 2.0 * (ac + bd);
  R.m[3] = 2.0 * (ad + bc); R.m[4] = aa - bb + cc - dd; R.m[5] = 2.0 * (cd - ab);
  R.m[6] = 2.0 * (bd - ac); R.m[7] = 2.0 * (ab + cd); R.m[8] = aa - bb - cc + dd;
  return R;
}
__device__ void R_to_quat(const Mat3& R, double* q) {        // rotation_helpers.h:253-314
  double S;
  const double tr = R.m[0] + R.m[4] + R.m[8] + 1.0;
  if (tr > 1e-7) {
    S = 0.5 / sqrt(tr);
    q[0] = 0.25 / S; q[1] = (R.m[7] - R.m[5]) * S; q[2] = (R.m[2] - R.m[6]) * S; q[3] = (R.m[3] - R.m[1]) * S;
  } else if (R.m[0] > R.m[4] && R.m[0] > R.m[8]) {
    S = sqrt(1.0 + R.m[0] - R.m[4] - R.m[8]) * 2.0;
    q[0] = (R.m[7] - R.m[5]) / S; q[1] = 0.25 * S; q[2] = (R.m[3] + R.m[1]) / S; q[3] = (R.m[2] + R.m[6]) / S;
  } else if (R.m[4] > R.m[8]) {
    S = sqrt(1.0 + R.m[4] - R.m[0] - R.m[8]) * 2.0;
    q[0] = (R.m[2] - R.m[6]) / S; q[1] = (R.m[3] + R.m[1]) / S; q[2] = 0.25 * S; q[3] = (R.m[7] + R.m[5]) / S;
  } else {
    S = sqrt(1.0 + R.m[8] - R.m[0] - R.m[4]) * 2.0;
    q[0] = (R.m[3] - R.m[1]) / S; q[1] = (R.m[2] + R.m[6]) / S; q[2] = (R.m[7] + R.m[5]) / S; q[3] = 0.25 * S;
  }
}
__device__ __forceinline__ bool pose_ok(const QPose& P) {     // Pose::isnormal, utils/cvl/pose.h:381-386
  bool fin = true;
#pragma unroll
  for (int i = 0; i < 4; ++i) fin &= isfinite(P.q[i]);
#pragma unroll
  for (int i = 0; i < 3; ++i) fin &= isfinite(P.t[i]);
  if (!fin) return false;
  const double len = sqrt(P.q[0] * P.q[0] + P.q[1] * P.q[1] + P.q[2] * P.q[2] + P.q[3] * P.q[3]);
  return !(len - 1.0 > 1e-5);
}

// P3P (lambdatwist.p3p.h:33-339) + 4th-point disambiguation (p4p.cpp:11-60).
// xs/ys live in shared memory.  Returns identity when no root survives.
__device__ QPose solve_p4p(const double* xs, const double* ys, const int (&idx)[4]) {
  D3 X[3], Y[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    X[i] = {xs[3 * idx[i]], xs[3 * idx[i] + 1], xs[3 * idx[i] + 2]};
    Y[i] = unit3(D3{ys[2 * idx[i]], ys[2 * idx[i] + 1], 1.0});
  }
  const double b12 = -2.0 * dot3(Y[0], Y[1]), b13 = -2.0 * dot3(Y[0], Y[2]), b23 = -2.0 * dot3(Y[1], Y[2]);
  const D3 d12 = X[0] - X[1], d13 = X[0] - X[2], d23 = X[1] - X[2];
  const D3 d12xd13 = cross3(d12, d13);
  const double a12 = dot3(d12, d12), a13 = dot3(d13, d13), a23 = dot3(d23, d23);
  const double c31 = -0.5 * b13, c23 = -0.5 * b23, c12 = -0.5 * b12;
  const double blob = c12 * c23 * c31 - 1.0;
  const double s31 = 1.0 - c31 * c31, s23 = 1.0 - c23 * c23, s12 = 1.0 - c12 * c12;
  double p3 = a13 * (a23 * s31 - a13 * s23);
  double p2 = 2.0 * blob * a23 * a13 + a13 * (2.0 * a12 + a13) * s23 + a23 * (a23 - a12) * s31;
  double p1 = a23 * (a13 - a23) * s12 - a12 * a12 * s23 - 2.0 * a12 * (blob * a23 + a13 * s23);
  double p0 = a12 * (a12 * s23 - a23 * s12);
  p3 = 1.0 / p3; p2 *= p3; p1 *= p3; p0 *= p3;
  const double gr = cubick(p2, p1, p0);

  Mat3 A;
  A.m[0] = a23 * (1.0 - gr); A.m[1] = (a23 * b12) * 0.5; A.m[2] = (a23 * b13 * gr) * (-0.5);
  A.m[4] = a23 - a12 + a13 * gr; A.m[5] = b23 * (a13 * gr - a12) * 0.5; A.m[8] = gr * (a13 - a23) - a12;
  A.m[3] = A.m[1]; A.m[6] = A.m[2]; A.m[7] = A.m[5];
  Mat3 V; double L0, L1;
  eig_known0(A, V, L0, L1);
  const double vroot = sqrt(fmax(0.0, -L1 / L0));

  double Ls[4][3];
  int valid = 0;
#pragma unroll
  for (int br = 0; br < 2; ++br) {
    const double s = br == 0 ? vroot : -vroot;
    const double w2 = 1.0 / (s * V.m[1] - V.m[0]);
    const double w0 = (V.m[3] - s * V.m[4]) * w2;
    const double w1 = (V.m[6] - s * V.m[7]) * w2;
    const double a = 1.0 / ((a13 - a12) * w1 * w1 - a12 * b13 * w1 - a12);
    const double b = (a13 * b12 * w1 - a12 * b13 * w0 - 2.0 * w0 * w1 * (a12 - a13)) * a;
    const double c = ((a13 - a12) * w0 * w0 + a13 * b12 * w0 + a13) * a;
    if (b * b - 4.0 * c >= 0) {
      double tau[2];
      root2real(b, c, tau[0], tau[1]);
#pragma unroll
      for (int ti = 0; ti < 2; ++ti) {
        if (tau[ti] > 0) {
          const double d = a23 / (tau[ti] * (b23 + tau[ti]) + 1.0);
          if (br == 1 && !(d > 0)) continue;          // only the -v branch guards d (p3p.h:252,266)
          const double l2 = sqrt(d), l3 = tau[ti] * l2, l1 = w0 * l2 + w1 * l3;
          if (l1 >= 0) { Ls[valid][0] = l1; Ls[valid][1] = l2; Ls[valid][2] = l3; ++valid; }
        }
      }
    }
  }
  Mat3 Xm;
  Xm.m[0] = d12.x; Xm.m[1] = d13.x; Xm.m[2] = d12xd13.x;
  Xm.m[3] = d12.y; Xm.m[4] = d13.y; Xm.m[5] = d12xd13.y;
  Xm.m[6] = d12.z; Xm.m[7] = d13.z; Xm.m[8] = d12xd13.z;
  Xm = inv3(Xm);

  const double y4x = ys[2 * idx[3]], y4y = ys[2 * idx[3] + 1];
  const D3 x4 = {xs[3 * idx[3]], xs[3 * idx[3] + 1], xs[3 * idx[3] + 2]};
  QPose best = {{1, 0, 0, 0}, {0, 0, 0}};
  double e0 = 1.7976931348623157e308;
  for (int i = 0; i < valid; ++i) {
    refine_lambdas(Ls[i], a12, a13, a23, b12, b13, b23);
    const D3 ry1 = Y[0] * Ls[i][0], ry2 = Y[1] * Ls[i][1], ry3 = Y[2] * Ls[i][2];
    const D3 yd1 = ry1 - ry2, yd2 = ry1 - ry3, ydx = cross3(yd1, yd2);
    Mat3 Ym;
    Ym.m[0] = yd1.x; Ym.m[1] = yd2.x; Ym.m[2] = ydx.x;
    Ym.m[3] = yd1.y; Ym.m[4] = yd2.y; Ym.m[5] = ydx.y;
    Ym.m[6] = yd1.z; Ym.m[7] = yd2.z; Ym.m[8] = ydx.z;
    const Mat3 R = mm(Ym, Xm);
    const D3 T = ry1 - mv(R, X[0]);
    QPose tmp;
    R_to_quat(R, tmp.q);
    const double nq = sqrt(tmp.q[0] * tmp.q[0] + tmp.q[1] * tmp.q[1] + tmp.q[2] * tmp.q[2] + tmp.q[3] * tmp.q[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) tmp.q[k] /= nq;
    tmp.t[0] = T.x; tmp.t[1] = T.y; tmp.t[2] = T.z;
    if (!pose_ok(tmp)) continue;
    const D3 xr = mv(quat_R(tmp.q), x4) + D3{tmp.t[0], tmp.t[1], tmp.t[2]};
    if (xr.z < 0) continue;
    const double dx = xr.x / xr.z - y4x, dy = xr.y / xr.z - y4y;
    const double e = dx * dx + dy * dy;
    if (isnan(e)) continue;
    if (e < e0) { best = tmp; e0 = e; }
  }
  return best;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// 4 distinct uniform indices, ascending (what get4RandomInRange0's std::set yields, pnp_ransac.cpp:161-183)
__device__ void sample4(uint64_t seed, uint64_t key, uint32_t iter, int n, int (&idx)[4]) {
  const uint64_t base = splitmix64(seed ^ splitmix64(key * 0xD1342543DE82EF95ull + iter));
  int cnt = 0;
  for (uint32_t j = 0; cnt < 4; ++j) {
    const int v = (int)(splitmix64(base + j) % (uint64_t)n);
    bool dup = false;
    for (int k = 0; k < cnt; ++k) dup |= (idx[k] == v);
    if (!dup) idx[cnt++] = v;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3 - a; ++b)
      if (idx[b] > idx[b + 1]) { const int t = idx[b]; idx[b] = idx[b + 1]; idx[b + 1] = t; }
}
__device__ int ransac_budget(double inlier_ratio) {             // parameters.h:76-102
  double p_inlier = fmin(0.9, inlier_ratio * 0.9);
  p_inlier = fmin(fmax(p_inlier, 1e-2), 1 - 1e-8);
  if (p_inlier < 0.01) return 1000;
  const double p_failure = fmin(fmax(1.0 - 0.99999, 1e-8), 0.01);
  const double p_good = pow(p_inlier, 4.0);
  const double it = ceil(log(p_failure) / log(1.0 - p_good)) + 50;
  if (it < 100) return 100;
  if (it > 1000) return 1000;
  return (int)it;
}
__device__ int count_inliers(const double* xs, const double* ys, int n, double thr2, const QPose& P) {  // pnp_ransac.cpp:41-87
  const Mat3 R = quat_R(P.q);
  int inl = 0;
  for (int i = 0; i < n; ++i) {
    const D3 r = mv(R, D3{xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]});
    const double x = r.x + P.t[0], y = r.y + P.t[1], z = r.z + P.t[2];
    const double iz = 1.0 / z;
    if (iz < 0) continue;
    const double e1 = x * iz - ys[2 * i], e2 = y * iz - ys[2 * i + 1];
    inl += (e1 * e1 + e2 * e2 < thr2) ? 1 : 0;
  }
  return inl;
}

// ---------------------------------------------------------------- LM refine (one warp)
// Cost of pnp_ransac.cpp:120-139 with parameter blocks q[4] (+QuaternionParameterization) and t[3]
// minimised by the trust-region LM that ceres::Solve runs with the options at :269-276,:311-319.
__device__ __forceinline__ QPose pose_plus(const QPose& P, const double* d) {
  QPose o = P;
  const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (nd > 0.0) {
    const double s = sin(nd) / nd, c = cos(nd);
    const double q1 = s * d[0], q2 = s * d[1], q3 = s * d[2];
    const double* x = P.q;
    o.q[0] = c * x[0] - q1 * x[1] - q2 * x[2] - q3 * x[3];
    o.q[1] = c * x[1] + q1 * x[0] + q2 * x[3] - q3 * x[2];
    o.q[2] = c * x[2] - q1 * x[3] + q2 * x[0] + q3 * x[1];
    o.q[3] = c * x[3] + q1 * x[2] - q2 * x[1] + q3 * x[0];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) o.t[k] = P.t[k] + d[3 + k];
  return o;
}
// residual + local 2x6 Jacobian of one point
__device__ __forceinline__ void point_rj(const QPose& P, const Mat3& R, const double* X3, const double* y2, double (&r)[2], double (&J)[2][6]) {
  const double a = P.q[0], b = P.q[1], c = P.q[2], d = P.q[3];
  const double X = X3[0], Y = X3[1], Z = X3[2];
  const D3 pr = mv(R, D3{X, Y, Z});
  const double x = pr.x + P.t[0], y = pr.y + P.t[1], z = pr.z + P.t[2];
  const double iz = 1.0 / z;
  r[0] = x * iz - y2[0]; r[1] = y * iz - y2[1];
  const double dP[3][4] = {
      {2 * (a * X - d * Y + c * Z), 2 * (b * X + c * Y + d * Z), 2 * (-c * X + b * Y + a * Z), 2 * (-d * X - a * Y + b * Z)},
      {2 * (d * X + a * Y - b * Z), 2 * (c * X - b * Y - a * Z), 2 * (b * X + c * Y + d * Z), 2 * (a * X - d * Y + c * Z)},
      {2 * (-c * X + b * Y + a * Z), 2 * (d * X + a * Y - b * Z), 2 * (-a * X + d * Y - c * Z), 2 * (b * X + c * Y + d * Z)}};
  const double du2 = -x * iz * iz, dv2 = -y * iz * iz;
  double Jq[2][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { Jq[0][k] = iz * dP[0][k] + du2 * dP[2][k]; Jq[1][k] = iz * dP[1][k] + dv2 * dP[2][k]; }
  const double PJ[4][3] = {{-b, -c, -d}, {a, d, -c}, {-d, a, b}, {c, -b, a}};   // QuaternionParameterization::ComputeJacobian
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
    for (int k = 0; k < 3; ++k) J[rr][k] = Jq[rr][0] * PJ[0][k] + Jq[rr][1] * PJ[1][k] + Jq[rr][2] * PJ[2][k] + Jq[rr][3] * PJ[3][k];
  }
  J[0][3] = iz; J[0][4] = 0; J[0][5] = du2;
  J[1][3] = 0; J[1][4] = iz; J[1][5] = dv2;
}
__device__ bool chol6_solve(double (&A)[36], double (&b)[6]) {
  for (int j = 0; j < 6; ++j) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; ++k) d -= A[j * 6 + k] * A[j * 6 + k];
    if (!(d > 0) || !isfinite(d)) return false;
    d = sqrt(d);
    A[j * 6 + j] = d;
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= A[i * 6 + k] * A[j * 6 + k];
      A[i * 6 + j] = s / d;
    }
  }
  for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[i * 6 + k] * b[k]; b[i] = s / A[i * 6 + i]; }
  for (int i = 5; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < 6; ++k) s -= A[k * 6 + i] * b[k]; b[i] = s / A[i * 6 + i]; }
  return true;
}

// Warp-collective LM.  sel[i] != 0 selects the residual points.  All lanes hold the same P.
__device__ int warp_lm(QPose& P, const double* xs, const double* ys, const uint8_t* sel, int n, int max_iter,
                       double ftol, double gtol) {
  const int lane = threadIdx.x & 31;
  const double kMaxD = 1.7976931348623157e308;
  double scale[6], diag[6];
  double radius = 1e4, decrease = 2.0, cost = 0, gmax = 0;
  bool reuse_diag = false;
  int invalid = 0, iter = 0;
  // accumulators shared by every evaluation: cost, g = J^T r (6), col sq norms via H diag, H (21 upper)
  auto evaluate = [&](const QPose& Q, const double* sc, double& cst, double (&g)[6], double (&H)[36], bool want_H) {
    const Mat3 R = quat_R(Q.q);
    double acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0;
    for (int i = lane; i < n; i += 32) {
      if (!sel[i]) continue;
      double r[2], J[2][6];
      point_rj(Q, R, xs + 3 * i, ys + 2 * i, r, J);
      acc[0] += r[0] * r[0] + r[1] * r[1];
      if (want_H) {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          const double ja0 = J[0][a] * sc[a], ja1 = J[1][a] * sc[a];
          acc[1 + a] += ja0 * r[0] + ja1 * r[1];
          int o = 7 + a * 6 - a * (a - 1) / 2 - a;   // start of row a in the packed upper triangle (relative index a..5)
#pragma unroll
          for (int b = a; b < 6; ++b) acc[o + b] += ja0 * (J[0][b] * sc[b]) + ja1 * (J[1][b] * sc[b]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = warp_sum(acc[k]);
    cst = 0.5 * acc[0];
    if (want_H) {
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        g[a] = acc[1 + a];
        int o = 7 + a * 6 - a * (a - 1) / 2 - a;
#pragma unroll
        for (int b = a; b < 6; ++b) { H[a * 6 + b] = acc[o + b]; H[b * 6 + a] = acc[o + b]; }
      }
    }
  };
  double ones[6] = {1, 1, 1, 1, 1, 1};
  double g[6], H[36];
  evaluate(P, ones, cost, g, H, true);          // unscaled: g = J^T r, diag(H) = column square norms
#pragma unroll
  for (int j = 0; j < 6; ++j) scale[j] = 1.0 / (1.0 + sqrt(H[j * 7]));
  auto grad_max = [&](const QPose& Q, const double (&gu)[6]) {   // || x - Plus(x, -g) ||_inf, g unscaled
    double ng[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) ng[j] = -gu[j];
    const QPose Qp = pose_plus(Q, ng);
    double mx = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) mx = fmax(mx, fabs(Q.q[k] - Qp.q[k]));
#pragma unroll
    for (int k = 0; k < 3; ++k) mx = fmax(mx, fabs(Q.t[k] - Qp.t[k]));
    return mx;
  };
  gmax = grad_max(P, g);
  // scaled system for the first step
  double gs[6], Hs[36];
#pragma unroll
  for (int a = 0; a < 6; ++a) { gs[a] = g[a] * scale[a]; for (int b = 0; b < 6; ++b) Hs[a * 6 + b] = H[a * 6 + b] * scale[a] * scale[b]; }
  while (true) {
    if (iter >= max_iter) break;
    if (gmax <= gtol) break;
    if (radius < 1e-32) break;
    ++iter;
    if (!reuse_diag) {
#pragma unroll
      for (int j = 0; j < 6; ++j) diag[j] = fmin(fmax(Hs[j * 7], 1e-6), 1e32);
    }
    double A[36], rhs[6];
#pragma unroll
    for (int k = 0; k < 36; ++k) A[k] = Hs[k];
#pragma unroll
    for (int j = 0; j < 6; ++j) { A[j * 7] += diag[j] / radius; rhs[j] = gs[j]; }
    reuse_diag = true;
    const bool ok = chol6_solve(A, rhs);
    double step[6], model_change = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) step[j] = -rhs[j];
    if (ok) {
      // -(Js s)^T (r + Js s / 2) = -(s^T gs) - 0.5 s^T Hs s
      double sg = 0, sHs = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) { sg += step[a] * gs[a]; double t = 0; for (int b = 0; b < 6; ++b) t += Hs[a * 6 + b] * step[b]; sHs += step[a] * t; }
      model_change = -sg - 0.5 * sHs;
    }
    if (!ok || !(model_change > 0.0)) {
      if (++invalid >= 5) break;
      radius /= decrease; decrease *= 2.0;
      continue;
    }
    invalid = 0;
    double delta[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) delta[j] = step[j] * scale[j];
    const QPose cand = pose_plus(P, delta);
    double ccost, gd[6], Hd[36];
    evaluate(cand, ones, ccost, gd, Hd, true);
    if (!isfinite(ccost)) ccost = kMaxD;
    double sn = 0, xn = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { sn += (P.q[k] - cand.q[k]) * (P.q[k] - cand.q[k]); xn += P.q[k] * P.q[k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { sn += (P.t[k] - cand.t[k]) * (P.t[k] - cand.t[k]); xn += P.t[k] * P.t[k]; }
    if (sqrt(sn) <= 1e-8 * (sqrt(xn) + 1e-8)) break;             // parameter tolerance
    if (fabs(cost - ccost) <= ftol * cost) break;                 // function tolerance (candidate NOT adopted)
    const double rel = (cost - ccost) / model_change;
    if (rel > 1e-3) {
      P = cand; cost = ccost;
#pragma unroll
      for (int a = 0; a < 6; ++a) { g[a] = gd[a]; gs[a] = gd[a] * scale[a]; for (int b = 0; b < 6; ++b) Hs[a * 6 + b] = Hd[a * 6 + b] * scale[a] * scale[b]; }
      gmax = grad_max(P, g);
      radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rel - 1.0, 3.0));
      radius = fmin(1e16, radius);
      decrease = 2.0; reuse_diag = false;
    } else {
      radius /= decrease; decrease *= 2.0; reuse_diag = true;
    }
  }
  return iter;
}

// NT threads = NT RANSAC hypotheses evaluated per batch.  The accept rule is replayed sequentially, so the result does not depend on NT:
// 128 for large batches of objects (two CTAs per SM), 256 when a call brings only a few objects (the drop-in's one-object pnp() calls:
// an object with outliers needs ~400 iterations = 2 batches instead of 4, and the kernel is latency-bound).
template <int NT>
__global__ void __launch_bounds__(NT)
pnp_kernel(const double* __restrict__ xs_all, const double* __restrict__ ys_all, const int32_t* __restrict__ offsets,
           const int32_t* __restrict__ npts, int max_pts, double threshold, uint64_t seed, const uint64_t* __restrict__ keys,
           double* __restrict__ T_out, int32_t* __restrict__ stats) {
  extern __shared__ double pnp_dyn[];
  double* xs = pnp_dyn;
  double* ys = xs + 3 * max_pts;
  uint8_t* sel = reinterpret_cast<uint8_t*>(ys + 2 * max_pts);
  uint8_t* sel0 = sel + max_pts;
  __shared__ int counts[NT];
  __shared__ int sh_best_inl, sh_best_it, sh_iters, sh_done;
  const int obj = blockIdx.x;
  const int off = offsets[obj];
  const int n = npts ? npts[obj] : offsets[obj + 1] - off;
  if (n > max_pts || n < 0) {     // never truncate: the object is reported as failed (identity, like the reference's failure) with best_inliers = -1
    if (threadIdx.x < 16) T_out[16 * obj + threadIdx.x] = (threadIdx.x % 5 == 0) ? 1.0 : 0.0;
    if (threadIdx.x == 0 && stats) { stats[5 * obj] = -1; stats[5 * obj + 1] = -1; stats[5 * obj + 2] = 0; stats[5 * obj + 3] = -1; stats[5 * obj + 4] = -1; }
    return;
  }
  const uint64_t key = keys ? keys[obj] : (uint64_t)obj;
  for (int i = threadIdx.x; i < 3 * n; i += NT) xs[i] = xs_all[3 * off + i];
  for (int i = threadIdx.x; i < 2 * n; i += NT) ys[i] = ys_all[2 * off + i];
  if (threadIdx.x == 0) { sh_best_inl = 0; sh_best_it = -1; sh_iters = n >= 4 ? ransac_budget(0.0) : 0; sh_done = 0; }
  __syncthreads();
  const double thr2 = threshold * threshold;

  // ---- RANSAC in batches of PNP_THREADS hypotheses, sequential accept rule replayed per batch ----
  int total = 0;
  for (int base = 0; base < MAX_RANSAC_ITERS; base += NT) {
    if (base >= sh_iters) break;
    const int it = base + threadIdx.x;
    int cnt = -1;
    if (it < sh_iters) {
      int idx[4];
      sample4(seed, key, (uint32_t)it, n, idx);
      const QPose P = solve_p4p(xs, ys, idx);
      if (pose_ok(P)) cnt = count_inliers(xs, ys, n, thr2, P);
    }
    counts[threadIdx.x] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      int iters = sh_iters, best = sh_best_inl, bit = sh_best_it, i = 0;
      for (; i < NT && base + i < iters; ++i) {
        if (counts[i] > best) { best = counts[i]; bit = base + i; iters = ransac_budget(best / (double)n); }
      }
      sh_iters = iters; sh_best_inl = best; sh_best_it = bit; sh_done = base + i;
    }
    __syncthreads();
  }
  total = sh_done;

  // ---- winner + refine: warp 0 ----
  if (threadIdx.x < 32) {
    QPose best = {{1, 0, 0, 0}, {0, 0, 0}};
    int rit0 = -1, rit1 = -1;
    if (sh_best_it >= 0) {
      int idx[4];
      sample4(seed, key, (uint32_t)sh_best_it, n, idx);
      best = solve_p4p(xs, ys, idx);        // every lane recomputes the same pose
    }
    if (sh_best_inl > 3) {
      // pnp_ransac.cpp:254-261 inlier selection under the RANSAC pose
      auto classify = [&](const QPose& P, uint8_t* out) {
        const Mat3 R = quat_R(P.q);
        for (int i = threadIdx.x; i < n; i += 32) {
          const D3 r = mv(R, D3{xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]});
          const double x = r.x + P.t[0], y = r.y + P.t[1], z = r.z + P.t[2];
          bool in = !(z < 0);
          const double dx = x / z - ys[2 * i], dy = y / z - ys[2 * i + 1];
          if (dx * dx + dy * dy > thr2) in = false;
          out[i] = in ? 1 : 0;
        }
        __syncwarp();
      };
      classify(best, sel0);
      rit0 = warp_lm(best, xs, ys, sel0, n, 5, 1e-6, 1e-6);
      classify(best, sel);
      int deltas = 0, nin = 0;
      for (int i = 0; i < n; ++i) { deltas += (sel[i] != sel0[i]); nin += sel[i]; }
      if (!(deltas < 0.05 * (double)nin)) rit1 = warp_lm(best, xs, ys, sel, n, 3, 1e-8, 1e-8);   // :303
    }
    if (threadIdx.x == 0) {
      const Mat3 R = quat_R(best.q);
      double* T = T_out + 16 * obj;
#pragma unroll
      for (int r = 0; r < 3; ++r) { T[4 * r] = R.m[3 * r]; T[4 * r + 1] = R.m[3 * r + 1]; T[4 * r + 2] = R.m[3 * r + 2]; T[4 * r + 3] = best.t[r]; }
      T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
      if (stats) {
        stats[5 * obj] = sh_best_inl; stats[5 * obj + 1] = sh_best_it; stats[5 * obj + 2] = total;
        stats[5 * obj + 3] = rit0; stats[5 * obj + 4] = rit1;
      }
    }
  }
}

}  // namespace

// counts == nullptr: object o owns rows offsets[o]..offsets[o+1]; else rows offsets[o]..offsets[o]+counts[o]
int launch_pnp_batch_counts(suo_ctx* ctx, const double* xs, const double* ys, const int32_t* offsets,
                            const int32_t* counts, int n_obj, double threshold, uint64_t seed, const uint64_t* obj_keys,
                            double* T_out, int32_t* stats, cudaStream_t s, int max_pts) {
  if (n_obj <= 0) return SUO_OK;
  max_pts = (std::max(max_pts, 4) + 7) & ~7;
  const size_t smem = (size_t)max_pts * PNP_SMEM_BYTES_PER_POINT;
  if (smem > 200 * 1024) { ctx->set_error("suo_pnp_batch: more than " + std::to_string(200 * 1024 / PNP_SMEM_BYTES_PER_POINT) + " points in one object", __FILE__, __LINE__); return SUO_E_INVALID; }
  static bool configured[64] = {};
  if (smem > 40 * 1024 && first_use_on_device(configured, nullptr)) {
    SUO_CUDA_TRY(ctx, cudaFuncSetAttribute(pnp_kernel<PNP_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SUO_CUDA_TRY(ctx, cudaFuncSetAttribute(pnp_kernel<2 * PNP_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  if (n_obj <= 64)
    pnp_kernel<2 * PNP_THREADS><<<n_obj, 2 * PNP_THREADS, smem, s>>>(xs, ys, offsets, counts, max_pts, threshold, seed, obj_keys, T_out, stats);
  else
    pnp_kernel<PNP_THREADS><<<n_obj, PNP_THREADS, smem, s>>>(xs, ys, offsets, counts, max_pts, threshold, seed, obj_keys, T_out, stats);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_pnp_batch(suo_ctx* ctx, const double* xs, const double* ys, const int32_t* offsets, int n_obj,
                     double threshold, uint64_t seed, const uint64_t* obj_keys, double* T_out, int32_t* stats,
                     cudaStream_t s, int max_pts) {
  return launch_pnp_batch_counts(ctx, xs, ys, offsets, nullptr, n_obj, threshold, seed, obj_keys, T_out, stats, s, max_pts);
}
