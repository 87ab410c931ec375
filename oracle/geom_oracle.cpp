// TEST INFRASTRUCTURE ONLY.  CPU FP64 restatement of the reference's geometry
// path: Lambda-Twist P3P -> P4P -> RANSAC -> Ceres
// LM refine, and the g2o Levenberg bundle adjustment with the two custom
// object-SLAM edges plus ObjectSLAM.optimize()'s round/chi2 orchestration.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.  The product path
// (suo_slam_b200/, libsuo_b200.so) never links or calls it.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).  Plain C++17, no dependencies, single-threaded (the
// reference's own defaults: g2o OpenMP OFF thirdparty/g2opy/CMakeLists.txt:126,
// Ceres num_threads=1).
//
// PARITY PINS
//  * P3P/P4P: validated against the reference's own p4p.cpp + lambdatwist/*.h
//    compiled in place into oracle/_ref/libref_p4p.so (oracle/Makefile) and
//    against the golden pose of thirdparty/lambdatwist/test_pnp.py:5-10.
//  * RANSAC driver: same law as pnp_ransac.cpp:188-232 but the 4-subsets come
//    from a counter-based generator (the reference uses a process-global
//    std::default_random_engine whose draw order depends on call history,
//    utils/random.h:68-86) => parity is on the refined pose / inlier set.
//  * Ceres (pnp_ransac.cpp:240-326) and Eigen/CHOLMOD (g2o) are external and
//    NOT under /root/reference: "parity unpinned" beyond the golden vector.
//    Ceres semantics restated are those of Ceres 1.14 / 2.x
//    trust_region_minimizer.cc + levenberg_marquardt_strategy.cc defaults.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

// ----------------------------------------------------------------- small math
struct V3 { double x, y, z; };
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline V3 normalized(V3 a) { double n = std::sqrt(dot(a, a)); return {a.x / n, a.y / n, a.z / n}; }

struct M3 { double m[9]; double& operator()(int r, int c) { return m[3 * r + c]; } double operator()(int r, int c) const { return m[3 * r + c]; } };
static inline V3 mul(const M3& A, V3 v) { return {A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z, A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z}; }
static inline M3 mul(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += A(i, k) * B(k, j); C(i, j) = s; }
  return C;
}
// utils/cvl/matrix.h:632-651 (adjugate / determinant)
static M3 inverse3(const M3& a) {
  M3 M;
  M(0, 0) = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1); M(0, 1) = a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2); M(0, 2) = a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1);
  M(1, 0) = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2); M(1, 1) = a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0); M(1, 2) = a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2);
  M(2, 0) = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0); M(2, 1) = a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1); M(2, 2) = a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0);
  double idet = 1.0 / (a(0, 0) * M(0, 0) + a(0, 1) * M(1, 0) + a(0, 2) * M(2, 0));
  for (double& v : M.m) v *= idet;
  return M;
}

// ------------------------------------------------- Lambda-Twist P3P building blocks
// lambdatwist/solve_cubic.h:13-33
static bool root2real(double b, double c, double& r1, double& r2) {
  double v = b * b - 4.0 * c;
  if (v < 0) { r1 = r2 = 0.5 * b; return false; }
  double y = std::sqrt(v);
  if (b < 0) { r1 = 0.5 * (-b + y); r2 = 0.5 * (-b - y); }
  else { r1 = 2.0 * c / (-b + y); r2 = 2.0 * c / (-b - y); }
  return true;
}

// lambdatwist/solve_cubic.h:134-209: one (sharpest) real root of r^3+b r^2+c r+d by Newton
static double cubick(double b, double c, double d) {
  double r0;
  if (b * b >= 3.0 * c) {
    double v = std::sqrt(b * b - 3.0 * c);
    double t1 = (-b - v) / 3.0;
    double k = ((t1 + b) * t1 + c) * t1 + d;
    if (k > 0.0) {
      r0 = t1 - std::sqrt(-k / (3.0 * t1 + b));
    } else {
      double t2 = (-b + v) / 3.0;
      k = ((t2 + b) * t2 + c) * t2 + d;
      r0 = t2 + std::sqrt(-k / (3.0 * t2 + b));
    }
  } else {
    r0 = -b / 3.0;
    if (std::fabs((3.0 * r0 + 2.0 * b) * r0 + c) < 1e-4) r0 += 1;
  }
  for (unsigned cnt = 0; cnt < 50; ++cnt) {       // KLAS_P3P_CUBIC_SOLVER_ITER
    double fx = ((r0 + b) * r0 + c) * r0 + d;
    if (cnt < 7 || std::fabs(fx) > 1e-13) {       // get_numeric_limit<double>()
      double fpx = (3.0 * r0 + 2.0 * b) * r0 + c;
      r0 -= fx / fpx;
    } else break;
  }
  return r0;
}

// lambdatwist/solve_eig0.h:11-84: eigen-decomposition of a symmetric 3x3 with a known 0 eigenvalue
static void eigwithknown0(const M3& x, M3& E, double L[3]) {
  L[2] = 0;
  V3 v3 = {x.m[3] * x.m[7] - x.m[6] * x.m[4], x.m[6] * x.m[1] - x.m[7] * x.m[0], x.m[4] * x.m[0] - x.m[3] * x.m[1]};
  v3 = normalized(v3);
  double x01_sq = x(0, 1) * x(0, 1);
  double b = -x(0, 0) - x(1, 1) - x(2, 2);
  double c = -x01_sq - x(0, 2) * x(0, 2) - x(1, 2) * x(1, 2) + x(0, 0) * (x(1, 1) + x(2, 2)) + x(1, 1) * x(2, 2);
  double e1, e2;
  root2real(b, c, e1, e2);
  if (std::fabs(e1) < std::fabs(e2)) std::swap(e1, e2);
  L[0] = e1; L[1] = e2;
  double mx0011 = -x(0, 0) * x(1, 1);
  double prec_0 = x(0, 1) * x(1, 2) - x(0, 2) * x(1, 1);
  double prec_1 = x(0, 1) * x(0, 2) - x(0, 0) * x(1, 2);
  double e = e1;
  double tmp = 1.0 / (e * (x(0, 0) + x(1, 1)) + mx0011 - e * e + x01_sq);
  double a1 = -(e * x(0, 2) + prec_0) * tmp;
  double a2 = -(e * x(1, 2) + prec_1) * tmp;
  double rnorm = 1.0 / std::sqrt(a1 * a1 + a2 * a2 + 1.0);
  a1 *= rnorm; a2 *= rnorm;
  V3 v1 = {a1, a2, rnorm};
  double tmp2 = 1.0 / (e2 * (x(0, 0) + x(1, 1)) + mx0011 - e2 * e2 + x01_sq);
  double a21 = -(e2 * x(0, 2) + prec_0) * tmp2;
  double a22 = -(e2 * x(1, 2) + prec_1) * tmp2;
  double rnorm2 = 1.0 / std::sqrt(a21 * a21 + a22 * a22 + 1.0);
  a21 *= rnorm2; a22 *= rnorm2;
  V3 v2 = {a21, a22, rnorm2};
  E = M3{{v1.x, v2.x, v3.x, v1.y, v2.y, v3.y, v1.z, v2.z, v3.z}};
}

// lambdatwist/refine_lambda.h:21-102: Gauss-Newton on the 3 distance constraints
static void gauss_newton_refineL(double L[3], double a12, double a13, double a23, double b12, double b13, double b23, int iterations) {
  for (int i = 0; i < iterations; ++i) {
    double l1 = L[0], l2 = L[1], l3 = L[2];
    double r1 = l1 * l1 + l2 * l2 + b12 * l1 * l2 - a12;
    double r2 = l1 * l1 + l3 * l3 + b13 * l1 * l3 - a13;
    double r3 = l2 * l2 + l3 * l3 + b23 * l2 * l3 - a23;
    if (std::fabs(r1) + std::fabs(r2) + std::fabs(r3) < 1e-10) break;
    double v0 = 2.0 * l1 + b12 * l2, v1 = 2.0 * l2 + b12 * l1;
    double v3 = 2.0 * l1 + b13 * l3, v5 = 2.0 * l3 + b13 * l1;
    double v7 = 2.0 * l2 + b23 * l3, v8 = 2.0 * l3 + b23 * l2;
    double det = 1.0 / (-v0 * v5 * v7 - v1 * v3 * v8);
    double J[9] = {-v5 * v7, -v1 * v8, v1 * v5, -v3 * v8, v0 * v8, -v0 * v5, v3 * v7, -v0 * v7, -v1 * v3};
    double n1 = l1 - det * (J[0] * r1 + J[1] * r2 + J[2] * r3);
    double n2 = l2 - det * (J[3] * r1 + J[4] * r2 + J[5] * r3);
    double n3 = l3 - det * (J[6] * r1 + J[7] * r2 + J[8] * r3);
    double r11 = n1 * n1 + n2 * n2 + b12 * n1 * n2 - a12;
    double r12 = n1 * n1 + n3 * n3 + b13 * n1 * n3 - a13;
    double r13 = n2 * n2 + n3 * n3 + b23 * n2 * n3 - a23;
    if (std::fabs(r11) + std::fabs(r12) + std::fabs(r13) > std::fabs(r1) + std::fabs(r2) + std::fabs(r3)) break;
    L[0] = n1; L[1] = n2; L[2] = n3;
  }
}

// lambdatwist/lambdatwist.p3p.h:33-339.  ys are homogeneous pinhole-normalised bearings.
static int p3p_lambdatwist(V3 y1, V3 y2, V3 y3, V3 x1, V3 x2, V3 x3, M3 Rs[4], V3 Ts[4]) {
  y1 = normalized(y1); y2 = normalized(y2); y3 = normalized(y3);
  double b12 = -2.0 * dot(y1, y2), b13 = -2.0 * dot(y1, y3), b23 = -2.0 * dot(y2, y3);
  V3 d12 = x1 - x2, d13 = x1 - x3, d23 = x2 - x3;
  V3 d12xd13 = cross(d12, d13);
  double a12 = dot(d12, d12), a13 = dot(d13, d13), a23 = dot(d23, d23);
  double c31 = -0.5 * b13, c23 = -0.5 * b23, c12 = -0.5 * b12;
  double blob = c12 * c23 * c31 - 1.0;
  double s31_sq = 1.0 - c31 * c31, s23_sq = 1.0 - c23 * c23, s12_sq = 1.0 - c12 * c12;
  double p3 = a13 * (a23 * s31_sq - a13 * s23_sq);
  double p2 = 2.0 * blob * a23 * a13 + a13 * (2.0 * a12 + a13) * s23_sq + a23 * (a23 - a12) * s31_sq;
  double p1 = a23 * (a13 - a23) * s12_sq - a12 * a12 * s23_sq - 2.0 * a12 * (blob * a23 + a13 * s23_sq);
  double p0 = a12 * (a12 * s23_sq - a23 * s12_sq);
  p3 = 1.0 / p3; p2 *= p3; p1 *= p3; p0 *= p3;     // the "||true" branch at :106
  double g = cubick(p2, p1, p0);

  double A00 = a23 * (1.0 - g), A01 = (a23 * b12) * 0.5, A02 = (a23 * b13 * g) * (-0.5);
  double A11 = a23 - a12 + a13 * g, A12 = b23 * (a13 * g - a12) * 0.5, A22 = g * (a13 - a23) - a12;
  M3 A{{A00, A01, A02, A01, A11, A12, A02, A12, A22}};
  M3 V; double Lam[3];
  eigwithknown0(A, V, Lam);
  double v = std::sqrt(std::max(0.0, -Lam[1] / Lam[0]));

  int valid = 0;
  double Ls[4][3];
  for (int branch = 0; branch < 2; ++branch) {
    double s = branch == 0 ? v : -v;
    double w2 = 1.0 / (s * V(0, 1) - V(0, 0));
    double w0 = (V(1, 0) - s * V(1, 1)) * w2;
    double w1 = (V(2, 0) - s * V(2, 1)) * w2;
    double a = 1.0 / ((a13 - a12) * w1 * w1 - a12 * b13 * w1 - a12);
    double b = (a13 * b12 * w1 - a12 * b13 * w0 - 2.0 * w0 * w1 * (a12 - a13)) * a;
    double c = ((a13 - a12) * w0 * w0 + a13 * b12 * w0 + a13) * a;
    if (b * b - 4.0 * c >= 0) {
      double tau[2];
      root2real(b, c, tau[0], tau[1]);
      for (int ti = 0; ti < 2; ++ti) {
        if (tau[ti] > 0) {
          double d = a23 / (tau[ti] * (b23 + tau[ti]) + 1.0);
          if (branch == 1 && !(d > 0)) continue;   // only the -v branch guards d (:252,:266)
          double l2 = std::sqrt(d);
          double l3 = tau[ti] * l2;
          double l1 = w0 * l2 + w1 * l3;
          if (l1 >= 0) { Ls[valid][0] = l1; Ls[valid][1] = l2; Ls[valid][2] = l3; ++valid; }
        }
      }
    }
  }
  for (int i = 0; i < valid; ++i) gauss_newton_refineL(Ls[i], a12, a13, a23, b12, b13, b23, 5);

  M3 X{{d12.x, d13.x, d12xd13.x, d12.y, d13.y, d12xd13.y, d12.z, d13.z, d12xd13.z}};
  X = inverse3(X);
  for (int i = 0; i < valid; ++i) {
    V3 ry1 = y1 * Ls[i][0], ry2 = y2 * Ls[i][1], ry3 = y3 * Ls[i][2];
    V3 yd1 = ry1 - ry2, yd2 = ry1 - ry3, yd1xd2 = cross(yd1, yd2);
    M3 Y{{yd1.x, yd2.x, yd1xd2.x, yd1.y, yd2.y, yd1xd2.y, yd1.z, yd2.z, yd1xd2.z}};
    Rs[i] = mul(Y, X);
    Ts[i] = ry1 - mul(Rs[i], x1);
  }
  return valid;
}

// ------------------------------------------------------------- cvl::Pose helpers
struct Pose { double q[4]; double t[3]; };   // q = (w,x,y,z), x' = R(q) x + t
static Pose pose_identity() { return {{1, 0, 0, 0}, {0, 0, 0}}; }

// utils/cvl/rotation_helpers.h:213-240 (q NOT renormalised)
static M3 rotmat_from_quat(const double q[4]) {
  double aa = q[0] * q[0], ab = q[0] * q[1], ac = q[0] * q[2], ad = q[0] * q[3];
  double bb = q[1] * q[1], bc = q[1] * q[2], bd = q[1] * q[3], cc = q[2] * q[2], cd = q[2] * q[3], dd = q[3] * q[3];
  return M3{{aa + bb - cc - dd, 2.0 * (bc - ad), 2.0 * (ac + bd),
             2.0 * (ad + bc), aa - bb + cc - dd, 2.0 * (cd - ab),
             2.0 * (bd - ac), 2.0 * (ab + cd), aa - bb - cc + dd}};
}
// utils/cvl/rotation_helpers.h:253-314
static void quat_from_rotmat(const M3& R, double q[4]) {
  double S, tr = R(0, 0) + R(1, 1) + R(2, 2) + 1.0;
  if (tr > 1e-7) {
    S = 0.5 / std::sqrt(tr);
    q[0] = 0.25 / S; q[1] = (R(2, 1) - R(1, 2)) * S; q[2] = (R(0, 2) - R(2, 0)) * S; q[3] = (R(1, 0) - R(0, 1)) * S;
  } else if (R(0, 0) > R(1, 1) && R(0, 0) > R(2, 2)) {
    S = std::sqrt(1.0 + R(0, 0) - R(1, 1) - R(2, 2)) * 2.0;
    q[0] = (R(2, 1) - R(1, 2)) / S; q[1] = 0.25 * S; q[2] = (R(1, 0) + R(0, 1)) / S; q[3] = (R(0, 2) + R(2, 0)) / S;
  } else if (R(1, 1) > R(2, 2)) {
    S = std::sqrt(1.0 + R(1, 1) - R(0, 0) - R(2, 2)) * 2.0;
    q[0] = (R(0, 2) - R(2, 0)) / S; q[1] = (R(1, 0) + R(0, 1)) / S; q[2] = 0.25 * S; q[3] = (R(2, 1) + R(1, 2)) / S;
  } else {
    S = std::sqrt(1.0 + R(2, 2) - R(0, 0) - R(1, 1)) * 2.0;
    q[0] = (R(1, 0) - R(0, 1)) / S; q[1] = (R(0, 2) + R(2, 0)) / S; q[2] = (R(2, 1) + R(1, 2)) / S; q[3] = 0.25 * S;
  }
}
// utils/cvl/pose.h:381-386 (one-sided norm test)
static bool pose_isnormal(const Pose& P) {
  for (double v : P.q) if (!std::isfinite(v)) return false;
  for (double v : P.t) if (!std::isfinite(v)) return false;
  double len = std::sqrt(P.q[0] * P.q[0] + P.q[1] * P.q[1] + P.q[2] * P.q[2] + P.q[3] * P.q[3]);
  return !(len - 1.0 > 1e-5);
}
static inline V3 pose_apply(const Pose& P, V3 x) { M3 R = rotmat_from_quat(P.q); V3 r = mul(R, x); return {r.x + P.t[0], r.y + P.t[1], r.z + P.t[2]}; }

// p4p.cpp:11-60
static Pose p4p(const double* xs, const double* ys, const int idx[4]) {
  auto X = [&](int i) { return V3{xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]}; };
  auto Yh = [&](int i) { return V3{ys[2 * i], ys[2 * i + 1], 1.0}; };
  M3 Rs[4]; V3 Ts[4];
  int valid = p3p_lambdatwist(Yh(idx[0]), Yh(idx[1]), Yh(idx[2]), X(idx[0]), X(idx[1]), X(idx[2]), Rs, Ts);
  double y0 = ys[2 * idx[3]], y1 = ys[2 * idx[3] + 1];
  V3 x = X(idx[3]);
  Pose P = pose_identity();
  double e0 = std::numeric_limits<double>::max();
  for (int v = 0; v < valid; ++v) {
    Pose tmp;
    quat_from_rotmat(Rs[v], tmp.q);
    double n = std::sqrt(tmp.q[0] * tmp.q[0] + tmp.q[1] * tmp.q[1] + tmp.q[2] * tmp.q[2] + tmp.q[3] * tmp.q[3]);
    for (double& c : tmp.q) c /= n;
    tmp.t[0] = Ts[v].x; tmp.t[1] = Ts[v].y; tmp.t[2] = Ts[v].z;
    if (!pose_isnormal(tmp)) continue;
    V3 xr = pose_apply(tmp, x);
    if (xr.z < 0) continue;
    double dx = xr.x / xr.z - y0, dy = xr.y / xr.z - y1;
    double e = dx * dx + dy * dy;
    if (std::isnan(e)) continue;
    if (e < e0) { P = tmp; e0 = e; }
  }
  return P;
}

// parameters.h:76-102
static int get_iterations(double estimated_inliers) {
  const double p_meets = 0.9, min_probability = 0.99999;
  const int max_iterations = 1000, min_iterations = 100;
  double p_inlier = std::min(0.9, estimated_inliers * p_meets);
  p_inlier = std::min(std::max(p_inlier, 1e-2), 1 - 1e-8);
  if (p_inlier < 0.01) return max_iterations;
  double p_failure = std::min(std::max(1.0 - min_probability, 1e-8), 0.01);
  double p_good = std::pow(p_inlier, 4);
  double iterations = std::ceil(std::log(p_failure) / std::log(1.0 - p_good)) + 50;
  if (iterations < min_iterations) return min_iterations;
  if (iterations > max_iterations) return max_iterations;
  return (int)iterations;
}

// Counter-based replacement for get4RandomInRange0 (pnp_ransac.cpp:161-183: 4 distinct
// uniform indices, returned sorted ascending because they pass through a std::set).
static inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void sample4(uint64_t seed, uint64_t obj_key, uint32_t iter, int n, int idx[4]) {
  uint64_t base = splitmix64(seed ^ splitmix64(obj_key * 0xD1342543DE82EF95ull + iter));
  int cnt = 0;
  for (uint32_t j = 0; cnt < 4; ++j) {
    int v = (int)(splitmix64(base + j) % (uint64_t)n);
    bool dup = false;
    for (int k = 0; k < cnt; ++k) dup |= (idx[k] == v);
    if (!dup) idx[cnt++] = v;
  }
  std::sort(idx, idx + 4);
}

// pnp_ransac.cpp:41-87 without the early break (which only under-counts
// hypotheses that cannot beat best_inliers, so the argmax is unchanged)
static int count_inliers(const double* xs, const double* ys, int n, double thr, const Pose& P) {
  M3 R = rotmat_from_quat(P.q);
  double thr2 = thr * thr;
  int inl = 0;
  for (int i = 0; i < n; ++i) {
    V3 X{xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]};
    V3 r = mul(R, X);
    double x = r.x + P.t[0], y = r.y + P.t[1], z = r.z + P.t[2];
    double iz = 1.0 / z;
    if (iz < 0) continue;
    double e1 = x * iz - ys[2 * i], e2 = y * iz - ys[2 * i + 1];
    inl += (e1 * e1 + e2 * e2 < thr2) ? 1 : 0;
  }
  return inl;
}

// ------------------------------------------------ dense SPD solve (Cholesky)
static bool chol_solve(int n, std::vector<double>& A, std::vector<double>& b) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[i * n + k] * b[k]; b[i] = s / A[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * b[k]; b[i] = s / A[i * n + i]; }
  return true;
}

// ------------------------------------------------ Ceres refine (pnp_ransac.cpp:240-326)
// Residual (pnp_ransac.cpp:120-139): r = pi(R(q) X + t) - y with the *unnormalised* q,
// parameter blocks q[4] (QuaternionParameterization, local size 3) and t[3].
static void pnp_residuals(const Pose& P, const std::vector<double>& xs, const std::vector<double>& ys, std::vector<double>& r) {
  M3 R = rotmat_from_quat(P.q);
  int n = (int)xs.size() / 3;
  r.resize(2 * n);
  for (int i = 0; i < n; ++i) {
    V3 p = mul(R, V3{xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]});
    double x = p.x + P.t[0], y = p.y + P.t[1], z = p.z + P.t[2];
    double iz = 1.0 / z;
    r[2 * i] = x * iz - ys[2 * i];
    r[2 * i + 1] = y * iz - ys[2 * i + 1];
  }
}
// Local (6-column) Jacobian: d r / d(delta_q[3], t[3]) = [J_q(2x4) * PlusJac(4x3) | J_t]
static void pnp_jacobian(const Pose& P, const std::vector<double>& xs, std::vector<double>& J) {
  const double a = P.q[0], b = P.q[1], c = P.q[2], d = P.q[3];
  M3 R = rotmat_from_quat(P.q);
  int n = (int)xs.size() / 3;
  J.assign((size_t)2 * n * 6, 0.0);
  // QuaternionParameterization::ComputeJacobian (public Ceres): rows w,x,y,z
  const double PJ[4][3] = {{-b, -c, -d}, {a, d, -c}, {-d, a, b}, {c, -b, a}};
  for (int i = 0; i < n; ++i) {
    double X = xs[3 * i], Y = xs[3 * i + 1], Z = xs[3 * i + 2];
    V3 p = mul(R, V3{X, Y, Z});
    double x = p.x + P.t[0], y = p.y + P.t[1], z = p.z + P.t[2];
    double iz = 1.0 / z;
    // d(RX)/dq, columns a,b,c,d (derivative of the polynomial rotation matrix above)
    double dP[3][4] = {
        {2 * (a * X - d * Y + c * Z), 2 * (b * X + c * Y + d * Z), 2 * (-c * X + b * Y + a * Z), 2 * (-d * X - a * Y + b * Z)},
        {2 * (d * X + a * Y - b * Z), 2 * (c * X - b * Y - a * Z), 2 * (b * X + c * Y + d * Z), 2 * (a * X - d * Y + c * Z)},
        {2 * (-c * X + b * Y + a * Z), 2 * (d * X + a * Y - b * Z), 2 * (-a * X + d * Y - c * Z), 2 * (b * X + c * Y + d * Z)}};
    double du[3] = {iz, 0, -x * iz * iz}, dv[3] = {0, iz, -y * iz * iz};
    double Jq[2][4];
    for (int k = 0; k < 4; ++k) {
      Jq[0][k] = du[0] * dP[0][k] + du[2] * dP[2][k];
      Jq[1][k] = dv[1] * dP[1][k] + dv[2] * dP[2][k];
    }
    for (int rr = 0; rr < 2; ++rr) {
      double* row = &J[(size_t)(2 * i + rr) * 6];
      for (int k = 0; k < 3; ++k) row[k] = Jq[rr][0] * PJ[0][k] + Jq[rr][1] * PJ[1][k] + Jq[rr][2] * PJ[2][k] + Jq[rr][3] * PJ[3][k];
      const double* dd = rr == 0 ? du : dv;
      row[3] = dd[0]; row[4] = dd[1]; row[5] = dd[2];
    }
  }
}
// QuaternionParameterization::Plus then t += dt
static Pose pnp_plus(const Pose& P, const double delta[6]) {
  Pose out = P;
  double nd = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (nd > 0.0) {
    double s = std::sin(nd) / nd;
    double qd[4] = {std::cos(nd), s * delta[0], s * delta[1], s * delta[2]};
    const double* x = P.q;
    out.q[0] = qd[0] * x[0] - qd[1] * x[1] - qd[2] * x[2] - qd[3] * x[3];
    out.q[1] = qd[0] * x[1] + qd[1] * x[0] + qd[2] * x[3] - qd[3] * x[2];
    out.q[2] = qd[0] * x[2] - qd[1] * x[3] + qd[2] * x[0] + qd[3] * x[1];
    out.q[3] = qd[0] * x[3] + qd[1] * x[2] - qd[2] * x[1] + qd[3] * x[0];
  }
  for (int k = 0; k < 3; ++k) out.t[k] = P.t[k] + delta[3 + k];
  return out;
}

struct CeresStats { int iterations; int successful; int termination; double final_cost; };
// Trust-region LM, Ceres defaults: radius 1e4, min_relative_decrease 1e-3, lm diag clamp
// [1e-6,1e32], jacobi scaling on, parameter_tolerance 1e-8, monotonic steps.
static CeresStats ceres_lm(Pose& P, const std::vector<double>& xs, const std::vector<double>& ys,
                           int max_iter, double ftol, double gtol) {
  CeresStats st{0, 0, 0, 0};
  const int m = (int)ys.size();
  std::vector<double> r, J, rc, Js((size_t)m * 6);
  pnp_residuals(P, xs, ys, r);
  pnp_jacobian(P, xs, J);
  double cost = 0; for (double v : r) cost += v * v; cost *= 0.5;
  double scale[6];
  for (int j = 0; j < 6; ++j) { double s = 0; for (int i = 0; i < m; ++i) s += J[(size_t)i * 6 + j] * J[(size_t)i * 6 + j]; scale[j] = 1.0 / (1.0 + std::sqrt(s)); }
  auto grad_max_norm = [&](const Pose& Q, const std::vector<double>& Jm, const std::vector<double>& rm) {
    double g[6];
    for (int j = 0; j < 6; ++j) { double s = 0; for (int i = 0; i < m; ++i) s += Jm[(size_t)i * 6 + j] * rm[i]; g[j] = -s; }
    Pose Qp = pnp_plus(Q, g);
    double mx = 0;
    for (int k = 0; k < 4; ++k) mx = std::max(mx, std::fabs(Q.q[k] - Qp.q[k]));
    for (int k = 0; k < 3; ++k) mx = std::max(mx, std::fabs(Q.t[k] - Qp.t[k]));
    return mx;
  };
  auto xnorm = [](const Pose& Q) { double s = 0; for (double v : Q.q) s += v * v; for (double v : Q.t) s += v * v; return std::sqrt(s); };
  double gmax = grad_max_norm(P, J, r);
  double radius = 1e4, decrease_factor = 2.0;
  double diag[6];
  bool reuse_diagonal = false;
  int invalid_steps = 0;
  int iter = 0;
  while (true) {
    if (iter >= max_iter) { st.termination = 1; break; }          // NO_CONVERGENCE (max iterations)
    if (gmax <= gtol) { st.termination = 2; break; }              // gradient tolerance
    if (radius < 1e-32) { st.termination = 3; break; }
    ++iter;
    for (int i = 0; i < m; ++i) for (int j = 0; j < 6; ++j) Js[(size_t)i * 6 + j] = J[(size_t)i * 6 + j] * scale[j];
    if (!reuse_diagonal) {
      for (int j = 0; j < 6; ++j) { double s = 0; for (int i = 0; i < m; ++i) s += Js[(size_t)i * 6 + j] * Js[(size_t)i * 6 + j]; diag[j] = std::min(std::max(s, 1e-6), 1e32); }
    }
    std::vector<double> H(36, 0.0), g(6, 0.0);
    for (int i = 0; i < m; ++i) for (int a = 0; a < 6; ++a) {
      g[a] += Js[(size_t)i * 6 + a] * r[i];
      for (int b = 0; b < 6; ++b) H[a * 6 + b] += Js[(size_t)i * 6 + a] * Js[(size_t)i * 6 + b];
    }
    for (int j = 0; j < 6; ++j) H[j * 6 + j] += diag[j] / radius;
    reuse_diagonal = true;
    bool ok = chol_solve(6, H, g);
    double step[6];
    for (int j = 0; j < 6; ++j) step[j] = -g[j];
    double model_change = 0;
    if (ok) {
      for (int i = 0; i < m; ++i) { double mr = 0; for (int j = 0; j < 6; ++j) mr += Js[(size_t)i * 6 + j] * step[j]; model_change -= mr * (r[i] + mr / 2.0); }
    }
    if (!ok || !(model_change > 0.0)) {                           // invalid step
      if (++invalid_steps >= 5) { st.termination = 4; break; }
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      continue;
    }
    invalid_steps = 0;
    double delta[6];
    for (int j = 0; j < 6; ++j) delta[j] = step[j] * scale[j];
    Pose cand = pnp_plus(P, delta);
    pnp_residuals(cand, xs, ys, rc);
    double ccost = 0; for (double v : rc) ccost += v * v; ccost *= 0.5;
    if (!std::isfinite(ccost)) ccost = std::numeric_limits<double>::max();
    // parameter tolerance
    double sn = 0; for (int k = 0; k < 4; ++k) sn += (P.q[k] - cand.q[k]) * (P.q[k] - cand.q[k]);
    for (int k = 0; k < 3; ++k) sn += (P.t[k] - cand.t[k]) * (P.t[k] - cand.t[k]);
    if (std::sqrt(sn) <= 1e-8 * (xnorm(P) + 1e-8)) { st.termination = 5; break; }
    // function tolerance (checked before the step is adopted, as in Ceres >= 1.12)
    if (std::fabs(cost - ccost) <= ftol * cost) { st.termination = 6; break; }
    double rel = (cost - ccost) / model_change;
    if (rel > 1e-3) {
      P = cand; r = rc; cost = ccost;
      pnp_jacobian(P, xs, J);
      gmax = grad_max_norm(P, J, r);
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(1e16, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      ++st.successful;
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
  }
  st.iterations = iter; st.final_cost = cost;
  return st;
}

// pnp_ransac.cpp:240-326
static void pnp_refine(Pose& best, const double* xs, const double* ys, int n, double threshold, int* refine_iters) {
  std::vector<double> ix, iy;
  std::vector<int> inl(n, 0);
  double thr = threshold * threshold;
  for (int i = 0; i < n; ++i) {
    V3 xr = pose_apply(best, V3{xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]});
    if (xr.z < 0) continue;
    double dx = xr.x / xr.z - ys[2 * i], dy = xr.y / xr.z - ys[2 * i + 1];
    if (dx * dx + dy * dy > thr) continue;
    ix.insert(ix.end(), {xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]});
    iy.insert(iy.end(), {ys[2 * i], ys[2 * i + 1]});
    inl[i] = 1;
  }
  CeresStats s1 = ceres_lm(best, ix, iy, 5, 1e-6, 1e-6);
  if (refine_iters) refine_iters[0] = s1.iterations;
  ix.clear(); iy.clear();
  int deltas = 0;
  for (int i = 0; i < n; ++i) {
    V3 xr = pose_apply(best, V3{xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]});
    bool inlier = true;
    if (xr.z < 0) inlier = false;
    double dx = xr.x / xr.z - ys[2 * i], dy = xr.y / xr.z - ys[2 * i + 1];
    if (dx * dx + dy * dy > thr) inlier = false;
    if (inlier ^ (inl[i] == 1)) deltas++;
    if (!inlier) continue;
    ix.insert(ix.end(), {xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]});
    iy.insert(iy.end(), {ys[2 * i], ys[2 * i + 1]});
  }
  if (refine_iters) refine_iters[1] = -1;
  if (deltas < 0.05 * (double)(ix.size() / 3)) return;
  CeresStats s2 = ceres_lm(best, ix, iy, 3, 1e-8, 1e-8);
  if (refine_iters) refine_iters[1] = s2.iterations;
}

static void pose_to_4x4(const Pose& P, double* T) {
  M3 R = rotmat_from_quat(P.q);
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T[4 * r + c] = R(r, c); T[4 * r + 3] = P.t[r]; }
  T[12] = T[13] = T[14] = 0; T[15] = 1;
}

// ======================================================================== g2o
// SE3Quat (types/slam3d/se3quat.h): unit quaternion (x,y,z,w) + t
struct SE3 { double qx, qy, qz, qw; double t[3]; };

// Eigen::Quaterniond(Matrix3d) (external Eigen, public algorithm) + normalizeRotation se3quat.h:277-282
static SE3 se3_from_Rt(const M3& m, const double t[3]) {
  double q[4];  // x y z w
  double tr = m(0, 0) + m(1, 1) + m(2, 2);
  if (tr > 0) {
    double s = std::sqrt(tr + 1.0);
    q[3] = 0.5 * s; s = 0.5 / s;
    q[0] = (m(2, 1) - m(1, 2)) * s; q[1] = (m(0, 2) - m(2, 0)) * s; q[2] = (m(1, 0) - m(0, 1)) * s;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * s; s = 0.5 / s;
    q[3] = (m(k, j) - m(j, k)) * s; q[j] = (m(j, i) + m(i, j)) * s; q[k] = (m(k, i) + m(i, k)) * s;
  }
  if (q[3] < 0) for (double& v : q) v = -v;
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  return SE3{q[0] / n, q[1] / n, q[2] / n, q[3] / n, {t[0], t[1], t[2]}};
}
static M3 se3_R(const SE3& T) {  // Eigen toRotationMatrix
  double tx = 2 * T.qx, ty = 2 * T.qy, tz = 2 * T.qz;
  double twx = tx * T.qw, twy = ty * T.qw, twz = tz * T.qw;
  double txx = tx * T.qx, txy = ty * T.qx, txz = tz * T.qx, tyy = ty * T.qy, tyz = tz * T.qy, tzz = tz * T.qz;
  return M3{{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)}};
}
static inline V3 se3_map(const SE3& T, V3 p) { V3 r = mul(se3_R(T), p); return {r.x + T.t[0], r.y + T.t[1], r.z + T.t[2]}; }

// SE3Quat::exp (se3quat.h:220-254) then exp(update) * T (types_six_dof_expmap.h:100-103)
static SE3 se3_oplus(const SE3& T, const double u[6]) {
  V3 om{u[0], u[1], u[2]}, up{u[3], u[4], u[5]};
  double theta = std::sqrt(dot(om, om));
  M3 Om{{0, -om.z, om.y, om.z, 0, -om.x, -om.y, om.x, 0}};
  M3 Om2 = mul(Om, Om);
  M3 R, V;
  if (theta < 0.00001) {
    for (int i = 0; i < 9; ++i) R.m[i] = (i % 4 == 0 ? 1.0 : 0.0) + Om.m[i] + Om2.m[i];
    V = R;
  } else {
    double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta), c = (theta - std::sin(theta)) / (theta * theta * theta);
    for (int i = 0; i < 9; ++i) { double I = (i % 4 == 0 ? 1.0 : 0.0); R.m[i] = I + a * Om.m[i] + b * Om2.m[i]; V.m[i] = I + b * Om.m[i] + c * Om2.m[i]; }
  }
  V3 tv = mul(V, up);
  double td[3] = {tv.x, tv.y, tv.z};
  SE3 E = se3_from_Rt(R, td);
  // SE3Quat operator* (se3quat.h:103-109): t = t1 + r1*t2 ; r = r1*r2 ; normalizeRotation
  V3 rt = mul(se3_R(E), V3{T.t[0], T.t[1], T.t[2]});
  SE3 out;
  out.t[0] = E.t[0] + rt.x; out.t[1] = E.t[1] + rt.y; out.t[2] = E.t[2] + rt.z;
  out.qw = E.qw * T.qw - E.qx * T.qx - E.qy * T.qy - E.qz * T.qz;
  out.qx = E.qw * T.qx + E.qx * T.qw + E.qy * T.qz - E.qz * T.qy;
  out.qy = E.qw * T.qy - E.qx * T.qz + E.qy * T.qw + E.qz * T.qx;
  out.qz = E.qw * T.qz + E.qx * T.qy - E.qy * T.qx + E.qz * T.qw;
  if (out.qw < 0) { out.qx = -out.qx; out.qy = -out.qy; out.qz = -out.qz; out.qw = -out.qw; }
  double n = std::sqrt(out.qx * out.qx + out.qy * out.qy + out.qz * out.qz + out.qw * out.qw);
  out.qx /= n; out.qy /= n; out.qz /= n; out.qw /= n;
  return out;
}

struct Graph {
  int n_vert;
  std::vector<SE3> est;
  std::vector<uint8_t> fixed;
  int n_edges;
  const int* e_obj; const int* e_cam;
  const double* cam_k; const double* p; const double* uv; const double* info;
  std::vector<uint8_t> level;      // 0 = in the optimisation, 1 = outlier
  std::vector<uint8_t> robust;     // Huber kernel attached
  std::vector<double> err;         // _error of every edge (2 per edge), as last computed
  double huber_delta;
};

// EdgeSE3ProjectFrom{,Fixed}Object::computeError, types_object_slam.cpp:45-60,156-169
static void edge_error(Graph& G, int e) {
  V3 p{G.p[3 * e], G.p[3 * e + 1], G.p[3 * e + 2]};
  if (G.e_obj[e] >= 0) p = se3_map(G.est[G.e_obj[e]], p);
  V3 pc = se3_map(G.est[G.e_cam[e]], p);
  const double* k = &G.cam_k[4 * e];
  G.err[2 * e] = G.uv[2 * e] - (k[0] * pc.x / pc.z + k[2]);
  G.err[2 * e + 1] = G.uv[2 * e + 1] - (k[1] * pc.y / pc.z + k[3]);
}
// BaseEdge::chi2, base_edge.h:56-59
static double edge_chi2(const Graph& G, int e) {
  const double* O = &G.info[4 * e];
  double e0 = G.err[2 * e], e1 = G.err[2 * e + 1];
  return e0 * (O[0] * e0 + O[1] * e1) + e1 * (O[2] * e0 + O[3] * e1);
}
// RobustKernelHuber::robustify, robust_kernel_impl.cpp:65-78
static void huber(double e, double delta, double rho[3]) {
  double dsqr = delta * delta;
  if (e <= dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
  else { double sq = std::sqrt(e); rho[0] = 2 * sq * delta - dsqr; rho[1] = delta / sq; rho[2] = -0.5 * rho[1] / e; }
}
// linearizeOplus, types_object_slam.cpp:70-123,177-201.  Ji: 2x6 wrt object, Jj: 2x6 wrt camera.
static void edge_jacobians(const Graph& G, int e, double Ji[12], double Jj[12]) {
  V3 p{G.p[3 * e], G.p[3 * e + 1], G.p[3 * e + 2]};
  V3 pw = p;
  if (G.e_obj[e] >= 0) pw = se3_map(G.est[G.e_obj[e]], p);
  const SE3& Tcw = G.est[G.e_cam[e]];
  M3 Rcw = se3_R(Tcw);
  V3 pc = se3_map(Tcw, pw);
  const double* k = &G.cam_k[4 * e];
  double pj[6] = {-(k[0] / pc.z), 0.0, k[0] * pc.x / (pc.z * pc.z), 0.0, -(k[1] / pc.z), k[1] * pc.y / (pc.z * pc.z)};
  auto se3deriv = [](V3 v, double D[18]) {
    double d[18] = {0, v.z, -v.y, 1, 0, 0, -v.z, 0, v.x, 0, 1, 0, v.y, -v.x, 0, 0, 0, 1};
    std::memcpy(D, d, sizeof(d));
  };
  double D[18];
  se3deriv(pc, D);
  for (int r = 0; r < 2; ++r) for (int c = 0; c < 6; ++c) Jj[6 * r + c] = pj[3 * r] * D[c] + pj[3 * r + 1] * D[6 + c] + pj[3 * r + 2] * D[12 + c];
  if (G.e_obj[e] >= 0) {
    double pjR[6];
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) pjR[3 * r + c] = pj[3 * r] * Rcw(0, c) + pj[3 * r + 1] * Rcw(1, c) + pj[3 * r + 2] * Rcw(2, c);
    se3deriv(pw, D);
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 6; ++c) Ji[6 * r + c] = pjR[3 * r] * D[c] + pjR[3 * r + 1] * D[6 + c] + pjR[3 * r + 2] * D[12 + c];
  }
}

struct LMState { double lambda; double ni; };

// One SparseOptimizer::optimize(iterations) call on the level-0 edges
// (sparse_optimizer.cpp:366-431 driving optimization_algorithm_levenberg.cpp:58-150).
// Returns the number of outer iterations performed; -1 if there is nothing to optimise.
static int g2o_optimize(Graph& G, int iterations, int* lm_trials_total) {
  // initializeOptimization(0): active edges = level 0; active vertices = non-fixed vertices touched by them
  std::vector<int> active;
  // edges whose vertices are all fixed never enter the active set (SparseOptimizer::initializeOptimization,
  // sparse_optimizer.cpp:201-272 skips e->allVerticesFixed())
  for (int e = 0; e < G.n_edges; ++e)
    if (G.level[e] == 0 && (!G.fixed[G.e_cam[e]] || (G.e_obj[e] >= 0 && !G.fixed[G.e_obj[e]]))) active.push_back(e);
  std::vector<int> col(G.n_vert, -1);
  int nfree = 0;
  {
    std::vector<uint8_t> touched(G.n_vert, 0);
    for (int e : active) { if (G.e_obj[e] >= 0) touched[G.e_obj[e]] = 1; touched[G.e_cam[e]] = 1; }
    for (int v = 0; v < G.n_vert; ++v) if (touched[v] && !G.fixed[v]) col[v] = nfree++;   // ascending vertex id (sparse_optimizer.cpp:493-498)
  }
  if (nfree == 0 || active.empty()) return -1;
  const int N = 6 * nfree;
  bool coupled = false;
  for (int e : active) if (G.e_obj[e] >= 0 && col[G.e_obj[e]] >= 0 && col[G.e_cam[e]] >= 0) coupled = true;

  auto compute_active_errors = [&]() { for (int e : active) edge_error(G, e); };
  auto active_robust_chi2 = [&]() {  // sparse_optimizer.cpp:102-116
    double chi = 0, rho[3];
    for (int e : active) { double c = edge_chi2(G, e); if (G.robust[e]) { huber(c, G.huber_delta, rho); chi += rho[0]; } else chi += c; }
    return chi;
  };

  LMState lm{-1.0, 2.0};
  std::vector<double> H, b(N), x(N), Hd;       // dense H only when camera and objects are both free
  std::vector<double> Hblk((size_t)nfree * 36);
  int done = 0;
  bool ok = true;
  for (int it = 0; it < iterations && ok; ++it) {
    compute_active_errors();
    double currentChi = active_robust_chi2();
    double tempChi = currentChi;
    // buildSystem (block_solver.hpp:464-523) via constructQuadraticForm (base_binary_edge.hpp:64-127, base_unary_edge.hpp:52-78)
    std::fill(b.begin(), b.end(), 0.0);
    std::fill(Hblk.begin(), Hblk.end(), 0.0);
    if (coupled) { H.assign((size_t)N * N, 0.0); }
    for (int e : active) {
      double Ji[12], Jj[12];
      edge_jacobians(G, e, Ji, Jj);
      const double* O = &G.info[4 * e];
      double e0 = G.err[2 * e], e1 = G.err[2 * e + 1];
      double w = 1.0;
      if (G.robust[e]) { double rho[3]; huber(edge_chi2(G, e), G.huber_delta, rho); w = rho[1]; }
      double Or[2] = {-(O[0] * e0 + O[1] * e1) * w, -(O[2] * e0 + O[3] * e1) * w};   // omega_r (*= rho[1])
      double Ow[4] = {O[0] * w, O[1] * w, O[2] * w, O[3] * w};                          // robustInformation, base_edge.h:94-100
      int ci = (G.e_obj[e] >= 0) ? col[G.e_obj[e]] : -1;
      int cj = col[G.e_cam[e]];
      auto add_block = [&](const double* A, const double* B, int ca, int cb) {
        // += A^T Ow B
        double AtO[12];
        for (int c = 0; c < 6; ++c) { AtO[c] = A[c] * Ow[0] + A[6 + c] * Ow[2]; AtO[6 + c] = A[c] * Ow[1] + A[6 + c] * Ow[3]; }
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) {
          double v = AtO[r] * B[c] + AtO[6 + r] * B[6 + c];
          if (ca == cb) Hblk[(size_t)ca * 36 + r * 6 + c] += v;
          if (coupled) { H[(size_t)(6 * ca + r) * N + 6 * cb + c] += v; if (ca != cb) H[(size_t)(6 * cb + c) * N + 6 * ca + r] += v; }
        }
      };
      if (ci >= 0) {
        for (int c = 0; c < 6; ++c) b[6 * ci + c] += Ji[c] * Or[0] + Ji[6 + c] * Or[1];
        add_block(Ji, Ji, ci, ci);
        if (cj >= 0) add_block(Ji, Jj, ci, cj);
      }
      if (cj >= 0) {
        for (int c = 0; c < 6; ++c) b[6 * cj + c] += Jj[c] * Or[0] + Jj[6 + c] * Or[1];
        add_block(Jj, Jj, cj, cj);
      }
    }
    if (it == 0) {  // computeLambdaInit, optimization_algorithm_levenberg.cpp:152-166
      double maxDiag = 0;
      for (int v = 0; v < nfree; ++v) for (int j = 0; j < 6; ++j) maxDiag = std::max(std::fabs(Hblk[(size_t)v * 36 + j * 7]), maxDiag);
      lm.lambda = 1e-5 * maxDiag;
      lm.ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    std::vector<SE3> backup;
    do {
      backup = G.est;                                   // push()
      bool ok2 = true;
      if (coupled) {
        Hd = H;
        for (int j = 0; j < N; ++j) Hd[(size_t)j * N + j] += lm.lambda;
        x = b;
        ok2 = chol_solve(N, Hd, x);
      } else {
        for (int v = 0; v < nfree; ++v) {
          std::vector<double> A(Hblk.begin() + (size_t)v * 36, Hblk.begin() + (size_t)(v + 1) * 36), rhs(b.begin() + 6 * v, b.begin() + 6 * v + 6);
          for (int j = 0; j < 6; ++j) A[j * 7] += lm.lambda;
          bool okv = chol_solve(6, A, rhs);
          ok2 = ok2 && okv;
          for (int j = 0; j < 6; ++j) x[6 * v + j] = okv ? rhs[j] : 0.0;
        }
      }
      for (int v = 0; v < G.n_vert; ++v) if (col[v] >= 0) G.est[v] = se3_oplus(G.est[v], &x[6 * col[v]]);   // update()
      compute_active_errors();
      tempChi = active_robust_chi2();
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = currentChi - tempChi;
      double scale = 0;                                  // computeScale :168-175
      for (int j = 0; j < N; ++j) scale += x[j] * (lm.lambda * x[j] + b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow(2 * rho - 1, 3);
        alpha = std::min(alpha, 2. / 3.);
        double scaleFactor = std::max(1. / 3., alpha);
        lm.lambda *= scaleFactor;
        lm.ni = 2;
        currentChi = tempChi;
      } else {
        lm.lambda *= lm.ni;
        lm.ni *= 2;
        G.est = backup;                                  // pop(); NOTE edge errors stay those of the rejected trial
        if (!std::isfinite(lm.lambda)) break;
      }
      qmax++;
      if (lm_trials_total) ++*lm_trials_total;
    } while (rho < 0 && qmax < 10);
    ++done;
    if (qmax == 10 || rho == 0 || !std::isfinite(lm.lambda)) ok = false;   // Terminate
  }
  return done;
}

}  // namespace

extern "C" {

// P3P alone: returns #solutions, Rs (row-major 4x9), Ts (4x3).  y* are 2-D pinhole-normalised.
int orc_p3p(const double* y2d /*3x2*/, const double* x3d /*3x3*/, double* Rs, double* Ts) {
  M3 R[4]; V3 T[4];
  int v = p3p_lambdatwist(V3{y2d[0], y2d[1], 1}, V3{y2d[2], y2d[3], 1}, V3{y2d[4], y2d[5], 1},
                          V3{x3d[0], x3d[1], x3d[2]}, V3{x3d[3], x3d[4], x3d[5]}, V3{x3d[6], x3d[7], x3d[8]}, R, T);
  for (int i = 0; i < v; ++i) { std::memcpy(Rs + 9 * i, R[i].m, 72); Ts[3 * i] = T[i].x; Ts[3 * i + 1] = T[i].y; Ts[3 * i + 2] = T[i].z; }
  return v;
}

// P4P: pose as 4x4 row-major; identity on degeneracy (p4p.cpp:35,60)
void orc_p4p(const double* xs, const double* ys, const int* idx4, double* T16) {
  Pose P = p4p(xs, ys, idx4);
  pose_to_4x4(P, T16);
}

void orc_sample4(uint64_t seed, uint64_t obj_key, uint32_t iter, int n, int* idx4) { sample4(seed, obj_key, iter, n, idx4); }
int orc_get_iterations(double inlier_ratio) { return get_iterations(inlier_ratio); }

// lambdatwist.pnp (pnp_python_binding.cpp:32-54 -> PNP::compute pnp_ransac.cpp:188-232).
// T16: 4x4 row-major (identity == failure).  stats: [best_inliers, best_iter, total_iters, refine1_iters, refine2_iters]
void orc_pnp(const double* xs, const double* ys, int n, double threshold, uint64_t seed, uint64_t obj_key,
             int do_refine, double* T16, int* stats) {
  Pose best = pose_identity();
  int best_inl = 0, best_it = -1;
  int iters = get_iterations(0.0);
  int i = 0;
  if (n >= 4) {
    for (i = 0; i < iters; ++i) {
      int idx[4];
      sample4(seed, obj_key, (uint32_t)i, n, idx);
      Pose P = p4p(xs, ys, idx);
      if (!pose_isnormal(P)) continue;
      int inl = count_inliers(xs, ys, n, threshold, P);
      if (inl > best_inl) {
        best_inl = inl; best = P; best_it = i;
        iters = get_iterations(best_inl / (double)n);
      }
    }
  }
  int rit[2] = {-1, -1};
  if (best_inl > 3 && do_refine) pnp_refine(best, xs, ys, n, threshold, rit);
  pose_to_4x4(best, T16);
  if (stats) { stats[0] = best_inl; stats[1] = best_it; stats[2] = i; stats[3] = rit[0]; stats[4] = rit[1]; }
}

// Refine alone from a given 4x4 pose (for unit tests of the Ceres restatement).
void orc_pnp_refine(const double* xs, const double* ys, int n, double threshold, double* T16_inout, int* iters2) {
  M3 R; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = T16_inout[4 * r + c];
  Pose P; quat_from_rotmat(R, P.q);
  double nq = std::sqrt(P.q[0] * P.q[0] + P.q[1] * P.q[1] + P.q[2] * P.q[2] + P.q[3] * P.q[3]);
  for (double& v : P.q) v /= nq;
  P.t[0] = T16_inout[3]; P.t[1] = T16_inout[7]; P.t[2] = T16_inout[11];
  pnp_refine(P, xs, ys, n, threshold, iters2);
  pose_to_4x4(P, T16_inout);
}

// ObjectSLAM.optimize() core (lib/object_slam.py:842-896) over a packed graph.
//  poses      n_vert x 12  row-major [R|t] (3x4), in/out
//  fixed      n_vert
//  e_obj      object vertex of each edge, or -1 for EdgeSE3ProjectFromFixedObject (p is then p_inG)
//  e_cam      camera vertex of each edge
//  inliers    n_edges in/out (detections[...]["inliers"])
//  its        iterations per round (n_rounds entries)
//  stats      [rounds_run, total_outer_iterations, total_lm_trials]
//  err_out    (orc_ba_optimize_err) 2 * n_edges: the error vector each edge holds when optimize() returns — for an edge of the active set
//             after a rejected LM trial that is the REJECTED state's error (g2o computes the trial's errors into the edges and pops only
//             the vertices, optimization_algorithm_levenberg.cpp:118-146), which ObjectSLAM.optimize() then reads through e.chi2()
//             without recomputing (lib/object_slam.py:881-883)
void orc_ba_optimize_err(int n_vert, double* poses, const uint8_t* fixed, int n_edges, const int* e_obj, const int* e_cam,
                         const double* cam_k, const double* p, const double* uv, const double* info, uint8_t* inliers,
                         const int* its, int n_rounds, double huber_delta, double chi2_gate, int init_with_outliers,
                         int* stats, double* err_out);
void orc_ba_optimize(int n_vert, double* poses, const uint8_t* fixed, int n_edges, const int* e_obj, const int* e_cam,
                     const double* cam_k, const double* p, const double* uv, const double* info, uint8_t* inliers,
                     const int* its, int n_rounds, double huber_delta, double chi2_gate, int init_with_outliers,
                     int* stats) {
  orc_ba_optimize_err(n_vert, poses, fixed, n_edges, e_obj, e_cam, cam_k, p, uv, info, inliers, its, n_rounds, huber_delta, chi2_gate,
                      init_with_outliers, stats, nullptr);
}
void orc_ba_optimize_err(int n_vert, double* poses, const uint8_t* fixed, int n_edges, const int* e_obj, const int* e_cam,
                         const double* cam_k, const double* p, const double* uv, const double* info, uint8_t* inliers,
                         const int* its, int n_rounds, double huber_delta, double chi2_gate, int init_with_outliers,
                         int* stats, double* err_out) {
  Graph G;
  G.n_vert = n_vert; G.n_edges = n_edges;
  G.e_obj = e_obj; G.e_cam = e_cam; G.cam_k = cam_k; G.p = p; G.uv = uv; G.info = info;
  G.huber_delta = huber_delta;
  G.est.resize(n_vert); G.fixed.assign(fixed, fixed + n_vert);
  for (int v = 0; v < n_vert; ++v) {
    M3 R; const double* T = poses + 12 * v;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = T[4 * r + c];
    double t[3] = {T[3], T[7], T[11]};
    G.est[v] = se3_from_Rt(R, t);
  }
  G.level.assign(n_edges, 0); G.robust.assign(n_edges, 1); G.err.assign((size_t)2 * n_edges, 0.0);
  int num_good = 0;
  if (init_with_outliers) {
    num_good = n_edges;      // object_slam.py:851-854
  } else {
    for (int e = 0; e < n_edges; ++e) {   // :856-866
      edge_error(G, e);
      if (edge_chi2(G, e) > chi2_gate) { G.level[e] = 1; inliers[e] = 0; } else { ++num_good; G.level[e] = 0; inliers[e] = 1; }
    }
  }
  int rounds = 0, outer = 0, trials = 0;
  for (int it = 0; it < n_rounds; ++it) {
    if (n_edges < 4 || num_good < 4) break;   // :869-871
    int d = g2o_optimize(G, its[it], &trials);
    if (d > 0) outer += d;
    ++rounds;
    num_good = 0;
    for (int e = 0; e < n_edges; ++e) {       // :878-896
      if (!inliers[e]) edge_error(G, e);
      if (edge_chi2(G, e) > chi2_gate) { G.level[e] = 1; inliers[e] = 0; } else { ++num_good; G.level[e] = 0; inliers[e] = 1; }
      if (it == std::max(1, n_rounds / 2)) G.robust[e] = 0;
    }
  }
  for (int v = 0; v < n_vert; ++v) {
    M3 R = se3_R(G.est[v]);
    double* T = poses + 12 * v;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T[4 * r + c] = R(r, c); T[4 * r + 3] = G.est[v].t[r]; }
  }
  if (stats) { stats[0] = rounds; stats[1] = outer; stats[2] = trials; }
  if (err_out) std::copy(G.err.begin(), G.err.end(), err_out);
}

// Edge residual + analytic Jacobians for one binary edge (finite-difference checks in tests).
void orc_edge_eval(const double* T_obj12, const double* T_cam12, const double* cam_k, const double* p, const double* uv,
                   double* err2, double* Ji12, double* Jj12) {
  Graph G; G.n_vert = 2; G.n_edges = 1;
  int eo = 0, ec = 1; G.e_obj = &eo; G.e_cam = &ec; G.cam_k = cam_k; G.p = p; G.uv = uv;
  double info[4] = {1, 0, 0, 1}; G.info = info; G.err.assign(2, 0.0);
  const double* Ts[2] = {T_obj12, T_cam12};
  G.est.resize(2);
  for (int v = 0; v < 2; ++v) { M3 R; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = Ts[v][4 * r + c]; double t[3] = {Ts[v][3], Ts[v][7], Ts[v][11]}; G.est[v] = se3_from_Rt(R, t); }
  edge_error(G, 0);
  err2[0] = G.err[0]; err2[1] = G.err[1];
  edge_jacobians(G, 0, Ji12, Jj12);
}

// exp(update) * T  -> 3x4 (for finite differences and SE3 tests)
void orc_se3_oplus(const double* T12, const double* upd6, double* out12) {
  M3 R; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = T12[4 * r + c];
  double t[3] = {T12[3], T12[7], T12[11]};
  SE3 T = se3_oplus(se3_from_Rt(R, t), upd6);
  M3 Ro = se3_R(T);
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) out12[4 * r + c] = Ro(r, c); out12[4 * r + 3] = T.t[r]; }
}

}  // extern "C"
