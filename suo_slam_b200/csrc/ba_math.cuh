// Device math shared by the bundle-adjustment kernels (ba.cu: block-diagonal graphs, ba_global.cu: coupled
// camera+object graphs): SE3Quat algebra, Huber kernel, 6x6 Cholesky, the packed-graph argument block.
#pragma once
#include "common.cuh"

namespace ba {

struct SE3q { double qx, qy, qz, qw, t[3]; };

__device__ inline void se3_from_Rt(const double* R, const double* t, SE3q& o) {   // Eigen::Quaterniond(R) + normalizeRotation (se3quat.h:55-57,277-282)
  double q[4];
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    double s = sqrt(tr + 1.0);
    q[3] = 0.5 * s; s = 0.5 / s;
    q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * s; s = 0.5 / s;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * s; q[j] = (R[3 * j + i] + R[3 * i + j]) * s; q[k] = (R[3 * k + i] + R[3 * i + k]) * s;
  }
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  o.qx = q[0] / n; o.qy = q[1] / n; o.qz = q[2] / n; o.qw = q[3] / n;
  o.t[0] = t[0]; o.t[1] = t[1]; o.t[2] = t[2];
}
__device__ __forceinline__ void se3_R(const SE3q& T, double* R) {          // Eigen toRotationMatrix
  const double tx = 2 * T.qx, ty = 2 * T.qy, tz = 2 * T.qz;
  const double twx = tx * T.qw, twy = ty * T.qw, twz = tz * T.qw;
  const double txx = tx * T.qx, txy = ty * T.qx, txz = tz * T.qx, tyy = ty * T.qy, tyz = tz * T.qy, tzz = tz * T.qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ __forceinline__ void se3_map(const SE3q& T, const double* p, double* o) {
  double R[9];
  se3_R(T, R);
  o[0] = R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + T.t[0];
  o[1] = R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + T.t[1];
  o[2] = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + T.t[2];
}
__device__ inline void se3_oplus(const SE3q& T, const double* u, SE3q& out) {     // exp(u) * T
  const double theta = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  const double Om[9] = {0, -u[2], u[1], u[2], 0, -u[0], -u[1], u[0], 0};
  double Om2[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Om2[3 * i + j] = Om[3 * i] * Om[j] + Om[3 * i + 1] * Om[3 + j] + Om[3 * i + 2] * Om[6 + j];
  double R[9], V[9];
  if (theta < 0.00001) {
#pragma unroll
    for (int i = 0; i < 9; ++i) { R[i] = (i % 4 == 0 ? 1.0 : 0.0) + Om[i] + Om2[i]; V[i] = R[i]; }
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / (theta * theta * theta);
#pragma unroll
    for (int i = 0; i < 9; ++i) { const double I = (i % 4 == 0 ? 1.0 : 0.0); R[i] = I + a * Om[i] + b * Om2[i]; V[i] = I + b * Om[i] + c * Om2[i]; }
  }
  const double td[3] = {V[0] * u[3] + V[1] * u[4] + V[2] * u[5], V[3] * u[3] + V[4] * u[4] + V[5] * u[5], V[6] * u[3] + V[7] * u[4] + V[8] * u[5]};
  SE3q E;
  se3_from_Rt(R, td, E);
  double RE[9];
  se3_R(E, RE);
  out.t[0] = E.t[0] + RE[0] * T.t[0] + RE[1] * T.t[1] + RE[2] * T.t[2];
  out.t[1] = E.t[1] + RE[3] * T.t[0] + RE[4] * T.t[1] + RE[5] * T.t[2];
  out.t[2] = E.t[2] + RE[6] * T.t[0] + RE[7] * T.t[1] + RE[8] * T.t[2];
  double w = E.qw * T.qw - E.qx * T.qx - E.qy * T.qy - E.qz * T.qz;
  double x = E.qw * T.qx + E.qx * T.qw + E.qy * T.qz - E.qz * T.qy;
  double y = E.qw * T.qy - E.qx * T.qz + E.qy * T.qw + E.qz * T.qx;
  double z = E.qw * T.qz + E.qx * T.qy - E.qy * T.qx + E.qz * T.qw;
  if (w < 0) { x = -x; y = -y; z = -z; w = -w; }
  const double n = sqrt(x * x + y * y + z * z + w * w);
  out.qx = x / n; out.qy = y / n; out.qz = z / n; out.qw = w / n;
}

// linearizeOplus of both edge types (types_object_slam.cpp:70-123: EdgeSE3ProjectFromObject, :177-201:
// EdgeSE3ProjectFromFixedObject).  Rcw = rotation of the camera vertex, pw = the keypoint in the world / camera-fixed frame
// (T_wo * p_O, or p_inG for the unary edge), pc = T_cw * pw, k = {fx, fy, cx, cy}.  Tangent order (omega, upsilon).
//   Ji [2x6] wrt the object vertex: projectJac * R_cw * [-[p_W]x | I]      (want_obj)
//   Jj [2x6] wrt the camera vertex: projectJac * [-[p_C]x | I]             (want_cam)
// with projectJac = -[[fx/z, 0, -fx x/z^2], [0, fy/z, -fy y/z^2]] at p_C.
__device__ __forceinline__ void edge_jacobians(const double* Rcw, const double* pw, const double* pc, const double* k,
                                               bool want_obj, bool want_cam, double* Ji, double* Jj) {
  const double iz = 1.0 / pc[2];
  const double pj[6] = {-(k[0] * iz), 0.0, k[0] * pc[0] * iz * iz, 0.0, -(k[1] * iz), k[1] * pc[1] * iz * iz};
  if (want_cam) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const double q0 = pj[3 * r], q1 = pj[3 * r + 1], q2 = pj[3 * r + 2];
      Jj[6 * r + 0] = -q1 * pc[2] + q2 * pc[1];
      Jj[6 * r + 1] = q0 * pc[2] - q2 * pc[0];
      Jj[6 * r + 2] = -q0 * pc[1] + q1 * pc[0];
      Jj[6 * r + 3] = q0; Jj[6 * r + 4] = q1; Jj[6 * r + 5] = q2;
    }
  }
  if (want_obj) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const double q0 = pj[3 * r] * Rcw[0] + pj[3 * r + 1] * Rcw[3] + pj[3 * r + 2] * Rcw[6];
      const double q1 = pj[3 * r] * Rcw[1] + pj[3 * r + 1] * Rcw[4] + pj[3 * r + 2] * Rcw[7];
      const double q2 = pj[3 * r] * Rcw[2] + pj[3 * r + 1] * Rcw[5] + pj[3 * r + 2] * Rcw[8];
      Ji[6 * r + 0] = -q1 * pw[2] + q2 * pw[1];
      Ji[6 * r + 1] = q0 * pw[2] - q2 * pw[0];
      Ji[6 * r + 2] = -q0 * pw[1] + q1 * pw[0];
      Ji[6 * r + 3] = q0; Ji[6 * r + 4] = q1; Ji[6 * r + 5] = q2;
    }
  }
}

__device__ __forceinline__ void huber(double e, double delta, double& rho0, double& rho1) {
  const double dsqr = delta * delta;
  if (e <= dsqr) { rho0 = e; rho1 = 1.0; }
  else { const double sq = sqrt(e); rho0 = 2 * sq * delta - dsqr; rho1 = delta / sq; }
}

__device__ inline bool chol6(double* A, double* b) {   // in place, row-major lower
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

struct BaArgs {
  const int32_t* prob_vert; const int32_t* prob_edge;
  const int32_t* vert_cnt; const int32_t* edge_cnt;   // optional explicit counts (else next offset - offset)
  double* poses; const uint8_t* fixed;
  const int32_t* e_obj; const int32_t* e_cam;
  const double* cam_k; const double* p; const double* uv; const double* info;
  uint8_t* inliers;
  const int32_t* its; int n_rounds;
  double huber_delta, chi2_gate; int init_with_outliers;
  int32_t* stats;
  double* err;        // scratch [n_edges,2]: the edge's _error as last computed (g2o keeps it in the edge)
  uint8_t* level;     // scratch [n_edges]
  int8_t* fv_kind;    // scratch [n_edges]: 0 = free vertex is the object, 1 = the camera, -1 = none / invalid
};

// Coupled camera+object graphs (ba_global.cu).  The host groups each problem's edges by (camera, object)
// pair; all index arrays use GLOBAL vertex / pair / edge numbers unless noted.
struct BgArgs {
  BaArgs g;
  const int32_t* perm;        // [n_edges] edge ids sorted by (camera, object) inside each problem
  const int32_t* pair_cam;    // [n_pairs]
  const int32_t* pair_obj;    // [n_pairs] object vertex or -1 (EdgeSE3ProjectFromFixedObject)
  const int32_t* pair_eoff;   // [n_pairs+1] ranges into perm
  const int32_t* prob_pair;   // [n_prob+1]
  const int32_t* cam_poff;    // [n_vert+1] pairs whose camera is vertex v (pairs are sorted by camera)
  const int32_t* obj_poff;    // [n_vert+1] ranges into obj_plist
  const int32_t* obj_plist;   // [n_pairs] pair ids grouped by object vertex, ascending camera
  const int32_t* obj_slot;    // [n_vert] row block of a free object vertex in its problem's reduced system, else -1
  const int32_t* slot_vert;   // per problem [n_obj]: vertex of each slot (offset prob_slot)
  const int32_t* prob_slot;   // [n_prob]
  const int32_t* peer;        // per problem [n_pairs_of_problem, n_obj]: pair (camera of p, object slot j) or -1
  const long long* prob_peer; // [n_prob] offsets into peer
  const int32_t* prob_nobj;   // [n_prob] free objects
  const long long* prob_S;    // [n_prob] offsets (doubles) into S: (6 n_obj)^2 + 2 * 6 n_obj per problem
  double* S;
  double* pairw;              // [n_pairs, 156]
  SE3q* est; SE3q* bak;       // [n_vert]
  double* Hv; double* bv; double* xv; double* Minv;   // [n_vert, 36 | 6 | 6 | 36]
  uint8_t* vact;              // [n_vert]
};

}  // namespace ba

int launch_ba_global(suo_ctx* ctx, int n_prob, const ba::BgArgs& args, cudaStream_t s);
// error + both Jacobians of a batch of edges at given vertex poses (test hook for the linearisation, ba.cu)
int launch_edge_linearize(suo_ctx* ctx, int n_edges, const double* T_obj, const double* T_cam, const double* cam_k, const double* p,
                          const double* uv, double* err, double* J_obj, double* J_cam, cudaStream_t s);
