// TEST INFRASTRUCTURE ONLY.  Thin C-ABI wrapper so tests can call the REFERENCE's own
// P3P/P4P (thirdparty/lambdatwist/p4p.cpp + lambdatwist/*.h, compiled where they lie
// under /root/reference by oracle/Makefile into oracle/_ref/libref_p4p.so).
// No reference source is copied: this file only #includes the reference headers.
#include <vector>
// single translation unit: the reference's p3p_timers.h defines a non-inline function,
// so p4p.cpp is compiled by inclusion (still from where it lies under /root/reference)
#include <p4p.cpp>

extern "C" {

void ref_p4p(const double* xs, const double* ys, int n, const int* idx4, double* T16) {
  std::vector<cvl::Vector3D> X(n);
  std::vector<cvl::Vector2D> Y(n);
  for (int i = 0; i < n; ++i) { X[i] = cvl::Vector3D(xs[3 * i], xs[3 * i + 1], xs[3 * i + 2]); Y[i] = cvl::Vector2D(ys[2 * i], ys[2 * i + 1]); }
  cvl::Vector4<uint> idx(idx4[0], idx4[1], idx4[2], idx4[3]);
  cvl::PoseD P = cvl::p4p(X, Y, idx);
  cvl::Matrix4x4D M = P.get4x4();
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T16[4 * r + c] = M(r, c);
}

int ref_p3p(const double* y2d, const double* x3d, double* Rs, double* Ts) {
  cvl::Vector<cvl::Matrix<double, 3, 3>, 4> R;
  cvl::Vector<cvl::Vector3<double>, 4> T;
  int v = cvl::p3p_lambdatwist<double, 5>(
      cvl::Vector3D(y2d[0], y2d[1], 1.0), cvl::Vector3D(y2d[2], y2d[3], 1.0), cvl::Vector3D(y2d[4], y2d[5], 1.0),
      cvl::Vector3D(x3d[0], x3d[1], x3d[2]), cvl::Vector3D(x3d[3], x3d[4], x3d[5]), cvl::Vector3D(x3d[6], x3d[7], x3d[8]), R, T);
  for (int i = 0; i < v; ++i) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Rs[9 * i + 3 * r + c] = R[i](r, c);
    for (int r = 0; r < 3; ++r) Ts[3 * i + r] = T[i][r];
  }
  return v;
}

}
