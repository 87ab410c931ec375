// Coupled camera+object bundle adjustment (SURVEY.md §8 rows a7/a8 global mode, row f3), FP64.
//
// ObjectSLAM.optimize(curr_only=False) in SLAM / SfM mode (lib/object_slam.py:736-778) frees every
// object vertex AND every camera but the first, adds one EdgeSE3ProjectFromObject per keypoint
// detection (types_object_slam.cpp:45-123), and lets g2o's Levenberg driver
// (optimization_algorithm_levenberg.cpp:58-175) solve the full sparse system with CHOLMOD
// (solvers/cholmod/linear_solver_cholmod.h:115-156).  The reference notes that a Schur complement
// "would make this faster" but segfaults in its g2o build (object_slam.py:775-776).
//
// Here the bipartite structure is used directly.  Every edge joins ONE object and ONE camera, so
//     H + lambda I = [ Hoo  Hoc ]      Hoo, Hcc block diagonal (6x6 per vertex)
//                    [ Hco  Hcc ]      Hoc: one 6x6 block per (object, camera) pair that shares edges
// and the cameras (the many) are eliminated:  S = Hoo - sum_c Hoc Hcc^-1 Hco,  g = bo - sum_c Hoc Hcc^-1 bc,
// S xo = g by a blocked dense Cholesky on the few objects, xc = Hcc^-1 (bc - Hco xo).  Mathematically
// the same step g2o computes (one lambda, one accept test per graph), so the LM trajectory matches the
// oracle's dense solve to rounding.
//
// One CTA per graph.  The host (api.cu) groups the edges by (camera, object) pair once; per LM
// iteration one warp per pair accumulates the pair's five blocks with a fixed shuffle tree, per-vertex
// totals and the Schur products are summed in pair order: deterministic run to run
// (lib/object_slam.py:440-442 insists on that).  Vertex state and blocks live in an L2-resident
// workspace (the graphs are a few hundred KB); latency / FP64-issue bound, no bandwidth claim.
#include "ba_math.cuh"
using namespace ba;

namespace {

constexpr int BG_THREADS = 256;
constexpr int BG_WARPS = BG_THREADS / 32;
constexpr int PW = 156;   // doubles per pair: A[36] B[36] C[36] bo[6] bc[6] Y[36]

__device__ double bg_block_sum(double v, double* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0;
#pragma unroll
  for (int w = 0; w < BG_WARPS; ++w) r += sh[w];
  return r;
}
__device__ double bg_block_max(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0;
#pragma unroll
  for (int w = 0; w < BG_WARPS; ++w) r = fmax(r, sh[w]);
  return r;
}

// lower Cholesky of a 6x6 block stored with row stride ld
__device__ bool factor6(double* A, int ld) {
  for (int j = 0; j < 6; ++j) {
    double d = A[j * ld + j];
    for (int k = 0; k < j; ++k) d -= A[j * ld + k] * A[j * ld + k];
    if (!(d > 0) || !isfinite(d)) return false;
    d = sqrt(d);
    A[j * ld + j] = d;
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i * ld + j];
      for (int k = 0; k < j; ++k) s -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = s / d;
    }
  }
  return true;
}

}  // namespace

namespace ba {

__global__ void __launch_bounds__(BG_THREADS)
ba_global_kernel(const BgArgs a) {
  __shared__ double red[BG_WARPS];
  __shared__ double s_lambda, s_ni, s_cur, s_rho;
  __shared__ int s_flag, s_ok;

  const BaArgs& g = a.g;
  const int prob = blockIdx.x;
  const int v0 = g.prob_vert[prob], nv = g.prob_vert[prob + 1] - v0;
  const int e0 = g.prob_edge[prob], ne = g.prob_edge[prob + 1] - e0;
  const int p0 = a.prob_pair[prob], np = a.prob_pair[prob + 1] - p0;
  const int No = a.prob_nobj[prob], n = 6 * No;
  double* S = a.S + a.prob_S[prob];
  double* gvec = S + (size_t)n * n;
  double* xo = gvec + n;
  const int32_t* peer = a.peer + a.prob_peer[prob];
  const int32_t* slot_vert = a.slot_vert + a.prob_slot[prob];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int v = tid; v < nv; v += BG_THREADS) {
    const double* T = g.poses + 12 * (size_t)(v0 + v);
    const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    const double t[3] = {T[3], T[7], T[11]};
    se3_from_Rt(R, t, a.est[v0 + v]);
  }
  __syncthreads();

  auto edge_error = [&](int ge) {   // computeError, types_object_slam.cpp:45-60,156-169
    double pw[3] = {g.p[3 * ge], g.p[3 * ge + 1], g.p[3 * ge + 2]}, pc[3];
    if (g.e_obj[ge] >= 0) { double tmp[3]; se3_map(a.est[g.e_obj[ge]], pw, tmp); pw[0] = tmp[0]; pw[1] = tmp[1]; pw[2] = tmp[2]; }
    se3_map(a.est[g.e_cam[ge]], pw, pc);
    const double* k = g.cam_k + 4 * ge;
    g.err[2 * ge] = g.uv[2 * ge] - (k[0] * pc[0] / pc[2] + k[2]);
    g.err[2 * ge + 1] = g.uv[2 * ge + 1] - (k[1] * pc[1] / pc[2] + k[3]);
  };
  auto edge_chi2 = [&](int ge) {
    const double* O = g.info + 4 * ge;
    const double r0 = g.err[2 * ge], r1 = g.err[2 * ge + 1];
    return r0 * (O[0] * r0 + O[1] * r1) + r1 * (O[2] * r0 + O[3] * r1);
  };
  // active = level 0 and not all vertices fixed (SparseOptimizer::initializeOptimization, sparse_optimizer.cpp:201-272)
  auto is_active = [&](int ge) {
    return g.level[ge] == 0 && (!g.fixed[g.e_cam[ge]] || (g.e_obj[ge] >= 0 && !g.fixed[g.e_obj[ge]]));
  };
  auto chi_sum = [&](int robust) {  // computeActiveErrors + activeRobustChi2 (sparse_optimizer.cpp:63-116)
    double part = 0;
    for (int e = tid; e < ne; e += BG_THREADS) {
      const int ge = e0 + e;
      if (!is_active(ge)) continue;
      edge_error(ge);
      const double c = edge_chi2(ge);
      if (robust) { double r0, r1; huber(c, g.huber_delta, r0, r1); part += r0; } else part += c;
    }
    return bg_block_sum(part, red);
  };

  // ---- initial chi2 classification (object_slam.py:848-866) ----
  int my_good = 0;
  for (int e = tid; e < ne; e += BG_THREADS) {
    const int ge = e0 + e;
    if (g.init_with_outliers) { g.level[ge] = 0; g.err[2 * ge] = 0.0; g.err[2 * ge + 1] = 0.0; my_good++; }   // an edge that never becomes active keeps a zero error
    else {
      edge_error(ge);
      if (edge_chi2(ge) > g.chi2_gate) { g.level[ge] = 1; g.inliers[ge] = 0; }
      else { g.level[ge] = 0; g.inliers[ge] = 1; my_good++; }
    }
  }
  int num_good = (int)(bg_block_sum((double)my_good, red) + 0.5);
  int robust = 1;
  int rounds = 0, outer_total = 0, trials_total = 0;

  for (int round = 0; round < g.n_rounds; ++round) {
    if (ne < 4 || num_good < 4) break;
    // ---------------- initializeOptimization(0) ----------------
    for (int v = tid; v < nv; v += BG_THREADS) a.vact[v0 + v] = 0;
    __syncthreads();
    int mine = 0;
    for (int e = tid; e < ne; e += BG_THREADS) {
      const int ge = e0 + e;
      if (!is_active(ge)) continue;
      if (!g.fixed[g.e_cam[ge]]) { a.vact[g.e_cam[ge]] = 1; mine = 1; }            // benign race: all write 1
      if (g.e_obj[ge] >= 0 && !g.fixed[g.e_obj[ge]]) { a.vact[g.e_obj[ge]] = 1; mine = 1; }
    }
    const int any = __syncthreads_or(mine);
    ++rounds;
    if (any) {
      bool ok = true;
      const int iters = g.its[round];
      for (int it = 0; it < iters && ok; ++it) {
        const double currentChi0 = chi_sum(robust);
        // ---------------- buildSystem: one warp per (camera, object) pair ----------------
        for (int p = warp; p < np; p += BG_WARPS) {
          const int gp = p0 + p;
          const int cam = a.pair_cam[gp], obj = a.pair_obj[gp];
          const bool fc = !g.fixed[cam], fo = obj >= 0 && !g.fixed[obj];
          double* W = a.pairw + (size_t)gp * PW;
          for (int k = lane; k < 120; k += 32) W[k] = 0.0;
          __syncwarp();
          if (!fc && !fo) continue;
          const int eb = a.pair_eoff[gp], ee = a.pair_eoff[gp + 1];
          const SE3q Tcw = a.est[cam];
          double Rcw[9];
          se3_R(Tcw, Rcw);
          for (int base = eb; base < ee; base += 32) {
            const int slot = base + lane;
            const int ge = slot < ee ? a.perm[slot] : -1;
            double Ji[12], Jj[12], or0 = 0, or1 = 0, w00 = 0, w01 = 0, w10 = 0, w11 = 0;
#pragma unroll
            for (int k = 0; k < 12; ++k) { Ji[k] = 0; Jj[k] = 0; }
            if (ge >= 0 && is_active(ge)) {
              // linearizeOplus, types_object_slam.cpp:70-123,177-201
              double pw[3] = {g.p[3 * ge], g.p[3 * ge + 1], g.p[3 * ge + 2]}, pc[3];
              if (obj >= 0) { double tmp[3]; se3_map(a.est[obj], pw, tmp); pw[0] = tmp[0]; pw[1] = tmp[1]; pw[2] = tmp[2]; }
              pc[0] = Rcw[0] * pw[0] + Rcw[1] * pw[1] + Rcw[2] * pw[2] + Tcw.t[0];
              pc[1] = Rcw[3] * pw[0] + Rcw[4] * pw[1] + Rcw[5] * pw[2] + Tcw.t[1];
              pc[2] = Rcw[6] * pw[0] + Rcw[7] * pw[1] + Rcw[8] * pw[2] + Tcw.t[2];
              const double* k = g.cam_k + 4 * ge;
              edge_jacobians(Rcw, pw, pc, k, fo, fc, Ji, Jj);
              const double* O = g.info + 4 * ge;
              const double r0 = g.err[2 * ge], r1 = g.err[2 * ge + 1];
              double w = 1.0;
              if (robust) { double h0; huber(r0 * (O[0] * r0 + O[1] * r1) + r1 * (O[2] * r0 + O[3] * r1), g.huber_delta, h0, w); }
              or0 = -(O[0] * r0 + O[1] * r1) * w; or1 = -(O[2] * r0 + O[3] * r1) * w;   // omega_r * rho'
              w00 = O[0] * w; w01 = O[1] * w; w10 = O[2] * w; w11 = O[3] * w;          // robustInformation
            }
            // J^T W (2 columns per Jacobian row index)
            double iW0[6], iW1[6], jW0[6], jW1[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              iW0[c] = Ji[c] * w00 + Ji[6 + c] * w10; iW1[c] = Ji[c] * w01 + Ji[6 + c] * w11;
              jW0[c] = Jj[c] * w00 + Jj[6 + c] * w10; jW1[c] = Jj[c] * w01 + Jj[6 + c] * w11;
            }
            // reduce every output entry over the 32 edges of this chunk; lane 0 adds it (chunk order fixed)
#pragma unroll
            for (int r = 0; r < 6; ++r) {
#pragma unroll
              for (int c = 0; c < 6; ++c) {
                const double vB = warp_sum(iW0[r] * Jj[c] + iW1[r] * Jj[6 + c]);
                if (lane == 0) W[36 + 6 * r + c] += vB;
                if (c >= r) {
                  const double vA = warp_sum(iW0[r] * Ji[c] + iW1[r] * Ji[6 + c]);
                  const double vC = warp_sum(jW0[r] * Jj[c] + jW1[r] * Jj[6 + c]);
                  if (lane == 0) { W[6 * r + c] += vA; W[72 + 6 * r + c] += vC; }
                }
              }
              const double vbo = warp_sum(Ji[r] * or0 + Ji[6 + r] * or1);
              const double vbc = warp_sum(Jj[r] * or0 + Jj[6 + r] * or1);
              if (lane == 0) { W[108 + r] += vbo; W[114 + r] += vbc; }
            }
          }
          __syncwarp();
          if (lane < 15) {   // mirror the upper triangles of A and C
            int r = 0, c = lane + 1, row = 5;
            while (c > row) { c -= row; ++r; --row; }
            c += r;           // (r, c) with c > r
            W[6 * c + r] = W[6 * r + c];
            W[72 + 6 * c + r] = W[72 + 6 * r + c];
          }
        }
        __syncthreads();
        // per-vertex totals in pair order
        for (int idx = tid; idx < nv * 42; idx += BG_THREADS) {
          const int gv = v0 + idx / 42, k = idx % 42;
          double s = 0;
          if (a.vact[gv]) {
            if (a.obj_slot[gv] >= 0) {
              for (int pl = a.obj_poff[gv]; pl < a.obj_poff[gv + 1]; ++pl) { const double* W = a.pairw + (size_t)a.obj_plist[pl] * PW; s += k < 36 ? W[k] : W[108 + k - 36]; }
            } else {
              for (int gp = a.cam_poff[gv]; gp < a.cam_poff[gv + 1]; ++gp) { const double* W = a.pairw + (size_t)gp * PW; s += k < 36 ? W[72 + k] : W[114 + k - 36]; }
            }
          }
          if (k < 36) a.Hv[(size_t)gv * 36 + k] = s; else a.bv[(size_t)gv * 6 + k - 36] = s;
        }
        __syncthreads();
        if (it == 0) {   // computeLambdaInit, optimization_algorithm_levenberg.cpp:152-166
          double mx = 0;
          for (int idx = tid; idx < nv * 6; idx += BG_THREADS) {
            const int gv = v0 + idx / 6;
            if (a.vact[gv]) mx = fmax(mx, fabs(a.Hv[(size_t)gv * 36 + 7 * (idx % 6)]));
          }
          mx = bg_block_max(mx, red);
          if (tid == 0) { s_lambda = 1e-5 * mx; s_ni = 2.0; }
        }
        if (tid == 0) s_cur = currentChi0;
        __syncthreads();
        double rho = 0;
        int qmax = 0;
        bool lam_bad = false;
        do {
          const double lambda = s_lambda;
          // push(); cameras: Minv = (Hcc + lambda I)^-1
          int my_ok = 1;
          for (int v = tid; v < nv; v += BG_THREADS) {
            const int gv = v0 + v;
            a.bak[gv] = a.est[gv];
            if (a.vact[gv] && a.obj_slot[gv] < 0) {
              double A[36], M[36];
              for (int k = 0; k < 36; ++k) A[k] = a.Hv[(size_t)gv * 36 + k];
              for (int j = 0; j < 6; ++j) A[7 * j] += lambda;
              const bool okv = factor6(A, 6);
              if (okv) {
                for (int c = 0; c < 6; ++c) {
                  double b[6];
                  for (int i = 0; i < 6; ++i) { double s = (i == c) ? 1.0 : 0.0; for (int k = 0; k < i; ++k) s -= A[i * 6 + k] * b[k]; b[i] = s / A[i * 6 + i]; }
                  for (int i = 5; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < 6; ++k) s -= A[k * 6 + i] * b[k]; b[i] = s / A[i * 6 + i]; }
                  for (int i = 0; i < 6; ++i) M[6 * i + c] = b[i];
                }
              } else { my_ok = 0; for (int k = 0; k < 36; ++k) M[k] = 0.0; }
              for (int k = 0; k < 36; ++k) a.Minv[(size_t)gv * 36 + k] = M[k];
            }
          }
          if (tid == 0) s_ok = 1;
          __syncthreads();
          // Y_p = Hoc_p * Minv_cam
          for (int idx = tid; idx < np * 36; idx += BG_THREADS) {
            const int gp = p0 + idx / 36, k = idx % 36, r = k / 6, c = k % 6;
            double* W = a.pairw + (size_t)gp * PW;
            const int cam = a.pair_cam[gp], obj = a.pair_obj[gp];
            double s = 0;
            if (obj >= 0 && !g.fixed[obj] && !g.fixed[cam] && a.vact[cam]) {
              const double* M = a.Minv + (size_t)cam * 36;
#pragma unroll
              for (int t = 0; t < 6; ++t) s += W[36 + 6 * r + t] * M[6 * t + c];
            }
            W[120 + k] = s;
          }
          __syncthreads();
          // reduced system on the objects
          for (int idx = tid; idx < No * No * 36; idx += BG_THREADS) {
            const int bi = idx / (No * 36), rem = idx % (No * 36), bj = rem / 36, k = rem % 36, r = k / 6, c = k % 6;
            if (bj > bi) continue;
            const int ov = slot_vert[bi];
            double s;
            if (!a.vact[ov]) s = (bi == bj && r == c) ? 1.0 : 0.0;
            else {
              s = (bi == bj) ? a.Hv[(size_t)ov * 36 + k] + (r == c ? lambda : 0.0) : 0.0;
              for (int pl = a.obj_poff[ov]; pl < a.obj_poff[ov + 1]; ++pl) {
                const int gp = a.obj_plist[pl];
                const int q = peer[(size_t)(gp - p0) * No + bj];
                if (q < 0) continue;
                const double* Y = a.pairw + (size_t)gp * PW + 120;
                const double* B = a.pairw + (size_t)q * PW + 36;
#pragma unroll
                for (int t = 0; t < 6; ++t) s -= Y[6 * r + t] * B[6 * c + t];
              }
            }
            S[(size_t)(6 * bi + r) * n + 6 * bj + c] = s;
          }
          for (int idx = tid; idx < n; idx += BG_THREADS) {
            const int bi = idx / 6, r = idx % 6, ov = slot_vert[bi];
            double s = 0;
            if (a.vact[ov]) {
              s = a.bv[(size_t)ov * 6 + r];
              for (int pl = a.obj_poff[ov]; pl < a.obj_poff[ov + 1]; ++pl) {
                const int gp = a.obj_plist[pl];
                const double* Y = a.pairw + (size_t)gp * PW + 120;
                const double* bc = a.bv + (size_t)a.pair_cam[gp] * 6;
#pragma unroll
                for (int t = 0; t < 6; ++t) s -= Y[6 * r + t] * bc[t];
              }
            }
            gvec[idx] = s;
          }
          __syncthreads();
          // blocked Cholesky S = L L^T (lower, in place), 6 columns per step
          for (int kb = 0; kb < No; ++kb) {
            const int c0 = 6 * kb;
            if (tid == 0 && !factor6(S + (size_t)c0 * n + c0, n)) s_ok = 0;
            __syncthreads();
            if (!s_ok) break;
            const double* Lkk = S + (size_t)c0 * n + c0;
            for (int i = c0 + 6 + tid; i < n; i += BG_THREADS) {
              double* row = S + (size_t)i * n + c0;
#pragma unroll
              for (int c = 0; c < 6; ++c) {
                double s = row[c];
                for (int t = 0; t < c; ++t) s -= row[t] * Lkk[(size_t)c * n + t];
                row[c] = s / Lkk[(size_t)c * n + c];
              }
            }
            __syncthreads();
            const int m = n - c0 - 6;
            for (int idx = tid; idx < m * m; idx += BG_THREADS) {
              const int i = c0 + 6 + idx / m, j = c0 + 6 + idx % m;
              if (j > i) continue;
              const double* li = S + (size_t)i * n + c0;
              const double* lj = S + (size_t)j * n + c0;
              double s = S[(size_t)i * n + j];
#pragma unroll
              for (int t = 0; t < 6; ++t) s -= li[t] * lj[t];
              S[(size_t)i * n + j] = s;
            }
            __syncthreads();
          }
          const int chol_ok = s_ok;
          if (chol_ok) {
            // L y = g
            for (int kb = 0; kb < No; ++kb) {
              const int c0 = 6 * kb;
              if (tid == 0) {
                for (int i = 0; i < 6; ++i) { double s = gvec[c0 + i]; for (int t = 0; t < i; ++t) s -= S[(size_t)(c0 + i) * n + c0 + t] * gvec[c0 + t]; gvec[c0 + i] = s / S[(size_t)(c0 + i) * n + c0 + i]; }
              }
              __syncthreads();
              for (int i = c0 + 6 + tid; i < n; i += BG_THREADS) {
                double s = gvec[i];
#pragma unroll
                for (int t = 0; t < 6; ++t) s -= S[(size_t)i * n + c0 + t] * gvec[c0 + t];
                gvec[i] = s;
              }
              __syncthreads();
            }
            // L^T x = y
            for (int kb = No - 1; kb >= 0; --kb) {
              const int c0 = 6 * kb;
              if (tid == 0) {
                for (int i = 5; i >= 0; --i) { double s = gvec[c0 + i]; for (int t = i + 1; t < 6; ++t) s -= S[(size_t)(c0 + t) * n + c0 + i] * xo[c0 + t]; xo[c0 + i] = s / S[(size_t)(c0 + i) * n + c0 + i]; }
              }
              __syncthreads();
              for (int i = tid; i < c0; i += BG_THREADS) {
                double s = gvec[i];
#pragma unroll
                for (int t = 0; t < 6; ++t) s -= S[(size_t)(c0 + t) * n + i] * xo[c0 + t];
                gvec[i] = s;
              }
              __syncthreads();
            }
          } else {
            for (int i = tid; i < n; i += BG_THREADS) xo[i] = 0.0;
            __syncthreads();
          }
          // back-substitute the cameras, update()
          double sc = 0;
          for (int v = tid; v < nv; v += BG_THREADS) {
            const int gv = v0 + v;
            if (!a.vact[gv]) continue;
            double x[6];
            const int sl = a.obj_slot[gv];
            if (sl >= 0) { for (int j = 0; j < 6; ++j) x[j] = xo[6 * sl + j]; }
            else {
              double rhs[6];
              for (int j = 0; j < 6; ++j) rhs[j] = a.bv[(size_t)gv * 6 + j];
              for (int gp = a.cam_poff[gv]; gp < a.cam_poff[gv + 1]; ++gp) {
                const int obj = a.pair_obj[gp];
                if (obj < 0 || g.fixed[obj] || !a.vact[obj]) continue;
                const double* B = a.pairw + (size_t)gp * PW + 36;
                const double* xs = xo + 6 * a.obj_slot[obj];
                for (int j = 0; j < 6; ++j) { double s = 0; for (int t = 0; t < 6; ++t) s += B[6 * t + j] * xs[t]; rhs[j] -= s; }
              }
              const double* M = a.Minv + (size_t)gv * 36;
              for (int j = 0; j < 6; ++j) { double s = 0; for (int t = 0; t < 6; ++t) s += M[6 * j + t] * rhs[t]; x[j] = chol_ok ? s : 0.0; }
            }
            for (int j = 0; j < 6; ++j) { a.xv[(size_t)gv * 6 + j] = x[j]; sc += x[j] * (lambda * x[j] + a.bv[(size_t)gv * 6 + j]); }   // computeScale
            SE3q nw;
            se3_oplus(a.est[gv], x, nw);
            a.est[gv] = nw;
          }
          const int all_ok = __syncthreads_and(my_ok) && chol_ok;
          const double scale = bg_block_sum(sc, red);
          double tempChi = chi_sum(robust);
          if (tid == 0) {
            if (!all_ok) tempChi = 1.7976931348623157e308;
            const double r = (s_cur - tempChi) / (scale + 1e-3);
            int flag;
            if (r > 0 && isfinite(tempChi)) {
              double alpha = 1. - pow(2 * r - 1, 3.0);
              alpha = fmin(alpha, 2. / 3.);
              s_lambda *= fmax(1. / 3., alpha);
              s_ni = 2; s_cur = tempChi; flag = 1;
            } else {
              s_lambda *= s_ni; s_ni *= 2; flag = 0;
              if (!isfinite(s_lambda)) flag = 2;
            }
            s_rho = r; s_flag = flag;
          }
          __syncthreads();
          rho = s_rho;
          const int flag = s_flag;
          if (flag != 1) {      // pop(): restore vertices; edge errors stay those of the rejected trial
            for (int v = tid; v < nv; v += BG_THREADS) a.est[v0 + v] = a.bak[v0 + v];
          }
          __syncthreads();
          ++trials_total;
          if (flag == 2) { lam_bad = true; break; }
          qmax++;
        } while (rho < 0 && qmax < 10);
        ++outer_total;
        if (qmax == 10 || rho == 0 || lam_bad) ok = false;   // Terminate
      }
    }
    // ---------------- chi2 re-classification (object_slam.py:878-896) ----------------
    my_good = 0;
    for (int e = tid; e < ne; e += BG_THREADS) {
      const int ge = e0 + e;
      if (!g.inliers[ge]) edge_error(ge);
      if (edge_chi2(ge) > g.chi2_gate) { g.level[ge] = 1; g.inliers[ge] = 0; }
      else { g.level[ge] = 0; g.inliers[ge] = 1; my_good++; }
    }
    num_good = (int)(bg_block_sum((double)my_good, red) + 0.5);
    if (round == max(1, g.n_rounds / 2)) robust = 0;
  }
  __syncthreads();
  for (int v = tid; v < nv; v += BG_THREADS) {
    double R[9];
    se3_R(a.est[v0 + v], R);
    double* T = g.poses + 12 * (size_t)(v0 + v);
    for (int r = 0; r < 3; ++r) { T[4 * r] = R[3 * r]; T[4 * r + 1] = R[3 * r + 1]; T[4 * r + 2] = R[3 * r + 2]; T[4 * r + 3] = a.est[v0 + v].t[r]; }
  }
  if (tid == 0 && g.stats) { g.stats[3 * prob] = rounds; g.stats[3 * prob + 1] = outer_total; g.stats[3 * prob + 2] = trials_total; }
}

}  // namespace ba

int launch_ba_global(suo_ctx* ctx, int n_prob, const ba::BgArgs& args, cudaStream_t s) {
  if (n_prob <= 0) return SUO_OK;
  ba::ba_global_kernel<<<n_prob, BG_THREADS, 0, s>>>(args);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}
