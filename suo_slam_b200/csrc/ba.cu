// Batched Levenberg-Marquardt object-pose bundle adjustment, FP64.
//
// Replaces, for a BATCH of independent graphs in one launch, what the reference does through
// thousands of pybind calls per frame in ObjectSLAM.optimize() (lib/object_slam.py:842-896):
//   * EdgeSE3ProjectFromObject / EdgeSE3ProjectFromFixedObject computeError + linearizeOplus
//     (thirdparty/g2opy/g2o/types/object_slam/types_object_slam.cpp:45-60,70-123,156-169,177-201)
//   * BaseBinaryEdge/BaseUnaryEdge::constructQuadraticForm with RobustKernelHuber
//     (g2o/core/base_binary_edge.hpp:64-127, base_unary_edge.hpp:52-78, robust_kernel_impl.cpp:65-78)
//   * OptimizationAlgorithmLevenberg::solve (g2o/core/optimization_algorithm_levenberg.cpp:58-175):
//     lambda0 = 1e-5 max diag(H); (H + lambda I) dx = b; T <- exp(dx) T; rho test with one lambda and
//     one accept/reject per graph; <= 10 trials per iteration
//   * VertexSE3Expmap::oplusImpl / SE3Quat::exp (types/sba/types_six_dof_expmap.h:100-103,
//     types/slam3d/se3quat.h:220-254) and the dense LDLT solve (solvers/dense/linear_solver_dense.h:65-113)
//   * the 4-round chi2 re-classification / Huber-stripping schedule of optimize() itself.
//
// One CTA per graph; the graph's vertex states, 6x6 Hessian blocks and LM scalars live in shared
// memory; edges are streamed from global memory.  Every edge has exactly one free vertex here
// (single-view mode: camera fixed; curr_only: objects folded into p_inG), so H is block-diagonal.
// Hessian blocks are accumulated by one warp per vertex in a fixed order (lane-strided partial sums
// + shuffle tree): results are deterministic run to run, as the reference insists on
// (lib/object_slam.py:440-442).  Latency / FP64-issue bound; no bandwidth claim.
#include "common.cuh"
#include "ba_math.cuh"
using namespace ba;

namespace {

constexpr int BA_THREADS = 128;
constexpr int BA_WARPS = BA_THREADS / 32;
constexpr int BA_MAXV = 64;     // vertices per graph held in shared memory


__device__ double block_sum_d(double v, double* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0;
#pragma unroll
  for (int w = 0; w < BA_WARPS; ++w) r += sh[w];
  return r;
}

__global__ void __launch_bounds__(BA_THREADS)
ba_kernel(const BaArgs a) {
  __shared__ SE3q est[BA_MAXV], bak[BA_MAXV];
  __shared__ double Hs[BA_MAXV][36], bs[BA_MAXV][6], xsol[BA_MAXV][6];
  __shared__ uint8_t vact[BA_MAXV], vok[BA_MAXV];
  __shared__ int vfirst[BA_MAXV], vlast[BA_MAXV];      // edge range that holds every edge whose free vertex is v (exact when the edges are grouped by vertex)
  __shared__ double red[BA_WARPS];
  __shared__ double s_lambda, s_ni, s_cur, s_rho;
  __shared__ int s_flag, s_bad;

  const int prob = blockIdx.x;
  const int v0 = a.prob_vert[prob], nv = a.vert_cnt ? a.vert_cnt[prob] : a.prob_vert[prob + 1] - v0;
  const int e0 = a.prob_edge[prob], ne = a.edge_cnt ? a.edge_cnt[prob] : a.prob_edge[prob + 1] - e0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  if (nv > BA_MAXV) { if (tid == 0 && a.stats) { a.stats[3 * prob] = -1; a.stats[3 * prob + 1] = 0; a.stats[3 * prob + 2] = 0; } return; }

  for (int v = tid; v < nv; v += BA_THREADS) {
    const double* T = a.poses + 12 * (size_t)(v0 + v);
    const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    const double t[3] = {T[3], T[7], T[11]};
    se3_from_Rt(R, t, est[v]);
  }
  for (int v = tid; v < nv; v += BA_THREADS) { vfirst[v] = 0x7fffffff; vlast[v] = -1; }
  __syncthreads();
  // which vertex of each edge is free
  for (int e = tid; e < ne; e += BA_THREADS) {
    const int ge = e0 + e;
    const int vo = a.e_obj[ge], vc = a.e_cam[ge];
    const bool fo = vo >= 0 && !a.fixed[vo], fc = !a.fixed[vc];
    int8_t k = -1;
    if (fo && fc) { k = -1; atomicExch(&s_bad, 1); }      // coupled camera+object graph: SURVEY §8 f3, not handled here
    else if (fo) k = 0;
    else if (fc) k = 1;
    a.fv_kind[ge] = k;
    if (k >= 0) { const int fv = (k == 0 ? vo : vc) - v0; atomicMin(&vfirst[fv], e); atomicMax(&vlast[fv], e); }
  }
  __syncthreads();
  if (s_bad) { if (tid == 0 && a.stats) { a.stats[3 * prob] = -2; a.stats[3 * prob + 1] = 0; a.stats[3 * prob + 2] = 0; } return; }

  auto edge_error = [&](int ge) {   // computeError
    double pw[3] = {a.p[3 * ge], a.p[3 * ge + 1], a.p[3 * ge + 2]}, pc[3];
    if (a.e_obj[ge] >= 0) { double tmp[3]; se3_map(est[a.e_obj[ge] - v0], pw, tmp); pw[0] = tmp[0]; pw[1] = tmp[1]; pw[2] = tmp[2]; }
    se3_map(est[a.e_cam[ge] - v0], pw, pc);
    const double* k = a.cam_k + 4 * ge;
    a.err[2 * ge] = a.uv[2 * ge] - (k[0] * pc[0] / pc[2] + k[2]);
    a.err[2 * ge + 1] = a.uv[2 * ge + 1] - (k[1] * pc[1] / pc[2] + k[3]);
  };
  auto edge_chi2 = [&](int ge) {
    const double* O = a.info + 4 * ge;
    const double r0 = a.err[2 * ge], r1 = a.err[2 * ge + 1];
    return r0 * (O[0] * r0 + O[1] * r1) + r1 * (O[2] * r0 + O[3] * r1);
  };
  // active = level 0 and has a free vertex (edges whose vertices are all fixed are dropped by
  // SparseOptimizer::initializeOptimization)
  auto is_active = [&](int ge) { return a.level[ge] == 0 && a.fv_kind[ge] >= 0; };

  // ---- initial chi2 classification (object_slam.py:848-866) ----
  int my_good = 0;
  for (int e = tid; e < ne; e += BA_THREADS) {
    const int ge = e0 + e;
    if (a.init_with_outliers) { a.level[ge] = 0; a.err[2 * ge] = 0.0; a.err[2 * ge + 1] = 0.0; my_good++; }   // an edge that never becomes active keeps a zero error
    else {
      edge_error(ge);
      if (edge_chi2(ge) > a.chi2_gate) { a.level[ge] = 1; a.inliers[ge] = 0; }
      else { a.level[ge] = 0; a.inliers[ge] = 1; my_good++; }
    }
  }
  int num_good = (int)(block_sum_d((double)my_good, red) + 0.5);
  int robust = 1;
  int rounds = 0, outer_total = 0, trials_total = 0;

  for (int round = 0; round < a.n_rounds; ++round) {
    if (ne < 4 || num_good < 4) break;
    // ---------------- initializeOptimization(0) ----------------
    for (int v = tid; v < nv; v += BA_THREADS) vact[v] = 0;
    __syncthreads();
    for (int e = tid; e < ne; e += BA_THREADS) {
      const int ge = e0 + e;
      if (is_active(ge)) vact[(a.fv_kind[ge] == 0 ? a.e_obj[ge] : a.e_cam[ge]) - v0] = 1;   // benign race: all write 1
    }
    __syncthreads();
    int any = 0;
    for (int v = 0; v < nv; ++v) any |= vact[v];
    ++rounds;
    if (any) {
      // ---------------- optimize(its[round]) ----------------
      bool ok = true;
      const int iters = a.its[round];
      for (int it = 0; it < iters && ok; ++it) {
        // computeActiveErrors + activeRobustChi2
        double part = 0;
        for (int e = tid; e < ne; e += BA_THREADS) {
          const int ge = e0 + e;
          if (!is_active(ge)) continue;
          edge_error(ge);
          const double c = edge_chi2(ge);
          if (robust) { double r0, r1; huber(c, a.huber_delta, r0, r1); part += r0; } else part += c;
        }
        const double currentChi0 = block_sum_d(part, red);
        // buildSystem: one warp per vertex, fixed summation order
        for (int v = warp; v < nv; v += BA_WARPS) {
          double acc[27];
#pragma unroll
          for (int k = 0; k < 27; ++k) acc[k] = 0;
          if (vact[v]) {
            for (int e = vfirst[v] + lane; e <= vlast[v]; e += 32) {
              const int ge = e0 + e;
              if (!is_active(ge)) continue;
              const int fvert = (a.fv_kind[ge] == 0 ? a.e_obj[ge] : a.e_cam[ge]) - v0;
              if (fvert != v) continue;
              // linearizeOplus
              double pw[3] = {a.p[3 * ge], a.p[3 * ge + 1], a.p[3 * ge + 2]}, pc[3];
              if (a.e_obj[ge] >= 0) { double tmp[3]; se3_map(est[a.e_obj[ge] - v0], pw, tmp); pw[0] = tmp[0]; pw[1] = tmp[1]; pw[2] = tmp[2]; }
              const SE3q& Tcw = est[a.e_cam[ge] - v0];
              se3_map(Tcw, pw, pc);
              const double* k = a.cam_k + 4 * ge;
              double J[12], Jx[12];
              double Rcw[9];
              se3_R(Tcw, Rcw);
              const bool wrt_obj = a.fv_kind[ge] == 0;
              edge_jacobians(Rcw, pw, pc, k, wrt_obj, !wrt_obj, wrt_obj ? J : Jx, wrt_obj ? Jx : J);
              const double* O = a.info + 4 * ge;
              const double r0 = a.err[2 * ge], r1 = a.err[2 * ge + 1];
              double w = 1.0;
              if (robust) { double h0; huber(r0 * (O[0] * r0 + O[1] * r1) + r1 * (O[2] * r0 + O[3] * r1), a.huber_delta, h0, w); }
              const double or0 = -(O[0] * r0 + O[1] * r1) * w, or1 = -(O[2] * r0 + O[3] * r1) * w;
              const double w00 = O[0] * w, w01 = O[1] * w, w10 = O[2] * w, w11 = O[3] * w;
              int idx = 6;
#pragma unroll
              for (int c = 0; c < 6; ++c) {
                acc[c] += J[c] * or0 + J[6 + c] * or1;
                const double a0 = J[c] * w00 + J[6 + c] * w10, a1 = J[c] * w01 + J[6 + c] * w11;   // row c of J^T W
#pragma unroll
                for (int d = c; d < 6; ++d) acc[idx++] += a0 * J[d] + a1 * J[6 + d];
              }
            }
#pragma unroll
            for (int k = 0; k < 27; ++k) acc[k] = warp_sum(acc[k]);
          }
          if (lane == 0) {
            int idx = 6;
            for (int c = 0; c < 6; ++c) {
              bs[v][c] = acc[c];
              for (int d = c; d < 6; ++d) { Hs[v][6 * c + d] = acc[idx]; Hs[v][6 * d + c] = acc[idx]; ++idx; }
            }
          }
        }
        __syncthreads();
        if (tid == 0) {
          s_cur = currentChi0;
          if (it == 0) {   // computeLambdaInit
            double mx = 0;
            for (int v = 0; v < nv; ++v) if (vact[v]) for (int j = 0; j < 6; ++j) mx = fmax(fabs(Hs[v][7 * j]), mx);
            s_lambda = 1e-5 * mx; s_ni = 2.0;
          }
        }
        __syncthreads();
        double rho = 0;
        int qmax = 0;
        bool lam_bad = false;
        do {
          // push + solve (H + lambda I) x = b per active vertex, update
          for (int v = tid; v < nv; v += BA_THREADS) {
            bak[v] = est[v];
            vok[v] = 1;
            if (vact[v]) {
              double A[36], rhs[6];
              for (int k = 0; k < 36; ++k) A[k] = Hs[v][k];
              for (int j = 0; j < 6; ++j) { A[7 * j] += s_lambda; rhs[j] = bs[v][j]; }
              const bool okv = chol6(A, rhs);
              vok[v] = okv;
              for (int j = 0; j < 6; ++j) xsol[v][j] = okv ? rhs[j] : 0.0;
              SE3q nw;
              se3_oplus(est[v], xsol[v], nw);
              est[v] = nw;
            }
          }
          __syncthreads();
          double part2 = 0;
          for (int e = tid; e < ne; e += BA_THREADS) {
            const int ge = e0 + e;
            if (!is_active(ge)) continue;
            edge_error(ge);
            const double c = edge_chi2(ge);
            if (robust) { double h0, h1; huber(c, a.huber_delta, h0, h1); part2 += h0; } else part2 += c;
          }
          double tempChi = block_sum_d(part2, red);
          if (tid == 0) {
            bool ok2 = true;
            double scale = 0;
            for (int v = 0; v < nv; ++v) if (vact[v]) {
              ok2 = ok2 && vok[v];
              for (int j = 0; j < 6; ++j) scale += xsol[v][j] * (s_lambda * xsol[v][j] + bs[v][j]);   // computeScale
            }
            if (!ok2) tempChi = 1.7976931348623157e308;
            double r = (s_cur - tempChi) / (scale + 1e-3);
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
            for (int v = tid; v < nv; v += BA_THREADS) est[v] = bak[v];
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
    for (int e = tid; e < ne; e += BA_THREADS) {
      const int ge = e0 + e;
      if (!a.inliers[ge]) edge_error(ge);
      if (edge_chi2(ge) > a.chi2_gate) { a.level[ge] = 1; a.inliers[ge] = 0; }
      else { a.level[ge] = 0; a.inliers[ge] = 1; my_good++; }
    }
    num_good = (int)(block_sum_d((double)my_good, red) + 0.5);
    if (round == max(1, a.n_rounds / 2)) robust = 0;
  }
  __syncthreads();
  for (int v = tid; v < nv; v += BA_THREADS) {
    double R[9];
    se3_R(est[v], R);
    double* T = a.poses + 12 * (size_t)(v0 + v);
    for (int r = 0; r < 3; ++r) { T[4 * r] = R[3 * r]; T[4 * r + 1] = R[3 * r + 1]; T[4 * r + 2] = R[3 * r + 2]; T[4 * r + 3] = est[v].t[r]; }
  }
  if (tid == 0 && a.stats) { a.stats[3 * prob] = rounds; a.stats[3 * prob + 1] = outer_total; a.stats[3 * prob + 2] = trials_total; }
}

// ---- one WARP per problem: graphs with a single free vertex and at most one fixed one -----------------------------------------
// BASELINE config 3 (512 objects, each its own LM problem against a fixed camera), single-object frames and the curr_only camera solve of
// SLAM mode (one camera vertex, unary edges) have ONE 6x6 block: a CTA per graph leaves 3 of 4 warps idle and pays ~12 block barriers per
// LM iteration.  Here the whole LM state of a problem (vertex estimate, backup, H, b, lambda) lives in the registers of one warp, edges are
// strided over the lanes, sums are shuffle trees (every lane ends up with the same value, so every lane takes the same branch), the 6x6
// solve runs on lane 0 and is broadcast: no shared memory, no barrier.  Same LM driver, same g2o semantics as ba_kernel (accept / reject rule,
// rejected-trial error left in the edges, rounds with chi2 re-classification, Huber strip); the summation order differs, i.e. results agree with
// ba_kernel and the oracle to FP64 rounding.
constexpr int BAW_THREADS = 128;

__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(BAW_THREADS)
ba_warp_kernel(const BaArgs a, int n_prob) {
  const int lane = threadIdx.x & 31;
  const int prob = blockIdx.x * (BAW_THREADS / 32) + (threadIdx.x >> 5);
  if (prob >= n_prob) return;
  const int v0 = a.prob_vert[prob], nv = a.vert_cnt ? a.vert_cnt[prob] : a.prob_vert[prob + 1] - v0;
  const int e0 = a.prob_edge[prob], ne = a.edge_cnt ? a.edge_cnt[prob] : a.prob_edge[prob + 1] - e0;
  auto fail = [&](int code) { if (lane == 0 && a.stats) { a.stats[3 * prob] = code; a.stats[3 * prob + 1] = 0; a.stats[3 * prob + 2] = 0; } };
  if (nv < 1 || nv > 2) { fail(-1); return; }
  const bool f0 = !a.fixed[v0], f1 = nv == 2 && !a.fixed[v0 + 1];
  if (f0 == f1) { fail(f0 ? -2 : -3); return; }          // two free vertices (coupled: ba_global.cu) / nothing to optimise
  const int vfree = f0 ? v0 : v0 + 1, vother = nv == 2 ? (f0 ? v0 + 1 : v0) : -1;
  auto load = [&](int v, SE3q& o) {
    const double* T = a.poses + 12 * (size_t)v;
    const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    const double t[3] = {T[3], T[7], T[11]};
    se3_from_Rt(R, t, o);
  };
  SE3q est, bak, other;
  load(vfree, est);
  if (vother >= 0) load(vother, other); else other = est;
  bak = est;
  // which vertex of each edge is free (0 = the object, 1 = the camera, -1 = none)
  int bad = 0;
  for (int e = lane; e < ne; e += 32) {
    const int ge = e0 + e, vo = a.e_obj[ge], vc = a.e_cam[ge];
    const bool in_range = (vo < 0 || vo == vfree || vo == vother) && (vc == vfree || vc == vother);
    int8_t k = -1;
    if (!in_range || (vo == vfree && vc == vfree)) bad = 1;
    else if (vo == vfree) k = 0;
    else if (vc == vfree) k = 1;
    a.fv_kind[ge] = k;
  }
  if (__any_sync(0xffffffffu, bad)) { fail(-2); return; }

  auto edge_error = [&](int ge) {   // computeError
    double pw[3] = {a.p[3 * ge], a.p[3 * ge + 1], a.p[3 * ge + 2]}, pc[3];
    if (a.e_obj[ge] >= 0) { double tmp[3]; se3_map(a.e_obj[ge] == vfree ? est : other, pw, tmp); pw[0] = tmp[0]; pw[1] = tmp[1]; pw[2] = tmp[2]; }
    se3_map(a.e_cam[ge] == vfree ? est : other, pw, pc);
    const double* k = a.cam_k + 4 * ge;
    a.err[2 * ge] = a.uv[2 * ge] - (k[0] * pc[0] / pc[2] + k[2]);
    a.err[2 * ge + 1] = a.uv[2 * ge + 1] - (k[1] * pc[1] / pc[2] + k[3]);
  };
  auto edge_chi2 = [&](int ge) {
    const double* O = a.info + 4 * ge;
    const double r0 = a.err[2 * ge], r1 = a.err[2 * ge + 1];
    return r0 * (O[0] * r0 + O[1] * r1) + r1 * (O[2] * r0 + O[3] * r1);
  };
  auto is_active = [&](int ge) { return a.level[ge] == 0 && a.fv_kind[ge] >= 0; };
  auto chi_sum = [&](int robust) {   // computeActiveErrors + activeRobustChi2
    double part = 0;
    for (int e = lane; e < ne; e += 32) {
      const int ge = e0 + e;
      if (!is_active(ge)) continue;
      edge_error(ge);
      const double c = edge_chi2(ge);
      if (robust) { double r0, r1; huber(c, a.huber_delta, r0, r1); part += r0; } else part += c;
    }
    return warp_sum(part);
  };

  // ---- initial chi2 classification (object_slam.py:848-866) ----
  int my_good = 0;
  for (int e = lane; e < ne; e += 32) {
    const int ge = e0 + e;
    if (a.init_with_outliers) { a.level[ge] = 0; a.err[2 * ge] = 0.0; a.err[2 * ge + 1] = 0.0; my_good++; }
    else {
      edge_error(ge);
      if (edge_chi2(ge) > a.chi2_gate) { a.level[ge] = 1; a.inliers[ge] = 0; }
      else { a.level[ge] = 0; a.inliers[ge] = 1; my_good++; }
    }
  }
  int num_good = warp_sum_i(my_good);
  int robust = 1, rounds = 0, outer_total = 0, trials_total = 0;

  for (int round = 0; round < a.n_rounds; ++round) {
    if (ne < 4 || num_good < 4) break;
    int any = 0;
    for (int e = lane; e < ne; e += 32) any |= is_active(e0 + e) ? 1 : 0;
    any = __any_sync(0xffffffffu, any);
    ++rounds;
    if (any) {
      bool ok = true;
      const int iters = a.its[round];
      double lambda = 0, ni = 2;
      for (int it = 0; it < iters && ok; ++it) {
        const double currentChi0 = chi_sum(robust);
        // buildSystem: b (6) and the upper triangle of H (21), lane-strided partial sums + shuffle tree
        double acc[27];
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[k] = 0;
        for (int e = lane; e < ne; e += 32) {
          const int ge = e0 + e;
          if (!is_active(ge)) continue;
          double pw[3] = {a.p[3 * ge], a.p[3 * ge + 1], a.p[3 * ge + 2]}, pc[3];
          if (a.e_obj[ge] >= 0) { double tmp[3]; se3_map(a.e_obj[ge] == vfree ? est : other, pw, tmp); pw[0] = tmp[0]; pw[1] = tmp[1]; pw[2] = tmp[2]; }
          const SE3q& Tcw = a.e_cam[ge] == vfree ? est : other;
          se3_map(Tcw, pw, pc);
          double Rcw[9], J[12], Jx[12];
          se3_R(Tcw, Rcw);
          const bool wrt_obj = a.fv_kind[ge] == 0;
          edge_jacobians(Rcw, pw, pc, a.cam_k + 4 * ge, wrt_obj, !wrt_obj, wrt_obj ? J : Jx, wrt_obj ? Jx : J);
          const double* O = a.info + 4 * ge;
          const double r0 = a.err[2 * ge], r1 = a.err[2 * ge + 1];
          double w = 1.0;
          if (robust) { double h0; huber(r0 * (O[0] * r0 + O[1] * r1) + r1 * (O[2] * r0 + O[3] * r1), a.huber_delta, h0, w); }
          const double or0 = -(O[0] * r0 + O[1] * r1) * w, or1 = -(O[2] * r0 + O[3] * r1) * w;
          const double w00 = O[0] * w, w01 = O[1] * w, w10 = O[2] * w, w11 = O[3] * w;
          int idx = 6;
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            acc[c] += J[c] * or0 + J[6 + c] * or1;
            const double a0 = J[c] * w00 + J[6 + c] * w10, a1 = J[c] * w01 + J[6 + c] * w11;
#pragma unroll
            for (int d = c; d < 6; ++d) acc[idx++] += a0 * J[d] + a1 * J[6 + d];
          }
        }
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[k] = warp_sum(acc[k]);
        if (it == 0) {   // computeLambdaInit
          double mx = 0;
          int idx = 6;
#pragma unroll
          for (int c = 0; c < 6; ++c) { mx = fmax(fabs(acc[idx]), mx); idx += 6 - c; }
          lambda = 1e-5 * mx; ni = 2.0;
        }
        double cur = currentChi0, rho = 0;
        int qmax = 0;
        bool lam_bad = false;
        do {
          bak = est;
          double x[6];
          int okv = 1;
          if (lane == 0) {
            double A[36], rhs[6];
            int idx = 6;
            for (int c = 0; c < 6; ++c) {
              rhs[c] = acc[c];
              for (int d = c; d < 6; ++d) { A[6 * c + d] = acc[idx]; A[6 * d + c] = acc[idx]; ++idx; }
            }
            for (int j = 0; j < 6; ++j) A[7 * j] += lambda;
            okv = chol6(A, rhs) ? 1 : 0;
            for (int j = 0; j < 6; ++j) x[j] = okv ? rhs[j] : 0.0;
          }
          okv = __shfl_sync(0xffffffffu, okv, 0);
#pragma unroll
          for (int j = 0; j < 6; ++j) x[j] = __shfl_sync(0xffffffffu, x[j], 0);
          SE3q nw;
          se3_oplus(est, x, nw);
          est = nw;
          double tempChi = chi_sum(robust);
          double scale = 0;
#pragma unroll
          for (int j = 0; j < 6; ++j) scale += x[j] * (lambda * x[j] + acc[j]);   // computeScale
          if (!okv) tempChi = 1.7976931348623157e308;
          const double r = (cur - tempChi) / (scale + 1e-3);
          int flag;
          if (r > 0 && isfinite(tempChi)) {
            double alpha = 1. - pow(2 * r - 1, 3.0);
            alpha = fmin(alpha, 2. / 3.);
            lambda *= fmax(1. / 3., alpha);
            ni = 2; cur = tempChi; flag = 1;
          } else {
            lambda *= ni; ni *= 2; flag = 0;
            if (!isfinite(lambda)) flag = 2;
          }
          rho = r;
          if (flag != 1) est = bak;      // pop(): restore the vertex; edge errors stay those of the rejected trial
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
    for (int e = lane; e < ne; e += 32) {
      const int ge = e0 + e;
      if (!a.inliers[ge]) edge_error(ge);
      if (edge_chi2(ge) > a.chi2_gate) { a.level[ge] = 1; a.inliers[ge] = 0; }
      else { a.level[ge] = 0; a.inliers[ge] = 1; my_good++; }
    }
    num_good = warp_sum_i(my_good);
    if (round == max(1, a.n_rounds / 2)) robust = 0;
  }
  if (lane == 0) {
    double R[9];
    se3_R(est, R);
    double* T = a.poses + 12 * (size_t)vfree;
    for (int r = 0; r < 3; ++r) { T[4 * r] = R[3 * r]; T[4 * r + 1] = R[3 * r + 1]; T[4 * r + 2] = R[3 * r + 2]; T[4 * r + 3] = est.t[r]; }
    if (vother >= 0) {      // (ba_kernel rewrites fixed vertices too: normalised through the quaternion)
      se3_R(other, R);
      T = a.poses + 12 * (size_t)vother;
      for (int r = 0; r < 3; ++r) { T[4 * r] = R[3 * r]; T[4 * r + 1] = R[3 * r + 1]; T[4 * r + 2] = R[3 * r + 2]; T[4 * r + 3] = other.t[r]; }
    }
    if (a.stats) { a.stats[3 * prob] = rounds; a.stats[3 * prob + 1] = outer_total; a.stats[3 * prob + 2] = trials_total; }
  }
}

// error and both Jacobians of independent edges at given vertex poses (suo_edge_linearize): the device functions the LM
// kernels use, exposed so that tests can compare them with central differences (the recipe left commented out in
// types_object_slam.cpp:108-122).
__global__ void edge_linearize_kernel(int n, const double* __restrict__ T_obj, const double* __restrict__ T_cam,
                                      const double* __restrict__ cam_k, const double* __restrict__ p, const double* __restrict__ uv,
                                      double* __restrict__ err, double* __restrict__ J_obj, double* __restrict__ J_cam) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  auto load = [](const double* T, SE3q& o) {
    const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    const double t[3] = {T[3], T[7], T[11]};
    se3_from_Rt(R, t, o);
  };
  SE3q To, Tc;
  load(T_cam + 12 * (size_t)e, Tc);
  double pw[3] = {p[3 * e], p[3 * e + 1], p[3 * e + 2]}, pc[3];
  if (T_obj) { load(T_obj + 12 * (size_t)e, To); double tmp[3]; se3_map(To, pw, tmp); pw[0] = tmp[0]; pw[1] = tmp[1]; pw[2] = tmp[2]; }
  se3_map(Tc, pw, pc);
  const double* k = cam_k + 4 * e;
  if (err) {
    err[2 * e] = uv[2 * e] - (k[0] * pc[0] / pc[2] + k[2]);
    err[2 * e + 1] = uv[2 * e + 1] - (k[1] * pc[1] / pc[2] + k[3]);
  }
  double Rcw[9], Ji[12], Jj[12];
  se3_R(Tc, Rcw);
  edge_jacobians(Rcw, pw, pc, k, T_obj != nullptr && J_obj != nullptr, J_cam != nullptr, Ji, Jj);
  if (T_obj && J_obj) for (int q = 0; q < 12; ++q) J_obj[12 * (size_t)e + q] = Ji[q];
  if (J_cam) for (int q = 0; q < 12; ++q) J_cam[12 * (size_t)e + q] = Jj[q];
}

}  // namespace

int launch_edge_linearize(suo_ctx* ctx, int n_edges, const double* T_obj, const double* T_cam, const double* cam_k, const double* p,
                          const double* uv, double* err, double* J_obj, double* J_cam, cudaStream_t s) {
  if (n_edges <= 0) return SUO_OK;
  edge_linearize_kernel<<<(n_edges + 127) / 128, 128, 0, s>>>(n_edges, T_obj, T_cam, cam_k, p, uv, err, J_obj, J_cam);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

int launch_ba_kernel(suo_ctx* ctx, int n_prob, const BaArgs& args, cudaStream_t s, int single_vertex) {
  if (n_prob <= 0) return SUO_OK;
  if (single_vertex) ba_warp_kernel<<<(n_prob + BAW_THREADS / 32 - 1) / (BAW_THREADS / 32), BAW_THREADS, 0, s>>>(args, n_prob);
  else ba_kernel<<<n_prob, BA_THREADS, 0, s>>>(args);
  ctx->launches++;
  SUO_CUDA_TRY(ctx, cudaGetLastError());
  return SUO_OK;
}

// Called by api.cu with scratch buffers sized for n_edges.
int launch_ba_batch_scratch(suo_ctx* ctx, int n_prob, const int32_t* prob_vert, const int32_t* prob_edge, double* poses,
                            const uint8_t* fixed, const int32_t* e_obj, const int32_t* e_cam, const double* cam_k,
                            const double* p, const double* uv, const double* info, uint8_t* inliers,
                            const int32_t* its, int n_rounds, double huber_delta, double chi2_gate,
                            int init_with_outliers, int32_t* stats, double* err_scratch, uint8_t* level_scratch,
                            int8_t* fv_scratch, cudaStream_t s, const int32_t* vert_cnt, const int32_t* edge_cnt, int single_vertex) {
  BaArgs a;
  a.prob_vert = prob_vert; a.prob_edge = prob_edge; a.vert_cnt = vert_cnt; a.edge_cnt = edge_cnt; a.poses = poses; a.fixed = fixed; a.e_obj = e_obj; a.e_cam = e_cam;
  a.cam_k = cam_k; a.p = p; a.uv = uv; a.info = info; a.inliers = inliers; a.its = its; a.n_rounds = n_rounds;
  a.huber_delta = huber_delta; a.chi2_gate = chi2_gate; a.init_with_outliers = init_with_outliers; a.stats = stats;
  a.err = err_scratch; a.level = level_scratch; a.fv_kind = fv_scratch;
  return launch_ba_kernel(ctx, n_prob, a, s, single_vertex);
}
