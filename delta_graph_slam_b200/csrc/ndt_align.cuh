// b200reg — NDT alignment: the B200 replacement of
// pclomp::NormalDistributionsTransform::computeTransformation and everything under it
// (computeDerivatives / updateDerivatives / computePointDerivatives / computeAngleDerivatives /
// computeStepLengthMT / trialValueSelectionMT / updateIntervalMT / computeHessian),
// reached through registration->align [REF apps/scan_matching_odometry_nodelet.cpp:218;
// include/hdl_graph_slam/loop_detector.hpp:145].
//
// One persistent cooperative kernel runs a whole registration: every derivative pass, the
// Newton solve and the More-Thuente line search happen on the device, so an iteration costs no
// host round trip.  A *group* of G CTAs owns one registration (G = all 148 SMs for a single
// odometry alignment, a handful for loop-closure batches where many groups run side by side):
//   stage  : the target's voxel hash + 48-byte voxel records are copied into shared memory once
//            per registration (they fit the 227 KB of an SM for HDL-64-sized scans; larger grids
//            are read through L2 instead)
//   pass   : each thread transforms its source points (float, same operation order as
//            pcl::transformPoint), probes the DIRECT1/7/26 (or KDTREE-radius) voxels and
//            accumulates score / gradient / Hessian; 32-point groups are dealt round-robin to
//            the group's warps so dense and empty regions of the scan are spread over all SMs
//   reduce : warp shuffles -> shared memory -> one 29-double partial per CTA in global memory
//   sync   : a counter barrier among the group's CTAs (co-residency guaranteed by the
//            cooperative launch), then EVERY CTA sums the partials in the same fixed order
//   step   : warp 0 of every CTA runs the Newton / More-Thuente state machine redundantly on the
//            identical totals — bit-identical decisions, so no broadcast and no second barrier.
//
// Per-hit arithmetic is float and the sums are double, as upstream.  The per-hit Hessian is
// factored through the point: with v = sum_hits e*(C q) and M = sum_hits e*(C - d2 (Cq)(Cq)^T),
//   g = J^T v,   H = J^T M J + [v . d2x/dpi dpj]
// which is the same polynomial as updateDerivatives' per-hit 6x6 update evaluated in a
// different (cheaper) order.
#pragma once
#include "../../include/b200reg.h"
#include "ndt_grid.cuh"
#include "small_solve.cuh"

namespace b200 {

constexpr int kAlignThreads = 512;
constexpr int kAlignWarps = kAlignThreads / 32;
constexpr int kNumAcc = 29;  // score, g[6], H upper triangle [21], hits
constexpr int kAccStride = 32;
constexpr int kStageBytes = 192 * 1024;  // dynamic shared memory for the staged grid
constexpr int kLaneGroup = 4;             // points a lane sums in float before the warp reduction (batch mode)

enum NdtPhase { PH_INIT = 0, PH_MT_FIRST = 1, PH_MT_ITER = 2, PH_DONE = 4, PH_EVAL_ONLY = 5 };

struct NdtParams {
  int search;  // b200reg_nn_search
  double resolution, step_size, outlier_ratio, trans_eps;
  int max_iterations;
};

struct NdtJob {
  const float4* src;
  int n_src;
  NdtGridView grid;
  float guess[12];  // row-major 3x4 of the initial guess
  double p0[6];     // [t, eulerXYZ] of the guess (host: Eigen eulerAngles(0,1,2) restated, A.6)
  int eval_only;    // 1: a single derivative pass at p0 (introspection), T built from p0
  b200reg_result* result;  // device
  // single registrations: a second copy of the record goes straight to page-locked host memory mapped
  // into the device address space, followed by a completion flag the host polls — no D2H copy, no
  // stream synchronisation on the critical path of a frame
  b200reg_result* result_host;
  unsigned int* done_flag;
  unsigned int done_seq;
  double* deriv_out;       // device, 1 + 6 + 36 doubles (eval_only)
  long long* prof;         // device, optional: SM cycles of CTA 0 per phase {pass, reduce, barrier, total, step, n, stage}
  double* trace;           // device, optional (profiled instantiation): kTraceDoubles per pass, at most kTraceCap passes; trace[-1 record] holds the count
};
constexpr int kTraceDoubles = 13, kTraceCap = 255;

// Full loop batches: maximal runs of consecutive jobs that share a target, and one counter per run.
struct NdtTargetQueue {
  const uint2* runs;   // (first job, number of jobs); nullptr = plain tickets, one job each
  unsigned int* next;  // per run: jobs handed out so far (zeroed by the host before the launch)
  int n_runs;
};

struct NdtShared {
  // inputs of the current pass
  float T[12];
  float j_ang[8][3];
  float h_ang[15][3];
  int need_hessian;
  int phase;
  double neg_g[6], dp[6];
  int new_pose;  // the step asked for an evaluation at a new x_t (T / tables must be rebuilt)
  // optimiser state (identical in every CTA of the group)
  double p[6], x_t[6], dir[6];
  double score, g[6], H[36];
  double phi_0, d_phi_0, a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t, psi_t, d_psi_t;
  int step_iterations, open_interval, interval_converged;
  int nr_iterations, converged, n_eval, n_pass;
  double hits;
  double gauss_d1, gauss_d2;
  double trig[6][2];  // sin, cos of: float-rounded angles (T) [0..2], double angles (tables) [3..5]
  double tot[kAccStride];
  int solve_path;  // developer trace: how the last H dp = -g was solved (0 none, 1 block closed form, 2 pivoted elimination, 3 SVD)
  long long pf[6];  // developer cycle counters inside ndt_step: {totals -> state, line-search update, trial value, Newton end, 6x6 solve, Newton begin rest}
  double red[kAlignWarps][kAccStride];
};

// staged (or global) view of the target grid used inside a pass.  When the grid sits in shared
// memory the loads are issued as explicit ld.shared on 32-bit shared-window addresses: through the
// generic pointers the compiler could not prove the address space and emitted generic LDs, which
// cost an address-space check and sit on the long scoreboard (ncu: the top stall of the pass).
struct NdtLookup {
  const uint2* table;
  uint32_t mask;
  const NdtVoxel* voxels;
  const float4* centroids;
  uint32_t s_table, s_voxels, s_centroids;  // shared-window byte addresses (staged grids)
};

// The hash is bucketised: two (key, record) entries share a 16-byte bucket, insertion fills entry 0,
// then entry 1, then moves on to the next bucket.  A look-up is ONE 16-byte load and two compares for
// all but the few keys whose bucket overflowed (a second entry left empty proves a miss), instead of
// a data-dependent walk over 8-byte entries that diverged for every fourth look-up.
template <bool STAGED>
__device__ __forceinline__ uint4 lk_bucket(const NdtLookup& g, uint32_t b) {
  if (STAGED) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(g.s_table + b * 16u));
    return v;
  }
  return reinterpret_cast<const uint4*>(g.table)[b];
}
// record of `key` (flags included) or -1, starting from an already loaded first bucket
template <bool STAGED>
__device__ __forceinline__ int ndt_resolve(const NdtLookup& g, uint32_t key, uint32_t b, uint4 e) {
  while (true) {
    if (e.x == key) return (int)e.y;
    if (e.z == key) return (int)e.w;
    if (e.z == kInvalidKey) return -1;
    b = (b + 1) & g.mask;
    e = lk_bucket<STAGED>(g, b);
  }
}
template <bool STAGED>
__device__ __forceinline__ float4 lk_f4(const NdtLookup& g, uint32_t s_base, const float4* p_base, uint32_t index16) {
  if (STAGED) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(s_base + index16 * 16u));
    return v;
  }
  return p_base[index16];
}

template <bool STAGED>
__device__ __forceinline__ int ndt_lookup(const NdtLookup& g, uint32_t key) {
  const uint32_t b = ndt_hash(key, g.mask);
  return ndt_resolve<STAGED>(g, key, b, lk_bucket<STAGED>(g, b));
}

// ---- the pose -> transform / angle-derivative tables ---------------------------------------
// T = Translation(p0..2) * Rx(p3) * Ry(p4) * Rz(p5) composed in float exactly like the oracle's
// m4f_from_xyz_euler (sequence of float 4x4 products with exact zeros and ones collapsed).
// trig[0..2] = {sin, cos} of the float-rounded angles, evaluated in double.
__device__ inline void pose_to_T(const double p[6], const double trig[6][2], float T[12]) {
  const float sx = (float)trig[0][0], cx = (float)trig[0][1];
  const float sy = (float)trig[1][0], cy = (float)trig[1][1];
  const float sz = (float)trig[2][0], cz = (float)trig[2][1];
  const float b10 = __fmul_rn(sx, sy), b12 = -__fmul_rn(sx, cy), b20 = -__fmul_rn(cx, sy), b22 = __fmul_rn(cx, cy);
  T[0] = __fmul_rn(cy, cz); T[1] = -__fmul_rn(cy, sz); T[2] = sy; T[3] = (float)p[0];
  T[4] = __fadd_rn(__fmul_rn(b10, cz), __fmul_rn(cx, sz));
  T[5] = __fadd_rn(__fmul_rn(b10, -sz), __fmul_rn(cx, cz));
  T[6] = b12; T[7] = (float)p[1];
  T[8] = __fadd_rn(__fmul_rn(b20, cz), __fmul_rn(sx, sz));
  T[9] = __fadd_rn(__fmul_rn(b20, -sz), __fmul_rn(sx, cz));
  T[10] = b22; T[11] = (float)p[2];
}

// computeAngleDerivatives (A.4): rows of j_ang (8) and h_ang (15) as float; evaluated in double
// without contraction, sharing sub-products in the order the reference forms them
// ((a*b)*c), so the float casts match a CPU evaluation of the original expressions.
// trig[3..5] = {sin, cos} of p[3..5] with the |angle| < 10e-5 -> (0, 1) shortcut applied.
__device__ inline void angle_tables(const double trig[6][2], float j_ang[8][3], float h_ang[15][3]) {
  const double sx = trig[3][0], cx = trig[3][1], sy = trig[4][0], cy = trig[4][1], sz = trig[5][0], cz = trig[5][1];
#define MU(a, b) __dmul_rn(a, b)
#define AD(a, b) __dadd_rn(a, b)
#define SB(a, b) __dsub_rn(a, b)
  const double cxsy = MU(cx, sy), sxsy = MU(sx, sy), cxcy = MU(cx, cy), sxcy = MU(sx, cy);
  const double cxsycz = MU(cxsy, cz), cxsysz = MU(cxsy, sz), sxsycz = MU(sxsy, cz), sxsysz = MU(sxsy, sz);
  const double cxcycz = MU(cxcy, cz), cxcysz = MU(cxcy, sz), sxcycz = MU(sxcy, cz), sxcysz = MU(sxcy, sz);
  const double sxsz = MU(sx, sz), sxcz = MU(sx, cz), cxsz = MU(cx, sz), cxcz = MU(cx, cz);
  const double sycz = MU(sy, cz), sysz = MU(sy, sz), cysz = MU(cy, sz), cycz = MU(cy, cz);
  const double j[8][3] = {{AD(-sxsz, cxsycz), SB(-sxcz, cxsysz), -cxcy},
                          {AD(cxsz, sxsycz), SB(cxcz, sxsysz), -sxcy},
                          {-sycz, sysz, cy},
                          {sxcycz, -sxcysz, sxsy},
                          {-cxcycz, cxcysz, -cxsy},
                          {-cysz, -cycz, 0.0},
                          {SB(cxcz, sxsysz), SB(-cxsz, sxsycz), 0.0},
                          {AD(sxcz, cxsysz), SB(cxsycz, sxsz), 0.0}};
  const double h[15][3] = {{SB(-cxsz, sxsycz), AD(-cxcz, sxsysz), sxcy},
                           {AD(-sxsz, cxsycz), SB(-cxsysz, sxcz), -cxcy},
                           {cxcycz, -cxcysz, cxsy},
                           {sxcycz, -sxcysz, sxsy},
                           {SB(-sxcz, cxsysz), SB(sxsz, cxsycz), 0.0},
                           {SB(cxcz, sxsysz), SB(-sxsycz, cxsz), 0.0},
                           {-cycz, cysz, sy},  // d1: upstream's (+sy), kept
                           {-sxsycz, sxsysz, sxcy},
                           {cxsycz, -cxsysz, -cxcy},
                           {sysz, sycz, 0.0},
                           {-sxcysz, -sxcycz, 0.0},
                           {cxcysz, cxcycz, 0.0},
                           {-cycz, cysz, 0.0},
                           {SB(-cxsz, sxsycz), AD(-cxcz, sxsysz), 0.0},
                           {AD(-sxsz, cxsycz), SB(-cxsysz, sxcz), 0.0}};
#undef MU
#undef AD
#undef SB
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) j_ang[r][c] = (float)j[r][c];
#pragma unroll
  for (int r = 0; r < 15; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) h_ang[r][c] = (float)h[r][c];
}

// (A warp-parallel form of angle_tables — pair / triple products on twelve / eight lanes, every lane combining three
// entries from a shared-memory table — was built, verified bit-identical, measured and dropped: 1.45 k cycles against the
// 0.85 k of one lane walking the expressions, the dependent shared-memory round trips between its stages cost more than
// the ~170 independent double operations they distribute.)
// the six sines and six cosines of a pose, one per lane
__device__ __forceinline__ void pose_trig_warp(NdtShared& s, const double x[6], int lane) {
  // six lanes, one sincos each.  (One sine OR cosine per lane on twelve lanes was measured and is slower: the two
  // code paths diverge inside the warp and run one after the other, 2.0 k cycles against 0.85 k.)
  if (lane < 6) {
    const int a = lane < 3 ? lane : lane - 3;
    double ang = x[3 + a], sn, cs;
    if (lane < 3) {
      ang = (double)(float)ang;
      sincos(ang, &sn, &cs);
    } else if (fabs(ang) < 10e-5) {
      sn = 0.0; cs = 1.0;
    } else {
      sincos(ang, &sn, &cs);
    }
    s.trig[lane][0] = sn;
    s.trig[lane][1] = cs;
  }
  __syncwarp();
}

// ---- More-Thuente helpers (A.4) ---------------------------------------------------------------
__device__ inline double mt_psi(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
__device__ inline double mt_dpsi(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

__device__ inline bool mt_update_interval(double& a_l, double& f_l, double& g_l, double& a_u, double& f_u, double& g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
  if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  return true;
}

__device__ inline double mt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    return fabs(a_c - a_l) < fabs(a_q - a_l) ? a_c : 0.5 * (a_q + a_c);
  }
  if (g_t * g_l < 0) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    return fabs(a_c - a_t) >= fabs(a_s - a_t) ? a_c : a_s;
  }
  if (fabs(g_t) <= fabs(g_l)) {
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    double a_n = fabs(a_c - a_t) < fabs(a_s - a_t) ? a_c : a_s;
    return a_t > a_l ? fmin(a_t + 0.66 * (a_u - a_t), a_n) : fmax(a_t + 0.66 * (a_u - a_t), a_n);
  }
  double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
  double w = sqrt(z * z - g_t * g_u);
  return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
}

// H upper-triangle index of (i, j), i <= j, inside the accumulator vector
__host__ __device__ constexpr int hidx(int i, int j) { return 7 + i * 6 - (i * (i - 1)) / 2 + (j - i); }

// ---- the optimiser state machine; one lane, after every derivative pass ----------------------
// Leaves the next phase in s.phase (PH_DONE when finished) and sets s.new_pose when the next pass
// must be evaluated at s.x_t (the caller then rebuilds T and the angle tables).
// The optimiser step, run by ALL 32 lanes of warp 0 after every derivative pass.
//
// The state machine of computeTransformation / computeStepLengthMT (A.4) used to run on one lane with its state in
// shared memory: ~8.7 k cycles of dependent shared-memory loads, stores and double-precision operations per pass
// (4.4 us of a 12.4 us pass) while 107 SMs x 16 warps waited at the next barrier.  Here the state is loaded into
// registers once, every lane executes the (warp-uniform) scalar control flow redundantly, the 6-vectors p, dir, g, x_t live
// one component per lane (lane i < 6; dot products are three-step butterflies), the Hessian one ROW per lane — which is the
// layout the warp-level elimination wants, so the 6x6 solve starts without touching memory — and the state goes back to
// shared memory once at the end.  Leaves the next phase in s.phase (PH_DONE when finished) and sets s.new_pose when the
// next pass must be evaluated at s.x_t.
__device__ __forceinline__ double warp_dot6(double a, double b, int lane) {
  double v = lane < 6 ? a * b : 0.0;
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return __shfl_sync(0xffffffffu, v, 0);  // one value for the whole warp (lanes 8.. hold other butterflies)
}

#ifndef B200_STEP_ATTR
#define B200_STEP_ATTR __forceinline__
#endif
template <bool PROF>
static __device__ B200_STEP_ATTR void ndt_step_warp(NdtShared& s, const NdtParams& prm, int lane) {
  const double mu = 1.e-4, nu = 0.9;
  const double step_max = prm.step_size, step_min = prm.trans_eps / 2;
  long long tk = PROF ? clock64() : 0;
  auto lap = [&](int slot) {
    if (PROF) { const long long now = clock64(); if (lane == 0) s.pf[slot] += now - tk; tk = now; }
  };
  const int li = lane < 6 ? lane : 0;
  // ---- state -> registers (independent shared-memory loads, all in flight together)
  const double* t = s.tot;
  const double score = t[0];
  const double g_i = lane < 6 ? t[1 + li] : 0.0;
  const double hits_pass = t[28];
  double p_i = s.p[li], dir_i = s.dir[li], x_t_i = 0.0;
  double phi_0 = s.phi_0, d_phi_0 = s.d_phi_0, a_l = s.a_l, f_l = s.f_l, g_l = s.g_l, a_u = s.a_u, f_u = s.f_u, g_u = s.g_u, a_t = s.a_t;
  int phase = s.phase, step_iterations = s.step_iterations, open_interval = s.open_interval, interval_converged = s.interval_converged;
  int nr_iterations = s.nr_iterations, converged = s.converged, n_eval = s.n_eval + 1;
  double phi_t = 0.0, d_phi_t = 0.0, psi_t = 0.0, d_psi_t = 0.0;
  int new_pose = 0;
  bool go_newton_begin = phase == PH_INIT, go_loop_check = phase != PH_INIT, go_newton_end = false;
  lap(0);
  if (go_loop_check) {
    // the evaluation at a_t just finished.  A trial pass (PH_MT_ITER) evaluated the Hessian as well: upstream ends the
    // line search with computeHessian(x_t) at the LAST trial point — a whole extra pass over the cloud at the pose that
    // was just evaluated.  Here that Hessian is already in the totals of the last trial pass (same transform, same
    // tables, same arithmetic as a pass of its own), so the extra pass never runs.
    phi_t = -score;
    d_phi_t = -warp_dot6(g_i, dir_i, lane);
    psi_t = mt_psi(a_t, phi_t, phi_0, d_phi_0, mu);
    d_psi_t = mt_dpsi(d_phi_t, d_phi_0, mu);
    if (phase == PH_MT_ITER) {
      if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
        open_interval = 0;
        f_l = f_l + phi_0 - mu * d_phi_0 * a_l;
        g_l = g_l + mu * d_phi_0;
        f_u = f_u + phi_0 - mu * d_phi_0 * a_u;
        g_u = g_u + mu * d_phi_0;
      }
      if (open_interval) interval_converged = mt_update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
      else interval_converged = mt_update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
      step_iterations++;
    }
  }
  lap(1);
  while (true) {  // warp-uniform: every lane holds the same scalars
    if (go_loop_check) {
      go_loop_check = false;
      if (!interval_converged && step_iterations < 10 && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
        if (open_interval) a_t = mt_trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
        else a_t = mt_trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
        a_t = fmin(a_t, step_max);
        a_t = fmax(a_t, step_min);
        x_t_i = p_i + dir_i * a_t;
        phase = PH_MT_ITER;
        new_pose = 1;
        lap(2);
        break;
      }
      // computeHessian at x_t: taken from the last trial pass; it counts as an evaluation of the reference's
      // algorithm, not as a pass of ours
      if (step_iterations) n_eval++;
      go_newton_end = true;
    }
    if (go_newton_end) {
      go_newton_end = false;
      p_i += dir_i * a_t;
      if (nr_iterations > prm.max_iterations || (nr_iterations && (fabs(a_t) < prm.trans_eps))) converged = 1;
      nr_iterations++;
      if (converged) { phase = PH_DONE; break; }
      go_newton_begin = true;
      lap(3);
    }
    if (go_newton_begin) {
      go_newton_begin = false;
      // H dp = -g.  Fast path: every lane solves the system by 3x3 blocks in closed form from the totals (broadcast
      // shared-memory loads, no exchange between lanes) and verifies the residual; if that refuses (degenerate or strongly
      // indefinite Hessian) the pivoted elimination runs on the warp, and if a pivot collapses there too, JacobiSVD's
      // minimum-norm answer on one lane (see small_solve.cuh).
      double dp_i = 0.0;
      {
        double x[6];
        bool solved = solve6_schur(t + 7, t + 1, x);
        int path = 1;
        if (!solved) {  // warp-uniform: every lane computed the same bits
          path = 2;
          double r[7];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int a = li < j ? li : j, b = li < j ? j : li;
            r[j] = lane < 6 ? t[7 + a * 6 - (a * (a - 1)) / 2 + (b - a)] : 0.0;
          }
          r[6] = -g_i;
          solved = warp_solve6_rows(r, x, lane);
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) dp_i = lane == j ? x[j] : dp_i;
        if (!solved) {
          path = 3;
          if (lane < 6) {
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const int a = li < j ? li : j, b = li < j ? j : li;
              s.H[6 * lane + j] = t[7 + a * 6 - (a * (a - 1)) / 2 + (b - a)];
            }
            s.neg_g[lane] = -g_i;
          }
          __syncwarp();
          if (lane == 0) svd_solve6(s.H, s.neg_g, s.dp);
          __syncwarp();
          dp_i = lane < 6 ? s.dp[li] : 0.0;
        }
        if (PROF && lane == 0) s.solve_path = path;
      }
      lap(4);
      const double nrm = sqrt(warp_dot6(dp_i, dp_i, lane));
      if (nrm == 0 || nrm != nrm) {
        converged = (nrm == nrm) ? 1 : 0;
        phase = PH_DONE;
        break;
      }
      const double inv_nrm = 1.0 / nrm;
      dir_i = dp_i * inv_nrm;
      // computeStepLengthMT prologue
      phi_0 = -score;
      d_phi_0 = -warp_dot6(g_i, dir_i, lane);
      if (d_phi_0 >= 0) {
        if (d_phi_0 == 0) {  // returns 0 without evaluating anything
          a_t = 0;
          go_newton_end = true;
          continue;
        }
        d_phi_0 *= -1;
        dir_i *= -1;
      }
      step_iterations = 0;
      a_l = 0; a_u = 0;
      f_l = mt_psi(a_l, phi_0, phi_0, d_phi_0, mu);
      g_l = mt_dpsi(d_phi_0, d_phi_0, mu);
      f_u = mt_psi(a_u, phi_0, phi_0, d_phi_0, mu);
      g_u = mt_dpsi(d_phi_0, d_phi_0, mu);
      interval_converged = (step_max - step_min) < 0;
      open_interval = 1;
      a_t = nrm;
      a_t = fmin(a_t, step_max);
      a_t = fmax(a_t, step_min);
      x_t_i = p_i + dir_i * a_t;
      phase = PH_MT_FIRST;
      new_pose = 1;
      lap(5);
      break;
    }
  }
  // ---- registers -> state
  if (lane < 6) { s.p[lane] = p_i; s.dir[lane] = dir_i; s.x_t[lane] = x_t_i; }
  if (lane == 0) {
    s.score = score;
    s.phi_0 = phi_0; s.d_phi_0 = d_phi_0; s.a_l = a_l; s.f_l = f_l; s.g_l = g_l; s.a_u = a_u; s.f_u = f_u; s.g_u = g_u; s.a_t = a_t;
    s.phase = phase; s.step_iterations = step_iterations; s.open_interval = open_interval; s.interval_converged = interval_converged;
    s.nr_iterations = nr_iterations; s.converged = converged; s.n_eval = n_eval; s.n_pass = s.n_pass + 1; s.hits += hits_pass;
    s.new_pose = new_pose;
    s.need_hessian = 1;
    if (PROF) { s.phi_t = phi_t; s.d_phi_t = d_phi_t; s.psi_t = psi_t; s.d_psi_t = d_psi_t; }
  }
  __syncwarp();
}

// ---- one source point ------------------------------------------------------------------------
// ADDS the point's 29 contributions (score, g[6], H upper triangle [21], hits) as floats to
// out[0..28]; the caller reduces them across the warp and accumulates in double.
// gauss_d1 arrives as an unevaluated float pair (d1h + d1l) so that the reference's
// float(double(d1) * e) products are reproduced without FP64 conversions in the per-hit path.
template <int MODE, bool STAGED>  // MODE: 1, 7, 27 (DIRECT*) or 0 (KDTREE radius search over voxel centroids)
__device__ __forceinline__ void ndt_point(const NdtShared& s, const NdtLookup& grid, const GridParams& gp, float4 pt, float d1h, float d1l, float gauss_d2, float res2,
                                          int need_hessian, float out[32]) {
  const float x0 = pt.x, x1 = pt.y, x2 = pt.z;
  const float xt0 = affine_row(s.T[0], s.T[1], s.T[2], s.T[3], x0, x1, x2);
  const float xt1 = affine_row(s.T[4], s.T[5], s.T[6], s.T[7], x0, x1, x2);
  const float xt2 = affine_row(s.T[8], s.T[9], s.T[10], s.T[11], x0, x1, x2);
  // getNeighborhoodAtPoint*: float DIVISION by the leaf size (A.3)
  const int c0 = (int)floorf(__fdiv_rn(xt0, gp.leaf[0])), c1 = (int)floorf(__fdiv_rn(xt1, gp.leaf[1])), c2 = (int)floorf(__fdiv_rn(xt2, gp.leaf[2]));
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;                                       // sum e (C q)
  float m00 = 0.f, m01 = 0.f, m02 = 0.f, m11 = 0.f, m12 = 0.f, m22 = 0.f;  // sum e (C - d2 Cq Cq^T)
  float score = 0.f;
  int hits = 0;
  constexpr int NOFF = MODE == 0 ? 27 : MODE;
  // DIRECT1 / DIRECT7: the keys and FIRST hash probes of all neighbours are issued before any of them
  // is consumed, so the shared-memory loads of the seven look-ups are in flight together instead of
  // seven dependent probe loops back to back (a pass has ~1 point per thread: latency is the cost).
  constexpr bool kBatchProbes = NOFF <= 7;
  int pre_slot[kBatchProbes ? NOFF : 1];
  if (kBatchProbes) {
    uint32_t pkey[NOFF], ph[NOFF];
    uint4 pe[NOFF];
    bool pin[NOFF];
    // the seven keys are the centre key +- one stride, and a neighbour is inside the lattice when its own axis
    // coordinate is (the other two are the centre's): one key and nine range tests instead of seven of each
    const int r0 = c0 - gp.min_b[0], r1 = c1 - gp.min_b[1], r2 = c2 - gp.min_b[2];
    const uint32_t kc = (uint32_t)(r0 * gp.mul[0] + r1 * gp.mul[1] + r2 * gp.mul[2]);
    const int e0 = gp.max_b[0] - gp.min_b[0], e1 = gp.max_b[1] - gp.min_b[1], e2 = gp.max_b[2] - gp.min_b[2];
    const bool x_c = r0 >= 0 && r0 <= e0, y_c = r1 >= 0 && r1 <= e1, z_c = r2 >= 0 && r2 <= e2;
#pragma unroll
    for (int o = 0; o < NOFF; ++o) {
      const int dx = NOFF == 1 ? 0 : (o == 1 ? 1 : o == 2 ? -1 : 0), dy = NOFF == 1 ? 0 : (o == 3 ? 1 : o == 4 ? -1 : 0), dz = NOFF == 1 ? 0 : (o == 5 ? 1 : o == 6 ? -1 : 0);
      const bool x_ok = dx == 0 ? x_c : (r0 + dx >= 0 && r0 + dx <= e0);
      const bool y_ok = dy == 0 ? y_c : (r1 + dy >= 0 && r1 + dy <= e1);
      const bool z_ok = dz == 0 ? z_c : (r2 + dz >= 0 && r2 + dz <= e2);
      pin[o] = x_ok && y_ok && z_ok;
      pkey[o] = kc + (uint32_t)(dx * gp.mul[0] + dy * gp.mul[1] + dz * gp.mul[2]);
      ph[o] = ndt_hash(pkey[o], grid.mask);
      pe[o] = pin[o] ? lk_bucket<STAGED>(grid, ph[o]) : make_uint4(kInvalidKey, 0u, kInvalidKey, 0u);
    }
#pragma unroll
    for (int o = 0; o < NOFF; ++o) {
      // both entries of the first bucket are compared without a branch; only an overflowed bucket walks on
      int sl = pe[o].x == pkey[o] ? (int)pe[o].y : (pe[o].z == pkey[o] ? (int)pe[o].w : -1);
      if (pin[o] && sl < 0 && pe[o].z != kInvalidKey) sl = ndt_resolve<STAGED>(grid, pkey[o], ph[o], pe[o]);
      pre_slot[o] = pin[o] ? sl : -1;
    }
  }
#pragma unroll
  for (int o = 0; o < NOFF; ++o) {
    int slot;
    if (kBatchProbes) {
      slot = pre_slot[o];
    } else {
      const int dx = o / 9 - 1, dy = (o / 3) % 3 - 1, dz = o % 3 - 1;
      const int i0 = c0 + dx, i1 = c1 + dy, i2 = c2 + dz;
      if (i0 < gp.min_b[0] || i0 > gp.max_b[0] || i1 < gp.min_b[1] || i1 > gp.max_b[1] || i2 < gp.min_b[2] || i2 > gp.max_b[2]) continue;
      const uint32_t key = (uint32_t)((i0 - gp.min_b[0]) * gp.mul[0] + (i1 - gp.min_b[1]) * gp.mul[1] + (i2 - gp.min_b[2]) * gp.mul[2]);
      slot = ndt_lookup<STAGED>(grid, key);
    }
    // DIRECT1 / DIRECT7 run branch-free: a neighbour without a usable voxel still goes through the
    // arithmetic on record 0 and every update is a select.  At warp level the body executed for all
    // seven neighbours anyway (some lane always has a hit); without the divergence bookkeeping
    // (BSSY / BSYNC / BRA were 14 % of the instructions) the seven bodies also overlap.
    bool valid = true;
    if (kBatchProbes) {
      valid = slot >= 0 && !(slot & (int)kNdtRejected);  // nr_points = -1 upstream: invisible to DIRECT*
      slot = valid ? slot : 0;
    } else {
      if (slot < 0) continue;
      if (slot & (int)kNdtRejected) {  // ... but still in the KDTREE cloud
        if (MODE != 0) continue;
        slot &= ~(int)kNdtRejected;
      }
      if (MODE == 0) {  // radiusSearch over the voxel centroids, d2 < resolution^2
        const float4 c = lk_f4<STAGED>(grid, grid.s_centroids, grid.centroids, (uint32_t)slot);
        if (!(l2_simple(xt0, xt1, xt2, c.x, c.y, c.z) < res2)) continue;
      }
    }
    const float4* vp = reinterpret_cast<const float4*>(grid.voxels);
    const uint32_t v16 = (uint32_t)slot * 3u;  // 48-byte records = 3 x 16 bytes
    const float4 ra = lk_f4<STAGED>(grid, grid.s_voxels, vp, v16), rb = lk_f4<STAGED>(grid, grid.s_voxels, vp, v16 + 1u),
                 rc = lk_f4<STAGED>(grid, grid.s_voxels, vp, v16 + 2u);  // hi0 hi1 hi2 lo0 | lo1 lo2 C00 C01 | C02 C11 C12 C22
    const float C00 = rb.z, C01 = rb.w, C02 = rc.x, C11 = rc.y, C12 = rc.z, C22 = rc.w;
    hits += valid ? 1 : 0;
    // q = float(double(x') - mean): (x' - hi) is exact for points within a few voxels of the mean
    const float q0 = __fsub_rn(__fsub_rn(xt0, ra.x), ra.w), q1 = __fsub_rn(__fsub_rn(xt1, ra.y), rb.x), q2 = __fsub_rn(__fsub_rn(xt2, ra.z), rb.y);
    const float cq0 = __fadd_rn(__fadd_rn(__fmul_rn(q0, C00), __fmul_rn(q1, C01)), __fmul_rn(q2, C02));
    const float cq1 = __fadd_rn(__fadd_rn(__fmul_rn(q0, C01), __fmul_rn(q1, C11)), __fmul_rn(q2, C12));
    const float cq2 = __fadd_rn(__fadd_rn(__fmul_rn(q0, C02), __fmul_rn(q1, C12)), __fmul_rn(q2, C22));
    const float qCq = __fadd_rn(__fadd_rn(__fmul_rn(q0, cq0), __fmul_rn(q1, cq1)), __fmul_rn(q2, cq2));
    const float ex = expf(__fmul_rn(__fmul_rn(-gauss_d2, qCq), 0.5f));
    // score_inc = float(-d1 * e) with d1 = d1h + d1l
    const float t1 = __fmul_rn(d1h, ex);
    const float sinc = __fadd_rn(t1, fmaf(d1l, ex, fmaf(d1h, ex, -t1)));
    float e = __fmul_rn(gauss_d2, ex);
    const bool ok = valid && !(e > 1.f || e < 0.f || e != e);  // upstream returns 0 for such a hit (score included)
    if (!kBatchProbes && !ok) continue;
    // e = float(e * d1)
    const float t2 = __fmul_rn(d1h, e);
    e = __fadd_rn(t2, fmaf(d1l, e, fmaf(d1h, e, -t2)));
    // a hit that does not count contributes with weight zero: two selects instead of one per accumulator (the voxel
    // record read for it is record 0, a real voxel, so every product below is finite and x + 0 * y == x)
    const float em = ok ? e : 0.f;
    score = __fsub_rn(score, ok ? sinc : 0.f);
    v0 = fmaf(em, cq0, v0); v1 = fmaf(em, cq1, v1); v2 = fmaf(em, cq2, v2);
    if (need_hessian) {
      const float ed = __fmul_rn(em, gauss_d2);
      m00 = m00 + (em * C00 - ed * cq0 * cq0); m01 = m01 + (em * C01 - ed * cq0 * cq1); m02 = m02 + (em * C02 - ed * cq0 * cq2);
      m11 = m11 + (em * C11 - ed * cq1 * cq1); m12 = m12 + (em * C12 - ed * cq1 * cq2); m22 = m22 + (em * C22 - ed * cq2 * cq2);
    }
  }
  if (!hits) return;
  out[28] += (float)hits;
  out[0] += score;
  // point_gradient columns 3..5: j3 = (0, a.x, b.x), j4 = (c.x, d.x, e.x), j5 = (f.x, g.x, h.x)
  auto dotj = [&](int r) { return s.j_ang[r][0] * x0 + s.j_ang[r][1] * x1 + s.j_ang[r][2] * x2; };
  const float j3y = dotj(0), j3z = dotj(1);
  const float j4x = dotj(2), j4y = dotj(3), j4z = dotj(4);
  const float j5x = dotj(5), j5y = dotj(6), j5z = dotj(7);
  out[1] += v0; out[2] += v1; out[3] += v2;
  out[4] += v1 * j3y + v2 * j3z;
  out[5] += v0 * j4x + v1 * j4y + v2 * j4z;
  out[6] += v0 * j5x + v1 * j5y + v2 * j5z;
  if (!need_hessian) return;
  // M J_k for the rotational columns
  const float a0 = m01 * j3y + m02 * j3z, a1 = m11 * j3y + m12 * j3z, a2 = m12 * j3y + m22 * j3z;                          // M j3
  const float b0 = m00 * j4x + m01 * j4y + m02 * j4z, b1 = m01 * j4x + m11 * j4y + m12 * j4z, b2 = m02 * j4x + m12 * j4y + m22 * j4z;  // M j4
  const float e0 = m00 * j5x + m01 * j5y + m02 * j5z, e1 = m01 * j5x + m11 * j5y + m12 * j5z, e2 = m02 * j5x + m12 * j5y + m22 * j5z;  // M j5
  auto doth = [&](int r) { return s.h_ang[r][0] * x0 + s.h_ang[r][1] * x1 + s.h_ang[r][2] * x2; };
  // second derivatives: a=(0,h0,h1) b=(0,h2,h3) c=(0,h4,h5) d=(h6,h7,h8) e=(h9,h10,h11) f=(h12,h13,h14)
  const float vh33 = v1 * doth(0) + v2 * doth(1);
  const float vh34 = v1 * doth(2) + v2 * doth(3);
  const float vh35 = v1 * doth(4) + v2 * doth(5);
  const float vh44 = v0 * doth(6) + v1 * doth(7) + v2 * doth(8);
  const float vh45 = v0 * doth(9) + v1 * doth(10) + v2 * doth(11);
  const float vh55 = v0 * doth(12) + v1 * doth(13) + v2 * doth(14);
  out[hidx(0, 0)] += m00; out[hidx(0, 1)] += m01; out[hidx(0, 2)] += m02;
  out[hidx(1, 1)] += m11; out[hidx(1, 2)] += m12; out[hidx(2, 2)] += m22;
  out[hidx(0, 3)] += a0; out[hidx(1, 3)] += a1; out[hidx(2, 3)] += a2;
  out[hidx(0, 4)] += b0; out[hidx(1, 4)] += b1; out[hidx(2, 4)] += b2;
  out[hidx(0, 5)] += e0; out[hidx(1, 5)] += e1; out[hidx(2, 5)] += e2;
  out[hidx(3, 3)] += j3y * a1 + j3z * a2 + vh33;
  out[hidx(3, 4)] += j3y * b1 + j3z * b2 + vh34;
  out[hidx(3, 5)] += j3y * e1 + j3z * e2 + vh35;
  out[hidx(4, 4)] += j4x * b0 + j4y * b1 + j4z * b2 + vh44;
  out[hidx(4, 5)] += j4x * e0 + j4y * e1 + j4z * e2 + vh45;
  out[hidx(5, 5)] += j5x * e0 + j5y * e1 + j5z * e2 + vh55;
}

// Transposing warp reduction of 32 floats per lane: afterwards lane l holds in v[0] the sum over
// the warp of value l.  31 shuffles instead of 32 x 5; the summation tree is fixed.
__device__ __forceinline__ void warp_transpose_reduce32(float v[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = up ? v[k] : v[k + off];
      const float keep = up ? v[k + off] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

// ---- bulk asynchronous copy global -> shared (the TMA engine's 1-D form, cp.async.bulk; SASS: UBLKCP) ---------------
// One elected thread posts the expected byte count on an mbarrier and issues the copies; the copy engine moves the data
// without passing through registers and completes the transaction count on the barrier, on whose phase every thread waits.
// Sizes and addresses are multiples of 16 bytes (cudaMalloc'd arrays, 16-byte records).
#ifndef B200_STAGE_TMA
#define B200_STAGE_TMA 1
#endif
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src_gmem), "r"(bytes),
               "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}

// group barrier: monotonically increasing counter, one arrival per CTA per use
__device__ __forceinline__ void group_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// PROF: developer cycle counters (b200reg_set_profile).  A template parameter, not a run-time switch: the
// counters and time stamps of the profiled build cost every thread a dozen registers, which the 128-register
// kernel pays for in spills — ncu showed 4.7 M local-memory sectors per registration (more than the
// algorithmic traffic) coming from the counters alone when they were always on.
template <int MODE, bool PROF>
__global__ void __launch_bounds__(kAlignThreads, 1) k_ndt_align(const NdtJob* __restrict__ jobs, int n_jobs, int ctas_per_group, NdtTargetQueue tq, NdtParams prm, double* partials_all, unsigned int* barriers,
                                                                unsigned int* queue, const __grid_constant__ NdtJob single) {
  __shared__ NdtShared s;
  __shared__ GridParams s_gp;
  __shared__ int s_job;
  __shared__ __align__(8) uint64_t s_stage_bar;  // mbarrier of the grid staging copies
  uint32_t stage_parity = 0;
  if (B200_STAGE_TMA && threadIdx.x == 0) mbar_init(&s_stage_bar, 1);
  __syncthreads();
  extern __shared__ __align__(16) unsigned char stage[];  // kStageBytes
  const int G = ctas_per_group;
  const int group = blockIdx.x / G, rank = blockIdx.x % G;
  double* partials = partials_all + (size_t)group * 2 * G * kAccStride;
  unsigned int* barrier = barriers + group * 32;  // one 128-byte line per group: [0] arrival counter, [1..2] job mailbox
  unsigned int epoch = 0;
  int parity = 0;
  unsigned int fetched = 0;
  int cur_run = -1;  // target run this CTA is working through (full batches, see NdtTargetQueue)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const void* staged_table = nullptr;  // which grid currently sits in shared memory

  // Jobs are handed out from one global counter: registrations differ in their number of passes,
  // so a static split would leave groups idle behind the slowest one.  Rank 0 of the group takes
  // the ticket; the other CTAs of the group read it from the mailbox after the group barrier.
  while (true) {
    int jb;
    if (jobs == nullptr) {
      // one registration riding in the kernel parameters: no ticket, no mailbox, no barrier to fetch it
      if (fetched) break;
      jb = 0;
      ++fetched;
    } else if (G == 1) {
      // Full batches: the pairs arrive grouped by target.  A CTA works through the pairs of ONE target run
      // (its 200 KB grid is staged once for all of them); runs are handed out from a ticket counter, pairs
      // inside a run from the run's own counter, and a CTA that finds no run left STEALS pairs from runs
      // still in progress — so the tail of the batch is one pair long, not one run long.
      if (tid == 0) {
        int got = n_jobs;
        if (tq.runs == nullptr) {
          got = (int)atomicAdd(queue, 1u);
        } else {
          while (true) {
            if (cur_run >= 0) {
              const uint2 rn = tq.runs[cur_run];
              const unsigned int i = atomicAdd(tq.next + cur_run, 1u);
              if (i < rn.y) { got = (int)(rn.x + i); break; }
            }
            const unsigned int t = atomicAdd(queue, 1u);
            if (t < (unsigned)tq.n_runs) { cur_run = (int)t; continue; }
            int found = -1;
            for (int k = 0; k < tq.n_runs; ++k) {
              const int g = (int)((blockIdx.x + (unsigned)k) % (unsigned)tq.n_runs);
              unsigned int v;
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(tq.next + g) : "memory");
              if (v < tq.runs[g].y) { found = g; break; }
            }
            if (found < 0) break;  // every pair of every run has been taken
            cur_run = found;
          }
        }
        s_job = got;
      }
      __syncthreads();
      jb = s_job;
      __syncthreads();
    } else {
      unsigned int* mailbox = barrier + 1 + (fetched & 1u);
      if (rank == 0 && tid == 0) {
        const unsigned int j = atomicAdd(queue, 1u);
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(mailbox), "r"(j) : "memory");
      }
      epoch += (unsigned)G;
      group_barrier(barrier, epoch);
      unsigned int j;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(j) : "l"(mailbox) : "memory");
      jb = (int)j;
      ++fetched;
    }
    if (jb >= n_jobs) break;
    const NdtJob& job = jobs ? jobs[jb] : single;  // one registration: the job rides in the kernel parameters
    // the lattice of the target (20 words) is read by every thread for every point: it lives in shared
    // memory, not in 20 registers per thread of a kernel that is already spilling (ncu: the first use of
    // gp.leaf was the top long-scoreboard stall of the batch configuration — a reload from local / global)
    __syncthreads();
    if (tid < (int)(sizeof(GridParams) / 4)) reinterpret_cast<uint32_t*>(&s_gp)[tid] = reinterpret_cast<const uint32_t*>(&job.grid.meta->grid)[tid];
    __syncthreads();
    const GridParams& gp = s_gp;
    const int n_src = job.n_src;
    long long prof[PROF ? 10 : 1] = {0};
    const long long ts0 = PROF ? clock64() : 0;
    // ---- stage the target grid in shared memory when it fits
    NdtLookup look;
    bool grid_staged = false;
    {
      const uint32_t n_rec = job.grid.gmeta->n_records, cap = job.grid.gmeta->table_cap;
      const size_t table_bytes = (size_t)cap * sizeof(uint2), vox_bytes = (size_t)n_rec * sizeof(NdtVoxel);
      const size_t cen_bytes = MODE == 0 ? (size_t)n_rec * sizeof(float4) : 0;
      look.mask = cap / 2 - 1;  // buckets of two entries
      if (table_bytes + vox_bytes + cen_bytes <= (size_t)kStageBytes) {
        if (staged_table != (const void*)job.grid.table) {
          __syncthreads();
#if B200_STAGE_TMA
          // three bulk copies (hash table, 48-byte voxel records, centroids for KDTREE) posted by one thread: ~200 KB per SM
          // without an instruction per 16 bytes; everybody waits on the barrier's phase
          if (tid == 0) {
            mbar_expect_tx(&s_stage_bar, (uint32_t)(table_bytes + vox_bytes + cen_bytes));
            // a single bulk copy moves at most 2^20 - 16 bytes here; the pieces are far below that
            bulk_g2s(stage, job.grid.table, (uint32_t)table_bytes, &s_stage_bar);
            if (vox_bytes) bulk_g2s(stage + table_bytes, job.grid.voxels, (uint32_t)vox_bytes, &s_stage_bar);
            if (cen_bytes) bulk_g2s(stage + table_bytes + vox_bytes, job.grid.centroids, (uint32_t)cen_bytes, &s_stage_bar);
          }
          mbar_wait(&s_stage_bar, stage_parity);
          stage_parity ^= 1u;
#else
          const uint4* src_t = reinterpret_cast<const uint4*>(job.grid.table);
          uint4* dst = reinterpret_cast<uint4*>(stage);
          const int n_t = (int)(table_bytes / 16), n_v = (int)(vox_bytes / 16), n_c = (int)(cen_bytes / 16);
          for (int i = tid; i < n_t; i += kAlignThreads) dst[i] = __ldg(src_t + i);
          const uint4* src_v = reinterpret_cast<const uint4*>(job.grid.voxels);
          for (int i = tid; i < n_v; i += kAlignThreads) dst[n_t + i] = __ldg(src_v + i);
          const uint4* src_c = reinterpret_cast<const uint4*>(job.grid.centroids);
          for (int i = tid; i < n_c; i += kAlignThreads) dst[n_t + n_v + i] = __ldg(src_c + i);
#endif
          staged_table = (const void*)job.grid.table;
        }
        look.table = reinterpret_cast<const uint2*>(stage);
        look.voxels = reinterpret_cast<const NdtVoxel*>(stage + table_bytes);
        look.centroids = reinterpret_cast<const float4*>(stage + table_bytes + vox_bytes);
        look.s_table = (uint32_t)__cvta_generic_to_shared(stage);
        look.s_voxels = look.s_table + (uint32_t)table_bytes;
        look.s_centroids = look.s_voxels + (uint32_t)vox_bytes;
        grid_staged = true;
      } else {
        look.table = job.grid.table;
        look.voxels = job.grid.voxels;
        look.centroids = job.grid.centroids;
        look.s_table = look.s_voxels = look.s_centroids = 0u;
        grid_staged = false;
      }
    }
    if (tid < 32) {
      if (tid == 0) {
        // gauss constants (A.4), recomputed per registration like upstream
        const double c1 = 10.0 * (1 - prm.outlier_ratio);
        const double c2 = prm.outlier_ratio / pow((double)(float)prm.resolution, 3);
        const double d3 = -log(c2);
        s.gauss_d1 = -log(c1 + c2) - d3;
        s.gauss_d2 = -2 * log((-log(c1 * exp(-0.5) + c2) - d3) / s.gauss_d1);
        for (int i = 0; i < 6; ++i) s.p[i] = job.p0[i];
        s.nr_iterations = 0; s.converged = 0; s.n_eval = 0; s.n_pass = 0; s.hits = 0.0;
        for (int i = 0; i < 6; ++i) s.pf[i] = 0;
        s.score = 0.0;
        s.need_hessian = 1;
        s.solve_path = 0;
        s.phase = job.eval_only ? PH_EVAL_ONLY : PH_INIT;
      }
      __syncwarp();
      pose_trig_warp(s, s.p, lane);
      if (tid == 0) {
        if (job.eval_only) pose_to_T(s.p, s.trig, s.T);
        else
          for (int i = 0; i < 12; ++i) s.T[i] = job.guess[i];  // first pass: the guess matrix itself
        angle_tables(s.trig, s.j_ang, s.h_ang);
      }
    }
    __syncthreads();
    if (PROF) prof[6] = clock64() - ts0;
    const int n_groups32 = (n_src + 31) >> 5;
    const float res2 = __fmul_rn((float)prm.resolution, (float)prm.resolution);
    while (true) {
      // ---- pass
      const long long t0 = PROF ? clock64() : 0;
      constexpr int need_h = 1;  // every pass evaluates the Hessian (see ndt_step, PH_MT_ITER)
      const float d1h = (float)s.gauss_d1, d1l = (float)(s.gauss_d1 - (double)d1h);
      const float gd2 = (float)s.gauss_d2;
      double accd = 0.0;  // lane l accumulates accumulator l of this warp's point groups
      if (gp.any && !gp.overflow) {
        // 32-point groups dealt round-robin over the group's warps.  A lane keeps float partial sums
        // of up to kLaneGroup of its points before the warp-transposed reduction folds them into the
        // double accumulator: with one SM per registration a lane sees ~90 points per pass and the
        // 31-shuffle reduction is paid once per kLaneGroup of them.  (With all 148 SMs on one
        // registration a warp has at most one group per pass: nothing changes there, bit for bit.)
        float vals[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) vals[k] = 0.f;
        int pending = 0;
        // the source point of the NEXT group is requested before the current one is processed: with
        // 16 warps per SM an L2 / HBM load issued at the point of use is not hidden (ncu: long
        // scoreboard was the top stall of the one-CTA-per-registration configuration)
        const int qstride = G * kAlignWarps;
        int q = warp * G + rank;
        // two groups ahead: with one SM per registration the 148 source clouds of a batch do not stay in
        // L2, and one group of look-ahead did not cover an HBM round trip (ncu: long scoreboard on the
        // first use of the point was the top stall of the batch configuration)
        auto fetch = [&](int qq) -> float4 {
          const int ii = (qq << 5) + lane;
          return (qq < n_groups32 && ii < n_src) ? __ldg(job.src + ii) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        float4 pt_n1 = fetch(q), pt_n2 = fetch(q + qstride);
        for (; q < n_groups32; q += qstride) {
          const int i = (q << 5) + lane;
          const float4 pt = pt_n1;
          pt_n1 = pt_n2;
          pt_n2 = fetch(q + 2 * qstride);
          if (i < n_src) {
            if (grid_staged) ndt_point<MODE, true>(s, look, gp, pt, d1h, d1l, gd2, res2, need_h, vals);
            else ndt_point<MODE, false>(s, look, gp, pt, d1h, d1l, gd2, res2, need_h, vals);
          }
          if (++pending == kLaneGroup) {
            warp_transpose_reduce32(vals, lane);
            accd += (double)vals[0];
#pragma unroll
            for (int k = 0; k < 32; ++k) vals[k] = 0.f;
            pending = 0;
          }
        }
        if (pending) {
          warp_transpose_reduce32(vals, lane);
          accd += (double)vals[0];
        }
      }
      const long long t1 = PROF ? clock64() : 0;
      // ---- block reduce: warp totals -> shared memory -> one partial row per CTA
      s.red[warp][lane] = accd;
      __syncthreads();
      if (tid < kAccStride) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kAlignWarps; ++w) v += s.red[w][tid];
        if (G == 1) s.tot[tid] = v;  // a one-CTA group (loop-closure batches): no global round trip, no barrier
        else partials[((size_t)parity * G + rank) * kAccStride + tid] = v;
      }
      // ---- group sync + redundant fixed-order reduction of the G partials.  (Tagging every row with
      // its own arrival flag instead of one counter barrier was measured: 2368 polling lanes cost more
      // than 148 spinners on one line — 288 us vs 223 us per registration.)
      const long long t2 = PROF ? clock64() : 0;
      long long t3 = t2;
      if (G > 1) {
        epoch += (unsigned)G;
        group_barrier(barrier, epoch);
        if (PROF) t3 = clock64();
        // warp w sums rows w, w+16, ... (each row one coalesced 256-byte read), then the 16 row
        // groups are combined through shared memory; fixed order -> every CTA gets the same bits
        const double* base = partials + (size_t)parity * G * kAccStride + lane;
        double tmp[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) {  // all loads in flight before the first add (G <= 160)
          const int r = warp + kAlignWarps * k;
          tmp[k] = r < G ? __ldcg(base + (size_t)r * kAccStride) : 0.0;
        }
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 10; ++k) v += tmp[k];
        for (int r = warp + kAlignWarps * 10; r < G; r += kAlignWarps) v += __ldcg(base + (size_t)r * kAccStride);
        __syncthreads();  // s.red is reused
        s.red[warp][lane] = v;
        __syncthreads();
        if (tid < kAccStride) {
          double t = 0.0;
#pragma unroll
          for (int w = 0; w < kAlignWarps; ++w) t += s.red[w][tid];
          s.tot[tid] = t;
        }
      }
      parity ^= 1;
      __syncthreads();
      const long long t4 = PROF ? clock64() : 0;
      // ---- optimiser step (warp 0 of every CTA, identical inputs -> identical state)
      if (tid < 32) {
        if (tid == 0) {
          if (s.phase == PH_EVAL_ONLY) {
            s.n_eval++;
            s.n_pass++;
            s.hits += s.tot[28];
            s.phase = PH_DONE;
            s.new_pose = 0;
          }
        }
        __syncwarp();
        if (s.phase != PH_EVAL_ONLY && s.phase != PH_DONE) ndt_step_warp<PROF>(s, prm, lane);  // warp-uniform (PH_EVAL_ONLY became PH_DONE just above)
        const long long tb = PROF ? clock64() : 0;
        __syncwarp();
        if (s.new_pose) {
          pose_trig_warp(s, s.x_t, lane);
          const long long tc = PROF ? clock64() : 0;
          if (tid == 0) {
            pose_to_T(s.x_t, s.trig, s.T);
            angle_tables(s.trig, s.j_ang, s.h_ang);
            if (PROF) { prof[8] += tc - tb; prof[9] += clock64() - tc; }
          }
        }
        if (PROF && tid == 0) prof[7] += tb - t4;
      }
      if (PROF && job.trace && rank == 0 && tid == 0 && s.n_pass <= kTraceCap) {
        // one record per pass, the quantities the oracle's trace holds (oracle.hpp NDT::trace)
        double* r = job.trace + (size_t)s.n_pass * kTraceDoubles;
        r[0] = s.nr_iterations; r[1] = s.step_iterations; r[2] = s.a_t; r[3] = s.tot[0]; r[4] = s.phi_t; r[5] = s.d_phi_t; r[6] = s.psi_t; r[7] = s.d_psi_t;
        r[8] = s.open_interval; r[9] = s.interval_converged; r[10] = s.phi_0; r[11] = s.d_phi_0; r[12] = s.solve_path;
        s.solve_path = 0;
        job.trace[0] = (double)s.n_pass;
      }
      __syncthreads();
      if (PROF) {
        const long long t5 = clock64();
        prof[0] += t1 - t0; prof[1] += t2 - t1; prof[2] += t3 - t2; prof[3] += t4 - t3; prof[4] += t5 - t4; prof[5] += 1;
      }
      if (s.phase == PH_DONE) break;
    }
    if (PROF && job.prof && rank == 0 && tid == 0) {
      for (int k = 0; k < 10; ++k) job.prof[k] = prof[k];
      for (int k = 0; k < 6; ++k) job.prof[10 + k] = s.pf[k];
    }
    // ---- result
    if (rank == 0 && tid == 0) {
      if (job.eval_only) {
        double* o = job.deriv_out;
        o[0] = s.tot[0];
        for (int i = 0; i < 6; ++i) o[1 + i] = s.tot[1 + i];
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) o[7 + 6 * i + j] = s.tot[i <= j ? hidx(i, j) : hidx(j, i)];
      }
      b200reg_result r;
      // column-major 4x4 from the row-major 3x4 of the last evaluated transform
      for (int c = 0; c < 4; ++c) {
        for (int rr = 0; rr < 3; ++rr) r.transformation[4 * c + rr] = s.T[4 * rr + c];
        r.transformation[4 * c + 3] = c == 3 ? 1.f : 0.f;
      }
      r.fitness = 0.0;
      r.score = n_src > 0 ? s.score / (double)n_src : 0.0;
      r.converged = s.converged;
      r.iterations = s.nr_iterations;
      r.evaluations = s.n_eval;
      r.passes = s.n_pass;
      r.hits = (long long)s.hits;
      *job.result = r;
      if (job.result_host) {
        *job.result_host = r;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int*>(job.done_flag) = job.done_seq;
      }
    }
    __syncthreads();
  }
  // the last CTA to leave puts the barrier lines, the job counter and the exit counter back to
  // zero, so the next launch needs no memset in front of it
  if (tid == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(queue + 1, 1u);
    if (prev == gridDim.x - 1) {
      const int n_groups = (int)gridDim.x / G;
      for (int gq = 0; gq < n_groups; ++gq) { barriers[gq * 32] = 0u; barriers[gq * 32 + 1] = 0u; barriers[gq * 32 + 2] = 0u; }
      queue[0] = 0u;
      queue[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace b200
