// b200reg — the host side of the two front-end nodelets, in C++ above the C ABI (include/b200reg.h, b200reg_odometry_*
// and b200reg_frontend_*).
//
// The reference's callers of the registration path are compiled C++: ScanMatchingOdometryNodelet::matching
// [REF apps/scan_matching_odometry_nodelet.cpp:173-270] and PrefilteringNodelet::cloud_callback
// [REF apps/prefiltering_nodelet.cpp:121-162], two nodelets of one manager joined by the /filtered_points topic
// [REF launch/delta_graph_slam.launch:26,46].  delta_graph_slam_b200/odometry.py mirrors them in Python for the parity
// tests; this file is the same state machine as host C++ — O(1) float 4x4 algebra per frame around the engine calls —
// so that a frame costs the engine's kernels plus a few microseconds, not an interpreter's ~50 us of small-array
// overhead.  Everything that touches points is a C-ABI call into the engine; nothing here computes on clouds.
//
//   b200reg_odometry_*  : matching(stamp, cloud) -> odom, one call per scan; keyframe promotion on the device
//   b200reg_frontend_*  : prefilter (distance gate + VoxelGrid) of scan k+1 in flight on its own handle / stream / SMs
//                         while scan k is matched — the two nodelets as the pipeline they are — plus a whole-sequence
//                         runner for device-resident scans (bench `value` leg)
//
// Float algebra is Eigen's: Matrix4f products accumulate k = 0..3 in order, no contraction (this file is compiled with
// -fmad=false for device code and the host compiler's default, which does not contract across statements).
#include <math.h>
#include <string.h>

#include <chrono>
#include <new>
#include <string>
#include <vector>

#include "../../include/b200reg.h"
#include "common.cuh"

// internal to the library (b200reg_api.cu): b200reg_align in two halves
extern "C" int b200reg_internal_align_begin(b200reg_handle* h, const float* guess, float* aligned_xyzw);
extern "C" int b200reg_internal_align_end(b200reg_handle* h);
extern "C" int b200reg_internal_set_defer_source_sync(b200reg_handle* h, int on);

namespace {

inline double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct M4 {
  float m[16];  // row-major
  float& operator()(int r, int c) { return m[4 * r + c]; }
  float operator()(int r, int c) const { return m[4 * r + c]; }
};
M4 identity4() {
  M4 I;
  for (int i = 0; i < 16; ++i) I.m[i] = (i % 5 == 0) ? 1.f : 0.f;
  return I;
}
M4 mul4(const M4& a, const M4& b) {
  M4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      volatile float acc = a(i, 0) * b(0, j);  // volatile: one rounding per product and per sum, in k order
      for (int k = 1; k < 4; ++k) {
        volatile float prod = a(i, k) * b(k, j);
        acc = acc + prod;
      }
      r(i, j) = acc;
    }
  return r;
}
M4 from_colmajor(const float* c) {
  M4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) r(i, j) = c[4 * j + i];
  return r;
}
void to_colmajor(const M4& a, float* c) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) c[4 * j + i] = a(i, j);
}
// inverse of a rigid transform's 4x4 by cofactors in double, rounded once (Eigen's Matrix4f::inverse() is a float
// cofactor expansion; the only caller is the optional jump gate, a threshold test on the result)
M4 inverse4(const M4& a) {
  double m[16], inv[16];
  for (int i = 0; i < 16; ++i) m[i] = a.m[i];
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  M4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = (float)(inv[i] / det);
  return r;
}
// w of Eigen::Quaternionf(R) (Eigen's rotation-matrix -> quaternion conversion), float
float quaternion_w(const M4& T) {
  float t = T(0, 0) + T(1, 1) + T(2, 2);
  if (t > 0.f) return 0.5f * sqrtf(t + 1.0f);
  int i = 0;
  if (T(1, 1) > T(0, 0)) i = 1;
  if (T(2, 2) > T(i, i)) i = 2;
  const int j = (i + 1) % 3, k = (j + 1) % 3;
  t = sqrtf(T(i, i) - T(j, j) - T(k, k) + 1.0f);
  return (T(k, j) - T(j, k)) * (0.5f / t);
}
float norm3(const M4& T) {
  volatile float a = T(0, 3) * T(0, 3), b = T(1, 3) * T(1, 3), c = T(2, 3) * T(2, 3);
  volatile float s = a + b;
  s = s + c;
  return sqrtf(s);
}
double clamp_acos(float w) {
  double v = (double)w;
  if (v > 1.0) v = 1.0;
  if (v < -1.0) v = -1.0;
  return acos(v);
}

}  // namespace

struct b200reg_odometry {
  b200reg_handle* reg = nullptr;  // borrowed
  b200reg_odometry_config cfg;
  bool have_keyframe = false;
  M4 keyframe_pose, prev_trans;
  double keyframe_stamp = 0.0, prev_time = 0.0;
  bool have_prev_time = false;
  int num_keyframes = 0, last_converged = 1, last_switched = 0;
  b200reg_result last{};
  std::string err;
  int prepare_promotion = 0;  // 0 off, 1 when the motion so far says this scan will probably become the keyframe, 2 every scan
  double last_step = 0.0;
  long long n_hints = 0;
  // host wall clock spent inside matching(), by phase (b200reg_frontend_get_timing): {set source, align launch -> result, keyframe promotion}
  double t_source = 0.0, t_align = 0.0, t_promote = 0.0;
  double pend_stamp = 0.0, pend_dist_before = 0.0, pend_t1 = 0.0;  // matching_begin .. matching_end
  long long n_match = 0, n_promote = 0;
};

extern "C" {

void b200reg_odometry_default_config(b200reg_odometry_config* c) {
  if (!c) return;
  // defaults of the nodelet [REF apps/scan_matching_odometry_nodelet.cpp:73-80]
  c->keyframe_delta_trans = 0.25;
  c->keyframe_delta_angle = 0.15;
  c->keyframe_delta_time = 1.0;
  c->transform_thresholding = 0;
  c->max_acceptable_trans = 1.0;
  c->max_acceptable_angle = 1.0;
}

int b200reg_odometry_create(b200reg_handle* registration, const b200reg_odometry_config* cfg, b200reg_odometry** out) {
  if (!registration || !cfg || !out) return B200REG_E_INVALID;
  b200reg_odometry* o = new (std::nothrow) b200reg_odometry();
  if (!o) return B200REG_E_INVALID;
  o->reg = registration;
  o->cfg = *cfg;
  o->keyframe_pose = identity4();
  o->prev_trans = identity4();
  *out = o;
  return B200REG_OK;
}

int b200reg_odometry_destroy(b200reg_odometry* o) {
  if (!o) return B200REG_E_INVALID;
  delete o;
  return B200REG_OK;
}

const char* b200reg_odometry_last_error(const b200reg_odometry* o) { return o ? o->err.c_str() : "null odometry object"; }

int b200reg_odometry_reset(b200reg_odometry* o) {
  if (!o) return B200REG_E_INVALID;
  o->have_keyframe = false;
  o->num_keyframes = 0;
  return B200REG_OK;
}

// matching(stamp, cloud): the cloud is the odometry nodelet's input (the prefiltered scan); `device` != 0 means xyzw is
// a device pointer on the registration handle's GPU.  guess_delta16 (column-major, may be NULL = identity) is the
// msf / robot-odometry delta the reference multiplies onto prev_trans (:190-214).
// matching() in two halves: _begin runs the state machine up to the launch of the registration, _end takes the result
// and finishes the frame.  *finished = true: the frame needed no registration (first keyframe) or failed early, odom16 is
// final and _end must not be called.
static int matching_begin(b200reg_odometry* o, double stamp, const float* xyzw, size_t n, size_t stride, int device, const float* guess_delta16, float* aligned_out, float* odom16, bool* finished) {
  *finished = true;
  if (!o || !odom16 || (n && !xyzw)) return B200REG_E_INVALID;
  b200reg_handle* reg = o->reg;
  auto fail = [&](int rc) {
    o->err = b200reg_last_error(reg);
    return rc;
  };
  int rc;
  o->last_switched = 0;
  if (!o->have_keyframe) {
    o->have_prev_time = false;
    o->prev_trans = identity4();
    o->keyframe_pose = identity4();
    o->keyframe_stamp = stamp;
    rc = device ? b200reg_set_target_device(reg, xyzw, n) : b200reg_set_target(reg, xyzw, n, stride);
    if (rc != B200REG_OK) return fail(rc);
    o->have_keyframe = true;
    o->num_keyframes = 1;
    o->last_switched = 1;
    to_colmajor(identity4(), odom16);
    return B200REG_OK;
  }
  const double t0 = now_us();
  rc = device ? b200reg_set_source_device(reg, xyzw, n) : b200reg_set_source(reg, xyzw, n, stride);
  if (rc != B200REG_OK) return fail(rc);
  const double t1 = now_us();
  o->t_source += t1 - t0;
  o->n_match += 1;
  M4 guess = o->prev_trans;
  if (guess_delta16) guess = mul4(o->prev_trans, from_colmajor(guess_delta16));
  float g16[16];
  to_colmajor(guess, g16);
  // scheduling hint for the engine (not part of the reference's logic, never changes a result): when the motion since the
  // keyframe says this scan will probably become the next keyframe, its target structures are built on a side stream
  // while it is being aligned, and the promotion below only swaps them in
  o->pend_dist_before = (double)norm3(o->prev_trans);
  if (o->prepare_promotion == 2 || (o->prepare_promotion == 1 && o->pend_dist_before + o->last_step > 0.95 * o->cfg.keyframe_delta_trans)) {
    b200reg_prepare_promotion(reg);
    o->n_hints += 1;
  }
  rc = b200reg_internal_align_begin(reg, g16, aligned_out);
  if (rc == B200REG_E_STATE) {  // PCL logs and returns with converged_ == false
    o->last_converged = 0;
    to_colmajor(mul4(o->keyframe_pose, o->prev_trans), odom16);
    return B200REG_OK;
  }
  if (rc != B200REG_OK) return fail(rc);
  o->pend_stamp = stamp;
  o->pend_t1 = t1;
  *finished = false;
  return B200REG_OK;
}

static int matching_end(b200reg_odometry* o, float* odom16) {
  b200reg_handle* reg = o->reg;
  auto fail = [&](int rc) {
    o->err = b200reg_last_error(reg);
    return rc;
  };
  const double stamp = o->pend_stamp, dist_before = o->pend_dist_before;
  int rc = b200reg_internal_align_end(reg);
  if (rc != B200REG_OK) return fail(rc);
  if ((rc = b200reg_get_result(reg, &o->last)) != B200REG_OK) return fail(rc);
  o->t_align += now_us() - o->pend_t1;
  o->last_converged = o->last.converged;
  if (!o->last.converged) {  // "scan matching has not converged!! ignore this frame": state untouched (:222-226)
    to_colmajor(mul4(o->keyframe_pose, o->prev_trans), odom16);
    return B200REG_OK;
  }
  const M4 trans = from_colmajor(o->last.transformation);
  const M4 odom = mul4(o->keyframe_pose, trans);
  if (o->cfg.transform_thresholding) {
    const M4 delta = mul4(inverse4(o->prev_trans), trans);
    const double dx = (double)norm3(delta), da = clamp_acos(quaternion_w(delta));
    if (dx > o->cfg.max_acceptable_trans || da > o->cfg.max_acceptable_angle) {
      to_colmajor(mul4(o->keyframe_pose, o->prev_trans), odom16);
      return B200REG_OK;
    }
  }
  o->prev_time = stamp;
  o->have_prev_time = true;
  o->prev_trans = trans;
  {
    const double step = (double)norm3(trans) - dist_before;
    o->last_step = step > 0.0 ? step : 0.0;
  }
  const double delta_trans = (double)norm3(trans), delta_angle = clamp_acos(quaternion_w(trans)), delta_time = stamp - o->keyframe_stamp;
  if (delta_trans > o->cfg.keyframe_delta_trans || delta_angle > o->cfg.keyframe_delta_angle || delta_time > o->cfg.keyframe_delta_time) {
    // keyframe = filtered; registration->setInputTarget(keyframe) (:252-254): the cloud just aligned changes role on the device
    const double tp = now_us();
    if ((rc = b200reg_promote_source_to_target(reg)) != B200REG_OK) return fail(rc);
    o->t_promote += now_us() - tp;
    o->n_promote += 1;
    o->keyframe_pose = odom;
    o->keyframe_stamp = stamp;
    o->prev_time = stamp;
    o->prev_trans = identity4();
    o->num_keyframes += 1;
    o->last_switched = 1;
  }
  to_colmajor(odom, odom16);
  return B200REG_OK;
}

static int matching_impl(b200reg_odometry* o, double stamp, const float* xyzw, size_t n, size_t stride, int device, const float* guess_delta16, float* aligned_out, float* odom16) {
  bool finished = true;
  const int rc = matching_begin(o, stamp, xyzw, n, stride, device, guess_delta16, aligned_out, odom16, &finished);
  if (rc != B200REG_OK || finished) return rc;
  return matching_end(o, odom16);
}

int b200reg_odometry_matching(b200reg_odometry* o, double stamp, const float* xyzw, size_t n, size_t stride_bytes, const float* guess_delta16, float* aligned_xyzw, float* odom16) {
  return matching_impl(o, stamp, xyzw, n, stride_bytes, 0, guess_delta16, aligned_xyzw, odom16);
}
int b200reg_odometry_matching_device(b200reg_odometry* o, double stamp, const float* d_xyzw, size_t n, const float* guess_delta16, float* odom16) {
  return matching_impl(o, stamp, d_xyzw, n, 16, 1, guess_delta16, nullptr, odom16);
}

int b200reg_odometry_get_state(b200reg_odometry* o, int* num_keyframes, int* last_converged, int* last_switched, float* keyframe_pose16, float* prev_trans16, b200reg_result* last_result) {
  if (!o) return B200REG_E_INVALID;
  if (num_keyframes) *num_keyframes = o->num_keyframes;
  if (last_converged) *last_converged = o->last_converged;
  if (last_switched) *last_switched = o->last_switched;
  if (keyframe_pose16) to_colmajor(o->keyframe_pose, keyframe_pose16);
  if (prev_trans16) to_colmajor(o->prev_trans, prev_trans16);
  if (last_result) *last_result = o->last;
  return B200REG_OK;
}

}  // extern "C"

// ---- the two nodelets as one pipeline -----------------------------------------------------------------------------
struct b200reg_frontend {
  b200reg_handle* filter = nullptr;  // owned: the prefiltering nodelet's engine handle
  b200reg_handle* reg = nullptr;     // owned: the registration object of the odometry nodelet
  b200reg_odometry* odo = nullptr;   // owned
  b200reg_frontend_config cfg;
  float leaf[3];
  // three filtered clouds in rotation (device): the registration copies its source / target out of them, and the
  // filter of the next scan must not overwrite the cloud the current registration is about to read
  float* d_out[3] = {nullptr, nullptr, nullptr};
  size_t cap = 0;
  int slot = 0;        // slot the filter in flight writes
  bool in_flight = false;
  double stamp_in_flight = 0.0;
  const float* host_out_in_flight = nullptr;  // two-nodelet form: the caller's filtered cloud of the scan in flight
  double t_filter_wait = 0.0, t_begin_next = 0.0, t_step = 0.0;  // host wall clock by phase (b200reg_frontend_get_timing)
  long long n_steps = 0;
  size_t n_last_filtered = 0;
  std::string err;
};

static int fe_reserve(b200reg_frontend* fe, size_t n) {
  if (n <= fe->cap) return B200REG_OK;
  cudaSetDevice(fe->cfg.device);
  const size_t want = n + n / 4 + 1024;
  for (int j = 0; j < 3; ++j) {
    if (fe->d_out[j]) cudaFree(fe->d_out[j]);
    fe->d_out[j] = nullptr;
    if (cudaMalloc((void**)&fe->d_out[j], want * 16) != cudaSuccess) {
      fe->err = "cudaMalloc of the filtered-cloud ring failed";
      fe->cap = 0;
      return B200REG_E_CUDA;
    }
  }
  fe->cap = want;
  return B200REG_OK;
}

extern "C" {

void b200reg_frontend_default_config(b200reg_frontend_config* c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->device = 0;
  b200reg_default_config(B200REG_METHOD_NDT, &c->registration);
  b200reg_odometry_default_config(&c->odometry);
  c->downsample_resolution = 0.1;   // prefiltering nodelet [REF apps/prefiltering_nodelet.cpp:56-57]
  c->use_distance_filter = 1;       // the gate runs on every scan [REF :150]
  c->distance_near_thresh = 1.0;    // [REF :101-102]
  c->distance_far_thresh = 100.0;
  c->filter_sms = 52;  // 36 CTAs for the filter's sort: 16 keys per thread of an HDL-64 scan (32 CTAs and fewer need 32), see DESIGN 5c
  c->prepare_promotion = 0;
  c->side_sms = 16;
}

int b200reg_frontend_create(const b200reg_frontend_config* cfg, b200reg_frontend** out) {
  if (!cfg || !out) return B200REG_E_INVALID;
  *out = nullptr;
  if (!(cfg->downsample_resolution > 0)) return B200REG_E_INVALID;
  b200reg_frontend* fe = new (std::nothrow) b200reg_frontend();
  if (!fe) return B200REG_E_INVALID;
  fe->cfg = *cfg;
  fe->leaf[0] = fe->leaf[1] = fe->leaf[2] = (float)cfg->downsample_resolution;
  b200reg_config fc;
  b200reg_default_config(B200REG_METHOD_NONE, &fc);
  fc.device = cfg->device;
  b200reg_config rc = cfg->registration;
  rc.device = cfg->device;
  int e;
  if ((e = b200reg_create(&fc, &fe->filter)) != B200REG_OK || (e = b200reg_create(&rc, &fe->reg)) != B200REG_OK || (e = b200reg_odometry_create(fe->reg, &cfg->odometry, &fe->odo)) != B200REG_OK) {
    if (fe->filter) b200reg_destroy(fe->filter);
    if (fe->reg) b200reg_destroy(fe->reg);
    delete fe;
    return e;
  }
  b200reg_set_distance_filter(fe->filter, cfg->use_distance_filter, cfg->distance_near_thresh, cfg->distance_far_thresh);
  fe->odo->prepare_promotion = cfg->prepare_promotion;
  if (cfg->filter_sms > 0) {
    // both persistent kernels are latency bound and neither needs the whole GPU: the filter handle gets a few SMs,
    // the registration the rest, so the filter of scan k+1 and the registration of scan k are resident together
    int total = 0;
    cudaDeviceGetAttribute(&total, cudaDevAttrMultiProcessorCount, cfg->device);
    if (total > b200::kNumSM) total = b200::kNumSM;
    // with prepared promotions a third persistent kernel (the side build's sort) may be resident: it takes side_sms of the filter's share
    const int side = (cfg->prepare_promotion && cfg->registration.method == B200REG_METHOD_NDT) ? (cfg->side_sms > 0 ? cfg->side_sms : 16) : 0;
    b200reg_set_sm_budget(fe->filter, cfg->filter_sms - side > 0 ? cfg->filter_sms - side : 1);
    b200reg_set_sm_budget(fe->reg, total - cfg->filter_sms);
    if (side) b200reg_set_side_budget(fe->reg, side);
  }
  // the caller's filtered clouds rotate (>= 3) and stay valid until two scans later: setInputSource does not wait for its DMA
  b200reg_internal_set_defer_source_sync(fe->reg, 1);
  *out = fe;
  return B200REG_OK;
}

int b200reg_frontend_destroy(b200reg_frontend* fe) {
  if (!fe) return B200REG_E_INVALID;
  if (fe->in_flight) {
    size_t n = 0;
    b200reg_voxelgrid_filter_end(fe->filter, &n);
  }
  b200reg_odometry_destroy(fe->odo);
  b200reg_destroy(fe->reg);
  b200reg_destroy(fe->filter);
  cudaSetDevice(fe->cfg.device);
  for (int j = 0; j < 3; ++j)
    if (fe->d_out[j]) cudaFree(fe->d_out[j]);
  delete fe;
  return B200REG_OK;
}

const char* b200reg_frontend_last_error(const b200reg_frontend* fe) { return fe ? fe->err.c_str() : "null front end"; }
b200reg_handle* b200reg_frontend_registration(b200reg_frontend* fe) { return fe ? fe->reg : nullptr; }
b200reg_handle* b200reg_frontend_filter(b200reg_frontend* fe) { return fe ? fe->filter : nullptr; }
b200reg_odometry* b200reg_frontend_odometry(b200reg_frontend* fe) { return fe ? fe->odo : nullptr; }

int b200reg_frontend_reset(b200reg_frontend* fe) {
  if (!fe) return B200REG_E_INVALID;
  if (fe->in_flight) {
    size_t n = 0;
    b200reg_voxelgrid_filter_end(fe->filter, &n);
    fe->in_flight = false;
  }
  return b200reg_odometry_reset(fe->odo);
}

// Enqueue the prefilter of one scan.  Where the filtered cloud goes decides how the scan reaches the registration:
//   filtered_out != NULL : the reference's two nodelets — the filtered cloud is written to the caller's host cloud (the
//                          /filtered_points message; a page-locked cloud is written by the centroid kernel itself) and the
//                          odometry side uploads it again with setInputSource, exactly as two separate nodelets would;
//   filtered_out == NULL : fused front end — the filtered cloud stays in the device ring and becomes the registration's
//                          source by a device-to-device copy (device-resident scans always take this form).
static int fe_begin(b200reg_frontend* fe, double stamp, const float* xyzw, size_t n, size_t stride, int device, float* filtered_out, size_t filtered_cap) {
  if (fe->in_flight) { fe->err = "a scan is already in flight (b200reg_frontend_step first)"; return B200REG_E_STATE; }
  int rc;
  const int slot = (fe->slot + 1) % 3;
  if (device) {
    if ((rc = fe_reserve(fe, n ? n : 1))) return rc;
    rc = b200reg_voxelgrid_filter_device_begin(fe->filter, xyzw, n, fe->leaf, 0, /*is_dense=*/0, fe->d_out[slot]);
  } else if (filtered_out) {
    rc = b200reg_voxelgrid_filter_begin(fe->filter, xyzw, n, stride, fe->leaf, 0, 0, filtered_out, filtered_cap);
  } else {
    if ((rc = fe_reserve(fe, n ? n : 1))) return rc;
    rc = b200reg_voxelgrid_filter_host_to_device_begin(fe->filter, xyzw, n, stride, fe->leaf, 0, 0, fe->d_out[slot]);
  }
  if (rc != B200REG_OK) { fe->err = b200reg_last_error(fe->filter); return rc; }
  fe->slot = slot;
  fe->in_flight = true;
  fe->stamp_in_flight = stamp;
  fe->host_out_in_flight = (!device && filtered_out) ? filtered_out : nullptr;
  return B200REG_OK;
}

int b200reg_frontend_begin(b200reg_frontend* fe, double stamp, const float* xyzw, size_t n, size_t stride_bytes, float* filtered_out, size_t filtered_capacity) {
  if (!fe || (n && !xyzw)) return B200REG_E_INVALID;
  return fe_begin(fe, stamp, xyzw, n, stride_bytes, 0, filtered_out, filtered_capacity);
}
int b200reg_frontend_begin_device(b200reg_frontend* fe, double stamp, const float* d_xyzw, size_t n) {
  if (!fe || (n && !d_xyzw)) return B200REG_E_INVALID;
  return fe_begin(fe, stamp, d_xyzw, n, 16, 1, nullptr, 0);
}

// collect the filter in flight, start the filter of the next scan (if any), match the collected scan
static int fe_step(b200reg_frontend* fe, double next_stamp, const float* next_xyzw, size_t next_n, size_t next_stride, int next_device, float* next_filtered_out, size_t next_filtered_cap,
                   size_t* n_filtered, float* aligned_out, float* odom16) {
  if (!fe || !odom16) return B200REG_E_INVALID;
  if (!fe->in_flight) { fe->err = "no scan in flight (b200reg_frontend_begin first)"; return B200REG_E_STATE; }
  size_t n = 0;
  const double t0 = now_us();
  int rc = b200reg_voxelgrid_filter_end(fe->filter, &n);
  fe->in_flight = false;
  if (rc != B200REG_OK) { fe->err = b200reg_last_error(fe->filter); return rc; }
  const double t1 = now_us();
  fe->t_filter_wait += t1 - t0;
  const int cur = fe->slot;
  const double stamp = fe->stamp_in_flight;
  const float* host_cloud = fe->host_out_in_flight;
  fe->n_last_filtered = n;
  if (n_filtered) *n_filtered = n;
  // the registration of this scan is launched first; the next scan's filter is enqueued while it runs (the host work of
  // that — ~11 us of launches — used to sit in front of the registration on the frame's critical path)
  bool finished = true;
  if (host_cloud) rc = matching_begin(fe->odo, stamp, host_cloud, n, 16, 0, nullptr, aligned_out, odom16, &finished);
  else rc = matching_begin(fe->odo, stamp, fe->d_out[cur], n, 16, 1, nullptr, aligned_out, odom16, &finished);
  if (rc != B200REG_OK) { fe->err = b200reg_odometry_last_error(fe->odo); return rc; }
  const double t2 = now_us();
  int rc_next = B200REG_OK;
  if (next_xyzw || next_n) rc_next = fe_begin(fe, next_stamp, next_xyzw, next_n, next_stride, next_device, next_filtered_out, next_filtered_cap);
  fe->t_begin_next += now_us() - t2;
  if (!finished) {
    rc = matching_end(fe->odo, odom16);
    if (rc != B200REG_OK) { fe->err = b200reg_odometry_last_error(fe->odo); return rc; }
  }
  if (rc_next != B200REG_OK) return rc_next;
  fe->t_step += now_us() - t0;
  fe->n_steps += 1;
  return B200REG_OK;
}

// host wall clock per phase since the last call (microseconds, summed): out8 = {steps, filter wait, begin of the next
// filter, set source, align launch -> result, keyframe promotions, number of promotions, whole steps}; resets the counters
int b200reg_frontend_get_timing(b200reg_frontend* fe, double* out8) {  // nine values
  if (!fe || !out8) return B200REG_E_INVALID;
  b200reg_odometry* o = fe->odo;
  out8[0] = (double)fe->n_steps; out8[1] = fe->t_filter_wait; out8[2] = fe->t_begin_next; out8[3] = o->t_source; out8[4] = o->t_align; out8[5] = o->t_promote;
  out8[6] = (double)o->n_promote; out8[7] = fe->t_step; out8[8] = (double)o->n_hints;
  o->n_hints = 0;
  fe->n_steps = 0; fe->t_filter_wait = fe->t_begin_next = fe->t_step = 0.0;
  o->t_source = o->t_align = o->t_promote = 0.0; o->n_match = o->n_promote = 0;
  return B200REG_OK;
}

int b200reg_frontend_step(b200reg_frontend* fe, double next_stamp, const float* next_xyzw, size_t next_n, size_t next_stride_bytes, float* next_filtered_out, size_t next_filtered_capacity,
                          size_t* n_filtered, float* aligned_out, float* odom16) {
  return fe_step(fe, next_stamp, next_xyzw, next_n, next_stride_bytes, 0, next_filtered_out, next_filtered_capacity, n_filtered, aligned_out, odom16);
}
int b200reg_frontend_step_device(b200reg_frontend* fe, double next_stamp, const float* next_d_xyzw, size_t next_n, size_t* n_filtered, float* odom16) {
  return fe_step(fe, next_stamp, next_d_xyzw, next_n, 16, 1, nullptr, 0, n_filtered, nullptr, odom16);
}

// a whole sequence of device-resident scans (bench `value` leg): odom16_out receives frames x 16 floats (column-major),
// results (optional) the registration record of every frame (zeroed for frame 0), n_filtered (optional) the filtered sizes
int b200reg_frontend_run_device(b200reg_frontend* fe, const float* const* d_scans, const size_t* n_points, const double* stamps, size_t frames, float* odom16_out, b200reg_result* results,
                                size_t* n_filtered, int* keyframes_out) {
  if (!fe || (frames && (!d_scans || !n_points || !odom16_out))) return B200REG_E_INVALID;
  int rc = b200reg_frontend_reset(fe);
  if (rc) return rc;
  if (!frames) return B200REG_OK;
  if ((rc = fe_begin(fe, stamps ? stamps[0] : 0.0, d_scans[0], n_points[0], 16, 1, nullptr, 0))) return rc;
  for (size_t k = 0; k < frames; ++k) {
    const bool more = k + 1 < frames;
    size_t nf = 0;
    rc = fe_step(fe, more ? (stamps ? stamps[k + 1] : 0.1 * (double)(k + 1)) : 0.0, more ? d_scans[k + 1] : nullptr, more ? n_points[k + 1] : 0, 16, 1, nullptr, 0, &nf, nullptr, odom16_out + 16 * k);
    if (rc != B200REG_OK) return rc;
    if (n_filtered) n_filtered[k] = nf;
    if (results) {
      if (k == 0) memset(&results[0], 0, sizeof(b200reg_result));
      else results[k] = fe->odo->last;
    }
  }
  if (keyframes_out) *keyframes_out = fe->odo->num_keyframes;
  return B200REG_OK;
}

}  // extern "C"
