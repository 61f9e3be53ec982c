// Synthetic HDL-64 / 128-beam lidar scans of a procedural street ("street_v1").
//
// Test and bench infrastructure, not part of the registration path: the
// reference ships no sample clouds, so every parity test and
// bench workload is driven by these deterministic scans.
// One implementation, compiled twice: by g++ into the oracle library (CPU
// tests) and by nvcc into libb200reg (bench generates the 1000-scan sequence
// on the device).  All geometry is double precision so both builds agree to
// rounding; the emitted points are float4 (x, y, z, 1) like pcl::PointXYZ.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SYNTH_HD __host__ __device__ inline
#else
#define SYNTH_HD inline
#endif

namespace synth {

struct Sensor {
  int beams;          // vertical channels
  int azimuth_steps;  // firings per revolution
  double elev_min_deg, elev_max_deg;
  double range_min, range_max;
  double range_sigma;  // gaussian range noise [m]
};

SYNTH_HD Sensor sensor_hdl64() { return Sensor{64, 2083, -24.8, 2.0, 1.0, 100.0, 0.02}; }
SYNTH_HD Sensor sensor_dense128() { return Sensor{128, 8192, -25.0, 15.0, 1.0, 100.0, 0.02}; }

SYNTH_HD uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
SYNTH_HD double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
// k-th uniform of stream (a, b)
SYNTH_HD double uni(uint64_t a, uint64_t b, uint64_t k) {
  return u01(splitmix64(splitmix64(a * 0x100000001B3ull + b) ^ (k * 0xD6E8FEB86659FD93ull)));
}
SYNTH_HD double gauss(uint64_t a, uint64_t b) {
  double u1 = uni(a, b, 0), u2 = uni(a, b, 1);
  if (u1 < 1e-300) u1 = 1e-300;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

// ---- scene -----------------------------------------------------------------
constexpr double kGroundZ = -1.73;
constexpr double kCellLen = 10.0;
constexpr double kGroundRange = 30.0;

struct Box { double lo[3], hi[3]; };

// slab test; returns entry distance or -1
SYNTH_HD double hit_box(const Box& b, const double o[3], const double d[3], double tmax) {
  double t0 = 0.0, t1 = tmax;
  for (int a = 0; a < 3; ++a) {
    if (fabs(d[a]) < 1e-12) {
      if (o[a] < b.lo[a] || o[a] > b.hi[a]) return -1.0;
    } else {
      double inv = 1.0 / d[a];
      double ta = (b.lo[a] - o[a]) * inv, tb = (b.hi[a] - o[a]) * inv;
      if (ta > tb) { double s = ta; ta = tb; tb = s; }
      if (ta > t0) t0 = ta;
      if (tb < t1) t1 = tb;
      if (t0 > t1) return -1.0;
    }
  }
  return t0 > 0.0 ? t0 : -1.0;
}

// vertical cylinder (cx, cy, r) from z0 to z1
SYNTH_HD double hit_pole(double cx, double cy, double r, double z0, double z1, const double o[3], const double d[3], double tmax) {
  double ox = o[0] - cx, oy = o[1] - cy;
  double a = d[0] * d[0] + d[1] * d[1];
  if (a < 1e-14) return -1.0;
  double b = ox * d[0] + oy * d[1];
  double c = ox * ox + oy * oy - r * r;
  double disc = b * b - a * c;
  if (disc < 0.0) return -1.0;
  double t = (-b - sqrt(disc)) / a;
  if (t <= 0.0 || t > tmax) return -1.0;
  double z = o[2] + t * d[2];
  if (z < z0 || z > z1) return -1.0;
  return t;
}

// sphere (bush); returns entry distance or -1
SYNTH_HD double hit_sphere(double cx, double cy, double cz, double r, const double o[3], const double d[3], double tmax) {
  double ox = o[0] - cx, oy = o[1] - cy, oz = o[2] - cz;
  double b = ox * d[0] + oy * d[1] + oz * d[2];
  double c = ox * ox + oy * oy + oz * oz - r * r;
  double disc = b * b - c;  // |d| = 1
  if (disc < 0.0) return -1.0;
  double t = -b - sqrt(disc);
  if (t <= 0.0 || t > tmax) return -1.0;
  return t;
}

// Nearest surface along a world-frame ray; returns range or -1 (miss).
// *fuzzy is set when the hit is vegetation (the caller adds a penetration depth).
// Street along +x, 10 m lots on both sides: a building segment with its own
// setback and height, a garden wall across the lot, a tree trunk, sometimes a
// bush (fuzzy sphere), sometimes a parked car; about one lot in seven is empty.
SYNTH_HD double cast_ray(uint64_t scene_seed, const double o[3], const double d[3], double tmax, bool* fuzzy) {
  double best = tmax;
  bool hit = false;
  *fuzzy = false;
  if (d[2] < -1e-9) {  // ground plane; returns at grazing incidence (beyond kGroundRange) are lost,
    double t = (kGroundZ - o[2]) / d[2];  // as on a real sensor, which also removes the far ring arcs
    if (t > 0.0 && t < best && t < kGroundRange) { best = t; hit = true; *fuzzy = false; }
  }
  long c0 = (long)floor((o[0] - tmax) / kCellLen), c1 = (long)floor((o[0] + tmax) / kCellLen);
  for (long c = c0; c <= c1; ++c) {
    for (int side = 0; side < 2; ++side) {
      double sgn = side ? 1.0 : -1.0;
      uint64_t key = (uint64_t)(c + 1000003) * 2u + (uint64_t)side;
      double x0 = (double)c * kCellLen;
      double face = 8.0 + 5.0 * uni(scene_seed, key, 2);
      bool empty_lot = uni(scene_seed, key, 8) < 0.15;
      Box b;
      if (!empty_lot) {  // building segment
        b.lo[0] = x0 + 1.5 * uni(scene_seed, key, 0);
        b.hi[0] = x0 + kCellLen - 1.5 * uni(scene_seed, key, 1);
        double hgt = 4.0 + 11.0 * uni(scene_seed, key, 3);
        if (side) { b.lo[1] = face; b.hi[1] = face + 12.0; } else { b.lo[1] = -face - 12.0; b.hi[1] = -face; }
        b.lo[2] = kGroundZ; b.hi[2] = kGroundZ + hgt;
        double t = hit_box(b, o, d, best);
        if (t > 0.0 && t < best) { best = t; hit = true; *fuzzy = false; }
      }
      {  // garden wall across the lot
        double wx = x0 + kCellLen * uni(scene_seed, key, 9);
        b.lo[0] = wx; b.hi[0] = wx + 0.3;
        if (side) { b.lo[1] = 5.5; b.hi[1] = face + 0.5; } else { b.lo[1] = -face - 0.5; b.hi[1] = -5.5; }
        b.lo[2] = kGroundZ; b.hi[2] = kGroundZ + 1.2 + uni(scene_seed, key, 10);
        double t = hit_box(b, o, d, best);
        if (t > 0.0 && t < best) { best = t; hit = true; *fuzzy = false; }
      }
      {  // tree trunk
        double px = x0 + kCellLen * uni(scene_seed, key, 4);
        double py = sgn * (5.0 + uni(scene_seed, key, 11));
        double t = hit_pole(px, py, 0.2 + 0.2 * uni(scene_seed, key, 12), kGroundZ, kGroundZ + 5.0, o, d, best);
        if (t > 0.0 && t < best) { best = t; hit = true; *fuzzy = false; }
      }
      {  // bush: a fuzzy sphere sitting on the ground
        double bx = x0 + kCellLen * uni(scene_seed, key, 13);
        double by = sgn * (4.5 + 3.0 * uni(scene_seed, key, 14));
        double br = 0.8 + 0.8 * uni(scene_seed, key, 15);
        double t = hit_sphere(bx, by, kGroundZ + 0.6 * br, br, o, d, best);
        if (t > 0.0 && t < best) { best = t; hit = true; *fuzzy = true; }
      }
      if (uni(scene_seed, key, 5) < 0.7) {  // parked car
        double cx = x0 + 5.0 * uni(scene_seed, key, 6);
        double cy = sgn * (3.6 + 0.6 * uni(scene_seed, key, 7));
        b.lo[0] = cx; b.hi[0] = cx + 4.5;
        b.lo[1] = cy - 0.9; b.hi[1] = cy + 0.9;
        b.lo[2] = kGroundZ; b.hi[2] = kGroundZ + 1.5;
        double t = hit_box(b, o, d, best);
        if (t > 0.0 && t < best) { best = t; hit = true; *fuzzy = false; }
      }
    }
  }
  return hit ? best : -1.0;
}

// One ray of a scan.  pose = row-major 4x4 sensor->world.  Returns true and
// writes the point in the SENSOR frame if the ray returns inside the range gate.
SYNTH_HD bool scan_ray(const Sensor& s, uint64_t scene_seed, uint64_t noise_seed, const double pose[16], long ray, float out[4]) {
  long az_i = ray / s.beams, beam = ray % s.beams;  // firing order: azimuth-major
  double az = 6.283185307179586 * (double)az_i / (double)s.azimuth_steps;
  double el = (s.elev_min_deg + (s.elev_max_deg - s.elev_min_deg) * (double)beam / (double)(s.beams - 1)) * 0.017453292519943295;
  double ce = cos(el);
  double ds[3] = {ce * cos(az), ce * sin(az), sin(el)};
  double d[3], o[3] = {pose[3], pose[7], pose[11]};
  for (int r = 0; r < 3; ++r) d[r] = pose[4 * r + 0] * ds[0] + pose[4 * r + 1] * ds[1] + pose[4 * r + 2] * ds[2];
  bool fuzzy;
  double t = cast_ray(scene_seed, o, d, s.range_max, &fuzzy);
  if (t < 0.0) return false;
  t += s.range_sigma * gauss(noise_seed, (uint64_t)ray);
  if (fuzzy) t += 0.8 * uni(noise_seed, (uint64_t)ray, 2);  // foliage penetration
  if (!(t > s.range_min && t < s.range_max)) return false;
  out[0] = (float)(ds[0] * t); out[1] = (float)(ds[1] * t); out[2] = (float)(ds[2] * t); out[3] = 1.0f;
  return true;
}

// "kitti_like" trajectory: 1 m per frame along +x with a gentle weave that keeps
// the vehicle inside the street (|y| < 2 m); small z / roll / pitch jitter.
// Host only (the device generator receives poses as input).
inline void rot_xyz(double rx, double ry, double rz, double R[9]) {
  double cx = cos(rx), sx = sin(rx), cy = cos(ry), sy = sin(ry), cz = cos(rz), sz = sin(rz);
  // R = Rz * Ry * Rx
  R[0] = cz * cy; R[1] = cz * sy * sx - sz * cx; R[2] = cz * sy * cx + sz * sx;
  R[3] = sz * cy; R[4] = sz * sy * sx + cz * cx; R[5] = sz * sy * cx - cz * sx;
  R[6] = -sy;     R[7] = cy * sx;                R[8] = cy * cx;
}
inline void pose_from_xyzrpy(const double v[6], double T[16]) {
  double R[9];
  rot_xyz(v[3], v[4], v[5], R);
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T[4 * r + c] = R[3 * r + c]; T[4 * r + 3] = v[r]; }
  T[12] = T[13] = T[14] = 0.0; T[15] = 1.0;
}
constexpr double kStep = 0.5;  // 5 m/s at 10 Hz
inline void traj_kitti_like(long k, uint64_t seed, double T[16]) {
  // yaw(k) = 0.1 sin(2 pi k / 100); position integrates the heading at kStep m / frame
  double x = 0.0, y = 0.0;
  for (long i = 0; i < k; ++i) {
    double yaw = 0.1 * sin(6.283185307179586 * (double)i / 100.0);
    x += kStep * cos(yaw); y += kStep * sin(yaw);
  }
  double v[6];
  v[0] = x; v[1] = y;
  v[2] = 0.002 * gauss(seed, (uint64_t)k * 3 + 0);
  v[3] = 0.0017453292519943296 * gauss(seed, (uint64_t)k * 3 + 1);
  v[4] = 0.0017453292519943296 * gauss(seed, (uint64_t)k * 3 + 2);
  v[5] = 0.1 * sin(6.283185307179586 * (double)k / 100.0);
  pose_from_xyzrpy(v, T);
}

}  // namespace synth
