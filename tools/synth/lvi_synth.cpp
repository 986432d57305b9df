// tools/synth/lvi_synth.cpp — deterministic synthetic LiDAR + IMU + mono-camera sequence generator
// (the reference ships no data, no simulator config and no tests, so bench.py and tests/
// drive both the CPU oracle and the CUDA library from this generator).  Not part of the product library and not
// part of the oracle: it only produces inputs in the reference's own layouts:
//   raw scans      licalib::PointXYZIT  (L/include/utils/pcl_utils.h:39-58), organised H x W, VLP-16 timing
//                  (L/include/utils/vlp_common.h:205-222,244-248)
//   LOAM poses     lidar pose in the first-scan lidar frame (src/aloam/src/laserMapping.cpp:890-900)
//   IMU samples    specific force convention of K/sensors/imu.h:25,61-101 (at rest accel = +9.79 along gravity-up)
//   ORB tracks     views (t0), observations (u,v,landmark), landmark reference + inverse depth
//                  (L/test/lvi_initialize_surfel_orb.cpp:337-455)
// Every random draw is a pure function of (seed, stream, index) so results do not depend on thread count.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
struct M3 { double m[9]; };
inline V3 mul(const M3& R, V3 v) { return {R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z, R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z}; }
inline V3 mulT(const M3& R, V3 v) { return {R.m[0] * v.x + R.m[3] * v.y + R.m[6] * v.z, R.m[1] * v.x + R.m[4] * v.y + R.m[7] * v.z, R.m[2] * v.x + R.m[5] * v.y + R.m[8] * v.z}; }
inline M3 mul(const M3& A, const M3& B) { M3 C; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C.m[r * 3 + c] = A.m[r * 3] * B.m[c] + A.m[r * 3 + 1] * B.m[3 + c] + A.m[r * 3 + 2] * B.m[6 + c]; return C; }
inline M3 tr(const M3& A) { M3 C; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C.m[r * 3 + c] = A.m[c * 3 + r]; return C; }
inline M3 rpy(double r, double p, double y) {  // Rz(y) Ry(p) Rx(r)
  const double cr = std::cos(r), sr = std::sin(r), cp = std::cos(p), sp = std::sin(p), cy = std::cos(y), sy = std::sin(y);
  return {{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr, -sp, cp * sr, cp * cr}};
}
inline void to_quat(const M3& R, double q[4]) {  // x,y,z,w
  const double* m = R.m;
  const double trc = m[0] + m[4] + m[8];
  double x, y, z, w;
  if (trc > 0) { double s = std::sqrt(trc + 1.0) * 2; w = 0.25 * s; x = (m[7] - m[5]) / s; y = (m[2] - m[6]) / s; z = (m[3] - m[1]) / s; }
  else if (m[0] > m[4] && m[0] > m[8]) { double s = std::sqrt(1.0 + m[0] - m[4] - m[8]) * 2; w = (m[7] - m[5]) / s; x = 0.25 * s; y = (m[1] + m[3]) / s; z = (m[2] + m[6]) / s; }
  else if (m[4] > m[8]) { double s = std::sqrt(1.0 + m[4] - m[0] - m[8]) * 2; w = (m[2] - m[6]) / s; x = (m[1] + m[3]) / s; y = 0.25 * s; z = (m[5] + m[7]) / s; }
  else { double s = std::sqrt(1.0 + m[8] - m[0] - m[4]) * 2; w = (m[3] - m[1]) / s; x = (m[2] + m[6]) / s; y = (m[5] + m[7]) / s; z = 0.25 * s; }
  const double n = std::sqrt(x * x + y * y + z * z + w * w);
  q[0] = x / n; q[1] = y / n; q[2] = z / n; q[3] = w / n;
}
inline M3 axis_angle(V3 a, double ang) {
  const double n = std::sqrt(dot(a, a)); a = (1.0 / n) * a;
  const double c = std::cos(ang), s = std::sin(ang), C = 1 - c;
  return {{c + a.x * a.x * C, a.x * a.y * C - a.z * s, a.x * a.z * C + a.y * s, a.y * a.x * C + a.z * s, c + a.y * a.y * C, a.y * a.z * C - a.x * s,
           a.z * a.x * C - a.y * s, a.z * a.y * C + a.x * s, c + a.z * a.z * C}};
}

inline uint64_t mix64(uint64_t z) { z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
inline double uni(uint64_t seed, uint64_t stream, uint64_t i) { return (static_cast<double>(mix64(mix64(seed ^ (stream << 48)) + i) >> 11) + 0.5) / 9007199254740992.0; }
inline double gauss(uint64_t seed, uint64_t stream, uint64_t i) {
  const double u1 = uni(seed, stream, 2 * i), u2 = uni(seed, stream, 2 * i + 1);
  return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
}

struct Box { double lo[3], hi[3]; };

// The three sensors free-run: their sample clocks have fixed phases against the LiDAR scan clock, so IMU / camera stamps are not
// commensurate with the 0.02 s knot grid anchored at the first scan (Kontiki's per-residual segment lookup throws for a time
// within round-off of a knot boundary, K/trajectories/spline_base.h:194-222).
constexpr double kImuPhase = 0.00137, kCamPhase = 0.00411;

}  // namespace

extern "C" {

struct synth_config {
  double t_start;        // time of the first scan (map time)
  double duration;       // seconds of LiDAR data
  int32_t rings;         // 16 (VLP-16) or 64
  int32_t az_steps;      // 1800 / 2000
  double scan_rate;      // 10 / 20 Hz
  double imu_rate;       // 200 Hz
  double cam_rate;       // 20 Hz frames
  int32_t keyframe_every;   // every 4th frame is a view
  int32_t n_landmarks;      // 4000
  int32_t max_track_views;  // ORB local-map horizon: a landmark is observed in at most this many consecutive views
  int32_t degenerate;       // C5 planar low-excitation motion
  double range_noise, loam_pos_noise, loam_rot_noise, gyro_noise, accel_noise, pixel_noise, rho_rel_noise;
  double pad_time;       // IMU/camera data extend this much before/after the LiDAR span (cfg time_offset_padding)
};

void synth_default_config(synth_config* c) {
  c->t_start = 10.0037;  /* not commensurate with the 0.02 s knot grid: exact knot-boundary times make the reference's segment lookup throw */ c->duration = 60.0; c->rings = 16; c->az_steps = 1800; c->scan_rate = 10.0; c->imu_rate = 200.0;
  c->cam_rate = 20.0; c->keyframe_every = 4; c->n_landmarks = 4000; c->max_track_views = 12; c->degenerate = 0;
  c->range_noise = 0.01; c->loam_pos_noise = 0.005; c->loam_rot_noise = 0.05 * M_PI / 180.0; c->gyro_noise = 2e-3;
  c->accel_noise = 2e-2; c->pixel_noise = 0.5; c->rho_rel_noise = 0.02; c->pad_time = 0.2;
}

// ground-truth extrinsics / biases
void synth_gt_extrinsics(double q_LtoI[4], double p_LinI[3], double q_CtoI[4], double p_CinI[3], double bg[3], double ba[3]) {
  const double D = M_PI / 180.0;
  to_quat(rpy(2 * D, -3 * D, 5 * D), q_LtoI);
  p_LinI[0] = 0.05; p_LinI[1] = -0.10; p_LinI[2] = 0.08;
  const M3 base = {{0, 0, 1, -1, 0, 0, 0, -1, 0}};  // camera (z fwd, x right, y down) -> body (x fwd, y left, z up)
  to_quat(mul(base, rpy(1 * D, -2 * D, 1.5 * D)), q_CtoI);
  p_CinI[0] = 0.10; p_CinI[1] = 0.05; p_CinI[2] = -0.03;
  bg[0] = 0.002; bg[1] = -0.001; bg[2] = 0.0015;
  ba[0] = 0.03; ba[1] = -0.02; ba[2] = 0.01;
}

}  // extern "C"

namespace {

struct Pose { M3 R; V3 p; };
struct Kin { M3 R; V3 p, v, a, w_body; };

// analytic IMU trajectory in the world frame (room frame, z up)
Kin gt_world(const synth_config& c, double t) {
  const double s = c.degenerate ? 0.3 : 1.0;
  const V3 ctr = {6.0, 4.5, 1.6};
  Kin k;
  k.p = {ctr.x + s * 1.5 * std::sin(0.5 * t), ctr.y + s * 1.0 * std::sin(0.7 * t + 0.3), ctr.z + (c.degenerate ? 0.0 : 0.3 * std::sin(1.1 * t))};
  k.v = {s * 1.5 * 0.5 * std::cos(0.5 * t), s * 1.0 * 0.7 * std::cos(0.7 * t + 0.3), c.degenerate ? 0.0 : 0.3 * 1.1 * std::cos(1.1 * t)};
  k.a = {-s * 1.5 * 0.25 * std::sin(0.5 * t), -s * 1.0 * 0.49 * std::sin(0.7 * t + 0.3), c.degenerate ? 0.0 : -0.3 * 1.21 * std::sin(1.1 * t)};
  double r, p, y, dr, dp, dy;
  if (c.degenerate) { r = 0; p = 0; y = 0.1 * std::sin(0.2 * t); dr = 0; dp = 0; dy = 0.02 * std::cos(0.2 * t); }
  else {
    r = 0.4 * std::sin(0.9 * t); dr = 0.36 * std::cos(0.9 * t);
    p = 0.3 * std::sin(1.3 * t + 1.0); dp = 0.39 * std::cos(1.3 * t + 1.0);
    y = 0.8 * std::sin(0.6 * t) + 0.2 * t; dy = 0.48 * std::cos(0.6 * t) + 0.2;
  }
  k.R = rpy(r, p, y);
  k.w_body = {dr - dy * std::sin(p), dp * std::cos(r) + dy * std::sin(r) * std::cos(p), -dp * std::sin(r) + dy * std::cos(r) * std::cos(p)};
  return k;
}

struct World {
  Box room; Box obs[4];
  M3 R_LI, R_CI; V3 p_LI, p_CI;
  World() {
    room = {{0, 0, 0}, {12.0, 9.0, 3.5}};
    obs[0] = {{1.37, 1.21, 0.0}, {2.61, 2.93, 1.83}};
    obs[1] = {{8.83, 1.57, 0.0}, {10.49, 2.71, 2.37}};
    obs[2] = {{1.93, 6.29, 0.0}, {3.77, 7.63, 1.41}};
    obs[3] = {{9.11, 6.07, 0.0}, {10.31, 7.89, 2.09}};
    double ql[4], pl[3], qc[4], pc[3], bg[3], ba[3];
    synth_gt_extrinsics(ql, pl, qc, pc, bg, ba);
    const double D = M_PI / 180.0;
    R_LI = rpy(2 * D, -3 * D, 5 * D);
    const M3 base = {{0, 0, 1, -1, 0, 0, 0, -1, 0}};
    R_CI = mul(base, rpy(1 * D, -2 * D, 1.5 * D));
    p_LI = {pl[0], pl[1], pl[2]}; p_CI = {pc[0], pc[1], pc[2]};
  }
  // first hit of ray o + s*dir, s > 0 ; returns range or inf
  double cast(V3 o, V3 d) const {
    double best = std::numeric_limits<double>::infinity();
    const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    // room: inside an axis-aligned box -> exit distance
    double tex = std::numeric_limits<double>::infinity();
    for (int k = 0; k < 3; ++k) {
      if (dd[k] > 1e-12) tex = std::min(tex, (room.hi[k] - oo[k]) / dd[k]);
      else if (dd[k] < -1e-12) tex = std::min(tex, (room.lo[k] - oo[k]) / dd[k]);
    }
    best = tex;
    for (int b = 0; b < 4; ++b) {  // slab test
      double t0 = 0, t1 = best; bool ok = true;
      for (int k = 0; k < 3 && ok; ++k) {
        if (std::fabs(dd[k]) < 1e-12) { if (oo[k] < obs[b].lo[k] || oo[k] > obs[b].hi[k]) ok = false; continue; }
        double ta = (obs[b].lo[k] - oo[k]) / dd[k], tb = (obs[b].hi[k] - oo[k]) / dd[k];
        if (ta > tb) std::swap(ta, tb);
        t0 = std::max(t0, ta); t1 = std::min(t1, tb);
        if (t0 > t1) ok = false;
      }
      if (ok && t0 > 1e-9) best = std::min(best, t0);
    }
    return best;
  }
};

const World& world() { static World w; return w; }

Pose lidar_pose_world(const synth_config& c, double t) {
  Kin k = gt_world(c, t);
  const World& w = world();
  return {mul(k.R, w.R_LI), k.p + mul(k.R, w.p_LI)};
}

}  // namespace

extern "C" {

struct synth_raw_point { float x, y, z, pad; float intensity; float pad2; double timestamp; };

int32_t synth_num_scans(const synth_config* c) { return static_cast<int32_t>(std::floor(c->duration * c->scan_rate + 1e-9)); }
double synth_scan_time(const synth_config* c, int32_t i) { return c->t_start + i / c->scan_rate; }

// GT IMU kinematics expressed in the estimator's world frame G = IMU frame at t_ref (the spline is anchored at
// identity there): out = p[3] q[4](x,y,z,w) v[3] a[3] w_body[3] ; gravity_G[3] = R_W_I(t_ref)^T (0,0,9.79)
void synth_gt_state(const synth_config* c, double t_ref, double t, double* out, double* gravity_G) {
  Kin k0 = gt_world(*c, t_ref), k = gt_world(*c, t);
  M3 R = mul(tr(k0.R), k.R);
  V3 p = mulT(k0.R, k.p - k0.p), v = mulT(k0.R, k.v), a = mulT(k0.R, k.a);
  double q[4]; to_quat(R, q);
  const double o[16] = {p.x, p.y, p.z, q[0], q[1], q[2], q[3], v.x, v.y, v.z, a.x, a.y, a.z, k.w_body.x, k.w_body.y, k.w_body.z};
  std::memcpy(out, o, sizeof(o));
  if (gravity_G) { V3 g = mulT(k0.R, V3{0, 0, 9.79}); gravity_G[0] = g.x; gravity_G[1] = g.y; gravity_G[2] = g.z; }
}

// raw organised scans [n_scans][rings][az_steps]; each point measured at its own firing time in the lidar frame AT
// THAT TIME (a real spinning sensor: the scan is motion-distorted).
void synth_scans(const synth_config* c, int32_t first_scan, int32_t n_scans, uint64_t seed, synth_raw_point* out) {
  const int H = c->rings, W = c->az_steps;
  const World& w = world();
  std::vector<double> vert(H);
  if (H == 16) { for (int r = 0; r < 16; ++r) vert[r] = (-15.0 + 2.0 * r) * M_PI / 180.0; }  // vlp_common.h:205-222 (sorted by elevation)
  else { for (int r = 0; r < H; ++r) vert[r] = (-24.8 + (26.8) * r / (H - 1)) * M_PI / 180.0; }
  const double az_dt = (1.0 / c->scan_rate) / W * (H == 16 ? 0.99533 : 1.0);  // 55.296 us at 10 Hz x 1800
  const double ring_dt = 2.304e-6;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int s = 0; s < n_scans; ++s)
    for (int wi = 0; wi < W; ++wi) {
      const int scan = first_scan + s;
      const double t_scan = synth_scan_time(c, scan);
      for (int h = 0; h < H; ++h) {
        const double t = t_scan + (wi + 0.5) * az_dt + h * ring_dt;  // firing centred in its azimuth slot: never on a knot boundary
        Pose L = lidar_pose_world(*c, t);
        const double az = -2.0 * M_PI * wi / W;  // clockwise like a Velodyne
        const double ce = std::cos(vert[h]);
        V3 dl = {ce * std::cos(az), ce * std::sin(az), std::sin(vert[h])};
        V3 dw = mul(L.R, dl);
        double rng = w.cast(L.p, dw);
        const uint64_t idx = (static_cast<uint64_t>(scan) * H + h) * W + wi;
        synth_raw_point& o = out[(static_cast<size_t>(s) * H + h) * W + wi];
        o.pad = 1.0f; o.pad2 = 0; o.intensity = static_cast<float>(10 + (h * 7 + wi) % 90);
        o.timestamp = t;
        if (!(rng >= 0.6 && rng <= 150.0)) { o.x = o.y = o.z = std::numeric_limits<float>::quiet_NaN(); continue; }  // vlp_common.h:186-187
        rng += c->range_noise * gauss(seed, 1, idx);
        o.x = static_cast<float>(rng * dl.x); o.y = static_cast<float>(rng * dl.y); o.z = static_cast<float>(rng * dl.z);
      }
    }
}

// C3: a room scene holds ~1.5 k occupied 0.5 m leaves, so the large NDT map of the 300 s configuration is synthesised
// directly in voxel space: `n_leaves` occupied leaves on a sparse lattice (every second cell of a cube, in the map frame = LiDAR frame
// at the first scan), each holding one planar patch (normal along a random axis, tilted by up to ~6 deg, offset by up to 0.1 m from the
// leaf centre, range noise sigma 0.01 m along the normal).  The scan stream re-draws the patches in time order: point g of the stream
// (scan-major, ring, azimuth; 16 x 1800 per scan) lies on leaf floor(g * n_leaves / n_points), so consecutive firings of a ring sweep
// one surface as on a real sensor.  Every raw point is the map point seen from the sensor's ground-truth pose at its own firing time,
// i.e. the stream is motion-distorted exactly like synth_scans' and the de-skew has real work to do.
//   raw_out [n_scans][rings][az_steps] PointXYZIT ; map_out (may be NULL) the same points in the map frame, PointXYZI-shaped 8 floats
void synth_lattice_scans(const synth_config* c, int64_t n_leaves, uint64_t seed, int32_t first_scan, int32_t n_scans, synth_raw_point* raw_out,
                         float* map_out) {
  const int H = c->rings, W = c->az_steps;
  const int total_scans = synth_num_scans(c);
  const int64_t n_points = static_cast<int64_t>(total_scans) * H * W;
  int side = 1;
  while (static_cast<int64_t>(side) * side * side < n_leaves) ++side;
  const double leaf = 0.5;
  const double org = -std::floor(side / 2.0) - 0.25 + 0.5;   // leaf centres at org + i (metres): cell centres of the absolute 0.5 m grid, never closer than 0.25 m to 0
  const double az_dt = (1.0 / c->scan_rate) / W * (H == 16 ? 0.99533 : 1.0);
  const double ring_dt = 2.304e-6;
  const Pose L0 = lidar_pose_world(*c, synth_scan_time(c, 0));
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int s = 0; s < n_scans; ++s)
    for (int h = 0; h < H; ++h) {
      const int scan = first_scan + s;
      const double t_scan = synth_scan_time(c, scan);
      for (int wi = 0; wi < W; ++wi) {
        const int64_t g = (static_cast<int64_t>(scan) * H + h) * W + wi;
        const int64_t l = static_cast<int64_t>((static_cast<__int128>(g) * n_leaves) / n_points);
        const int ix = static_cast<int>(l % side), iy = static_cast<int>((l / side) % side), iz = static_cast<int>(l / (static_cast<int64_t>(side) * side));
        const double ctr[3] = {org + ix, org + iy, org + iz};
        const int k = static_cast<int>(mix64(seed ^ (0x51ull << 40) ^ static_cast<uint64_t>(l)) % 3);   // dominant axis of the normal
        const double tu = 0.1 * (2 * uni(seed, 7, 4 * l) - 1), tv = 0.1 * (2 * uni(seed, 7, 4 * l + 1) - 1);   // slopes of the patch
        const double off = 0.1 * (2 * uni(seed, 7, 4 * l + 2) - 1);
        const double a = 0.22 * (2 * uni(seed, 8, 2 * g) - 1), b = 0.22 * (2 * uni(seed, 8, 2 * g + 1) - 1);
        const double hgt = off + tu * a + tv * b + c->range_noise * gauss(seed, 9, g) / std::sqrt(1.0 + tu * tu + tv * tv);
        double pm[3];
        pm[k] = ctr[k] + hgt; pm[(k + 1) % 3] = ctr[(k + 1) % 3] + a; pm[(k + 2) % 3] = ctr[(k + 2) % 3] + b;
        (void)leaf;
        const double t = t_scan + (wi + 0.5) * az_dt + h * ring_dt;
        const Pose Lk = lidar_pose_world(*c, t);
        const V3 pw = mul(L0.R, V3{pm[0], pm[1], pm[2]}) + L0.p;
        const V3 pr = mulT(Lk.R, pw - Lk.p);
        const size_t o = (static_cast<size_t>(s) * H + h) * W + wi;
        synth_raw_point& r = raw_out[o];
        r.x = static_cast<float>(pr.x); r.y = static_cast<float>(pr.y); r.z = static_cast<float>(pr.z);
        r.pad = 1.0f; r.pad2 = 0; r.intensity = static_cast<float>(10 + (h * 7 + wi) % 90); r.timestamp = t;
        if (map_out) {
          float* m = map_out + 8 * o;
          m[0] = static_cast<float>(pm[0]); m[1] = static_cast<float>(pm[1]); m[2] = static_cast<float>(pm[2]); m[3] = 1.0f;
          m[4] = r.intensity; m[5] = m[6] = m[7] = 0.0f;
        }
      }
    }
}

// LOAM-style poses: lidar pose at scan i expressed in the lidar frame of scan 0, row-major 4x4, with noise
void synth_loam_poses(const synth_config* c, int32_t n_scans, uint64_t seed, double* T44) {
  Pose L0 = lidar_pose_world(*c, synth_scan_time(c, 0));
  for (int i = 0; i < n_scans; ++i) {
    Pose L = lidar_pose_world(*c, synth_scan_time(c, i));
    M3 R = mul(tr(L0.R), L.R);
    V3 p = mulT(L0.R, L.p - L0.p);
    if (i > 0) {
      V3 ax = {gauss(seed, 2, 6 * i), gauss(seed, 2, 6 * i + 1), gauss(seed, 2, 6 * i + 2)};
      const double ang = std::sqrt(dot(ax, ax)) * c->loam_rot_noise;
      if (ang > 0) R = mul(R, axis_angle(ax, ang));
      p = p + V3{c->loam_pos_noise * gauss(seed, 2, 6 * i + 3), c->loam_pos_noise * gauss(seed, 2, 6 * i + 4), c->loam_pos_noise * gauss(seed, 2, 6 * i + 5)};
    }
    double* T = T44 + 16 * i;
    for (int r = 0; r < 3; ++r) { for (int cc = 0; cc < 3; ++cc) T[r * 4 + cc] = R.m[r * 3 + cc]; }
    T[3] = p.x; T[7] = p.y; T[11] = p.z; T[12] = T[13] = T[14] = 0; T[15] = 1;
  }
}

int32_t synth_num_imu(const synth_config* c) { return static_cast<int32_t>(std::floor((c->duration + 2 * c->pad_time + 0.2) * c->imu_rate)); }
// IMU samples over [t_start - pad - 0.1, ...): t[n], gyro[n*3], accel[n*3]
void synth_imu(const synth_config* c, uint64_t seed, double* t, double* gyro, double* accel) {
  const int n = synth_num_imu(c);
  double ql[4], pl[3], qc[4], pc[3], bg[3], ba[3];
  synth_gt_extrinsics(ql, pl, qc, pc, bg, ba);
  for (int i = 0; i < n; ++i) {
    const double ti = c->t_start - c->pad_time - 0.1 + kImuPhase + i / c->imu_rate;
    Kin k = gt_world(*c, ti);
    V3 f = mulT(k.R, k.a + V3{0, 0, 9.79});
    t[i] = ti;
    gyro[3 * i + 0] = k.w_body.x + bg[0] + c->gyro_noise * gauss(seed, 3, 6 * i);
    gyro[3 * i + 1] = k.w_body.y + bg[1] + c->gyro_noise * gauss(seed, 3, 6 * i + 1);
    gyro[3 * i + 2] = k.w_body.z + bg[2] + c->gyro_noise * gauss(seed, 3, 6 * i + 2);
    accel[3 * i + 0] = f.x + ba[0] + c->accel_noise * gauss(seed, 3, 6 * i + 3);
    accel[3 * i + 1] = f.y + ba[1] + c->accel_noise * gauss(seed, 3, 6 * i + 4);
    accel[3 * i + 2] = f.z + ba[2] + c->accel_noise * gauss(seed, 3, 6 * i + 5);
  }
}

// Camera: pinhole 1280x720 (cfg/lvi.yaml:54-78), rolling shutter readout 0.0666 s.
// Views = every keyframe_every-th frame. Landmarks on the room surfaces; a landmark's reference is the first view
// that sees it and it is tracked in at most max_track_views consecutive views.
// Returns the number of observations; arrays may be NULL for a sizing call.
//   view_t0[n_views] ; obs_view[n_obs], obs_landmark[n_obs], obs_uv[n_obs*2] ; lm_ref_obs[n_landmarks] (index into obs or -1),
//   lm_rho[n_landmarks] initial inverse depth (1/z in the reference camera frame, with relative noise)
int32_t synth_num_views(const synth_config* c) {
  const int frames = static_cast<int>(std::floor(c->duration * c->cam_rate + 1e-9));
  return (frames + c->keyframe_every - 1) / c->keyframe_every;
}
int64_t synth_camera(const synth_config* c, uint64_t seed, double* view_t0, int32_t* obs_view, int32_t* obs_landmark, double* obs_uv,
                     int32_t* lm_ref_obs, double* lm_rho, int64_t cap) {
  const double fx = 530.175, fy = 530.095, cx = 635.12, cy = 356.522, readout = 0.0666;
  const int rows = 720, cols = 1280;
  const World& w = world();
  const int nv = synth_num_views(c);
  const int nl = c->n_landmarks;
  // landmark positions: uniform over the 6 room faces by area
  std::vector<V3> X(nl);
  const double Lx = 12.0, Ly = 9.0, Lz = 3.5;
  const double areas[3] = {Lx * Ly, Lx * Lz, Ly * Lz};
  const double tot = 2 * (areas[0] + areas[1] + areas[2]);
  for (int l = 0; l < nl; ++l) {
    double u = uni(seed, 4, 3 * l) * tot, a = uni(seed, 4, 3 * l + 1), b = uni(seed, 4, 3 * l + 2);
    if (u < areas[0]) X[l] = {a * Lx, b * Ly, 0};
    else if (u < 2 * areas[0]) X[l] = {a * Lx, b * Ly, Lz};
    else if (u < 2 * areas[0] + areas[1]) X[l] = {a * Lx, 0, b * Lz};
    else if (u < 2 * areas[0] + 2 * areas[1]) X[l] = {a * Lx, Ly, b * Lz};
    else if (u < 2 * areas[0] + 2 * areas[1] + areas[2]) X[l] = {0, a * Ly, b * Lz};
    else X[l] = {Lx, a * Ly, b * Lz};
  }
  std::vector<int> first_view(nl, -1), n_track(nl, 0);
  int64_t n_obs = 0;
  for (int l = 0; l < nl; ++l) if (lm_ref_obs) lm_ref_obs[l] = -1;
  for (int v = 0; v < nv; ++v) {
    const double t0 = c->t_start + kCamPhase + (v * c->keyframe_every) / c->cam_rate;
    if (view_t0) view_t0[v] = t0;
    for (int l = 0; l < nl; ++l) {
      if (first_view[l] >= 0 && (v - first_view[l] >= c->max_track_views)) continue;
      // rolling shutter: fixed-point iteration on the row time
      double vv = cy, uu = 0, z = 0;
      bool ok = true;
      for (int itn = 0; itn < 3; ++itn) {
        const double t = t0 + vv * readout / rows;
        Kin k = gt_world(*c, t);
        M3 Rc = mul(k.R, w.R_CI);
        V3 pc = k.p + mul(k.R, w.p_CI);
        V3 Xc = mulT(Rc, X[l] - pc);
        z = Xc.z;
        if (!(z > 0.5 && z < 20.0)) { ok = false; break; }
        uu = fx * Xc.x / z + cx; vv = fy * Xc.y / z + cy;
        if (!(uu > 10 && uu < cols - 10 && vv > 10 && vv < rows - 10)) { ok = false; break; }  // border filter T:412-417
      }
      if (!ok) { if (first_view[l] >= 0) first_view[l] = -1000000; continue; }  // track lost: never re-acquired
      if (first_view[l] < -1) continue;
      // occlusion by the obstacles
      {
        const double t = t0 + vv * readout / rows;
        Kin k = gt_world(*c, t);
        V3 pc = k.p + mul(k.R, w.p_CI);
        V3 d = X[l] - pc; const double dist = std::sqrt(dot(d, d));
        const double hit = w.cast(pc, (1.0 / dist) * d);
        if (hit < dist - 1e-3) { if (first_view[l] >= 0) first_view[l] = -1000000; continue; }
      }
      const bool is_ref = first_view[l] == -1;
      if (is_ref) first_view[l] = v;
      const uint64_t oi = static_cast<uint64_t>(l) * nv + v;
      // the reference observation defines the bearing: keep it noise-free in the generator's GT sense but noisy as data
      const double un = uu + c->pixel_noise * gauss(seed, 5, 2 * oi), vn = vv + c->pixel_noise * gauss(seed, 5, 2 * oi + 1);
      if (obs_view && n_obs < cap) { obs_view[n_obs] = v; obs_landmark[n_obs] = l; obs_uv[2 * n_obs] = un; obs_uv[2 * n_obs + 1] = vn; }
      if (is_ref) {
        if (lm_ref_obs) lm_ref_obs[l] = static_cast<int32_t>(n_obs);
        if (lm_rho) lm_rho[l] = (1.0 / z) * (1.0 + c->rho_rel_noise * gauss(seed, 6, l));
      }
      ++n_track[l];
      ++n_obs;
    }
  }
  return n_obs;
}

}  // extern "C"
