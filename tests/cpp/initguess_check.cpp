// The reference's initial-guess call sequence (LIinitializer::ComputeIntegrationForFrames / EstimateInitExtrinsicLI / EstimateInitExtrinsicCI,
// L/test/lvi_initialize_surfel_orb.cpp:952-1148) against the vi_init headers under include/lvi_exc_b200/compat.  Host code only.
// usage: initguess_check <blob>   -> one JSON line
#include <cstdint>
#include <cstdio>
#include <deque>
#include <memory>
#include <vector>

#include <vi_init/initial_alignment.h>
#include <vi_init/initial_ex_rotation.h>
#include <vi_init/integration_base.h>

struct IntegrationFrame {
  double timestamp = 0;   // seconds (the reference keeps ns and multiplies by 1e-9)
  Eigen::Matrix4d Tcw;
  std::shared_ptr<IntegrationBase> integrator;
};
struct Imu { double t; Eigen::Vector3d gyr, acc; };

static void PopOldIMU(double stamp, std::deque<Imu>& imu_cache) {
  while (!imu_cache.empty() && imu_cache.front().t < stamp) imu_cache.pop_front();
}
static void RemoveOverTimeFrames(const std::deque<Imu>& imu_cache, std::deque<IntegrationFrame>& frames) {
  while (!frames.empty() && frames.front().timestamp < imu_cache.front().t) frames.pop_front();
  while (!frames.empty() && frames.back().timestamp > imu_cache.back().t) frames.pop_back();
}
static void ComputeIntegrationForFrames(std::deque<IntegrationFrame>& integration_frames, std::deque<Imu> imu_cache) {
  RemoveOverTimeFrames(imu_cache, integration_frames);
  const size_t size = integration_frames.size();
  Eigen::Vector3d zero_vec(0, 0, 0);
  double last_imu_time = -1.;
  for (size_t i = 1; i < size; ++i) {
    auto& intframe = integration_frames[i];
    PopOldIMU(integration_frames[i - 1].timestamp, imu_cache);
    std::shared_ptr<IntegrationBase> integrator = std::make_shared<IntegrationBase>(zero_vec, zero_vec, IMUNoise());
    if (!imu_cache.empty()) {
      const double cur_stamp = intframe.timestamp;
      while (1) {
        if (imu_cache.empty()) break;
        if (imu_cache.front().t < cur_stamp) {
          const Imu imu_msg = imu_cache.front();
          if (last_imu_time < 0) integrator->push_back(0.001, imu_msg.acc, imu_msg.gyr);
          else integrator->push_back(imu_msg.t - last_imu_time, imu_msg.acc, imu_msg.gyr);
          last_imu_time = imu_msg.t;
          imu_cache.pop_front();
        } else break;
      }
    } else return;
    intframe.integrator = integrator;
  }
}

struct Result { bool rot_ok = false, ok = false; Eigen::Matrix3d R; Eigen::Vector3d T; Eigen::Vector4d g_t; double scale = 1.0; };

static Result EstimateInitExtrinsic(std::deque<IntegrationFrame>& frames, bool fix_scale) {
  Result out;
  InitialEXRotation rot_estimator_;
  for (size_t i = 0; i + 1 < frames.size(); ++i) {
    Eigen::Matrix3d relative_rot = frames[i].Tcw.block<3, 3>(0, 0).transpose() * frames[i + 1].Tcw.block<3, 3>(0, 0);
    out.rot_ok = rot_estimator_.CalibrationExRotationLiDAR(relative_rot, frames[i + 1].integrator->delta_q, out.R);
    if (out.rot_ok) break;
  }
  if (!out.rot_ok) return out;
  const int winsize = 10;
  Eigen::Vector3d Bgs, g;
  Eigen::VectorXd x;
  for (size_t i = winsize; i < frames.size(); ++i) {
    Eigen::Matrix4d T_inv = frames[i - winsize].Tcw.inverse();
    std::deque<ImageFrame> win_frames = {};
    for (size_t k = i - winsize; k < i; ++k) {
      Eigen::Matrix4d Tk = T_inv * frames[k].Tcw;
      ImageFrame kf;
      kf.R = Tk.block<3, 3>(0, 0) * out.R.transpose();
      kf.T = Tk.block<3, 1>(0, 3).col(0);
      kf.pre_integration = frames[k].integrator.get();
      win_frames.push_back(kf);
    }
    out.ok = VisualIMUAlignment(win_frames, &Bgs, g, out.T, x, fix_scale);
    if (out.ok) {
      const Eigen::Vector3d gi = out.R * g;
      out.g_t = Eigen::Vector4d(gi(0), gi(1), gi(2), frames[i - winsize].timestamp);
      if (!fix_scale) out.scale = (x.tail<1>())(0);
      return out;
    }
  }
  return out;
}

static std::vector<double> read_doubles(FILE* f, size_t n) {
  std::vector<double> v(n);
  if (n && fread(v.data(), 8, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}
static std::deque<IntegrationFrame> frames_of(const std::vector<double>& stamps, const std::vector<double>& poses) {
  std::deque<IntegrationFrame> fr;
  for (size_t i = 0; i < stamps.size(); ++i) {
    IntegrationFrame f;
    f.timestamp = stamps[i];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) f.Tcw(r, c) = poses[16 * i + 4 * r + c];
    fr.push_back(f);
  }
  return fr;
}
static void print(const char* name, const Result& r, bool last) {
  const Eigen::Quaterniond q(r.R);
  printf("\"%s\": {\"rot_ok\": %d, \"ok\": %d, \"q\": [%.17g, %.17g, %.17g, %.17g], \"T\": [%.17g, %.17g, %.17g], \"g\": [%.17g, %.17g, %.17g, %.17g], \"scale\": %.17g}%s",
         name, r.rot_ok, r.ok, q.x(), q.y(), q.z(), q.w(), r.T(0), r.T(1), r.T(2), r.g_t(0), r.g_t(1), r.g_t(2), r.g_t(3), r.scale, last ? "" : ", ");
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int32_t hdr[4];
  if (fread(hdr, 4, 4, f) != 4 || hdr[0] != 0x4C564932) return 2;
  const int nl = hdr[1], nc = hdr[2], ni = hdr[3];
  const auto ls = read_doubles(f, nl), lp = read_doubles(f, 16 * static_cast<size_t>(nl)), cs = read_doubles(f, nc), cp = read_doubles(f, 16 * static_cast<size_t>(nc));
  const auto it = read_doubles(f, ni), gy = read_doubles(f, 3 * static_cast<size_t>(ni)), ac = read_doubles(f, 3 * static_cast<size_t>(ni));
  fclose(f);
  std::deque<Imu> imu;
  for (int i = 0; i < ni; ++i) imu.push_back(Imu{it[i], Eigen::Vector3d(gy[3 * i], gy[3 * i + 1], gy[3 * i + 2]), Eigen::Vector3d(ac[3 * i], ac[3 * i + 1], ac[3 * i + 2])});
  auto fl = frames_of(ls, lp), fc = frames_of(cs, cp);
  ComputeIntegrationForFrames(fl, imu);
  ComputeIntegrationForFrames(fc, imu);
  const Result rl = EstimateInitExtrinsic(fl, true), rc = EstimateInitExtrinsic(fc, false);
  printf("{");
  print("lidar", rl, false);
  print("camera", rc, true);
  printf("}\n");
  return 0;
}
