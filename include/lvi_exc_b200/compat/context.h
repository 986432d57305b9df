// context.h — the process-wide default device context of the compat headers and the status -> exception mapping
#ifndef LVI_EXC_B200_COMPAT_CONTEXT_H
#define LVI_EXC_B200_COMPAT_CONTEXT_H
#include <stdexcept>
#include <string>

#include "../../lvi_exc_b200.h"

namespace lvi_exc_b200 {
// The context every estimator / map object of this process works on unless told otherwise (the reference has no such notion: Ceres and
// PCL are process-global too).  Created lazily on device 0; SetDefaultContext installs a caller-owned one (another device, NCCL).
inline lvi_ctx*& default_context_slot() { static lvi_ctx* ctx = nullptr; return ctx; }
inline void SetDefaultContext(lvi_ctx* ctx) { default_context_slot() = ctx; }
inline lvi_ctx* DefaultContext() {
  lvi_ctx*& c = default_context_slot();
  if (!c) {
    const int rc = lvi_ctx_create(0, nullptr, 0, 1, &c);
    if (rc != LVI_OK) throw std::runtime_error(std::string("lvi_exc_b200: cannot create a CUDA context: ") + lvi_last_error());
  }
  return c;
}
inline void throw_status(int rc) {
  if (rc == LVI_OK) return;
  const std::string msg = lvi_last_error();
  if (rc == LVI_ERR_RANGE) throw std::range_error(msg);     // K/kontiki/trajectory_estimator.h:111-122, spline_base.h:221
  if (rc == LVI_ERR_DOMAIN) throw std::domain_error(msg);   // K/kontiki/trajectories/uniform_so3_spline_trajectory.h:23-27
  throw std::runtime_error(msg);
}
}  // namespace lvi_exc_b200
#endif
