// PinholeCamera of the compat headers: Project / Unproject with the radial-tangential distortion (K/sensors/pinhole_camera.h:96-240).  Host only.
#include <cstdio>

#include <kontiki/sensors/pinhole_camera.h>

int main() {
  kontiki::sensors::PinholeCamera cam(720, 1280, 0.0666, -0.28, 0.07, 1.0e-3, -5.0e-4, 0.01, 530.175, 530.095, 635.12, 356.522);
  kontiki::sensors::PinholeCamera plain(720, 1280, 0.0666, 0, 0, 0, 0, 0, 530.175, 530.095, 635.12, 356.522);
  const double pts[4][3] = {{0.3, -0.2, 2.0}, {-1.1, 0.4, 3.0}, {0.05, 0.6, 1.5}, {0.0, 0.0, 4.0}};
  printf("{\"do_distortion\": [%d, %d], \"points\": [", cam.do_distortion() ? 1 : 0, plain.do_distortion() ? 1 : 0);
  for (int i = 0; i < 4; ++i) {
    const Eigen::Vector3d X(pts[i][0], pts[i][1], pts[i][2]);
    const Eigen::Vector2d y = cam.Project(X), y0 = plain.Project(X);
    const Eigen::Vector3d r = cam.Unproject(y), r0 = plain.Unproject(y0);
    printf("%s{\"y\": [%.17g, %.17g], \"y_plain\": [%.17g, %.17g], \"ray\": [%.17g, %.17g, %.17g], \"ray_plain\": [%.17g, %.17g, %.17g]}", i ? ", " : "", y(0), y(1), y0(0),
           y0(1), r(0), r(1), r(2), r0(0), r0(1), r0(2));
  }
  printf("]}\n");
  return 0;
}
