// setup_kernels.cu — FP64 setup solvers on the device (compiled with -fmad=false so that double
// arithmetic is evaluated as written, like the reference's host code).
//
// Reference: camera_model_specific_setup /root/reference/src/lentil.h:1568-1670.  Its two searches
// are embarrassingly parallel: logarithmic_focus_search evaluates 20 001 independent sensor shifts
// (lentil.h:1445-1460, lens.h:395-407) and trace_backwards_for_fstop 999 independent ray heights
// (lentil.h:1390-1441).  Each candidate is one thread here; the (order-dependent) selection among
// the candidates stays on the host (lentil_host.cpp).
#include "lens_device.cuh"
#include "lentil_internal.h"

namespace lb {

// line_plane_intersection(...)(2), lens.h:412-419
__device__ double line_plane_z(const double o[3], const double d_in[3]) {
  const double dn = sqrt(d_in[0] * d_in[0] + d_in[1] * d_in[1] + d_in[2] * d_in[2]);
  const double d[3] = {d_in[0] / dn, d_in[1] / dn, d_in[2] / dn};
  const double cn = sqrt(100.0 * 100.0 + 0.0 * 0.0 + 100.0 * 100.0);
  const double coord_y = 0.0 / cn;
  const double s = coord_y - o[1];  // coord.dot(planeNormal) - planeNormal.dot(rayOrigin), planeNormal = (0,1,0)
  return o[2] + (d[2] * s) / d[1];
}

__global__ void k_focus_distances(const __grid_constant__ LensTable lens, const __grid_constant__ CamConsts<double> cam,
                                  double aperture_y, const double *__restrict__ shifts, int n, double4 *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const TableEval<double> ev(lens);
  const double shift = shifts[i];
  double x = 0.0, y = 0.0, dx = 0.0, dy = 0.0;
  pt_sample_aperture(ev, x, y, dx, dy, cam.lambda, 0.0, aperture_y, shift);
  x += dx * shift;
  y += dy * shift;
  const double b[5] = {x, y, dx, dy, cam.lambda};
  double o[4];
  ev.out4(b, o);
  const double T = t_max(0.0, ev.transmittance(b));
  double pos[3], dir[3];
  outer_to_cs(cam, o, pos, dir);
  const double px = x + dx * cam.bfl, py = y + dy * cam.bfl;
  out[i] = make_double4(line_plane_z(pos, dir), T, o[0] * o[0] + o[1] * o[1], px * px + py * py);
}

__global__ void k_fstop_rays(const __grid_constant__ LensTable lens, const __grid_constant__ CamConsts<double> cam, int n,
                             double outer_pupil_radius, double4 *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || i < 1) return;
  const TableEval<double> ev(lens);
  const double h = ((double)i / (double)n) * outer_pupil_radius;
  const double scene[3] = {0.0, h, (double)1.0e12f};  // AI_BIG
  double sensor[4], o[4];
  const double T = lt_sample_aperture(ev, cam, scene, 0.01, h, cam.lambda, sensor, o, (int *)nullptr);
  double valid = T > 0.0 ? 1.0 : 0.0;
  const double px = sensor[0] + sensor[2] * cam.bfl, py = sensor[1] + sensor[3] * cam.bfl;
  if (px * px + py * py > cam.inner_pupil_r2) valid = 0.0;
  double pos[3], dir[3];
  const double center = -cam.inner_R + cam.bfl;
  if (cam.inner_geom == 0) sphere_to_cs_center(o[0], o[1], o[2], o[3], center, cam.inner_R, pos, dir);
  else cylinder_to_cs(o[0], o[1], o[2], o[3], center, cam.inner_R, cam.inner_geom == 1, pos, dir);
  out[i] = make_double4(valid, pos[1], pos[2], 0.0);
}

cudaError_t launch_focus_distances(const LensTable &lens, const CamConsts<double> &cam, double aperture_y, const double *shifts,
                                   int n, double4 *out, cudaStream_t stream) {
  k_focus_distances<<<(n + 127) / 128, 128, 0, stream>>>(lens, cam, aperture_y, shifts, n, out);
  return cudaGetLastError();
}
cudaError_t launch_fstop_rays(const LensTable &lens, const CamConsts<double> &cam, int n, double outer_pupil_radius, double4 *out,
                              cudaStream_t stream) {
  k_fstop_rays<<<(n + 63) / 64, 64, 0, stream>>>(lens, cam, n, outer_pupil_radius, out);
  return cudaGetLastError();
}

}  // namespace lb
