// Coordinate front-end (row f1 of SURVEY section 8): particle positions / velocities in the
// galaxy's Cartesian frame -> what the hot path reads.  One fused O(N) pass instead of the
// reference's SkyCoord machinery (sph_source.py:265-362) and the host numpy mirror of it:
//   rotate to the source's (ra, dec), translate by its distance, add the peculiar velocity along
//   the line of sight and the Hubble flow of every particle (sph_source.py:288-316);
//   spherical representation: RA, Dec, distance, radial velocity (:318-326);
//   WCS of the cube (RA---TAN / DEC--TAN / VRAD or FREQ, datacube.py:426-486): 0-indexed pixel
//   coordinates, pad included (:351-361);
//   smoothing lengths in pixels: arctan(hsm / D) through the pixel scale (sph_kernels.py:250-253).
// ICRS frame and specsys only (as the host mirror).  HBM-bound: 56 B read, 48 B written per particle.
#pragma once

#include "common.cuh"

namespace mtn {

struct FrontEndArgs {
  int64_t n;
  const double* xyz;    // (n, 3) kpc
  const double* vxyz;   // (n, 3) km/s
  const double* hsm;    // (n) kpc or null
  double hsm_scalar;
  double R[9];          // rotation to the source's direction, row-major
  double unit[3];       // unit vector towards (ra, dec)
  double distance_kpc, vpeculiar, hubble;  // hubble = h * 100 km/s/Mpc
  double sin_d0, cos_d0, a0;               // cube centre (radians)
  double rad_to_px;                        // 1 / (px_size in radians)
  double crpix_x, crpix_y, crpix_z;        // 1-indexed reference pixels (pad included)
  int freq_mode;                           // 1: FREQ axis (Hz), 0: VRAD (km/s)
  double spectral_centre, channel_width;   // in the axis' own unit
  double* px;
  double* py;
  double* pz;
  double* v;
  double* D;
  double* sm_length;
};

__global__ void __launch_bounds__(256) sky_to_pix_kernel(FrontEndArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double x0 = a.xyz[3 * i], y0 = a.xyz[3 * i + 1], z0 = a.xyz[3 * i + 2];
  const double u0 = a.vxyz[3 * i], v0 = a.vxyz[3 * i + 1], w0 = a.vxyz[3 * i + 2];
  const double shift = a.distance_kpc;
  const double x = a.R[0] * x0 + a.R[1] * y0 + a.R[2] * z0 + a.unit[0] * shift;
  const double y = a.R[3] * x0 + a.R[4] * y0 + a.R[5] * z0 + a.unit[1] * shift;
  const double z = a.R[6] * x0 + a.R[7] * y0 + a.R[8] * z0 + a.unit[2] * shift;
  // peculiar velocity along the line of sight, then the Hubble flow of the particle's position
  const double vx = a.R[0] * u0 + a.R[1] * v0 + a.R[2] * w0 + a.unit[0] * a.vpeculiar + a.hubble * (x * 1.0e-3);
  const double vy = a.R[3] * u0 + a.R[4] * v0 + a.R[5] * w0 + a.unit[1] * a.vpeculiar + a.hubble * (y * 1.0e-3);
  const double vz = a.R[6] * u0 + a.R[7] * v0 + a.R[8] * w0 + a.unit[2] * a.vpeculiar + a.hubble * (z * 1.0e-3);
  const double r = sqrt(x * x + y * y + z * z);
  const double ra = atan2(y, x), dec = asin(z / r);
  const double vr = (x * vx + y * vy + z * vz) / r;
  const double dist = r * 1.0e-3;  // Mpc
  // gnomonic (TAN) projection about the cube centre, wcs_world2pix(..., origin=0)
  double sd, cd, sa, ca;
  sincos(dec, &sd, &cd);
  sincos(ra - a.a0, &sa, &ca);
  const double cosc = a.sin_d0 * sd + a.cos_d0 * cd * ca;
  const double xi = cd * sa / cosc, eta = (a.cos_d0 * sd - a.sin_d0 * cd * ca) / cosc;
  a.px[i] = -xi * a.rad_to_px + a.crpix_x - 1.0;
  a.py[i] = eta * a.rad_to_px + a.crpix_y - 1.0;
  if (a.freq_mode) {
    const double f = 1.420405751e9 * (1.0 - vr / 299792.458);
    a.pz[i] = (f - a.spectral_centre) / a.channel_width + a.crpix_z - 1.0;
  } else {
    a.pz[i] = -(vr - a.spectral_centre) / a.channel_width + a.crpix_z - 1.0;
  }
  a.v[i] = vr;
  a.D[i] = dist;
  const double h = a.hsm ? a.hsm[i] : a.hsm_scalar;
  a.sm_length[i] = atan(h / dist * 1.0e-3) * a.rad_to_px;
}

}  // namespace mtn
