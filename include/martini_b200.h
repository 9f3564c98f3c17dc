/*
 * martini_b200 -- C ABI of the B200-native particle->datacube projection.
 *
 * This is the drop-in boundary for ONE hot path of MARTINI (kyleaoman/martini,
 * astromartini 2.1.18): the work of Martini.insert_source_in_cube.  The reference has no
 * FFI of its own (it is pure Python); each entry point below names the reference function
 * it replaces (file:line relative to the reference tree).  The Python host side
 * (martini_b200/_lib.py) binds these with ctypes; INTEGRATION.md shows the stub a MARTINI
 * maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer (e.g. torch.Tensor.data_ptr()) unless named *_host;
 *  - the library never owns device memory: scratch comes from caller-provided workspaces
 *    whose sizes are queried first, so the caller's allocator (torch) owns everything;
 *  - all work is enqueued on the caller's CUDA stream; only mtn_plan synchronises (it
 *    returns its counters to the host);
 *  - functions return 0 on success, a negative MTN_ERR_* code otherwise, and
 *    mtn_last_error() then returns a thread-local message;
 *  - units: positions / smoothing lengths in pixels (pad included, 0-indexed pixel
 *    centres at integers), velocities and channel edges in km/s, distances in Mpc, HI
 *    masses in Msun, pixel size in arcsec; the cube is float64 (nx, ny, n_channels),
 *    channel fastest, Jy/arcsec^2 on return.
 */
#ifndef MARTINI_B200_H
#define MARTINI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTN_VERSION 200 /* 0.2.0 */

/* error codes */
#define MTN_OK 0
#define MTN_ERR_INVALID -1   /* bad argument */
#define MTN_ERR_CUDA -2      /* CUDA runtime error */
#define MTN_ERR_WORKSPACE -3 /* workspace too small */
#define MTN_ERR_LIMIT -4     /* problem exceeds a 32-bit index limit */

/* SPH kernel kinds -- the primitive kernels of martini/sph_kernels.py */
#define MTN_KERNEL_WENDLANDC2 0    /* _WendlandC2Kernel   :340-478   */
#define MTN_KERNEL_WENDLANDC6 1    /* _WendlandC6Kernel   :481-721   */
#define MTN_KERNEL_CUBICSPLINE 2   /* _CubicSplineKernel  :724-894   */
#define MTN_KERNEL_GAUSSIAN 3      /* _GaussianKernel     :897-1080  */
#define MTN_KERNEL_DIRACDELTA 4    /* DiracDeltaKernel    :1083-1204 */
#define MTN_KERNEL_QUARTICSPLINE 5 /* _QuarticSplineKernel:1402-1600 */

/* spectral models -- martini/spectral_models.py */
#define MTN_SPECTRUM_GAUSSIAN 0   /* GaussianSpectrum.spectral_function   :347-425 */
#define MTN_SPECTRUM_DIRACDELTA 1 /* DiracDeltaSpectrum.spectral_function :510-570 */

/* prune flags -- the keyword arguments of _BaseMartini._prune_particles, martini.py:168-173 */
#define MTN_PRUNE_SPATIAL 1
#define MTN_PRUNE_SPECTRAL 2
#define MTN_PRUNE_MASS 4

/* cube flags */
#define MTN_CUBE_ACCUMULATE 0 /* out = (in + inserted) / px_area  (martini.py:338,364-366) */
#define MTN_CUBE_ZEROED 1     /* caller guarantees the slab is all zeros: voxels no particle
                                 reaches are left untouched, the rest are written once */

#define MTN_MAX_KERNELS 8

/* One entry per primitive kernel in use.  A simple kernel has one entry; the adaptive
 * kernels (_AdaptiveKernel and WendlandC2Kernel/... , sph_kernels.py:1207-1903) have
 * one per member of `.kernels`, in priority order.  All values are read from the host
 * kernel objects (the FWHM rescale comes from a numeric fsolve there, :14-48). */
typedef struct MtnKernelEntry {
  int32_t kind;         /* MTN_KERNEL_* */
  int32_t valid_is_max; /* 1: valid iff sm_length*rescale <= valid_size (DiracDelta :1189);
                           0: valid iff sm_length*rescale >= valid_size (e.g. :461) */
  double rescale;       /* K._rescale */
  double size_in_fwhm;  /* K.size_in_fwhm (may be +inf for DiracDelta, martini.py:1537) */
  double valid_size;    /* K.min_valid_size or K.max_valid_size */
  double truncate;      /* _GaussianKernel.truncate, else 0 */
  double norm;          /* _GaussianKernel.norm (:949-951), else 1 */
} MtnKernelEntry;

typedef struct MtnKernelTable {
  int32_t n;        /* number of entries, 1..MTN_MAX_KERNELS */
  int32_t adaptive; /* 1: choose per particle the first entry that validates (:1254-1272) */
  MtnKernelEntry k[MTN_MAX_KERNELS];
} MtnKernelTable;

/* Particle arrays (length n, float64).  For sigma, mHI and D a NULL pointer selects the
 * scalar beside it (the reference allows scalar mHI_g / T_g / sigma). */
typedef struct MtnParticles {
  int64_t n;
  const double* px;        /* source.pixcoords[0], sph_source.py:351-361 */
  const double* py;        /* source.pixcoords[1] */
  const double* h_eff;     /* sm_length * rescale: the h handed to _kernel_integral (:116) */
  const double* sm_range;  /* sph_kernel.sm_ranges (:257-262) */
  const uint8_t* kernel_id;/* sph_kernel.kernel_indices with -1 -> 0 (:1360); NULL = all 0 */
  const double* v;         /* source.skycoords.radial_velocity [km/s], spectral_models.py:92 */
  const double* sigma;     /* spectral_model.half_width(source) [km/s] or NULL */
  double sigma_scalar;
  const double* mHI;       /* source.mHI_g [Msun] or NULL */
  double mHI_scalar;
  const double* D;         /* source.skycoords.distance [Mpc] (:93) or NULL */
  double D_scalar;
  const uint8_t* accept;   /* optional prune mask from mtn_prune (1 = keep); NULL = keep all */
} MtnParticles;

/* The cube slab this call fills: rows [x_lo, x_hi) of the full (nx, ny, n_channels) cube.
 * Single GPU: x_lo = 0, x_hi = nx.  Multi-GPU: one contiguous x-slab per rank. */
typedef struct MtnCube {
  int32_t nx, ny, n_channels; /* full padded cube: datacube._array.shape[:3] */
  int32_t x_lo, x_hi;         /* slab rows owned by this call */
  int32_t spectrum;           /* MTN_SPECTRUM_* */
  int32_t flags;              /* MTN_CUBE_* */
  double px_size_arcsec;      /* datacube.px_size: final Jy/pix^2 -> Jy/arcsec^2 */
  const double* edges;        /* n_channels+1 velocity_channel_edges [km/s], monotone */
  double* slab;               /* (x_hi-x_lo, ny, n_channels) float64, in/out */
  int32_t edges_direction;    /* +1: edges increase with channel index, -1: decrease (the usual
                                 velocity-mode cube, datacube.py:469-475), 0: unknown -- the
                                 library then reads two edges back and synchronises */
  int32_t reserved;
} MtnCube;

/* What mtn_plan reports back (host memory). */
typedef struct MtnPlan {
  int64_t n_kept;          /* particles that reach at least one voxel of the slab */
  int64_t n_pairs;         /* (particle, brick) pairs to be sorted */
  int64_t n_bricks;        /* bricks in the slab grid (tiles x channel blocks) */
  int64_t updates_dense;   /* U_dense: (particle,pixel,channel) terms the reference loop
                              executes for this slab = C * sum_p n_x(p) n_y(p) */
  int64_t chunk;           /* particles per work item chosen for mtn_project */
  int64_t n_pairs2;        /* pairs of the second stream: (particle, pixel x channel superblock)
                              for DiracDelta-kernel particles under a Gaussian line (column
                              kernel), (particle, tile x channel) for every particle under a
                              DiracDelta spectrum (splat kernel); 0 if the insertion has none */
  int64_t chunk2;          /* particles per work item of the second stream */
  int32_t edges_increasing;/* 1 if channel edges increase with channel index */
  int32_t route2;          /* 0: none, 1: column kernel, 2: splat kernel */
  size_t workspace_bytes;  /* device workspace mtn_project needs */
} MtnPlan;

int mtn_version(void);
const char* mtn_last_error(void);

/* Number of SMs / device name of the current device (diagnostics for bench.py). */
int mtn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/*
 * K0 -- smoothing setup.  Replaces _BaseSPHKernel._init_sm_ranges (sph_kernels.py:257-262)
 * and, for adaptive kernels, the per-particle selection of
 * _AdaptiveKernel._init_sm_lengths (:1254-1272).  Outputs, per particle:
 *   kernel_id_out  first table entry whose _validate passes, 0 if none (:1264-1272, :1360)
 *   valid_out      1 if some entry validated (what _AdaptiveKernel._validate reports, :1383;
 *                  for a simple kernel: its own _validate) -- may be NULL
 *   sm_range_out   ceil(sm_length * size_in_fwhm[kernel])
 *   h_eff_out      sm_length * rescale[kernel]
 */
int mtn_smoothing_setup(int64_t n, const double* sm_length, const MtnKernelTable* table,
                        uint8_t* kernel_id_out, uint8_t* valid_out, double* sm_range_out,
                        double* h_eff_out, void* stream);

/*
 * K1 -- prune.  Replaces the accept-mask computation of _BaseMartini._prune_particles
 * (martini.py:198-232), bit-exact: a particle is rejected if
 *   spatial : px+r<0 or py+r<0 or px-r>nx_tot or py-r>ny_tot or isnan(px) or isnan(py)
 *   spectral: pz+4w<0 or pz-4w>n_channels, w = half_width/max_abs_dv
 *   mass    : mHI == 0
 * half_width NULL selects half_width_scalar; mHI NULL selects mHI_scalar.
 * n_accept_out (device int64, may be NULL) receives the number kept.
 */
int mtn_prune(int64_t n0, const double* px, const double* py, const double* pz,
              const double* sm_range, const double* mHI, double mHI_scalar,
              const double* half_width, double half_width_scalar, double max_abs_dv,
              int32_t nx_tot, int32_t ny_tot, int32_t n_channels, int32_t flags,
              uint8_t* accept_out, int64_t* n_accept_out, void* stream);

/* Bytes of device scratch mtn_plan needs for n particles and this cube slab.  The same
 * scratch buffer must be handed, untouched, to the mtn_project call that follows. */
size_t mtn_plan_scratch_bytes(int64_t n, const MtnCube* cube);

/*
 * Plan the projection of `p` into `cube`: footprints, live channel windows, brick
 * overlap counts.  Synchronises the stream and fills *plan_host.  Replaces the
 * O(n_pix * N) candidate scan of _evaluate_pixel_spectrum (martini.py:272-274) by an
 * O(N) footprint pass that also routes every particle to the kernel that computes it (the
 * brick kernel; the column kernel for DiracDelta-kernel particles, which reach one pixel; the
 * splat kernel under a DiracDelta spectrum, where a particle reaches one channel) -- which is
 * why it needs the kernel table.  Limits (MTN_ERR_LIMIT): fewer than 2^32 - 1 pairs per stream
 * and kept particles per slab, at most 65535 channels.
 */
int mtn_plan(const MtnParticles* p, const MtnKernelTable* table, const MtnCube* cube, void* scratch,
             size_t scratch_bytes, MtnPlan* plan_host, void* stream);

/*
 * Project.  Replaces spectral_model.init_spectra (spectral_models.py:63-147), the pixel
 * loop of _insert_source_in_cube (martini.py:327-362) with _evaluate_pixel_spectrum
 * (:271-283) and sph_kernel._px_weight (sph_kernels.py:85-119), and the final unit
 * conversion (martini.py:364-366):
 *   slab[i,j,c] = (slab[i,j,c] + sum_p W_p(i,j) * S_p(c)) / px_size_arcsec^2
 * `plan` must come from mtn_plan on the same inputs.  Asynchronous on `stream`.
 */
int mtn_project(const MtnParticles* p, const MtnKernelTable* table, const MtnCube* cube,
                const MtnPlan* plan, void* scratch, size_t scratch_bytes, void* workspace,
                size_t workspace_bytes, void* stream);

/* Number of projection-path kernel launches the last mtn_project on this thread made. */
int mtn_last_launch_count(void);

/*
 * Coordinate front-end: replaces SPHSource._init_skycoords and
 * _init_pixcoords (martini/sources/sph_source.py:265-362: rotate to (ra, dec), translate by the
 * distance, peculiar velocity + Hubble flow, spherical representation, WCS pixel coordinates)
 * and _BaseSPHKernel._init_sm_lengths (martini/sph_kernels.py:235-255) for the ICRS frame /
 * specsys and the cube's RA---TAN / DEC--TAN / VRAD-or-FREQ system (martini/datacube.py:426-486).
 * xyz, vxyz: (n, 3) row-major, kpc and km/s in the galaxy's frame (after SPHSource.rotate);
 * hsm: kpc (NULL selects hsm_scalar).  Outputs (device, length n): pixel coordinates (0-indexed,
 * pad included), radial velocity [km/s], distance [Mpc], smoothing length [pixels].
 */
typedef struct MtnFrontEnd {
  double rotation[9];      /* row-major: galaxy frame -> frame whose x axis points to (ra, dec) */
  double direction[3];     /* unit vector towards (ra, dec) */
  double distance_mpc;     /* source.distance */
  double vpeculiar;        /* km/s, along the line of sight */
  double hubble;           /* h * 100 km/s/Mpc */
  double ra0_rad, dec0_rad;/* cube centre (datacube.ra, .dec) */
  double px_size_arcsec;   /* datacube.px_size */
  double crpix[3];         /* 1-indexed reference pixels: n/2 + 0.5 (+ pad) (datacube.py:469-475) */
  double spectral_centre;  /* km/s (VRAD) or Hz (FREQ) */
  double channel_width;    /* |cdelt3| in the same unit */
  int32_t freq_mode;       /* 1: FREQ axis (frequency increases with channel), 0: VRAD */
  int32_t reserved;
} MtnFrontEnd;
int mtn_sky_to_pix(const MtnFrontEnd* fe, int64_t n, const double* xyz, const double* vxyz,
                   const double* hsm, double hsm_scalar, double* px, double* py, double* pz, double* v,
                   double* D, double* sm_length, void* stream);

/*
 * Multi-GPU input exchange: route particles to the ranks whose x-slab their candidate box can
 * reach (martini.py:272-274 decides per pixel; a slab needs every particle with
 * [px - r, px + r] intersecting its rows -- halo particles go to both neighbours).  The reference
 * has no counterpart (one process, threads over pixels, martini.py:345-362).
 *   mtn_route_count    per destination rank d (slab rows [bounds[d], bounds[d+1])): how many of
 *                      the n particles go there -> totals_out[world] (device int64); leaves the
 *                      per-block offsets mtn_route_scatter needs in `scratch`.
 *   mtn_route_scatter  stores every routed particle's n_fields quantities into the destination
 *                      rank's inbox: inboxes[d] is a DEVICE pointer valid on this device -- the
 *                      peer-mapped inbox of rank d (symmetric memory over NVLink), laid out
 *                      [field][capacity] -- at position src_offsets[d] (device int64[world]: the
 *                      totals of the lower-numbered source ranks for d) + the particle's rank
 *                      among this source's particles for d.  Order inside an inbox: source rank,
 *                      then particle index.  fields / inboxes are HOST arrays of device pointers.
 * The destination test is conservative (one pixel of slack); mtn_plan applies the exact
 * predicate on the receiving rank.  world <= 16, n_fields <= 12.
 */
size_t mtn_route_scratch_bytes(int64_t n, int32_t world);
int mtn_route_count(int64_t n, const double* px, const double* sm_range, int32_t world,
                    const int32_t* bounds_host, void* scratch, size_t scratch_bytes,
                    int64_t* totals_out, void* stream);
int mtn_route_scatter(int64_t n, const double* px, const double* sm_range, int32_t world,
                      const int32_t* bounds_host, int32_t n_fields, const double* const* fields_host,
                      double* const* inboxes_host, int64_t capacity, const int64_t* src_offsets,
                      const void* scratch, void* stream);

/*
 * Measurement hooks (bench.py).  With timing enabled, mtn_project records CUDA events on
 * the caller's stream at its stage boundaries; mtn_last_timing then returns the seven stage
 * durations in ms: [emit, sort, items, project kernel, partial reduce, finalize, second stream
 * (column / splat kernel + its reduce)].
 * With exec counting enabled, mtn_project launches a diagnostic instantiation of the
 * projection kernel that also tallies the algorithmic work actually executed
 * (mtn_last_exec_counts: [particle-channel updates with non-zero weight and spectrum,
 * kernel integrals evaluated, unsaturated edge erfs]) and synchronises; never time that one.
 */
int mtn_set_timing(int enable);
int mtn_last_timing(float* ms_out, int n);
int mtn_set_count_exec(int enable);
int mtn_last_exec_counts(int64_t* out3);

/*
 * Beam convolution, the step after the projection.  Replaces the
 * per-channel scipy.signal.fftconvolve(slice, beam.kernel, mode="same") loop of
 * Martini.convolve_beam (martini/martini.py:863-901):
 *   cube_out[x, y, c] = scale * sum_{a,b} cube_in[x + ka/2 - a, y + kb/2 - b, c] * kernel[a, b]
 * with zeros outside the (nx, ny, nc) cube; kernel is the (ka, kb) beam image (odd sizes,
 * device memory, row-major); scale carries the Jy/arcsec^2 -> Jy/beam factor (beam area).
 * Out of place.
 */
int mtn_convolve_beam(const double* cube_in, double* cube_out, int32_t nx, int32_t ny, int32_t nc,
                      const double* kernel, int32_t ka, int32_t kb, double scale, void* stream);

/*
 * FP64 FMA throughput microbenchmark (register-resident dependent-chain FMAs), used by
 * bench.py as the measured roofline denominator for the projection kernel.  Returns the
 * achieved TFLOP/s (2 flops per FMA) in *tflops_out (host).  Synchronises.
 */
int mtn_fp64_peak(double* tflops_out, double* ms_out, void* stream);

/*
 * Device-function probes used by the parity tests: evaluate the kernel integral
 * (w_out[i] = W(dx[i], dy[i], h[i])) or the per-particle channel spectrum
 * (s_out[i*C + c]) exactly as the projection kernel does.  closed_form != 0 evaluates the
 * reference's closed-form expression instead of the tabulated one (Wendland C2 and the cubic
 * spline are tabulated, csrc/tables.cuh).  mtn_table_error returns the worst deviation of a
 * kind's table from its closed form, relative to W(0), found when the table was built (host
 * memory; 0 for kinds without a table).
 */
int mtn_table_error(int32_t kind, double* err_out);
int mtn_probe_kernel_integral(const MtnKernelEntry* entry, int32_t closed_form, int64_t n,
                              const double* dx, const double* dy, const double* h, double* w_out,
                              void* stream);
int mtn_probe_spectra(int32_t spectrum, int64_t n, const double* v, const double* sigma,
                      double sigma_scalar, const double* amp, int32_t n_channels,
                      const double* edges, double* s_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MARTINI_B200_H */
