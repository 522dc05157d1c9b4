/*
 * sphb200.h -- C ABI of the B200-native SPH particle -> grid scatter library
 * (libsphb200.so, built for sm_100a from sarracen_b200/csrc).
 *
 * This is the drop-in boundary for Sarracen's one data-parallel hot path.  Each
 * entry point replaces one compiled routine behind the reference's backend
 * interface `BaseBackend` (sarracen/interpolate/base_backend.py:7-177); the
 * reference-side binding (a ctypes stub inside a `BaseBackend` subclass) is
 * shown in INTEGRATION.md and implemented in sarracen_b200/backend.py.
 *
 * Conventions
 *  - Every array pointer is a DEVICE pointer owned by the caller (PyTorch
 *    tensors in practice).  The library never allocates or frees device
 *    memory: scratch space is a caller-provided workspace.  Calls that need a
 *    data-dependent amount report it through `ws_needed` and return
 *    SPHB_ERR_WORKSPACE when `ws_bytes` is too small; the caller grows the
 *    buffer and calls again.
 *  - Particle columns are structure-of-arrays, all of one dtype (SPHB_F32 or
 *    SPHB_F64).  SPHB_F64 computes and accumulates in double; SPHB_F32
 *    computes in float with tile-local positions, accumulates float per tile
 *    chunk and flushes into the double output.  Output images are always
 *    float64, row-major [y][x] / [z][y][x], like the reference's
 *    (cpu_backend.py:243).
 *  - `stream` is a cudaStream_t passed as void*.  Work is stream-ordered;
 *    the scatter calls synchronise the stream once (they read the bin totals
 *    back to size the pair list and the cell-ordered record stream), the
 *    others are fully asynchronous.
 *  - Return value 0 on success, negative SPHB_ERR_* otherwise;
 *    sphb_last_error() returns a thread-local message.
 *  - No global mutable state; re-entrant across streams and devices.
 */
#ifndef SPHB200_H
#define SPHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHB_VERSION 200
#define SPHB_MAX_WEIGHTS 4
#define SPHB_MAX_STAGES 64

enum { SPHB_F32 = 0, SPHB_F64 = 1 };

/* Weight functions.  The first three are the analytic kernels of
 * sarracen/kernels/{cubic,quartic,quintic}_spline.py; SPHB_COLUMN_LUT is the
 * tabulated column-integrated kernel evaluated by linear interpolation
 * (sarracen/kernels/base_kernel.py:68-99). */
enum { SPHB_CUBIC = 0, SPHB_QUARTIC = 1, SPHB_QUINTIC = 2, SPHB_COLUMN_LUT = 3 };

enum {
    SPHB_OK = 0,
    SPHB_ERR_ARGUMENT = -1,  /* invalid argument */
    SPHB_ERR_WORKSPACE = -2, /* workspace too small; *ws_needed holds the size */
    SPHB_ERR_CUDA = -3,      /* CUDA runtime error; see sphb_last_error() */
    SPHB_ERR_TOO_MANY = -4   /* more than 2^31-1 (particle, tile) pairs: split the particle set */
};

typedef void *sphb_stream;

/* Particle columns (device pointers).  `z` may be NULL for 2-D data.  `w[k]`
 * are the per-particle weights A m / rho (interpolate.py:448-472); one
 * footprint walk accumulates all `n_weights` of them (the `_vec` methods and
 * the normalisation pass of the reference are separate passes,
 * cpu_backend.py:58-63, interpolate.py:587-594). */
typedef struct sphb_particles {
    int64_t n;
    int dtype; /* SPHB_F32 | SPHB_F64, of x, y, z, h and w[] alike */
    const void *x, *y, *z, *h;
    int n_weights;
    const void *w[SPHB_MAX_WEIGHTS];
} sphb_particles;

/* Weight function.  `ndim` is the exponent of h in the particle term
 * w / h^ndim and the normalisation dimension of the analytic kernels
 * (cpu_backend.py:251 and :34, :120, :168, :218).  For SPHB_COLUMN_LUT,
 * `lut` is a device array of `lut_samples` float64 and `ndim` must be 2. */
typedef struct sphb_kernel {
    int kind;
    int ndim;
    double radius;
    const double *lut;
    int lut_samples;
} sphb_kernel;

/* Output raster: pixel (i, j[, k]) has its centre at x_min + (i + 1/2) pwx. */
typedef struct sphb_raster {
    int nx, ny, nz; /* nz = 1 for images */
    double x_min, x_max, y_min, y_max, z_min, z_max;
} sphb_raster;

/* Optional counters filled by the scatter calls (host memory). */
typedef struct sphb_stats {
    int64_t n_pairs;      /* out: (particle, tile) pairs binned and sorted (tile path) */
    int64_t n_direct;     /* out: small-footprint particles rendered by the cell path */
    int64_t n_tiles;      /* out: tiles in the raster */
    int32_t tile_x, tile_y, tile_z; /* out: tile size in pixels */
    int32_t time_phases;  /* in: non-zero -> bracket the phases with CUDA events on `stream`
                             and synchronise at the end of the call to read them */
    float ms_bin;         /* out: count + cell scan + emit kernels */
    float ms_sort;        /* out: radix sort of the pairs + segment boundaries */
    float ms_render;      /* out: the tile kernel (large footprints) alone */
    float ms_direct;      /* out: the cell kernel (small footprints) alone */
    int32_t launches;     /* out: kernels of this library launched by the call (CUB's not counted) */
    float ms_count;       /* out: the count kernel + cell scan alone (part of ms_bin) */
} sphb_stats;

int sphb_version(void);
const char *sphb_last_error(void);

/* A4 -- column-integrated kernel table, sarracen/kernels/base_kernel.py:39-66
 * and :102-117.  out: device float64[samples]. */
int sphb_column_kernel(int kind, int samples, double *out, sphb_stream stream);

/* B3 -- sampled scatter onto an image, replaces CPUBackend._fast_2d
 * (cpu_backend.py:226-313) as used by interpolate_2d_render[_vec],
 * interpolate_3d_projection[_vec] (use_z = 0) and interpolate_3d_cross[_vec]
 * (use_z = 1: cull |z_slice - z| >= R h and add dz^2 to the distance).
 * out[k]: device float64[ny * nx] per weight.  If `accumulate` is zero the
 * images are cleared first, otherwise the result is added to them (for
 * particle sets processed in several batches). */
int sphb_scatter2d(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, int use_z,
                   double z_slice, int accumulate, double *const *out, void *ws, size_t ws_bytes,
                   size_t *ws_needed, sphb_stats *stats, sphb_stream stream);

/* B2/B3 -- sampled scatter onto a voxel grid in ONE pass over the particles,
 * replaces the per-slice loop of CPUBackend.interpolate_3d_grid
 * (cpu_backend.py:193-220).  out[k]: device float64[nz * ny * nx]. */
int sphb_scatter3d(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, int accumulate,
                   double *const *out, void *ws, size_t ws_bytes, size_t *ws_needed,
                   sphb_stats *stats, sphb_stream stream);

/* B6 -- line cut through 2-D data, replaces CPUBackend._fast_2d_line
 * (cpu_backend.py:450-538).  out[k]: device float64[pixels]. */
int sphb_line2d(const sphb_particles *p, const sphb_kernel *k, int pixels, double x1, double x2,
                double y1, double y2, int accumulate, double *const *out, sphb_stream stream);

/* B7 -- line cut through 3-D data, replaces CPUBackend._fast_3d_line
 * (cpu_backend.py:540-609). */
int sphb_line3d(const sphb_particles *p, const sphb_kernel *k, int pixels, double x1, double x2,
                double y1, double y2, double z1, double z2, int accumulate, double *const *out,
                sphb_stream stream);

/* Same, rendered in `n_stages` z-slab stages for the multi-GPU reduction along z
 * (SURVEY 8(e)): stage s finishes every z-plane below stage_planes[s] (ascending; the last
 * stage finishes the grid whatever its entry says), and right after the kernels of stage s have
 * been enqueued the library calls on_stage(user, s) on the calling thread -- the caller records an
 * event on `stream` there and starts reducing that slab on another stream while the later stages
 * render.  The kernels of consecutive stages run on two library-owned streams, forked from `stream`
 * and joined to it before each callback (so that the tail of one stage overlaps the head of the
 * next): work the caller enqueues on `stream` in or after the callback of stage s is ordered after
 * the stages 0..s and may overlap the later ones.  n_stages = 1 with on_stage = NULL is
 * sphb_scatter3d. */
typedef void (*sphb_stage_fn)(void *user, int stage);
int sphb_scatter3d_staged(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, int accumulate,
                          double *const *out, void *ws, size_t ws_bytes, size_t *ws_needed, sphb_stats *stats,
                          int n_stages, const int *stage_planes, sphb_stage_fn on_stage, void *user,
                          sphb_stream stream);

/* Plan / execute form of the sampled scatter: NO hidden synchronisation.
 *
 * sphb_scatter2d / 3d read four bin totals back to size their lists, which costs one stream
 * synchronisation per call.  A caller that wants a fully stream-asynchronous (CUDA-graph
 * capturable) call splits it:
 *   1. sphb_scatter_plan     count pass only; the totals are copied to `plan` (host memory,
 *                            page-locked for a truly asynchronous copy) with cudaMemcpyAsync.  The
 *                            caller synchronises whenever it likes -- or keeps the plan of an earlier
 *                            call on inputs of the same statistics.
 *   2. sphb_scatter_planned  the whole pipeline (count, emit, sort, render) with every list sized
 *                            from `plan`; it enqueues work and returns.  ws_needed reports the exact
 *                            workspace for that plan (call with ws = NULL to query).  If the particle
 *                            set holds MORE than the plan says, the excess is dropped and a flag is
 *                            raised in the workspace: sphb_scatter_overflowed() reads it (that call
 *                            synchronises).  With the plan of the same inputs the result is that of
 *                            sphb_scatter2d / 3d.
 * dim3 != 0: voxel grid (use_z ignored). */
typedef struct sphb_plan {
    unsigned long long n_small;        /* particles of the cell path */
    unsigned long long n_pairs;        /* (tile, particle) pairs of the tile path */
    unsigned long long max_small_edge; /* largest box edge among the cell-path particles */
    unsigned long long reserved0;
    unsigned long long n_large;        /* particles of the tile path */
    unsigned long long reserved[3];
} sphb_plan;

int sphb_scatter_plan(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, int dim3, int use_z,
                      double z_slice, void *ws, size_t ws_bytes, size_t *ws_needed, sphb_plan *plan,
                      sphb_stream stream);
int sphb_scatter_planned(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, int dim3, int use_z,
                         double z_slice, int accumulate, double *const *out, void *ws, size_t ws_bytes,
                         size_t *ws_needed, const sphb_plan *plan, sphb_stream stream);
int sphb_scatter_overflowed(const void *ws, int *flag, sphb_stream stream);

/* B4 -- exact pixel-area integral of the 2-D cubic spline, replaces
 * CPUBackend._exact_2d_render (cpu_backend.py:317-447) with the integrals of
 * sarracen/kernels/cubic_spline_exact.py:7-217.  Always computes in float64.
 * Workspace: 16 (n + 1) bytes + scan scratch + 16 bytes per lattice line of the particles' pixel
 * boxes (reported through ws_needed in up to two rounds); synchronises the stream once (the line
 * count sizes the table of per-line terms). */
int sphb_exact2d(const sphb_particles *p, const sphb_raster *r, int accumulate, double *const *out,
                 void *ws, size_t ws_bytes, size_t *ws_needed, sphb_stream stream);

/* B5 -- exact column integral of the 3-D cubic spline, replaces
 * CPUBackend._exact_3d_project (cpu_backend.py:611-720) with the integrals of
 * sarracen/kernels/cubic_spline_exact.py:220-483.  Always computes in float64.
 * Workspace: 8 (n + 1) bytes + scan scratch (size reported through ws_needed); fully asynchronous. */
int sphb_exact3d_proj(const sphb_particles *p, const sphb_raster *r, int accumulate,
                      double *const *out, void *ws, size_t ws_bytes, size_t *ws_needed, sphb_stream stream);

/* Measurement aid: the number of (particle, pixel) pairs for which
 * the reference evaluates the weight function and accumulates, i.e. pixel
 * centres with q <= R.  nz > 1 counts voxels.  out: device uint64 (one value,
 * overwritten). */
int sphb_count_contributions(const sphb_particles *p, double radius, const sphb_raster *r, int use_z,
                             double z_slice, unsigned long long *out, sphb_stream stream);

/* SURVEY 8(f)-4 -- radial binning behind Sarracen's disc analysis (sarracen/disc/utils.py:25-86:
 * pandas.cut of the particle radii).  index[i] = ring of particle i, i.e. the i such that
 * edges[i] < r <= edges[i + 1] (right-closed, like pandas.cut), -1 outside (edges[0], edges[bins]];
 * r = |(x, y) - origin| or, with `spherical`, |(x, y, z) - origin|; counts[ring] = its population.
 * `edges`: device float64[bins + 1], ascending; `origin`: HOST double[3]; index: device int32[n];
 * counts: device uint64[bins] (overwritten).  1 <= bins <= 6000. */
int sphb_radial_index(int dtype, const void *x, const void *y, const void *z, int64_t n, const double *origin,
                      int spherical, const double *edges, int bins, int *index, unsigned long long *counts,
                      sphb_stream stream);

/* Sums of `n_values` (1..4) value columns per ring (the groupby(...).sum() of
 * sarracen/disc/surface_density.py:148-151, :188-201): out[k][ring] = sum of values[k][i] over
 * the particles with index[i] == ring.  values: HOST array of device pointers of `dtype`;
 * out: HOST array of device float64[bins] (overwritten). */
int sphb_bin_sums(int dtype, const int *index, int n_values, const void *const *values, int64_t n, int bins,
                  double *const *out, sphb_stream stream);

/* N5 -- out = nan_to_num(num / den) elementwise on the device
 * (interpolate.py:594): 0/0 -> 0, +-inf -> +-DBL_MAX. */
int sphb_normalize(double *num, const double *den, int64_t n, sphb_stream stream);

/* Receive side of the multi-GPU exchange along z (no reference counterpart: the reference is one process;
 * SURVEY 8(e)).  dst[i] = own[i] + sum over slots s != skip of slots[s * slot_stride + i], i < n.  `slots`
 * are this rank's receive buffers -- peer-mapped device memory the other ranks' copy engines have written
 * their partial z-planes into -- `own` this rank's partial planes.  16-byte accesses when the pointers are 16-byte aligned
 * and slot_stride is even, scalar otherwise.  Asynchronous. */
int sphb_sum_slots(double *dst, const double *own, const double *slots, int n_slots, int64_t slot_stride,
                   int skip, int64_t n, sphb_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* SPHB200_H */
