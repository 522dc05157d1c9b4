// Sampled SPH scatter onto pixel / voxel rasters for B200 (sm_100a).
//
// Replaces CPUBackend._fast_2d (sarracen/interpolate/cpu_backend.py:226-313)
// and the per-slice loop of CPUBackend.interpolate_3d_grid (:193-220):
//
//   out[j,i] = sum_p (w_p / h_p^d) W( |r_ij - r_p| / h_p )
//
// The bounding box the reference builds per particle is only a cull (every W
// is exactly zero at and beyond its radius), so the result is the plain sum
// over all pixel centres inside the support -- computed here in a different
// order and a different decomposition:
//
//  1. bin    every particle's support box is classified: boxes of at most
//            small_edge pixels per axis are listed once (direct path), larger
//            ones are keyed by every raster tile they cover (count ->
//            exclusive scan -> emit (tile, particle) pairs); the emit pass also
//            writes one packed record per particle,
//  2. sort   cub::DeviceRadixSort on the tile key (only the bits in use),
//  3. render the sorted pair list is cut into fixed chunks, one CTA per
//            chunk.  A CTA walks the tile segments of its chunk; per tile it
//            stages the particles' derived records into shared memory, and
//            every warp owns an exclusive strip (2-D: rows, 3-D: z-planes) of
//            the tile, so accumulation needs no atomics: wide footprints go to
//            per-lane register accumulators on fixed pixels, narrow ones to a
//            plain load-add-store on the shared-memory tile (render_kernel),
//  4. flush  a finished tile segment is added to the float64 output with
//            red.global.add.f64 -- the only global atomics of the tile path,
//            once per tile segment, coalesced along x,
//  5. direct the listed small-footprint particles are rendered straight into
//            the raster, a lane per in-plane pixel, one red.global.add.f64 per
//            contribution (direct_kernel).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdlib>
#include <mutex>
#include <set>
#include <utility>

#include "sphb_common.cuh"

namespace sphb {

// CTA shape of the tile render kernel: WARPS warps, one staged particle per thread per round.
constexpr int kChunkPairs = 2048; // sorted pairs per CTA

// Geometry of one call, shared by every kernel of the pipeline.
struct View {
    int nx, ny, nz;
    int tx, ty, tz;    // tile size in pixels (powers of two)
    int tx_log, ty_log, tz_log;
    int ntx, nty, ntz; // tiles per axis
    double x_min, y_min, z_min;
    double inv_pwx, inv_pwy, inv_pwz;
    double pwx, pwy, pwz;
    double radius;
    double z_slice;
    int use_z; // 2-D raster, particles carry a depth (cross-section)
    int dim3;  // voxel raster
    int ndim;  // exponent of h in the particle term
    double norm;
    int rec_stride;    // elements per packed particle record (4 or 8)
    int small_shift;   // direct-path sort key = tile id >> small_shift (coarse buckets when nothing is tiled)
    unsigned long long n_out;   // raster elements per weight (debug index checks)
    int small_edge;    // footprint boxes with every edge <= this many pixels take the direct path
    uint32_t flag_bit; // key bit that marks a direct-path particle (sorts after all tile pairs)
};

// Packed particle record, written once per call by the emit kernel (which streams the
// columns anyway) so that the render kernels fetch one particle with one or two 16-byte
// loads from a single place instead of one 8-byte gather per column -- every random gather
// costs a full DRAM burst, and the sorted lists visit particles in random order.
//   stride 4 elements: { x, y, h, w0 }                    (no depth column, one weight)
//   stride 6 elements: { x, y, h, z, w0, w1 }             (double only, up to two weights)
//   stride 8 elements: { x, y, h, z, w0, w1, w2, w3 }
template <typename T> struct Rec {
    T x, y, h, z, w[4];
};

template <typename T>
__device__ __forceinline__ Rec<T> load_rec(const T *__restrict__ recs, int stride, int64_t p, int nw)
{
    Rec<T> r;
    const T *b = recs + (size_t)p * stride;
    if (sizeof(T) == 8) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(b));
        const double2 c = __ldg(reinterpret_cast<const double2 *>(b) + 1);
        r.x = (T)a.x; r.y = (T)a.y; r.h = (T)c.x;
        if (stride == 4) {
            r.z = T(0); r.w[0] = (T)c.y;
        } else {
            r.z = (T)c.y;
            const double2 d = __ldg(reinterpret_cast<const double2 *>(b) + 2);
            r.w[0] = (T)d.x; r.w[1] = (T)d.y;
            if (stride == 8 && nw > 2) {
                const double2 e = __ldg(reinterpret_cast<const double2 *>(b) + 3);
                r.w[2] = (T)e.x; r.w[3] = (T)e.y;
            }
        }
    } else {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(b));
        r.x = (T)a.x; r.y = (T)a.y; r.h = (T)a.z;
        if (stride == 4) {
            r.z = T(0); r.w[0] = (T)a.w;
        } else {
            r.z = (T)a.w;
            const float4 d = __ldg(reinterpret_cast<const float4 *>(b) + 1);
            r.w[0] = (T)d.x; r.w[1] = (T)d.y; r.w[2] = (T)d.z; r.w[3] = (T)d.w;
        }
    }
    return r;
}

struct TileBox {
    int x0, x1, y0, y1, z0, z1; // tile index ranges, inclusive
};

__device__ __forceinline__ int clamp_to_int(double v, int lo, int hi)
{
    // also maps NaN to lo
    return (int)fmin(fmax(v, (double)lo), (double)hi);
}

// Pixel box [lo, hi) that holds every pixel centre within rad of c.
__device__ __forceinline__ void pixel_range(double c, double rad, double origin, double inv_pw, int n,
                                            int &lo, int &hi)
{
    // centre of pixel i sits at origin + (i + 1/2) pw
    lo = clamp_to_int(ceil((c - rad - origin) * inv_pw - 0.5), 0, n);
    hi = clamp_to_int(floor((c + rad - origin) * inv_pw - 0.5) + 1.0, 0, n);
}

// Tiles covered by a particle's support box; false if it touches no pixel.
// `small` is set when the whole box is at most small_edge pixels along every
// axis: such a particle is not binned per tile but listed once (under the tile
// of its box origin) for the direct path.
template <typename T>
__device__ __forceinline__ bool particle_tiles(const View &v, const T *x, const T *y, const T *z,
                                               const T *h, int64_t i, TileBox &b, bool &small)
{
    double hp = (double)__ldg(h + i);
    double xp = (double)__ldg(x + i), yp = (double)__ldg(y + i);
    if (!(hp > 0.0) || !isfinite(hp) || !isfinite(xp) || !isfinite(yp)) return false;
    double rad = v.radius * hp;
    int lo, hi, edge = 0;
    b.z0 = b.z1 = 0;
    if (v.dim3) {
        double zp = (double)__ldg(z + i);
        if (!isfinite(zp)) return false;
        pixel_range(zp, rad, v.z_min, v.inv_pwz, v.nz, lo, hi);
        if (hi <= lo) return false;
        b.z0 = lo >> v.tz_log;
        b.z1 = (hi - 1) >> v.tz_log;
        edge = hi - lo;
    } else if (v.use_z) {
        double dz = v.z_slice - (double)__ldg(z + i);
        if (!(fabs(dz) < rad)) return false; // cpu_backend.py:264
    }
    pixel_range(xp, rad, v.x_min, v.inv_pwx, v.nx, lo, hi);
    if (hi <= lo) return false;
    b.x0 = lo >> v.tx_log;
    b.x1 = (hi - 1) >> v.tx_log;
    edge = max(edge, hi - lo);
    pixel_range(yp, rad, v.y_min, v.inv_pwy, v.ny, lo, hi);
    if (hi <= lo) return false;
    b.y0 = lo >> v.ty_log;
    b.y1 = (hi - 1) >> v.ty_log;
    edge = max(edge, hi - lo);
    small = edge <= v.small_edge;
    return true;
}

template <typename T>
__global__ void __launch_bounds__(256)
count_tiles_kernel(View v, const T *x, const T *y, const T *z, const T *h, int64_t n, uint32_t *counts,
                   unsigned long long *totals)
{
    // Grid-stride over the particles; the two totals -- direct-path particles (totals[0]) and list
    // entries (totals[1], 64-bit: the 32-bit per-particle counts are scanned in 32 bits, which is
    // only valid once the host has seen that the total is below 2^31) -- reach global memory as
    // one atomic each per CTA of a few-hundred-CTA grid.
    unsigned mine_small = 0;
    unsigned long long mine_all = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        TileBox b;
        uint32_t c = 0; // counts[n] = 0 so that the exclusive scan leaves the total there
        bool small = false;
        if (i < n && particle_tiles(v, x, y, z, h, i, b, small))
            c = small ? 1u
                      : (uint32_t)(b.x1 - b.x0 + 1) * (uint32_t)(b.y1 - b.y0 + 1) * (uint32_t)(b.z1 - b.z0 + 1);
        counts[i] = c; // fits: the raster has fewer than 2^30 tiles
        mine_small += (small && c) ? 1u : 0u;
        mine_all += c;
    }
    __shared__ unsigned long long block_small, block_all;
    if (threadIdx.x == 0) block_small = block_all = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mine_small += __shfl_xor_sync(0xffffffffu, mine_small, o);
        mine_all += __shfl_xor_sync(0xffffffffu, mine_all, o);
    }
    if ((threadIdx.x & 31) == 0 && mine_all) {
        atomicAdd(&block_small, (unsigned long long)mine_small);
        atomicAdd(&block_all, mine_all);
    }
    __syncthreads();
    if (threadIdx.x == 0 && block_all) {
        if (block_small) atomicAdd(totals, block_small);
        atomicAdd(totals + 1, block_all);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
emit_pairs_kernel(View v, const T *x, const T *y, const T *z, const T *h, const T *w0, const T *w1,
                  const T *w2, const T *w3, int nw, int64_t n, const uint32_t *offsets, uint32_t *keys,
                  uint32_t *vals, T *recs)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TileBox b;
    bool small;
    if (!particle_tiles(v, x, y, z, h, i, b, small)) return;
    {
        // packed record, written as 16-byte vectors (a scalar store per field touched every
        // 32-byte sector of the record array 4-8 times)
        T f[8] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
        f[0] = x[i];
        f[1] = y[i];
        f[2] = h[i];
        if (v.rec_stride == 4) {
            f[3] = w0[i];
        } else {
            f[3] = z ? z[i] : T(0);
            f[4] = w0[i];
            f[5] = nw > 1 ? w1[i] : T(0);
            f[6] = (v.rec_stride == 8 && nw > 2) ? w2[i] : T(0);
            f[7] = (v.rec_stride == 8 && nw > 3) ? w3[i] : T(0);
        }
        T *r = recs + (size_t)i * v.rec_stride;
        if (sizeof(T) == 8) {
            double2 *q = reinterpret_cast<double2 *>(r);
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (2 * k < v.rec_stride) q[k] = make_double2((double)f[2 * k], (double)f[2 * k + 1]);
        } else {
            float4 *q = reinterpret_cast<float4 *>(r);
#pragma unroll
            for (int k = 0; k < 2; k++)
                if (4 * k < v.rec_stride)
                    q[k] = make_float4((float)f[4 * k], (float)f[4 * k + 1], (float)f[4 * k + 2], (float)f[4 * k + 3]);
        }
    }
    uint64_t o = offsets[i];
    SPHB_ASSERT(o < offsets[n]);
    if (small) {
        keys[o] = ((uint32_t)((b.z0 * v.nty + b.y0) * v.ntx + b.x0) >> v.small_shift) | v.flag_bit;
        vals[o] = (uint32_t)i;
        return;
    }
    for (int tz = b.z0; tz <= b.z1; tz++)
        for (int ty = b.y0; ty <= b.y1; ty++)
            for (int tx = b.x0; tx <= b.x1; tx++) {
                SPHB_ASSERT(o < offsets[n] && tx < v.ntx && ty < v.nty && tz < v.ntz);
                keys[o] = (uint32_t)((tz * v.nty + ty) * v.ntx + tx);
                vals[o] = (uint32_t)i;
                o++;
            }
}

__global__ void tile_end_kernel(const uint32_t *keys, int64_t n_pairs, uint32_t *tile_end)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    if (i == n_pairs - 1 || keys[i] != keys[i + 1]) tile_end[keys[i]] = (uint32_t)(i + 1);
}

// Column tables of up to kLutSharedMax samples are staged in shared memory by every CTA; larger
// ones (integral_samples is a user argument, base_kernel.py:68) are expanded once per call into
// segment lines (A_i, B_i) (sphb_common.cuh, lut_index) in the workspace and read through L1/L2.
constexpr int kLutSharedMax = 2048;

template <typename R>
__global__ void lut_pairs_kernel(const double *lut, int samples, double scale, typename Real2<R>::type *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= samples) return;
    out[i] = lut_make_entry<R>(lut, samples, i, scale);
}

// ---------------------------------------------------------------------------
// Render.

// Staged particle record, tile-local.  Coordinates are in pixel units with the
// tile origin subtracted; in float mode the centre is split into the nearest
// pixel (ic) and a fraction in [-1/2, 1/2] so that position differences keep
// full float precision however large the raster is.
template <typename R, int NW, int BATCH, bool DIM3> struct Staged {
    using R2 = typename Real2<R>::type;
    // read one per lane when a warp scans the batch for its strip: the box along the split axis
    // (rows in 2-D, z-planes in 3-D), lo | hi << 16 -- kept as its own array so that the scan is
    // conflict-free
    uint32_t key[BATCH];
    // the rest is read by whole warps (broadcast) when a particle hits: packed so that a hit costs
    // 4 (2-D) or 6 (3-D) 16-byte loads instead of 10-14 scalar ones
    int4 ibox[BATCH];            // bx, by (lo | hi << 16), cx, cy (nearest pixel in float mode, else 0)
    int4 ibox3[DIM3 ? BATCH : 1]; // bz, cz, -, -   (3-D only)
    R2 fxy[BATCH];               // centre (double mode) / fraction (float mode), tile-local pixel units
    R2 sxy[BATCH];               // (pixel width / h)^2
    R2 fsz[DIM3 ? BATCH : 1];    // fz, sz         (3-D only)
    R2 ct[(NW + 2) / 2][BATCH];  // c (dz^2/h^2 of a cross-section), term[0] | term[1], term[2] | term[3], -
};

template <typename R> struct Split;
template <> struct Split<double> {
    static __device__ __forceinline__ void apply(double u, double &f, int &ic)
    {
        f = u;
        ic = 0;
    }
};
template <> struct Split<float> {
    static __device__ __forceinline__ void apply(double u, float &f, int &ic)
    {
        double r = rint(u);
        r = fmin(fmax(r, -2097152.0), 2097152.0);
        ic = (int)r;
        f = (float)(u - r);
    }
};

// Tile render.  Work split inside a CTA of kWarps warps:
//   2-D tile TX x TY pixels: warp w owns the strip of rows [w TY/8, (w+1) TY/8);
//   3-D tile TX x TY x TZ voxels with TZ == kWarps: warp w owns z-plane w.
// Ownership is exclusive, so accumulation needs no atomics.  Per (warp,
// particle) hit there are two code paths:
//   wide   (footprint box wider than TX/2 inside the tile): lanes sit on fixed
//          pixels of the warp's box -- column lane % TX, rows lane / TX + k RI --
//          and accumulate in REGISTERS (K values per lane and weight);
//   narrow lanes are laid over the footprint box as LW x (32 / LW) pixels, LW the
//          power of two >= box width, and accumulate into the warp's part of the
//          shared-memory tile (plain load / fma / store).
// Registers and shared memory are merged when the tile segment is flushed.
// Particles staged per round: one per thread.
template <typename R, bool DIM3, int WARPS> struct RenderBatch {
    static constexpr int value = WARPS * 32;
};

// Row padding of the shared-memory tile.  2-D (32 pixels per row): 4 elements, so that consecutive
// rows start 8 banks apart in double -- as good for the narrow path's few-rows-by-few-columns
// accesses as the former 8, and what lets two CTAs of the two-weight double kernel (2 x 36.9 KB of
// tile each) share an SM.
template <bool DIM3> struct RowPad {
    static constexpr int value = DIM3 ? 8 : 4;
};

template <typename T, typename R, int KIND, int NW, bool DIM3, int TX, int TY, int TZ, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, (DIM3 || NW > 2 || (NW > 1 && sizeof(R) == 8)) ? 2 : 3)
render_kernel(View v, const T *__restrict__ recs, const double *__restrict__ lut, int lut_samples,
              const void *__restrict__ lut_global, const uint32_t *__restrict__ keys,
              const uint32_t *__restrict__ vals, const uint32_t *__restrict__ tile_end, int64_t n_pairs,
              double *o0, double *o1, double *o2, double *o3)
{
    using R2 = typename Real2<R>::type;
    constexpr int kWarps = WARPS, kThreads = WARPS * 32, kBatch = RenderBatch<R, DIM3, WARPS>::value;
    constexpr int SX = TX + RowPad<DIM3>::value;  // padded row stride
    constexpr int PLANE = SX * TY;                // elements per z-plane
    constexpr int ACC = PLANE * TZ;               // elements per weight
    constexpr int ROWS = DIM3 ? TY : TY / kWarps; // rows of the warp's box
    constexpr int RI = 32 / TX;                   // rows one warp iteration covers in the wide path
    constexpr int K = ROWS / RI;                  // register accumulators per lane and weight
    static_assert(!DIM3 || TZ == kWarps, "one z-plane per warp");
    static_assert(DIM3 || TY % kWarps == 0, "strip split");
    static_assert(32 % TX == 0 && ROWS % RI == 0, "lane layout");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    R *acc = reinterpret_cast<R *>(smem_raw);
    Staged<R, NW, kBatch, DIM3> *st = reinterpret_cast<Staged<R, NW, kBatch, DIM3> *>(acc + (size_t)ACC * NW);
    R2 *lut_s = reinterpret_cast<R2 *>(st + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *outs[4] = {o0, o1, o2, o3};

    // Column table in shared memory as segment lines (A_i, B_i), see lut_index().  (Replicating it so that every
    // lane of an access phase owns a bank group was measured on config 2: 1, 2, 4, 8 copies ->
    // 124, 127, 145, 216 ms, every doubling costs a resident CTA; the random index is not the
    // bottleneck.)
    LutView<R> lv;
    lut_set_table(lv, lut_s);
    lv.gtab = static_cast<const R2 *>(lut_global);
    lv.scale = R(0);
    lv.last = 0;
    if (is_lut<KIND>()) {
        const double lut_scale = (double)(lut_samples - 1) / v.radius;
        for (int i = tid; i < (KIND == SPHB_COLUMN_LUT_GLOBAL ? 0 : lut_samples); i += kThreads)
            lut_s[i] = lut_make_entry<R>(lut, lut_samples, i, lut_scale);
        lv.scale = (R)lut_scale;
        lv.last = lut_samples - 1;
    }
    for (int i = tid; i < ACC * NW; i += kThreads) acc[i] = R(0);
    __syncthreads();

    const R rad2 = (R)(v.radius * v.radius);
    // the warp's box inside the tile
    const int own_y0 = DIM3 ? 0 : warp * ROWS;
    const int own_lo = DIM3 ? warp : own_y0, own_hi = DIM3 ? warp + 1 : own_y0 + ROWS;
    R *const my_acc = acc + (DIM3 ? warp * PLANE : 0);
    // wide-path lane position
    const int lx = lane % TX, ly = lane / TX;

    R racc[NW][K];
#pragma unroll
    for (int k = 0; k < NW; k++)
#pragma unroll
        for (int r = 0; r < K; r++) racc[k][r] = R(0);

    const int64_t chunk_lo = (int64_t)blockIdx.x * kChunkPairs;
    const int64_t chunk_hi = min(chunk_lo + (int64_t)kChunkPairs, n_pairs);
    int64_t pos = chunk_lo;

    while (pos < chunk_hi) {
        const uint32_t tile = keys[pos];
        SPHB_ASSERT(tile < (uint32_t)(v.ntx * v.nty * v.ntz) && (int64_t)tile_end[tile] > pos);
        const int64_t seg_hi = min((int64_t)tile_end[tile], chunk_hi);
        const int ttx = tile % v.ntx, tty = (tile / v.ntx) % v.nty, ttz = tile / (v.ntx * v.nty);
        const int ox = ttx * TX, oy = tty * TY, oz = ttz * TZ; // tile origin in pixels
        const int lim_x = min(TX, v.nx - ox), lim_y = min(TY, v.ny - oy), lim_z = min(TZ, v.nz - oz);

        for (int64_t base = pos; base < seg_hi; base += kBatch) {
            const int nb = (int)min((int64_t)kBatch, seg_hi - base);
            // ---- stage: one particle per thread --------------------------------
            if (tid < nb) {
                const int64_t p = vals[base + tid];
                const Rec<T> pr = load_rec<T>(recs, v.rec_stride, p, NW);
                const double hp = (double)pr.h;
                const double xp = (double)pr.x, yp = (double)pr.y;
                const double rad = v.radius * hp;
                const double hinv = 1.0 / hp, hinv2 = hinv * hinv;
                int lo, hi;
                // pixel-unit coordinate in which pixel i has its centre at i
                const double ux = (xp - v.x_min) * v.inv_pwx - 0.5 - ox;
                const double uy = (yp - v.y_min) * v.inv_pwy - 0.5 - oy;
                pixel_range(xp, rad, v.x_min, v.inv_pwx, v.nx, lo, hi);
                lo = max(lo - ox, 0);
                hi = min(hi - ox, lim_x);
                const uint32_t bx = (uint32_t)lo | ((uint32_t)max(hi, lo) << 16);
                pixel_range(yp, rad, v.y_min, v.inv_pwy, v.ny, lo, hi);
                lo = max(lo - oy, 0);
                hi = min(hi - oy, lim_y);
                const uint32_t by = (uint32_t)lo | ((uint32_t)max(hi, lo) << 16);
                R2 f, sc;
                int cx, cy;
                Split<R>::apply(ux, f.x, cx);
                Split<R>::apply(uy, f.y, cy);
                sc.x = (R)(v.pwx * hinv); // pixel -> q along x and y
                sc.y = (R)(v.pwy * hinv);
                st->ibox[tid] = make_int4((int)bx, (int)by, cx, cy);
                st->fxy[tid] = f;
                st->sxy[tid] = sc;
                uint32_t key = by;
                double cc = 0.0;
                if (DIM3) {
                    const double zp = (double)pr.z;
                    const double uz = (zp - v.z_min) * v.inv_pwz - 0.5 - oz;
                    pixel_range(zp, rad, v.z_min, v.inv_pwz, v.nz, lo, hi);
                    lo = max(lo - oz, 0);
                    hi = min(hi - oz, lim_z);
                    key = (uint32_t)lo | ((uint32_t)max(hi, lo) << 16);
                    R2 fz;
                    int cz;
                    Split<R>::apply(uz, fz.x, cz);
                    fz.y = (R)(v.pwz * v.pwz * hinv2);
                    st->fsz[tid] = fz;
                    st->ibox3[tid] = make_int4((int)key, cz, 0, 0);
                } else if (v.use_z) {
                    const double dz = v.z_slice - (double)pr.z;
                    cc = dz * dz * hinv2;
                }
                SPHB_ASSERT(tid < kBatch && p >= 0);
                st->key[tid] = key;
                const double scale = v.norm * (v.ndim == 2 ? hinv2 : hinv2 * hinv);
                R tv[4] = {R(0), R(0), R(0), R(0)};
#pragma unroll
                for (int k = 0; k < NW; k++) tv[k] = (R)((double)pr.w[k] * scale);
                R2 e0;
                e0.x = (R)cc + sqrt_guard(R(0)); // every q^2 built on it is > 0 (sqrt_pos)
                e0.y = tv[0];
                st->ct[0][tid] = e0;
                if (NW > 1) {
                    R2 e1;
                    e1.x = tv[1];
                    e1.y = tv[2];
                    st->ct[1][tid] = e1;
                }
                if (NW > 3) {
                    R2 e2;
                    e2.x = tv[3];
                    e2.y = R(0);
                    st->ct[2][tid] = e2;
                }
            }
            __syncthreads();

            // ---- accumulate: each warp scans the batch for its box --------------
            for (int g = 0; g < nb; g += 32) {
                const int me = g + lane;
                bool hit = false;
                if (me < nb) {
                    const uint32_t b = st->key[me];
                    const int lo = (int)(b & 0xffff), hi = (int)(b >> 16);
                    hit = max(lo, own_lo) < min(hi, own_hi);
                }
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int s = g + __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int4 ib = st->ibox[s];
                    const uint32_t bx = (uint32_t)ib.x, by = (uint32_t)ib.y;
                    const int x0 = (int)(bx & 0xffff), x1 = (int)(bx >> 16);
                    const int y0 = max((int)(by & 0xffff), own_y0), y1 = min((int)(by >> 16), own_y0 + ROWS);
                    const int wx = x1 - x0;
                    if (wx <= 0 || y1 <= y0) continue;
                    const R2 f = st->fxy[s], sc = st->sxy[s], e0 = st->ct[0][s];
                    const R fx = f.x, fy = f.y, sx = sc.x, sy = sc.y;
                    // double mode stages whole coordinates (Split<double>: nearest pixel 0): keeping that
                    // a compile-time fact makes the lane's own pixel coordinate loop-invariant
                    const int cx = sizeof(R) == 8 ? 0 : ib.z, cy = sizeof(R) == 8 ? 0 : ib.w;
                    R cc = e0.x;
                    if (DIM3) {
                        const R2 fz = st->fsz[s];
                        const R dz = int_to_real(warp - st->ibox3[s].y, R(0)) - fz.x;
                        cc = fma(dz * dz, fz.y, sqrt_guard(R(0)));
                    }
                    R term[NW];
                    term[0] = e0.y;
                    if (NW > 1) {
                        const R2 e1 = st->ct[1][s];
                        term[1] = e1.x;
                        if (NW > 2) term[2] = e1.y;
                    }
                    if (NW > 3) term[3] = st->ct[2][s].x;

                    // wide when the box spans more than half the tile width (a quarter in float mode,
                    // where idle lanes are cheaper than the narrow path's shared-memory updates:
                    // 65.5 -> 63.0 ms on config 2; no gain in double)
                    if (wx > TX / (sizeof(R) == 4 ? 4 : 2)) {
                        // ---- wide: fixed pixels, register accumulators -------------
                        // Every lane evaluates its own pixels: a pixel outside the
                        // footprint box has q >= R and W(q) = 0 there (the table is
                        // clamped to its zero last entry, the analytic shapes are
                        // selected to 0), so no box tests are needed.  Rows are handled G at
                        // a time without branches so the G dependent square-root /
                        // table chains overlap; groups that cannot contribute are skipped.
                        const R dx = (int_to_real(lx - cx, R(0)) - fx) * sx;
                        const R qx = fma(dx, dx, cc); // cc carries sqrt_guard()
                        const R dy0 = (int_to_real(own_y0 + ly - cy, R(0)) - fy) * sy;
                        constexpr int G = K % 4 == 0 ? 4 : (K % 2 == 0 ? 2 : 1);
#pragma unroll
                        for (int r0 = 0; r0 < K; r0 += G) {
                            R q[G];
                            bool any = false;
#pragma unroll
                            for (int j = 0; j < G; j++) {
                                // q^2 is evaluated directly per row.  Second differences down the
                                // strip (one add less) carry an absolute error of ulp(q^2_max of the
                                // strip): too much for float against the 1e-5 budget, and in double
                                // it shows as an error of sqrt(1e-14) in q right at q = 0, which a
                                // coarse column table (integral_samples = 2: a cusp at 0) turns into
                                // 4e-9 -- found by tests/test_gpu_fuzz.py.
                                const R dy = (r0 + j) == 0 ? dy0 : fma(R((r0 + j) * RI), sy, dy0);
                                q[j] = fma(dy, dy, qx);
                                if (sizeof(R) == 4) any = any || (q[j] < rad2);
                            }
                            // Skip a group that cannot contribute.  double: rows that miss the
                            // footprint box (warp-uniform integer test; a DSETP per row costs more
                            // than the rare all-outside group).  float: exact warp vote on q^2
                            // (comparisons are cheap there, skipped corners are not).  Measured on
                            // config 2: 124 -> 117 ms (double), 78 vs 83 ms (float).
                            bool live;
                            if (sizeof(R) == 4)
                                live = __any_sync(0xffffffffu, any);
                            else
                                live = own_y0 + r0 * RI < y1 && own_y0 + (r0 + G) * RI > y0;
                            if (live) {
#pragma unroll
                                for (int j = 0; j < G; j++) {
                                    const R wv = kernel_eval_unguarded<KIND, R>(q[j], rad2, lv);
#pragma unroll
                                    for (int k = 0; k < NW; k++)
                                        racc[k][r0 + j] = fma(term[k], wv, racc[k][r0 + j]);
                                }
                            }
                        }
                    } else {
                        // ---- narrow: lanes over the footprint box, shared memory ----
                        const int lw_log = 32 - __clz(wx - 1); // smallest power of two >= wx (wx >= 1)
                        const int rows_per_it = 32 >> lw_log;
                        const int px = x0 + (lane & ((1 << lw_log) - 1));
                        const int pr = lane >> lw_log;
                        const bool in_x = px < x1;
                        const R dx = (int_to_real(px - cx, R(0)) - fx) * sx;
                        const R qx = fma(dx, dx, cc); // cc carries sqrt_guard()
                        R dy = int_to_real(y0 + pr - cy, R(0)) - fy;
                        const R dstep = int_to_real(rows_per_it, R(0));
                        R *a = my_acc + ((y0 + pr) * SX + px);
                        for (int py = y0 + pr; py < y1; py += rows_per_it) {
                            if (in_x) {
                                const R dys = dy * sy;
                                const R q2 = fma(dys, dys, qx);
                                if (q2 < rad2) {
                                    const R wv = kernel_eval_pos<KIND, R>(q2, lv);
                                    SPHB_ASSERT(px >= 0 && px < TX && py >= own_y0 && py < own_y0 + ROWS &&
                                                a >= acc && a + (NW - 1) * ACC < acc + (size_t)ACC * NW);
#pragma unroll
                                    for (int k = 0; k < NW; k++) a[k * ACC] = fma(term[k], wv, a[k * ACC]);
                                }
                            }
                            dy += dstep;
                            a += rows_per_it * SX;
                        }
                        __syncwarp();
                    }
                }
            }
            __syncthreads();
        }

        // ---- merge registers into the tile, flush it, clear both ---------------
#pragma unroll
        for (int r = 0; r < K; r++) {
            R *a = my_acc + ((own_y0 + ly + r * RI) * SX + lx);
#pragma unroll
            for (int k = 0; k < NW; k++) {
                // lanes beyond the raster edge hold values of pixels that do not exist
                if (racc[k][r] != R(0) && lx < lim_x && own_y0 + ly + r * RI < lim_y) a[k * ACC] += racc[k][r];
                racc[k][r] = R(0);
            }
        }
        __syncthreads();
        for (int i = tid; i < TX * TY * TZ; i += kThreads) {
            const int ix = i % TX, iy = (i / TX) % TY, iz = i / (TX * TY);
            if (ix < lim_x && iy < lim_y && iz < lim_z) {
                const int a = iz * PLANE + iy * SX + ix;
                const size_t o = ((size_t)(oz + iz) * v.ny + (oy + iy)) * v.nx + (ox + ix);
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    const R val = acc[k * ACC + a];
                    if (val != R(0)) {
                        SPHB_ASSERT(o < v.n_out && a < ACC);
                        red_add(outs[k] + o, (double)val);
                        acc[k * ACC + a] = R(0);
                    }
                }
            }
        }
        __syncthreads();
        pos = seg_hi;
    }
}

// ---------------------------------------------------------------------------
// Direct path for small footprints.
//
// A particle whose support box is at most small_edge pixels along every axis
// covers a handful of pixels; binning it per tile, staging it per tile and
// scanning it per warp costs far more than its few kernel evaluations.  Such
// particles are listed once (sorted by the tile of their box origin, so that
// consecutive particles touch the same few kilobytes of the raster and the
// reductions are served by L2) and rendered here: one warp takes 32 of them,
// every lane works out the record of one, then sub-groups of LPP lanes -- LPP
// the power of two >= the largest in-plane box area in the batch -- walk one
// particle each, a lane per in-plane pixel and a loop over z, adding each
// contribution to the float64 raster with red.global.add.f64.
constexpr int kDirectWarps = 8;

template <typename R, int NW> struct DirectRec {
    R fx[32], fy[32], fz[32]; // centre in pixel units relative to the box origin
    R sx[32], sy[32], sz[32];
    R cq[32];                 // dz^2 / h^2 of a cross-section + sqrt guard
    R term[NW][32];
    unsigned long long base[32]; // raster offset of the box origin
    uint32_t inv_wx[32];         // 65536 / wx + 1
    uint32_t dims[32];           // wx | wy << 8 | wz << 16 ; 0 = nothing to do
};

template <typename T, typename R, int KIND, int NW, bool DIM3>
__global__ void __launch_bounds__(kDirectWarps * 32)
direct_kernel(View v, const T *__restrict__ recs, const double *__restrict__ lut, int lut_samples,
              const void *__restrict__ lut_global, const uint32_t *__restrict__ order, int64_t n_small, double *o0, double *o1, double *o2,
              double *o3)
{
    using R2 = typename Real2<R>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DirectRec<R, NW> *drecs = reinterpret_cast<DirectRec<R, NW> *>(smem_raw);
    R2 *lut_s = reinterpret_cast<R2 *>(drecs + kDirectWarps);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DirectRec<R, NW> &rec = drecs[warp];
    double *outs[4] = {o0, o1, o2, o3};

    LutView<R> lv;
    lut_set_table(lv, lut_s);
    lv.gtab = static_cast<const R2 *>(lut_global);
    lv.scale = R(0);
    lv.last = 0;
    if (is_lut<KIND>()) {
        const double lut_scale = (double)(lut_samples - 1) / v.radius;
        for (int i = tid; i < (KIND == SPHB_COLUMN_LUT_GLOBAL ? 0 : lut_samples); i += kDirectWarps * 32)
            lut_s[i] = lut_make_entry<R>(lut, lut_samples, i, lut_scale);
        lv.scale = (R)lut_scale;
        lv.last = lut_samples - 1;
        __syncthreads();
    }
    const R rad2 = (R)(v.radius * v.radius);
    const int64_t warps_total = (int64_t)gridDim.x * kDirectWarps;

    for (int64_t base = ((int64_t)blockIdx.x * kDirectWarps + warp) * 32; base < n_small;
         base += warps_total * 32) {
        // ---- every lane prepares one particle --------------------------------
        int area = 0;
        {
            uint32_t dims = 0;
            if (base + lane < n_small) {
                const int64_t p = order[base + lane];
                const Rec<T> pr = load_rec<T>(recs, v.rec_stride, p, NW);
                const double hp = (double)pr.h;
                const double xp = (double)pr.x, yp = (double)pr.y;
                const double rad = v.radius * hp;
                const double hinv = 1.0 / hp, hinv2 = hinv * hinv;
                int x0, x1, y0, y1, z0 = 0, z1 = 1;
                pixel_range(xp, rad, v.x_min, v.inv_pwx, v.nx, x0, x1);
                pixel_range(yp, rad, v.y_min, v.inv_pwy, v.ny, y0, y1);
                rec.fx[lane] = (R)((xp - v.x_min) * v.inv_pwx - 0.5 - x0);
                rec.fy[lane] = (R)((yp - v.y_min) * v.inv_pwy - 0.5 - y0);
                rec.sx[lane] = (R)(v.pwx * v.pwx * hinv2);
                rec.sy[lane] = (R)(v.pwy * v.pwy * hinv2);
                double cc = 0.0;
                if (DIM3) {
                    const double zp = (double)pr.z;
                    pixel_range(zp, rad, v.z_min, v.inv_pwz, v.nz, z0, z1);
                    rec.fz[lane] = (R)((zp - v.z_min) * v.inv_pwz - 0.5 - z0);
                    rec.sz[lane] = (R)(v.pwz * v.pwz * hinv2);
                } else if (v.use_z) {
                    const double dz = v.z_slice - (double)pr.z;
                    cc = dz * dz * hinv2;
                }
                rec.cq[lane] = (R)cc + sqrt_guard(R(0));
                const double scale = v.norm * (v.ndim == 2 ? hinv2 : hinv2 * hinv);
#pragma unroll
                for (int k = 0; k < NW; k++) rec.term[k][lane] = (R)((double)pr.w[k] * scale);
                rec.base[lane] = ((unsigned long long)z0 * v.ny + (unsigned long long)y0) * v.nx + x0;
                const int wx = x1 - x0, wy = y1 - y0, wz = z1 - z0;
                rec.inv_wx[lane] = 65536u / (unsigned)max(wx, 1) + 1u;
                if (wx > 0 && wy > 0 && wz > 0) {
                    dims = (uint32_t)wx | ((uint32_t)wy << 8) | ((uint32_t)wz << 16);
                    area = wx * wy;
                }
            }
            rec.dims[lane] = dims;
        }
        // largest in-plane area of the batch -> lanes per particle
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) area = max(area, __shfl_xor_sync(0xffffffffu, area, o));
        __syncwarp();
        if (area == 0) continue;
        int lpp_log = 0;
        while ((1 << lpp_log) < area && lpp_log < 5) lpp_log++;
        const int lpp = 1 << lpp_log, sub = lane & (lpp - 1), grp = lane >> lpp_log;

        for (int g = 0; g < 32; g += 32 >> lpp_log) {
            const int s = g + grp;
            const uint32_t dims = rec.dims[s];
            if (dims) {
                const int wx = (int)(dims & 0xff), wy = (int)((dims >> 8) & 0xff), wz = (int)(dims >> 16);
                const R fx = rec.fx[s], fy = rec.fy[s], sx = rec.sx[s], sy = rec.sy[s], cq = rec.cq[s];
                R fz = R(0), sz = R(0);
                if (DIM3) {
                    fz = rec.fz[s];
                    sz = rec.sz[s];
                }
                R term[NW];
#pragma unroll
                for (int k = 0; k < NW; k++) term[k] = rec.term[k][s];
                const unsigned inv_wx = rec.inv_wx[s];
                const unsigned long long base_o = rec.base[s];
                for (int e = sub; e < wx * wy; e += lpp) {
                    const int iy = (int)(((unsigned)e * inv_wx) >> 16), ix = e - iy * wx;
                    const R dx = int_to_real(ix, R(0)) - fx, dy = int_to_real(iy, R(0)) - fy;
                    const R qxy = fma(dy * dy, sy, fma(dx * dx, sx, cq));
                    const size_t o = (size_t)base_o + (size_t)(iy * v.nx + ix);
                    if (DIM3) {
                        size_t oo = o;
                        for (int iz = 0; iz < wz; iz++) {
                            const R dz = int_to_real(iz, R(0)) - fz; // direct evaluation (see render_kernel)
                            const R q2 = fma(dz * dz, sz, qxy);
                            if (q2 < rad2) {
                                const R wv = kernel_eval_pos<KIND, R>(q2, lv);
                                SPHB_ASSERT(oo < v.n_out);
#pragma unroll
                                for (int k = 0; k < NW; k++) red_add(outs[k] + oo, (double)(term[k] * wv));
                            }
                            oo += (size_t)v.ny * v.nx;
                        }
                    } else if (qxy < rad2) {
                        const R wv = kernel_eval_pos<KIND, R>(qxy, lv);
                        SPHB_ASSERT(o < v.n_out);
#pragma unroll
                        for (int k = 0; k < NW; k++) red_add(outs[k] + o, (double)(term[k] * wv));
                    }
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Host side.

// Opt a kernel in to the full 227 KB of dynamic shared memory, once per instantiation and
// device (driver calls per launch are kept to a minimum: they serialise with other
// driver clients such as NVML monitors).
template <typename K> static cudaError_t allow_large_smem(K kern)
{
    // keyed by (device, kernel address): instantiations share their function-pointer type
    static std::mutex mu;
    static std::set<std::pair<int, const void *>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const std::pair<int, const void *> key(dev, reinterpret_cast<const void *>(kern));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (done.count(key)) return cudaSuccess;
    }
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lock(mu);
        done.insert(key);
    }
    return e;
}

// SM count of the current device (cached per device).
static cudaError_t sm_count(int *out)
{
    static std::mutex mu;
    static int cached[64] = {0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 0 || dev >= 64 || cached[dev] == 0) {
        int val = 0;
        e = cudaDeviceGetAttribute(&val, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64) {
            *out = val;
            return cudaSuccess;
        }
        cached[dev] = val;
    }
    *out = cached[dev];
    return cudaSuccess;
}

// Timing events are created once per host thread and device and reused.
static cudaError_t phase_events(cudaEvent_t *ev, int count)
{
    static thread_local cudaEvent_t cache[64][8] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || count > 8) return cudaErrorInvalidValue;
    for (int i = 0; i < count; i++) {
        if (!cache[dev][i]) {
            e = cudaEventCreate(&cache[dev][i]);
            if (e != cudaSuccess) return e;
        }
        ev[i] = cache[dev][i];
    }
    return cudaSuccess;
}

struct Workspace {
    unsigned char *base;
    size_t used;
    explicit Workspace(void *p) : base(static_cast<unsigned char *>(p)), used(0) {}
    template <typename U> U *take(size_t count)
    {
        used = align_up(used, 256);
        U *r = reinterpret_cast<U *>(base + used);
        used += count * sizeof(U);
        return r;
    }
};

template <typename T, typename R, int KIND, int NW, bool DIM3, int TX, int TY, int TZ, int WARPS>
static int launch_render(const View &v, const T *recs, const sphb_kernel *k, const void *lut_global,
                         const uint32_t *keys, const uint32_t *vals, const uint32_t *tile_end, int64_t n_pairs,
                         double *const *out, cudaStream_t stream)
{
    using R2 = typename Real2<R>::type;
    auto kern = render_kernel<T, R, KIND, NW, DIM3, TX, TY, TZ, WARPS>;
    const size_t fixed = sizeof(R) * (size_t)(TX + RowPad<DIM3>::value) * TY * TZ * NW + sizeof(Staged<R, NW, RenderBatch<R, DIM3, WARPS>::value, DIM3>);
    const size_t limit = 227 * 1024;
    size_t smem = fixed;
    if (KIND == SPHB_COLUMN_LUT) smem += sizeof(R2) * (size_t)k->lut_samples;
    if (smem > limit) {
        set_error("column table of %d samples does not fit in shared memory", k->lut_samples);
        return SPHB_ERR_ARGUMENT;
    }
    SPHB_CUDA(allow_large_smem(kern));
    const int grid = div_up(n_pairs, kChunkPairs);
    double *o[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < NW; i++) o[i] = out[i];
    kern<<<grid, WARPS * 32, smem, stream>>>(v, recs, k->lut, k->lut_samples, lut_global, keys, vals,
                                             tile_end, n_pairs, o[0], o[1], o[2], o[3]);
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}

template <typename T, typename R, int KIND, int NW, bool DIM3>
static int launch_direct(const View &v, const T *recs, const sphb_kernel *k, const void *lut_global,
                         const uint32_t *order, int64_t n_small, double *const *out, cudaStream_t stream)
{
    using R2 = typename Real2<R>::type;
    auto kern = direct_kernel<T, R, KIND, NW, DIM3>;
    size_t smem = sizeof(DirectRec<R, NW>) * kDirectWarps;
    if (KIND == SPHB_COLUMN_LUT) smem += sizeof(R2) * (size_t)k->lut_samples;
    if (smem > 227 * 1024) {
        set_error("column table of %d samples does not fit in shared memory", k->lut_samples);
        return SPHB_ERR_ARGUMENT;
    }
    SPHB_CUDA(allow_large_smem(kern));
    int grid = div_up(n_small, kDirectWarps * 32);
    grid = grid > 148 * 64 ? 148 * 64 : grid;
    double *o[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < NW; i++) o[i] = out[i];
    kern<<<grid, kDirectWarps * 32, smem, stream>>>(v, recs, k->lut, k->lut_samples, lut_global, order, n_small,
                                                    o[0], o[1], o[2], o[3]);
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}

template <typename T, typename R, int KIND, bool DIM3, typename... A>
static int direct_nw(int nw, A &&...a)
{
    switch (nw) {
    case 1: return launch_direct<T, R, KIND, 1, DIM3>(a...);
    case 2: return launch_direct<T, R, KIND, 2, DIM3>(a...);
    case 3: return launch_direct<T, R, KIND, 3, DIM3>(a...);
    default: return launch_direct<T, R, KIND, 4, DIM3>(a...);
    }
}

template <typename T, typename R, bool DIM3, typename... A> static int direct_kind(int kind, int nw, A &&...a)
{
    switch (kind) {
    case SPHB_CUBIC: return direct_nw<T, R, SPHB_CUBIC, DIM3>(nw, a...);
    case SPHB_QUARTIC: return direct_nw<T, R, SPHB_QUARTIC, DIM3>(nw, a...);
    case SPHB_QUINTIC: return direct_nw<T, R, SPHB_QUINTIC, DIM3>(nw, a...);
    case SPHB_COLUMN_LUT: return direct_nw<T, R, SPHB_COLUMN_LUT, DIM3>(nw, a...);
    default: return direct_nw<T, R, SPHB_COLUMN_LUT_GLOBAL, DIM3>(nw, a...);
    }
}

template <typename T, typename R, int KIND, bool DIM3, int TX, int TY, int TZ, int WARPS, typename... A>
static int dispatch_nw(int nw, A &&...a)
{
    switch (nw) {
    case 1: return launch_render<T, R, KIND, 1, DIM3, TX, TY, TZ, WARPS>(a...);
    case 2: return launch_render<T, R, KIND, 2, DIM3, TX, TY, TZ, WARPS>(a...);
    // three and four weights in double: half-height 2-D tiles (see tile2y())
    case 3: return launch_render<T, R, KIND, 3, DIM3, TX, (!DIM3 && sizeof(R) == 8) ? TY / 2 : TY, TZ, WARPS>(a...);
    default: return launch_render<T, R, KIND, 4, DIM3, TX, (!DIM3 && sizeof(R) == 8) ? TY / 2 : TY, TZ, WARPS>(a...);
    }
}

template <typename T, typename R, bool DIM3, int TX, int TY, int TZ, int WARPS, typename... A>
static int dispatch_kind(int kind, int nw, A &&...a)
{
    switch (kind) {
    case SPHB_CUBIC: return dispatch_nw<T, R, SPHB_CUBIC, DIM3, TX, TY, TZ, WARPS>(nw, a...);
    case SPHB_QUARTIC: return dispatch_nw<T, R, SPHB_QUARTIC, DIM3, TX, TY, TZ, WARPS>(nw, a...);
    case SPHB_QUINTIC: return dispatch_nw<T, R, SPHB_QUINTIC, DIM3, TX, TY, TZ, WARPS>(nw, a...);
    case SPHB_COLUMN_LUT: return dispatch_nw<T, R, SPHB_COLUMN_LUT, DIM3, TX, TY, TZ, WARPS>(nw, a...);
    default: return dispatch_nw<T, R, SPHB_COLUMN_LUT_GLOBAL, DIM3, TX, TY, TZ, WARPS>(nw, a...);
    }
}

// 2-D tiles: 32 x 128 pixels, 8 warps with a 16-row strip each (16 register accumulators per lane
// and weight): per-hit set-up is amortised over twice the rows of a 32 x 64 tile and the pair list
// shrinks by 28 %; measured 104 -> 99.5 ms (double), 68.9 -> 63.6 ms (float) on config 2 at
// 3 CTAs/SM (80 registers).  16 warps x 8 rows on the same tile was 20 % slower.
constexpr int kTile2X = 32, kTile2Y = 128;

// Footprint boxes up to this many pixels per edge take the direct path.  The
// default can be overridden with SPHB_SMALL_EDGE (0 sends everything through
// the tile path) for experiments.
static int small_edge_setting()
{
    const char *e = getenv("SPHB_SMALL_EDGE");
    if (e && *e) {
        int val = atoi(e);
        return val < 0 ? 0 : val > 15 ? 15 : val;
    }
    return 8;
}
constexpr int kTile3X = 16, kTile3Y = 16, kTile3Z = 8;

// Height of the 2-D tile.  With three or four weight columns in double a 32 x 128 tile needs
// 111 / 148 KB of accumulators and 96 / 128 accumulator registers per lane: one CTA per SM, with
// spills.  Half the height brings back two CTAs per SM and spill-free code.
static int tile2y(bool f32, int n_weights) { return (!f32 && n_weights >= 3) ? kTile2Y / 2 : kTile2Y; }

static int validate(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, bool dim3,
                    int use_z, double *const *out)
{
    if (!p || !k || !r || !out) {
        set_error("null argument");
        return SPHB_ERR_ARGUMENT;
    }
    if (p->n < 0 || p->n > 0xffffffffLL) {
        set_error("particle count %lld out of range", (long long)p->n);
        return SPHB_ERR_ARGUMENT;
    }
    if (p->dtype != SPHB_F32 && p->dtype != SPHB_F64) {
        set_error("dtype must be SPHB_F32 or SPHB_F64");
        return SPHB_ERR_ARGUMENT;
    }
    if (p->n_weights < 1 || p->n_weights > SPHB_MAX_WEIGHTS) {
        set_error("n_weights must be 1..%d", SPHB_MAX_WEIGHTS);
        return SPHB_ERR_ARGUMENT;
    }
    if (p->n > 0 && (!p->x || !p->y || !p->h || ((dim3 || use_z) && !p->z))) {
        set_error("missing particle column");
        return SPHB_ERR_ARGUMENT;
    }
    for (int i = 0; i < p->n_weights; i++)
        if (!out[i] || (p->n > 0 && !p->w[i])) {
            set_error("missing weight or output %d", i);
            return SPHB_ERR_ARGUMENT;
        }
    if (r->nx <= 0 || r->ny <= 0 || (dim3 && r->nz <= 0) || !(r->x_max > r->x_min) ||
        !(r->y_max > r->y_min) || (dim3 && !(r->z_max > r->z_min))) {
        set_error("invalid raster");
        return SPHB_ERR_ARGUMENT;
    }
    if (k->kind < SPHB_CUBIC || k->kind > SPHB_COLUMN_LUT || !(k->radius > 0.0)) {
        set_error("invalid kernel");
        return SPHB_ERR_ARGUMENT;
    }
    if (k->kind == SPHB_COLUMN_LUT && (!k->lut || k->lut_samples < 2)) {
        set_error("column table needs >= 2 samples");
        return SPHB_ERR_ARGUMENT;
    }
    if (k->ndim != 2 && k->ndim != 3) {
        set_error("kernel ndim must be 2 or 3");
        return SPHB_ERR_ARGUMENT;
    }
    return SPHB_OK;
}

template <typename T>
static int scatter_impl(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, bool dim3,
                        int use_z, double z_slice, int accumulate, double *const *out, void *ws,
                        size_t ws_bytes, size_t *ws_needed, sphb_stats *stats, cudaStream_t stream)
{
    View v;
    v.nx = r->nx;
    v.ny = r->ny;
    v.nz = dim3 ? r->nz : 1;
    v.tx = dim3 ? kTile3X : kTile2X;
    v.ty = dim3 ? kTile3Y : tile2y(p->dtype == SPHB_F32, p->n_weights);
    v.tz = dim3 ? kTile3Z : 1;
    auto ilog2 = [](int t) { int l = 0; while ((1 << l) < t) l++; return l; };
    v.tx_log = ilog2(v.tx);
    v.ty_log = ilog2(v.ty);
    v.tz_log = ilog2(v.tz);
    v.ntx = div_up(v.nx, v.tx);
    v.nty = div_up(v.ny, v.ty);
    v.ntz = div_up(v.nz, v.tz);
    v.x_min = r->x_min;
    v.y_min = r->y_min;
    v.z_min = dim3 ? r->z_min : 0.0;
    v.pwx = (r->x_max - r->x_min) / r->nx;
    v.pwy = (r->y_max - r->y_min) / r->ny;
    v.pwz = dim3 ? (r->z_max - r->z_min) / r->nz : 1.0;
    v.inv_pwx = 1.0 / v.pwx;
    v.inv_pwy = 1.0 / v.pwy;
    v.inv_pwz = 1.0 / v.pwz;
    v.radius = k->radius;
    v.z_slice = z_slice;
    v.use_z = use_z;
    v.dim3 = dim3 ? 1 : 0;
    v.ndim = k->ndim;
    v.norm = kernel_norm(k->kind, k->ndim);
    v.small_edge = small_edge_setting();
    v.rec_stride = (!dim3 && !use_z && p->n_weights == 1) ? 4 : (sizeof(T) == 8 && p->n_weights <= 2) ? 6 : 8;
    const int64_t n_tiles = (int64_t)v.ntx * v.nty * v.ntz;
    if (n_tiles > 0x3fffffffLL) {
        set_error("raster has too many tiles");
        return SPHB_ERR_ARGUMENT;
    }
    int end_bit = 1;
    while ((1LL << end_bit) < n_tiles) end_bit++;
    v.flag_bit = 1u << end_bit; // direct-path entries sort behind every tile pair
    v.small_shift = 0;
    end_bit++;
    const int64_t n = p->n;
    const size_t n_out = (size_t)v.nx * v.ny * v.nz;
    v.n_out = n_out;

    if (!accumulate)
        for (int i = 0; i < p->n_weights; i++)
            SPHB_CUDA(cudaMemsetAsync(out[i], 0, n_out * sizeof(double), stream));
    const bool timed = stats && stats->time_phases;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (stats) {
        stats->n_pairs = 0;
        stats->n_tiles = n_tiles;
        stats->tile_x = v.tx;
        stats->tile_y = v.ty;
        stats->tile_z = v.tz;
        stats->ms_bin = stats->ms_sort = stats->ms_render = stats->ms_direct = 0.f;
        stats->n_direct = 0;
        stats->launches = 0;
    }
    if (timed) SPHB_CUDA(phase_events(ev, 5));
    if (ws_needed) *ws_needed = 0;
    if (n == 0) return SPHB_OK;

    // ---- phase A: count and scan ------------------------------------------
    size_t scan_tmp = 0;
    SPHB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                            n + 1, stream));
    Workspace a(ws);
    uint32_t *offsets = a.take<uint32_t>(n + 1); // counts, scanned in place
    unsigned long long *n_small_dev = a.take<unsigned long long>(2); // direct-path particles, list entries
    unsigned char *scan_ws = a.take<unsigned char>(scan_tmp);
    const size_t need_a = align_up(a.used, 256);
    if (ws_needed) *ws_needed = need_a;
    if (!ws || ws_bytes < need_a) {
        set_error("workspace of %zu bytes needed for binning, %zu given", need_a, ws_bytes);
        return SPHB_ERR_WORKSPACE;
    }
    const T *x = static_cast<const T *>(p->x), *y = static_cast<const T *>(p->y);
    const T *z = static_cast<const T *>(p->z), *h = static_cast<const T *>(p->h);
    const int blocks = div_up(n + 1, 256);
    if (timed) SPHB_CUDA(cudaEventRecord(ev[0], stream));
    SPHB_CUDA(cudaMemsetAsync(n_small_dev, 0, 2 * sizeof(unsigned long long), stream));
    {
        int sms = 0;
        SPHB_CUDA(sm_count(&sms));
        const int count_blocks = std::min<int64_t>(blocks, (int64_t)sms * 16);
        count_tiles_kernel<T><<<count_blocks, 256, 0, stream>>>(v, x, y, z, h, n, offsets, n_small_dev);
    }
    SPHB_LAUNCH_CHECK();
    SPHB_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan_tmp, offsets, offsets, n + 1, stream));
    unsigned long long totals_host[2] = {0, 0};
    SPHB_CUDA(cudaMemcpyAsync(totals_host, n_small_dev, sizeof(totals_host), cudaMemcpyDeviceToHost, stream));
    SPHB_CUDA(cudaStreamSynchronize(stream));
    const unsigned long long n_small_host = totals_host[0];
    const uint64_t total = totals_host[1]; // the 32-bit scan is meaningful only if this passes the check below
    if (stats) {
        stats->n_pairs = (int64_t)total - (int64_t)n_small_host;
        stats->n_direct = (int64_t)n_small_host;
    }
    if (total == 0) return SPHB_OK;
    if (total > 0x7fffffffULL) {
        set_error("%llu (particle, tile) pairs exceed 2^31-1: split the particle set",
                  (unsigned long long)total);
        if (ws_needed) *ws_needed = 0;
        return SPHB_ERR_TOO_MANY;
    }
    const int64_t n_pairs = (int64_t)total;                     // entries in the sorted list
    const int64_t n_small = (int64_t)n_small_host;              // its tail: direct-path particles
    const int64_t n_large = n_pairs - n_small;                  // its head: (tile, particle) pairs

    // ---- phase B: emit, sort, segment ------------------------------------
    // Key bits to sort on.  The flag is the top bit: with no direct-path entries it is never set.
    // With ONLY direct-path entries the order matters just for the locality of the reductions in
    // L2, so the flag is dropped and the keys are coarsened to buckets of tiles covering at most
    // 128 KiB of raster (SPHB_SMALL_BUCKET_KB overrides).  Measured on config 4 (1 GiB grid, 16-bit
    // tile ids + flag = three 8-bit radix passes): 16 / 128 KiB buckets = two passes, sort 2.64 ->
    // 1.9 ms, direct kernel unchanged; 1 MiB: hmin variant's direct kernel 5.2 -> 8.5 ms; 4 MiB (one
    // pass): 5.2 -> 14.4 ms and the float32 grid 28.7 -> 66 ms -- the reductions of neighbouring
    // particles must meet in the same L2 sectors.  A raster of <= 128 KiB is not sorted at all.
    v.small_shift = 0;
    if (n_small == 0) {
        end_bit--;
    } else if (n_large == 0) {
        const int tile_bits = end_bit - 1;
        const char *bk = getenv("SPHB_SMALL_BUCKET_KB");
        const double bucket_bytes = 1024.0 * (bk && *bk ? atof(bk) : 128.0);
        const double buckets = (double)n_out * 8.0 * p->n_weights / bucket_bytes;
        int want = 0;
        while ((double)(1 << want) < buckets && want < tile_bits) want++;
        v.small_shift = tile_bits - want;
        v.flag_bit = 0;
        end_bit = want;
    }
    const bool sorted = end_bit > 0;
    size_t sort_tmp = 0;
    SPHB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                              (uint32_t *)nullptr, (uint32_t *)nullptr, n_pairs, 0,
                                              sorted ? end_bit : 1, stream));
    uint32_t *keys_a = a.take<uint32_t>(n_pairs);
    uint32_t *vals_a = a.take<uint32_t>(n_pairs);
    uint32_t *keys_b = a.take<uint32_t>(n_pairs);
    uint32_t *vals_b = a.take<uint32_t>(n_pairs);
    uint32_t *tile_end = a.take<uint32_t>(n_tiles);
    T *recs = a.take<T>((size_t)n * v.rec_stride);
    const bool big_lut = k->kind == SPHB_COLUMN_LUT && k->lut_samples > kLutSharedMax;
    void *lut_global = big_lut ? static_cast<void *>(a.take<double2>(k->lut_samples)) : nullptr;
    unsigned char *sort_ws = a.take<unsigned char>(sort_tmp);
    const size_t need_b = align_up(a.used, 256);
    if (ws_needed) *ws_needed = need_b;
    if (ws_bytes < need_b) {
        set_error("workspace of %zu bytes needed for %lld pairs, %zu given", need_b, (long long)n_pairs,
                  ws_bytes);
        return SPHB_ERR_WORKSPACE;
    }
    emit_pairs_kernel<T><<<blocks, 256, 0, stream>>>(
        v, x, y, z, h, static_cast<const T *>(p->w[0]), static_cast<const T *>(p->w[1]),
        static_cast<const T *>(p->w[2]), static_cast<const T *>(p->w[3]), p->n_weights, n, offsets, keys_a, vals_a,
        recs);
    SPHB_LAUNCH_CHECK();
    if (timed) SPHB_CUDA(cudaEventRecord(ev[1], stream));
    if (sorted)
        SPHB_CUDA(cub::DeviceRadixSort::SortPairs(sort_ws, sort_tmp, keys_a, keys_b, vals_a, vals_b, n_pairs, 0,
                                                  end_bit, stream));
    if (n_large > 0) {
        tile_end_kernel<<<div_up(n_large, 256), 256, 0, stream>>>(keys_b, n_large, tile_end);
        SPHB_LAUNCH_CHECK();
    }

    if (big_lut) {
        if (p->dtype == SPHB_F32)
            lut_pairs_kernel<float><<<div_up(k->lut_samples, 256), 256, 0, stream>>>(
                k->lut, k->lut_samples, (double)(k->lut_samples - 1) / v.radius, static_cast<float2 *>(lut_global));
        else
            lut_pairs_kernel<double><<<div_up(k->lut_samples, 256), 256, 0, stream>>>(
                k->lut, k->lut_samples, (double)(k->lut_samples - 1) / v.radius, static_cast<double2 *>(lut_global));
        SPHB_LAUNCH_CHECK();
    }

    // ---- phase C: render ---------------------------------------------------
    if (timed) SPHB_CUDA(cudaEventRecord(ev[2], stream));
    const bool f32 = p->dtype == SPHB_F32;
    const int kind = big_lut ? SPHB_COLUMN_LUT_GLOBAL : k->kind;
    int rc = SPHB_OK;
    if (n_large > 0) {
        if (dim3) {
            if (f32)
                rc = dispatch_kind<T, float, true, kTile3X, kTile3Y, kTile3Z, 8>(
                    kind, p->n_weights, v, recs, k, lut_global, keys_b, vals_b, tile_end, n_large, out, stream);
            else
                rc = dispatch_kind<T, double, true, kTile3X, kTile3Y, kTile3Z, 8>(
                    kind, p->n_weights, v, recs, k, lut_global, keys_b, vals_b, tile_end, n_large, out, stream);
        } else {
            if (f32)
                rc = dispatch_kind<T, float, false, kTile2X, kTile2Y, 1, 8>(kind, p->n_weights, v, recs, k, lut_global, keys_b,
                                                                           vals_b, tile_end, n_large, out, stream);
            else
                rc = dispatch_kind<T, double, false, kTile2X, kTile2Y, 1, 8>(kind, p->n_weights, v, recs, k, lut_global, keys_b,
                                                                            vals_b, tile_end, n_large, out, stream);
        }
        if (rc) return rc;
    }
    if (timed) SPHB_CUDA(cudaEventRecord(ev[4], stream));
    if (n_small > 0) {
        const uint32_t *order = (sorted ? vals_b : vals_a) + n_large;
        if (dim3)
            rc = f32 ? direct_kind<T, float, true>(kind, p->n_weights, v, recs, k, lut_global, order, n_small, out, stream)
                     : direct_kind<T, double, true>(kind, p->n_weights, v, recs, k, lut_global, order, n_small, out, stream);
        else
            rc = f32 ? direct_kind<T, float, false>(kind, p->n_weights, v, recs, k, lut_global, order, n_small, out, stream)
                     : direct_kind<T, double, false>(kind, p->n_weights, v, recs, k, lut_global, order, n_small, out, stream);
    }
    if (rc) return rc;
    if (stats) stats->launches = 2 + (n_large > 0 ? 2 : 0) + (n_small > 0 ? 1 : 0); // count, emit, tile_end, render, direct
    if (timed) {
        SPHB_CUDA(cudaEventRecord(ev[3], stream));
        SPHB_CUDA(cudaEventSynchronize(ev[3]));
        SPHB_CUDA(cudaEventElapsedTime(&stats->ms_bin, ev[0], ev[1]));
        SPHB_CUDA(cudaEventElapsedTime(&stats->ms_sort, ev[1], ev[2]));
        SPHB_CUDA(cudaEventElapsedTime(&stats->ms_render, ev[2], ev[4]));
        SPHB_CUDA(cudaEventElapsedTime(&stats->ms_direct, ev[4], ev[3]));
    }
    return SPHB_OK;
}

} // namespace sphb

using namespace sphb;

extern "C" int sphb_scatter2d(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, int use_z,
                              double z_slice, int accumulate, double *const *out, void *ws, size_t ws_bytes,
                              size_t *ws_needed, sphb_stats *stats, sphb_stream stream)
{
    int rc = validate(p, k, r, false, use_z, out);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (p->dtype == SPHB_F32)
        return scatter_impl<float>(p, k, r, false, use_z, z_slice, accumulate, out, ws, ws_bytes, ws_needed,
                                   stats, s);
    return scatter_impl<double>(p, k, r, false, use_z, z_slice, accumulate, out, ws, ws_bytes, ws_needed,
                                stats, s);
}

extern "C" int sphb_scatter3d(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r,
                              int accumulate, double *const *out, void *ws, size_t ws_bytes,
                              size_t *ws_needed, sphb_stats *stats, sphb_stream stream)
{
    int rc = validate(p, k, r, true, 0, out);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (p->dtype == SPHB_F32)
        return scatter_impl<float>(p, k, r, true, 0, 0.0, accumulate, out, ws, ws_bytes, ws_needed, stats, s);
    return scatter_impl<double>(p, k, r, true, 0, 0.0, accumulate, out, ws, ws_bytes, ws_needed, stats, s);
}
