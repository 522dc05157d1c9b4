// Sampled SPH scatter onto pixel / voxel rasters for B200 (sm_100a): binning and host side.
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
//  1. count  one streaming pass over x, y, z, h classifies every particle's support box.
//            Boxes of at most small_edge pixels per axis are SMALL: they are histogrammed by
//            the cell that holds their box origin.  Larger ones are counted by the raster tiles
//            they cover.  Nothing per particle is written.
//  2. emit   a second streaming pass writes one packed record per particle.  Small particles
//            are placed directly at their final, cell-ordered position (counting sort: scanned
//            cell histogram + one warp-aggregated atomic per cell and warp; the append
//            frontiers of all cells stay L2-resident, so DRAM sees full-sector writes).  Large
//            particles emit one (tile, particle) pair per covered tile into a list claimed
//            with one atomic per warp -- no per-particle offsets, no scan over the particles.
//  3. sort   cub::DeviceRadixSort on the tile key of the pairs only (on the bits in use).
//  4. render large footprints: tile kernel over the sorted pairs (sphb_render.cu);
//            small footprints: cell kernel over the cell-ordered records (sphb_cell.cu).
//            Both accumulate locally (registers / shared memory) and touch the raster with
//            red.global.add.f64 only when a tile segment / cell is finished.
//
// A voxel grid can be rendered in z-slab STAGES (sphb_scatter3d_staged): both work lists are
// z-major, so stage s finishes every plane below its boundary and a caller can start reducing
// that slab across GPUs while the later stages still render.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sphb_scatter.cuh"

namespace sphb {

struct Box {
    int x0, x1, y0, y1, z0, z1; // pixel ranges [lo, hi)
};

// Support box of a particle in raster pixels; false if it touches no pixel (or is degenerate:
// h <= 0 or non-finite input -- such particles contribute nothing).
__device__ __forceinline__ bool particle_box(const View &v, double xp, double yp, double zp, double hp, Box &b)
{
    if (!(hp > 0.0) || !isfinite(hp) || !isfinite(xp) || !isfinite(yp)) return false;
    const double rad = v.radius * hp;
    b.z0 = 0;
    b.z1 = 1;
    if (v.dim3) {
        if (!isfinite(zp)) return false;
        pixel_range(zp, rad, v.z_min, v.inv_pwz, v.nz, b.z0, b.z1);
        if (b.z1 <= b.z0) return false;
    } else if (v.use_z) {
        const double dz = v.z_slice - zp;
        if (!(fabs(dz) < rad)) return false; // cpu_backend.py:264
    }
    pixel_range(xp, rad, v.x_min, v.inv_pwx, v.nx, b.x0, b.x1);
    if (b.x1 <= b.x0) return false;
    pixel_range(yp, rad, v.y_min, v.inv_pwy, v.ny, b.y0, b.y1);
    return b.y1 > b.y0;
}

__device__ __forceinline__ int box_edge(const Box &b)
{
    return max(max(b.x1 - b.x0, b.y1 - b.y0), b.z1 - b.z0);
}

__device__ __forceinline__ uint32_t box_cell(const View &v, const Box &b)
{
    return (uint32_t)((((b.z0 >> v.cz_log) * v.ncy) + (b.y0 >> v.cy_log)) * v.ncx + (b.x0 >> v.cx_log));
}

struct TileBox {
    int x0, x1, y0, y1, z0, z1; // tile index ranges, inclusive
};

__device__ __forceinline__ TileBox box_tiles(const View &v, const Box &b)
{
    TileBox t;
    t.x0 = b.x0 >> v.tx_log;
    t.x1 = (b.x1 - 1) >> v.tx_log;
    t.y0 = b.y0 >> v.ty_log;
    t.y1 = (b.y1 - 1) >> v.ty_log;
    t.z0 = b.z0 >> v.tz_log;
    t.z1 = (b.z1 - 1) >> v.tz_log;
    return t;
}

__device__ __forceinline__ uint32_t tile_count(const TileBox &t)
{
    return (uint32_t)(t.x1 - t.x0 + 1) * (uint32_t)(t.y1 - t.y0 + 1) * (uint32_t)(t.z1 - t.z0 + 1);
}

// totals[0] small particles, [1] (tile, particle) pairs, [2] largest small box edge,
// [3] cursor of the pair list (emit), [4] large particles, [5] set when a planned call met more than its plan
enum { kTotSmall = 0, kTotPairs = 1, kTotEdge = 2, kTotCursor = 3, kTotLarge = 4, kTotOverflow = 5, kTotCount = 8 };

template <typename T>
__global__ void __launch_bounds__(256)
count_kernel(View v, const T *x, const T *y, const T *z, const T *h, int64_t n, uint32_t *cell_count,
             unsigned long long *totals)
{
    unsigned mine_small = 0, mine_large = 0, mine_edge = 0;
    unsigned long long mine_pairs = 0;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_up = (n + 31) / 32 * 32; // whole warps stay in the loop (match_any below)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_up; i += stride) {
        Box b;
        bool seen = false;
        if (i < n)
            seen = particle_box(v, (double)__ldg(x + i), (double)__ldg(y + i),
                                (v.dim3 || v.use_z) ? (double)__ldg(z + i) : 0.0, (double)__ldg(h + i), b);
        const int edge = seen ? box_edge(b) : 0;
        const bool small = seen && edge <= v.small_edge;
        // histogram of the small particles by cell, one atomic per distinct cell and warp
        const uint32_t key = small ? box_cell(v, b) : (0x80000000u | (uint32_t)lane);
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (small) {
            if (lane == __ffs(peers) - 1) atomicAdd(cell_count + key, (uint32_t)__popc(peers));
            mine_small++;
            mine_edge = max(mine_edge, (unsigned)edge);
        } else if (seen) {
            mine_pairs += tile_count(box_tiles(v, b));
            mine_large++;
        }
    }
    __shared__ unsigned long long block_tot[kTotCount];
    if (threadIdx.x < kTotCount) block_tot[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mine_small += __shfl_xor_sync(0xffffffffu, mine_small, o);
        mine_large += __shfl_xor_sync(0xffffffffu, mine_large, o);
        mine_pairs += __shfl_xor_sync(0xffffffffu, mine_pairs, o);
        mine_edge = max(mine_edge, __shfl_xor_sync(0xffffffffu, mine_edge, o));
    }
    if (lane == 0 && (mine_small | mine_large)) {
        atomicAdd(&block_tot[kTotSmall], (unsigned long long)mine_small);
        atomicAdd(&block_tot[kTotLarge], (unsigned long long)mine_large);
        atomicAdd(&block_tot[kTotPairs], mine_pairs);
        atomicMax(&block_tot[kTotEdge], (unsigned long long)mine_edge);
    }
    __syncthreads();
    if (threadIdx.x == 0 && (block_tot[kTotSmall] | block_tot[kTotLarge])) {
        if (block_tot[kTotSmall]) {
            atomicAdd(totals + kTotSmall, block_tot[kTotSmall]);
            atomicMax(totals + kTotEdge, block_tot[kTotEdge]);
        }
        if (block_tot[kTotLarge]) {
            atomicAdd(totals + kTotLarge, block_tot[kTotLarge]);
            atomicAdd(totals + kTotPairs, block_tot[kTotPairs]);
        }
    }
}

template <typename T>
__device__ __forceinline__ void write_rec(const View &v, T *r, T xv, T yv, T zv, T hv, const T *w0, const T *w1,
                                          const T *w2, const T *w3, int nw, int64_t i)
{
    // whole 32-byte sectors, one 256-bit store each (see Rec)
    T f[8] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
    f[0] = xv;
    f[1] = yv;
    f[2] = hv;
    if (v.rec_stride == 4) {
        f[3] = __ldg(w0 + i);
    } else {
        f[3] = zv;
        f[4] = __ldg(w0 + i);
        f[5] = nw > 1 ? __ldg(w1 + i) : T(0);
        f[6] = nw > 2 ? __ldg(w2 + i) : T(0);
        f[7] = nw > 3 ? __ldg(w3 + i) : T(0);
    }
    stg256(r, f);
    if (sizeof(T) == 8 && v.rec_stride == 8) stg256(r + 4, f + 4);
}

// Pairs of one particle are written by its own lane up to this many; a particle that covers
// more tiles (h >> tile: up to every tile of the raster) is written by the whole warp.
constexpr uint32_t kOwnPairs = 16;

// Particles per thread of the emit kernel: their column loads, and then their cell-cursor atomics,
// are issued back to back, so that several global round trips are in flight per thread (with
// one particle per thread the kernel sat at 15 % issue utilisation waiting for them).  Measured on the hmin
// variant of config 4 (per thread x CTAs per SM): 4 x 3 (80 registers) 6.15 ms, 4 x 4 (64, spills) 6.69,
// 2 x 4 (64, no spills) 5.96, 2 x 5 (48) 6.08, 8 x 2 (128) 7.12.
#ifndef SPHB_EMIT_PER
#define SPHB_EMIT_PER 2
#endif
#ifndef SPHB_EMIT_MINB
#define SPHB_EMIT_MINB 4
#endif
constexpr int kEmitPer = SPHB_EMIT_PER;

template <typename T>
__global__ void __launch_bounds__(256, SPHB_EMIT_MINB)
emit_kernel(View v, const T *x, const T *y, const T *z, const T *h, const T *w0, const T *w1, const T *w2,
            const T *w3, int nw, int64_t n, uint32_t *cell_cursor,
            unsigned long long *totals, uint32_t *keys, uint32_t *vals, T *recs_large, T *recs_small,
            int64_t n_pairs, int64_t n_small, int dbg_mode)
{
    const int lane = threadIdx.x & 31;
    // whole warps: the grid rounds n up; particle k of this thread is base + 256 k (coalesced columns)
    const int64_t base = (int64_t)blockIdx.x * (256 * kEmitPer) + threadIdx.x;
    const bool has_z = v.dim3 || v.use_z;
    T xs[kEmitPer], ys[kEmitPer], zs[kEmitPer], hs[kEmitPer];
#pragma unroll
    for (int k = 0; k < kEmitPer; k++) {
        const int64_t i = base + 256 * k;
        const bool in = i < n;
        xs[k] = in ? __ldg(x + i) : T(0);
        ys[k] = in ? __ldg(y + i) : T(0);
        zs[k] = (in && has_z) ? __ldg(z + i) : T(0);
        hs[k] = in ? __ldg(h + i) : T(0); // h = 0: not seen
    }
    Box b[kEmitPer];
    bool seen[kEmitPer], small[kEmitPer];
    uint32_t key[kEmitPer], slot[kEmitPer];
    unsigned peers[kEmitPer];
#pragma unroll
    for (int k = 0; k < kEmitPer; k++) {
        seen[k] = particle_box(v, (double)xs[k], (double)ys[k], (double)zs[k], (double)hs[k], b[k]);
        small[k] = seen[k] && box_edge(b[k]) <= v.small_edge;
        key[k] = small[k] ? box_cell(v, b[k]) : (0x80000000u | (uint32_t)lane);
    }
    // ---- small: final position in the cell-ordered record stream ------------------------
#pragma unroll
    for (int k = 0; k < kEmitPer; k++) {
        peers[k] = __match_any_sync(0xffffffffu, key[k]);
        slot[k] = 0;
        if (small[k]) {
            // (the cursors were initialised with the scanned cell starts: the atomic returns the position)
            if (lane == __ffs(peers[k]) - 1 && dbg_mode != 1)
                slot[k] = atomicAdd(cell_cursor + key[k], (uint32_t)__popc(peers[k]));
        }
    }
#pragma unroll
    for (int k = 0; k < kEmitPer; k++) {
        slot[k] = __shfl_sync(0xffffffffu, slot[k], __ffs(peers[k]) - 1);
        if (small[k]) {
            int64_t pos = (int64_t)slot[k] + __popc(peers[k] & ((1u << lane) - 1u));
            if (dbg_mode == 1) pos = (int64_t)(((uint64_t)(base + 256 * k) * 2654435761ull) % (uint64_t)n_small);
            if (dbg_mode == 3) pos = (base + 256 * k) % n_small;
            if (pos >= n_small) {   // only with a stale plan: drop and report
                totals[kTotOverflow] = 1;
                continue;
            }
            if (dbg_mode != 2)
            write_rec<T>(v, recs_small + (size_t)pos * v.rec_stride, xs[k], ys[k], zs[k], hs[k], w0, w1, w2, w3, nw,
                         base + 256 * k);
        }
    }

    // ---- large: (tile, particle) pairs ----------------------------------------------------
#pragma unroll
    for (int k = 0; k < kEmitPer; k++) {
        const int64_t i = base + 256 * k;
        const bool large = seen[k] && !small[k];
        if (!__any_sync(0xffffffffu, large)) continue;
        TileBox t = {0, 0, 0, 0, 0, 0};
        uint32_t c = 0;
        if (large) {
            t = box_tiles(v, b[k]);
            c = tile_count(t);
            write_rec<T>(v, recs_large + (size_t)i * v.rec_stride, xs[k], ys[k], zs[k], hs[k], w0, w1, w2, w3, nw, i);
        }
        // the warp claims one range of the list: exclusive scan of the counts + one atomic
        unsigned long long incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        unsigned long long first = 0;
        if (lane == 31) first = atomicAdd(totals + kTotCursor, incl);
        first = __shfl_sync(0xffffffffu, first, 31);
        const uint64_t o0 = first + incl - c;
        if (large && o0 + c > (uint64_t)n_pairs) {   // only with a stale plan: drop and report
            totals[kTotOverflow] = 1;
            c = 0;
        }
        if (large && c > 0 && c <= kOwnPairs) {
            uint64_t o = o0;
            for (int tz = t.z0; tz <= t.z1; tz++)
                for (int ty = t.y0; ty <= t.y1; ty++)
                    for (int tx = t.x0; tx <= t.x1; tx++) {
                        SPHB_ASSERT((int64_t)o < n_pairs && tx < v.ntx && ty < v.nty && tz < v.ntz);
                        keys[o] = (uint32_t)((tz * v.nty + ty) * v.ntx + tx);
                        vals[o] = (uint32_t)i;
                        o++;
                    }
        }
        unsigned wide = __ballot_sync(0xffffffffu, large && c > kOwnPairs);
        while (wide) {
            const int src = __ffs(wide) - 1;
            wide &= wide - 1;
            const uint64_t ob = __shfl_sync(0xffffffffu, o0, src);
            const uint32_t cc = __shfl_sync(0xffffffffu, c, src);
            const int tx0 = __shfl_sync(0xffffffffu, t.x0, src), ty0 = __shfl_sync(0xffffffffu, t.y0, src);
            const int tz0 = __shfl_sync(0xffffffffu, t.z0, src);
            const int nx_ = __shfl_sync(0xffffffffu, t.x1, src) - tx0 + 1;
            const int ny_ = __shfl_sync(0xffffffffu, t.y1, src) - ty0 + 1;
            const uint32_t pid = (uint32_t)__shfl_sync(0xffffffffu, (unsigned long long)i, src);
            for (uint32_t e = lane; e < cc; e += 32) {
                const uint32_t tz = e / (uint32_t)(nx_ * ny_), r = e - tz * (uint32_t)(nx_ * ny_);
                const uint32_t ty = r / (uint32_t)nx_, tx = r - ty * (uint32_t)nx_;
                SPHB_ASSERT((int64_t)(ob + e) < n_pairs);
                keys[ob + e] = (uint32_t)(((tz0 + tz) * v.nty + (ty0 + ty)) * v.ntx + (tx0 + tx));
                vals[ob + e] = pid;
            }
        }
    }
}

// keys >= n_tiles are the 0xffffffff padding of a planned call whose list is shorter than planned
__global__ void tile_end_kernel(const uint32_t *keys, int64_t n_pairs, uint32_t n_tiles, uint32_t *tile_end)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs || keys[i] >= n_tiles) return;
    if (i == n_pairs - 1 || keys[i] != keys[i + 1]) tile_end[keys[i]] = (uint32_t)(i + 1);
}

// bounds[s] = first sorted pair whose tile id is >= first_tile[s] (s = 0 .. n_bounds-1)
__global__ void stage_bounds_kernel(const uint32_t *keys, int64_t n_pairs, const uint32_t *first_tile, int n_bounds,
                                    uint32_t *bounds)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_bounds) return;
    const uint32_t want = first_tile[s];
    int64_t lo = 0, hi = n_pairs;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    bounds[s] = (uint32_t)lo;
}

template <typename R>
__global__ void lut_pairs_kernel(const double *lut, int samples, double scale, typename Real2<R>::type *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= samples) return;
    out[i] = lut_make_entry<R>(lut, samples, i, scale);
}

// ---------------------------------------------------------------------------
// Host side.

cudaError_t sm_count(int *out)
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

// Staged renders run consecutive stages on two internal streams (forked from and joined to the caller's
// stream with events), so that the tail of stage s -- its last, partly filled wave of CTAs -- overlaps the
// head of stage s + 1; the stages touch different tiles / cells and add with reductions, so they need no
// order among themselves.  Launched back to back on one stream, every extra stage cost ~0.19 ms of tail at
// the 340 us a cell-kernel CTA lives (2 GPUs, 4 stages: 20.1 -> 20.85 ms).  Streams and events are created
// once per host thread and device.
struct StageLanes {
    cudaStream_t lane[2];
    cudaEvent_t fork;
    cudaEvent_t done[SPHB_MAX_STAGES];
};
static cudaError_t stage_lanes(StageLanes **out)
{
    static thread_local StageLanes *cache[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidValue;
    if (!cache[dev]) {
        StageLanes *l = new StageLanes();
        for (int i = 0; i < 2; i++) {
            e = cudaStreamCreateWithFlags(&l->lane[i], cudaStreamNonBlocking);
            if (e != cudaSuccess) return e;
        }
        e = cudaEventCreateWithFlags(&l->fork, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        for (int i = 0; i < SPHB_MAX_STAGES; i++) {
            e = cudaEventCreateWithFlags(&l->done[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        cache[dev] = l;
    }
    *out = cache[dev];
    return cudaSuccess;
}

// Page-locked landing zone of the bin totals, one per host thread and device.
static cudaError_t totals_landing(unsigned long long **out)
{
    static thread_local unsigned long long *cache[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidValue;
    if (!cache[dev]) {
        e = cudaHostAlloc(reinterpret_cast<void **>(&cache[dev]), kTotCount * sizeof(unsigned long long),
                          cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
    }
    *out = cache[dev];
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

// Footprint boxes up to this many pixels per edge take the cell path.  Bounded by the
// shared-memory accumulator of a warp ((cell + edge - 1)^d values per weight); the
// default can be lowered with SPHB_SMALL_EDGE (0 sends everything through the tile path).
static int small_edge_setting(bool dim3, bool f32)
{
    int cap = dim3 ? (f32 ? 8 : 6) : 8;
    const char *e = getenv("SPHB_SMALL_EDGE");
    if (e && *e) {
        int val = atoi(e);
        cap = val < 0 ? 0 : val > cap ? cap : val;
    }
    return cap;
}

// Cell size (log2) of the small-footprint path.  Images: 32 x 32 pixels (16 x 16 with three or four
// weights in double).  Grids: 4 x 4 x 4 voxels -- a warp's accumulator is 4 KB at the 5-voxel boxes of
// BASELINE config 4 and 20 warps share an SM (with 8 x 8 x 8 it is 14-27 KB and 8-11 warps fit: 53 against
// 42 ms when that was measured), although it flushes 8 instead of 3.4 reductions per voxel.  Sparse sets
// (fewer than 32 particles per 4^3 cell: a 1/8 shard of config 4 holds 8) flush too often with that;
// their cells are twice as deep, 4 x 4 x 8 (profiles/exp_r2_cellshape.sh, 1/8 shard: 4^3 5.61, 4x4x8 5.12,
// 4x8x8 5.13, 8x8x4 5.45, 4x4x16 5.63, 8^3 6.21 ms; full set: 4^3 30.0, 4x4x8 31.3, 8x8x4 33.8 ms).
static int cell_log(bool dim3, bool f32, int nw, int small_edge)
{
    const char *e = getenv("SPHB_CELL_LOG");
    if (e && *e) return std::min(std::max(atoi(e), 1), 5);
    if (dim3) return 2;
    const size_t elem = f32 ? 4 : 8;
    const size_t s = 32 + (small_edge > 0 ? small_edge - 1 : 0);
    return s * s * elem * nw <= 28 * 1024 ? 5 : 4;
}

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
    if ((int64_t)r->nx * r->ny * (dim3 ? r->nz : 1) > 0x7fffffff00LL || r->nx > 0x7fff00 || r->ny > 0x7fff00 ||
        (dim3 && r->nz > 0x7fff00)) {
        set_error("raster too large");
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

struct Stages {
    int n;                    // 1 = one launch covers everything
    const int *planes;        // planes[s] = first z-plane NOT finished by stage s (ascending; last >= nz)
    sphb_stage_fn fn;
    void *user;
};

// Geometry of a call: tiles, cells, record stride.
static int build_view(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, bool dim3, int use_z,
                      double z_slice, View &v, int64_t &n_tiles, int64_t &n_cells, int &end_bit)
{
    const bool f32 = p->dtype == SPHB_F32;
    v.nx = r->nx;
    v.ny = r->ny;
    v.nz = dim3 ? r->nz : 1;
    v.tx = dim3 ? kTile3X : kTile2X;
    v.ty = dim3 ? kTile3Y : tile2y(f32, p->n_weights);
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
    v.small_edge = small_edge_setting(dim3, f32);
    v.rec_stride = rec_stride_for(f32, dim3 || use_z, p->n_weights);
    const int clog = cell_log(dim3, f32, p->n_weights, v.small_edge);
    v.cx_log = v.cy_log = clog;
    v.cz_log = dim3 ? clog : 0;
    if (dim3 && (double)p->n * 64.0 < 32.0 * ((double)v.nx * v.ny * v.nz)) v.cz_log = clog + 1; // sparse set
    if (const char *e = getenv("SPHB_CELL_ZLOG"))
        if (*e && dim3) v.cz_log = std::min(std::max(atoi(e), 1), 5);
    if (const char *e = getenv("SPHB_CELL_YLOG"))
        if (*e && dim3) v.cy_log = std::min(std::max(atoi(e), 1), 5);
    v.ncx = div_up(v.nx, 1 << v.cx_log);
    v.ncy = div_up(v.ny, 1 << v.cy_log);
    v.ncz = dim3 ? div_up(v.nz, 1 << v.cz_log) : 1;
    v.halo = v.small_edge > 0 ? v.small_edge - 1 : 0; // tightened below to the largest box seen
    n_tiles = (int64_t)v.ntx * v.nty * v.ntz;
    n_cells = (int64_t)v.ncx * v.ncy * v.ncz;
    if (n_tiles > 0x3fffffffLL || n_cells > 0x3fffffffLL) {
        set_error("raster has too many tiles");
        return SPHB_ERR_ARGUMENT;
    }
    end_bit = 1;
    while ((1LL << end_bit) < n_tiles) end_bit++;
    v.n_out = (size_t)v.nx * v.ny * v.nz;

    return SPHB_OK;
}

template <typename T>
static int scatter_impl(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, bool dim3,
                        int use_z, double z_slice, int accumulate, double *const *out, void *ws,
                        size_t ws_bytes, size_t *ws_needed, sphb_stats *stats, const Stages &stages,
                        const sphb_plan *plan, sphb_plan *plan_out, cudaStream_t stream)
{
    const bool f32 = p->dtype == SPHB_F32;
    View v;
    int64_t n_tiles = 0, n_cells = 0;
    int end_bit = 1;
    {
        const int rc = build_view(p, k, r, dim3, use_z, z_slice, v, n_tiles, n_cells, end_bit);
        if (rc) return rc;
    }
    const int64_t n = p->n;
    const size_t n_out = (size_t)v.n_out;

    if (!accumulate && !plan_out)
        for (int i = 0; i < p->n_weights; i++)
            SPHB_CUDA(cudaMemsetAsync(out[i], 0, n_out * sizeof(double), stream));
    const bool timed = stats && stats->time_phases && !plan && !plan_out;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (stats) {
        stats->n_pairs = 0;
        stats->n_tiles = n_tiles;
        stats->tile_x = v.tx;
        stats->tile_y = v.ty;
        stats->tile_z = v.tz;
        stats->ms_bin = stats->ms_sort = stats->ms_render = stats->ms_direct = stats->ms_count = 0.f;
        stats->n_direct = 0;
        stats->launches = 0;
    }
    if (timed) SPHB_CUDA(phase_events(ev, 6));
    if (ws_needed) *ws_needed = 0;
    auto all_stages_done = [&]() {
        if (stages.fn)
            for (int s = 0; s < stages.n; s++) stages.fn(stages.user, s);
    };
    if (n == 0) {
        if (plan_out) memset(plan_out, 0, sizeof(*plan_out));
        all_stages_done();
        return SPHB_OK;
    }

    // ---- phase A: count -------------------------------------------------------
    size_t scan_tmp = 0;
    SPHB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                            n_cells + 1, stream));
    Workspace a(ws);
    // one contiguous block cleared by one memset: totals, cell histogram, cell cursors
    const size_t zero_words = 2 * kTotCount + (size_t)(n_cells + 1) + (size_t)n_cells;
    uint32_t *zero_block = a.take<uint32_t>(zero_words);
    unsigned long long *totals = reinterpret_cast<unsigned long long *>(zero_block);
    uint32_t *cell_start = zero_block + 2 * kTotCount; // histogram, scanned in place
    uint32_t *cell_cursor = cell_start + (n_cells + 1);
    unsigned char *scan_ws = a.take<unsigned char>(scan_tmp);
    const size_t need_a = align_up(a.used, 256);
    if (ws_needed) *ws_needed = need_a;
    if (!plan && (!ws || ws_bytes < need_a)) {
        set_error("workspace of %zu bytes needed for binning, %zu given", need_a, ws_bytes);
        return SPHB_ERR_WORKSPACE;
    }
    const T *x = static_cast<const T *>(p->x), *y = static_cast<const T *>(p->y);
    const T *z = static_cast<const T *>(p->z), *h = static_cast<const T *>(p->h);
    const int blocks = div_up(n, 256);
    int sms = 0;
    SPHB_CUDA(sm_count(&sms));
    auto run_count = [&]() -> int {
        if (timed) SPHB_CUDA(cudaEventRecord(ev[0], stream));
        SPHB_CUDA(cudaMemsetAsync(zero_block, 0, zero_words * sizeof(uint32_t), stream));
        count_kernel<T><<<std::min<int64_t>(blocks, (int64_t)sms * 16), 256, 0, stream>>>(v, x, y, z, h, n,
                                                                                        cell_start, totals);
        SPHB_LAUNCH_CHECK();
        SPHB_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan_tmp, cell_start, cell_start, n_cells + 1, stream));
        // the cursors start at the cell starts: the emit kernel's one atomic per particle then returns the final
        // position, with no second random read of cell_start
        SPHB_CUDA(cudaMemcpyAsync(cell_cursor, cell_start, (size_t)n_cells * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                  stream));
        if (timed) SPHB_CUDA(cudaEventRecord(ev[5], stream));
        return SPHB_OK;
    };
    if (!plan) {   // a planned call counts after its whole workspace has been checked (below)
        const int rc = run_count();
        if (rc) return rc;
    }
    if (plan_out) {
        // plan call: the totals travel to the caller's (page-locked) plan, asynchronously; nothing else
        static_assert(sizeof(sphb_plan) == kTotCount * sizeof(unsigned long long), "plan mirrors the totals");
        SPHB_CUDA(cudaMemcpyAsync(plan_out, totals, sizeof(sphb_plan), cudaMemcpyDeviceToHost, stream));
        return SPHB_OK;
    }
    unsigned long long plan_tot[kTotCount] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned long long *totals_host = plan_tot;
    if (plan) {
        // planned call: list sizes and the accumulator halo come from the caller; no read-back, no sync
        plan_tot[kTotSmall] = plan->n_small;
        plan_tot[kTotPairs] = plan->n_pairs;
        plan_tot[kTotEdge] = plan->max_small_edge;
        plan_tot[kTotLarge] = plan->n_large;
    } else {
        unsigned long long *landing = nullptr;
        SPHB_CUDA(totals_landing(&landing));
        SPHB_CUDA(cudaMemcpyAsync(landing, totals, kTotCount * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                  stream));
        SPHB_CUDA(cudaStreamSynchronize(stream));
        totals_host = landing;
    }
    const int64_t n_small = (int64_t)totals_host[kTotSmall];
    const int64_t n_large = (int64_t)totals_host[kTotLarge];
    const uint64_t total_pairs = totals_host[kTotPairs];
    if (totals_host[kTotEdge] > 0) v.halo = (int)totals_host[kTotEdge] - 1;
    if (stats) {
        stats->n_pairs = (int64_t)total_pairs;
        stats->n_direct = n_small;
    }
    if (total_pairs == 0 && n_small == 0) {
        all_stages_done();
        return SPHB_OK;
    }
    if (total_pairs > 0x7fffffffULL) {
        set_error("%llu (particle, tile) pairs exceed 2^31-1: split the particle set",
                  (unsigned long long)total_pairs);
        if (ws_needed) *ws_needed = 0;
        return SPHB_ERR_TOO_MANY;
    }
    const int64_t n_pairs = (int64_t)total_pairs;

    // ---- phase B: emit, sort, segment ------------------------------------
    size_t sort_tmp = 0;
    if (n_pairs > 0)
        SPHB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (uint32_t *)nullptr, (uint32_t *)nullptr,
                                                  (uint32_t *)nullptr, (uint32_t *)nullptr, n_pairs, 0,
                                                  plan ? 32 : end_bit, stream));
    uint32_t *keys_a = a.take<uint32_t>(n_pairs);
    uint32_t *vals_a = a.take<uint32_t>(n_pairs);
    uint32_t *keys_b = a.take<uint32_t>(n_pairs);
    uint32_t *vals_b = a.take<uint32_t>(n_pairs);
    uint32_t *tile_end = a.take<uint32_t>(n_pairs > 0 ? n_tiles : 0);
    uint32_t *stage_tile = a.take<uint32_t>(2 * (stages.n + 1)); // first tile of each stage, then pair bounds
    uint32_t *stage_pos = stage_tile + (stages.n + 1);
    T *recs_large = a.take<T>(n_large > 0 ? (size_t)n * v.rec_stride : 0);
    T *recs_small = a.take<T>((size_t)n_small * v.rec_stride);
    const bool big_lut = k->kind == SPHB_COLUMN_LUT && k->lut_samples > kLutSharedMax;
    void *lut_global = big_lut ? static_cast<void *>(a.take<double2>(k->lut_samples)) : nullptr;
    unsigned char *sort_ws = a.take<unsigned char>(sort_tmp);
    const size_t need_b = align_up(a.used, 256);
    if (ws_needed) *ws_needed = need_b;
    if (!ws || ws_bytes < need_b) {
        set_error("workspace of %zu bytes needed for %lld pairs and %lld cell-path particles, %zu given", need_b,
                  (long long)n_pairs, (long long)n_small, ws_bytes);
        return SPHB_ERR_WORKSPACE;
    }
    if (plan) {
        const int rc = run_count();
        if (rc) return rc;
    }
    if (plan && n_pairs > 0)   // entries the emit kernel does not fill sort behind every tile
        SPHB_CUDA(cudaMemsetAsync(keys_a, 0xff, (size_t)n_pairs * sizeof(uint32_t), stream));
    emit_kernel<T><<<div_up(n, 256 * kEmitPer), 256, 0, stream>>>(
        v, x, y, z, h, static_cast<const T *>(p->w[0]), static_cast<const T *>(p->w[1]),
        static_cast<const T *>(p->w[2]), static_cast<const T *>(p->w[3]), p->n_weights, n, cell_cursor,
        totals, keys_a, vals_a, recs_large, recs_small, n_pairs, n_small,
        getenv("SPHB_EMIT_MODE") ? atoi(getenv("SPHB_EMIT_MODE")) : 0);
    SPHB_LAUNCH_CHECK();
    if (timed) SPHB_CUDA(cudaEventRecord(ev[1], stream));
    if (n_pairs > 0) {
        SPHB_CUDA(cub::DeviceRadixSort::SortPairs(sort_ws, sort_tmp, keys_a, keys_b, vals_a, vals_b, n_pairs, 0,
                                                  plan ? 32 : end_bit, stream));
        tile_end_kernel<<<div_up(n_pairs, 256), 256, 0, stream>>>(keys_b, n_pairs, (uint32_t)n_tiles, tile_end);
        SPHB_LAUNCH_CHECK();
    }
    if (big_lut) {
        if (f32)
            lut_pairs_kernel<float><<<div_up(k->lut_samples, 256), 256, 0, stream>>>(
                k->lut, k->lut_samples, (double)(k->lut_samples - 1) / v.radius, static_cast<float2 *>(lut_global));
        else
            lut_pairs_kernel<double><<<div_up(k->lut_samples, 256), 256, 0, stream>>>(
                k->lut, k->lut_samples, (double)(k->lut_samples - 1) / v.radius, static_cast<double2 *>(lut_global));
        SPHB_LAUNCH_CHECK();
    }

    // ---- stages: z-planes -> tile and cell boundaries ---------------------------
    // Stage s renders every tile / cell whose lowest plane lies below planes[s]; a box listed under
    // a tile / cell never reaches planes below it, so the planes below the boundary are final.
    int launches = 2 + (n_pairs > 0 ? 1 : 0);
    std::vector<int64_t> cell_bound(stages.n + 1, 0);
    if (stages.n > 1) {
        std::vector<uint32_t> first(stages.n + 1, 0);
        for (int s = 0; s < stages.n; s++) {
            const int plane = s + 1 < stages.n ? std::min(std::max(stages.planes[s], 0), v.nz) : v.nz;
            const int64_t tz = std::min<int64_t>(div_up(plane, v.tz), v.ntz);
            const int64_t cz = std::min<int64_t>(div_up(plane, 1 << v.cz_log), v.ncz);
            first[s + 1] = (uint32_t)(tz * v.nty * v.ntx);
            cell_bound[s + 1] = cz * v.ncy * v.ncx;
        }
        first[stages.n] = (uint32_t)n_tiles;
        cell_bound[stages.n] = n_cells;
        if (n_pairs > 0) {
            SPHB_CUDA(cudaMemcpyAsync(stage_tile, first.data(), (stages.n + 1) * sizeof(uint32_t),
                                      cudaMemcpyHostToDevice, stream));
            stage_bounds_kernel<<<1, 64, 0, stream>>>(keys_b, n_pairs, stage_tile, stages.n + 1, stage_pos);
            SPHB_LAUNCH_CHECK();
            launches++;
        }
    } else {
        cell_bound[1] = n_cells;
    }

    // ---- phase C: render ---------------------------------------------------
    if (timed) SPHB_CUDA(cudaEventRecord(ev[2], stream));
    const int kind = big_lut ? SPHB_COLUMN_LUT_GLOBAL : k->kind;
    // the two kernels of a stage; with one stage the tile kernel's time is bracketed on its own
    StageLanes *lanes = nullptr;
    if (stages.n > 1 && !timed) {
        SPHB_CUDA(stage_lanes(&lanes));
        SPHB_CUDA(cudaEventRecord(lanes->fork, stream));
        SPHB_CUDA(cudaStreamWaitEvent(lanes->lane[0], lanes->fork, 0));
        SPHB_CUDA(cudaStreamWaitEvent(lanes->lane[1], lanes->fork, 0));
    }
    for (int s = 0; s < stages.n; s++) {
        cudaStream_t ss = lanes ? lanes->lane[s & 1] : stream;
        if (n_pairs > 0) {
            int rc = launch_render(v, f32, kind, p->n_weights, recs_large, k, lut_global, keys_b, vals_b, tile_end,
                                   n_pairs, stages.n > 1 ? stage_pos + s : nullptr, out, ss);
            if (rc) return rc;
            launches++;
        }
        if (timed && s == 0) SPHB_CUDA(cudaEventRecord(ev[4], stream));
        if (n_small > 0 && cell_bound[s + 1] > cell_bound[s]) {
            int rc = launch_cells(v, f32, kind, p->n_weights, recs_small, k, lut_global, cell_start, n_small,
                                  cell_bound[s], cell_bound[s + 1], out, ss);
            if (rc) return rc;
            launches++;
        }
        if (lanes) {
            // the caller's stream joins stage s (and has joined every earlier one): what the callback
            // records there marks the planes below this stage's boundary final
            SPHB_CUDA(cudaEventRecord(lanes->done[s], ss));
            SPHB_CUDA(cudaStreamWaitEvent(stream, lanes->done[s], 0));
        }
        if (stages.fn) stages.fn(stages.user, s);
    }
    if (stats) stats->launches = launches;
    if (timed) {
        SPHB_CUDA(cudaEventRecord(ev[3], stream));
        SPHB_CUDA(cudaEventSynchronize(ev[3]));
        SPHB_CUDA(cudaEventElapsedTime(&stats->ms_bin, ev[0], ev[1]));
        SPHB_CUDA(cudaEventElapsedTime(&stats->ms_count, ev[0], ev[5]));
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
    const Stages one = {1, nullptr, nullptr, nullptr};
    if (p->dtype == SPHB_F32)
        return scatter_impl<float>(p, k, r, false, use_z, z_slice, accumulate, out, ws, ws_bytes, ws_needed,
                                   stats, one, nullptr, nullptr, s);
    return scatter_impl<double>(p, k, r, false, use_z, z_slice, accumulate, out, ws, ws_bytes, ws_needed,
                                stats, one, nullptr, nullptr, s);
}

extern "C" int sphb_scatter_plan(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, int dim3,
                                 int use_z, double z_slice, void *ws, size_t ws_bytes, size_t *ws_needed,
                                 sphb_plan *plan, sphb_stream stream)
{
    double *dummy[SPHB_MAX_WEIGHTS] = {reinterpret_cast<double *>(8), reinterpret_cast<double *>(8),
                                       reinterpret_cast<double *>(8), reinterpret_cast<double *>(8)};
    int rc = validate(p, k, r, dim3 != 0, dim3 ? 0 : use_z, dummy);
    if (rc) return rc;
    if (!plan) {
        set_error("null plan");
        return SPHB_ERR_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const Stages one = {1, nullptr, nullptr, nullptr};
    if (p->dtype == SPHB_F32)
        return scatter_impl<float>(p, k, r, dim3 != 0, dim3 ? 0 : use_z, z_slice, 1, dummy, ws, ws_bytes, ws_needed,
                                   nullptr, one, nullptr, plan, s);
    return scatter_impl<double>(p, k, r, dim3 != 0, dim3 ? 0 : use_z, z_slice, 1, dummy, ws, ws_bytes, ws_needed,
                                nullptr, one, nullptr, plan, s);
}

extern "C" int sphb_scatter_planned(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r, int dim3,
                                    int use_z, double z_slice, int accumulate, double *const *out, void *ws,
                                    size_t ws_bytes, size_t *ws_needed, const sphb_plan *plan, sphb_stream stream)
{
    int rc = validate(p, k, r, dim3 != 0, dim3 ? 0 : use_z, out);
    if (rc) return rc;
    if (!plan || plan->n_pairs > 0x7fffffffULL) {
        set_error(plan ? "plan holds more than 2^31-1 pairs: split the particle set" : "null plan");
        return plan ? SPHB_ERR_TOO_MANY : SPHB_ERR_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const Stages one = {1, nullptr, nullptr, nullptr};
    if (p->dtype == SPHB_F32)
        return scatter_impl<float>(p, k, r, dim3 != 0, dim3 ? 0 : use_z, z_slice, accumulate, out, ws, ws_bytes,
                                   ws_needed, nullptr, one, plan, nullptr, s);
    return scatter_impl<double>(p, k, r, dim3 != 0, dim3 ? 0 : use_z, z_slice, accumulate, out, ws, ws_bytes,
                                ws_needed, nullptr, one, plan, nullptr, s);
}

extern "C" int sphb_scatter_overflowed(const void *ws, int *flag, sphb_stream stream)
{
    if (!ws || !flag) {
        set_error("null argument");
        return SPHB_ERR_ARGUMENT;
    }
    // the totals block opens the workspace (see scatter_impl)
    unsigned long long v = 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SPHB_CUDA(cudaMemcpyAsync(&v, static_cast<const unsigned long long *>(ws) + kTotOverflow, sizeof(v),
                              cudaMemcpyDeviceToHost, s));
    SPHB_CUDA(cudaStreamSynchronize(s));
    *flag = v != 0;
    return SPHB_OK;
}

extern "C" int sphb_scatter3d_staged(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r,
                                     int accumulate, double *const *out, void *ws, size_t ws_bytes,
                                     size_t *ws_needed, sphb_stats *stats, int n_stages, const int *stage_planes,
                                     sphb_stage_fn on_stage, void *user, sphb_stream stream)
{
    int rc = validate(p, k, r, true, 0, out);
    if (rc) return rc;
    if (n_stages < 1 || n_stages > SPHB_MAX_STAGES || (n_stages > 1 && !stage_planes)) {
        set_error("n_stages must be 1..%d (with the stage planes given)", SPHB_MAX_STAGES);
        return SPHB_ERR_ARGUMENT;
    }
    for (int s = 1; s < n_stages; s++)
        if (stage_planes[s] < stage_planes[s - 1]) {
            set_error("stage planes must ascend");
            return SPHB_ERR_ARGUMENT;
        }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const Stages st = {n_stages, stage_planes, on_stage, user};
    if (p->dtype == SPHB_F32)
        return scatter_impl<float>(p, k, r, true, 0, 0.0, accumulate, out, ws, ws_bytes, ws_needed, stats, st, nullptr,
                                   nullptr, s);
    return scatter_impl<double>(p, k, r, true, 0, 0.0, accumulate, out, ws, ws_bytes, ws_needed, stats, st, nullptr,
                                nullptr, s);
}

extern "C" int sphb_scatter3d(const sphb_particles *p, const sphb_kernel *k, const sphb_raster *r,
                              int accumulate, double *const *out, void *ws, size_t ws_bytes,
                              size_t *ws_needed, sphb_stats *stats, sphb_stream stream)
{
    return sphb_scatter3d_staged(p, k, r, accumulate, out, ws, ws_bytes, ws_needed, stats, 1, nullptr, nullptr,
                                 nullptr, stream);
}
