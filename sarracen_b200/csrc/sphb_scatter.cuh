// Shared declarations of the sampled scatter pipeline (sphb_scatter.cu: binning and host
// orchestration; sphb_render.cu: tile kernel for large footprints; sphb_cell.cu: cell kernel for
// small footprints).
#pragma once
#include <mutex>
#include <set>
#include <utility>

#include "sphb_common.cuh"

namespace sphb {

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
    int rec_stride;    // elements per packed particle record (4, 6 or 8)
    unsigned long long n_out;   // raster elements per weight (debug index checks)
    int small_edge;    // footprint boxes with every edge <= this many pixels take the cell path
    // cells of the small-footprint path: a particle is listed under the cell that holds the
    // ORIGIN (lowest corner) of its footprint box
    int cx_log, cy_log, cz_log; // cell size in pixels (powers of two)
    int ncx, ncy, ncz;          // cells per axis
    int halo;                   // largest small box edge - 1: how far a box reaches beyond its cell
};

// Packed particle record, written once per call by the emit kernel (which streams the
// columns anyway).  Large-footprint particles keep their input order (the sorted (tile,
// particle) pairs point at them); small-footprint particles are written in CELL ORDER, so the
// cell kernel reads them as one coalesced stream.
//   32 bytes: double { x, y, h, w0 }                   (no depth column, one weight)
//             float  { x, y, h, z, w0, w1, w2, w3 }    (always)
//   64 bytes: double { x, y, h, z, w0, w1, w2, w3 }
// A record is one or two whole, aligned 32-byte sectors and is moved with 256-bit accesses
// (sm_100: LDG / STG.E.ENL2.256): a record placed on its own in the cell-ordered stream then
// never leaves a partially written sector behind -- with 48-byte records written as three
// 16-byte stores the placement pass ran at 1.5 TB/s, the scattered sub-sector writes
// costing 6 of its 8.8 ms at 2^27 particles.
template <typename T> struct Rec {
    T x, y, h, z, w[4];
};

__device__ __forceinline__ void ldg256(const double *p, double (&f)[4])
{
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(f[0]), "=d"(f[1]), "=d"(f[2]), "=d"(f[3]) : "l"(p));
}
__device__ __forceinline__ void ldg256(const float *p, float (&f)[8])
{
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]), "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7])
        : "l"(p));
}
__device__ __forceinline__ void stg256(double *p, const double *f)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(f[0]), "d"(f[1]), "d"(f[2]), "d"(f[3])
                 : "memory");
}
__device__ __forceinline__ void stg256(float *p, const float *f)
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(f[0]), "f"(f[1]), "f"(f[2]),
                 "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7])
                 : "memory");
}

// Record stride in elements: 4 (double, image without depth, one weight) or 8.
inline int rec_stride_for(bool f32, bool has_z, int n_weights) { return (!f32 && !has_z && n_weights == 1) ? 4 : 8; }

__device__ __forceinline__ Rec<double> load_rec(const double *__restrict__ recs, int stride, int64_t p, int nw)
{
    Rec<double> r;
    const double *b = recs + (size_t)p * stride;
    double a[4];
    ldg256(b, a);
    r.x = a[0]; r.y = a[1]; r.h = a[2];
    if (stride == 4) {
        r.z = 0.0; r.w[0] = a[3];
        r.w[1] = r.w[2] = r.w[3] = 0.0;
    } else {
        r.z = a[3];
        double c[4];
        ldg256(b + 4, c);
        r.w[0] = c[0]; r.w[1] = c[1]; r.w[2] = c[2]; r.w[3] = c[3];
    }
    (void)nw;
    return r;
}
__device__ __forceinline__ Rec<float> load_rec(const float *__restrict__ recs, int stride, int64_t p, int nw)
{
    Rec<float> r;
    float a[8];
    ldg256(recs + (size_t)p * 8, a);
    r.x = a[0]; r.y = a[1]; r.h = a[2]; r.z = a[3];
    r.w[0] = a[4]; r.w[1] = a[5]; r.w[2] = a[6]; r.w[3] = a[7];
    (void)stride; (void)nw;
    return r;
}

// Pixel box [lo, hi) that holds every pixel centre within rad of c.  The roundings are the
// converting ones (cvt.rpi / cvt.rmi.s32.f64: they saturate, and map NaN to 0) followed by an
// integer clamp -- one conversion-pipe instruction per bound instead of a rounding, two double
// min / max and a conversion.
__device__ __forceinline__ void pixel_range(double c, double rad, double origin, double inv_pw, int n,
                                            int &lo, int &hi)
{
    // centre of pixel i sits at origin + (i + 1/2) pw
    lo = min(max(__double2int_ru((c - rad - origin) * inv_pw - 0.5), 0), n);
    hi = min(max(__double2int_rd((c + rad - origin) * inv_pw - 0.5), -1), n - 1) + 1;
}

// Opt a kernel in to the full 227 KB of dynamic shared memory, once per instantiation and
// device (driver calls per launch are kept to a minimum: they serialise with other
// driver clients such as NVML monitors).
template <typename K> inline cudaError_t allow_large_smem(K kern)
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

// Column tables of up to kLutSharedMax samples are staged in shared memory by every CTA; larger
// ones (integral_samples is a user argument, base_kernel.py:68) are expanded once per call into
// segment lines (A_i, B_i) (sphb_common.cuh, lut_index) in the workspace and read through L1/L2.
constexpr int kLutSharedMax = 2048;

// Tile sizes of the large-footprint path.  2-D: 32 x 128 pixels, 8 warps with a 16-row strip each
// (16 register accumulators per lane and weight); 3-D: 16 x 16 x 8 voxels, a z-plane per warp.
constexpr int kTile2X = 32, kTile2Y = 128;
constexpr int kTile3X = 16, kTile3Y = 16, kTile3Z = 8;
constexpr int kChunkPairs = 2048; // sorted pairs per CTA of the tile kernel

// Height of the 2-D tile.  With three or four weight columns in double a 32 x 128 tile needs
// 111 / 148 KB of accumulators and 96 / 128 accumulator registers per lane: one CTA per SM, with
// spills.  Half the height brings back two CTAs per SM and spill-free code.
inline int tile2y(bool f32, int n_weights) { return (!f32 && n_weights >= 3) ? kTile2Y / 2 : kTile2Y; }

// Cells of the small-footprint path: 32 x 32 pixels (images), 8 x 8 x 8 voxels (grids).
constexpr int kCell2 = 32, kCell3 = 8;
#ifndef SPHB_CELL_CHUNK
#define SPHB_CELL_CHUNK 512
#endif
constexpr int kCellChunk = SPHB_CELL_CHUNK; // cell-ordered particles per warp work unit

// sphb_render.cu: tile kernel over sorted pairs [range[0], range[1]) (range == nullptr: all n_pairs).
int launch_render(const View &v, bool f32, int kind, int nw, const void *recs, const sphb_kernel *k,
                  const void *lut_global, const uint32_t *keys, const uint32_t *vals, const uint32_t *tile_end,
                  int64_t n_pairs, const uint32_t *range, double *const *out, cudaStream_t stream);

// sphb_cell.cu: cell kernel over the cell-ordered records of cells [cell_lo, cell_hi).
int launch_cells(const View &v, bool f32, int kind, int nw, const void *recs_small, const sphb_kernel *k,
                 const void *lut_global, const uint32_t *cell_start, int64_t n_small, int64_t cell_lo,
                 int64_t cell_hi, double *const *out, cudaStream_t stream);

} // namespace sphb
