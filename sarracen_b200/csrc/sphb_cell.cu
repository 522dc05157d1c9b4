// Cell kernel of the sampled scatter: small footprints (sm_100a).
//
// A particle whose support box is at most small_edge pixels along every axis covers a handful of
// pixels; binning it per tile, staging it per tile and scanning it per warp costs far more than
// its few kernel evaluations, and one global reduction per contribution is bound by the
// reduction rate of the SM (1.29 cycles per lane).  Such particles are therefore counting-sorted
// by the CELL that holds the origin of their box (32 x 32 pixels / 4 x 4 x 4 voxels, 4 x 4 x 8 for
// sparse sets) -- the emit kernel writes their packed records in cell order -- and rendered here
// with a LOCAL ACCUMULATOR:
//
//   * a warp takes kCellChunk consecutive records of that stream, 32 at a time.  The stream is
//     contiguous, so a batch is ONE 1-D TMA bulk copy (cp.async.bulk global -> shared, completion
//     on an mbarrier; SASS UBLKCP).  Every lane derives one particle's box, scales and terms from
//     the staged record and parks them in the warp's staging slots; that empties the raw buffer,
//     so lane 0 issues the bulk copy of the next batch right then and it lands while this batch is
//     evaluated (one buffer per warp);
//   * the warp owns a private shared-memory accumulator that covers one cell plus the halo a
//     box can reach beyond it ((cell + halo)^d values per weight).  The batch is taken apart into
//     runs of one cell and, inside a run, classes of equal z-extent; each class is a tight loop
//     over its particles, one particle per pass with the lanes laid over the box plane and the
//     z-column unrolled at compile time, accumulated with plain load / fma / store: the
//     accumulator is exclusive to the warp, lanes of one particle touch distinct pixels, so no
//     atomics are needed.  Images and narrow / wide box planes take a generic per-particle path;
//   * when the stream moves on to another cell (or the chunk ends) the non-zero part of the
//     accumulator is added to the float64 raster with red.global.add.f64, coalesced along x,
//     and cleared: global atomics only at cell boundaries -- 8 per voxel instead of 57.6 per
//     particle for BASELINE config 4;
//   * a batch whose boxes all hold at most 16 pixels (h <~ pixel: 4 contributions per particle)
//     skips the accumulator: sub-groups of lanes take one particle each and reduce straight
//     into the raster.  When EVERY box of the call is that small the accumulator-free
//     instantiation runs instead, in which each lane walks the pixels of its own particle
//     straight from registers.
//
// Replaces the pixel loop of CPUBackend._fast_2d (cpu_backend.py:263-308) and the per-slice loop
// of CPUBackend.interpolate_3d_grid (:193-220) for small footprints.
#include "sphb_scatter.cuh"

namespace sphb {

#ifndef SPHB_CELL_MAXW
#define SPHB_CELL_MAXW 20
#endif
constexpr int kCellMaxWarps = SPHB_CELL_MAXW;
constexpr int kTinyVolume = 16; // a batch whose largest box holds <= this many pixels reduces directly

// ---- TMA: 1-D bulk copies global -> shared, completion on an mbarrier -------------------------
// The cell-ordered records of a warp's chunk are one contiguous stream, so the next batch of 32
// records (1-2 KB) is fetched by ONE cp.async.bulk issued by lane 0 (SASS: UBLKCP) while the warp
// evaluates the current batch; the lanes then read their record from shared memory.
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\t"
                 "bra WAIT_%=;\n\t"
                 "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// record of lane `i` out of a staged batch in shared memory (same layouts as load_rec)
__device__ __forceinline__ Rec<double> smem_rec(const double *raw, int stride, int i)
{
    Rec<double> r;
    const double2 *b = reinterpret_cast<const double2 *>(raw + (size_t)i * stride);
    const double2 a = b[0], c = b[1];
    r.x = a.x; r.y = a.y; r.h = c.x;
    if (stride == 4) {
        r.z = 0.0; r.w[0] = c.y;
        r.w[1] = r.w[2] = r.w[3] = 0.0;
    } else {
        r.z = c.y;
        const double2 d = b[2], e = b[3];
        r.w[0] = d.x; r.w[1] = d.y; r.w[2] = e.x; r.w[3] = e.y;
    }
    return r;
}
__device__ __forceinline__ Rec<float> smem_rec(const float *raw, int stride, int i)
{
    Rec<float> r;
    const float4 *b = reinterpret_cast<const float4 *>(raw + (size_t)i * 8);
    const float4 a = b[0], c = b[1];
    r.x = a.x; r.y = a.y; r.h = a.z; r.z = a.w;
    r.w[0] = c.x; r.w[1] = c.y; r.w[2] = c.z; r.w[3] = c.w;
    (void)stride;
    return r;
}

// Staging slots of one warp: one batch of 32 particles.  Everything a particle's pass needs is
// derived here once, lane-parallel, so that the pass itself starts with a few broadcast loads.
template <typename R, int NW> struct CellBatch {
    using R2 = typename Real2<R>::type;
    // pairs, so that a particle's pass starts with a few 16-byte broadcast loads:
    R2 xs[32], ys[32], zs[32]; // (s, o) per axis: pixel -> q scale (pixel width / h) and the q-space offset of
                               // the box origin, -(centre - origin) * s
    R2 ct[32];                 // (dz^2 / h^2 of a cross-section + sqrt guard, term[0])
    R term[NW][32];            // norm * w / h^ndim (all weights; term[0] also in ct)
    // x: box origin inside its cell, lx | ly << 8 | lz << 16;  y: box size wx | wy << 8 | wz << 16
    // (0 = nothing to do);  z: cell id;  w: lanes per z-slice (log2) | (65536 / wx + 1) << 8
    int4 geo[32];
    // ready-made for the accumulator pass.  x: element offset of the box origin in the accumulator;
    // y: wx wy | z-passes << 8;  z: 65536 / (wx wy) + 1 (tiny path);  w: unused
    int4 use[32];
    int bx[32], by[32], bz[32]; // box origin (raster pixels)
};

template <typename R> struct CellAcc {
    R *acc;         // the warp's accumulator, NW planes of `plane` elements
    int s1, s2;     // row and z-plane strides (elements)
    int rows;       // rows of one weight: plane / s1
    int plane;      // elements per weight
    int cell;       // cell being accumulated (-1: none)
    int cx0, cy0, cz0; // its origin in raster pixels
    bool dirty;
};

// Add the warp's accumulator to the raster and clear it.  Lanes are laid over x (several rows per
// pass when a row is shorter than a warp) and walk the rows of the accumulator; only non-zero
// entries reach the raster, so untouched entries -- including those that would lie outside it -- cost a
// shared-memory read and nothing else.
template <typename R, int NW>
__device__ __forceinline__ void cell_flush(CellAcc<R> &a, const View &v, int lane, double *const *outs)
{
    if (!a.dirty) return;
    __syncwarp();
    const int xlog = min(5, 32 - __clz(a.s1 - 1)); // lanes per row: power of two >= s1 (at most 32)
    const int rp = 32 >> xlog;                     // rows per pass
    const int sy = a.s2 / a.s1;                    // rows per z-plane
    for (int ix = lane & ((1 << xlog) - 1); ix < a.s1; ix += 32) {
        int iy = lane >> xlog, iz = 0;
        for (int r = iy; r < a.rows; r += rp) {
            while (iy >= sy) {
                iy -= sy;
                iz++;
            }
            const int la = r * a.s1 + ix;
#pragma unroll
            for (int k = 0; k < NW; k++) {
                const R val = a.acc[k * a.plane + la];
                if (val != R(0)) {
                    const size_t o = ((size_t)(a.cz0 + iz) * v.ny + (size_t)(a.cy0 + iy)) * v.nx + (size_t)(a.cx0 + ix);
                    SPHB_ASSERT(o < v.n_out && la < a.plane);
                    red_add(outs[k] + o, (double)val);
                    a.acc[k * a.plane + la] = R(0);
                }
            }
            iy += rp;
        }
    }
    __syncwarp();
    a.dirty = false;
}

// One lane's column of a particle's box: U z-slices (iz = zs, zs + sp, ...) evaluated without
// branches, so that the U dependent chains (q^2 -> square root -> polynomial / table) overlap; a
// slice outside the support evaluates to W = 0 and is not stored.
template <typename R, int KIND, int NW, int U, bool CHECKZ>
__device__ __forceinline__ void cell_column(R qxy, R dz0, R step, int zs, int sp, int wz, R rad2, const R *term,
                                            R *ap, int zstride, int plane, const LutView<R> &lv)
{
    // opaque copies: without them ptxas evaluates the multiples of step and zstride for all eight
    // instantiations ahead of the switch that selects one
    if (sizeof(R) == 8)
        asm volatile("" : "+d"(*reinterpret_cast<double *>(&step)));
    else
        asm volatile("" : "+f"(*reinterpret_cast<float *>(&step)));
    asm volatile("" : "+r"(zstride));
    R q2[U], wv[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const R dzs = u == 0 ? dz0 : fma(R(u), step, dz0);
        q2[u] = fma(dzs, dzs, qxy); // qxy carries sqrt_guard()
    }
#pragma unroll
    for (int u = 0; u < U; u++) wv[u] = kernel_eval_unguarded<KIND, R>(q2[u], rad2, lv);
#pragma unroll
    for (int u = 0; u < U; u++) {
        // CHECKZ = false: one z-slice per pass (sp = 1), the U passes are exactly the wz slices
        if ((!CHECKZ || zs + u * sp < wz) && lt_pos(q2[u], rad2)) {
            R *p = ap + u * zstride;
#pragma unroll
            for (int k = 0; k < NW; k++) p[k * plane] = fma(term[k], wv[u], p[k * plane]);
        }
    }
}

// The particles `m` (lanes of the batch) of one cell and one class: voxel boxes whose x-y plane takes the
// whole warp in one round (17 .. 32 pixels), U z-slices per lane.
template <typename R, int NW> struct CellPrep {
    using R2 = typename Real2<R>::type;
    R qxy;      // in-plane q^2 of this lane's pixel (+ sqrt guard)
    R2 zz;      // z scale and offset
    R term[NW];
    R *ap;      // this lane's accumulator column
    bool on;    // lane inside the box plane
};

template <typename R, int NW>
__device__ __forceinline__ CellPrep<R, NW> cell_prep(int s, const CellBatch<R, NW> &st, const CellAcc<R> &a, int lane)
{
    using R2 = typename Real2<R>::type;
    CellPrep<R, NW> p;
    const int4 geo = st.geo[s];
    const int4 use = st.use[s];
    const int wx = geo.y & 0xff, wxy = use.y & 0xff;
    SPHB_ASSERT((geo.x & 0xff) + wx <= a.s1 && ((geo.x >> 16) + (geo.y >> 16)) * a.s2 <= a.plane && wxy <= 32);
    const unsigned inv_wx = (unsigned)geo.w >> 8;
    const R2 xs = st.xs[s], ys = st.ys[s], ct = st.ct[s];
    p.zz = st.zs[s];
    p.term[0] = ct.y;
#pragma unroll
    for (int k = 1; k < NW; k++) p.term[k] = st.term[k][s];
    const int iy = (int)(((unsigned)lane * inv_wx) >> 16), ix = lane - iy * wx;
    const R dxs = fma((R)ix, xs.x, xs.y), dys = fma((R)iy, ys.x, ys.y);
    p.qxy = fma(dys, dys, fma(dxs, dxs, ct.x)); // ct.x carries sqrt_guard()
    p.ap = a.acc + (use.x + iy * a.s1 + ix);
    p.on = lane < wxy;
    return p;
}

template <typename R, int KIND, int NW, int U>
__device__ __forceinline__ void cell_class(unsigned m, const CellBatch<R, NW> &st, const CellAcc<R> &a, int lane,
                                           R rad2, int plane, const LutView<R> &lv)
{
    // (fetching the next particle's staged values ahead of this particle's column -- a software pipeline
    // across the __syncwarp -- was measured: 118 registers and 33.6 instead of 32.3 ms, not kept)
    while (m) {
        const CellPrep<R, NW> cur = cell_prep<R, NW>(__ffs(m) - 1, st, a, lane);
        m &= m - 1;
        if (cur.on)
            cell_column<R, KIND, NW, U, false>(cur.qxy, cur.zz.y, cur.zz.x, 0, 1, U, rad2, cur.term, cur.ap, a.s2, plane, lv);
        __syncwarp();
    }
}

// ACC = false: every box of the call holds at most kTinyVolume pixels, so the accumulator is never
// used; that variant is built without it (fewer registers, 8-warp CTAs, several CTAs per SM: the
// direct reductions need many warps in flight, not shared memory).
template <typename T, typename R, int KIND, int NW, bool DIM3, bool ACC>
__global__ void __launch_bounds__(ACC ? kCellMaxWarps * 32 : 256, ACC ? 1 : 4)
cell_kernel(View v, const T *__restrict__ recs, const double *__restrict__ lut, int lut_samples,
            const void *__restrict__ lut_global, const uint32_t *__restrict__ cell_start, int64_t cell_lo,
            int64_t cell_hi, int plane, int lut_bytes, int warp_bytes, double *o0, double *o1, double *o2,
            double *o3)
{
    using R2 = typename Real2<R>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, warps = blockDim.x >> 5;
    R2 *lut_s = reinterpret_cast<R2 *>(smem_raw);
    unsigned char *mine = smem_raw + lut_bytes + (size_t)warp * warp_bytes;
    CellBatch<R, NW> &st = *reinterpret_cast<CellBatch<R, NW> *>(mine);
    double *outs[4] = {o0, o1, o2, o3};

    LutView<R> lv;
    lut_set_table(lv, lut_s);
    lv.gtab = static_cast<const R2 *>(lut_global);
    lv.scale = R(0);
    lv.last = 0;
    if (is_lut<KIND>()) {
        const double lut_scale = (double)(lut_samples - 1) / v.radius;
        for (int i = tid; i < (KIND == SPHB_COLUMN_LUT_GLOBAL ? 0 : lut_samples); i += blockDim.x)
            lut_s[i] = lut_make_entry<R>(lut, lut_samples, i, lut_scale);
        lv.scale = (R)lut_scale;
        lv.last = lut_samples - 1;
        __syncthreads();
    }

    // this warp's slice of the cell-ordered stream
    const int64_t p_lo = cell_start[cell_lo], p_hi = cell_start[cell_hi];
    const int64_t start = p_lo + ((int64_t)blockIdx.x * warps + warp) * kCellChunk;
    if (start >= p_hi) return;
    const int64_t end = min(start + (int64_t)kCellChunk, p_hi);

    // one raw record batch + its mbarrier, behind the staging slots.  The staging step empties the buffer
    // (every lane turns its record into staging slots), so the bulk copy of the next batch is issued right
    // after it and lands while this batch's particles are evaluated: one buffer is enough, and 2 KB less per
    // warp is what lets 20 warps share an SM's shared memory.
    const int raw_bytes = 32 * v.rec_stride * (int)sizeof(T);
    T *const raw0 = reinterpret_cast<T *>(mine + sizeof(CellBatch<R, NW>));
    uint64_t *bar = reinterpret_cast<uint64_t *>(mine + sizeof(CellBatch<R, NW>) + raw_bytes);
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto fetch = [&](int64_t first) {   // lane 0: records [first, first + 32) of the stream
        const unsigned bytes = (unsigned)(min((int64_t)32, end - first) * v.rec_stride * (int)sizeof(T));
        mbar_expect_tx(bar, bytes);
        bulk_g2s(raw0, recs + (size_t)first * v.rec_stride, bytes, bar);
    };
    if (lane == 0) fetch(start);

    CellAcc<R> a;
    a.acc = reinterpret_cast<R *>(mine + sizeof(CellBatch<R, NW>) + raw_bytes + 16);
    a.s1 = (1 << v.cx_log) + v.halo;
    a.s2 = a.s1 * ((1 << v.cy_log) + v.halo);
    a.plane = plane;
    a.rows = plane / a.s1;
    a.cell = -1;
    a.cx0 = a.cy0 = a.cz0 = 0;
    a.dirty = false;
    if (ACC) {
        for (int i = lane; i < NW * plane; i += 32) a.acc[i] = R(0);
        __syncwarp();
    }

    const R rad2 = (R)(v.radius * v.radius);
    const int cmx = (1 << v.cx_log) - 1, cmy = (1 << v.cy_log) - 1, cmz = (1 << v.cz_log) - 1;

    int batch = 0;
    for (int64_t base = start; base < end; base += 32, batch++) {
        // ---- this batch has landed; every lane derives one particle ----------------------------
        mbar_wait(bar, (unsigned)batch & 1u);
        if (!ACC) {
            // ---- every box of the call holds <= kTinyVolume pixels: each lane walks the few pixels of ITS
            // particle straight from registers into the raster -- no staging slots, no shuffles (the sub-group
            // scheme below, which this instantiation used before, spent 31 warp instructions per particle on
            // re-reading the slots; this is ~10)
            if (base + lane < end) {
                const Rec<T> pr = smem_rec(raw0, v.rec_stride, lane);
                const double hp = (double)pr.h;
                const double xp = (double)pr.x, yp = (double)pr.y;
                const double rad = v.radius * hp;
                const double hinv = 1.0 / hp, hinv2 = hinv * hinv;
                int x0, x1, y0, y1, z0 = 0, z1 = 1;
                pixel_range(xp, rad, v.x_min, v.inv_pwx, v.nx, x0, x1);
                pixel_range(yp, rad, v.y_min, v.inv_pwy, v.ny, y0, y1);
                const R sx = (R)(v.pwx * hinv), sy = (R)(v.pwy * hinv);
                const R ox = (R)(-((xp - v.x_min) * v.inv_pwx - 0.5 - x0) * (v.pwx * hinv));
                const R oy = (R)(-((yp - v.y_min) * v.inv_pwy - 0.5 - y0) * (v.pwy * hinv));
                R sz = R(0), oz = R(0);
                double cc = 0.0;
                if (DIM3) {
                    const double zp = (double)pr.z;
                    pixel_range(zp, rad, v.z_min, v.inv_pwz, v.nz, z0, z1);
                    sz = (R)(v.pwz * hinv);
                    oz = (R)(-((zp - v.z_min) * v.inv_pwz - 0.5 - z0) * (v.pwz * hinv));
                } else if (v.use_z) {
                    const double dz = v.z_slice - (double)pr.z;
                    cc = dz * dz * hinv2;
                }
                const double scale = v.norm * (v.ndim == 2 ? hinv2 : hinv2 * hinv);
                R term[NW];
#pragma unroll
                for (int k = 0; k < NW; k++) term[k] = (R)((double)pr.w[k] * scale);
                const R cq = (R)cc + sqrt_guard(R(0));
                const int wx = x1 - x0, wy = y1 - y0, wz = z1 - z0;
                // a box beyond the tiny volume can only come with a stale plan (sphb_scatter_planned): dropped
                if (wx > 0 && wy > 0 && wz > 0 && wx * wy * wz <= kTinyVolume) {
                    for (int iz = 0; iz < wz; iz++) {
                        const R dzs = fma((R)iz, sz, oz);
                        const R qz = fma(dzs, dzs, cq);
                        for (int iy = 0; iy < wy; iy++) {
                            const R dys = fma((R)iy, sy, oy);
                            const R qyz = fma(dys, dys, qz);
                            const size_t row = ((size_t)(z0 + iz) * v.ny + (size_t)(y0 + iy)) * v.nx + (size_t)x0;
                            for (int ix = 0; ix < wx; ix++) {
                                const R dxs = fma((R)ix, sx, ox);
                                const R q2 = fma(dxs, dxs, qyz);
                                if (q2 < rad2) {
                                    const R wv = kernel_eval_pos<KIND, R>(q2, lv);
                                    SPHB_ASSERT(row + ix < v.n_out);
#pragma unroll
                                    for (int k = 0; k < NW; k++) red_add(outs[k] + row + ix, (double)(term[k] * wv));
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && base + 32 < end) fetch(base + 32); // the raw buffer is free again
            continue;
        }
        int vol = 0, my_cell = -1, my_class = 0;
        {
            int4 geo = make_int4(0, 0, -1, 0);
            if (base + lane < end) {
                const Rec<T> pr = smem_rec(raw0, v.rec_stride, lane);
                const double hp = (double)pr.h;
                const double xp = (double)pr.x, yp = (double)pr.y;
                const double rad = v.radius * hp;
                const double hinv = 1.0 / hp, hinv2 = hinv * hinv;
                int x0, x1, y0, y1, z0 = 0, z1 = 1;
                pixel_range(xp, rad, v.x_min, v.inv_pwx, v.nx, x0, x1);
                pixel_range(yp, rad, v.y_min, v.inv_pwy, v.ny, y0, y1);
                const double sx = v.pwx * hinv, sy = v.pwy * hinv;
                R2 pair;
                pair.x = (R)sx;
                pair.y = (R)(-((xp - v.x_min) * v.inv_pwx - 0.5 - x0) * sx);
                st.xs[lane] = pair;
                pair.x = (R)sy;
                pair.y = (R)(-((yp - v.y_min) * v.inv_pwy - 0.5 - y0) * sy);
                st.ys[lane] = pair;
                double cc = 0.0;
                if (DIM3) {
                    const double zp = (double)pr.z;
                    pixel_range(zp, rad, v.z_min, v.inv_pwz, v.nz, z0, z1);
                    const double sz = v.pwz * hinv;
                    pair.x = (R)sz;
                    pair.y = (R)(-((zp - v.z_min) * v.inv_pwz - 0.5 - z0) * sz);
                    st.zs[lane] = pair;
                } else if (v.use_z) {
                    const double dz = v.z_slice - (double)pr.z;
                    cc = dz * dz * hinv2;
                }
                const double scale = v.norm * (v.ndim == 2 ? hinv2 : hinv2 * hinv);
#pragma unroll
                for (int k = 0; k < NW; k++) st.term[k][lane] = (R)((double)pr.w[k] * scale);
                pair.x = (R)cc + sqrt_guard(R(0));
                pair.y = (R)((double)pr.w[0] * scale);
                st.ct[lane] = pair;
                st.bx[lane] = x0;
                st.by[lane] = y0;
                st.bz[lane] = z0;
                const int wx = x1 - x0, wy = y1 - y0, wz = z1 - z0;
                // a box wider than the halo can only come with a stale plan (sphb_scatter_planned): dropped
                if (wx > 0 && wy > 0 && wz > 0 && max(max(wx, wy), wz) <= v.halo + 1) {
                    vol = wx * wy * wz;
                    geo.x = (x0 & cmx) | ((y0 & cmy) << 8) | ((z0 & cmz) << 16);
                    geo.y = wx | (wy << 8) | (wz << 16);
                    geo.z = (((z0 >> v.cz_log) * v.ncy) + (y0 >> v.cy_log)) * v.ncx + (x0 >> v.cx_log);
                    const int lpl_log = DIM3 ? min(5, 32 - __clz(wx * wy - 1)) : 5; // ceil(log2(wx wy))
                    geo.w = lpl_log | (int)((65536u / (unsigned)wx + 1u) << 8);
                    const int s1 = (1 << v.cx_log) + v.halo, s2 = s1 * ((1 << v.cy_log) + v.halo);
                    const int nit = (wz + (32 >> lpl_log) - 1) >> (5 - lpl_log); // z-passes of a lane
                    st.use[lane] = make_int4((z0 & cmz) * s2 + (y0 & cmy) * s1 + (x0 & cmx), (wx * wy) | (nit << 8),
                                             (int)(65536u / (unsigned)(wx * wy) + 1u), 0);
                    my_cell = geo.z;
                    // class U = 1..8: a voxel box whose plane takes the whole warp in one round (one z-slice per
                    // pass, U = wz slices per lane); 0: everything else (images, planes of <= 16 or > 32 pixels)
                    my_class = DIM3 && lpl_log == 5 && wx * wy <= 32 ? nit : 0;
                }
            }
            st.geo[lane] = geo;
        }
        int vmax = vol;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        __syncwarp();
        if (lane == 0 && base + 32 < end) fetch(base + 32); // the raw buffer is free again: next batch on its way
        if (vmax == 0) continue;

        if (!ACC || vmax <= kTinyVolume) {
            // ---- tiny boxes: a sub-group of lanes per particle, straight into the raster -----
            int lpp_log = 0;
            while ((1 << lpp_log) < vmax) lpp_log++;
            const int sub = lane & ((1 << lpp_log) - 1), grp = lane >> lpp_log;
            for (int g = 0; g < 32; g += 32 >> lpp_log) {
                const int s = g + grp;
                const int4 geo = st.geo[s];
                const int dims = geo.y;
                const int wx = dims & 0xff, wy = (dims >> 8) & 0xff, wz = dims >> 16;
                const int wxy = wx * wy;
                if (sub < wxy * wz) {
                    // sub < 16: exact small divisions by multiply-shift (reciprocals from the staging step)
                    const int iz = (int)(((unsigned)sub * (unsigned)st.use[s].z) >> 16), r = sub - iz * wxy;
                    const int iy = (int)(((unsigned)r * ((unsigned)geo.w >> 8)) >> 16), ix = r - iy * wx;
                    const R2 xs = st.xs[s], ys = st.ys[s];
                    const R dxs = fma((R)ix, xs.x, xs.y), dys = fma((R)iy, ys.x, ys.y);
                    R q2 = fma(dys, dys, fma(dxs, dxs, st.ct[s].x));
                    if (DIM3) {
                        const R2 zz = st.zs[s];
                        const R dzs = fma((R)iz, zz.x, zz.y);
                        q2 = fma(dzs, dzs, q2);
                    }
                    if (q2 < rad2) {
                        const R wv = kernel_eval_pos<KIND, R>(q2, lv);
                        const size_t o = ((size_t)(st.bz[s] + iz) * v.ny + (size_t)(st.by[s] + iy)) * v.nx +
                                         (size_t)(st.bx[s] + ix);
                        SPHB_ASSERT(o < v.n_out);
#pragma unroll
                        for (int k = 0; k < NW; k++) red_add(outs[k] + o, (double)(st.term[k][s] * wv));
                    }
                }
            }
            __syncwarp();
            continue;
        }

        // ---- one particle per pass, lanes over its box, local accumulator --------------
        // The batch is taken apart into runs of one cell (the stream is cell-ordered: usually one or two per
        // batch) and, inside a run, into classes of equal z-extent: each class is a tight loop over its
        // particles with the z-column unrolled at compile time -- no per-particle dispatch (the indirect
        // branch, the cell test and the loop control of the particle-by-particle version took a fifth of the
        // kernel's stall samples at 4 warps per scheduler).
        unsigned todo = ACC ? __ballot_sync(0xffffffffu, vol > 0) : 0u;
        while (todo) {
            const int first = __ffs(todo) - 1;
            const int c = __shfl_sync(0xffffffffu, my_cell, first);
            const unsigned run = todo & __ballot_sync(0xffffffffu, my_cell == c);
            todo &= ~run;
            if (c != a.cell) {
                cell_flush<R, NW>(a, v, lane, outs);
                const int gx = st.geo[first].x;
                a.cell = c;
                a.cx0 = st.bx[first] - (gx & 0xff);
                a.cy0 = st.by[first] - ((gx >> 8) & 0xff);
                a.cz0 = st.bz[first] - (gx >> 16);
            }
            a.dirty = true;
            if (DIM3) {
#define SPHB_CLASS(U) \
    cell_class<R, KIND, NW, U>(run & __ballot_sync(0xffffffffu, my_class == U), st, a, lane, rad2, plane, lv)
                SPHB_CLASS(5);
                SPHB_CLASS(4);
                SPHB_CLASS(6);
                SPHB_CLASS(3);
                SPHB_CLASS(7);
                SPHB_CLASS(2);
                SPHB_CLASS(8);
                SPHB_CLASS(1);
#undef SPHB_CLASS
            }
            unsigned rest = run & __ballot_sync(0xffffffffu, my_class == 0);
            while (rest) {
                const int s = __ffs(rest) - 1;
                rest &= rest - 1;
                const int4 geo = st.geo[s];
                const int4 use = st.use[s];
                const int wx = geo.y & 0xff, wz = geo.y >> 16;
                const int wxy = use.y & 0xff, nit = use.y >> 8;
                SPHB_ASSERT((geo.x & 0xff) + wx <= a.s1 && ((geo.x >> 16) + wz) * a.s2 <= a.plane);
                const R2 xs = st.xs[s], ys = st.ys[s], ct = st.ct[s];
                const R sx = xs.x, ox = xs.y, sy = ys.x, oy = ys.y, cq = ct.x;
                R term[NW];
                term[0] = ct.y;
#pragma unroll
                for (int k = 1; k < NW; k++) term[k] = st.term[k][s];
                const unsigned inv_wx = (unsigned)geo.w >> 8;
                if (DIM3) {
                    // lanes: in-plane pixel lane % lpl, z-slice lane / lpl (lpl = power of two >= wx wy, at
                    // most 32; a wider plane is covered in several rounds)
                    const R2 zz = st.zs[s];
                    const R sz = zz.x, oz = zz.y;
                    const int lpl_log = geo.w & 0xff, lpl = 1 << lpl_log, sp = 32 >> lpl_log, zs = lane >> lpl_log;
                    const R dz0 = fma((R)zs, sz, oz), step = (R)sp * sz;
                    const int zstride = sp * a.s2;
                    for (int a0 = lane & (lpl - 1); a0 < wxy; a0 += lpl) {
                        const int iy = (int)(((unsigned)a0 * inv_wx) >> 16), ix = a0 - iy * wx;
                        const R dxs = fma((R)ix, sx, ox), dys = fma((R)iy, sy, oy);
                        const R qxy = fma(dys, dys, fma(dxs, dxs, cq)); // cq carries sqrt_guard()
                        R *ap = a.acc + (use.x + zs * a.s2 + iy * a.s1 + ix);
#define SPHB_COLUMN(U) cell_column<R, KIND, NW, U, true>(qxy, dz0, step, zs, sp, wz, rad2, term, ap, zstride, plane, lv)
                        switch (nit) {
                        case 1: SPHB_COLUMN(1); break;
                        case 2: SPHB_COLUMN(2); break;
                        case 3: SPHB_COLUMN(3); break;
                        case 4: SPHB_COLUMN(4); break;
                        case 5: SPHB_COLUMN(5); break;
                        case 6: SPHB_COLUMN(6); break;
                        case 7: SPHB_COLUMN(7); break;
                        default: SPHB_COLUMN(8); break;
                        }
#undef SPHB_COLUMN
                    }
                } else {
                    for (int a0 = lane; a0 < wxy; a0 += 32) {
                        const int iy = (int)(((unsigned)a0 * inv_wx) >> 16), ix = a0 - iy * wx;
                        const R dxs = fma((R)ix, sx, ox), dys = fma((R)iy, sy, oy);
                        const R qxy = fma(dys, dys, fma(dxs, dxs, cq));
                        if (qxy < rad2) {
                            const R wv = kernel_eval_pos<KIND, R>(qxy, lv);
                            R *ap = a.acc + (use.x + iy * a.s1 + ix);
#pragma unroll
                            for (int k = 0; k < NW; k++) ap[k * plane] = fma(term[k], wv, ap[k * plane]);
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    if (ACC) cell_flush<R, NW>(a, v, lane, outs);
}

// ---------------------------------------------------------------------------
template <typename T, typename R, int KIND, int NW, bool DIM3>
static int launch_one(const View &v, const T *recs, const sphb_kernel *k, const void *lut_global,
                      const uint32_t *cell_start, int64_t n_small, int64_t cell_lo, int64_t cell_hi,
                      double *const *out, cudaStream_t stream)
{
    using R2 = typename Real2<R>::type;
    const int s1 = (1 << v.cx_log) + v.halo, s2 = s1 * ((1 << v.cy_log) + v.halo);
    const int edge = v.halo + 1;
    const bool acc = (DIM3 ? edge * edge * edge : edge * edge) > kTinyVolume;
    const int plane = !acc ? 0 : DIM3 ? s2 * ((1 << v.cz_log) + v.halo) : s2;
    size_t lut_bytes = 0;
    if (KIND == SPHB_COLUMN_LUT) lut_bytes = align_up(sizeof(R2) * (size_t)k->lut_samples, 16);
    // staging slots, one raw record batch, its mbarrier, accumulator
    const size_t warp_bytes = align_up(sizeof(CellBatch<R, NW>) + 32 * (size_t)v.rec_stride * sizeof(T) + 16 +
                                           sizeof(R) * (size_t)plane * NW, 16);
    const size_t limit = 227 * 1024;
    if (lut_bytes + warp_bytes > limit) {
        set_error("cell accumulator (%zu bytes) and column table (%zu bytes) do not fit in shared memory",
                  warp_bytes, lut_bytes);
        return SPHB_ERR_ARGUMENT;
    }
    int warps = (int)((limit - lut_bytes) / warp_bytes);
    const int cap = acc ? kCellMaxWarps : 8;
    warps = warps > cap ? cap : warps;
    const int64_t units = (n_small + kCellChunk - 1) / kCellChunk + 1; // upper bound for any cell range
    const int grid = div_up(units, warps);
    double *o[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < NW; i++) o[i] = out[i];
    const size_t smem = lut_bytes + warp_bytes * warps;
    if (acc) {
        auto kern = cell_kernel<T, R, KIND, NW, DIM3, true>;
        SPHB_CUDA(allow_large_smem(kern));
        kern<<<grid, warps * 32, smem, stream>>>(v, recs, k->lut, k->lut_samples, lut_global, cell_start, cell_lo,
                                                 cell_hi, plane, (int)lut_bytes, (int)warp_bytes, o[0], o[1], o[2],
                                                 o[3]);
    } else {
        auto kern = cell_kernel<T, R, KIND, NW, DIM3, false>;
        SPHB_CUDA(allow_large_smem(kern));
        kern<<<grid, warps * 32, smem, stream>>>(v, recs, k->lut, k->lut_samples, lut_global, cell_start, cell_lo,
                                                 cell_hi, plane, (int)lut_bytes, (int)warp_bytes, o[0], o[1], o[2],
                                                 o[3]);
    }
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}

template <typename T, typename R, int KIND, bool DIM3, typename... A> static int cells_nw(int nw, A &&...a)
{
    switch (nw) {
    case 1: return launch_one<T, R, KIND, 1, DIM3>(a...);
    case 2: return launch_one<T, R, KIND, 2, DIM3>(a...);
    case 3: return launch_one<T, R, KIND, 3, DIM3>(a...);
    default: return launch_one<T, R, KIND, 4, DIM3>(a...);
    }
}

template <typename T, typename R, bool DIM3, typename... A> static int cells_kind(int kind, int nw, A &&...a)
{
    switch (kind) {
    case SPHB_CUBIC: return cells_nw<T, R, SPHB_CUBIC, DIM3>(nw, a...);
    case SPHB_QUARTIC: return cells_nw<T, R, SPHB_QUARTIC, DIM3>(nw, a...);
    case SPHB_QUINTIC: return cells_nw<T, R, SPHB_QUINTIC, DIM3>(nw, a...);
    case SPHB_COLUMN_LUT: return cells_nw<T, R, SPHB_COLUMN_LUT, DIM3>(nw, a...);
    default: return cells_nw<T, R, SPHB_COLUMN_LUT_GLOBAL, DIM3>(nw, a...);
    }
}

int launch_cells(const View &v, bool f32, int kind, int nw, const void *recs_small, const sphb_kernel *k,
                 const void *lut_global, const uint32_t *cell_start, int64_t n_small, int64_t cell_lo,
                 int64_t cell_hi, double *const *out, cudaStream_t stream)
{
    if (v.dim3)
        return f32 ? cells_kind<float, float, true>(kind, nw, v, static_cast<const float *>(recs_small), k, lut_global,
                                                    cell_start, n_small, cell_lo, cell_hi, out, stream)
                   : cells_kind<double, double, true>(kind, nw, v, static_cast<const double *>(recs_small), k,
                                                      lut_global, cell_start, n_small, cell_lo, cell_hi, out, stream);
    return f32 ? cells_kind<float, float, false>(kind, nw, v, static_cast<const float *>(recs_small), k, lut_global,
                                                 cell_start, n_small, cell_lo, cell_hi, out, stream)
               : cells_kind<double, double, false>(kind, nw, v, static_cast<const double *>(recs_small), k, lut_global,
                                                   cell_start, n_small, cell_lo, cell_hi, out, stream);
}

} // namespace sphb
