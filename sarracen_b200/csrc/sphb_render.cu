// Tile kernel of the sampled scatter: large footprints (sm_100a).
//
// The sorted (tile, particle) pair list is cut into fixed chunks, one CTA per chunk.  A CTA walks
// the tile segments of its chunk; per tile it stages the particles' derived records into shared
// memory, and every warp owns an exclusive strip (2-D: rows, 3-D: a z-plane) of the tile, so
// accumulation needs no atomics: wide footprints go to per-lane register accumulators on fixed
// pixels, narrow ones to a plain load-add-store on the shared-memory tile.  A finished tile
// segment is added to the float64 output with red.global.add.f64 -- the only global atomics of
// this path, once per tile segment, coalesced along x.  Replaces the pixel loop of
// CPUBackend._fast_2d (sarracen/interpolate/cpu_backend.py:263-308) for footprints larger than a
// few pixels.
#include "sphb_scatter.cuh"

namespace sphb {

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
              const uint32_t *__restrict__ range, double *o0, double *o1, double *o2, double *o3)
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

    // This launch covers the sorted pairs [range[0], range[1]) (a z-slab stage of a voxel grid
    // rendered slab by slab for the multi-GPU reduction), or all of them.
    const int64_t range_lo = range ? (int64_t)range[0] : 0, range_hi = range ? (int64_t)range[1] : n_pairs;
    const int64_t chunk_lo = range_lo + (int64_t)blockIdx.x * kChunkPairs;
    if (chunk_lo >= range_hi) return;
    const int64_t chunk_hi = min(chunk_lo + (int64_t)kChunkPairs, range_hi);

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

    int64_t pos = chunk_lo;

    while (pos < chunk_hi) {
        const uint32_t tile = keys[pos];
        if (tile >= (uint32_t)(v.ntx * v.nty * v.ntz)) break; // 0xffffffff padding of a planned call: nothing follows
        SPHB_ASSERT((int64_t)tile_end[tile] > pos);
        const int64_t seg_hi = min((int64_t)tile_end[tile], chunk_hi);
        const int ttx = tile % v.ntx, tty = (tile / v.ntx) % v.nty, ttz = tile / (v.ntx * v.nty);
        const int ox = ttx * TX, oy = tty * TY, oz = ttz * TZ; // tile origin in pixels
        const int lim_x = min(TX, v.nx - ox), lim_y = min(TY, v.ny - oy), lim_z = min(TZ, v.nz - oz);

        for (int64_t base = pos; base < seg_hi; base += kBatch) {
            const int nb = (int)min((int64_t)kBatch, seg_hi - base);
            // ---- stage: one particle per thread --------------------------------
            if (tid < nb) {
                const int64_t p = vals[base + tid];
                const Rec<T> pr = load_rec(recs, v.rec_stride, p, NW);
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


template <typename T, typename R, int KIND, int NW, bool DIM3, int TX, int TY, int TZ, int WARPS>
static int launch_one(const View &v, const T *recs, const sphb_kernel *k, const void *lut_global,
                      const uint32_t *keys, const uint32_t *vals, const uint32_t *tile_end, int64_t n_pairs,
                      const uint32_t *range, double *const *out, cudaStream_t stream)
{
    using R2 = typename Real2<R>::type;
    auto kern = render_kernel<T, R, KIND, NW, DIM3, TX, TY, TZ, WARPS>;
    const size_t fixed = sizeof(R) * (size_t)(TX + RowPad<DIM3>::value) * TY * TZ * NW +
                         sizeof(Staged<R, NW, RenderBatch<R, DIM3, WARPS>::value, DIM3>);
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
                                             tile_end, n_pairs, range, o[0], o[1], o[2], o[3]);
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}

template <typename T, typename R, int KIND, bool DIM3, int TX, int TY, int TZ, int WARPS, typename... A>
static int dispatch_nw(int nw, A &&...a)
{
    switch (nw) {
    case 1: return launch_one<T, R, KIND, 1, DIM3, TX, TY, TZ, WARPS>(a...);
    case 2: return launch_one<T, R, KIND, 2, DIM3, TX, TY, TZ, WARPS>(a...);
    // three and four weights in double: half-height 2-D tiles (see tile2y())
    case 3: return launch_one<T, R, KIND, 3, DIM3, TX, (!DIM3 && sizeof(R) == 8) ? TY / 2 : TY, TZ, WARPS>(a...);
    default: return launch_one<T, R, KIND, 4, DIM3, TX, (!DIM3 && sizeof(R) == 8) ? TY / 2 : TY, TZ, WARPS>(a...);
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

int launch_render(const View &v, bool f32, int kind, int nw, const void *recs, const sphb_kernel *k,
                  const void *lut_global, const uint32_t *keys, const uint32_t *vals, const uint32_t *tile_end,
                  int64_t n_pairs, const uint32_t *range, double *const *out, cudaStream_t stream)
{
    if (v.dim3) {
        if (f32)
            return dispatch_kind<float, float, true, kTile3X, kTile3Y, kTile3Z, 8>(
                kind, nw, v, static_cast<const float *>(recs), k, lut_global, keys, vals, tile_end, n_pairs, range,
                out, stream);
        return dispatch_kind<double, double, true, kTile3X, kTile3Y, kTile3Z, 8>(
            kind, nw, v, static_cast<const double *>(recs), k, lut_global, keys, vals, tile_end, n_pairs, range, out,
            stream);
    }
    if (f32)
        return dispatch_kind<float, float, false, kTile2X, kTile2Y, 1, 8>(
            kind, nw, v, static_cast<const float *>(recs), k, lut_global, keys, vals, tile_end, n_pairs, range, out,
            stream);
    return dispatch_kind<double, double, false, kTile2X, kTile2Y, 1, 8>(
        kind, nw, v, static_cast<const double *>(recs), k, lut_global, keys, vals, tile_end, n_pairs, range, out,
        stream);
}

} // namespace sphb
