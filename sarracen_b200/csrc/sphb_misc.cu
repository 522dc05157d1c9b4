// Error reporting, the column-integrated kernel table (A4), line cuts (B6, B7)
// and the on-device normalisation (N5) of libsphb200.
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstdarg>
#include <cstdio>

#include "sphb_common.cuh"

namespace sphb {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------------------
// A4: out[i] = 2 * trapezoid_{q_z in linspace(0, sqrt(R^2 - q_xy^2), S)}
//              W_3D( sqrt(q_xy^2 + q_z^2) ),  q_xy = R i / (S - 1)
// (sarracen/kernels/base_kernel.py:39-66, :102-117).  One thread per entry,
// the S-point sum runs sequentially in double.
template <int KIND> __device__ double w3d(double q, double radius)
{
    return kernel_norm(KIND, 3) * kernel_eval_any<KIND>(q, radius, nullptr, 0);
}

template <int KIND> __global__ void column_kernel_kernel(int samples, double *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= samples) return;
    const double R = kernel_default_radius(KIND);
    const double qxy = R * i / (samples - 1);
    const double bound = sqrt(R * R - qxy * qxy);
    const double step = bound / (samples - 1);
    double acc = 0.0, zprev = 0.0, yprev = w3d<KIND>(sqrt(qxy * qxy), R);
    for (int k = 1; k < samples; k++) {
        const double z = (k == samples - 1) ? bound : k * step;
        const double y = w3d<KIND>(sqrt(qxy * qxy + z * z), R);
        acc += (z - zprev) * (y + yprev) / 2.0;
        zprev = z;
        yprev = y;
    }
    out[i] = 2.0 * acc;
}

// ---------------------------------------------------------------------------
// Line cuts.  One warp takes 32 particles at a time; each lane works out its
// particle's pixel span, then the warp walks the non-empty spans with its
// lanes on consecutive pixels.  Each warp owns a private copy of the line in
// shared memory (plain adds), flushed once with red.global.add.f64; lines too
// long for that are accumulated with global reductions directly.
constexpr int kLineWarps = 8;
constexpr int kLineMaxSmemPixels = 1024;

struct LineGeom {
    int pixels;
    double x1, x2, y1, y2, z1, z2;
    // 2-D: slope / intercept form
    double gradient, yint, pixwidth, xpixwidth, aa;
    // 3-D: unit direction and length
    double ux, uy, uz, length;
    double radius, norm;
    int lut_samples;
};

template <typename T, int KIND, int NW, bool DIM3>
__global__ void __launch_bounds__(kLineWarps * 32)
line_kernel(LineGeom g, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z,
            const T *__restrict__ h, const T *__restrict__ w0, const T *__restrict__ w1,
            const T *__restrict__ w2, const T *__restrict__ w3, const double *__restrict__ lut, int64_t n,
            double *o0, double *o1, double *o2, double *o3, int use_smem)
{
    extern __shared__ double line_s[]; // [warp][NW][pixels]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T *wsrc[4] = {w0, w1, w2, w3};
    double *outs[4] = {o0, o1, o2, o3};
    double *mine = line_s + (size_t)warp * NW * g.pixels;
    if (use_smem) {
        for (int i = lane; i < NW * g.pixels; i += 32) mine[i] = 0.0;
        __syncwarp();
    }
    const int64_t warps_total = (int64_t)gridDim.x * kLineWarps;
    for (int64_t base = ((int64_t)blockIdx.x * kLineWarps + warp) * 32; base < n; base += warps_total * 32) {
        const int64_t p = base + lane;
        int i0 = 0, i1 = 0;
        double xp = 0, yp = 0, zp = 0, hp = 1, term[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) term[k] = 0.0;
        if (p < n) {
            xp = (double)__ldg(x + p);
            yp = (double)__ldg(y + p);
            hp = (double)__ldg(h + p);
            const double rad = g.radius * hp;
            if (DIM3) {
                // cpu_backend.py:573-585
                zp = (double)__ldg(z + p);
                const double dx = g.x1 - xp, dy = g.y1 - yp, dz = g.z1 - zp;
                const double proj = g.ux * dx + g.uy * dy + g.uz * dz;
                const double delta = proj * proj - (dx * dx + dy * dy + dz * dz) + rad * rad;
                if (delta >= 0.0 && hp > 0.0) {
                    const double sq = sqrt(delta);
                    const double d1 = -proj - sq, d2 = -proj + sq;
                    i0 = (int)fmin(fmax(rint(d1 / g.length * g.pixels), 0.0), (double)g.pixels);
                    i1 = (int)fmin(fmax(rint(d2 / g.length * g.pixels), 0.0), (double)g.pixels);
                }
            } else {
                // cpu_backend.py:480-506
                const double bb = 2.0 * g.gradient * (g.yint - yp) - 2.0 * xp;
                const double cc = xp * xp + yp * yp - 2.0 * g.yint * yp + g.yint * g.yint - rad * rad;
                double det = bb * bb - 4.0 * g.aa * cc;
                if (det >= 0.0 && hp > 0.0) {
                    det = sqrt(det);
                    double rs = (-bb - det) / (2.0 * g.aa), re = (-bb + det) / (2.0 * g.aa);
                    rs = fmin(fmax(rs, g.x1), g.x2);
                    re = fmin(fmax(re, g.x1), g.x2);
                    const double ys = g.gradient * rs + g.yint - g.y1, ye = g.gradient * re + g.yint - g.y1;
                    const double ds = sqrt((rs - g.x1) * (rs - g.x1) + ys * ys);
                    const double de = sqrt((re - g.x1) * (re - g.x1) + ye * ye);
                    i0 = (int)fmin(fmax(rint(ds / g.pixwidth), 0.0), (double)g.pixels);
                    i1 = (int)fmin(fmax(rint(de / g.pixwidth), 0.0), (double)g.pixels);
                }
            }
            if (i1 > i0) {
                const double scale = g.norm / (DIM3 ? hp * hp * hp : hp * hp);
#pragma unroll
                for (int k = 0; k < NW; k++) term[k] = (double)__ldg(wsrc[k] + p) * scale;
            }
        }
        unsigned mask = __ballot_sync(0xffffffffu, i1 > i0);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const int a = __shfl_sync(0xffffffffu, i0, src), b = __shfl_sync(0xffffffffu, i1, src);
            const double sx = __shfl_sync(0xffffffffu, xp, src), sy = __shfl_sync(0xffffffffu, yp, src);
            const double sz = __shfl_sync(0xffffffffu, zp, src), sh = __shfl_sync(0xffffffffu, hp, src);
            double st[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) st[k] = __shfl_sync(0xffffffffu, term[k], src);
            for (int i = a + lane; i < b; i += 32) {
                double q;
                if (DIM3) {
                    const double xd = g.x1 + (i + 0.5) * (g.x2 - g.x1) / g.pixels - sx;
                    const double yd = g.y1 + (i + 0.5) * (g.y2 - g.y1) / g.pixels - sy;
                    const double zd = g.z1 + (i + 0.5) * (g.z2 - g.z1) / g.pixels - sz;
                    q = sqrt((xd * xd + yd * yd + zd * zd) / (sh * sh));
                } else {
                    const double xq = g.x1 + (i + 0.5) * g.xpixwidth;
                    const double yq = g.gradient * xq + g.yint;
                    const double dx = xq - sx, dy = yq - sy;
                    q = sqrt((dx * dx + dy * dy) / (sh * sh));
                }
                const double wv = kernel_eval_any<KIND>(q, g.radius, lut, g.lut_samples);
                SPHB_ASSERT(i >= 0 && i < g.pixels);
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    if (use_smem)
                        mine[k * g.pixels + i] += st[k] * wv;
                    else
                        red_add(outs[k] + i, st[k] * wv);
                }
            }
            __syncwarp();
        }
    }
    if (use_smem) {
        __syncwarp();
        for (int i = lane; i < NW * g.pixels; i += 32) {
            const double val = mine[i];
            if (val != 0.0) red_add(outs[i / g.pixels] + (i % g.pixels), val);
        }
    }
}

template <typename T, int KIND, int NW, bool DIM3>
static int launch_line(const LineGeom &g, const sphb_particles *p, const sphb_kernel *k, double *const *out,
                       cudaStream_t stream)
{
    auto kern = line_kernel<T, KIND, NW, DIM3>;
    // per-warp private lines in shared memory when they fit in the 227 KB a CTA may have (up to 4
    // weights x 909 pixels); longer lines reduce into global memory directly
    const size_t want = sizeof(double) * kLineWarps * NW * (size_t)g.pixels;
    const int use_smem = g.pixels <= kLineMaxSmemPixels && want <= 227 * 1024;
    const size_t smem = use_smem ? want : 0;
    if (smem > 48 * 1024)
        SPHB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0;
    SPHB_CUDA(sm_count(&sms));
    int grid = div_up(p->n, kLineWarps * 32 * 4);
    grid = grid < 1 ? 1 : grid > sms * 4 ? sms * 4 : grid;
    const T *w[4] = {nullptr, nullptr, nullptr, nullptr};
    double *o[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < NW; i++) {
        w[i] = static_cast<const T *>(p->w[i]);
        o[i] = out[i];
    }
    kern<<<grid, kLineWarps * 32, smem, stream>>>(
        g, static_cast<const T *>(p->x), static_cast<const T *>(p->y), static_cast<const T *>(p->z),
        static_cast<const T *>(p->h), w[0], w[1], w[2], w[3], k->lut, p->n, o[0], o[1], o[2], o[3], use_smem);
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}

template <typename T, int KIND, bool DIM3, typename... A> static int line_nw(int nw, A &&...a)
{
    switch (nw) {
    case 1: return launch_line<T, KIND, 1, DIM3>(a...);
    case 2: return launch_line<T, KIND, 2, DIM3>(a...);
    case 3: return launch_line<T, KIND, 3, DIM3>(a...);
    default: return launch_line<T, KIND, 4, DIM3>(a...);
    }
}

template <typename T, bool DIM3, typename... A> static int line_kind(int kind, int nw, A &&...a)
{
    switch (kind) {
    case SPHB_CUBIC: return line_nw<T, SPHB_CUBIC, DIM3>(nw, a...);
    case SPHB_QUARTIC: return line_nw<T, SPHB_QUARTIC, DIM3>(nw, a...);
    case SPHB_QUINTIC: return line_nw<T, SPHB_QUINTIC, DIM3>(nw, a...);
    default: return line_nw<T, SPHB_COLUMN_LUT, DIM3>(nw, a...);
    }
}

static int line_impl(const sphb_particles *p, const sphb_kernel *k, bool dim3, int pixels, double x1,
                     double x2, double y1, double y2, double z1, double z2, int accumulate,
                     double *const *out, cudaStream_t stream)
{
    if (!p || !k || !out || pixels <= 0 || p->n < 0 || p->n_weights < 1 || p->n_weights > SPHB_MAX_WEIGHTS ||
        (p->dtype != SPHB_F32 && p->dtype != SPHB_F64) || k->kind < SPHB_CUBIC || k->kind > SPHB_COLUMN_LUT ||
        (k->kind == SPHB_COLUMN_LUT && (!k->lut || k->lut_samples < 2)) || !(k->radius > 0.0)) {
        set_error("invalid argument to line cut");
        return SPHB_ERR_ARGUMENT;
    }
    for (int i = 0; i < p->n_weights; i++)
        if (!out[i] || (p->n > 0 && !p->w[i])) {
            set_error("missing weight or output %d", i);
            return SPHB_ERR_ARGUMENT;
        }
    if (p->n > 0 && (!p->x || !p->y || !p->h || (dim3 && !p->z))) {
        set_error("missing particle column");
        return SPHB_ERR_ARGUMENT;
    }
    if (!accumulate)
        for (int i = 0; i < p->n_weights; i++)
            SPHB_CUDA(cudaMemsetAsync(out[i], 0, sizeof(double) * (size_t)pixels, stream));
    if (p->n == 0) return SPHB_OK;
    LineGeom g = {};
    g.pixels = pixels;
    g.x1 = x1; g.x2 = x2; g.y1 = y1; g.y2 = y2; g.z1 = z1; g.z2 = z2;
    g.radius = k->radius;
    g.lut_samples = k->lut_samples;
    g.norm = kernel_norm(k->kind, dim3 ? 3 : 2);
    if (dim3) {
        const double dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
        g.length = sqrt(dx * dx + dy * dy + dz * dz);
        g.ux = dx / g.length; g.uy = dy / g.length; g.uz = dz / g.length;
    } else {
        // cpu_backend.py:463-472
        g.gradient = (x2 - x1 != 0.0) ? (y2 - y1) / (x2 - x1) : 0.0;
        g.yint = y2 - g.gradient * x2;
        const double xlength = sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
        g.pixwidth = xlength / pixels;
        g.xpixwidth = (x2 - x1) / pixels;
        g.aa = 1.0 + g.gradient * g.gradient;
    }
    if (p->dtype == SPHB_F32)
        return dim3 ? line_kind<float, true>(k->kind, p->n_weights, g, p, k, out, stream)
                    : line_kind<float, false>(k->kind, p->n_weights, g, p, k, out, stream);
    return dim3 ? line_kind<double, true>(k->kind, p->n_weights, g, p, k, out, stream)
                : line_kind<double, false>(k->kind, p->n_weights, g, p, k, out, stream);
}

// ---------------------------------------------------------------------------
// Measurement aid: exact number of pixel centres with q <= R, summed over the
// particles.  One thread per particle walks the rows of its support box and
// counts the centres inside the row's chord.
struct CountView {
    int nx, ny, nz;
    double x_min, y_min, z_min, pwx, pwy, pwz, radius, z_slice;
    int use_z, dim3;
};

__device__ __forceinline__ void centre_range(double c, double rad, double origin, double pw, int n, int &lo,
                                             int &hi)
{
    lo = (int)fmin(fmax(ceil((c - rad - origin) / pw - 0.5), 0.0), (double)n);
    hi = (int)fmin(fmax(floor((c + rad - origin) / pw - 0.5) + 1.0, 0.0), (double)n);
}

template <typename T>
__global__ void count_contributions_kernel(CountView v, const T *x, const T *y, const T *z, const T *h,
                                           int64_t n, unsigned long long *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long c = 0;
    if (i < n) {
        const double hp = (double)h[i], xp = (double)x[i], yp = (double)y[i];
        const double rad = v.radius * hp;
        if (hp > 0.0 && isfinite(hp) && isfinite(xp) && isfinite(yp)) {
            double r2 = rad * rad;
            int k0 = 0, k1 = 1;
            double zp = 0.0;
            bool ok = true;
            if (v.dim3) {
                zp = (double)z[i];
                centre_range(zp, rad, v.z_min, v.pwz, v.nz, k0, k1);
            } else if (v.use_z) {
                const double dz = v.z_slice - (double)z[i];
                ok = fabs(dz) < rad;
                r2 -= dz * dz;
            }
            for (int k = k0; ok && k < k1; k++) {
                double rr = r2;
                if (v.dim3) {
                    const double dz = v.z_min + (k + 0.5) * v.pwz - zp;
                    rr -= dz * dz;
                }
                if (rr < 0.0) continue;
                int j0, j1;
                centre_range(yp, sqrt(rr), v.y_min, v.pwy, v.ny, j0, j1);
                for (int j = j0; j < j1; j++) {
                    const double dy = v.y_min + (j + 0.5) * v.pwy - yp;
                    const double rx = rr - dy * dy;
                    if (rx < 0.0) continue;
                    int i0, i1;
                    centre_range(xp, sqrt(rx), v.x_min, v.pwx, v.nx, i0, i1);
                    if (i1 > i0) c += (unsigned long long)(i1 - i0);
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// ---------------------------------------------------------------------------
// N5: num <- nan_to_num(num / den)   (interpolate.py:594)
__global__ void normalize_kernel(double *num, const double *den, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = num[i] / den[i];
    if (isnan(v))
        v = 0.0;
    else if (isinf(v))
        v = v > 0 ? DBL_MAX : -DBL_MAX;
    num[i] = v;
}

// ---------------------------------------------------------------------------
// Multi-GPU exchange, receive side: dst = own + sum over the slots s != skip of slots[s * stride + i].
// The slots are this rank's receive buffers in peer-mapped memory: the other ranks' copy engines have
// written their partial z-planes there (sarracen_b200/parallel.py), this kernel is the reduction of the
// reduce-scatter.  Two doubles per thread (16-byte accesses), grid-stride.
template <bool VEC>
__global__ void __launch_bounds__(256)
sum_slots_kernel(double *__restrict__ dst, const double *__restrict__ own, const double *__restrict__ slots,
                 int n_slots, int64_t stride, int skip, int64_t n)
{
    const int64_t step = (int64_t)gridDim.x * blockDim.x * 2;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += step) {
        if (VEC && i + 1 < n) {
            double2 a = *reinterpret_cast<const double2 *>(own + i);
            for (int s = 0; s < n_slots; s++) {
                if (s == skip) continue;
                const double2 b = *reinterpret_cast<const double2 *>(slots + (size_t)s * stride + i);
                a.x += b.x;
                a.y += b.y;
            }
            *reinterpret_cast<double2 *>(dst + i) = a;
        } else {
            for (int64_t j = i; j < min(i + 2, n); j++) {
                double a = own[j];
                for (int s = 0; s < n_slots; s++)
                    if (s != skip) a += slots[(size_t)s * stride + j];
                dst[j] = a;
            }
        }
    }
}

} // namespace sphb

using namespace sphb;

extern "C" int sphb_version(void) { return SPHB_VERSION; }

extern "C" const char *sphb_last_error(void) { return g_error; }

extern "C" int sphb_column_kernel(int kind, int samples, double *out, sphb_stream stream)
{
    if (samples < 2 || !out || kind < SPHB_CUBIC || kind > SPHB_QUINTIC) {
        set_error("column kernel needs an analytic kernel and >= 2 samples");
        return SPHB_ERR_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = div_up(samples, 64);
    if (kind == SPHB_CUBIC)
        column_kernel_kernel<SPHB_CUBIC><<<grid, 64, 0, s>>>(samples, out);
    else if (kind == SPHB_QUARTIC)
        column_kernel_kernel<SPHB_QUARTIC><<<grid, 64, 0, s>>>(samples, out);
    else
        column_kernel_kernel<SPHB_QUINTIC><<<grid, 64, 0, s>>>(samples, out);
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}

extern "C" int sphb_line2d(const sphb_particles *p, const sphb_kernel *k, int pixels, double x1, double x2,
                           double y1, double y2, int accumulate, double *const *out, sphb_stream stream)
{
    return line_impl(p, k, false, pixels, x1, x2, y1, y2, 0.0, 0.0, accumulate, out,
                     static_cast<cudaStream_t>(stream));
}

extern "C" int sphb_line3d(const sphb_particles *p, const sphb_kernel *k, int pixels, double x1, double x2,
                           double y1, double y2, double z1, double z2, int accumulate, double *const *out,
                           sphb_stream stream)
{
    return line_impl(p, k, true, pixels, x1, x2, y1, y2, z1, z2, accumulate, out,
                     static_cast<cudaStream_t>(stream));
}

extern "C" int sphb_count_contributions(const sphb_particles *p, double radius, const sphb_raster *r, int use_z,
                                        double z_slice, unsigned long long *out, sphb_stream stream)
{
    if (!p || !r || !out || p->n < 0 || (p->dtype != SPHB_F32 && p->dtype != SPHB_F64) || r->nx <= 0 ||
        r->ny <= 0 || r->nz <= 0 || !(radius > 0.0)) {
        set_error("invalid argument to count_contributions");
        return SPHB_ERR_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    SPHB_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned long long), s));
    if (p->n == 0) return SPHB_OK;
    CountView v;
    v.nx = r->nx; v.ny = r->ny; v.nz = r->nz;
    v.x_min = r->x_min; v.y_min = r->y_min; v.z_min = r->z_min;
    v.pwx = (r->x_max - r->x_min) / r->nx;
    v.pwy = (r->y_max - r->y_min) / r->ny;
    v.pwz = r->nz > 1 ? (r->z_max - r->z_min) / r->nz : 1.0;
    v.radius = radius;
    v.z_slice = z_slice;
    v.use_z = use_z;
    v.dim3 = r->nz > 1;
    if ((v.dim3 || use_z) && !p->z) {
        set_error("missing z column");
        return SPHB_ERR_ARGUMENT;
    }
    const int grid = div_up(p->n, 256);
    if (p->dtype == SPHB_F32)
        count_contributions_kernel<float><<<grid, 256, 0, s>>>(
            v, static_cast<const float *>(p->x), static_cast<const float *>(p->y),
            static_cast<const float *>(p->z), static_cast<const float *>(p->h), p->n, out);
    else
        count_contributions_kernel<double><<<grid, 256, 0, s>>>(
            v, static_cast<const double *>(p->x), static_cast<const double *>(p->y),
            static_cast<const double *>(p->z), static_cast<const double *>(p->h), p->n, out);
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}

extern "C" int sphb_normalize(double *num, const double *den, int64_t n, sphb_stream stream)
{
    if (!num || !den || n < 0) {
        set_error("invalid argument to normalize");
        return SPHB_ERR_ARGUMENT;
    }
    if (n == 0) return SPHB_OK;
    normalize_kernel<<<div_up(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(num, den, n);
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}

extern "C" int sphb_sum_slots(double *dst, const double *own, const double *slots, int n_slots, int64_t slot_stride,
                              int skip, int64_t n, sphb_stream stream)
{
    if (!dst || !own || n < 0 || n_slots < 0 || (n_slots > 0 && !slots) || slot_stride < 0) {
        set_error("sum_slots: null or negative argument");
        return SPHB_ERR_ARGUMENT;
    }
    // 16-byte accesses when everything is aligned for them (the usual case: whole z-planes), scalar otherwise
    const bool vec = !(slot_stride & 1) && !((uintptr_t)dst & 15) && !((uintptr_t)own & 15) && !((uintptr_t)slots & 15);
    if (n == 0) return SPHB_OK;
    int sms = 0;
    SPHB_CUDA(sm_count(&sms));
    const int grid = (int)std::min<int64_t>(div_up(n, 512), (int64_t)sms * 8);
    if (vec)
        sum_slots_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dst, own, slots, n_slots,
                                                                                  slot_stride, skip, n);
    else
        sum_slots_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(dst, own, slots, n_slots,
                                                                                   slot_stride, skip, n);
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}
