/*
 * sph_oracle.c -- CPU restatement of Sarracen's particle -> grid scatter.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the B200
 * kernels in sarracen_b200/csrc.  Nothing in the product path may link, load
 * or call it: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do.
 *
 * Parity status: PINNED.  the .npz files under tests/golden/ hold outputs of the unmodified
 * reference (sarracen 1.3.1 CPUBackend, imported from /root/reference by
 * tests/golden/make_golden.py); tests/test_oracle_golden.py checks every
 * routine below against them, plus the closed-form / 10-digit literals of the
 * reference's own tests/kernels/test_kernels.py.
 *
 * Each routine cites the reference file:line whose arithmetic it restates
 * (paths relative to the reference root, sarracen/...): plain C, double
 * precision, OpenMP threads with private images like the reference's prange
 * blocks.  The sampled scatter, the line cuts and the kernels are restated
 * from their definition; the exact-integral functions (full_integral_3d,
 * get_i_terms, line_int3d, surface_int and the 2-D f1 / f2 / f3 family) FOLLOW
 * THE REFERENCE LINE BY LINE, keeping its variable names
 * (kernels/cubic_spline_exact.py:7-483) -- the closed forms of Petkova et al.
 * (2018) have to be evaluated in the same order for the oracle to agree with
 * the reference to rounding.  Being a checker, that is what it is for.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_CUBIC 0
#define ORC_QUARTIC 1
#define ORC_QUINTIC 2
#define ORC_LUT 3

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct {
    int kind;          /* ORC_* */
    int ndim;          /* normalisation dimension handed to w(q, ndim) */
    double radius;     /* compact support radius in units of h */
    const double *lut; /* column-integrated table (kind == ORC_LUT) */
    int samples;
} orc_kernel;

/* ---------------------------------------------------------------- kernels */

/* kernels/cubic_spline.py:14-22 */
static double w_cubic(double q, int ndim)
{
    double norm = ndim == 1 ? 2.0 / 3.0 : ndim == 2 ? 10.0 / (7.0 * M_PI) : 1.0 / M_PI;
    double v = 0.0;
    if (q >= 0.0 && q < 1.0)
        v = 1.0 - 1.5 * q * q + 0.75 * q * q * q;
    else if (q >= 1.0 && q < 2.0)
        v = 0.25 * (2.0 - q) * (2.0 - q) * (2.0 - q);
    return norm * v;
}

static double p4(double a) { double s = a * a; return s * s; }
static double p5(double a) { double s = a * a; return s * s * a; }

/* kernels/quartic_spline.py:14-23 */
static double w_quartic(double q, int ndim)
{
    double norm = ndim == 1 ? 1.0 / 24.0 : ndim == 2 ? 96.0 / (1199.0 * M_PI) : 1.0 / (20.0 * M_PI);
    double v = 0.0;
    if (q < 0.0) return 0.0;
    if (q < 2.5) v += p4(2.5 - q);
    if (q < 1.5) v -= 5.0 * p4(1.5 - q);
    if (q < 0.5) v += 10.0 * p4(0.5 - q);
    return norm * v;
}

/* kernels/quintic_spline.py:14-23 */
static double w_quintic(double q, int ndim)
{
    double norm = ndim == 1 ? 1.0 / 120.0 : ndim == 2 ? 7.0 / (478.0 * M_PI) : 1.0 / (120.0 * M_PI);
    double v = 0.0;
    if (q < 0.0) return 0.0;
    if (q < 3.0) v += p5(3.0 - q);
    if (q < 2.0) v -= 6.0 * p5(2.0 - q);
    if (q < 1.0) v += 15.0 * p5(1.0 - q);
    return norm * v;
}

double orc_w(int kind, double q, int ndim)
{
    switch (kind) {
    case ORC_CUBIC: return w_cubic(q, ndim);
    case ORC_QUARTIC: return w_quartic(q, ndim);
    default: return w_quintic(q, ndim);
    }
}

double orc_radius(int kind)
{
    return kind == ORC_CUBIC ? 2.0 : kind == ORC_QUARTIC ? 2.5 : 3.0;
}

/* kernels/base_kernel.py:39-66 and :102-117 -- column-integrated table:
 * out[i] = 2 * trapezoid over q_z in linspace(0, sqrt(R^2 - q_xy^2), S) of
 * W3D(sqrt(q_xy^2 + q_z^2)), q_xy = R i / (S - 1). */
void orc_column_kernel(int kind, int samples, double *out)
{
    double R = orc_radius(kind);
    for (int i = 0; i < samples; i++) {
        double qxy = R * i / (samples - 1);
        double bound = sqrt(R * R - qxy * qxy);
        double step = bound / (samples - 1);
        double acc = 0.0, zprev = 0.0, yprev = orc_w(kind, sqrt(qxy * qxy), 3);
        for (int k = 1; k < samples; k++) {
            double z = (k == samples - 1) ? bound : k * step;
            double y = orc_w(kind, sqrt(qxy * qxy + z * z), 3);
            acc += (z - zprev) * (y + yprev) / 2.0;
            zprev = z;
            yprev = y;
        }
        out[i] = 2.0 * acc;
    }
}

/* kernels/base_kernel.py:89-97 -- linear interpolation in the table, index
 * clamped to [0, S-1]; the normalisation dimension is ignored. */
static double w_lut(const orc_kernel *k, double q)
{
    double t = q * (k->samples - 1) / k->radius;
    long i0 = (long)floor(t), i1 = (long)ceil(t);
    if (i0 < 0) i0 = 0;
    if (i0 > k->samples - 1) i0 = k->samples - 1;
    if (i1 < 0) i1 = 0;
    if (i1 > k->samples - 1) i1 = k->samples - 1;
    double f = t - (double)i0;
    /* written as a + f (b - a): equal to a (1 - f) + b f up to rounding, and
     * exact when both entries coincide (q < 0 or q >= R), which the
     * reference's own test relies on (tests/kernels/test_kernels.py:191-193) */
    return k->lut[i0] + f * (k->lut[i1] - k->lut[i0]);
}

static double kern_w(const orc_kernel *k, double q)
{
    if (k->kind == ORC_LUT) return w_lut(k, q);
    return orc_w(k->kind, q, k->ndim);
}

double orc_kernel_eval(const orc_kernel *k, double q) { return kern_w(k, q); }

static int nthreads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_num_threads(void) { return nthreads(); }

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static void reduce_private(double *out, const double *priv, int nt, size_t len)
{
    memset(out, 0, len * sizeof(double));
    for (int t = 0; t < nt; t++) {
        const double *p = priv + (size_t)t * len;
        for (size_t i = 0; i < len; i++) out[i] += p[i];
    }
}

static long clampl(long v, long lo, long hi) { return v < lo ? lo : v > hi ? hi : v; }

/* --------------------------------------------------- sampled 2-D scatter */

/* interpolate/cpu_backend.py:226-313 (_fast_2d).  z == NULL means the 2-D /
 * column case (dz = 0); n_dims is the exponent of h in the particle weight
 * and the normalisation dimension of the analytic kernels. */
int orc_fast_2d(const double *x, const double *y, const double *z, double z_slice,
                const double *w, const double *h, int64_t n, const orc_kernel *kern,
                int nx, int ny, double x_min, double x_max, double y_min, double y_max,
                int n_dims, double *out, int64_t *contributions)
{
    const double pwx = (x_max - x_min) / nx, pwy = (y_max - y_min) / ny;
    const double R = kern->radius;
    const size_t len = (size_t)nx * ny;
    const int nt = nthreads();
    double *priv = (double *)calloc((size_t)nt * len, sizeof(double));
    if (!priv) return -1;
    int64_t total = 0;

#pragma omp parallel for schedule(static, 1) reduction(+ : total)
    for (int t = 0; t < nt; t++) {
        double *img = priv + (size_t)t * len;
        int64_t lo = (int64_t)((double)t * ((double)n / nt));
        int64_t hi = (t == nt - 1) ? n : (int64_t)((double)(t + 1) * ((double)n / nt));
        for (int64_t p = lo; p < hi; p++) {
            double hp = h[p];
            double dz = z ? z_slice - z[p] : 0.0;
            if (fabs(dz) >= R * hp) continue; /* :264 */
            double rad = R * hp;
            long i0 = lrint((x[p] - rad - x_min) / pwx); /* :270-273 */
            long j0 = lrint((y[p] - rad - y_min) / pwy);
            long i1 = lrint((x[p] + rad - x_min) / pwx);
            long j1 = lrint((y[p] + rad - y_min) / pwy);
            if (i1 < 0 || i0 > nx) continue; /* :275-278 */
            if (j1 < 0 || j0 > ny) continue;
            i0 = clampl(i0, 0, nx); i1 = clampl(i1, 0, nx);
            j0 = clampl(j0, 0, ny); j1 = clampl(j1, 0, ny);
            double term = w[p] / (n_dims == 2 ? hp * hp : hp * hp * hp); /* :251 */
            double h2 = hp * hp;
            for (long j = j0; j < j1; j++) {
                double dy = y_min + (j + 0.5) * pwy - y[p];
                double dy2 = dy * dy * (1.0 / h2); /* :296 */
                for (long i = i0; i < i1; i++) {
                    double dx = x_min + (i + 0.5) * pwx - x[p];
                    double q = sqrt((dx * dx + dz * dz) / h2 + dy2); /* :290-291, :299 */
                    if (q > R) continue;                             /* :303 */
                    img[(size_t)j * nx + i] += term * kern_w(kern, q);
                    total++;
                }
            }
        }
    }
    reduce_private(out, priv, nt, len);
    free(priv);
    if (contributions) *contributions = total;
    return 0;
}

/* interpolate/cpu_backend.py:193-220 -- the grid is nz independent
 * cross-sections at the voxel-centre depths. */
int orc_grid_3d(const double *x, const double *y, const double *z, const double *w,
                const double *h, int64_t n, const orc_kernel *kern, int nx, int ny, int nz,
                double x_min, double x_max, double y_min, double y_max, double z_min,
                double z_max, double *out, int64_t *contributions)
{
    const double pwz = (z_max - z_min) / nz;
    int64_t total = 0;
    for (int k = 0; k < nz; k++) {
        int64_t c = 0;
        double zc = z_min + (k + 0.5) * pwz;
        int rc = orc_fast_2d(x, y, z, zc, w, h, n, kern, nx, ny, x_min, x_max, y_min, y_max, 3,
                             out + (size_t)k * nx * ny, &c);
        if (rc) return rc;
        total += c;
    }
    if (contributions) *contributions = total;
    return 0;
}

/* ------------------------------------------------------------- line cuts */

/* interpolate/cpu_backend.py:450-538 (_fast_2d_line): slope/intercept form of
 * the cut, support-circle intersection by the quadratic formula, roots
 * clipped to [x1, x2], pixel span by rint, weight w/h^2, W(q, 2). */
int orc_fast_2d_line(const double *x, const double *y, const double *w, const double *h,
                     int64_t n, const orc_kernel *kern, int pixels, double x1, double x2,
                     double y1, double y2, double *out)
{
    double gradient = 0.0;
    if (x2 - x1 != 0.0) gradient = (y2 - y1) / (x2 - x1);
    const double yint = y2 - gradient * x2;
    const double xlength = sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
    const double pixwidth = xlength / pixels, xpixwidth = (x2 - x1) / pixels;
    const double aa = 1.0 + gradient * gradient;
    const double R = kern->radius;
    const int nt = nthreads();
    double *priv = (double *)calloc((size_t)nt * pixels, sizeof(double));
    if (!priv) return -1;
    orc_kernel k2 = *kern;
    k2.ndim = 2;

#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        double *img = priv + (size_t)t * pixels;
        int64_t lo = (int64_t)((double)t * ((double)n / nt));
        int64_t hi = (t == nt - 1) ? n : (int64_t)((double)(t + 1) * ((double)n / nt));
        for (int64_t p = lo; p < hi; p++) {
            double bb = 2.0 * gradient * (yint - y[p]) - 2.0 * x[p];
            double cc = x[p] * x[p] + y[p] * y[p] - 2.0 * yint * y[p] + yint * yint -
                        (R * h[p]) * (R * h[p]);
            double det = bb * bb - 4.0 * aa * cc;
            if (!(det >= 0.0)) continue;
            det = sqrt(det);
            double rs = (-bb - det) / (2.0 * aa), re = (-bb + det) / (2.0 * aa);
            rs = fmin(fmax(rs, x1), x2); /* clip(a_min=x1, a_max=x2) */
            re = fmin(fmax(re, x1), x2);
            double ds = sqrt((rs - x1) * (rs - x1) + (gradient * rs + yint - y1) * (gradient * rs + yint - y1));
            double de = sqrt((re - x1) * (re - x1) + (gradient * re + yint - y1) * (gradient * re + yint - y1));
            long i0 = clampl(lrint(ds / pixwidth), 0, pixels);
            long i1 = clampl(lrint(de / pixwidth), 0, pixels);
            double term = w[p] / (h[p] * h[p]);
            for (long i = i0; i < i1; i++) {
                double xp = x1 + (i + 0.5) * xpixwidth;
                double yp = gradient * xp + yint;
                double dx = xp - x[p], dy = yp - y[p];
                double q = sqrt((dx * dx + dy * dy) / (h[p] * h[p]));
                img[i] += term * kern_w(&k2, q);
            }
        }
    }
    reduce_private(out, priv, nt, (size_t)pixels);
    free(priv);
    return 0;
}

/* Python's round(): half to even, like rint in the default rounding mode. */
static long pyround(double v) { return lrint(v); }

/* interpolate/cpu_backend.py:540-609 (_fast_3d_line): ray / support-sphere
 * discriminant, pixel span round(d / length * pixels), weight w/h^3, W(q, 3). */
int orc_fast_3d_line(const double *x, const double *y, const double *z, const double *w,
                     const double *h, int64_t n, const orc_kernel *kern, int pixels, double x1,
                     double x2, double y1, double y2, double z1, double z2, double *out)
{
    const double ddx = x2 - x1, ddy = y2 - y1, ddz = z2 - z1;
    const double length = sqrt(ddx * ddx + ddy * ddy + ddz * ddz);
    const double ux = ddx / length, uy = ddy / length, uz = ddz / length;
    const double R = kern->radius;
    const int nt = nthreads();
    double *priv = (double *)calloc((size_t)nt * pixels, sizeof(double));
    if (!priv) return -1;

#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; t++) {
        double *img = priv + (size_t)t * pixels;
        int64_t lo = (int64_t)((double)t * ((double)n / nt));
        int64_t hi = (t == nt - 1) ? n : (int64_t)((double)(t + 1) * ((double)n / nt));
        orc_kernel k3 = *kern;
        k3.ndim = 3;
        for (int64_t p = lo; p < hi; p++) {
            double dx = x1 - x[p], dy = y1 - y[p], dz = z1 - z[p];
            double proj = ux * dx + uy * dy + uz * dz;
            double delta = proj * proj - (dx * dx + dy * dy + dz * dz) + (R * h[p]) * (R * h[p]);
            if (delta < 0.0) continue;
            double d1 = -proj - sqrt(delta), d2 = -proj + sqrt(delta);
            long i0 = clampl(pyround(d1 / length * pixels), 0, pixels);
            long i1 = clampl(pyround(d2 / length * pixels), 0, pixels);
            double term = w[p] / (h[p] * h[p] * h[p]);
            for (long i = i0; i < i1; i++) {
                double xd = x1 + (i + 0.5) * (x2 - x1) / pixels - x[p];
                double yd = y1 + (i + 0.5) * (y2 - y1) / pixels - y[p];
                double zd = z1 + (i + 0.5) * (z2 - z1) / pixels - z[p];
                double q = sqrt((xd * xd + yd * yd + zd * zd) / (h[p] * h[p]));
                img[i] += term * kern_w(&k3, q);
            }
        }
    }
    reduce_private(out, priv, nt, (size_t)pixels);
    free(priv);
    return 0;
}

/* ------------------------------------------- exact cubic-spline integrals */
/* Petkova, Laibe & Bonnell (2018) as laid out in
 * kernels/cubic_spline_exact.py. */

/* cubic_spline_exact.py:124-155 */
static double f1_2d(double phi, double q0)
{
    double c = cos(phi), c2 = c * c, tn = tan(phi);
    double logs = log(tan(phi / 2.0 + M_PI / 4.0));
    double i2 = tn;
    double i4 = 1.0 / 3.0 * tn * (2.0 + 1.0 / c2);
    double i5 = 1.0 / 16.0 * (0.5 * (11.0 * sin(phi) + 3.0 * sin(3.0 * phi)) / c2 / c2 + 6.0 * logs);
    double q02 = q0 * q0;
    return 5.0 / 7.0 * q02 / M_PI * (i2 - 0.75 * q02 * i4 + 0.3 * q02 * q0 * i5);
}

/* cubic_spline_exact.py:158-196 */
static double f2_2d(double phi, double q0)
{
    double c = cos(phi), c2 = c * c, tn = tan(phi);
    double q02 = q0 * q0, q03 = q02 * q0;
    double logs = log(tan(phi / 2.0 + M_PI / 4.0));
    double i0 = phi, i2 = tn;
    double i3 = 0.5 * (tn / c + logs);
    double i4 = 1.0 / 3.0 * tn * (2.0 + 1.0 / c2);
    double i5 = 1.0 / 16.0 * (0.5 * (11.0 * sin(phi) + 3.0 * sin(3.0 * phi)) / c2 / c2 + 6.0 * logs);
    return 5.0 / 7.0 * q02 / M_PI *
           (2.0 * i2 - 2.0 * q0 * i3 + 0.75 * q02 * i4 - 0.1 * q03 * i5 - 0.1 / q02 * i0);
}

/* cubic_spline_exact.py:199-217 */
static double f3_2d(double phi) { return 0.5 / M_PI * phi; }

/* cubic_spline_exact.py:58-121 */
static double full_2d_mod(double phi, double q0)
{
    if (q0 <= 1.0) {
        double q = q0 / cos(phi);
        if (q <= 1.0) return f1_2d(phi, q0);
        double pa = acos(q0);
        if (q <= 2.0) return f2_2d(phi, q0) - f2_2d(pa, q0) + f1_2d(pa, q0);
        double pb = acos(0.5 * q0);
        return f3_2d(phi) - f3_2d(pb) + f2_2d(pb, q0) - f2_2d(pa, q0) + f1_2d(pa, q0);
    }
    if (q0 <= 2.0) {
        double q = q0 / cos(phi);
        if (q <= 2.0) return f2_2d(phi, q0);
        double pb = acos(0.5 * q0);
        return f3_2d(phi) - f3_2d(pb) + f2_2d(pb, q0);
    }
    return f3_2d(phi);
}

/* cubic_spline_exact.py:7-55 */
double orc_line_int(double r0, double d1, double d2, double h)
{
    if (r0 == 0.0) return 0.0;
    double sign = r0 > 0.0 ? 1.0 : -1.0, ar0 = fabs(r0);
    double q0 = ar0 / h;
    double pa = atan(fabs(d1) / ar0), pb = atan(fabs(d2) / ar0);
    if (d1 * d2 >= 0.0) return sign * (full_2d_mod(pa, q0) + full_2d_mod(pb, q0));
    if (fabs(d1) < fabs(d2)) return sign * (full_2d_mod(pb, q0) - full_2d_mod(pa, q0));
    return sign * (full_2d_mod(pa, q0) - full_2d_mod(pb, q0));
}

/* cubic_spline_exact.py:457-483 */
static void get_i_terms(double cosp, double a2, double a, double I[6])
{
    double cosp2 = cosp * cosp;
    double p = acos(cosp);
    double tn = sqrt(1.0 - cosp2) / cosp;
    double mu2_1 = 1.0 / (1.0 + cosp2 / a2);
    double u2 = (1.0 - cosp2) * mu2_1, u = sqrt(u2);
    double logs = log((1.0 + u) / (1.0 - u));
    double I1 = atan2(u, a);
    double fac = 1.0 / (1.0 - u2);
    double I_1 = 0.5 * a * logs + I1;
    double I_3 = I_1 + a * 0.25 * (1.0 + a2) * (2.0 * u * fac + logs);
    double I_5 = I_3 + a * (1.0 + a2) * (1.0 + a2) / 16.0 *
                           ((10.0 * u - 6.0 * u * u2) * fac * fac + 3.0 * logs);
    I[0] = p;
    I[1] = I1;
    I[2] = p + a2 * tn;
    I[3] = I_3;
    I[4] = p + 2.0 * a2 * tn + 1.0 / 3.0 * a2 * a2 * tn * (2.0 + 1.0 / cosp2);
    I[5] = I_5;
}

/* cubic_spline_exact.py:353-454 */
static double full_integral_3d(double d, double r0, double r1, double h)
{
    double r0h = r0 / h;
    double phi = atan(fabs(d) / r1);
    if (fabs(r0h) == 0.0 || fabs(r1 / h) == 0.0 || fabs(phi) == 0.0) return 0.0;

    double h2 = h * h, r03 = r0 * r0 * r0;
    double r0h2 = r0h * r0h, r0h3 = r0h2 * r0h;
    double r0h_2 = 1.0 / r0h2, r0h_3 = 1.0 / r0h3;
    double b1 = 0.0, b2 = 0.0, b3;
    if (r0 > 2.0 * h) {
        b3 = 0.25 * h2 * h;
    } else if (r0 > h) {
        b3 = 0.25 * r03 * (-4.0 / 3.0 + r0h - 0.3 * r0h2 + 1.0 / 30.0 * r0h3 - 1.0 / 15.0 * r0h_3 + 8.0 / 5.0 * r0h_2);
        b2 = 0.25 * r03 * (-4.0 / 3.0 + r0h - 0.3 * r0h2 + 1.0 / 30.0 * r0h3 - 1.0 / 15.0 * r0h_3);
    } else {
        b3 = 0.25 * r03 * (-2.0 / 3.0 + 0.3 * r0h2 - 0.1 * r0h3 + 1.4 * r0h_2);
        b2 = 0.25 * r03 * (-2.0 / 3.0 + 0.3 * r0h2 - 0.1 * r0h3 - 0.2 * r0h_2);
        b1 = 0.25 * r03 * (-2.0 / 3.0 + 0.3 * r0h2 - 0.1 * r0h3);
    }

    double a = r1 / r0, a2 = a * a;
    double linedist2 = r0 * r0 + r1 * r1;
    double r_ = r1 / cos(phi);
    double r2 = r0 * r0 + r_ * r_;
    double d2 = 0.0, d3 = 0.0, I[6];

    if (linedist2 < h2) {
        get_i_terms(r1 / sqrt(h2 - r0 * r0), a2, a, I);
        d2 = -1.0 / 6.0 * I[2] + 0.25 * r0h * I[3] - 0.15 * r0h2 * I[4];
        d2 += 1.0 / 30.0 * r0h3 * I[5] - 1.0 / 60.0 * r0h_3 * I[1];
        d2 += (b1 - b2) / r03 * I[0];
    }
    if (linedist2 < 4.0 * h2) {
        get_i_terms(r1 / sqrt(4.0 * h2 - r0 * r0), a2, a, I);
        d3 = 1.0 / 3.0 * I[2] - 0.25 * r0h * I[3] + 3.0 / 40.0 * r0h2 * I[4];
        d3 += -1.0 / 120.0 * r0h3 * I[5] + 4.0 / 15.0 * r0h_3 * I[1];
        d3 += (b2 - b3) / r03 * I[0] + d2;
    }
    get_i_terms(cos(phi), a2, a, I);
    if (r2 <= h2)
        return r0h3 / M_PI * (1.0 / 6.0 * I[2] - 3.0 / 40.0 * r0h2 * I[4] + 1.0 / 40.0 * r0h3 * I[5] + b1 / r03 * I[0]);
    if (r2 <= 4.0 * h2)
        return r0h3 / M_PI *
               (0.25 * (4.0 / 3.0 * I[2] - (r0 / h) * I[3] + 0.3 * r0h2 * I[4] - 1.0 / 30.0 * r0h3 * I[5] + 1.0 / 15.0 * r0h_3 * I[1]) +
                b2 / r03 * I[0] + d2);
    return r0h3 / M_PI * (-0.25 * r0h_3 * I[1] + b3 / r03 * I[0] + d3);
}

/* cubic_spline_exact.py:284-350 (diagnostic prints omitted) */
static double line_int3d(double r0, double r1, double d1, double d2, double h)
{
    if (fabs(r0) == 0.0) return 0.0;
    double sign = r0 > 0.0 ? 1.0 : -1.0, ar0 = fabs(r0), ar1;
    if (r1 > 0.0) {
        ar1 = r1;
    } else {
        sign = -sign;
        ar1 = -r1;
    }
    double int1 = full_integral_3d(d1, ar0, ar1, h);
    double int2 = full_integral_3d(d2, ar0, ar1, h);
    if (int1 < 0.0) int1 = 0.0;
    if (int2 < 0.0) int2 = 0.0;
    if (d1 * d2 >= 0.0) return sign * (int1 + int2);
    if (fabs(d1) < fabs(d2)) return sign * (int2 - int1);
    return sign * (int1 - int2);
}

/* cubic_spline_exact.py:220-281 */
double orc_surface_int(double r0, double x1, double y1, double x2, double y2, double wx,
                       double wy, double h)
{
    double dx = x2 - x1, dy = y2 - y1, res = 0.0;
    res += line_int3d(r0, 0.5 * wy + dy, 0.5 * wx - dx, 0.5 * wx + dx, h);
    res += line_int3d(r0, 0.5 * wy - dy, 0.5 * wx + dx, 0.5 * wx - dx, h);
    res += line_int3d(r0, 0.5 * wx + dx, 0.5 * wy + dy, 0.5 * wy - dy, h);
    res += line_int3d(r0, 0.5 * wx - dx, 0.5 * wy - dy, 0.5 * wy + dy, h);
    return res;
}

/* ---------------------------------------------------- exact 2-D rendering */

/* interpolate/cpu_backend.py:317-447 (_exact_2d_render): pixel-area integral
 * of the 2-D cubic spline from the four edge integrals; interior edges are
 * evaluated once and shared with opposite sign.  The bounds test uses the
 * GPU backend's `>=` (gpu_backend.py:383): the CPU `>` at cpu_backend.py:353
 * writes out of bounds (SURVEY note N3) and is not reproduced. */
int orc_exact_2d(const double *x, const double *y, const double *w, const double *h, int64_t n,
                 int nx, int ny, double x_min, double x_max, double y_min, double y_max,
                 double *out, int64_t *contributions)
{
    const double pwx = (x_max - x_min) / nx, pwy = (y_max - y_min) / ny;
    const size_t len = (size_t)nx * ny;
    const int nt = nthreads();
    double *priv = (double *)calloc((size_t)nt * len, sizeof(double));
    if (!priv) return -1;
    int64_t total = 0;

#pragma omp parallel for schedule(static, 1) reduction(+ : total)
    for (int t = 0; t < nt; t++) {
        double *img = priv + (size_t)t * len;
        int64_t lo = (int64_t)((double)t * ((double)n / nt));
        int64_t hi = (t == nt - 1) ? n : (int64_t)((double)(t + 1) * ((double)n / nt));
        for (int64_t p = lo; p < hi; p++) {
            double hp = h[p], rad = 2.0 * hp;
            long i0 = (long)floor((x[p] - rad - x_min) / pwx);
            long j0 = (long)floor((y[p] - rad - y_min) / pwy);
            long i1 = (long)ceil((x[p] + rad - x_min) / pwx);
            long j1 = (long)ceil((y[p] + rad - y_min) / pwy);
            if (i1 < 0 || i0 >= nx) continue;
            if (j1 < 0 || j0 >= ny) continue;
            i0 = clampl(i0, 0, nx); i1 = clampl(i1, 0, nx);
            j0 = clampl(j0, 0, ny); j1 = clampl(j1, 0, ny);
            /* A particle whose clamped pixel range is empty in either
             * direction overlaps no pixel.  The reference still adds the lone
             * top-row / left-column edge terms for it (:370, :387), leaving an
             * unbalanced edge integral in the border pixels; that artefact is
             * not reproduced ("reference behaviours not
             * copied"). */
            if (i1 <= i0 || j1 <= j0) continue;
            double term = w[p] / (hp * hp);
            double denom = 1.0 / fabs(pwx * pwy) * hp * hp;

            /* top edges of the first row (:370-385) */
            {
                double dy = y_min + (j0 + 0.5) * pwy - y[p];
                for (long i = i0; i < i1; i++) {
                    double dx = x_min + (i + 0.5) * pwx - x[p];
                    double v = orc_line_int(0.5 * pwy - dy, 0.5 * pwx + dx, 0.5 * pwx - dx, hp);
                    img[(size_t)j0 * nx + i] += term * (v * denom);
                }
            }
            /* left edges of the first column (:387-402) */
            {
                double dx = x_min + (i0 + 0.5) * pwx - x[p];
                for (long j = j0; j < j1; j++) {
                    double dy = y_min + (j + 0.5) * pwy - y[p];
                    double v = orc_line_int(0.5 * pwx - dx, 0.5 * pwy - dy, 0.5 * pwy + dy, hp);
                    img[(size_t)j * nx + i0] += term * (v * denom);
                }
            }
            /* bottom and right edge of every pixel (:404-440) */
            for (long j = j0; j < j1; j++) {
                double dy = y_min + (j + 0.5) * pwy - y[p];
                for (long i = i0; i < i1; i++) {
                    double dx = x_min + (i + 0.5) * pwx - x[p];
                    double v = orc_line_int(0.5 * pwy + dy, 0.5 * pwx - dx, 0.5 * pwx + dx, hp) * denom;
                    img[(size_t)j * nx + i] += term * v;
                    if (j < j1 - 1) img[(size_t)(j + 1) * nx + i] -= term * v;
                    v = orc_line_int(0.5 * pwx + dx, 0.5 * pwy + dy, 0.5 * pwy - dy, hp) * denom;
                    img[(size_t)j * nx + i] += term * v;
                    if (i < i1 - 1) img[(size_t)j * nx + i + 1] -= term * v;
                    total++;
                }
            }
        }
    }
    reduce_private(out, priv, nt, len);
    free(priv);
    if (contributions) *contributions = total;
    return 0;
}

/* -------------------------------------------- exact 3-D column projection */

/* interpolate/cpu_backend.py:611-720 (_exact_3d_project): volume integral of
 * the 3-D cubic spline over the pixel column pwx x pwy x 4h from its six face
 * integrals; per-pixel cull on the pixel centre (:675). */
int orc_exact_3d_project(const double *x, const double *y, const double *w, const double *h,
                         int64_t n, int nx, int ny, double x_min, double x_max, double y_min,
                         double y_max, double *out, int64_t *contributions)
{
    const double pwx = (x_max - x_min) / nx, pwy = (y_max - y_min) / ny;
    const double norm3d = 1.0 / M_PI;
    const size_t len = (size_t)nx * ny;
    const int nt = nthreads();
    double *priv = (double *)calloc((size_t)nt * len, sizeof(double));
    if (!priv) return -1;
    int64_t total = 0;

#pragma omp parallel for schedule(static, 1) reduction(+ : total)
    for (int t = 0; t < nt; t++) {
        double *img = priv + (size_t)t * len;
        int64_t lo = (int64_t)((double)t * ((double)n / nt));
        int64_t hi = (t == nt - 1) ? n : (int64_t)((double)(t + 1) * ((double)n / nt));
        for (int64_t p = lo; p < hi; p++) {
            double hp = h[p], rad = 2.0 * hp, xp = x[p], yp = y[p];
            long i0 = (long)floor((xp - rad - x_min) / pwx);
            long j0 = (long)floor((yp - rad - y_min) / pwy);
            long i1 = (long)ceil((xp + rad - x_min) / pwx);
            long j1 = (long)ceil((yp + rad - y_min) / pwy);
            double pwz = 4.0 * hp;
            if (i1 < 0 || i0 >= nx) continue;
            if (j1 < 0 || j0 >= ny) continue;
            i0 = clampl(i0, 0, nx); i1 = clampl(i1, 0, nx);
            j0 = clampl(j0, 0, ny); j1 = clampl(j1, 0, ny);
            double h3 = hp * hp * hp;
            double dfac = h3 / (pwx * pwy * norm3d);
            double term = norm3d * w[p] / h3;
            for (long j = j0; j < j1; j++) {
                double ypix = y_min + (j + 0.5) * pwy, dy = ypix - yp;
                for (long i = i0; i < i1; i++) {
                    double xpix = x_min + (i + 0.5) * pwx, dx = xpix - xp;
                    double q2 = (dx * dx + dy * dy) / (hp * hp);
                    if (!(q2 < 4.0 + 3.0 * pwx * pwy / (hp * hp))) continue;
                    double v = 2.0 * orc_surface_int(0.5 * pwz, xp, yp, xpix, ypix, pwx, pwy, hp);
                    v += orc_surface_int(ypix - yp + 0.5 * pwy, xp, 0.0, xpix, 0.0, pwx, pwz, hp);
                    v += orc_surface_int(yp - ypix + 0.5 * pwy, xp, 0.0, xpix, 0.0, pwx, pwz, hp);
                    v += orc_surface_int(xpix - xp + 0.5 * pwx, 0.0, yp, 0.0, ypix, pwz, pwy, hp);
                    v += orc_surface_int(xp - xpix + 0.5 * pwx, 0.0, yp, 0.0, ypix, pwz, pwy, hp);
                    img[(size_t)j * nx + i] += term * (v * dfac);
                    total++;
                }
            }
        }
    }
    reduce_private(out, priv, nt, len);
    free(priv);
    if (contributions) *contributions = total;
    return 0;
}
