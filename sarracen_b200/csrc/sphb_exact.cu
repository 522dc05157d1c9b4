// Exact (pixel-integrated) cubic-spline rendering for B200 (sm_100a).
//
// Replaces CPUBackend._exact_2d_render (sarracen/interpolate/cpu_backend.py:
// 317-447) and CPUBackend._exact_3d_project (:611-720).  The integrals are the
// closed forms of Petkova, Laibe & Bonnell (2018) as the reference lays them
// out in sarracen/kernels/cubic_spline_exact.py; they are restated here from
// those formulas, in double precision (no fast-math), because the pixel value
// is a difference of edge/face terms that cancel to many digits.
//
// Decomposition: these paths cost 10^3 - 10^4 double-precision operations per
// (particle, pixel), so accumulation is not the bottleneck; load balance (boxes
// range from 1 to thousands of pixels) and redundant transcendental work are.
//
//  - LATTICE FORM.  The pixel edges (2-D) / column faces (3-D) of a particle's
//    box lie on the lines of a lattice, and every term of the closed forms
//    belongs either to a lattice LINE (its distance from the particle) or to a
//    lattice VERTEX (an end point on a line), shared by the two edges / faces
//    that meet there.  The lanes of a warp take consecutive vertices of a line,
//    evaluate the vertex terms once, and form each edge / face from a lane and its
//    neighbour with shuffles; an edge / face value is added to the pixel on one
//    side of the line and subtracted from the pixel on the other (the integrals
//    are odd in the signed distance), as the reference does for its shared 2-D
//    edges (:404-440).  2-D: one closed-form evaluation per edge instead of the
//    reference's four per pixel edge pair; 3-D: 2 edge set-ups and 6 half-segments
//    per pixel instead of the reference's ~40 `_full_integral_3d` calls.
//  - WORK UNITS of 32 vertices (31 edges; consecutive units overlap by one
//    vertex) are counted per particle in chunks of 16 and scanned; warps take
//    chunks grid-stride, find the chunk's particle by a 33-way search in the
//    scanned offsets, and run its units one after the other.  The 3-D path trims
//    the lattice to the reference's cull disc first (see "Lattice form" below).
//  - The region structure of the closed forms (q <= 1, <= 2, > 2) is evaluated
//    without data-dependent branches: all three forms share their integrals and
//    the result is selected (see line_crossings / half_segment).
//  - A particle whose clamped pixel range is empty contributes nothing; the
//    reference's lone top-row / left-column edge terms for such a particle
//    (:370, :387) and its out-of-bounds row write (:353) are not reproduced.
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "sphb_common.cuh"

namespace sphb {

namespace ex {

// The closed forms below need sin, cos and tan of an angle phi plus
// L = ln tan(phi/2 + pi/4) and, in two of the three regions, phi itself.  Every
// angle on this path is either atan(t) of a known tangent t (segment end
// points) or acos(c) of a known cosine c (where the segment crosses q = 1 or
// q = 2), so the trigonometric values are algebraic in t resp. c:
//   cos = 1/sqrt(1+t^2), sin = t cos, L = ln((1 + sin)/cos) = ln(t + sqrt(1+t^2)),
//   sin 3phi = 3 sin - 4 sin^3.
// Only atan / acos (for phi) and one log per angle remain -- the same numbers
// the reference gets from cos(atan(t)) etc., to rounding.
// 2-D closed forms: one shared copy of each double-precision transcendental (each is ~100 SASS
// instructions).  Inlined at every use they put the kernel far beyond the instruction cache (ncu: 3.9
// 'no instruction' stalls per issue, profiles/r1_exact2d_v2.txt); shared: 13.8 -> 12.2 ms.  The 3-D
// face integrals keep the inlined library calls: there the calls cost more in spills than the
// instruction fetch saves (62.7 -> 65.3 ms).
__device__ __noinline__ double m_log(double a) { return log(a); }
__device__ __noinline__ double m_atan(double a) { return atan(a); }
__device__ __noinline__ double m_acos(double a) { return acos(a); }
__device__ __noinline__ double m_atan2(double a, double b) { return atan2(a, b); }

struct Angle {
    double s, c, tn, logs, phi;
};

__device__ __forceinline__ Angle angle_from_tan(double t)
{
    Angle a;
    t = fmin(t, 1e150);
    const double r = sqrt(fma(t, t, 1.0));
    a.c = 1.0 / r;
    a.s = t * a.c;
    a.tn = t;
    a.logs = m_log(t + r);
    a.phi = m_atan(t);
    return a;
}

__device__ __forceinline__ Angle angle_from_cos(double c)
{
    Angle a;
    a.c = c;
    a.s = sqrt(fmax(1.0 - c * c, 0.0));
    a.tn = a.s / c;
    a.logs = m_log((1.0 + a.s) / c);
    a.phi = m_acos(c);
    return a;
}

// The three closed forms of a half-segment, cubic_spline_exact.py:124-155 (region 0 < q <= 1),
// :158-196 (1 < q <= 2) and :199-217 (q > 2), evaluated TOGETHER: they are polynomials in q0 of
// the same few integrals I_0 = phi, I_2 = tan, I_3, I_4, I_5, all algebraic in the Angle, so the
// three values cost little more than one and the caller selects without branching.
struct F123 {
    double f1, f2, f3;
};

__device__ __forceinline__ F123 f123_2d(const Angle &a, double q0)
{
    const double ic = 1.0 / a.c, ic2 = ic * ic;
    const double q02 = q0 * q0, q03 = q02 * q0;
    const double i0 = a.phi, i2 = a.tn;
    const double i3 = 0.5 * (a.tn * ic + a.logs);
    const double i4 = (1.0 / 3.0) * a.tn * (2.0 + ic2);
    const double s3 = a.s * (20.0 - 12.0 * a.s * a.s); // 11 sin + 3 sin 3phi
    const double i5 = (1.0 / 16.0) * (0.5 * s3 * ic2 * ic2 + 6.0 * a.logs);
    const double k = 5.0 / 7.0 * q02 / kPi;
    F123 f;
    f.f1 = k * (i2 - 0.75 * q02 * i4 + 0.3 * q03 * i5);
    f.f2 = k * (2.0 * i2 - 2.0 * q0 * i3 + 0.75 * q02 * i4 - 0.1 * q03 * i5 - 0.1 / q02 * i0);
    f.f3 = 0.5 / kPi * a.phi;
    return f;
}

// cubic_spline_exact.py:7-55 with :58-121 (`_full_2d_mod`) regrouped.  The reference walks a
// half-segment from the foot of the perpendicular (q = q0) to its end point (q = q0 / cos phi) and
// switches closed form where the segment crosses q = 1 (at phi_a = acos q0) and q = 2
// (phi_b = acos q0/2):
//     q0 <= 1:  F = f1(e)                                            q <= 1
//               F = f2(e) - f2(pa) + f1(pa)                          q <= 2
//               F = f3(e) - f3(pb) + f2(pb) - f2(pa) + f1(pa)        q >  2     (and likewise for q0 > 1)
// The crossing terms c1 = f1(pa) - f2(pa) and c2 = f2(pb) - f3(pb) depend on q0 only -- one value per
// edge, shared by its two half-segments, and (the lanes of a warp walk a column of edges) nearly
// warp-uniform -- so they are worked out once per edge under q0 tests, and each half-segment is
//     F = q <= 1 ? f1(e) : q <= 2 ? f2(e) + c1 : f3(e) + c2 + c1
// (c1 = 0 for q0 > 1, c2 = 0 for q0 > 2): no data-dependent branch on the end point's region, which
// is what made the lanes of a warp diverge (ncu r1: 17.8 of 32 threads active).  Same terms as the
// reference, summed in a different order.
// crossing terms of a lattice line at normalised distance q0 (see above): (c1, c2)
__device__ __noinline__ double2 line_crossings(double q0)
{
    double c1 = 0.0, c2 = 0.0;
    if (q0 <= 1.0) {
        const F123 fa = f123_2d(angle_from_cos(q0), q0);
        c1 = fa.f1 - fa.f2;
    }
    if (q0 <= 2.0) {
        const F123 fb = f123_2d(angle_from_cos(0.5 * q0), q0);
        c2 = fb.f2 - fb.f3;
    }
    return make_double2(c1, c2);
}

// value of the half-segment from the foot of the perpendicular to an end point at tangent t = |d| / |r0|
__device__ __noinline__ double half_segment(double t, double q0, double c1, double c2)
{
    const Angle e = angle_from_tan(t);
    const F123 f = f123_2d(e, q0);
    const double q = q0 / e.c;
    return (q <= 1.0 && q0 <= 1.0) ? f.f1 : (q <= 2.0 && q0 <= 2.0) ? f.f2 + c1 : f.f3 + c2 + c1;
}

// the two half-segments of an edge combined (cubic_spline_exact.py:43-53): `ga`, `gb` are their
// values, `da`, `db` the signed offsets of the end points from the foot of the perpendicular
// measured along the line in ONE direction (the foot lies between them iff da db <= 0)
__device__ __forceinline__ double combine_halves(double ga, double gb, double da, double db)
{
    if (da * db <= 0.0) return ga + gb;
    return fabs(da) < fabs(db) ? gb - ga : ga - gb;
}

// cubic_spline_exact.py:457-483
struct ITerms {
    double i0, i1, i2, i3, i4, i5;
};

// `p` = acos(cosp) and `tn` = tan(p) are passed in: for the segment end point they are known
// (p = atan(t), tn = t) and need no inverse cosine.
__device__ __noinline__ ITerms get_i_terms(double cosp, double p, double tn, double a2, double a)
{
    ITerms t;
    const double cosp2 = cosp * cosp;
    const double mu2_1 = 1.0 / (1.0 + cosp2 / a2);
    const double u2 = (1.0 - cosp2) * mu2_1, u = sqrt(u2);
    const double logs = m_log((1.0 + u) / (1.0 - u));
    const double i1 = m_atan2(u, a);
    const double fac = 1.0 / (1.0 - u2);
    const double im1 = 0.5 * a * logs + i1;
    const double im3 = im1 + a * 0.25 * (1.0 + a2) * (2.0 * u * fac + logs);
    const double im5 = im3 + a * (1.0 + a2) * (1.0 + a2) / 16.0 * ((10.0 * u - 6.0 * u * u2) * fac * fac + 3.0 * logs);
    t.i0 = p;
    t.i1 = i1;
    t.i2 = p + a2 * tn;
    t.i3 = im3;
    t.i4 = p + 2.0 * a2 * tn + (1.0 / 3.0) * a2 * a2 * tn * (2.0 + 1.0 / cosp2);
    t.i5 = im5;
    return t;
}

__device__ __noinline__ ITerms get_i_terms_cos(double cosp, double a2, double a)
{
    return get_i_terms(cosp, m_acos(cosp), sqrt(1.0 - cosp * cosp) / cosp, a2, a);
}

// I_1 alone (the only term besides I_0 = phi that the region outside the kernel support needs):
// no logarithm, no higher terms.
__device__ __forceinline__ double get_i1(double cosp, double a2, double a)
{
    const double cosp2 = cosp * cosp;
    const double mu2_1 = 1.0 / (1.0 + cosp2 / a2);
    const double u = sqrt((1.0 - cosp2) * mu2_1);
    return m_atan2(u, a);
}

// cubic_spline_exact.py:353-454, split in two.  The reference evaluates `full_integral_3d(d, r0, r1, h)`
// for both half-segments d1, d2 of an edge (cubic_spline_exact.py:284-350); everything that does not
// depend on d -- the B terms, a = r1 / r0 and the two inner-boundary corrections D2, D3 with their
// inverse cosines, logarithms and arctangents -- is the same for both calls and is worked out once per
// edge here (Edge3), leaving the phi-dependent part (edge_integral) to be evaluated twice.
struct Edge3 {
    double r0, r1, h2, r03, r0h, r0h2, r0h3, r0h_3;
    double b1, b2, b3, a, a2, d2, d3;
    bool zero; // r0 = 0 or r1 = 0: the integral vanishes for every d
};

__device__ __forceinline__ void edge_setup(double r0, double r1, double h, Edge3 &s)
{
    const double r0h = r0 / h;
    s.zero = fabs(r0h) == 0.0 || fabs(r1 / h) == 0.0;
    if (s.zero) return;
    const double h2 = h * h, r03 = r0 * r0 * r0;
    const double r0h2 = r0h * r0h, r0h3 = r0h2 * r0h;
    const double r0h_2 = 1.0 / r0h2, r0h_3 = 1.0 / r0h3;
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
    const double a = r1 / r0, a2 = a * a;
    const double linedist2 = r0 * r0 + r1 * r1;
    double d2 = 0.0, d3 = 0.0;
    if (linedist2 < h2) {
        const ITerms t = get_i_terms_cos(r1 / sqrt(h2 - r0 * r0), a2, a);
        d2 = -1.0 / 6.0 * t.i2 + 0.25 * r0h * t.i3 - 0.15 * r0h2 * t.i4;
        d2 += 1.0 / 30.0 * r0h3 * t.i5 - 1.0 / 60.0 * r0h_3 * t.i1;
        d2 += (b1 - b2) / r03 * t.i0;
    }
    if (linedist2 < 4.0 * h2) {
        const ITerms t = get_i_terms_cos(r1 / sqrt(4.0 * h2 - r0 * r0), a2, a);
        d3 = 1.0 / 3.0 * t.i2 - 0.25 * r0h * t.i3 + 3.0 / 40.0 * r0h2 * t.i4;
        d3 += -1.0 / 120.0 * r0h3 * t.i5 + 4.0 / 15.0 * r0h_3 * t.i1;
        d3 += (b2 - b3) / r03 * t.i0 + d2;
    }
    s.r0 = r0; s.r1 = r1; s.h2 = h2; s.r03 = r03;
    s.r0h = r0h; s.r0h2 = r0h2; s.r0h3 = r0h3; s.r0h_3 = r0h_3;
    s.b1 = b1; s.b2 = b2; s.b3 = b3; s.a = a; s.a2 = a2; s.d2 = d2; s.d3 = d3;
}

__device__ __forceinline__ double edge_integral(double d, const Edge3 &s)
{
    // phi = atan(|d| / r1): its cosine and tangent are algebraic in t = |d| / r1
    const double tphi = fabs(d) / s.r1;
    if (s.zero || tphi == 0.0) return 0.0;
    const double phi = m_atan(tphi);
    const double tcl = fmin(tphi, 1e150);
    const double sec = sqrt(fma(tcl, tcl, 1.0)); // 1 / cos(phi)
    const double cphi = 1.0 / sec;
    const double r_ = s.r1 * sec;
    const double r2 = s.r0 * s.r0 + r_ * r_;
    const bool outside = !(r2 <= 4.0 * s.h2);
    if (__all_sync(__activemask(), outside)) // the whole warp is beyond the support: I_0 and I_1 only
        return s.r0h3 / kPi * (-0.25 * s.r0h_3 * get_i1(cphi, s.a2, s.a) + s.b3 / s.r03 * phi + s.d3);
    // One evaluation of the I terms serves the three regions: they differ in the coefficients of the same
    // integrals, selected per lane, instead of in code the lanes of a warp would run one after the other
    // (the end points of a column of faces straddle r = h and r = 2h).
    const ITerms t = get_i_terms(cphi, phi, tcl, s.a2, s.a);
    const bool inner = r2 <= s.h2;
    double c0, c1, c2, c3, c4, c5, add;
    if (outside) {
        c0 = s.b3 / s.r03; c1 = -0.25 * s.r0h_3; c2 = c3 = c4 = c5 = 0.0; add = s.d3;
    } else if (inner) {
        c0 = s.b1 / s.r03; c1 = 0.0; c2 = 1.0 / 6.0; c3 = 0.0; c4 = -3.0 / 40.0 * s.r0h2; c5 = 1.0 / 40.0 * s.r0h3;
        add = 0.0;
    } else {
        c0 = s.b2 / s.r03; c1 = 0.25 / 15.0 * s.r0h_3; c2 = 1.0 / 3.0; c3 = -0.25 * s.r0h; c4 = 0.075 * s.r0h2;
        c5 = -0.25 / 30.0 * s.r0h3; add = s.d2;
    }
    // outside the support I_2 .. I_5 are not needed (and may be huge): the zero coefficients must not meet them
    const double hi = outside ? 0.0 : c2 * t.i2 + c3 * t.i3 + c4 * t.i4 + c5 * t.i5;
    return s.r0h3 / kPi * (hi + c1 * t.i1 + c0 * t.i0 + add);
}

// The edge along the line of sight through a lattice vertex: cubic_spline_exact.py:284-350 for a
// segment centred on the foot of the perpendicular (the column is centred on the particle), whose two
// half-segments are equal: line_int3d(r0, r1, dhalf, dhalf, h) = sign * 2 * max(I(dhalf), 0).  Not
// inlined: one copy serves the kernel (instruction cache).
__device__ __noinline__ double los_edge(double r0, double r1, double dhalf, double h)
{
    if (fabs(r0) == 0.0) return 0.0;
    double sign = r0 > 0.0 ? 1.0 : -1.0;
    double ar1 = r1;
    if (!(r1 > 0.0)) {
        sign = -sign;
        ar1 = -r1;
    }
    Edge3 s;
    edge_setup(fabs(r0), ar1, h, s);
    const double v = edge_integral(dhalf, s);
    return v > 0.0 ? sign * (v + v) : 0.0;
}

// One clamped half-segment integral of a line that lies wholly outside the kernel support,
// r0^2 + r1^2 >= (2h)^2 -- the top edges of the side faces (r1 = 2h) and the edges of the top faces
// (r0 = 2h) of a column of height 4h.  There D2 = D3 = 0 and only I_0 = phi and I_1 appear
// (cubic_spline_exact.py:391-397, :438-447):
//     F = (r0/h)^3 / pi * ( -(h/r0)^3 / 4 * I_1 + B3 / r0^3 * phi ) = ( B3 / h^3 * phi - I_1 / 4 ) / pi,
// with phi = atan(|d| / r1) and, from cubic_spline_exact.py:457-483 with cos(phi) = 1 / sqrt(1 + t^2),
// t = |d| / r1, a = r1 / r0:   I_1 = atan2(u, a),  u^2 = a^2 t^2 / (a^2 (1 + t^2) + 1), i.e.
//     I_1 = atan( |d| r0 / (r1 sqrt(r0^2 + r1^2 + d^2)) ).
// B3 / h^3 is a polynomial in r0 / h (the reference's negative powers of r0/h cancel against r0^3):
// two arctangents, two divisions and a square root, no logarithm, no inverse cosine.
__device__ __noinline__ double half_outside(double ar0, double r1, double d, double h)
{
    const double ad = fabs(d);
    if (ar0 == 0.0 || r1 == 0.0 || ad == 0.0) return 0.0;
    const double q = ar0 / h, q2 = q * q, q3 = q2 * q;
    double b3; // B3 / h^3
    if (ar0 > 2.0 * h)
        b3 = 0.25;
    else if (ar0 > h)
        b3 = 0.25 * (q3 * (-4.0 / 3.0 + q - 0.3 * q2 + 1.0 / 30.0 * q3) - 1.0 / 15.0 + 1.6 * q);
    else
        b3 = 0.25 * (q3 * (-2.0 / 3.0 + 0.3 * q2 - 0.1 * q3) + 1.4 * q);
    const double rinv = 1.0 / r1;
    const double phi = m_atan(ad * rinv);
    const double i1 = m_atan(ad * rinv * (ar0 * rsqrt(fma(ar0, ar0, fma(r1, r1, ad * ad)))));
    const double v = (1.0 / kPi) * fma(b3, phi, -0.25 * i1);
    return v < 0.0 ? 0.0 : v;
}

} // namespace ex

struct ExactView {
    int nx, ny;
    double x_min, y_min, pwx, pwy;
    double inv_area; // 1 / (pwx pwy)
};

constexpr int kExactWarps = 4;

// Pixel box of a particle (cpu_backend.py:343-363 / :639-664); false if empty.
template <typename T>
__device__ __forceinline__ bool exact_box(const ExactView &v, const T *x, const T *y, const T *h, int64_t p,
                                          double &xp, double &yp, double &hp, int &i0, int &i1, int &j0, int &j1)
{
    xp = (double)__ldg(x + p);
    yp = (double)__ldg(y + p);
    hp = (double)__ldg(h + p);
    if (!(hp > 0.0) || !isfinite(hp) || !isfinite(xp) || !isfinite(yp)) return false;
    const double rad = 2.0 * hp;
    const double a0 = floor((xp - rad - v.x_min) / v.pwx), a1 = ceil((xp + rad - v.x_min) / v.pwx);
    const double b0 = floor((yp - rad - v.y_min) / v.pwy), b1 = ceil((yp + rad - v.y_min) / v.pwy);
    if (a1 < 0.0 || a0 >= v.nx || b1 < 0.0 || b0 >= v.ny) return false;
    i0 = (int)fmax(a0, 0.0);
    i1 = (int)fmin(a1, (double)v.nx);
    j0 = (int)fmax(b0, 0.0);
    j1 = (int)fmin(b1, (double)v.ny);
    return i1 > i0 && j1 > j0;
}

// ---------------------------------------------------------------------------
// Lattice form of both exact paths.
//
// The pixel edges (2-D) / column faces (3-D) of a particle's box lie on the lines of a lattice: bw + 1
// vertical lines with bh + 1 vertices each, bh + 1 horizontal lines with bw + 1 vertices each.  Everything
// the closed forms need factors over that lattice.
//
// 2-D render (cpu_backend.py:317-447):
//   * the crossing terms (c1, c2) -- two inverse cosines, two logarithms, two sets of closed forms --
//     depend on the LINE only (its distance from the particle): `exact2_lines_kernel` works them out once
//     per line, one warp per particle with the lanes over its lines;
//   * the half-segment value G -- one arctangent, one logarithm, one set of closed forms -- belongs to a
//     VERTEX and serves the two edges that meet there; an edge is formed from a lane and its neighbour
//     (one shuffle).  One closed-form evaluation per edge instead of the reference's four.
//
// 3-D projection (cpu_backend.py:611-720): the column of a pixel has a top and a bottom face (equal
// integrals) and four side faces, each shared with the neighbouring column.  Take a lattice line at signed
// distance r0 from the particle and its slot between the vertices k and k + 1 (offsets d_k, d_{k+1} along
// the line).  Two things live on that slot, and both feed the pixel after the line with + and the pixel
// before it with -:
//   the side face in the plane of the line (surface_int with the face centred on the particle along the
//   line of sight):   S = Z(d_{k+1}) - Z(d_k) + 2 sign(r0) combine(E_k, E_{k+1})
//       Z(d) = line_int3d(r0, r1 = d, 2h, 2h)   the edge along the line of sight through the vertex (los_edge)
//       E_k  = clamped half-segment of the face's top edge: line (|r0|, r1 = 2h), end point d_k
//   the edge of the top (= bottom) faces on the line:   2 sign(r0) combine(T_k, T_{k+1})
//       T_k  = clamped half-segment of the line (2h, r1 = |r0|), end point d_k
// Per pixel that is 2 full edge set-ups and 6 half-segments instead of the reference's 10 and 16.
//
// Which vertices.  2-D: all of them, line by line (vertical lines first, each orientation cut into its own
// units).  3-D: the reference evaluates a pixel only if its centre passes q^2 < 4 + 3 pwx pwy / h^2
// (cpu_backend.py:675), a disc inside the box, so ~ 1 - pi/4 of a large box is never needed.  The lines of
// each orientation are cut into at most 32 GROUPS of G consecutive lines; a group keeps, of every line, only
// the vertex range the widest chord of that disc beside its lines can reach (widened by one pixel, so
// rounding cannot lose a pixel the cull accepts; the cull itself is still applied per pixel).  Inside a
// group the (line, vertex) numbering is rectangular and the groups are laid end to end.  (The same trimming
// of the 2-D lattice to the pixels that touch the support was measured and costs more than it saves there:
// 4.0 against 3.8 ms, a vertex is four times cheaper.)
// Either sequence is cut into UNITS of 32 vertices = 31 slots, consecutive units overlapping by one vertex.
// A warp takes kExactChunk consecutive units of one particle, so the search for the particle, its box and
// the group table are worked out once per chunk (they were two thirds of the 2-D kernel's instructions
// when every unit did them).
constexpr int kExactChunk = 16;
constexpr int kExactGroups = 32;

struct LatBox {
    double sx, sy, sh, r2cull;
    int i0, j0, bw, bh;
};

template <typename T, bool PROJ3>
__device__ __forceinline__ bool lat_box(const ExactView &v, const T *x, const T *y, const T *h, int64_t p, LatBox &b)
{
    int i1, j1;
    if (!exact_box(v, x, y, h, p, b.sx, b.sy, b.sh, b.i0, i1, b.j0, j1)) return false;
    b.bw = i1 - b.i0;
    b.bh = j1 - b.j0;
    b.r2cull = 4.0 * b.sh * b.sh + 3.0 * v.pwx * v.pwy; // (4 + 3 pwx pwy / h^2) h^2: the 3-D cull
    return true;
}

// Lines [g G, g G + nl) of one orientation: the vertex range [klo, klo + cnt) kept of each (cnt = 0: none).
__device__ __forceinline__ void lat_group(const ExactView &v, const LatBox &b, bool vertical, int G, int g, int &nl,
                                          int &klo, int &cnt)
{
    const int nlines = (vertical ? b.bw : b.bh) + 1, len = vertical ? b.bh : b.bw;
    const int l0 = g * G;
    nl = min(G, nlines - l0);
    klo = cnt = 0;
    if (nl <= 0) {
        nl = 0;
        return;
    }
    // across the lines: the pixel columns beside them, centres a0 .. a1; the one nearest to the particle
    const double pa = vertical ? v.pwx : v.pwy, oa = vertical ? v.x_min + b.i0 * v.pwx : v.y_min + b.j0 * v.pwy;
    const double sa = vertical ? b.sx : b.sy;
    const int c0 = max(l0 - 1, 0), c1 = min(l0 + nl - 1, nlines - 2);
    const double a0 = oa + (c0 + 0.5) * pa, a1 = oa + (c1 + 0.5) * pa;
    const double dmin = sa < a0 ? a0 - sa : sa > a1 ? sa - a1 : 0.0;
    const double c = b.r2cull - dmin * dmin;
    if (!(c > 0.0)) return;
    // along the lines: pixels whose centre lies within the chord, one more on either side
    const double half = sqrt(c);
    const double pl = vertical ? v.pwy : v.pwx, ol = vertical ? v.y_min + b.j0 * v.pwy : v.x_min + b.i0 * v.pwx;
    const double sl = vertical ? b.sy : b.sx;
    const double lo = floor((sl - half - ol) / pl - 0.5), hi = ceil((sl + half - ol) / pl - 0.5);
    const int jlo = (int)fmax(lo, 0.0), jhi = (int)fmin(hi, (double)(len - 1));
    if (jlo > jhi) return;
    klo = jlo;
    cnt = jhi - jlo + 2;
}

__device__ __forceinline__ int units_of(unsigned long long vertices)
{
    return vertices > 1 ? (int)((vertices - 2) / 31 + 1) : 0;
}

// Per particle: the number of chunks (entry n stays 0: the exclusive scan leaves the total there) and,
// for the 2-D render, the number of lattice lines (slots of the crossing-term table).
template <typename T, bool PROJ3>
__global__ void exact_count_kernel(ExactView v, const T *x, const T *y, const T *h, int64_t n,
                                   unsigned long long *chunks, unsigned long long *lines)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n) return;
    unsigned long long total = 0, l = 0;
    LatBox b;
    int units = 0;
    if (p < n && lat_box<T, PROJ3>(v, x, y, h, p, b)) {
        if (PROJ3) {
            for (int o = 0; o < 2; o++) {
                const int nlines = (o == 0 ? b.bw : b.bh) + 1, G = (nlines + kExactGroups - 1) / kExactGroups;
                for (int g = 0; g * G < nlines; g++) {
                    int nl, klo, cnt;
                    lat_group(v, b, o == 0, G, g, nl, klo, cnt);
                    total += (unsigned long long)nl * cnt;
                }
            }
            units = units_of(total);
        } else {
            // whole lines, vertical then horizontal, each orientation cut into its own units
            units = 2 * units_of((unsigned long long)(b.bw + 1) * (b.bh + 1));
        }
        l = (unsigned long long)(b.bw + b.bh + 2);
    }
    chunks[p] = (unsigned long long)((units + kExactChunk - 1) / kExactChunk);
    if (!PROJ3) lines[p] = l;
}

template <typename T>
__global__ void __launch_bounds__(128)
exact2_lines_kernel(ExactView v, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ h,
                    int64_t n, const unsigned long long *__restrict__ line_off, double2 *__restrict__ cross)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < n; p += warps) {
        double sx, sy, sh;
        int i0, i1, j0, j1;
        if (!exact_box(v, x, y, h, p, sx, sy, sh, i0, i1, j0, j1)) continue;
        const int nv = i1 - i0 + 1, nl = nv + (j1 - j0 + 1);
        double2 *dst = cross + line_off[p];
        for (int t = lane; t < nl; t += 32) {
            const double r0 = t < nv ? sx - (v.x_min + (i0 + t) * v.pwx) : sy - (v.y_min + (j0 + t - nv) * v.pwy);
            dst[t] = ex::line_crossings(fabs(r0) / sh);
        }
    }
}

template <typename T, int NW, bool PROJ3>
__global__ void __launch_bounds__(kExactWarps * 32, 8)
exact_lattice_kernel(ExactView v, const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ h,
                     const T *__restrict__ w0, const T *__restrict__ w1, const T *__restrict__ w2,
                     const T *__restrict__ w3, int64_t n, const unsigned long long *__restrict__ offsets,
                     const unsigned long long *__restrict__ line_off, const double2 *__restrict__ cross, double *o0,
                     double *o1, double *o2, double *o3)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T *wsrc[4] = {w0, w1, w2, w3};
    double *outs[4] = {o0, o1, o2, o3};
    const unsigned long long total = offsets[n];
    const unsigned long long stride = (unsigned long long)gridDim.x * kExactWarps;
    for (unsigned long long u = (unsigned long long)blockIdx.x * kExactWarps + warp; u < total; u += stride) {
        // largest p with offsets[p] <= u: a 33-way search, the lanes probe 32 positions at once
        int64_t lo = 0, hi = n;
        while (hi - lo > 1) {
            const int64_t step = (hi - lo + 32) / 33;
            const int64_t probe = lo + (int64_t)(lane + 1) * step;
            const bool le = probe < hi && __ldg(offsets + probe) <= u;
            const int k = __popc(__ballot_sync(0xffffffffu, le));
            hi = min(hi, lo + (int64_t)(k + 1) * step);
            lo = lo + (int64_t)k * step;
        }
        const int64_t p = lo;
        LatBox b;
        if (!lat_box<T, PROJ3>(v, x, y, h, p, b)) continue;
        // 3-D: group table, lane g holds group g of either orientation; start = number of its first vertex
        const int Gv = (b.bw + kExactGroups) / kExactGroups, Gh = (b.bh + kExactGroups) / kExactGroups;
        int klo_v = 0, cnt_v = 0, klo_h = 0, cnt_h = 0, start_v = 0, start_h = 0, total_v = 0, total_all = 0;
        int units;
        if (PROJ3) {
            int nl_v, nl_h;
            lat_group(v, b, true, Gv, lane, nl_v, klo_v, cnt_v);
            lat_group(v, b, false, Gh, lane, nl_h, klo_h, cnt_h);
            start_v = nl_v * cnt_v;
            start_h = nl_h * cnt_h;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int tv = __shfl_up_sync(0xffffffffu, start_v, o), th = __shfl_up_sync(0xffffffffu, start_h, o);
                if (lane >= o) {
                    start_v += tv;
                    start_h += th;
                }
            }
            total_v = __shfl_sync(0xffffffffu, start_v, 31);
            total_all = total_v + __shfl_sync(0xffffffffu, start_h, 31);
            start_v -= nl_v * cnt_v; // exclusive
            start_h += total_v - nl_h * cnt_h;
            units = units_of((unsigned long long)total_all);
        } else {
            total_v = units_of((unsigned long long)(b.bw + 1) * (b.bh + 1)); // units per orientation
            units = 2 * total_v;
        }

        // 2-D: term * denom = (w / h^2) * (h^2 / |pwx pwy|)  (cpu_backend.py:333, :365)
        // 3-D: term * dfac  = (w / (pi h^3)) * (pi h^3 / (pwx pwy))  (cpu_backend.py:628-629)
        double wp[NW];
#pragma unroll
        for (int q = 0; q < NW; q++) wp[q] = (double)__ldg(wsrc[q] + p) * v.inv_area;
        const unsigned long long lbase = PROJ3 ? 0ull : line_off[p];
        const double two_h = 2.0 * b.sh;
        const int ul0 = (int)(u - offsets[p]) * kExactChunk, ul1 = min(ul0 + kExactChunk, units);
        for (int ul = ul0; ul < ul1; ul++) {
            bool vertical, live, has_a, has_b, needed;
            int line, k;
            if (PROJ3) {
                const int e = ul * 31 + lane;
                vertical = e < total_v;
                // the group holding vertex e: the last one whose start is <= e
                int g = 0;
#pragma unroll
                for (int st = 16; st > 0; st >>= 1) {
                    const int cand = g + st;
                    const int sv = __shfl_sync(0xffffffffu, start_v, cand), sh_ = __shfl_sync(0xffffffffu, start_h, cand);
                    if ((vertical ? sv : sh_) <= e) g = cand;
                }
                const int gs_v = __shfl_sync(0xffffffffu, start_v, g), gs_h = __shfl_sync(0xffffffffu, start_h, g);
                const int gk_v = __shfl_sync(0xffffffffu, klo_v, g), gk_h = __shfl_sync(0xffffffffu, klo_h, g);
                const int gc_v = __shfl_sync(0xffffffffu, cnt_v, g), gc_h = __shfl_sync(0xffffffffu, cnt_h, g);
                const int gstart = vertical ? gs_v : gs_h, klo = vertical ? gk_v : gk_h;
                const int cnt = max(vertical ? gc_v : gc_h, 1);
                const int f = e - gstart, fl = f / cnt;
                line = g * (vertical ? Gv : Gh) + fl;
                k = klo + (f - fl * cnt);
                live = e < total_all;
                const bool last = k == klo + cnt - 1; // the line's last kept vertex: no slot after it
                const int len = vertical ? b.bh : b.bw, nlines = (vertical ? b.bw : b.bh) + 1;
                // the four pixels around the vertex: a / b = after / before the line, rows k and k - 1 (centre
                // offsets as functions of the pixel index alone: the two lanes that share a slot must agree)
                const double pa = vertical ? v.pwx : v.pwy, pl = vertical ? v.pwy : v.pwx;
                const double oa = vertical ? v.x_min - b.sx : v.y_min - b.sy, ol = vertical ? v.y_min - b.sy : v.x_min - b.sx;
                const int ca = (vertical ? b.i0 : b.j0) + line, cl = (vertical ? b.j0 : b.i0) + k;
                const double xa = fma(ca + 0.5, pa, oa), xb = fma(ca - 0.5, pa, oa);
                const double y1 = fma(cl + 0.5, pl, ol), y0 = fma(cl - 0.5, pl, ol);
                const double xa2 = xa * xa, xb2 = xb * xb, y12 = y1 * y1, y02 = y0 * y0;
                const bool col_a = live && line < nlines - 1, col_b = live && line > 0;
                has_a = col_a && k < len && !last && xa2 + y12 < b.r2cull;
                has_b = col_b && k < len && !last && xb2 + y12 < b.r2cull;
                // a vertex is evaluated only if one of the two slots that meet there feeds a pixel inside the disc
                needed = has_a || has_b ||
                         (k > klo && ((col_a && xa2 + y02 < b.r2cull) || (col_b && xb2 + y02 < b.r2cull)));
            } else {
                vertical = ul < total_v; // first the vertical lines, then the horizontal ones
                const int len = vertical ? b.bh : b.bw, nlines = (vertical ? b.bw : b.bh) + 1;
                const int e = (vertical ? ul : ul - total_v) * 31 + lane;
                line = e / (len + 1);
                k = e - line * (len + 1);
                live = needed = line < nlines;
                has_a = live && k < len && line < nlines - 1;
                has_b = live && k < len && line > 0;
            }
            const bool mine = has_a || has_b;
            // signed distance of the particle from the line, offset of the vertex along it
            const double r0 = vertical ? b.sx - (v.x_min + (b.i0 + line) * v.pwx) : b.sy - (v.y_min + (b.j0 + line) * v.pwy);
            const double d = vertical ? (v.y_min + (b.j0 + k) * v.pwy) - b.sy : (v.x_min + (b.i0 + k) * v.pwx) - b.sx;
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            if (needed && r0 != 0.0) {
                const double ar0 = fabs(r0);
                if (PROJ3) {
                    t0 = ex::los_edge(r0, d, two_h, b.sh);       // side face: edge along the line of sight
                    t1 = ex::half_outside(ar0, two_h, d, b.sh);  // side face: its top edge
                    t2 = ex::half_outside(two_h, ar0, d, b.sh);  // top face: the edge on this line
                } else {
                    const double2 c = __ldg(cross + lbase + (vertical ? line : b.bw + 1 + line));
                    t0 = ex::half_segment(fabs(d) / ar0, ar0 / b.sh, c.x, c.y); // same q0 as the lines kernel
                }
            }
            const double n0 = __shfl_down_sync(0xffffffffu, t0, 1), d_next = __shfl_down_sync(0xffffffffu, d, 1);
            double n1 = 0.0, n2 = 0.0;
            if (PROJ3) {
                n1 = __shfl_down_sync(0xffffffffu, t1, 1);
                n2 = __shfl_down_sync(0xffffffffu, t2, 1);
            }
            if (!mine || lane == 31 || r0 == 0.0) continue;
            const double sign = r0 > 0.0 ? 1.0 : -1.0;
            // 3-D: side face + top and bottom face edges on this slot (same two pixels, same signs)
            const double val = PROJ3 ? n0 - t0 + 2.0 * sign * (ex::combine_halves(t1, n1, d, d_next) +
                                                               ex::combine_halves(t2, n2, d, d_next))
                                     : sign * ex::combine_halves(t0, n0, d, d_next);
            if (val == 0.0) continue;
            // the pixel after the line gains the integral, the pixel before it loses it (cpu_backend.py:404-440)
            const int ia = vertical ? b.i0 + line : b.i0 + k, ja = vertical ? b.j0 + k : b.j0 + line;
            const int ib = vertical ? ia - 1 : ia, jb = vertical ? ja : ja - 1;
#pragma unroll
            for (int q = 0; q < NW; q++) {
                const double wv = wp[q] * val;
                if (has_a) {
                    SPHB_ASSERT(ia >= 0 && ia < v.nx && ja >= 0 && ja < v.ny);
                    red_add(outs[q] + (size_t)ja * v.nx + ia, wv);
                }
                if (has_b) {
                    SPHB_ASSERT(ib >= 0 && ib < v.nx && jb >= 0 && jb < v.ny);
                    red_add(outs[q] + (size_t)jb * v.nx + ib, -wv);
                }
            }
        }
    }
}

template <typename T, int NW, bool PROJ3>
static int launch_lattice(const ExactView &v, const sphb_particles *p, const unsigned long long *offsets,
                          const unsigned long long *line_off, const double2 *cross, double *const *out,
                          cudaStream_t stream)
{
    int sms = 0;
    SPHB_CUDA(sm_count(&sms));
    const T *w[4] = {nullptr, nullptr, nullptr, nullptr};
    double *o[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < NW; i++) {
        w[i] = static_cast<const T *>(p->w[i]);
        o[i] = out[i];
    }
    exact_lattice_kernel<T, NW, PROJ3><<<sms * 32, kExactWarps * 32, 0, stream>>>(
        v, static_cast<const T *>(p->x), static_cast<const T *>(p->y), static_cast<const T *>(p->h), w[0], w[1], w[2],
        w[3], p->n, offsets, line_off, cross, o[0], o[1], o[2], o[3]);
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}

template <typename T, bool PROJ3>
static int launch_lattice_nw(const ExactView &v, const sphb_particles *p, const unsigned long long *offsets,
                             const unsigned long long *line_off, const double2 *cross, double *const *out,
                             cudaStream_t stream)
{
    switch (p->n_weights) {
    case 1: return launch_lattice<T, 1, PROJ3>(v, p, offsets, line_off, cross, out, stream);
    case 2: return launch_lattice<T, 2, PROJ3>(v, p, offsets, line_off, cross, out, stream);
    case 3: return launch_lattice<T, 3, PROJ3>(v, p, offsets, line_off, cross, out, stream);
    default: return launch_lattice<T, 4, PROJ3>(v, p, offsets, line_off, cross, out, stream);
    }
}

// Page-locked landing zone of the 2-D line total, one per host thread and device.
static cudaError_t exact_landing(unsigned long long **out)
{
    static thread_local unsigned long long *cache[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidValue;
    if (!cache[dev]) {
        e = cudaHostAlloc(reinterpret_cast<void **>(&cache[dev]), sizeof(unsigned long long), cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
    }
    *out = cache[dev];
    return cudaSuccess;
}

template <typename T>
static int exact_typed(const ExactView &v, const sphb_particles *p, bool proj3d, double *const *out, void *ws,
                       size_t ws_bytes, size_t *ws_needed, cudaStream_t stream)
{
    const int64_t n = p->n;
    size_t scan_tmp = 0;
    SPHB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (unsigned long long *)nullptr,
                                            (unsigned long long *)nullptr, n + 1, stream));
    const size_t off_bytes = align_up(sizeof(unsigned long long) * (size_t)(n + 1), 256);
    // 3-D projection: unit offsets + scan scratch.  2-D render: also the per-particle line offsets, and
    // (sized after the count, below) the crossing terms of every lattice line.
    const size_t need = (proj3d ? 1 : 2) * off_bytes + align_up(scan_tmp, 256);
    if (ws_needed) *ws_needed = need;
    if (!ws || ws_bytes < need) {
        set_error("workspace of %zu bytes needed, %zu given", need, ws_bytes);
        return SPHB_ERR_WORKSPACE;
    }
    unsigned long long *offsets = static_cast<unsigned long long *>(ws);
    const T *x = static_cast<const T *>(p->x), *y = static_cast<const T *>(p->y), *h = static_cast<const T *>(p->h);
    if (proj3d) {
        void *scan_ws = static_cast<unsigned char *>(ws) + off_bytes;
        exact_count_kernel<T, true><<<div_up(n + 1, 256), 256, 0, stream>>>(v, x, y, h, n, offsets, nullptr);
        SPHB_LAUNCH_CHECK();
        SPHB_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan_tmp, offsets, offsets, n + 1, stream));
        return launch_lattice_nw<T, true>(v, p, offsets, nullptr, nullptr, out, stream);
    }
    unsigned long long *line_off = reinterpret_cast<unsigned long long *>(static_cast<unsigned char *>(ws) + off_bytes);
    void *scan_ws = static_cast<unsigned char *>(ws) + 2 * off_bytes;
    exact_count_kernel<T, false><<<div_up(n + 1, 256), 256, 0, stream>>>(v, x, y, h, n, offsets, line_off);
    SPHB_LAUNCH_CHECK();
    SPHB_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan_tmp, offsets, offsets, n + 1, stream));
    SPHB_CUDA(cub::DeviceScan::ExclusiveSum(scan_ws, scan_tmp, line_off, line_off, n + 1, stream));
    // the one synchronisation of the call: the number of lattice lines sizes the table of crossing terms
    unsigned long long *total_lines = nullptr;
    SPHB_CUDA(exact_landing(&total_lines));
    SPHB_CUDA(cudaMemcpyAsync(total_lines, line_off + n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
    SPHB_CUDA(cudaStreamSynchronize(stream));
    const size_t cross_off = align_up(need, 256);
    const size_t need2 = cross_off + align_up(sizeof(double2) * (size_t)*total_lines, 256);
    if (ws_needed) *ws_needed = need2;
    if (ws_bytes < need2) {
        set_error("workspace of %zu bytes needed for %llu lattice lines, %zu given", need2, *total_lines, ws_bytes);
        return SPHB_ERR_WORKSPACE;
    }
    if (*total_lines == 0) return SPHB_OK;
    double2 *cross = reinterpret_cast<double2 *>(static_cast<unsigned char *>(ws) + cross_off);
    int sms = 0;
    SPHB_CUDA(sm_count(&sms));
    exact2_lines_kernel<T><<<(int)std::min<int64_t>(div_up(n, 4), (int64_t)sms * 16), 128, 0, stream>>>(v, x, y, h, n,
                                                                                                     line_off, cross);
    SPHB_LAUNCH_CHECK();
    return launch_lattice_nw<T, false>(v, p, offsets, line_off, cross, out, stream);
}

static int exact_impl(const sphb_particles *p, const sphb_raster *r, bool proj3d, int accumulate,
                      double *const *out, void *ws, size_t ws_bytes, size_t *ws_needed, cudaStream_t stream)
{
    if (!p || !r || !out || p->n < 0 || p->n_weights < 1 || p->n_weights > SPHB_MAX_WEIGHTS ||
        (p->dtype != SPHB_F32 && p->dtype != SPHB_F64) || r->nx <= 0 || r->ny <= 0 ||
        !(r->x_max > r->x_min) || !(r->y_max > r->y_min)) {
        set_error("invalid argument to exact render");
        return SPHB_ERR_ARGUMENT;
    }
    // the work items of one particle (edges / faces of its pixel box, at most ~3 per pixel of the
    // raster) are indexed with 32-bit integers
    if ((int64_t)(r->nx + 1) * (int64_t)(r->ny + 1) > (1LL << 28)) {
        set_error("raster of %d x %d pixels is too large for the exact paths (limit 2^28 pixels)", r->nx, r->ny);
        return SPHB_ERR_ARGUMENT;
    }
    for (int i = 0; i < p->n_weights; i++)
        if (!out[i] || (p->n > 0 && !p->w[i])) {
            set_error("missing weight or output %d", i);
            return SPHB_ERR_ARGUMENT;
        }
    if (p->n > 0 && (!p->x || !p->y || !p->h)) {
        set_error("missing particle column");
        return SPHB_ERR_ARGUMENT;
    }
    if (ws_needed) *ws_needed = 0;
    if (!accumulate)
        for (int i = 0; i < p->n_weights; i++)
            SPHB_CUDA(cudaMemsetAsync(out[i], 0, sizeof(double) * (size_t)r->nx * r->ny, stream));
    if (p->n == 0) return SPHB_OK;
    ExactView v;
    v.nx = r->nx;
    v.ny = r->ny;
    v.x_min = r->x_min;
    v.y_min = r->y_min;
    v.pwx = (r->x_max - r->x_min) / r->nx;
    v.pwy = (r->y_max - r->y_min) / r->ny;
    v.inv_area = 1.0 / (v.pwx * v.pwy);
    if (p->dtype == SPHB_F32) return exact_typed<float>(v, p, proj3d, out, ws, ws_bytes, ws_needed, stream);
    return exact_typed<double>(v, p, proj3d, out, ws, ws_bytes, ws_needed, stream);
}

} // namespace sphb

using namespace sphb;

extern "C" int sphb_exact2d(const sphb_particles *p, const sphb_raster *r, int accumulate, double *const *out,
                            void *ws, size_t ws_bytes, size_t *ws_needed, sphb_stream stream)
{
    return exact_impl(p, r, false, accumulate, out, ws, ws_bytes, ws_needed, static_cast<cudaStream_t>(stream));
}

extern "C" int sphb_exact3d_proj(const sphb_particles *p, const sphb_raster *r, int accumulate,
                                 double *const *out, void *ws, size_t ws_bytes, size_t *ws_needed,
                                 sphb_stream stream)
{
    return exact_impl(p, r, true, accumulate, out, ws, ws_bytes, ws_needed, static_cast<cudaStream_t>(stream));
}
