// Shared device math and host helpers for libsphb200 (sm_100a).
//
// Kernel functions follow their mathematical definition in
// sarracen/kernels/{cubic,quartic,quintic}_spline.py and the table lookup of
// sarracen/kernels/base_kernel.py:89-97.  The normalisation constant is NOT
// applied here: it is folded into the per-particle term w / h^ndim when a
// particle is staged, so the per-contribution work is polynomial only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sphb200.h"

namespace sphb {

void set_error(const char *fmt, ...);

// SM count of the current device (cached per device; sphb_scatter.cu).
cudaError_t sm_count(int *out);

#define SPHB_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            sphb::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call,                  \
                            cudaGetErrorString(e_));                                     \
            return SPHB_ERR_CUDA;                                                        \
        }                                                                                \
    } while (0)

#define SPHB_LAUNCH_CHECK() SPHB_CUDA(cudaGetLastError())

// Index checks for debug builds (python -m sarracen_b200.build --debug-bounds builds
// lib/libsphb200_dbg.so; run the GPU tests against it with SPHB_LIB_OVERRIDE).  compute-sanitizer
// is not available on the target pool, so every shared-memory / raster / table index the kernels
// form is asserted in that build instead.
#ifdef SPHB_DEBUG_BOUNDS
#include <cassert>
#define SPHB_ASSERT(cond) assert(cond)
#else
#define SPHB_ASSERT(cond) ((void)0)
#endif

constexpr double kPi = 3.14159265358979323846;

// sigma_d of each kernel (cubic_spline.py:15-17, quartic_spline.py:15-17,
// quintic_spline.py:15-17); 1 for the already normalised column table.
__host__ __device__ inline double kernel_norm(int kind, int ndim)
{
    switch (kind) {
    case SPHB_CUBIC: return ndim == 1 ? 2.0 / 3.0 : ndim == 2 ? 10.0 / (7.0 * kPi) : 1.0 / kPi;
    case SPHB_QUARTIC:
        return ndim == 1 ? 1.0 / 24.0 : ndim == 2 ? 96.0 / (1199.0 * kPi) : 1.0 / (20.0 * kPi);
    case SPHB_QUINTIC:
        return ndim == 1 ? 1.0 / 120.0 : ndim == 2 ? 7.0 / (478.0 * kPi) : 1.0 / (120.0 * kPi);
    default: return 1.0;
    }
}

__host__ __device__ inline double kernel_default_radius(int kind)
{
    return kind == SPHB_CUBIC ? 2.0 : kind == SPHB_QUARTIC ? 2.5 : 3.0;
}

template <typename R> struct Real2;
template <> struct Real2<float> { using type = float2; };
template <> struct Real2<double> { using type = double2; };

// Column table as the device sees it: entry i holds (T[i], T[i+1] - T[i]) so
// one vector load serves the lerp T[i] + f (T[i+1] - T[i]).
template <typename R> struct LutView {
    // 32-bit shared-memory address of the table minus lut_bias() entries: the (biased) index the
    // lookup computes is scaled and added to it directly (unsigned arithmetic wraps, so the sum is
    // the address of the entry)
    unsigned tab_biased;
    const typename Real2<R>::type *gtab; // global memory instead (tables too large for shared memory)
    R scale;                             // (S - 1) / radius
    int last;                            // S - 1
};

// Internal kernel id: column table read from global memory (tables too large for shared
// memory).  A separate template value, not a run-time test, so the common shared-memory
// lookup stays a single LDS.
constexpr int SPHB_COLUMN_LUT_GLOBAL = 4;

template <int KIND> __device__ __host__ constexpr bool is_lut() { return KIND == SPHB_COLUMN_LUT || KIND == SPHB_COLUMN_LUT_GLOBAL; }

__device__ __forceinline__ double2 lds_pair(unsigned addr, double)
{
    double2 e;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(e.x), "=d"(e.y) : "r"(addr));
    return e;
}
__device__ __forceinline__ float2 lds_pair(unsigned addr, float)
{
    float2 e;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(e.x), "=f"(e.y) : "r"(addr));
    return e;
}

// Index bias of lut_index(): see there.
__device__ __forceinline__ constexpr int lut_bias(double) { return 0; }
__device__ __forceinline__ constexpr int lut_bias(float) { return 0x4B400000; }

template <typename R> __device__ __forceinline__ void lut_set_table(LutView<R> &lut, const void *shared_table)
{
    lut.tab_biased = (unsigned)__cvta_generic_to_shared(shared_table) -
                     (unsigned)lut_bias(R(0)) * (unsigned)sizeof(typename Real2<R>::type);
    // Opaque to ptxas (a shuffle result cannot be constant-folded): otherwise it re-derives the
    // address from the shared-memory window at every lookup and pays an extra add for the bias,
    // which does not fit an LDS immediate.  Called once per kernel by all threads.
    lut.tab_biased = __shfl_sync(0xffffffffu, lut.tab_biased, 0);
}

// Entry of the table for a BIASED index (already clamped to last + bias).
template <int KIND, typename R>
__device__ __forceinline__ typename Real2<R>::type lut_entry(const LutView<R> &lut, int ib)
{
    SPHB_ASSERT(ib - lut_bias(R(0)) >= 0 && ib - lut_bias(R(0)) <= lut.last);
    if (KIND == SPHB_COLUMN_LUT_GLOBAL) return __ldg(lut.gtab + (ib - lut_bias(R(0))));
    return lds_pair(lut.tab_biased + (unsigned)ib * (unsigned)sizeof(typename Real2<R>::type), R(0));
}

// ---- square root -----------------------------------------------------------
// double: MUFU.RSQ64H seed (reads the high word only: relative error < 2^-19.9)
// followed by one coupled Goldschmidt step -> relative error < 1.5 * 2^-39.8
// ~ 1.6e-12, far inside the 1e-10 parity budget (it moves W by < |W'| q 1.6e-12).
// Arguments must be > 0: callers add sqrt_guard() once per particle so that a == 0 cannot
// reach rsqrt.
__device__ __forceinline__ double sqrt_guard(double) { return 1e-290; }
__device__ __forceinline__ float sqrt_guard(float) { return 1e-35f; }

__device__ __forceinline__ double sqrt_pos(double a)
{
    // a > 0 is the caller's job: every call site forms q^2 as fma(d * d, s, base) with
    // base >= sqrt_guard() > 0, so no floor is needed here.
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    // y / 2 by an integer decrement of the exponent (y is a normal number): one FP64-pipe
    // instruction less per square root
    const double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
    const double g = a * y;
    const double r = fma(-g, h, 0.5);
    return fma(g, r, g);
}
__device__ __forceinline__ float sqrt_pos(float a)
{
    // one MUFU.RSQ: the argument is a normal number (callers add sqrt_guard), so the
    // denormal rescaling that rsqrtf() wraps around it is not needed
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));
    return a * y;
}

// ---- unnormalised kernel shapes, valid for 0 <= q < radius ----------------
template <int KIND, typename R> __device__ __forceinline__ R kernel_shape(R q)
{
    if (KIND == SPHB_CUBIC) {
        // (2 - q)^3 / 4 as (c (2 - q))^3 with c = 4^(-1/3): one multiply less
        const R a = fma(q, R(-0.62996052494743658), R(1.25992104989487316));
        const R w = a * a * a;
        const R w1 = fma(q * q, fma(R(0.75), q, R(-1.5)), R(1));
        return q < R(1) ? w1 : w;
    } else if (KIND == SPHB_QUARTIC) {
        R a = R(2.5) - q, b = R(1.5) - q, c = R(0.5) - q;
        R a2 = a * a, b2 = b * b, c2 = c * c;
        R w = a2 * a2;
        if (q < R(1.5)) w = fma(R(-5), b2 * b2, w);
        if (q < R(0.5)) w = fma(R(10), c2 * c2, w);
        return w;
    } else {
        R a = R(3) - q, b = R(2) - q, c = R(1) - q;
        R a2 = a * a, b2 = b * b, c2 = c * c;
        R w = a2 * a2 * a;
        if (q < R(2)) w = fma(R(-6), b2 * b2 * b, w);
        if (q < R(1)) w = fma(R(15), c2 * c2 * c, w);
        return w;
    }
}

// a < b for POSITIVE finite numbers, off the FP64 pipe: positive doubles order like their bit patterns, so the
// compare is two integer instructions on the ALU pipe instead of a DSETP on the (half-rate, contended) FP64 pipe.
__device__ __forceinline__ bool lt_pos(double a, double b)
{
    return __double_as_longlong(a) < __double_as_longlong(b);
}
__device__ __forceinline__ bool lt_pos(float a, float b) { return a < b; }

// The same from q and q^2 when the caller has both (it formed q^2 first): the cubic's inner branch is
// 1 + q^2 (0.75 q - 1.5) -- the exact q^2 instead of the rounded square of its root, one multiply less.
template <int KIND, typename R> __device__ __forceinline__ R kernel_shape2(R q, R q2)
{
    if (KIND == SPHB_CUBIC) {
        const R a = fma(q, R(-0.62996052494743658), R(1.25992104989487316));
        const R w = a * a * a;
        const R w1 = fma(q2, fma(R(0.75), q, R(-1.5)), R(1));
        return lt_pos(q2, R(1)) ? w1 : w;
    }
    return kernel_shape<KIND>(q);
}

// Column-table lookup.  The reference evaluates T[i] + f (T[i+1] - T[i]) with i = floor(t),
// f = t - i, t = q (S - 1) / R (sarracen/kernels/base_kernel.py:88-110 via np.interp).  On each
// segment that is the straight line A_i + B_i q with B_i = (T[i+1] - T[i]) (S - 1) / R and
// A_i = T[i] - i (T[i+1] - T[i]); the table holds (A_i, B_i), so a lookup is one fused
// multiply-add for the index -- q * scale + 1.5 * 2^52 rounded DOWN leaves floor(t) in the low
// mantissa bits -- and one for the value.  Values agree with the reference's form to a few ulp
// of max|T| (A_i carries the cancellation i dT_i ~ q T'(q)).
__device__ __forceinline__ int lut_index(double q, double scale)
{
    return __double2loint(__fma_rd(q, scale, 6755399441055744.0));
}
// float: the bit pattern of 1.5 * 2^23 + i is 0x4B400000 + i (i < 2^22).  The bias is not masked off:
// it is carried through the clamp (compare against last + bias) and folds into the table address.
__device__ __forceinline__ int lut_index(float q, float scale)
{
    return __float_as_int(__fmaf_rd(q, scale, 12582912.0f));
}
// Table entry i from T (double, S samples): shared by every kernel that builds the table.
template <typename R>
__device__ __forceinline__ typename Real2<R>::type lut_make_entry(const double *lut, int samples, int i, double scale)
{
    typename Real2<R>::type e;
    if (i >= samples - 1) {
        // the clamp target of out-of-support lookups: W(q >= R) = 0
        e.x = e.y = R(0);
        return e;
    }
    const double a = lut[i], d = lut[i + 1] - a;
    e.x = (R)(a - (double)i * d);
    e.y = (R)(d * scale);
    return e;
}
template <int KIND, typename R>
__device__ __forceinline__ R lut_eval(R q, const LutView<R> &lut)
{
    int i = lut_index(q, lut.scale);              // biased, see lut_bias()
    i = min(i, lut.last + lut_bias(R(0)));
    typename Real2<R>::type e = lut_entry<KIND>(lut, i);
    return fma(e.y, q, e.x);
}

// W(q) from 0 < q^2 < radius^2 (the caller has tested the range and added
// sqrt_guard()).
template <int KIND, typename R>
__device__ __forceinline__ R kernel_eval_pos(R q2, const LutView<R> &lut)
{
    R q = sqrt_pos(q2);
    if (is_lut<KIND>()) {
        return lut_eval<KIND, R>(q, lut);
    } else {
        return kernel_shape2<KIND>(q, q2);
    }
}

// W(q) from q^2 > 0 with no caller-side range test: zero for q >= radius.  The
// table path needs no test at all -- the index clamps to the last entry, which
// is (0, 0) because every kernel vanishes at its radius.
template <int KIND, typename R>
__device__ __forceinline__ R kernel_eval_unguarded(R q2, R rad2, const LutView<R> &lut)
{
    R q = sqrt_pos(q2);
    if (is_lut<KIND>()) {
        return lut_eval<KIND, R>(q, lut);
    } else {
        R w = kernel_shape2<KIND>(q, q2);
        return lt_pos(q2, rad2) ? w : R(0);
    }
}

// W(q) for arbitrary q (line cuts: no q <= R guard in the reference,
// cpu_backend.py:527-528, :598-599): zero outside [0, R).
template <int KIND>
__device__ __forceinline__ double kernel_eval_any(double q, double radius, const double *lut, int samples)
{
    if (KIND == SPHB_COLUMN_LUT) {
        double t = q * (samples - 1) / radius;
        double fl = floor(t);
        int i0 = (int)fmin(fmax(fl, 0.0), (double)(samples - 1));
        int i1 = (int)fmin(fmax(ceil(t), 0.0), (double)(samples - 1));
        double f = t - (double)i0;
        return fma(f, lut[i1] - lut[i0], lut[i0]);
    } else {
        if (!(q >= 0.0) || q >= radius) return 0.0;
        return kernel_shape<KIND>(q);
    }
}

// small exact int -> real without the conversion pipe
__device__ __forceinline__ double int_to_real(int i, double)
{
    return __hiloint2double(0x43300000, i ^ 0x80000000) - 4503601774854144.0; // 2^52 + 2^31
}
__device__ __forceinline__ float int_to_real(int i, float)
{
    return __int_as_float(0x4B400000 + i) - 12582912.0f; // 1.5 * 2^23, |i| < 2^22
}

__device__ __forceinline__ void red_add(double *addr, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

} // namespace sphb
