// Radial binning of particles on the device: the segmented histogram behind
// Sarracen's disc analysis -- surface_density, azimuthal_average, angular_momentum,
// scale_height, honH (sarracen/disc/surface_density.py:10-459), which bin particles into rings
// with pandas.cut + groupby (sarracen/disc/utils.py:25-86).
//
//   radial_index_kernel  r = |(x, y[, z]) - origin| per particle, ring i such that
//                        edges[i] < r <= edges[i + 1] (pandas.cut: right-closed intervals, values
//                        outside (edges[0], edges[bins]] belong to no ring: index -1), found by
//                        binary search on the caller's edge array -- the same comparisons pandas
//                        makes, so the assignment is identical; and the ring populations.
//   bin_sum_kernel       sums of up to four value columns per ring: every CTA accumulates into a
//                        private copy of the rings in shared memory and adds it to the result once
//                        at the end (rings beyond the shared-memory budget: global reductions).
#include <algorithm>

#include "sphb_common.cuh"

namespace sphb {

template <typename T>
__global__ void __launch_bounds__(256)
radial_index_kernel(const T *__restrict__ x, const T *__restrict__ y, const T *__restrict__ z, int64_t n,
                    double ox, double oy, double oz, int spherical, const double *__restrict__ edges, int bins,
                    int *__restrict__ index, unsigned long long *__restrict__ counts)
{
    extern __shared__ unsigned long long cnt_s[];
    for (int i = threadIdx.x; i < bins; i += blockDim.x) cnt_s[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double dx = (double)__ldg(x + i) - ox, dy = (double)__ldg(y + i) - oy;
        // no contraction: the radius has to round like NumPy's (x-ox)**2 + (y-oy)**2 [+ (z-oz)**2]
        double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (spherical) {
            const double dz = (double)__ldg(z + i) - oz;
            r2 = __dadd_rn(r2, __dmul_rn(dz, dz));
        }
        const double r = sqrt(r2);
        // first edge >= r (numpy.searchsorted(edges, r, 'left')); ring = that - 1
        int lo = 0, hi = bins + 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(edges + mid) < r) lo = mid + 1; else hi = mid;
        }
        const int ring = (lo >= 1 && lo <= bins) ? lo - 1 : -1;   // NaN compares false everywhere: lo = 0
        index[i] = ring;
        if (ring >= 0) atomicAdd(&cnt_s[ring], 1ull);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x)
        if (cnt_s[i]) atomicAdd(counts + i, cnt_s[i]);
}

template <typename T, int NV>
__global__ void __launch_bounds__(256)
bin_sum_kernel(const int *__restrict__ index, const T *v0, const T *v1, const T *v2, const T *v3, int64_t n, int bins,
               int use_smem, double *o0, double *o1, double *o2, double *o3)
{
    extern __shared__ double sum_s[];   // [NV][bins]
    const T *vals[4] = {v0, v1, v2, v3};
    double *outs[4] = {o0, o1, o2, o3};
    if (use_smem) {
        for (int i = threadIdx.x; i < NV * bins; i += blockDim.x) sum_s[i] = 0.0;
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int ring = __ldg(index + i);
        if (ring < 0) continue;
        SPHB_ASSERT(ring < bins);
#pragma unroll
        for (int k = 0; k < NV; k++) {
            const double v = (double)__ldg(vals[k] + i);
            if (use_smem)
                atomicAdd(&sum_s[k * bins + ring], v);
            else
                red_add(outs[k] + ring, v);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < NV * bins; i += blockDim.x)
            if (sum_s[i] != 0.0) red_add(outs[i / bins] + (i % bins), sum_s[i]);
    }
}

template <typename T>
static int radial_index(const void *x, const void *y, const void *z, int64_t n, const double *origin, int spherical,
                        const double *edges, int bins, int *index, unsigned long long *counts, cudaStream_t stream)
{
    int sms = 0;
    SPHB_CUDA(sm_count(&sms));
    const int grid = (int)std::min<int64_t>(div_up(std::max<int64_t>(n, 1), 256), (int64_t)sms * 8);
    SPHB_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * bins, stream));
    radial_index_kernel<T><<<grid, 256, sizeof(unsigned long long) * bins, stream>>>(
        static_cast<const T *>(x), static_cast<const T *>(y), static_cast<const T *>(z), n, origin[0], origin[1],
        origin[2], spherical, edges, bins, index, counts);
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}

template <typename T, int NV>
static int bin_sums(const int *index, const void *const *values, int64_t n, int bins, double *const *out,
                    cudaStream_t stream)
{
    int sms = 0;
    SPHB_CUDA(sm_count(&sms));
    const size_t smem = sizeof(double) * (size_t)NV * bins;
    const int use_smem = smem <= 96 * 1024;
    auto kern = bin_sum_kernel<T, NV>;
    if (use_smem && smem > 48 * 1024)
        SPHB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)std::min<int64_t>(div_up(std::max<int64_t>(n, 1), 256), (int64_t)sms * 4);
    const T *v[4] = {nullptr, nullptr, nullptr, nullptr};
    double *o[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < NV; k++) {
        v[k] = static_cast<const T *>(values[k]);
        o[k] = out[k];
        SPHB_CUDA(cudaMemsetAsync(out[k], 0, sizeof(double) * bins, stream));
    }
    kern<<<grid, 256, use_smem ? smem : 0, stream>>>(index, v[0], v[1], v[2], v[3], n, bins, use_smem, o[0], o[1], o[2],
                                                     o[3]);
    SPHB_LAUNCH_CHECK();
    return SPHB_OK;
}

} // namespace sphb

using namespace sphb;

extern "C" int sphb_radial_index(int dtype, const void *x, const void *y, const void *z, int64_t n,
                                 const double *origin, int spherical, const double *edges, int bins, int *index,
                                 unsigned long long *counts, sphb_stream stream)
{
    if (n < 0 || bins < 1 || bins > 6000 || !origin || !edges || !index || !counts ||
        (n > 0 && (!x || !y || (spherical && !z))) || (dtype != SPHB_F32 && dtype != SPHB_F64)) {
        set_error("invalid argument to radial_index (1 <= bins <= 6000)");
        return SPHB_ERR_ARGUMENT;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return dtype == SPHB_F32 ? radial_index<float>(x, y, z, n, origin, spherical, edges, bins, index, counts, s)
                             : radial_index<double>(x, y, z, n, origin, spherical, edges, bins, index, counts, s);
}

extern "C" int sphb_bin_sums(int dtype, const int *index, int n_values, const void *const *values, int64_t n, int bins,
                             double *const *out, sphb_stream stream)
{
    if (n < 0 || bins < 1 || n_values < 1 || n_values > 4 || !index || !values || !out ||
        (dtype != SPHB_F32 && dtype != SPHB_F64)) {
        set_error("invalid argument to bin_sums");
        return SPHB_ERR_ARGUMENT;
    }
    for (int k = 0; k < n_values; k++)
        if (!out[k] || (n > 0 && !values[k])) {
            set_error("missing value column or output %d", k);
            return SPHB_ERR_ARGUMENT;
        }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define SPHB_BINS_CASE(T)                                                    \
    switch (n_values) {                                                      \
    case 1: return bin_sums<T, 1>(index, values, n, bins, out, s);           \
    case 2: return bin_sums<T, 2>(index, values, n, bins, out, s);           \
    case 3: return bin_sums<T, 3>(index, values, n, bins, out, s);           \
    default: return bin_sums<T, 4>(index, values, n, bins, out, s);          \
    }
    if (dtype == SPHB_F32) {
        SPHB_BINS_CASE(float)
    }
    SPHB_BINS_CASE(double)
#undef SPHB_BINS_CASE
}
