// Confusion-matrix counts of a simulated flood raster against a reference water mask: Stats._set_values and
// Stats._totals of the reference (stats.py:21-25, :63-86).  One streaming pass over both
// rasters, six counters; the arithmetic is the reference's NumPy arithmetic LITERALLY, per sample type of the mask:
//   mask = file > 0.0001                              (compared in the raster's own type: a Python float is a weak scalar)
//   TP   = count_nonzero(mask * ndwi)                 FN = count_nonzero(((ndwi - mask) > 0) * 1)
//   FP   = count_nonzero(((mask - ndwi) > 0) * 1)     TN = count_nonzero((1 - mask) * (1 - ndwi))
//   total positives = count_nonzero(ndwi)             total negatives = count_nonzero(1 - ndwi)
// -- including what that arithmetic does outside {0, 1}: uint8 differences wrap (0 - 1 = 255 > 0), 0 * NaN is NaN and
// counts as non-zero.  HBM bound: 5 - 8 B/cell read, nothing written.
#include "common.cuh"

namespace {

template <typename S> __device__ __forceinline__ bool wet(S v, double thr);
template <> __device__ __forceinline__ bool wet<float>(float v, double thr) { return v > (float)thr; }
template <> __device__ __forceinline__ bool wet<double>(double v, double thr) { return v > thr; }

template <typename S, typename T>
__global__ void __launch_bounds__(256) confusion_kernel(const S* __restrict__ sim, int64_t sim_pitch,
                                                        const T* __restrict__ truth, int64_t truth_pitch, int64_t ny,
                                                        int64_t nx, double thr, unsigned long long* __restrict__ counts)
{
    __shared__ unsigned int block[6];
    if (threadIdx.x < 6) block[threadIdx.x] = 0u;
    __syncthreads();
    unsigned int c[6] = {0u, 0u, 0u, 0u, 0u, 0u};
    for (CellIter it(nx); it.y < ny; it.next()) {
        const S v = sim[it.y * sim_pitch + it.x];
        const T t = truth[it.y * truth_pitch + it.x];
        const bool m = wet<S>(v, thr);
        const T mt = (T)(m ? 1 : 0);
        const T tc = (T)((T)1 - t);                        // ndwi_complement                      stats.py:66
        c[0] += ((T)(mt * t) != (T)0);                     // TP                                   :69-70
        c[1] += ((T)(t - mt) > (T)0);                      // FN                                   :72-73
        c[2] += ((T)(mt - t) > (T)0);                      // FP                                   :78-79
        c[3] += ((T)((T)(m ? 0 : 1) * tc) != (T)0);        // TN                                   :81-83
        c[4] += (t != (T)0);                               // total positives                      :22
        c[5] += (tc != (T)0);                              // total negatives                      :23-24
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const unsigned int w = __reduce_add_sync(0xffffffffu, c[k]);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&block[k], w);
    }
    __syncthreads();
    if (threadIdx.x < 6 && block[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)block[threadIdx.x]);
}

template <typename S>
int launch(const void* sim, int64_t sim_pitch, const void* truth, int truth_dtype, int64_t truth_pitch, int64_t ny,
           int64_t nx, double thr, unsigned long long* counts, cudaStream_t s)
{
    const int64_t total = ny * nx;
    // at most 2^32 - 1 cells per thread and per block counter: 148 x 16 blocks of 256 threads cover 1.5e15 cells
    const int64_t want = (total + 255) / 256;
    const int blocks = (int)(want < (int64_t)hd_num_sms() * 16 ? want : (int64_t)hd_num_sms() * 16);
    hd_prof_begin("confusion_kernel", s);
    if (truth_dtype == HD_U8)
        confusion_kernel<S, uint8_t><<<blocks, 256, 0, s>>>((const S*)sim, sim_pitch, (const uint8_t*)truth, truth_pitch, ny, nx,
                                                           thr, counts);
    else if (truth_dtype == HD_I16)
        confusion_kernel<S, int16_t><<<blocks, 256, 0, s>>>((const S*)sim, sim_pitch, (const int16_t*)truth, truth_pitch, ny, nx,
                                                           thr, counts);
    else if (truth_dtype == HD_F32)
        confusion_kernel<S, float><<<blocks, 256, 0, s>>>((const S*)sim, sim_pitch, (const float*)truth, truth_pitch, ny, nx, thr,
                                                         counts);
    else
        return HD_ERR_UNSUPPORTED;
    HD_LAUNCH_CHECK();
    hd_count_launch();
    return HD_OK;
}

}  // namespace

// counts (DEVICE, 6 x uint64, zeroed here): TP, FN, FP, TN, total positives, total negatives.  sim: F32 / F64, truth: U8 /
// I16 / F32.  No host synchronisation.
extern "C" int hd_confusion_counts(const void* sim, int sim_dtype, int64_t sim_pitch, const void* truth, int truth_dtype,
                                   int64_t truth_pitch, int64_t ny, int64_t nx, double threshold, void* counts, void* stream)
{
    if (!sim || !truth || !counts) return HD_ERR_NULL;
    if (ny < 1 || nx < 1 || sim_pitch < nx || truth_pitch < nx) return HD_ERR_ARG;
    if (ny * nx / ((int64_t)hd_num_sms() * 16) >= 0xffffffffLL) return HD_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    HD_CUDA_OK(cudaMemsetAsync(counts, 0, 6 * sizeof(unsigned long long), s));
    if (sim_dtype == HD_F32)
        return launch<float>(sim, sim_pitch, truth, truth_dtype, truth_pitch, ny, nx, threshold, (unsigned long long*)counts, s);
    if (sim_dtype == HD_F64)
        return launch<double>(sim, sim_pitch, truth, truth_dtype, truth_pitch, ny, nx, threshold, (unsigned long long*)counts, s);
    return HD_ERR_UNSUPPORTED;
}
