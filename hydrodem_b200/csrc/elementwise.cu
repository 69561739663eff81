// Elementwise filters of the reference (filters/simple_filters.py, extension_filters.py:12-130) as one
// dtype-generic kernel.  These are HBM-bound streaming ops; inside the fused chain they never run alone
// (they are folded into the neighbouring stencil kernels) -- this entry point exists so that each
// reference class has a device implementation behind the same Filter.apply() API.
#include "common.cuh"
#include "tile_common.cuh"

namespace {

struct cplx { double re, im; };

__device__ __forceinline__ cplx load_as(const void* p, int dtype, int64_t idx)
{
    switch (dtype) {
        case HD_U8: return {(double)((const uint8_t*)p)[idx], 0.0};
        case HD_F32: return {(double)((const float*)p)[idx], 0.0};
        case HD_F64: return {((const double*)p)[idx], 0.0};
        case HD_I64: return {(double)((const int64_t*)p)[idx], 0.0};
        case HD_I32: return {(double)((const int32_t*)p)[idx], 0.0};
        case HD_I16: return {(double)((const int16_t*)p)[idx], 0.0};
        case HD_C64: { float2 v = ((const float2*)p)[idx]; return {(double)v.x, (double)v.y}; }
        default: { double2 v = ((const double2*)p)[idx]; return {v.x, v.y}; }
    }
}
__device__ __forceinline__ int64_t load_int(const void* p, int dtype, int64_t idx)
{
    switch (dtype) {
        case HD_U8: return ((const uint8_t*)p)[idx];
        case HD_I32: return ((const int32_t*)p)[idx];
        case HD_I16: return ((const int16_t*)p)[idx];
        default: return ((const int64_t*)p)[idx];
    }
}
__device__ __forceinline__ void store_as(void* p, int dtype, int64_t idx, cplx v)
{
    switch (dtype) {
        case HD_U8: ((uint8_t*)p)[idx] = (uint8_t)v.re; break;
        case HD_F32: ((float*)p)[idx] = (float)v.re; break;
        case HD_F64: ((double*)p)[idx] = v.re; break;
        case HD_I64: ((int64_t*)p)[idx] = (int64_t)v.re; break;
        case HD_I32: ((int32_t*)p)[idx] = (int32_t)v.re; break;
        case HD_I16: ((int16_t*)p)[idx] = (int16_t)(int32_t)v.re; break;
        case HD_C64: ((float2*)p)[idx] = make_float2((float)v.re, (float)v.im); break;
        default: ((double2*)p)[idx] = make_double2(v.re, v.im); break;
    }
}

__global__ void __launch_bounds__(256) elementwise_kernel(int op, const void* __restrict__ a, int a_dtype,
                                                          int64_t a_pitch, const void* __restrict__ b, int b_dtype,
                                                          int64_t b_pitch, double b_scalar, void* __restrict__ out,
                                                          int out_dtype, int64_t out_pitch, int64_t ny, int64_t nx)
{
    for (CellIter it(nx); it.y < ny; it.next()) {
        const int64_t y = (int64_t)it.y, x = (int64_t)it.x;
        const int64_t ia = y * a_pitch + x, io = y * out_pitch + x;
        if (op == HD_OP_XOR) {
            const int64_t va = load_int(a, a_dtype, ia);
            const int64_t vb = b ? load_int(b, b_dtype, y * b_pitch + x) : (int64_t)b_scalar;
            const int64_t r = va ^ vb;
            if (out_dtype == HD_U8) ((uint8_t*)out)[io] = (uint8_t)r;
            else if (out_dtype == HD_I32) ((int32_t*)out)[io] = (int32_t)r;
            else if (out_dtype == HD_I16) ((int16_t*)out)[io] = (int16_t)r;
            else ((int64_t*)out)[io] = r;
            continue;
        }
        const cplx va = load_as(a, a_dtype, ia);
        const cplx vb = b ? load_as(b, b_dtype, y * b_pitch + x) : cplx{b_scalar, 0.0};
        cplx r{0.0, 0.0};
        switch (op) {
            case HD_OP_COPY: r = va; break;
            case HD_OP_MUL:
                r.re = __dsub_rn(__dmul_rn(vb.re, va.re), __dmul_rn(vb.im, va.im));
                r.im = __dadd_rn(__dmul_rn(vb.re, va.im), __dmul_rn(vb.im, va.re));
                break;
            case HD_OP_ADD: r = {vb.re + va.re, vb.im + va.im}; break;
            case HD_OP_RSUB: r = {vb.re - va.re, vb.im - va.im}; break;
            case HD_OP_LT: r.re = va.re < vb.re ? 1.0 : 0.0; break;
            case HD_OP_GT: r.re = va.re > vb.re ? 1.0 : 0.0; break;
            case HD_OP_ABS: r.re = (va.im == 0.0) ? fabs(va.re) : hypot(va.re, va.im); break;
            case HD_OP_RINT: r.re = rint(va.re); break;
            case HD_OP_TRUNC: r.re = (va.re == va.re) ? trunc(va.re) : 0.0; break;
            default: break;
        }
        store_as(out, out_dtype, io, r);
    }
}

}  // namespace

extern "C" int hd_elementwise(int op, const void* a, int a_dtype, int64_t a_pitch, const void* b, int b_dtype,
                              int64_t b_pitch, double b_scalar, void* out, int out_dtype, int64_t out_pitch, int64_t ny,
                              int64_t nx, void* stream)
{
    if (!a || !out) return HD_ERR_NULL;
    if (op < HD_OP_COPY || op > HD_OP_TRUNC) return HD_ERR_ARG;
    if (!hd_dtype_size(a_dtype) || !hd_dtype_size(out_dtype) || (b && !hd_dtype_size(b_dtype))) return HD_ERR_ARG;
    if (ny < 0 || nx < 0 || a_pitch < nx || out_pitch < nx || (b && b_pitch < nx)) return HD_ERR_ARG;
    if (op == HD_OP_XOR) {
        auto is_int = [](int d) { return d == HD_U8 || d == HD_I64 || d == HD_I32 || d == HD_I16; };
        if (!is_int(a_dtype) || (b && !is_int(b_dtype)) || !is_int(out_dtype)) return HD_ERR_UNSUPPORTED;
    }
    if (ny == 0 || nx == 0) return HD_OK;
    const int64_t total = ny * nx;
    const int blocks = (int)((total + 255) / 256 < (int64_t)hd_num_sms() * 16 ? (total + 255) / 256 : hd_num_sms() * 16);
    hd_prof_begin("elementwise_kernel", (cudaStream_t)stream);
    elementwise_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(op, a, a_dtype, a_pitch, b, b_dtype, b_pitch, b_scalar,
                                                                 out, out_dtype, out_pitch, ny, nx);
    HD_LAUNCH_CHECK();
    hd_count_launch();
    return HD_OK;
}

// ---- HydroDEMProcess._prepare_final_terms + the three-term sum (hydro_dem_process.py:60-91, :148) ----------
//   mask   = lagoon_values > 0                         (LagoonsDetection.mask_lagoons, custom_filters.py:659)
//   first  = srtm * (1 - (mask + rivers))
//   third  = hsheds_fixed * rivers
//   out    = (first + lagoon_values) + third           float64 arithmetic, rounded once to the output dtype
namespace {
template <typename SrtmT, typename OutT>
__global__ void __launch_bounds__(256) final_terms_kernel(const SrtmT* __restrict__ srtm, int64_t srtm_pitch,
                                                          const float* __restrict__ lagoons, int64_t lag_pitch,
                                                          const float* __restrict__ hsheds, int64_t hs_pitch,
                                                          const float* __restrict__ rivers, int64_t riv_pitch,
                                                          OutT* __restrict__ out, int64_t out_pitch, int64_t ny, int64_t nx)
{
    // four consecutive cells per thread: 16-byte loads, enough of them in flight to cover the HBM latency
    const int64_t nxq = (nx + 3) / 4;
    for (CellIter it(nxq); it.y < ny; it.next()) {
        const int64_t y = it.y, x = 4 * it.x;
        SrtmT sv[4];
        float lv4[4], hv[4], rv[4] = {0.f, 0.f, 0.f, 0.f};
        gload4(srtm + y * srtm_pitch + x, x, nx, sv);
        gload4(lagoons + y * lag_pitch + x, x, nx, lv4);
        gload4(hsheds + y * hs_pitch + x, x, nx, hv);
        if (rivers) gload4(rivers + y * riv_pitch + x, x, nx, rv);
        OutT res[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double s = (double)sv[j];
            const double lv = (double)lv4[j];
            const double r = rivers ? (double)rv[j] : 0.0;
            const double mask = lv > 0.0 ? 1.0 : 0.0;
            const double first = __dmul_rn(s, 1.0 - (mask + r));
            double acc = __dadd_rn(first, lv);
            if (rivers) acc = __dadd_rn(acc, __dmul_rn((double)hv[j], r));
            else acc = __dadd_rn(acc, 0.0 * (double)hv[j]);
            res[j] = (OutT)acc;
        }
        store4v<OutT>(out, out_pitch, y, x, nx, res);
    }
}
}  // namespace

extern "C" int hd_final_terms(const void* srtm, int srtm_dtype, int64_t srtm_pitch, const void* lagoon_values,
                              int64_t lag_pitch, const void* hsheds_fixed, int64_t hs_pitch, const void* rivers,
                              int64_t riv_pitch, void* out, int out_dtype, int64_t out_pitch, int64_t ny, int64_t nx,
                              void* stream)
{
    if (!srtm || !lagoon_values || !hsheds_fixed || !out) return HD_ERR_NULL;
    if (ny < 1 || nx < 1 || srtm_pitch < nx || lag_pitch < nx || hs_pitch < nx || out_pitch < nx || (rivers && riv_pitch < nx))
        return HD_ERR_ARG;
    const int64_t total = ny * ((nx + 3) / 4);
    const int blocks = (int)((total + 255) / 256 < (int64_t)hd_num_sms() * 16 ? (total + 255) / 256 : hd_num_sms() * 16);
    cudaStream_t s = (cudaStream_t)stream;
#define HD_FT(ST, STAG, OT, OTAG)                                                                                      \
    if (srtm_dtype == STAG && out_dtype == OTAG) {                                                                     \
        hd_prof_begin("final_terms_kernel", s);                                                                        \
        final_terms_kernel<ST, OT><<<blocks, 256, 0, s>>>((const ST*)srtm, srtm_pitch, (const float*)lagoon_values,    \
                                                          lag_pitch, (const float*)hsheds_fixed, hs_pitch,             \
                                                          (const float*)rivers, riv_pitch, (OT*)out, out_pitch, ny, nx); \
        HD_LAUNCH_CHECK();                                                                                             \
        hd_count_launch();                                                                                             \
        return HD_OK;                                                                                                  \
    }
    HD_FT(float, HD_F32, float, HD_F32)
    HD_FT(float, HD_F32, double, HD_F64)
    HD_FT(double, HD_F64, double, HD_F64)
#undef HD_FT
    return HD_ERR_UNSUPPORTED;
}

// ---- lossless int16 transport of integer-valued float rasters (host API) -----------------------------------------------
namespace {
__global__ void __launch_bounds__(256) pack_i16_kernel(const float* __restrict__ src, int64_t src_pitch,
                                                       int16_t* __restrict__ dst, int64_t ny, int64_t nx,
                                                       int* __restrict__ inexact)
{
    const int64_t nxq = (nx + 3) / 4;
    bool bad = false;
    for (CellIter it(nxq); it.y < ny; it.next()) {
        const int64_t y = it.y, x = 4 * it.x;
        float v[4];
        gload4(src + y * src_pitch + x, x, nx, v);
        int16_t* q = dst + y * nx + x;                        // dense rows
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (x + j >= nx) break;
            const float r = rintf(v[j]);
            bad |= !(r == v[j] && r >= -32768.f && r <= 32767.f);      // NaN, fractions and out-of-range values
            q[j] = (int16_t)(int)fminf(fmaxf(r, -32768.f), 32767.f);
        }
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) *inexact = 1;
}
}  // namespace

extern "C" int hd_pack_i16(const void* src, int64_t src_pitch, void* dst_dense, int64_t ny, int64_t nx, int* inexact_flag,
                           void* stream)
{
    if (!src || !dst_dense || !inexact_flag) return HD_ERR_NULL;
    if (ny < 1 || nx < 1 || src_pitch < nx) return HD_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    HD_CUDA_OK(cudaMemsetAsync(inexact_flag, 0, sizeof(int), s));
    const int64_t total = ny * ((nx + 3) / 4);
    const int blocks = (int)((total + 255) / 256 < (int64_t)hd_num_sms() * 16 ? (total + 255) / 256 : hd_num_sms() * 16);
    hd_prof_begin("pack_i16_kernel", s);
    pack_i16_kernel<<<blocks, 256, 0, s>>>((const float*)src, src_pitch, (int16_t*)dst_dense, ny, nx, inexact_flag);
    HD_LAUNCH_CHECK();
    hd_count_launch();
    return HD_OK;
}
