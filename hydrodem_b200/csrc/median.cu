// NEW stage N1 -- median filter (no reference code: parity is against this repo's own oracle,
// oracle/stencils.py:median, "parity unpinned").  Definition: out = dem.copy(); every
// interior cell gets np.nanmedian of its float32 ws*ws window (ws = 3 or 5; corner-less when `circular`);
// the ws/2 border keeps the input (QuadraticFilter convention, custom_filters.py:249).
//
// Each thread sorts its window in registers with a pruned Batcher network on order-preserving integer keys
// (NaN -> largest key), then picks the middle of the k non-NaN values (mean of the two middles when k is
// even).  This stage is ALU bound (min/max pipe), not HBM bound: 8 B/cell of traffic against ~2*140
// min/max per cell for 5x5.
#include "common.cuh"
#include "tile_common.cuh"
#include "sort_networks.cuh"

namespace {

__device__ __forceinline__ uint32_t key_of(float v)
{
    const uint32_t u = __float_as_uint(v);
    if (v != v) return 0xffffffffu;                         // NaN sorts last, like np.sort
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float val_of(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

#define HD_CE(a, b)                    \
    do {                               \
        const uint32_t _lo = min(a, b); \
        b = max(a, b);                 \
        a = _lo;                       \
    } while (0)

template <int N> __device__ __forceinline__ void sort_keys(uint32_t (&v)[N]);
template <> __device__ __forceinline__ void sort_keys<5>(uint32_t (&v)[5]) { HD_SORT_NET_5(v) }
template <> __device__ __forceinline__ void sort_keys<9>(uint32_t (&v)[9]) { HD_SORT_NET_9(v) }
template <> __device__ __forceinline__ void sort_keys<21>(uint32_t (&v)[21]) { HD_SORT_NET_21(v) }
template <> __device__ __forceinline__ void sort_keys<25>(uint32_t (&v)[25]) { HD_SORT_NET_25(v) }

template <int H, bool CIRC>
__global__ void __launch_bounds__(NT) median_kernel(const __grid_constant__ CUtensorMap tm_in, float* __restrict__ out,
                                                    int64_t out_pitch, int64_t ny, int64_t nx, int tiles_x, int ntiles)
{
    constexpr int WS = 2 * H + 1;
    constexpr int N = WS * WS - (CIRC ? 4 : 0);
    constexpr int HX = hd_halo_x(H, 4);
    constexpr int IN_W = TW + 2 * HX, IN_H = TH + 2 * H;
    constexpr uint32_t STAGE = (IN_W * IN_H * 4 + 127) / 128 * 128;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[2];
    const TilePlane planes[1] = {{&tm_in, 0u, (uint32_t)(IN_W * IN_H * 4), HX, H}};
    tile_loop<1>(smem, STAGE, bars, planes, TW, TH, tiles_x, ntiles, [&](unsigned char* st, int ty0, int tx0) {
        const float* tile = reinterpret_cast<const float*>(st);
#pragma unroll 1
        for (int rep = 0; rep < TH * TW / NT; ++rep) {
            const int idx = rep * NT + threadIdx.x;
            const int ro = idx / TW, xo = idx % TW;
            const int64_t y = ty0 + ro, x = tx0 + xo;
            if (y >= ny || x >= nx) continue;
            float result = tile[(ro + H) * IN_W + xo + HX];
            if (y >= H && y < ny - H && x >= H && x < nx - H) {
                uint32_t v[N];
                int n = 0, valid = 0;
                const float* w = tile + ro * IN_W + xo + HX - H;
#pragma unroll
                for (int dy = 0; dy < WS; ++dy)
#pragma unroll
                    for (int dx = 0; dx < WS; ++dx) {
                        if (CIRC && (dy == 0 || dy == WS - 1) && (dx == 0 || dx == WS - 1)) continue;
                        const float f = w[dy * IN_W + dx];
                        valid += (f == f);
                        v[n++] = key_of(f);
                    }
                sort_keys<N>(v);
                if (valid == 0) {
                    result = __int_as_float(0x7fc00000);
                } else {
                    const int lo = (valid - 1) >> 1, hi = valid >> 1;
                    uint32_t klo = v[0], khi = v[0];
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        if (i == lo) klo = v[i];
                        if (i == hi) khi = v[i];
                    }
                    const float a = val_of(klo), b = val_of(khi);
                    // np.nanmedian: middle value, or float32 mean of the two middles
                    result = (lo == hi) ? a : __fdiv_rn(__fadd_rn(a, b), 2.0f);
                }
            }
            out[y * out_pitch + x] = result;
        }
    });
}

template <int H, bool CIRC>
int launch(const void* in, int64_t in_pitch, void* out, int64_t out_pitch, int64_t ny, int64_t nx, cudaStream_t stream)
{
    constexpr int IN_W = TW + 2 * hd_halo_x(H, 4), IN_H = TH + 2 * H;
    constexpr size_t STAGE = (IN_W * IN_H * 4 + 127) / 128 * 128;
    CUtensorMap tm;
    if (int e = hd_make_tmap_2d(&tm, in, HD_F32, ny, nx, in_pitch, IN_W, IN_H, false)) return e;
    const int tiles_x = hd_cdiv(nx, TW), tiles_y = hd_cdiv(ny, TH), ntiles = tiles_x * tiles_y;
    hd_prof_begin("median_kernel", stream);
    median_kernel<H, CIRC><<<grid_for(ntiles, 4), NT, 2 * STAGE, stream>>>(tm, (float*)out, out_pitch, ny, nx, tiles_x,
                                                                         ntiles);
    HD_LAUNCH_CHECK();
    hd_count_launch();
    return HD_OK;
}

}  // namespace

extern "C" int hd_median(const void* in, int64_t in_pitch, void* out, int64_t out_pitch, int64_t ny, int64_t nx, int ws,
                         int circular, void* stream)
{
    if (!in || !out) return HD_ERR_NULL;
    if (int e = check_window(ny, nx, ws)) return e;
    if (in_pitch < nx || out_pitch < nx) return HD_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    if (ws == 3) return circular ? launch<1, true>(in, in_pitch, out, out_pitch, ny, nx, s)
                                 : launch<1, false>(in, in_pitch, out, out_pitch, ny, nx, s);
    if (ws == 5) return circular ? launch<2, true>(in, in_pitch, out, out_pitch, ny, nx, s)
                                 : launch<2, false>(in, in_pitch, out, out_pitch, ny, nx, s);
    return HD_ERR_UNSUPPORTED;
}
