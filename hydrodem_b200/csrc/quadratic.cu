// QuadraticFilter.apply (filters/custom_filters.py:226-257) and the fused GrovesCorrection tail (:704-732).
//
// The reference evaluates, per window (cast to float32),
//     s1 = sum w,  s2 = sum w x^2,  s3 = sum w y^2,
//     smoothed = ((s2 + s3) r1 - s1 (r2 + r3)) / (2 r1^2 - r0 (r2 + r3))
// with x, y = linspace(-ws/2 + 1, ws/2, ws) (half-pixel asymmetric offsets).  All three sums are separable:
//     column pass : V1[y][x] = sum_dy w,  Vyy[y][x] = sum_dy y^2 w        (one pass down each tile column)
//     row pass    : s1 = sum_dx V1,  s3 = sum_dx Vyy,  s2 = sum_dx x^2 V1
// i.e. 2*ws taps instead of ws^2 per sum.  The window values are cast to float32 like the reference's windows
// (sliding_window.py:132) and every sum is accumulated in DOUBLE, tap by tap in a fixed order: a cell's result is a
// fixed sequence of operations on its own window and nothing else -- it does not depend on where tile or band
// boundaries fall, so a row-band sharded mosaic gives the same bits as one GPU (round 1 re-centred float32 sums on a
// per-tile reference elevation, which made borderline `dem - smooth > 1.5` tests flip with the partition).  The
// result agrees with the reference (float32 s1, float64 s2 / s3) to ~1e-7 relative: inside the 1e-5 tolerance class.
//
// Groves tail (one GrovesCorrection iteration, fused -- the six elementwise filters never touch HBM):
//     hi = dem - smooth;  tall = hi > 1.5;  keep = 1 - groves * tall;  out = keep * hi + smooth
//
// Algorithmic HBM traffic: 4 B read + 4 B written per cell (+1 B groves mask) for float32 rasters.
#include "common.cuh"
#include "tile_common.cuh"

namespace {

constexpr int STRIP = 8;
constexpr int MAX_WS = 15;

struct QuadParams {
    double c2[MAX_WS];   // squared coordinate per tap
    double r1, r23, den; // r1, r2 + r3, 2 r1^2 - r0 (r2 + r3)
    double thr;          // tall-groves threshold (1.5)
};

constexpr int HALF = TH / 2;          // partial sums are kept for half a tile at a time (shared memory: 2 CTAs per SM)

template <int H, typename T, typename OutT, bool GROVES>
__global__ void __launch_bounds__(NT) quadratic_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                       const __grid_constant__ CUtensorMap tm_groves,
                                                       OutT* __restrict__ out, int64_t out_pitch, int64_t ny, int64_t nx,
                                                       QuadParams p, int tiles_x, int ntiles,
                                                       uint8_t* __restrict__ tile_flags = nullptr, int sparse = 0)
{
    constexpr int WS = 2 * H + 1;
    constexpr int CWU = TW + 2 * H;                 // window columns touched by a tile
    constexpr int CW = (CWU + 1) / 2 * 2 + 2;       // row stride of the partial-sum arrays (doubles): 16-byte aligned rows
    constexpr int HX = hd_halo_x(H, sizeof(T));
    constexpr int XOFF = HX - H;
    constexpr int IN_W = TW + 2 * HX;
    constexpr int IN_H = TH + 2 * H;
    constexpr uint32_t IN_BYTES = (IN_W * IN_H * sizeof(T) + 127) / 128 * 128;
    constexpr uint32_t STAGE = IN_BYTES + (GROVES ? TH * TW : 0);     // + the tile's groves mask (uint8), staged by TMA too
    constexpr int NP = GROVES ? 2 : 1;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[2];
    double* v1 = reinterpret_cast<double*>(smem + 2 * STAGE);   // [HALF][CW]  sum_dy w
    double* vyy = v1 + HALF * CW;                                // [HALF][CW]  sum_dy y^2 w

    const TilePlane planes[2] = {{&tm_in, 0u, (uint32_t)(IN_W * IN_H * sizeof(T)), HX, H},
                                 {&tm_groves, IN_BYTES, (uint32_t)(TH * TW), 0, 0}};
    const TilePlane (&used)[NP] = reinterpret_cast<const TilePlane (&)[NP]>(planes);
    // Groves iterations 2, 3 ... (`sparse`): a cell without groves never changes, so the output raster -- the buffer the
    // PREVIOUS iteration read from -- already holds it; only the tiles that contain a groves cell (tile_flags, written by the
    // first iteration) are loaded at all, and inside them only the quads with a groves cell are stored.
    auto body = [&](unsigned char* st, int ty0, int tx0, int tile_id) {
        const T* tile = reinterpret_cast<const T*>(st);
        const uint32_t* gwords = reinterpret_cast<const uint32_t*>(st + IN_BYTES);   // [TH][TW / 4], zero outside the raster
        // Groves tail: where groves == 0 the reference evaluates smooth + 1 * (dem - smooth), i.e. dem to the last bit or
        // two -- those cells are copied, and a tile without a single groves cell (most of them: groves cover about one
        // per cent of a scene) skips both filter passes.
        bool any_half[2] = {true, true};
        if (GROVES) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                bool any = false;
#pragma unroll
                for (int rep = 0; rep < HALF * TW / 4 / NT; ++rep) {
                    const int idx = hf * (HALF * TW / 4) + rep * NT + threadIdx.x;
                    const int ro = idx >> 5, c4 = idx & 31;
                    const int64_t y = ty0 + ro, x = tx0 + 4 * c4;
                    const uint32_t g = (y < ny && x < nx) ? gwords[idx] : 0u;
                    any |= g != 0u;
                }
                any_half[hf] = __syncthreads_or(any);
            }
            if (tile_flags && !sparse && threadIdx.x == 0) tile_flags[tile_id] = (any_half[0] || any_half[1]) ? 1 : 0;
        }
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
            const int r0 = hf * HALF;                               // first tile row of this half
            if (any_half[hf]) {
                // ---- column pass: fixed tap order, double accumulators ------------------------------------------
                for (int item = threadIdx.x; item < CWU * (HALF / STRIP); item += NT) {
                    const int c = item % CWU, sq = item / CWU;
                    float v[STRIP + 2 * H];
#pragma unroll
                    for (int r = 0; r < STRIP + 2 * H; ++r)
                        v[r] = (float)tile[(r0 + sq * STRIP + r) * IN_W + c + XOFF];          // window cast to float32
#pragma unroll
                    for (int o = 0; o < STRIP; ++o) {
                        double a = 0.0, b = 0.0;
#pragma unroll
                        for (int r = 0; r < WS; ++r) {
                            a += (double)v[o + r];
                            b = fma(p.c2[r], (double)v[o + r], b);
                        }
                        v1[(sq * STRIP + o) * CW + c] = a;
                        vyy[(sq * STRIP + o) * CW + c] = b;
                    }
                }
                __syncthreads();
            }
            // ---- row pass: 4 consecutive outputs per thread -----------------------------------------------------
#pragma unroll 1
            for (int rep = 0; rep < HALF * TW / 4 / NT; ++rep) {
                const int idx = r0 * (TW / 4) + rep * NT + threadIdx.x;
                const int ro = idx >> 5, c4 = idx & 31;
                const int64_t y = ty0 + ro, x = tx0 + 4 * c4;
                if (y >= ny || x >= nx) continue;
                const uint32_t gq = GROVES ? gwords[idx] : 0u;
                if (GROVES && gq == 0u) {                     // no groves cell in this quad: copy
                    if (sparse) continue;                     // (the output raster already holds these cells)
                    OutT cp[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) cp[j] = (OutT)tile[(ro + H) * IN_W + 4 * c4 + j + HX];
                    store4v<OutT>(out, out_pitch, y, x, nx, cp);
                    continue;
                }
                const double* pa = v1 + (ro - r0) * CW + 4 * c4;
                const double* pb = vyy + (ro - r0) * CW + 4 * c4;
                OutT res[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const T ctr = tile[(ro + H) * IN_W + 4 * c4 + j + HX];
                    const bool interior = y >= H && y < ny - H && x + j >= H && x + j < nx - H;
                    const uint32_t gbyte = (gq >> (8 * j)) & 0xffu;
                    if (GROVES && gbyte == 0u) { res[j] = (OutT)ctr; continue; }
                    T smooth = ctr;                                    // smoothed = dem.copy()  (:249)
                    if (interior) {
                        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                        for (int d = 0; d < WS; ++d) {
                            s1 += pa[j + d];
                            s2 = fma(p.c2[d], pa[j + d], s2);
                            s3 += pb[j + d];
                        }
                        smooth = (T)(((s2 + s3) * p.r1 - s1 * p.r23) / p.den);        // (:255-256), stored in dem's dtype
                    }
                    if (GROVES) {
                        const double g = (double)gbyte;
                        const T hi = ctr - smooth;                                        // SubtractionFilter (:725)
                        const double tall = (hi > (T)p.thr) ? 1.0 : 0.0;                  // MaskTallGroves
                        const double keep = 1.0 - g * tall;                               // Product, 1 - .
                        res[j] = (OutT)__dadd_rn(__dmul_rn(keep, (double)hi), (double)smooth);   // x hi, + smooth (:729-731)
                    } else {
                        res[j] = (OutT)smooth;
                    }
                }
                store4v<OutT>(out, out_pitch, y, x, nx, res);
            }
            if (any_half[hf] && hf == 0) __syncthreads();       // the partial sums are rewritten for the second half
        }
    };
    if (GROVES && sparse)
        tile_loop_flagged<NP>(smem, STAGE, bars, used, TW, TH, tiles_x, ntiles, tile_flags, body);
    else
        tile_loop<NP>(smem, STAGE, bars, used, TW, TH, tiles_x, ntiles,
                      [&](unsigned char* st, int ty0, int tx0) { body(st, ty0, tx0, (ty0 / TH) * tiles_x + tx0 / TW); });
}

template <int H, typename T, typename OutT, bool GROVES>
int launch(const void* in, int dtype, int64_t in_pitch, void* out, int64_t out_pitch, const void* groves,
           int64_t groves_pitch, int64_t ny, int64_t nx, const QuadParams& p, cudaStream_t stream,
           uint8_t* tile_flags = nullptr, int sparse = 0)
{
    constexpr int CW = (TW + 2 * H + 1) / 2 * 2 + 2, IN_W = TW + 2 * hd_halo_x(H, sizeof(T)), IN_H = TH + 2 * H;
    constexpr size_t STAGE = (IN_W * IN_H * sizeof(T) + 127) / 128 * 128 + (GROVES ? TH * TW : 0);
    constexpr size_t SMEM = 2 * STAGE + 2 * HALF * CW * sizeof(double);
    CUtensorMap tm, tm_g;
    if (int e = hd_make_tmap_2d(&tm, in, dtype, ny, nx, in_pitch, IN_W, IN_H, false)) return e;
    tm_g = tm;
    if (GROVES)
        if (int e = hd_make_tmap_2d(&tm_g, groves, HD_U8, ny, nx, groves_pitch, TW, TH, false)) return e;
    const int tiles_x = hd_cdiv(nx, TW), tiles_y = hd_cdiv(ny, TH), ntiles = tiles_x * tiles_y;
    auto kern = quadratic_kernel<H, T, OutT, GROVES>;
    HD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    hd_prof_begin("quadratic_kernel", stream);
    kern<<<grid_for(ntiles, 2), NT, SMEM, stream>>>(tm, tm_g, (OutT*)out, out_pitch, ny, nx, p, tiles_x, ntiles, tile_flags,
                                                    sparse);
    HD_LAUNCH_CHECK();
    hd_count_launch();
    return HD_OK;
}

template <typename T, typename OutT, bool GROVES>
int dispatch_h(int h, const void* in, int dtype, int64_t in_pitch, void* out, int64_t out_pitch, const void* groves,
               int64_t groves_pitch, int64_t ny, int64_t nx, const QuadParams& p, cudaStream_t s,
               uint8_t* tile_flags = nullptr, int sparse = 0)
{
    switch (h) {
#define HD_Q_CASE(HH) \
    case HH: return launch<HH, T, OutT, GROVES>(in, dtype, in_pitch, out, out_pitch, groves, groves_pitch, ny, nx, p, s, \
                                                tile_flags, sparse);
        HD_Q_CASE(1) HD_Q_CASE(2) HD_Q_CASE(3) HD_Q_CASE(4) HD_Q_CASE(5) HD_Q_CASE(6) HD_Q_CASE(7)
#undef HD_Q_CASE
        default: return HD_ERR_UNSUPPORTED;
    }
}

// r0..r3 of custom_filters.py:240-246, evaluated like numpy: linspace(-ws/2 + 1, ws/2, ws)
QuadParams make_params(int ws, double thr)
{
    QuadParams p{};
    const double start = -ws / 2.0 + 1.0, stop = ws / 2.0;
    const double step = (stop - start) / (ws - 1);
    double vals[MAX_WS];
    for (int k = 0; k < ws; ++k) vals[k] = (k == ws - 1) ? stop : start + k * step;
    double r1 = 0, r2 = 0, r3 = 0;
    for (int yy = 0; yy < ws; ++yy)
        for (int xx = 0; xx < ws; ++xx) {
            const double x2 = vals[xx] * vals[xx], y2 = vals[yy] * vals[yy];
            r1 += x2;
            r2 += x2 * x2;
            r3 += x2 * y2;
        }
    for (int k = 0; k < ws; ++k) p.c2[k] = vals[k] * vals[k];
    const double r0 = (double)ws * ws;
    p.r1 = r1;
    p.r23 = r2 + r3;
    p.den = 2 * r1 * r1 - r0 * (r2 + r3);
    p.thr = thr;
    return p;
}

int run(const void* in, int64_t in_pitch, void* out, int64_t out_pitch, int in_dtype, int out_dtype, const void* groves,
        int64_t groves_pitch, int64_t ny, int64_t nx, int ws, double thr, void* stream, uint8_t* tile_flags = nullptr,
        int sparse = 0)
{
    if (!in || !out) return HD_ERR_NULL;
    if (int e = check_window(ny, nx, ws)) return e;
    if (ws < 3 || ws > MAX_WS) return HD_ERR_UNSUPPORTED;
    if (in_pitch < nx || out_pitch < nx) return HD_ERR_ARG;
    const QuadParams p = make_params(ws, thr);
    const int h = ws / 2;
    cudaStream_t s = (cudaStream_t)stream;
    const bool g = groves != nullptr;
#define HD_Q_T(TT, TTAG, OT, OTAG)                                                                                      \
    if (in_dtype == TTAG && out_dtype == OTAG)                                                                          \
        return g ? dispatch_h<TT, OT, true>(h, in, in_dtype, in_pitch, out, out_pitch, groves, groves_pitch, ny, nx, p, s, \
                                            tile_flags, sparse)                                                         \
                 : dispatch_h<TT, OT, false>(h, in, in_dtype, in_pitch, out, out_pitch, groves, groves_pitch, ny, nx, p, s);
    HD_Q_T(float, HD_F32, float, HD_F32)
    HD_Q_T(float, HD_F32, double, HD_F64)
    HD_Q_T(double, HD_F64, double, HD_F64)
#undef HD_Q_T
    return HD_ERR_UNSUPPORTED;
}

}  // namespace

extern "C" int hd_quadratic(const void* in, int64_t in_pitch, void* out, int64_t out_pitch, int dtype, int64_t ny,
                            int64_t nx, int ws, void* stream)
{
    return run(in, in_pitch, out, out_pitch, dtype, dtype, nullptr, 0, ny, nx, ws, 0.0, stream);
}

extern "C" int hd_groves_correction(const void* in, int in_dtype, int64_t in_pitch, const void* groves,
                                    int64_t groves_pitch, void* out, int out_dtype, int64_t out_pitch, int64_t ny,
                                    int64_t nx, int ws, double threshold, void* stream)
{
    if (!groves) return HD_ERR_NULL;
    if (groves_pitch < nx) return HD_ERR_ARG;
    return run(in, in_pitch, out, out_pitch, in_dtype, out_dtype, groves, groves_pitch, ny, nx, ws, threshold, stream);
}

// GrovesCorrectionsIter (custom_filters.py:735-767) without re-copying the cells that cannot change.  A cell outside the
// groves class comes out of GrovesCorrection as it went in, so from the second iteration on only groves cells need work:
//   sparse = 0  (first iteration): full pass in -> out, and tile_flags[tile] = "this 32 x 128 tile holds a groves cell"
//   sparse = 1  (later iterations): `out` must already hold the previous iteration's INPUT (ping-pong between two rasters:
//               iteration k writes into the buffer iteration k-1 read); only flagged tiles are loaded, only quads with a
//               groves cell are stored.  Same bits as dense iterations.
// tile_flags: hd_groves_tile_count(ny, nx) bytes.  F32 rasters only.
extern "C" int64_t hd_groves_tile_count(int64_t ny, int64_t nx) { return (int64_t)hd_cdiv(ny, TH) * hd_cdiv(nx, TW); }

extern "C" int hd_groves_correction_tiles(const void* in, int64_t in_pitch, const void* groves, int64_t groves_pitch, void* out,
                                          int64_t out_pitch, int64_t ny, int64_t nx, int ws, double threshold, void* tile_flags,
                                          int sparse, void* stream)
{
    if (!groves || !tile_flags) return HD_ERR_NULL;
    if (groves_pitch < nx || in == out) return HD_ERR_ARG;
    return run(in, in_pitch, out, out_pitch, HD_F32, HD_F32, groves, groves_pitch, ny, nx, ws, threshold, stream,
               (uint8_t*)tile_flags, sparse ? 1 : 0);
}
