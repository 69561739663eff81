// MajorityFilter.apply  (filters/custom_filters.py:48-73) on the corner-less "circular" window
// (sliding_window.py:475-499).
//
// The reference builds a Counter over the ws*ws window (4 corners NaN'ed, each NaN its own key) and keeps
// the mode only if count > (ws^2-1)*0.7, i.e. the winner always holds more than half of the window.  A value
// with an absolute majority is the Boyer-Moore candidate, and Boyer-Moore summaries (candidate, counter)
// are mergeable, so the kernel
//   1. builds one summary per window COLUMN (2h-1 inner rows, and all 2h+1 rows) -- shared by the 2h+1
//      windows that contain that column,
//   2. merges 2h+1 column summaries per output cell (the two outer columns use the inner-row summary:
//      that is exactly the corner-less footprint); aligned pairs of rows (step 1) and of columns (step 2) are
//      merged once and shared, which roughly halves the dependent update chains,
//   3. verifies the candidate with an exact count only where the summary counter leaves the threshold
//      reachable: true_count <= (n_window + counter) / 2 -- warp-cooperatively, by column counts.
// This kernel is ALU / shared-memory bound, not HBM bound: ~8 B/cell of HBM traffic
// against a few hundred shared-memory operations per cell.
#include "common.cuh"
#include "tile_common.cuh"

namespace {

constexpr int STRIP = 8;   // output rows per column-summary work item
constexpr int MNT = 576;   // threads per CTA: the (TW + 2h) x 4 column-summary items of a tile in ONE pass (h <= 7)
constexpr int MCELL = 512; // of which the first 16 warps merge / verify (8 output cells each)

struct BM { float c; int n; };

// Both updates are written as selects: with `if` chains the compiler emits a divergent branch (BSSY / BRA / BSYNC) per
// element, ~15 instructions per push; the kernel is bound by its instruction count (ncu: 360 instructions per cell).
__device__ __forceinline__ void bm_push(BM& s, float v)
{
    const bool nz = (s.n != 0);
    // NaN != x is true: every NaN is its own key (custom_filters.py:69); an empty summary adopts v, NaN included
    const bool ne = nz && (v != s.c);
    s.c = nz ? s.c : v;
    s.n += ne ? -1 : 1;
}
__device__ __forceinline__ void bm_merge(BM& s, float c2, int n2)
{
    const bool eq = (c2 == s.c);
    const int d = s.n - n2;
    s.c = (!eq && d < 0) ? c2 : s.c;
    s.n = eq ? s.n + n2 : abs(d);
}

template <int H, typename OutT>
__global__ void __launch_bounds__(MNT) majority_kernel(const __grid_constant__ CUtensorMap tm_in, OutT* __restrict__ out,
                                                      int64_t out_pitch, int64_t ny, int64_t nx, int min_count,
                                                      int tiles_x, int ntiles)
{
    constexpr int WS = 2 * H + 1;
    constexpr int CW = TW + 2 * H;                 // window columns touched by a tile
    constexpr int CP = CW / 2;                     // aligned column pairs
    constexpr int HX = hd_halo_x(H, 4);            // x halo of the staged box (16-byte TMA rule)
    constexpr int XOFF = HX - H;
    constexpr int IN_W = TW + 2 * HX;
    constexpr int IN_H = TH + 2 * H;
    constexpr int NWIN = WS * WS - 4;              // cells of the corner-less window
    constexpr uint32_t STAGE = (IN_W * IN_H * 4 + 127) / 128 * 128;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[2];
    __shared__ uint32_t colcnt[MCELL / 32][32 + 2 * H + 2];             // per warp: candidate counts of 32 + 2H window columns
    float* cand_in = reinterpret_cast<float*>(smem + 2 * STAGE);       // [TH][CW] summary over the 2H-1 inner rows
    float* cand_fu = cand_in + TH * CW;                                // [TH][CW] summary over all 2H+1 rows
    float* cand_pr = cand_fu + TH * CW;                                // [TH][CP] full-column summaries of aligned column pairs
    uint8_t* cnt_in = reinterpret_cast<uint8_t*>(cand_pr + TH * CP);
    uint8_t* cnt_fu = cnt_in + TH * CW;
    uint8_t* cnt_pr = cnt_fu + TH * CW;

    const TilePlane planes[1] = {{&tm_in, 0u, (uint32_t)(IN_W * IN_H * 4), HX, H}};
    tile_loop<1>(smem, STAGE, bars, planes, TW, TH, tiles_x, ntiles, [&](unsigned char* st, int ty0, int tx0) {
        const float* tile = reinterpret_cast<const float*>(st);
        // ---- 1. column summaries ---------------------------------------------------------------
        for (int item = threadIdx.x; item < CW * (TH / STRIP); item += MNT) {
            const int c = item % CW, s = item / CW;
            float v[STRIP + 2 * H];
#pragma unroll
            for (int r = 0; r < STRIP + 2 * H; ++r) v[r] = tile[(s * STRIP + r) * IN_W + c + XOFF];
            // Summaries are mergeable: aligned PAIRS of rows (equal -> (v, 2), different -> empty) are shared by the
            // outputs of the strip, so a 2H-1 row summary is one push + H-1 merges instead of 2H-1 dependent pushes.
            float pc[(STRIP + 2 * H) / 2];
            int pn[(STRIP + 2 * H) / 2];
#pragma unroll
            for (int k = 0; k < (STRIP + 2 * H) / 2; ++k) {
                pc[k] = v[2 * k];
                pn[k] = (v[2 * k] == v[2 * k + 1]) ? 2 : 0;
            }
#pragma unroll
            for (int o = 0; o < STRIP; ++o) {
                BM b{0.f, 0};
                int r = o + 1;                                         // inner rows o+1 .. o+2H-1 (compile-time after unrolling)
                if (r & 1) { bm_push(b, v[r]); ++r; }
#pragma unroll
                for (int q = 0; q < H; ++q)
                    if (r + 1 <= o + 2 * H - 1) { bm_merge(b, pc[r / 2], pn[r / 2]); r += 2; }
                if (r <= o + 2 * H - 1) bm_push(b, v[r]);
                const int ro = s * STRIP + o;
                cand_in[ro * CW + c] = b.c;
                cnt_in[ro * CW + c] = (uint8_t)b.n;
                bm_push(b, v[o]);
                bm_push(b, v[o + 2 * H]);
                cand_fu[ro * CW + c] = b.c;
                cnt_fu[ro * CW + c] = (uint8_t)b.n;
            }
        }
        __syncthreads();
        // ---- 1b. aligned pairs of full-column summaries, shared by the cells of a row ----------------------------
        for (int item = threadIdx.x; item < TH * CP; item += MNT) {
            const int ro = item / CP, k = item - ro * CP;
            const int base = ro * CW + 2 * k;
            BM b{cand_fu[base], (int)cnt_fu[base]};
            bm_merge(b, cand_fu[base + 1], (int)cnt_fu[base + 1]);
            cand_pr[item] = b.c;
            cnt_pr[item] = (uint8_t)b.n;
        }
        __syncthreads();
        // ---- 2. merge + 3. verify -------------------------------------------------------------------
        // A warp works on 32 consecutive cells of one output row.  The verification is WARP-COOPERATIVE: for each distinct
        // candidate among the lanes that need one (almost always a single value: the level of the plateau the warp
        // touches), lane j counts the candidate in window column j (and lanes 0 .. 2H-1 in column 32 + j): 42 column
        // counts of 11 cells instead of 32 window counts of 117, then every lane adds up its 2H+1 columns.
        if (threadIdx.x < MCELL) {
            const int lane = threadIdx.x & 31;
            uint32_t* cc = colcnt[threadIdx.x >> 5];
#pragma unroll 1
            for (int rep = 0; rep < TH * TW / MCELL; ++rep) {
                const int idx = rep * MCELL + threadIdx.x;
                const int ro = idx / TW, xo = idx % TW;
                const int64_t y = ty0 + ro, x = tx0 + xo;
                const bool valid = y < ny && x < nx;
                float result = 0.f;                                    // np.zeros border / no majority (:66)
                BM b{0.f, 0};
                bool plausible = false;
                if (valid && y >= H && y < ny - H && x >= H && x < nx - H) {
                    const int base = ro * CW + xo;
                    b.c = cand_in[base];
                    b.n = (int)cnt_in[base];
                    // full columns xo+1 .. xo+2H-1: one single column + H-1 aligned pairs (same code for both parities)
                    const int lo = xo + 1, single = (lo & 1) ? lo : lo + 2 * H - 2;
                    bm_merge(b, cand_fu[ro * CW + single], (int)cnt_fu[ro * CW + single]);
                    const int k0 = ro * CP + ((lo + 1) >> 1);
#pragma unroll
                    for (int q = 0; q < H - 1; ++q) bm_merge(b, cand_pr[k0 + q], (int)cnt_pr[k0 + q]);
                    bm_merge(b, cand_in[base + 2 * H], (int)cnt_in[base + 2 * H]);
                    // true count of the candidate <= (NWIN + counter) / 2; a NaN is its own key (count 1 < min_count)
                    plausible = (NWIN + b.n >= 2 * min_count) && (b.c == b.c);
                }
                unsigned need = __ballot_sync(0xffffffffu, plausible);
                const float* w0 = tile + ro * IN_W + (xo - lane) + XOFF;       // window column 0 of the warp's first cell
#pragma unroll 1
                while (need) {
                    const float c = __shfl_sync(0xffffffffu, b.c, __ffs(need) - 1);
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        const int j = part == 0 ? lane : 32 + (lane < 2 * H ? lane : 0);
                        const float* col = w0 + j;
                        int inner = 0;
#pragma unroll
                        for (int dy = 1; dy < WS - 1; ++dy) inner += (col[dy * IN_W] == c);
                        const int full = inner + (col[0] == c) + (col[(WS - 1) * IN_W] == c);
                        if (part == 0 || lane < 2 * H) cc[j] = (uint32_t)full | ((uint32_t)inner << 16);
                    }
                    __syncwarp();
                    uint32_t sum = 0;
#pragma unroll
                    for (int d = 1; d <= 2 * H - 1; ++d) sum += cc[lane + d];
                    // the two outer columns without their end cells: the corner-less footprint
                    const int count = (int)(sum & 0xffffu) + (int)(cc[lane] >> 16) + (int)(cc[lane + 2 * H] >> 16);
                    __syncwarp();
                    const bool mine = plausible && (b.c == c);
                    if (mine && count >= min_count) {                  // count > (ws^2-1)*0.7  (:71-72)
                        result = b.c;
                        if (b.c == 0.f) {
                            // Counter keeps the first-inserted key of an == class: +0.0 / -0.0 by the scan order
                            const float* w = w0 + lane;
                            bool found = false;
#pragma unroll 1
                            for (int k = 1; k < WS * WS - 1 && !found; ++k) {
                                const int dy = k / WS, dx = k - dy * WS;
                                if ((dy == 0 || dy == WS - 1) && (dx == 0 || dx == WS - 1)) continue;   // NaN corners
                                const float v = w[dy * IN_W + dx];
                                if (v == 0.f) { result = v; found = true; }
                            }
                        }
                    }
                    need &= ~__ballot_sync(0xffffffffu, mine);
                }
                if (valid) out[y * out_pitch + x] = (OutT)result;
            }
        }
    });
}

template <int H, typename OutT>
int launch(const void* in, int64_t in_pitch, void* out, int64_t out_pitch, int64_t ny, int64_t nx, int min_count,
           cudaStream_t stream)
{
    constexpr int CW = TW + 2 * H, IN_W = TW + 2 * hd_halo_x(H, 4), IN_H = TH + 2 * H;
    constexpr size_t STAGE = (IN_W * IN_H * 4 + 127) / 128 * 128;
    constexpr size_t SMEM = 2 * STAGE + 2 * TH * CW * 4 + 2 * TH * CW + TH * (CW / 2) * 5;
    CUtensorMap tm;
    if (int e = hd_make_tmap_2d(&tm, in, HD_F32, ny, nx, in_pitch, IN_W, IN_H, false)) return e;
    const int tiles_x = hd_cdiv(nx, TW), tiles_y = hd_cdiv(ny, TH), ntiles = tiles_x * tiles_y;
    HD_CUDA_OK(cudaFuncSetAttribute(majority_kernel<H, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    hd_prof_begin("majority_kernel", stream);
    majority_kernel<H, OutT><<<grid_for(ntiles, 2), MNT, SMEM, stream>>>(tm, (OutT*)out, out_pitch, ny, nx, min_count,
                                                                        tiles_x, ntiles);
    HD_LAUNCH_CHECK();
    hd_count_launch();
    return HD_OK;
}

template <typename OutT>
int dispatch(int h, const void* in, int64_t in_pitch, void* out, int64_t out_pitch, int64_t ny, int64_t nx, int min_count,
             cudaStream_t s)
{
    switch (h) {
        case 1: return launch<1, OutT>(in, in_pitch, out, out_pitch, ny, nx, min_count, s);
        case 2: return launch<2, OutT>(in, in_pitch, out, out_pitch, ny, nx, min_count, s);
        case 3: return launch<3, OutT>(in, in_pitch, out, out_pitch, ny, nx, min_count, s);
        case 4: return launch<4, OutT>(in, in_pitch, out, out_pitch, ny, nx, min_count, s);
        case 5: return launch<5, OutT>(in, in_pitch, out, out_pitch, ny, nx, min_count, s);
        case 6: return launch<6, OutT>(in, in_pitch, out, out_pitch, ny, nx, min_count, s);
        case 7: return launch<7, OutT>(in, in_pitch, out, out_pitch, ny, nx, min_count, s);
        default: return HD_ERR_UNSUPPORTED;
    }
}

}  // namespace

extern "C" int hd_majority(const void* in, int64_t in_pitch, void* out, int out_dtype, int64_t out_pitch, int64_t ny,
                           int64_t nx, int ws, int min_count, void* stream)
{
    if (!in || !out) return HD_ERR_NULL;
    if (int e = check_window(ny, nx, ws)) return e;
    if (in_pitch < nx || out_pitch < nx || min_count < 1) return HD_ERR_ARG;
    // Boyer-Moore needs the winner to hold an absolute majority of the ws*ws-4 window cells; the
    // reference threshold (ws^2-1)*0.7 always does.
    if (2 * min_count <= ws * ws - 4) return HD_ERR_UNSUPPORTED;
    const int h = ws / 2;
    cudaStream_t s = (cudaStream_t)stream;
    if (out_dtype == HD_F32) return dispatch<float>(h, in, in_pitch, out, out_pitch, ny, nx, min_count, s);
    if (out_dtype == HD_F64) return dispatch<double>(h, in, in_pitch, out, out_pitch, ny, nx, min_count, s);
    return HD_ERR_UNSUPPORTED;
}
