// Spectrum-peak detector of the Fourier stage and the mask assembly.
//
//   hd_hollow_mean_detect    BlanksFourier.apply        filters/custom_filters.py:395-427
//   hd_fourier_mask_assemble FourierProcessQuarters     filters/custom_filters.py:968-1050
//
// BlanksFourier visits EVERY cell of a spectrum quarter with a 55x55 window clipped to the quarter (NaN
// padding, sliding_window.py:400-418) whose central 5x5 block is blanked (InnerWindow + NoCenterWindow),
// takes np.nanmean of what is left and flags the cell when centre > 4 * mean.
//
// Kernel: a 64x64 tile with a 27-cell halo is staged by TMA (zero fill outside the quarter = "not counted"),
// a float64 integral image of the 118x120 box is built in shared memory (column sums as 4 row segments x 120
// columns with segment offsets, then one thread per row along the columns -- both conflict free), and each output
// needs 8 integral-image reads: big box minus inner box.  The tile's own cells are copied aside first, so the TMA
// load of the next tile overlaps the row pass and the outputs.  The number of valid
// cells is analytic (clipped 55x55 minus clipped 5x5).  float64 sums of float32 data are exact to ~1e-16, the
// reference's float32 pairwise nanmean to ~1e-7: the stage is tolerance class, the mask is expected to be
// identical (mismatch count reported by the tests).  NaN cells INSIDE the quarter are not skipped (they
// poison the windows that contain them); |F| is NaN-free unless the DEM itself has NaN.
//
// Algorithmic HBM traffic per pass: 4 B read + 4 B (modified image) + 1 B (mask) written (+1 B previous mask).
#include "common.cuh"
#include "tile_common.cuh"

namespace {

constexpr int BT = 64;          // tile edge (outputs)
constexpr int BNT = 512;        // threads per CTA
constexpr int WS = 55, INNER = 5, H = WS / 2, HI = INNER / 2;
constexpr int HX = hd_halo_x(H, 4);                 // 28
constexpr int XOFF = HX - H;                        // 1
constexpr int IN_W = BT + 2 * HX;                   // 120
constexpr int IN_H = BT + 2 * H;                    // 118
constexpr int IS = IN_W + 1;                        // integral-image row stride (odd -> conflict-free column scans)
constexpr uint32_t STAGE = (IN_W * IN_H * 4 + 127) / 128 * 128;
constexpr int NSEG = 4, SEG_ROWS = (IN_H + NSEG - 1) / NSEG;          // column scans run as NSEG row segments
constexpr size_t I_BYTES = (size_t)(IN_H + 1) * IS * sizeof(double);
constexpr size_t CTR_BYTES = (size_t)BT * BT * sizeof(float);
constexpr size_t OFF_BYTES = (size_t)NSEG * IN_W * sizeof(double);
constexpr int CSEG = 4, SEG_COLS = IN_W / CSEG;                       // row scans run as CSEG column segments too
constexpr size_t ROFF_BYTES = (size_t)CSEG * IN_H * sizeof(double);
constexpr size_t SMEM = STAGE + I_BYTES + CTR_BYTES + OFF_BYTES + ROFF_BYTES;
static_assert(NSEG * IN_W <= BNT, "one thread per (segment, column)");
static_assert(CSEG * IN_H <= BNT && IN_W % CSEG == 0, "one thread per (row, column segment)");

template <typename MaskT>
__global__ void __launch_bounds__(BNT, 1) hollow_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                        const uint8_t* __restrict__ mask_prev, int64_t prev_pitch,
                                                        MaskT* __restrict__ mask_out, int64_t mask_pitch,
                                                        float* __restrict__ modified, int64_t mod_pitch, int64_t ny,
                                                        int64_t nx, float factor, int tiles_x, int ntiles,
                                                        uint8_t* __restrict__ flags_out = nullptr,
                                                        const uint8_t* __restrict__ flags_in = nullptr)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    const float* tile = reinterpret_cast<const float*>(smem);
    double* I = reinterpret_cast<double*>(smem + STAGE);       // [(IN_H + 1)][IS], I[r][c] = sum tile[<r][<c]
    float* ctrs = reinterpret_cast<float*>(smem + STAGE + I_BYTES);             // the tile's own 64 x 64 cells
    double* off = reinterpret_cast<double*>(smem + STAGE + I_BYTES + CTR_BYTES);   // [NSEG][IN_W] segment offsets
    double* roff = off + NSEG * IN_W;                                              // [CSEG][IN_H] row-segment totals
    // Second BlanksFourier pass (flags_in): it runs on image * (1 - mask of the first pass), which differs from the image
    // only at the first pass's hits.  A cell whose 55 x 55 window holds no such hit sees the same centre and the same mean as
    // in the first pass, where it was not a hit -- so new hits can only appear within 27 cells of an old one, and a 64 x 64
    // tile needs the second pass only if its 3 x 3 tile neighbourhood had a hit in the first (flags_out of that pass).
    // Peaks are rare (1e-7 .. 1e-4 of the bins): every other tile keeps the previous mask, which mask_widen_kernel (a
    // streaming pass at full occupancy, launched right before this kernel) has already written into mask_out.  Same bits.
    const int tiles_y = ntiles / tiles_x;
    auto active = [&](int t) -> bool {
        if (!flags_in) return true;
        const int ty = t / tiles_x, tx = t - ty * tiles_x;
        int any = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int yy = ty + dy, xx = tx + dx;
                if (yy >= 0 && yy < tiles_y && xx >= 0 && xx < tiles_x) any |= __ldg(flags_in + yy * tiles_x + xx);
            }
        return any != 0;
    };
    auto advance = [&](int t) -> int {                       // skipped tiles keep the previous mask (mask_widen_kernel)
        while (t < ntiles && !active(t)) t += gridDim.x;
        return t;
    };
    for (int t = threadIdx.x; t < IS; t += BNT) I[t] = 0.0;                    // row 0
    for (int t = threadIdx.x; t <= IN_H; t += BNT) I[t * IS] = 0.0;            // column 0
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tm_in);
    }
    __syncthreads();
    auto issue = [&](int tl) {
        const int ty0 = (tl / tiles_x) * BT, tx0 = (tl % tiles_x) * BT;
        mbar_arrive_expect_tx(&bar, (uint32_t)(IN_W * IN_H * 4));
        tma_load_2d(smem, &tm_in, tx0 - HX, ty0 - H, &bar);
    };
    int tl = advance(blockIdx.x);
    if (threadIdx.x == 0 && tl < ntiles) issue(tl);
    for (int k = 0; tl < ntiles; ++k) {
        const int ty0 = (tl / tiles_x) * BT, tx0 = (tl % tiles_x) * BT;
        // the previous pass's mask for this thread's outputs: requested now, needed at the very end of the tile
        uint8_t prevs[BT * BT / BNT];
#pragma unroll
        for (int rep = 0; rep < BT * BT / BNT; ++rep) {
            const int idx = rep * BNT + threadIdx.x;
            const int64_t y = ty0 + idx / BT, x = tx0 + idx % BT;
            prevs[rep] = (mask_prev && y < ny && x < nx) ? __ldg(mask_prev + y * prev_pitch + x) : (uint8_t)0;
        }
        mbar_wait(&bar, k & 1);
        // ---- pass 1: column sums, one thread per (row segment, column); lanes = consecutive columns -----------------
        //      I[r+1][c+1] = sum of tile[r'][c] over the rows r' <= r of the SAME segment, off = totals of the segments above
        if (threadIdx.x < NSEG * IN_W) {
            const int seg = threadIdx.x / IN_W, c = threadIdx.x - seg * IN_W;
            const int r0 = seg * SEG_ROWS, r1 = r0 + SEG_ROWS < IN_H ? r0 + SEG_ROWS : IN_H;
            double acc = 0.0;
#pragma unroll 6
            for (int r = r0; r < r1; ++r) {
                acc += (double)tile[r * IN_W + c];
                I[(r + 1) * IS + c + 1] = acc;
            }
            off[seg * IN_W + c] = acc;
        }
#pragma unroll
        for (int rep = 0; rep < BT * BT / BNT; ++rep) {
            const int idx = rep * BNT + threadIdx.x;
            ctrs[idx] = tile[(idx / BT + H) * IN_W + (idx % BT) + XOFF + H];
        }
        __syncthreads();
        // the staged tile is no longer needed: fetch the next one underneath the rest of this tile's work
        const int tl_next = advance(tl + gridDim.x);
        if (threadIdx.x == 0 && tl_next < ntiles) issue(tl_next);
        if (threadIdx.x < IN_W) {                                  // segment totals -> exclusive offsets
            double run = 0.0;
#pragma unroll
            for (int sg = 0; sg < NSEG; ++sg) {
                const double tot = off[sg * IN_W + threadIdx.x];
                off[sg * IN_W + threadIdx.x] = run;
                run += tot;
            }
        }
        __syncthreads();
        // ---- pass 2: row prefix sums as CSEG column segments per row (472 threads x 30 dependent adds instead of 118
        //      threads x 120: this scan was 40 % of a tile's time); the column offsets of pass 1 are added on the way, the
        //      totals of the segments to the left are added in a third, fully parallel sweep
        if (threadIdx.x < CSEG * IN_H) {
            const int sg = threadIdx.x / IN_H, r = threadIdx.x - sg * IN_H;          // lanes = consecutive rows
            double* row = I + (r + 1) * IS + 1 + sg * SEG_COLS;
            const double* o = off + (r / SEG_ROWS) * IN_W + sg * SEG_COLS;
            double acc = 0.0;
#pragma unroll 6
            for (int c = 0; c < SEG_COLS; ++c) {
                acc += row[c] + o[c];
                row[c] = acc;
            }
            roff[sg * IN_H + r] = acc;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < IN_H * (IN_W - SEG_COLS); t += BNT) {
            const int r = t / (IN_W - SEG_COLS), c = SEG_COLS + (t - r * (IN_W - SEG_COLS));
            const int sg = c / SEG_COLS;
            double add = roff[r];
            if (sg > 1) add += roff[IN_H + r];
            if (sg > 2) add += roff[2 * IN_H + r];
            I[(r + 1) * IS + 1 + c] += add;
        }
        __syncthreads();
        // ---- outputs ----------------------------------------------------------------------------------------
        bool hit_any = false;
        // (a tile whose windows are never clipped has the same number of valid cells everywhere: 55^2 - 5^2)
        const bool interior = ty0 >= H && ty0 + BT - 1 + H <= ny - 1 && tx0 >= H && tx0 + BT - 1 + H <= nx - 1;
#pragma unroll
        for (int rep = 0; rep < BT * BT / BNT; ++rep) {
            const int idx = rep * BNT + threadIdx.x;
            const int ro = idx / BT, xo = idx % BT;
            const int64_t y = ty0 + ro, x = tx0 + xo;
            if (y >= ny || x >= nx) continue;
            const int r0 = ro, r1 = ro + WS, c0 = xo + XOFF, c1 = c0 + WS;
            const double big = (I[r1 * IS + c1] - I[r0 * IS + c1]) - (I[r1 * IS + c0] - I[r0 * IS + c0]);
            const int q0 = r0 + H - HI, q1 = q0 + INNER, d0 = c0 + H - HI, d1 = d0 + INNER;
            const double inner = (I[q1 * IS + d1] - I[q0 * IS + d1]) - (I[q1 * IS + d0] - I[q0 * IS + d0]);
            // valid cells: clipped 55x55 minus clipped 5x5 (the NaN padding is never counted by nanmean)
            auto span = [](int64_t p, int64_t n, int h) {
                const int64_t lo = p - h < 0 ? 0 : p - h, hi = p + h > n - 1 ? n - 1 : p + h;
                return (int)(hi - lo + 1);
            };
            const int cnt = interior ? WS * WS - INNER * INNER
                                     : span(y, ny, H) * span(x, nx, H) - span(y, ny, HI) * span(x, nx, HI);
            const float mean = (float)((big - inner) / (double)cnt);               // np.nanmean -> float32  (:421)
            const float ctr = ctrs[idx];
            const bool hit = ctr > __fmul_rn(factor, mean);                        // centre > 4 * mean      (:424)
            hit_any |= hit;
            const uint8_t prev = prevs[rep];
            mask_out[y * mask_pitch + x] = (MaskT)(prev + (hit ? 1 : 0));          // final_mask += filtered (:461)
            if (modified) modified[y * mod_pitch + x] = __fmul_rn(ctr, hit ? 0.f : 1.f);   // image * (1 - mask) (:426)
        }
        const int tile_hit = __syncthreads_or(hit_any);            // (also: I, ctrs and off are rewritten by the next tile)
        if (flags_out && threadIdx.x == 0) flags_out[tl] = tile_hit ? 1 : 0;
        tl = tl_next;
    }
}

// mask_out = (MaskT) mask_prev (or 0) for the whole raster: eight cells per thread, one 8-byte load, two 16-byte stores
template <typename MaskT>
__global__ void __launch_bounds__(256) mask_widen_kernel(const uint8_t* __restrict__ prev, int64_t prev_pitch,
                                                         MaskT* __restrict__ out, int64_t out_pitch, int64_t ny, int64_t nx)
{
    const int64_t nx8 = (nx + 7) / 8;
    for (CellIter it(nx8); it.y < ny; it.next()) {
        const int64_t y = it.y, x = 8 * it.x;
        MaskT* q = out + y * out_pitch + x;
        const uint8_t* pp = prev ? prev + y * prev_pitch + x : nullptr;
        if (x + 7 < nx && sizeof(MaskT) == 4 && ((reinterpret_cast<uintptr_t>(q) & 15) == 0) &&
            (!pp || (reinterpret_cast<uintptr_t>(pp) & 7) == 0)) {
            const uint2 b = pp ? __ldg(reinterpret_cast<const uint2*>(pp)) : make_uint2(0u, 0u);
            st_cs_f32x4(reinterpret_cast<float*>(q), (float)(b.x & 255u), (float)((b.x >> 8) & 255u),
                        (float)((b.x >> 16) & 255u), (float)(b.x >> 24));
            st_cs_f32x4(reinterpret_cast<float*>(q) + 4, (float)(b.y & 255u), (float)((b.y >> 8) & 255u),
                        (float)((b.y >> 16) & 255u), (float)(b.y >> 24));
        } else {
            for (int j = 0; j < 8 && x + j < nx; ++j) q[j] = (MaskT)(pp ? __ldg(pp + j) : (uint8_t)0);
        }
    }
}

// ---- mask assembly ------------------------------------------------------------------------------------------
// quarters[0] -> top-left block [:my-m, :mx-m]; quarters[1] -> top-right block, shifted by the margin;
// bottom blocks are the point mirrors; the middle row / column of odd sizes stays 0.  (:986-991, :1002-1049)
// Row-band variant (hydrodem_b200/sharding.py): only the `out_rows` rows listed by the K layout (a, b) are produced --
// local row t holds the spectrum row ky(t), i.e. the shifted row (ky + ny/2) % ny -- from quarter-mask slabs that start
// at quarter row q0.  (out_rows == 0: the whole mask, row t = shifted row t.)
template <typename OutT>
__global__ void __launch_bounds__(256) assemble_kernel(const uint8_t* __restrict__ m1, int64_t p1,
                                                       const uint8_t* __restrict__ m2, int64_t p2, OutT* __restrict__ out,
                                                       int64_t out_pitch, int ny, int nx, int margin, int invert,
                                                       int out_rows = 0, int ka = 0, int kb = 0, int q0 = 0,
                                                       int q_rows = 0x7fffffff)
{
    const int my = ny / 2, y_odd = ny & 1, mx = nx / 2, x_odd = nx & 1;
    const int nxq = (nx + 3) / 4;                                  // four consecutive cells of a row per thread
    const int nrows = out_rows > 0 ? out_rows : ny;
    // K layout (see csrc/fft.cu): rows ky in [ka, kb), then their mirrors ny - ky in ascending order
    const int kmin = ka > 1 ? ka : 1, kmax = (kb - 1) < (ny - 1) / 2 ? (kb - 1) : (ny - 1) / 2, hlo = ny - kmax;
    for (CellIter it(nxq); it.y < nrows; it.next()) {
        const int t = (int)it.y, x4 = 4 * (int)it.x;
        int y = t;
        if (out_rows > 0) {
            const int ky = t < kb - ka ? ka + t : hlo + (t - (kb - ka));
            y = ky + ny / 2; if (y >= ny) y -= ny;
        }
        (void)kmin;
        // row part of the block lookup: position inside one of the (my, mx) blocks, or -1 on the odd middle row
        int by = -1, top = 0;
        if (y < my) { by = y; top = 1; } else if (y >= my + y_odd) { by = y - my - y_odd; }
        // bottom blocks: c3 = flip(c2) sits bottom-left, c4 = flip(c1) bottom-right
        const int qy = top ? by : my - 1 - by;
        const bool row_ok = by >= 0 && qy < my - margin && qy >= q0 && qy - q0 < q_rows;
        OutT res[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x4 + j;
            int v = 0;
            if (row_ok && x < nx) {
                int bx = -1, left = 0;
                if (x < mx) { bx = x; left = 1; } else if (x >= mx + x_odd) { bx = x - mx - x_odd; }
                if (bx >= 0) {
                    const int qx = top ? bx : mx - 1 - bx;
                    const bool use_first = top ? left : !left;
                    if (use_first) {
                        if (qx < mx - margin) v = m1[(int64_t)(qy - q0) * p1 + qx];
                    } else {
                        if (qx >= margin) v = m2[(int64_t)(qy - q0) * p2 + (qx - margin)];
                    }
                }
            }
            res[j] = invert ? (OutT)(1 - v) : (OutT)v;
        }
        store4v<OutT>(out, out_pitch, t, x4, nx, res);
    }
}

}  // namespace

static int hollow_run(const void* in, int64_t in_pitch, const void* mask_prev, int64_t prev_pitch, void* mask_out,
                      int mask_dtype, int64_t mask_pitch, void* modified, int64_t mod_pitch, int64_t ny, int64_t nx, int ws,
                      int inner, double factor, void* flags_out, const void* flags_in, void* stream)
{
    if (!in || !mask_out) return HD_ERR_NULL;
    if (ws > ny || ws > nx) return HD_ERR_WINDOW_HIGH;
    if (ws % 2 != 1) return HD_ERR_WINDOW_EVEN;
    if (ws != WS || inner != INNER) return HD_ERR_UNSUPPORTED;     // the reference hard-codes 55 / 5 (:417-419, :457)
    if (mask_dtype != HD_U8 && mask_dtype != HD_F32) return HD_ERR_UNSUPPORTED;
    if (in_pitch < nx || mask_pitch < nx || (modified && mod_pitch < nx) || (mask_prev && prev_pitch < nx)) return HD_ERR_ARG;
    CUtensorMap tm;
    if (int e = hd_make_tmap_2d(&tm, in, HD_F32, ny, nx, in_pitch, IN_W, IN_H, false)) return e;
    const int tiles_x = hd_cdiv(nx, BT), tiles_y = hd_cdiv(ny, BT), ntiles = tiles_x * tiles_y;
    const int grid = ntiles < hd_num_sms() ? ntiles : hd_num_sms();
    if (flags_in) {                                            // the restricted pass leaves untouched tiles to this copy
        const int64_t total = ny * ((nx + 7) / 8);
        const int blocks = (int)((total + 255) / 256 < (int64_t)hd_num_sms() * 16 ? (total + 255) / 256 : hd_num_sms() * 16);
        hd_prof_begin("mask_widen_kernel", (cudaStream_t)stream);
        if (mask_dtype == HD_U8)
            mask_widen_kernel<uint8_t><<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)mask_prev, prev_pitch,
                                                                               (uint8_t*)mask_out, mask_pitch, ny, nx);
        else
            mask_widen_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)mask_prev, prev_pitch,
                                                                             (float*)mask_out, mask_pitch, ny, nx);
        HD_LAUNCH_CHECK(); hd_count_launch();
    }
    hd_prof_begin("hollow_kernel", (cudaStream_t)stream);
    if (mask_dtype == HD_U8) {
        HD_CUDA_OK(cudaFuncSetAttribute(hollow_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        hollow_kernel<uint8_t><<<grid, BNT, SMEM, (cudaStream_t)stream>>>(
            tm, (const uint8_t*)mask_prev, prev_pitch, (uint8_t*)mask_out, mask_pitch, (float*)modified, mod_pitch, ny, nx,
            (float)factor, tiles_x, ntiles, (uint8_t*)flags_out, (const uint8_t*)flags_in);
    } else {
        HD_CUDA_OK(cudaFuncSetAttribute(hollow_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        hollow_kernel<float><<<grid, BNT, SMEM, (cudaStream_t)stream>>>(
            tm, (const uint8_t*)mask_prev, prev_pitch, (float*)mask_out, mask_pitch, (float*)modified, mod_pitch, ny, nx,
            (float)factor, tiles_x, ntiles, (uint8_t*)flags_out, (const uint8_t*)flags_in);
    }
    HD_LAUNCH_CHECK();
    hd_count_launch();
    return HD_OK;
}

extern "C" int hd_hollow_mean_detect(const void* in, int64_t in_pitch, const void* mask_prev, int64_t prev_pitch,
                                     void* mask_out, int mask_dtype, int64_t mask_pitch, void* modified, int64_t mod_pitch,
                                     int64_t ny, int64_t nx, int ws, int inner, double factor, void* stream)
{
    return hollow_run(in, in_pitch, mask_prev, prev_pitch, mask_out, mask_dtype, mask_pitch, modified, mod_pitch, ny, nx, ws,
                      inner, factor, nullptr, nullptr, stream);
}

// The two passes of DetectBlanksFourier (custom_filters.py:441-462) with the second one restricted to where it can find
// anything: flags_out (pass 1) receives one byte per 64 x 64 tile, "a hit in this tile"; flags_in (pass 2, modified must be
// NULL) makes the pass run only on tiles whose 3 x 3 tile neighbourhood was flagged -- everywhere else the accumulated mask
// is the previous mask (see hollow_kernel).  hd_hollow_tile_count(ny, nx) bytes per flag array.
extern "C" int64_t hd_hollow_tile_count(int64_t ny, int64_t nx) { return (int64_t)hd_cdiv(ny, BT) * hd_cdiv(nx, BT); }

extern "C" int hd_hollow_mean_detect_tiles(const void* in, int64_t in_pitch, const void* mask_prev, int64_t prev_pitch,
                                           void* mask_out, int mask_dtype, int64_t mask_pitch, void* modified,
                                           int64_t mod_pitch, int64_t ny, int64_t nx, int ws, int inner, double factor,
                                           void* flags_out, const void* flags_in, void* stream)
{
    if (flags_in && modified) return HD_ERR_ARG;
    return hollow_run(in, in_pitch, mask_prev, prev_pitch, mask_out, mask_dtype, mask_pitch, modified, mod_pitch, ny, nx, ws,
                      inner, factor, flags_out, flags_in, stream);
}

extern "C" int hd_fourier_mask_assemble(const void* q1, int64_t q1_pitch, const void* q2, int64_t q2_pitch, void* out,
                                        int out_dtype, int64_t out_pitch, int64_t ny, int64_t nx, int margin, int invert,
                                        void* stream)
{
    if (!q1 || !q2 || !out) return HD_ERR_NULL;
    if (ny < 2 || nx < 2 || ny > 0x7fffffff || nx > 0x7fffffff || out_pitch < nx || margin < 0) return HD_ERR_ARG;
    if (ny / 2 - margin < 1 || nx / 2 - margin < 1) return HD_ERR_ARG;
    if (q1_pitch < nx / 2 - margin || q2_pitch < nx / 2 - margin) return HD_ERR_ARG;
    const int64_t total = ny * ((nx + 3) / 4);
    const int blocks = (int)((total + 255) / 256 < (int64_t)hd_num_sms() * 16 ? (total + 255) / 256 : hd_num_sms() * 16);
    cudaStream_t s = (cudaStream_t)stream;
#define HD_ASM(T, TAG)                                                                                              \
    if (out_dtype == TAG) {                                                                                         \
        hd_prof_begin("assemble_kernel", s);                                                                           \
        assemble_kernel<T><<<blocks, 256, 0, s>>>((const uint8_t*)q1, q1_pitch, (const uint8_t*)q2, q2_pitch, (T*)out, \
                                                  out_pitch, (int)ny, (int)nx, margin, invert);                   \
        HD_LAUNCH_CHECK();                                                                                          \
        hd_count_launch();                                                                                          \
        return HD_OK;                                                                                               \
    }
    HD_ASM(uint8_t, HD_U8)
    HD_ASM(float, HD_F32)
    HD_ASM(double, HD_F64)
#undef HD_ASM
    return HD_ERR_UNSUPPORTED;
}

// The same assembly for the rows of one rank of the sharded Fourier stage: out (U8, rows x nx) holds the mask rows of
// the K layout (a, b) (csrc/fft.cu); q1 / q2 are slabs of the quarter masks starting at quarter row q0 (q_rows rows).
extern "C" int hd_fourier_mask_assemble_rows(const void* q1, int64_t q1_pitch, const void* q2, int64_t q2_pitch, int64_t q0,
                                             int64_t q_rows, void* out, int64_t out_pitch, int64_t out_rows, int64_t a,
                                             int64_t b, int64_t ny, int64_t nx, int margin, void* stream)
{
    if (!q1 || !q2 || !out) return HD_ERR_NULL;
    if (ny < 2 || nx < 2 || ny > 0x7fffffff || nx > 0x7fffffff || out_pitch < nx || margin < 0 || out_rows < 1) return HD_ERR_ARG;
    if (ny / 2 - margin < 1 || nx / 2 - margin < 1 || a < 0 || b <= a || b > ny / 2 + 1 || q0 < 0 || q_rows < 0) return HD_ERR_ARG;
    if (q1_pitch < nx / 2 - margin || q2_pitch < nx / 2 - margin) return HD_ERR_ARG;
    const int64_t total = out_rows * ((nx + 3) / 4);
    const int blocks = (int)((total + 255) / 256 < (int64_t)hd_num_sms() * 16 ? (total + 255) / 256 : hd_num_sms() * 16);
    cudaStream_t s = (cudaStream_t)stream;
    hd_prof_begin("assemble_kernel", s);
    assemble_kernel<uint8_t><<<blocks, 256, 0, s>>>((const uint8_t*)q1, q1_pitch, (const uint8_t*)q2, q2_pitch, (uint8_t*)out,
                                                    out_pitch, (int)ny, (int)nx, margin, 0, (int)out_rows, (int)a, (int)b, (int)q0,
                                                    (int)q_rows);
    HD_LAUNCH_CHECK();
    hd_count_launch();
    return HD_OK;
}
