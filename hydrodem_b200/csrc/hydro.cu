// NEW stages N2 / N3 -- sink-fill and D8 flow direction.  Neither exists in the reference:
// parity is against this repo's own oracle (oracle/hydrology.py, oracle/c/hydro_oracle.c), "parity unpinned".
//
// Sink-fill: fixed point of Planchon & Darboux (2001) with eps = 0 and 8-connectivity,
//     W = z on the raster frame, NaN cells are outlets (-inf while iterating, NaN in the result),
//     W(c) = max(z(c), min_{n in N8} W(n))  wherever that lowers W(c).
// The fixed point is unique (it is the minimax path elevation to an outlet), so any update order -- including
// the racy, in-place one used here -- converges to the same bits: the result does not depend on the
// schedule, the tile size or the number of GPUs.
//
// Kernel: one CTA per ACTIVE 64x64 tile.  z (no halo) and W (one-cell halo, NaN fill outside the raster so
// that fminf ignores it) are staged by TMA; four groups of 64 threads march down / up / right / left through
// the tile simultaneously (a marching sweep carries a level across the whole tile in one pass) until a
// __syncthreads_or sees no change.  A changed tile writes W back and re-activates itself and its 8
// neighbours for the next global sweep; the host stops when a sweep changes nothing.  Warp-aggregated: one
// vote per thread group, one atomic per changed tile.
//
// Algorithmic HBM traffic: 4 B (z) + 4 B (W) read + 4 B (W) written per cell of an active tile per sweep.
#include "common.cuh"

namespace {

constexpr int FT = 64;                    // tile edge
constexpr int FNT = 256;
constexpr int WHX = 4;                    // x halo of the W box (16-byte TMA rule), 1 needed
constexpr int WBOX_W = FT + 2 * WHX;      // 72
constexpr int WBOX_H = FT + 2;            // 66
constexpr int WS_STRIDE = FT + 3;         // 67: odd stride -> row marches are bank-conflict free
constexpr uint32_t Z_BYTES = FT * FT * 4;
constexpr uint32_t W_BYTES = WBOX_W * WBOX_H * 4;

struct FillCounters { int changed_tiles; int pad[3]; };

__global__ void __launch_bounds__(256) fill_init_kernel(const float* __restrict__ z, int64_t z_pitch, float* __restrict__ w,
                                                        int64_t w_pitch, int64_t ny, int64_t nx)
{
    const int64_t total = ny * nx;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t y = t / nx, x = t - y * nx;
        const float v = z[y * z_pitch + x];
        float r = __int_as_float(0x7f800000);                                    // +inf inside
        if (v != v) r = __int_as_float(0xff800000);                              // nodata: outlet at -inf
        else if (y == 0 || x == 0 || y == ny - 1 || x == nx - 1) r = v;          // frame: W = z
        w[y * w_pitch + x] = r;
    }
}

__global__ void __launch_bounds__(256) fill_finish_kernel(const float* __restrict__ z, int64_t z_pitch,
                                                          float* __restrict__ w, int64_t w_pitch, int64_t ny, int64_t nx)
{
    const int64_t total = ny * nx;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t y = t / nx, x = t - y * nx;
        const float v = z[y * z_pitch + x];
        if (v != v) w[y * w_pitch + x] = v;                                      // nodata stays nodata
    }
}

__device__ __forceinline__ float relax(const float* ws, const float* zs, int r, int c)
{
    // ws is the padded (FT+2) x WS_STRIDE array, cell (r, c) of the tile lives at ws[(r+1)*WS_STRIDE + c+1]
    const float* p = ws + (r + 1) * WS_STRIDE + (c + 1);
    float m = fminf(fminf(p[-WS_STRIDE - 1], p[-WS_STRIDE]), fminf(p[-WS_STRIDE + 1], p[-1]));
    m = fminf(m, fminf(fminf(p[1], p[WS_STRIDE - 1]), fminf(p[WS_STRIDE], p[WS_STRIDE + 1])));
    return fmaxf(zs[r * FT + c], m);            // fminf / fmaxf skip NaN operands
}

__global__ void __launch_bounds__(FNT) fill_sweep_kernel(const __grid_constant__ CUtensorMap tm_z,
                                                         const __grid_constant__ CUtensorMap tm_w, float* __restrict__ w,
                                                         int64_t w_pitch, int64_t ny, int64_t nx, int tiles_x, int tiles_y,
                                                         const int* __restrict__ active_in, int* __restrict__ active_out,
                                                         FillCounters* counters)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    const int tile = blockIdx.x;
    if (!active_in[tile]) return;
    float* zs = reinterpret_cast<float*>(smem);                           // [FT][FT]
    float* wbox = reinterpret_cast<float*>(smem + Z_BYTES);               // [WBOX_H][WBOX_W] as loaded
    float* ws = reinterpret_cast<float*>(smem + Z_BYTES + W_BYTES);       // [(FT+2)][WS_STRIDE] padded copy
    const int ty0 = (tile / tiles_x) * FT, tx0 = (tile % tiles_x) * FT;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(&bar, Z_BYTES + W_BYTES);
        tma_load_2d(zs, &tm_z, tx0, ty0, &bar);
        tma_load_2d(wbox, &tm_w, tx0 - WHX, ty0 - 1, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int t = threadIdx.x; t < WBOX_H * (FT + 2); t += FNT) {
        const int r = t / (FT + 2), c = t - r * (FT + 2);
        ws[r * WS_STRIDE + c] = wbox[r * WBOX_W + c + WHX - 1];
    }
    __syncthreads();

    const int group = threadIdx.x >> 6, lane64 = threadIdx.x & 63;
    bool tile_changed = false;
    for (int iter = 0; iter < 4096; ++iter) {
        bool changed = false;
        for (int step = 0; step < FT; ++step) {
            int r, c;
            if (group == 0) { r = step; c = lane64; }                     // marching down
            else if (group == 1) { r = FT - 1 - step; c = lane64; }       // up
            else if (group == 2) { r = lane64; c = step; }                // right
            else { r = lane64; c = FT - 1 - step; }                       // left
            float* cell = ws + (r + 1) * WS_STRIDE + (c + 1);
            const float cand = relax(ws, zs, r, c);
            if (cand < *cell) { *cell = cand; changed = true; }           // only ever lowers W
        }
        if (!__syncthreads_or(changed)) break;
        tile_changed = true;
    }
    if (tile_changed) {
        for (int t = threadIdx.x; t < FT * FT; t += FNT) {
            const int r = t >> 6, c = t & 63;
            const int64_t y = ty0 + r, x = tx0 + c;
            if (y < ny && x < nx) w[y * w_pitch + x] = ws[(r + 1) * WS_STRIDE + c + 1];
        }
        if (threadIdx.x < 9) {
            const int dy = (int)threadIdx.x / 3 - 1, dx = (int)threadIdx.x % 3 - 1;
            const int tyy = tile / tiles_x + dy, txx = tile % tiles_x + dx;
            if (tyy >= 0 && tyy < tiles_y && txx >= 0 && txx < tiles_x) active_out[tyy * tiles_x + txx] = 1;
        }
        if (threadIdx.x == 0) atomicAdd(&counters->changed_tiles, 1);
    }
}

// ---- D8 ------------------------------------------------------------------------------------------------------
// ESRI codes E=1 SE=2 S=4 SW=8 W=16 NW=32 N=64 NE=128; steepest positive drop, diagonal drops scaled by
// 0.70710678f in float32; ties keep the first in that order; frame cells, NaN centres, no drop -> 0.
__global__ void __launch_bounds__(256) d8_kernel(const float* __restrict__ w, int64_t w_pitch, uint8_t* __restrict__ out,
                                                 int64_t out_pitch, int64_t ny, int64_t nx)
{
    const int64_t total = ny * nx;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t y = t / nx, x = t - y * nx;
        uint8_t code = 0;
        if (y > 0 && x > 0 && y < ny - 1 && x < nx - 1) {
            const float* p = w + y * w_pitch + x;
            const float c = p[0];
            const float nb[8] = {p[1], p[w_pitch + 1], p[w_pitch], p[w_pitch - 1], p[-1], p[-w_pitch - 1], p[-w_pitch],
                                 p[-w_pitch + 1]};
            float best = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float drop = __fsub_rn(c, nb[k]);
                if (k & 1) drop = __fmul_rn(drop, 0.70710678f);
                if (drop > best) { best = drop; code = (uint8_t)(1u << k); }
            }
        }
        out[y * out_pitch + x] = code;
    }
}

int stream_grid(int64_t total)
{
    const int64_t b = (total + 255) / 256, cap = (int64_t)hd_num_sms() * 16;
    return (int)(b < cap ? b : cap);
}

}  // namespace

extern "C" int64_t hd_pdfill_workspace_bytes(int64_t ny, int64_t nx)
{
    const int64_t ntiles = (int64_t)hd_cdiv(ny, FT) * hd_cdiv(nx, FT);
    return 2 * ntiles * (int64_t)sizeof(int) + 256;
}

extern "C" int hd_pdfill(const void* z, int64_t z_pitch, void* w, int64_t w_pitch, int64_t ny, int64_t nx, void* workspace,
                         int64_t workspace_bytes, int max_sweeps, int* sweeps_out, void* stream)
{
    if (!z || !w || !workspace) return HD_ERR_NULL;
    if (ny < 1 || nx < 1 || z_pitch < nx || w_pitch < nx) return HD_ERR_ARG;
    if (workspace_bytes < hd_pdfill_workspace_bytes(ny, nx)) return HD_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int tiles_x = hd_cdiv(nx, FT), tiles_y = hd_cdiv(ny, FT), ntiles = tiles_x * tiles_y;
    FillCounters* counters = (FillCounters*)workspace;
    int* flags_a = (int*)((char*)workspace + 256);
    int* flags_b = flags_a + ntiles;
    CUtensorMap tm_z, tm_w;
    if (int e = hd_make_tmap_2d(&tm_z, z, HD_F32, ny, nx, z_pitch, FT, FT, false)) return e;
    if (int e = hd_make_tmap_2d(&tm_w, w, HD_F32, ny, nx, w_pitch, WBOX_W, WBOX_H, true)) return e;
    const size_t smem = Z_BYTES + W_BYTES + (size_t)(FT + 2) * WS_STRIDE * 4;
    HD_CUDA_OK(cudaFuncSetAttribute(fill_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    hd_prof_begin("fill_init_kernel", s);
    fill_init_kernel<<<stream_grid(ny * nx), 256, 0, s>>>((const float*)z, z_pitch, (float*)w, w_pitch, ny, nx);
    HD_LAUNCH_CHECK(); hd_count_launch();
    // every tile starts active
    HD_CUDA_OK(cudaMemsetAsync(flags_a, 1, (size_t)ntiles * sizeof(int), s));   // any non-zero byte pattern = active
    static int* h_changed = nullptr;
    if (!h_changed) HD_CUDA_OK(cudaHostAlloc((void**)&h_changed, sizeof(int), cudaHostAllocDefault));
    int sweeps = 0, rc = HD_OK;
    int* fin = flags_a;
    int* fout = flags_b;
    if (max_sweeps <= 0) max_sweeps = 1 << 30;
    const int batch = 4;                      // sweeps between host convergence checks
    for (;;) {
        int last_changed_ptr_valid = 0;
        for (int b = 0; b < batch && sweeps < max_sweeps; ++b) {
            HD_CUDA_OK(cudaMemsetAsync(fout, 0, (size_t)ntiles * sizeof(int), s));
            HD_CUDA_OK(cudaMemsetAsync(counters, 0, sizeof(FillCounters), s));
            hd_prof_begin("fill_sweep_kernel", s);
            fill_sweep_kernel<<<ntiles, FNT, smem, s>>>(tm_z, tm_w, (float*)w, w_pitch, ny, nx, tiles_x, tiles_y, fin, fout,
                                                       counters);
            HD_LAUNCH_CHECK(); hd_count_launch();
            int* tmp = fin; fin = fout; fout = tmp;
            ++sweeps;
            last_changed_ptr_valid = 1;
        }
        if (!last_changed_ptr_valid) { rc = HD_OK; break; }
        // the counter of the LAST sweep of the batch: zero means that sweep found a global fixed point
        HD_CUDA_OK(cudaMemcpyAsync(h_changed, &counters->changed_tiles, sizeof(int), cudaMemcpyDeviceToHost, s));
        HD_CUDA_OK(cudaStreamSynchronize(s));
        if (*h_changed == 0) break;
        if (sweeps >= max_sweeps) { rc = HD_ERR_UNSUPPORTED; break; }        // did not converge within max_sweeps
    }
    hd_prof_begin("fill_finish_kernel", s);
    fill_finish_kernel<<<stream_grid(ny * nx), 256, 0, s>>>((const float*)z, z_pitch, (float*)w, w_pitch, ny, nx);
    HD_LAUNCH_CHECK(); hd_count_launch();
    if (sweeps_out) *sweeps_out = sweeps;
    return rc;
}

extern "C" int hd_d8(const void* w, int64_t w_pitch, void* out, int64_t out_pitch, int64_t ny, int64_t nx, void* stream)
{
    if (!w || !out) return HD_ERR_NULL;
    if (ny < 1 || nx < 1 || w_pitch < nx || out_pitch < nx) return HD_ERR_ARG;
    hd_prof_begin("d8_kernel", (cudaStream_t)stream);
    d8_kernel<<<stream_grid(ny * nx), 256, 0, (cudaStream_t)stream>>>((const float*)w, w_pitch, (uint8_t*)out, out_pitch, ny,
                                                                    nx);
    HD_LAUNCH_CHECK(); hd_count_launch();
    return HD_OK;
}
