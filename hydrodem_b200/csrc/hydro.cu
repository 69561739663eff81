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
#include <cstdio>
#include <cstdlib>

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
                                                        int64_t w_pitch, int64_t ny, int64_t nx, int top_is_halo = 0,
                                                        int bottom_is_halo = 0, int* __restrict__ tile_has_nodata = nullptr,
                                                        int tiles_x = 0)
{
    const int64_t total = ny * nx;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t y = t / nx, x = t - y * nx;
        const float v = z[y * z_pitch + x];
        float r = __int_as_float(0x7f800000);                                    // +inf inside
        if (v != v) {
            r = __int_as_float(0xff800000);                                      // nodata: outlet at -inf
            if (tile_has_nodata) tile_has_nodata[(y / FT) * tiles_x + (x / FT)] = 1;   // benign race: every writer stores 1
        }
        else if ((y == 0 && !top_is_halo) || x == 0 || (y == ny - 1 && !bottom_is_halo) || x == nx - 1)
            r = v;                                                               // frame: W = z (a band's halo rows are not frame)
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

template <int ZSTRIDE = FT>
__device__ __forceinline__ float relax(const float* ws, const float* zs, int r, int c)
{
    // ws is the padded (FT+2) x WS_STRIDE array, cell (r, c) of the tile lives at ws[(r+1)*WS_STRIDE + c+1]
    const float* p = ws + (r + 1) * WS_STRIDE + (c + 1);
    float m = fminf(fminf(p[-WS_STRIDE - 1], p[-WS_STRIDE]), fminf(p[-WS_STRIDE + 1], p[-1]));
    m = fminf(m, fminf(fminf(p[1], p[WS_STRIDE - 1]), fminf(p[WS_STRIDE], p[WS_STRIDE + 1])));
    return fmaxf(zs[r * ZSTRIDE + c], m);       // fminf / fmaxf skip NaN operands
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

// ---- asynchronous worklist variant (default on one GPU) --------------------------------------------------------------
// The sweep kernel above needs one launch per "ring" of tiles the fill wave crosses (39 sweeps on a 3601^2 tile) and
// visits every active tile once per sweep.  Here a persistent grid of co-resident CTAs pulls tiles from a device-side
// FIFO: a tile that changed pushes exactly the neighbours whose halo it changed, so the wave advances as fast as single
// tiles finish and nothing waits for a grid-wide barrier.  Because the fixed point is unique, the (racy) order in which
// tiles are processed cannot change the result.
//   queue  : ring of tile ids; producers reserve a slot with atomicAdd(tail), consumers take tickets with
//            atomicAdd(head) and wait for "their" slot; `queued[tile]` de-duplicates; `pending` = tiles queued or in
//            flight, 0 means the global fixed point is reached.
//   memory : W is read with ld.global.cg (L2, never a stale L1 line) and written with plain stores followed by
//            __threadfence() before the neighbour is published.
struct FillCtl {
    int head, tail, pending, error;
    unsigned long long visits, changed_visits, iterations;
    int qcap, pad_[3];
};
constexpr int SLOT_EMPTY = -1;
constexpr int SPIN_LIMIT = 1 << 22;
// per-tile state: a tile is never processed by two CTAs at once (a second writer could overwrite a lower W with a
// higher one); a tile that is poked while running is marked dirty and re-queued by its own worker when it finishes
enum { T_IDLE = 0, T_QUEUED = 1, T_RUNNING = 2, T_DIRTY = 3 };

__device__ __forceinline__ void fill_push(FillCtl* ctl, int* slots, int qcap, int tile)
{
    atomicAdd(&ctl->pending, 1);
    const int t = atomicAdd(&ctl->tail, 1);
    *(volatile int*)(slots + (t % qcap)) = tile;
}
__device__ __forceinline__ void fill_poke(FillCtl* ctl, int* slots, int qcap, int* state, int tile)
{
    for (;;) {
        const int st = atomicCAS(&state[tile], T_IDLE, T_QUEUED);
        if (st == T_IDLE) { fill_push(ctl, slots, qcap, tile); return; }
        if (st == T_QUEUED || st == T_DIRTY) return;
        if (atomicCAS(&state[tile], T_RUNNING, T_DIRTY) == T_RUNNING) return;     // else the state moved: retry
    }
}

__global__ void __launch_bounds__(256) fill_seed_kernel(int tiles_x, int tiles_y, int64_t ny, FillCtl* ctl,
                                                        int* __restrict__ slots, int* __restrict__ queued, int edge_rows_only)
{
    // A tile is seeded when it touches the raster frame or contains a nodata cell (fill_init_kernel left a 1 in
    // queued[tile] for those).  Continuing a banded fill, only tiles that hold a refreshed halo row, or read it as
    // their own halo (the row next to it), can change.
    const int ntiles = tiles_x * tiles_y;
    for (int tile = blockIdx.x * blockDim.x + threadIdx.x; tile < ntiles; tile += gridDim.x * blockDim.x) {
        const int ty = tile / tiles_x, tx = tile % tiles_x;
        bool seed;
        if (edge_rows_only)
            seed = ((edge_rows_only & 2) && ty == 0) || ((edge_rows_only & 4) && ty >= (int)((ny - 2) / FT));
        else
            seed = ty == 0 || tx == 0 || ty == tiles_y - 1 || tx == tiles_x - 1 || queued[tile] != 0;
        queued[tile] = seed ? T_QUEUED : T_IDLE;
        if (seed) fill_push(ctl, slots, ctl->qcap, tile);
    }
}

__global__ void fill_ctl_init_kernel(FillCtl* ctl, int qcap) { ctl->qcap = qcap; }

constexpr int ZS_STRIDE = FT + 1;         // 65: odd stride, the row-marching groups read z down a column

__global__ void __launch_bounds__(FNT) fill_async_kernel(const float* __restrict__ z, int64_t z_pitch, float* __restrict__ w,
                                                         int64_t w_pitch, int64_t ny, int64_t nx, int tiles_x, int tiles_y,
                                                         FillCtl* ctl, int* slots, int* queued)
{
    // This kernel does not use TMA: W must be read L2-coherently (ld.global.cg) while other CTAs update it, and z
    // has to land in a padded (conflict-free) layout that a dense TMA box cannot produce.
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_tile;
    __shared__ unsigned s_edges;
    float* zs = reinterpret_cast<float*>(smem);                           // [FT][ZS_STRIDE]
    float* wold = zs + FT * ZS_STRIDE;                                    // [(FT+2)][WS_STRIDE] as loaded
    float* ws = wold + (FT + 2) * WS_STRIDE;                              // [(FT+2)][WS_STRIDE] working copy
    const int qcap = ctl->qcap;
    const int group = threadIdx.x >> 6, lane64 = threadIdx.x & 63;
    const float qnan = __int_as_float(0x7fc00000);
    for (;;) {
        // ---- take a ticket and wait for its slot --------------------------------------------------------------------
        if (threadIdx.x == 0) {
            int tile = -1;
            const int my = atomicAdd(&ctl->head, 1);
            volatile int* slot = slots + (my % qcap);
            for (int spin = 0;; ++spin) {
                const int v = *slot;
                if (v != SLOT_EMPTY) { *slot = SLOT_EMPTY; tile = v; break; }
                if (*(volatile int*)&ctl->pending <= 0 || *(volatile int*)&ctl->error) break;
                if (spin > SPIN_LIMIT) { atomicExch(&ctl->error, 1); break; }
                __nanosleep(200);
            }
            if (tile >= 0) { atomicExch(&queued[tile], T_RUNNING); __threadfence(); }
            s_tile = tile;
            s_edges = 0u;
        }
        __syncthreads();
        const int tile = s_tile;
        if (tile < 0) return;
        const int ty0 = (tile / tiles_x) * FT, tx0 = (tile % tiles_x) * FT;
        // z tile (read-only path), zero outside the raster like a TMA box
        const bool zvec_ok = ((z_pitch & 3) == 0) && ((((uintptr_t)z) & 15) == 0);
#pragma unroll 2
        for (int t = threadIdx.x; t < FT * 16; t += FNT) {
            const int r = t >> 4, k = t & 15;
            const int64_t y = (int64_t)ty0 + r, x = (int64_t)tx0 + 4 * k;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y < ny) {
                if (zvec_ok && x + 3 < nx) {
                    v = __ldg(reinterpret_cast<const float4*>(z + y * z_pitch + x));
                } else {
                    if (x < nx) v.x = __ldg(z + y * z_pitch + x);
                    if (x + 1 < nx) v.y = __ldg(z + y * z_pitch + x + 1);
                    if (x + 2 < nx) v.z = __ldg(z + y * z_pitch + x + 2);
                    if (x + 3 < nx) v.w = __ldg(z + y * z_pitch + x + 3);
                }
            }
            float* pz = zs + r * ZS_STRIDE + 4 * k;
            pz[0] = v.x; pz[1] = v.y; pz[2] = v.z; pz[3] = v.w;
        }
        // W with a one-cell halo straight from L2; outside the raster = NaN (ignored by fminf)
        const bool vec_ok = ((w_pitch & 3) == 0) && ((((uintptr_t)w) & 15) == 0);
#pragma unroll 2
        for (int t = threadIdx.x; t < (FT + 2) * 18; t += FNT) {
            // per row of the box: 16 aligned float4 (the 64 tile columns) + the two halo columns
            const int r = t / 18, k = t - r * 18;
            const int64_t y = (int64_t)ty0 - 1 + r;
            const bool yin = y >= 0 && y < ny;
            float* po = wold + r * WS_STRIDE;
            float* pw = ws + r * WS_STRIDE;
            if (k < 16) {
                const int64_t x = (int64_t)tx0 + 4 * k;
                float4 v = make_float4(qnan, qnan, qnan, qnan);
                if (yin && vec_ok && x + 3 < nx) {
                    v = __ldcg(reinterpret_cast<const float4*>(w + y * w_pitch + x));
                } else if (yin) {
                    if (x < nx) v.x = __ldcg(w + y * w_pitch + x);
                    if (x + 1 < nx) v.y = __ldcg(w + y * w_pitch + x + 1);
                    if (x + 2 < nx) v.z = __ldcg(w + y * w_pitch + x + 2);
                    if (x + 3 < nx) v.w = __ldcg(w + y * w_pitch + x + 3);
                }
                const int c = 1 + 4 * k;
                po[c] = v.x; po[c + 1] = v.y; po[c + 2] = v.z; po[c + 3] = v.w;
                pw[c] = v.x; pw[c + 1] = v.y; pw[c + 2] = v.z; pw[c + 3] = v.w;
            } else {
                const int c = (k == 16) ? 0 : FT + 1;
                const int64_t x = (int64_t)tx0 - 1 + c;
                float v = qnan;
                if (yin && x >= 0 && x < nx) v = __ldcg(w + y * w_pitch + x);
                po[c] = v;
                pw[c] = v;
            }
        }
        __syncthreads();
        bool tile_changed = false;
        int iters = 0;
        // marching geometry of this thread group: f = step along the march, l = step to the neighbouring thread
        const int f = (group == 0) ? WS_STRIDE : (group == 1) ? -WS_STRIDE : (group == 2) ? 1 : -1;
        const int l = (group < 2) ? 1 : WS_STRIDE;
        const int zf = (group == 0) ? ZS_STRIDE : (group == 1) ? -ZS_STRIDE : (group == 2) ? 1 : -1;
        const int p0 = (group == 0) ? (1 * WS_STRIDE + lane64 + 1) : (group == 1) ? (FT * WS_STRIDE + lane64 + 1)
                     : (group == 2) ? ((lane64 + 1) * WS_STRIDE + 1) : ((lane64 + 1) * WS_STRIDE + FT);
        const int z0 = (group == 0) ? lane64 : (group == 1) ? ((FT - 1) * ZS_STRIDE + lane64)
                     : (group == 2) ? (lane64 * ZS_STRIDE) : (lane64 * ZS_STRIDE + FT - 1);
        for (int iter = 0; iter < 4096; ++iter) {
            bool changed = false;
            int p = p0, zp = z0;
            // the row behind (b*), the current row's lateral neighbours (cm, cp) and the row ahead (f*): the row
            // ahead of step k is the current row of step k+1 and this thread's own result is the next "behind"
            // centre, so 7 shared-memory reads per cell instead of 10
            float b0 = ws[p - f], cm = ws[p - l], cp = ws[p + l];
            for (int step = 0; step < FT; ++step) {
                const float bm = ws[p - f - l], bp = ws[p - f + l];
                const float fm = ws[p + f - l], f0 = ws[p + f], fp = ws[p + f + l];
                float m = fminf(fminf(fminf(bm, b0), fminf(bp, cm)), fminf(fminf(cp, fm), fminf(f0, fp)));
                const float cand = fmaxf(zs[zp], m);                       // fminf / fmaxf skip NaN operands
                float own = ws[p];
                if (cand < own) { ws[p] = cand; own = cand; changed = true; }   // only ever lowers W
                b0 = own; cm = fm; cp = fp;
                p += f; zp += zf;
            }
            ++iters;
            if (!__syncthreads_or(changed)) break;
            tile_changed = true;
        }
        if (threadIdx.x == 0) atomicAdd(&ctl->iterations, (unsigned long long)iters);
        if (tile_changed) {
            unsigned edges = 0u;
            for (int t = threadIdx.x; t < FT * FT; t += FNT) {
                const int r = t >> 6, c = t & 63;
                const int64_t y = ty0 + r, x = tx0 + c;
                const float nv = ws[(r + 1) * WS_STRIDE + c + 1], ov = wold[(r + 1) * WS_STRIDE + c + 1];
                if (y < ny && x < nx && nv != ov && !(nv != nv)) {
                    w[y * w_pitch + x] = nv;
                    // Poke a neighbouring tile only if this cell can still lower one of ITS cells: the halo ring holds the
                    // neighbour's edge values as loaded (they only ever decrease), so nv >= that value means
                    // max(z, nv) cannot undercut it.  This drops the useless "poke back" to the tile the level came from.
                    // bit = (dy+1)*3 + (dx+1)
                    if (r == 0 || r == FT - 1 || c == 0 || c == FT - 1) {
#pragma unroll
                        for (int a = -1; a <= 1; ++a)
#pragma unroll
                            for (int b = -1; b <= 1; ++b) {
                                const int rr = r + a, cc = c + b;
                                const int dy = rr < 0 ? -1 : (rr >= FT ? 1 : 0), dx = cc < 0 ? -1 : (cc >= FT ? 1 : 0);
                                if ((dy | dx) && nv < wold[(rr + 1) * WS_STRIDE + cc + 1])     // NaN (outside the raster): false
                                    edges |= 1u << ((dy + 1) * 3 + (dx + 1));
                            }
                    }
                }
            }
            if (edges) atomicOr(&s_edges, edges);
            __threadfence();                                   // W stores visible device-wide before neighbours are published
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned edges = s_edges;
            const int ty = tile / tiles_x, tx = tile % tiles_x;
            for (int k = 0; k < 9; ++k) {
                if (!((edges >> k) & 1u)) continue;
                const int tyy = ty + k / 3 - 1, txx = tx + k % 3 - 1;
                if (tyy < 0 || tyy >= tiles_y || txx < 0 || txx >= tiles_x) continue;
                fill_poke(ctl, slots, qcap, queued, tyy * tiles_x + txx);
            }
            atomicAdd(&ctl->visits, 1ull);
            if (tile_changed) atomicAdd(&ctl->changed_visits, 1ull);
            __threadfence();
            if (atomicCAS(&queued[tile], T_RUNNING, T_IDLE) != T_RUNNING) {     // poked while running: go again
                atomicExch(&queued[tile], T_QUEUED);
                fill_push(ctl, slots, qcap, tile);
            }
            __threadfence();
            atomicSub(&ctl->pending, 1);
        }
        __syncthreads();
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


// flags: 1 = W already initialised (continue a banded fill), 2 / 4 = the top / bottom raster row is a halo row of a
// neighbouring band (not frame); finish = restore NaN at nodata cells afterwards
static int pdfill_async(const void* z, int64_t z_pitch, void* w, int64_t w_pitch, int64_t ny, int64_t nx, void* workspace,
                        int* visits_out, cudaStream_t s, int flags = 0, bool finish = true)
{
    const int tiles_x = hd_cdiv(nx, FT), tiles_y = hd_cdiv(ny, FT), ntiles = tiles_x * tiles_y;
    FillCtl* ctl = (FillCtl*)workspace;
    const int qcap = ntiles + 8192;
    int* slots = (int*)((char*)workspace + 256);
    int* queued = slots + qcap;
    const size_t smem = (size_t)FT * ZS_STRIDE * 4 + 2 * (size_t)(FT + 2) * WS_STRIDE * 4;
    HD_CUDA_OK(cudaFuncSetAttribute(fill_async_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    HD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fill_async_kernel, FNT, smem));
    int grid = hd_num_sms() * (per_sm < 1 ? 1 : per_sm);       // every CTA must be co-resident: they wait on each other
    if (grid > ntiles) grid = ntiles;
    // (no host-to-device copy of a stack object here: the whole call must be capturable in a CUDA graph)
    HD_CUDA_OK(cudaMemsetAsync(ctl, 0, sizeof(FillCtl), s));
    HD_CUDA_OK(cudaMemsetAsync(slots, 0xff, (size_t)qcap * sizeof(int), s));        // SLOT_EMPTY = -1
    HD_CUDA_OK(cudaMemsetAsync(queued, 0, (size_t)ntiles * sizeof(int), s));
    fill_ctl_init_kernel<<<1, 1, 0, s>>>(ctl, qcap);
    HD_LAUNCH_CHECK();
    if (!(flags & 1)) {
        hd_prof_begin("fill_init_kernel", s);
        fill_init_kernel<<<stream_grid(ny * nx), 256, 0, s>>>((const float*)z, z_pitch, (float*)w, w_pitch, ny, nx, flags & 2,
                                                             flags & 4, queued, tiles_x);
        HD_LAUNCH_CHECK(); hd_count_launch();
    }
    hd_prof_begin("fill_seed_kernel", s);
    fill_seed_kernel<<<hd_cdiv(ntiles, 256) < 1184 ? hd_cdiv(ntiles, 256) : 1184, 256, 0, s>>>(
        tiles_x, tiles_y, ny, ctl, slots, queued, (flags & 1) ? (flags & 6) : 0);
    HD_LAUNCH_CHECK(); hd_count_launch();
    hd_prof_begin("fill_async_kernel", s);
    fill_async_kernel<<<grid, FNT, smem, s>>>((const float*)z, z_pitch, (float*)w, w_pitch, ny, nx, tiles_x, tiles_y, ctl,
                                              slots, queued);
    HD_LAUNCH_CHECK(); hd_count_launch();
    if (finish) {
        hd_prof_begin("fill_finish_kernel", s);
        fill_finish_kernel<<<stream_grid(ny * nx), 256, 0, s>>>((const float*)z, z_pitch, (float*)w, w_pitch, ny, nx);
        HD_LAUNCH_CHECK(); hd_count_launch();
    }
    if (!visits_out && !getenv("HD_FILL_TRACE")) return HD_OK;      // fully asynchronous when nobody asks for statistics
    static FillCtl* h_ctl = nullptr;
    if (!h_ctl) HD_CUDA_OK(cudaHostAlloc((void**)&h_ctl, sizeof(FillCtl), cudaHostAllocDefault));
    HD_CUDA_OK(cudaMemcpyAsync(h_ctl, ctl, sizeof(FillCtl), cudaMemcpyDeviceToHost, s));
    HD_CUDA_OK(cudaStreamSynchronize(s));
    if (getenv("HD_FILL_TRACE"))
        fprintf(stderr, "pdfill async: %llu tile visits (%llu changed, %llu in-tile iterations) over %d tiles, grid %d\n",
                h_ctl->visits, h_ctl->changed_visits, h_ctl->iterations, ntiles, grid);
    if (visits_out) *visits_out = (int)h_ctl->visits;
    if (h_ctl->error || h_ctl->pending != 0) return HD_ERR_UNSUPPORTED;    // worklist stalled (should not happen)
    return HD_OK;
}

extern "C" int64_t hd_pdfill_workspace_bytes(int64_t ny, int64_t nx)
{
    const int64_t ntiles = (int64_t)hd_cdiv(ny, FT) * hd_cdiv(nx, FT);
    // control block + FIFO ring (ntiles + slack) + per-tile flags; the sweep variant uses two flag arrays
    return 256 + (2 * ntiles + 8192) * (int64_t)sizeof(int) + 2 * ntiles * (int64_t)sizeof(int);
}

extern "C" int hd_pdfill(const void* z, int64_t z_pitch, void* w, int64_t w_pitch, int64_t ny, int64_t nx, void* workspace,
                         int64_t workspace_bytes, int max_sweeps, int* sweeps_out, void* stream)
{
    if (!z || !w || !workspace) return HD_ERR_NULL;
    if (ny < 1 || nx < 1 || z_pitch < nx || w_pitch < nx) return HD_ERR_ARG;
    if (workspace_bytes < hd_pdfill_workspace_bytes(ny, nx)) return HD_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int tiles_x = hd_cdiv(nx, FT), tiles_y = hd_cdiv(ny, FT), ntiles = tiles_x * tiles_y;
    const char* mode = getenv("HD_FILL_MODE");
    if (!(mode && mode[0] == 's') && max_sweeps <= 0)
        return pdfill_async(z, z_pitch, w, w_pitch, ny, nx, workspace, sweeps_out, s);
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
    const bool trace = getenv("HD_FILL_TRACE") != nullptr;
    const int batch = trace ? 1 : 4;          // sweeps between host convergence checks
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
        if (trace) fprintf(stderr, "pdfill sweep %d: %d of %d tiles changed\n", sweeps, *h_changed, ntiles);
        if (*h_changed == 0) break;
        if (sweeps >= max_sweeps) { rc = HD_ERR_UNSUPPORTED; break; }        // did not converge within max_sweeps
    }
    hd_prof_begin("fill_finish_kernel", s);
    fill_finish_kernel<<<stream_grid(ny * nx), 256, 0, s>>>((const float*)z, z_pitch, (float*)w, w_pitch, ny, nx);
    HD_LAUNCH_CHECK(); hd_count_launch();
    if (sweeps_out) *sweeps_out = sweeps;
    return rc;
}

extern "C" int hd_pdfill_band(const void* z, int64_t z_pitch, void* w, int64_t w_pitch, int64_t ny, int64_t nx,
                              void* workspace, int64_t workspace_bytes, int flags, int* visits_out, void* stream)
{
    if (!z || !w || !workspace) return HD_ERR_NULL;
    if (ny < 1 || nx < 1 || z_pitch < nx || w_pitch < nx || (flags & ~7)) return HD_ERR_ARG;
    if (workspace_bytes < hd_pdfill_workspace_bytes(ny, nx)) return HD_ERR_WORKSPACE;
    return pdfill_async(z, z_pitch, w, w_pitch, ny, nx, workspace, visits_out, (cudaStream_t)stream, flags, false);
}

extern "C" int hd_pdfill_finish(const void* z, int64_t z_pitch, void* w, int64_t w_pitch, int64_t ny, int64_t nx,
                                void* stream)
{
    if (!z || !w) return HD_ERR_NULL;
    if (ny < 1 || nx < 1 || z_pitch < nx || w_pitch < nx) return HD_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    hd_prof_begin("fill_finish_kernel", s);
    fill_finish_kernel<<<stream_grid(ny * nx), 256, 0, s>>>((const float*)z, z_pitch, (float*)w, w_pitch, ny, nx);
    HD_LAUNCH_CHECK(); hd_count_launch();
    return HD_OK;
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
