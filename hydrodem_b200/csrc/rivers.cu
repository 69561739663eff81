// RouteRivers.apply (filters/custom_filters.py:165-199) -- the one ORDER-DEPENDENT filter of the HydroSHEDS branch.
//
// Reference: a raster scan over the interior cells of the rivers mask with int(v) == 1 (a float32 snapshot of the mask,
// sliding_window.py:132, :192).  Each visit looks at the 3x3 window of a float32 working copy G of the reference DEM,
// takes its minimum (np.amin: a NaN in the window makes the minimum NaN and nothing matches), marks every window cell
// equal to the minimum as river and overwrites it with 10000 in G -- so later visits see the cells consumed by earlier
// ones.  A visit at (j, i) touches rows j-1..j+1 and columns i-1..i+1: it conflicts with every visit within Chebyshev
// distance 2, and the reference's result is the one of the raster order.
//
// Order-preserving wavefront: ONE WARP PER ROW, pipelined.  Row j may work on the columns of a chunk once row j-1 has
// finished everything up to two columns past it (per-row progress counters in global memory, release / acquire by
// __threadfence + volatile accesses).  Inside a chunk the warp holds the 3 x 32 window cells in registers (one column
// per lane), runs the chunk's visits in column order with shuffles -- no memory access per visit -- and writes back only
// the cells it changed.  Rows behind cannot have touched the chunk yet, the row ahead is done with it: the register copy
// is exact.  Every warp of a launch is co-resident (rows are processed in stripes of that many), and a warp only waits
// for the warp of the row above, so the waits cannot deadlock.
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int CHUNK = 30;            // columns visited per step: lanes hold columns c0-1 .. c0+30
constexpr int RNT = 256;

__global__ void __launch_bounds__(RNT) route_rivers_kernel(const float* __restrict__ mask, int64_t mask_pitch,
                                                           float* __restrict__ g, int64_t g_pitch, uint8_t* __restrict__ out,
                                                           int64_t out_pitch, int ny, int nx, int row0, int nrows,
                                                           int* __restrict__ progress)
{
    const int warp = (int)((blockIdx.x * (unsigned)RNT + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    const int j = row0 + warp;                         // interior row 1 .. ny-2
    volatile int* above = progress + (j - 1);
    const float BIG = 10000.0f;
    for (int c0 = 1; c0 < nx - 1; c0 += CHUNK) {
        const int c_end = min(c0 + CHUNK, nx - 1);                       // visits at columns [c0, c_end)
        // the row above must be done with every visit at a column <= c_end + 1 (progress = all columns below it are done)
        if (lane == 0) {
            const int need = min(c_end + 2, nx - 1);
            while (*above < need) __nanosleep(20);
        }
        __syncwarp();
        __threadfence();
        const int x = c0 - 1 + lane;                                     // this lane's column
        const bool in = x < nx;
        float v[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = in ? __ldcg(g + (int64_t)(j - 1 + d) * g_pitch + x) : BIG;
        const float mk = (lane >= 1 && x < c_end && in) ? __ldg(mask + (int64_t)j * mask_pitch + x) : 0.f;
        unsigned todo = __ballot_sync(0xffffffffu, mk == mk && (int)mk == 1);      // iter_over_ones: int(v) == 1
        unsigned changed = 0u;                                           // bit d: v[d] was consumed in this chunk
        while (todo) {
            const int L = __ffs(todo) - 1;                               // lane of the visited column (1 .. 30)
            todo &= todo - 1;
            const bool nan = v[0] != v[0] || v[1] != v[1] || v[2] != v[2];
            const float cm = fminf(v[0], fminf(v[1], v[2]));
            const bool nan3 = __shfl_sync(0xffffffffu, nan, L - 1) | __shfl_sync(0xffffffffu, nan, L) |
                              __shfl_sync(0xffffffffu, nan, L + 1);
            const float m = fminf(__shfl_sync(0xffffffffu, cm, L - 1),
                                  fminf(__shfl_sync(0xffffffffu, cm, L), __shfl_sync(0xffffffffu, cm, L + 1)));
            if (!nan3 && lane >= L - 1 && lane <= L + 1) {
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    if (v[d] == m) { v[d] = BIG; changed |= 1u << d; }
            }
        }
        if (in) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
                if (changed & (1u << d)) {
                    g[(int64_t)(j - 1 + d) * g_pitch + x] = BIG;
                    out[(int64_t)(j - 1 + d) * out_pitch + x] = 1;
                }
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) *(volatile int*)(progress + j) = c_end;           // every visit of this row at a column < c_end is done
    }
    __threadfence();
    if (lane == 0) *(volatile int*)(progress + j) = nx;
}

__global__ void progress_init_kernel(int* progress, int ny, int nx)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ny; j += gridDim.x * blockDim.x)
        progress[j] = (j == 0 || j >= ny - 1) ? nx : 0;
}

}  // namespace

// mask: F32 rivers mask (visits where int(v) == 1).  g: F32 working copy of the reference DEM (dem.astype('float32'),
// sliding_window.py:132) -- MODIFIED in place like dem_sliding.grid (:198).  out: U8, zero-initialised here, 1 = routed river
// (the reference returns float64 zeros / ones, :187).  workspace: >= ny ints (per-row progress counters).
extern "C" int hd_route_rivers(const void* mask, int64_t mask_pitch, void* g, int64_t g_pitch, void* out, int64_t out_pitch,
                               int64_t ny, int64_t nx, int ws, void* workspace, int64_t workspace_bytes, void* stream)
{
    if (!mask || !g || !out || !workspace) return HD_ERR_NULL;
    if (ws > ny || ws > nx) return HD_ERR_WINDOW_HIGH;
    if (ws % 2 != 1) return HD_ERR_WINDOW_EVEN;
    if (ws != 3) return HD_ERR_UNSUPPORTED;                     // ProcessRivers uses window_size = 3 (custom_filters.py:796)
    if (ny > 0x7fffffff || nx > 0x7fffffff || mask_pitch < nx || g_pitch < nx || out_pitch < nx) return HD_ERR_ARG;
    if (workspace_bytes < ny * (int64_t)sizeof(int)) return HD_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    HD_CUDA_OK(cudaMemset2DAsync(out, (size_t)out_pitch, 0, (size_t)nx, (size_t)ny, s));
    int* progress = (int*)workspace;
    progress_init_kernel<<<64, 256, 0, s>>>(progress, (int)ny, (int)nx);
    HD_LAUNCH_CHECK();
    int per_sm = 1;
    HD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, route_rivers_kernel, RNT, 0));
    int capacity = hd_num_sms_total() * (per_sm < 1 ? 1 : per_sm) * (RNT / 32);           // co-resident warps = rows per launch
    if (const char* e = getenv("HD_RIVERS_ROWS_PER_LAUNCH")) {                            // tests: several stripes on a small raster
        const int v = atoi(e);
        if (v >= 8 && v < capacity) capacity = v / 8 * 8;
    }
    for (int64_t row0 = 1; row0 < ny - 1; row0 += capacity) {
        const int nrows = (int)((ny - 1 - row0) < capacity ? (ny - 1 - row0) : capacity);
        hd_prof_begin("route_rivers_kernel", s);
        route_rivers_kernel<<<hd_cdiv(nrows, RNT / 32), RNT, 0, s>>>((const float*)mask, mask_pitch, (float*)g, g_pitch,
                                                                     (uint8_t*)out, out_pitch, (int)ny, (int)nx, (int)row0, nrows,
                                                                     progress);
        HD_LAUNCH_CHECK(); hd_count_launch();
    }
    return HD_OK;
}
