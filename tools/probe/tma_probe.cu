// Standalone probe of the TMA tile-load path (debugging aid, not part of the product).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hydrodem_b200/csrc/common.cuh"

#include "../../hydrodem_b200/csrc/runtime.cu"

__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, float* out, int box_w, int box_h, int x0, int y0,
                             int do_prefetch)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        if (do_prefetch) tma_prefetch_desc(&tm);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, box_w * box_h * 4);
        tma_load_2d(smem, &tm, x0, y0, &bar);
    }
    mbar_wait(&bar, 0);
    const float* t = (const float*)smem;
    for (int i = threadIdx.x; i < box_w * box_h; i += blockDim.x) out[i] = t[i];
}

int main(int argc, char** argv)
{
    const int box_w = argc > 1 ? atoi(argv[1]) : 132, box_h = argc > 2 ? atoi(argv[2]) : 34;
    const int prefetch = argc > 3 ? atoi(argv[3]) : 1;
    const int x0 = argc > 4 ? atoi(argv[4]) : -1, y0 = argc > 5 ? atoi(argv[5]) : -1;
    const int ny = 100, nx = 300, pitch = 320;
    std::vector<float> h(ny * pitch);
    for (int i = 0; i < ny * pitch; ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, box_w * box_h * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    int e = hd_make_tmap_2d(&tm, d, HD_F32, ny, nx, pitch, box_w, box_h, false);
    printf("box %dx%d prefetch %d origin (%d,%d): encode status %d\n", box_w, box_h, prefetch, x0, y0, e);
    if (e) return 1;
    probe_kernel<<<1, 128, box_w * box_h * 4>>>(tm, o, box_w, box_h, x0, y0, prefetch);
    cudaError_t err = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(err));
    if (err == cudaSuccess) {
        std::vector<float> r(box_w * box_h);
        cudaMemcpy(r.data(), o, r.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < box_h; ++y)
            for (int x = 0; x < box_w; ++x) {
                int gy = y0 + y, gx = x0 + x;
                float want = (gy >= 0 && gy < ny && gx >= 0 && gx < nx) ? (float)(gy * pitch + gx) : 0.f;
                if (r[y * box_w + x] != want) ++bad;
            }
        printf("  mismatches: %d\n", bad);
    }
    return 0;
}
