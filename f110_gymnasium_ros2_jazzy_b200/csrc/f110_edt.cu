// f110_edt.cu -- exact Euclidean distance transform on the device.
//
// The reference builds its map with scipy.ndimage.distance_transform_edt(img) * resolution on the host
// (laser_models.py:40-53,383-427; ~0.7 s for a 2000 x 2000 map).  scipy returns sqrt(d2) of the exact integer squared
// distance d2 to the nearest obstacle pixel, so dt = resolution * sqrt((double)d2) with IEEE sqrt and multiply is
// bit-identical to the reference's array.  d2 is computed exactly, in integers, with the two-phase algorithm of
// Meijster, Roerdink & Hesselink (2000): (1) per column, the vertical distance g to the nearest obstacle; (2) per row,
// the lower envelope of the parabolas (x - i)^2 + g(i)^2.
#include <cuda_runtime.h>
#include <stdint.h>

#include "f110_kernels.cuh"

namespace {

// phase 1: one thread per column (adjacent threads read adjacent bytes: coalesced)
__global__ void edt_columns_kernel(const uint8_t* __restrict__ freemask, int H, int W, int* __restrict__ g, int inf) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= W) return;
    int d = inf;
    for (int r = 0; r < H; ++r) {
        d = freemask[(size_t)r * W + c] ? (d < inf ? d + 1 : inf) : 0;
        g[(size_t)r * W + c] = d;
    }
    d = inf;
    for (int r = H - 1; r >= 0; --r) {
        d = freemask[(size_t)r * W + c] ? (d < inf ? d + 1 : inf) : 0;
        const int up = g[(size_t)r * W + c];
        g[(size_t)r * W + c] = d < up ? d : up;
    }
}

__device__ __forceinline__ long long parab(long long x, long long i, long long gi) { return (x - i) * (x - i) + gi * gi; }

// phase 2: one thread per row; s/t are the row's envelope stacks (global scratch)
__global__ void edt_rows_kernel(const int* __restrict__ g, int H, int W, int* __restrict__ s_all, int* __restrict__ t_all,
                                double resolution, double* __restrict__ dt) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= H) return;
    const int* gr = g + (size_t)r * W;
    int* s = s_all + (size_t)r * W;
    int* t = t_all + (size_t)r * W;
    int q = 0;
    s[0] = 0; t[0] = 0;
    for (int u = 1; u < W; ++u) {
        const long long gu = gr[u];
        while (q >= 0 && parab(t[q], s[q], gr[s[q]]) > parab(t[q], u, gu)) --q;
        if (q < 0) { q = 0; s[0] = u; }
        else {
            // Sep(i, u) = (u^2 - i^2 + g(u)^2 - g(i)^2) div (2 (u - i)), all non-negative here
            const long long i = s[q], gi = gr[i];
            const long long w = 1 + ((long long)u * u - i * i + gu * gu - gi * gi) / (2 * ((long long)u - i));
            if (w < W) { ++q; s[q] = u; t[q] = (int)w; }
        }
    }
    for (int u = W - 1; u >= 0; --u) {
        const long long d2 = parab(u, s[q], gr[s[q]]);
        dt[(size_t)r * W + u] = resolution * sqrt((double)d2);    // -fmad=false: IEEE sqrt then IEEE multiply
        if (u == t[q]) --q;
    }
}

}  // namespace

// freemask: DEVICE [H][W], non-zero = free space (pixel > 128 after the reference's bottom-up flip); dt: DEVICE fp64 [H][W]
int edt_device(const uint8_t* freemask, int H, int W, double resolution, double* dt, cudaStream_t stream) {
    int *g = nullptr, *s = nullptr, *t = nullptr;
    const size_t cells = (size_t)H * W;
    if (cudaMalloc(&g, 3 * cells * sizeof(int)) != cudaSuccess) return -1;
    s = g + cells; t = s + cells;
    const int inf = H + W + 1;     // larger than any real distance; (2 inf)^2 fits easily in 64 bits
    edt_columns_kernel<<<(W + 127) / 128, 128, 0, stream>>>(freemask, H, W, g, inf);
    edt_rows_kernel<<<(H + 63) / 64, 64, 0, stream>>>(g, H, W, s, t, resolution, dt);
    const cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(g);
    return e == cudaSuccess && cudaGetLastError() == cudaSuccess ? 0 : -1;
}
