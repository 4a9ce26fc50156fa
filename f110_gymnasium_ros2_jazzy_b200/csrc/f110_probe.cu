// f110_probe.cu -- the empirical gather roofline bench.py quotes next to the lidar kernel:
// every thread performs `chain` DEPENDENT 8-byte loads at pseudo-random cells of the handle's distance-transform
// map (the next address depends on the value just loaded), with the lidar kernel's launch shape and no other work.
// It answers "how fast can this chip chase `chain` fp64 cells per thread through a map of this size".
#include <cuda_runtime.h>
#include <stdint.h>

#include "f110_b200.h"

namespace {
__global__ void __launch_bounds__(128) gather_probe_kernel(const double* __restrict__ map, unsigned mask, unsigned offset,
                                                           int chain, unsigned total, double* __restrict__ sink) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    unsigned idx = (t * 2654435761u) & mask;
    double acc = 0.0;
    for (int k = 0; k < chain; ++k) {
        const double v = __ldg(map + offset + idx);
        acc += v;
        idx = (idx * 1664525u + 1013904223u + (unsigned)__double_as_longlong(v)) & mask;
    }
    if (acc == -1.0) sink[0] = acc;   // never true: keeps the chain alive
}
}  // namespace

extern "C" int f110_gather_probe(const double* map_dev, int64_t num_cells, int64_t window_cells, int32_t chain,
                                 int64_t num_threads, int32_t repeats, double* sink_dev, float* ms_out, void* stream) {
    if (!map_dev || !sink_dev || !ms_out || num_cells < 2 || chain < 1 || num_threads < 1 || repeats < 1) return F110_ERR_INVALID;
    int64_t w = window_cells > 0 && window_cells < num_cells ? window_cells : num_cells;
    unsigned mask = 1;
    while ((int64_t)(mask << 1) <= w) mask <<= 1;
    mask -= 1;
    const unsigned offset = (unsigned)((num_cells - (int64_t)mask - 1) / 2);
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return F110_ERR_CUDA;
    const unsigned total = (unsigned)num_threads;
    const unsigned blocks = (total + 127) / 128;
    gather_probe_kernel<<<blocks, 128, 0, s>>>(map_dev, mask, offset, chain, total, sink_dev);   // warm-up
    cudaEventRecord(e0, s);
    for (int r = 0; r < repeats; ++r) gather_probe_kernel<<<blocks, 128, 0, s>>>(map_dev, mask, offset, chain, total, sink_dev);
    cudaEventRecord(e1, s);
    if (cudaEventSynchronize(e1) != cudaSuccess) return F110_ERR_CUDA;
    cudaEventElapsedTime(ms_out, e0, e1);
    *ms_out /= (float)repeats;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return F110_OK;
}
