// f110_consumers.cu -- device-side forms of the consumers that sit directly around the env in the reference's
// training loop, so a rollout never leaves the GPU.  Paths are relative to /root/reference/.
//
//   gap_follow_kernel : rl_training/utils/gap_follow.py:3-58, the rule-based opponent train_ddpg.py:168 drives
//                       from info["scans"][1].  One CTA per scan; float32 arithmetic in numpy's order, so the
//                       chosen beam index -- and with it (steer, speed) -- is identical to the reference's.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "f110_b200.h"

namespace {

constexpr int GF_THREADS = 128;

struct RunSummary { int all_true, prefix, suffix, best_len, best_start; };

__global__ void __launch_bounds__(GF_THREADS) gap_follow_kernel(const float* __restrict__ scans, long long scan_stride, int n,
                                                                float* __restrict__ actions, long long action_stride,
                                                                double angle_min, double angle_increment, float max_distance,
                                                                int window_size, int bubble_radius, float threshold) {
    extern __shared__ float sm[];
    float* clipped = sm;           // [n]
    float* proc = sm + n;          // [n]
    __shared__ float s_wv[GF_THREADS / 32];
    __shared__ int s_wi[GF_THREADS / 32];
    __shared__ int s_closest;
    __shared__ RunSummary s_run[GF_THREADS];
    const int tid = threadIdx.x;
    const float* scan = scans + (size_t)blockIdx.x * scan_stride;

    // preprocess_lidar :3-12: mean of the clipped ranges over [i - w/2, i + w/2] cut at the ends
    for (int i = tid; i < n; i += GF_THREADS) {
        float v = scan[i];
        v = v < 0.f ? 0.f : v;
        v = v > max_distance ? max_distance : v;
        clipped[i] = v;
    }
    __syncthreads();
    const int half = window_size / 2;
    float best_v = INFINITY;
    int best_i = 0x7fffffff;
    for (int i = tid; i < n; i += GF_THREADS) {
        const int s = max(0, i - half), e = min(n - 1, i + half);
        float acc = 0.f;
        for (int k = s; k <= e; ++k) acc += clipped[k];          // numpy: sequential float32 sum for < 8 elements
        const float m = __fdiv_rn(acc, (float)(e - s + 1));
        proc[i] = m;
        if (m < best_v) { best_v = m; best_i = i; }              // i increases: the first minimum of this thread
    }
    // create_bubble :14-19: np.argmin = first minimum over the whole scan
    for (int o = 16; o > 0; o >>= 1) {
        const float v2 = __shfl_down_sync(0xffffffffu, best_v, o);
        const int i2 = __shfl_down_sync(0xffffffffu, best_i, o);
        if (v2 < best_v || (v2 == best_v && i2 < best_i)) { best_v = v2; best_i = i2; }
    }
    if ((tid & 31) == 0) { s_wv[tid >> 5] = best_v; s_wi[tid >> 5] = best_i; }
    __syncthreads();
    if (tid == 0) {
        float v = s_wv[0];
        int idx = s_wi[0];
        for (int w = 1; w < GF_THREADS / 32; ++w)
            if (s_wv[w] < v || (s_wv[w] == v && s_wi[w] < idx)) { v = s_wv[w]; idx = s_wi[w]; }
        s_closest = idx == 0x7fffffff ? 0 : idx;
    }
    __syncthreads();
    {
        const int s = max(s_closest - bubble_radius, 0), e = min(s_closest + bubble_radius, n - 1);
        for (int i = s + tid; i <= e; i += GF_THREADS) proc[i] = 0.f;
    }
    __syncthreads();

    // find_max_gap :21-38: the FIRST longest run of proc > threshold.  Each thread summarises a contiguous chunk,
    // thread 0 stitches the chunks in order.
    const int chunk = (n + GF_THREADS - 1) / GF_THREADS;
    {
        const int c0 = min(tid * chunk, n), c1 = min(c0 + chunk, n);
        RunSummary r;
        r.all_true = 1; r.prefix = 0; r.suffix = 0; r.best_len = -1; r.best_start = 0;
        int start = -1, seen_false = 0;
        for (int i = c0; i < c1; ++i) {
            const bool val = proc[i] > threshold;
            if (val) {
                if (start < 0) start = i;
            } else {
                if (start >= 0) {
                    if (!seen_false && start == c0) r.prefix = i - c0;                 // run touching the chunk start
                    else if (i - 1 - start > r.best_len) { r.best_len = i - 1 - start; r.best_start = start; }
                    start = -1;
                }
                seen_false = 1;
                r.all_true = 0;
            }
        }
        if (start >= 0) {
            if (r.all_true) r.prefix = c1 - c0;
            r.suffix = c1 - start;
        }
        if (c0 == c1) { r.all_true = 1; r.prefix = 0; r.suffix = 0; }
        s_run[tid] = r;
    }
    __syncthreads();
    if (tid == 0) {
        int best_s = 0, best_e = n - 1, best_len = -1, open = -1;
        for (int t = 0; t < GF_THREADS; ++t) {
            const int c0 = min(t * chunk, n), c1 = min(c0 + chunk, n);
            if (c0 == c1) break;
            const RunSummary r = s_run[t];
            if (r.all_true) { if (open < 0) open = c0; continue; }
            // a run entering from the left (or starting at c0) ends inside this chunk
            if (open >= 0 || r.prefix > 0) {
                const int st = open >= 0 ? open : c0;
                const int en = c0 + r.prefix - 1;
                if (en >= st && en - st > best_len) { best_len = en - st; best_s = st; best_e = en; }
                open = -1;
            }
            if (r.best_len > best_len) { best_len = r.best_len; best_s = r.best_start; best_e = r.best_start + r.best_len; }
            if (r.suffix > 0) open = c1 - r.suffix;
        }
        if (open >= 0 && n - 1 - open > best_len) { best_len = n - 1 - open; best_s = open; best_e = n - 1; }
        const int best = (best_s + best_e) / 2;                                   // find_best_point :40-41
        const double steering = angle_min + best * angle_increment;               // :49
        const double d10 = 10 * (3.141592653589793 / 180.0), d20 = 20 * (3.141592653589793 / 180.0);
        const double speed = fabs(steering) < d10 ? 2.5 : (fabs(steering) < d20 ? 2.0 : 1.5);
        float* out = actions + (size_t)blockIdx.x * action_stride;
        out[0] = (float)steering;
        out[1] = (float)speed;
    }
}

}  // namespace

extern "C" int f110_gap_follow(const float* scans, int64_t num_scans, int64_t scan_stride, int32_t num_beams,
                               float* actions, int64_t action_stride, double angle_min, double angle_increment,
                               float max_distance, int32_t window_size, int32_t bubble_radius, float threshold, void* stream) {
    if (!scans || !actions || num_scans < 0 || num_beams < 1 || num_beams > 8192 || window_size < 1 || window_size > 15)
        return F110_ERR_INVALID;
    if (num_scans == 0) return F110_OK;
    gap_follow_kernel<<<(unsigned)num_scans, GF_THREADS, 2 * sizeof(float) * num_beams, (cudaStream_t)stream>>>(
        scans, scan_stride, num_beams, actions, action_stride, angle_min, angle_increment, max_distance, window_size,
        bubble_radius, threshold);
    return cudaPeekAtLastError() == cudaSuccess ? F110_OK : F110_ERR_CUDA;
}
