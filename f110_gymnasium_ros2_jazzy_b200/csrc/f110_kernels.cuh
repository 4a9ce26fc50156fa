// f110_kernels.cuh -- device-side data model shared by the step kernels and the C ABI.
//
// Everything on the step path is fp64 evaluated in the reference's operator order with FMA
// contraction disabled (-fmad=false): numba's code for the reference contains no fused
// multiply-adds, and the ray-march's cell lookups are discontinuous in the last bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "f110_b200.h"

#define F110_PI 3.141592653589793  // numpy.pi

// Distance-transform map (laser_models.py:383-427).  Row 0 is the BOTTOM image row.
struct MapView {
    const double* dt;   // [H][W] metres to the nearest obstacle
    int H, W;
    int last;           // (H-1)*W + (W-1): the cell numba's negative-index wrap lands on
    double res;         // metres per cell
    double inv_fx;      // 2^fx_bits / res: quotient in 2^-fx_bits cell units for the guarded fast cell index
    unsigned w_fx, h_fx;  // W << fx_bits, H << fx_bits
    unsigned fx_bits;   // fraction bits: 32 - bit_length(max(W, H)), at most 24 (W << fx_bits must fit 32 bits)
    unsigned fx_mask;   // (1 << fx_bits) - 1
    double ox, oy, oc, os;
    double wres, hres;  // W*res, H*res
};

// Persistent simulation state: one arena, SoA over s = env*A + agent (coalesced for one thread
// per vehicle) and over env.  Everything a checkpoint needs lives between arena and arena+nbytes.
struct SimState {
    double* x[7];       // [NA] x, y, steer, v, yaw, yaw_rate, slip        base_classes.py:97-98
    double* steer_buf0; // [NA] newest queued steering command             base_classes.py:270-278
    double* steer_buf1; // [NA] oldest queued steering command
    double* start_x;    // [NA] f110_env.py:448-450
    double* start_y;
    double* start_th;
    double* lap_times;  // [NA] f110_env.py:346-348
    double* lap_counts; // [NA]
    double* time;       // [N]  f110_env.py:406
    double* rot_c;      // [N]  cos(-start_theta[ego])                      f110_env.py:451
    double* rot_s;      // [N]  sin(-start_theta[ego])
    int32_t* steer_cnt; // [NA] entries queued in the delay FIFO (0..2)
    int32_t* toggles;   // [NA] f110_env.py:339-346
    uint32_t* step_count; // [N] steps since the env's reset (Philox counter)
    uint8_t* near_start;  // [NA]
    uint8_t* collisions;  // [NA] last step's GJK | iTTC
};

// Scratch carried between the three kernels of one step (not part of a checkpoint).
struct StepScratch {
    double* scan_x;     // [NA] lidar pose for this step (base_classes.py:420-422)
    double* scan_y;
    double* pre_yaw;    // [NA] yaw after dynamics, BEFORE iTTC zeroing (Simulator.agent_poses, :587)
    double* theta0;     // [NA] wrapped theta index of beam 0 (laser_models.py:167-172)
    int32_t* ttc_hit;   // [NA] set by the lidar kernel
    double* scan;       // [NA][B] noisy map scan, before the opponent ray-cast
    unsigned long long* lookups;  // [3] dt lookups, rays, longest ray (only with F110_FLAG_COUNT_LOOKUPS)
    double* stats;      // [F110_NUM_STATS]
    // launch-order history of the lidar kernel (see lidar_kernel): [0] = this step's order, [1] = being recorded
    unsigned num_units;        // ceil(NA*B / 32) warp-sized work units, padded to a multiple of 4
    unsigned front_units;      // capacity of the heavy-first front region (multiple of 4)
    unsigned* heavy_cnt;       // [2]
    unsigned* heavy_list;      // [2][front_units]
    uint8_t* unit_heavy;       // [2][num_units]
};

struct FastDiv { uint32_t mul, sh1, sh2; };   // n / d == (t + ((n - t) >> sh1)) >> sh2 with t = umulhi(n, mul)

struct SimConst {
    int N, A, B, NA;
    int theta_dis, integrator, ego;
    double fov, eps, max_range, timestep, lidar_dist, ttc_thresh, noise_std;
    double theta_inc;   // theta_dis * (fov/(B-1)) / (2 pi)               laser_models.py:367-368
    float lidar_max;
    uint64_t seed;
    uint32_t noise_key;  // Philox key derived from seed
    FastDiv div_B, div_A;
    const double* params;      // [A][18]
    const double* sim_params;  // [18] Simulator.params (construction time; base_classes.py:562)
    const double* sines;       // [theta_dis]
    const double* cosines;     // [theta_dis]
    const double* scan_angles; // [B]
    const double* beam_cos;    // [B]
    const double* side_dist;   // [B]
};

void launch_dynamics(const SimConst& c, const SimState& st, const StepScratch& sc, const F110StepIO& io, cudaStream_t s);
void launch_lidar(const SimConst& c, const MapView& m, const SimState& st, const StepScratch& sc, const F110StepIO& io,
                  bool count_lookups, int threads_per_block, cudaStream_t s);
void launch_post(const SimConst& c, const SimState& st, const StepScratch& sc, const F110StepIO& io, cudaStream_t s);
void launch_sim_reset(const SimConst& c, const SimState& st, const double* poses, const uint8_t* mask, cudaStream_t s);

// exact EDT on the device (f110_edt.cu): freemask DEVICE [H][W] (non-zero = free), dt DEVICE fp64 [H][W]; synchronises
int edt_device(const uint8_t* freemask, int H, int W, double resolution, double* dt, cudaStream_t stream);
int f110_set_error(int code, const char* msg);   // sets the thread-local message of f110_last_error(); returns code
