// f110_kernels.cu -- the three sm_100a kernels of one batched F110Env step.
//
//   K1 dynamics_kernel : one thread per vehicle.  Steering-delay FIFO, bang-bang/P controller,
//                        fp64 single-track RK4 (or Euler), post clamps, lidar pose.
//                        (RaceCar.update_pose, base_classes.py:256-422)
//   K2 lidar_kernel    : one thread per beam, flat over all N*A*B rays.  Ray-march over the
//                        distance transform, + noise, per-beam iTTC test.
//                        (get_scan/trace_ray laser_models.py:106-186, scan :429-454,
//                         check_ttc_jit :188-217)
//   K3 post_kernel     : one CTA per env.  iTTC state zeroing, GJK over all pairs, opponent
//                        ray-cast, finish-zone/lap bookkeeping, done, observation packing.
//                        (base_classes.py:229-254,206-227,549-563,592-602; f110_env.py:310-352,552-602)
//
// Reference paths are relative to f110_gymnasium/gym/f110_gym/envs/.
// Compile with -fmad=false: the operator order below is the reference's, without contraction.
#include "f110_kernels.cuh"

#include <math.h>

namespace {

enum { P_MU = F110_P_MU, P_CSF = F110_P_C_SF, P_CSR = F110_P_C_SR, P_LF = F110_P_LF, P_LR = F110_P_LR,
       P_H = F110_P_H, P_M = F110_P_M, P_I = F110_P_I, P_SMIN = F110_P_S_MIN, P_SMAX = F110_P_S_MAX,
       P_SVMIN = F110_P_SV_MIN, P_SVMAX = F110_P_SV_MAX, P_VSWITCH = F110_P_V_SWITCH, P_AMAX = F110_P_A_MAX,
       P_VMIN = F110_P_V_MIN, P_VMAX = F110_P_V_MAX, P_WIDTH = F110_P_WIDTH, P_LENGTH = F110_P_LENGTH };

// ---------------------------------------------------------------- numpy scalar semantics

// np.clip(x, lo, hi) = min(max(x, lo), hi), NaN propagating
__device__ __forceinline__ double clipd(double x, double lo, double hi) {
    x = (x < lo) ? lo : x;
    x = (x > hi) ? hi : x;
    return x;
}

// Python float `a % b` for b > 0 (floored modulo)
__device__ __forceinline__ double floored_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) {
        if (m < 0.0) m += b;
    } else {
        m = 0.0;
    }
    return m;
}

__device__ __forceinline__ double wrap_angle(double a) {  // base_classes.py:408, f110_env.py:548-550
    return floored_mod(a + F110_PI, 2 * F110_PI) - F110_PI;
}

// ---------------------------------------------------------------- vehicle model

struct VehParams {
    double mu, C_Sf, C_Sr, lf, lr, h, m, I, s_min, s_max, sv_min, sv_max, v_switch, a_max, v_min, v_max;
};

__device__ __forceinline__ VehParams load_params(const double* __restrict__ p) {
    VehParams v;
    v.mu = __ldg(p + P_MU); v.C_Sf = __ldg(p + P_CSF); v.C_Sr = __ldg(p + P_CSR); v.lf = __ldg(p + P_LF);
    v.lr = __ldg(p + P_LR); v.h = __ldg(p + P_H); v.m = __ldg(p + P_M); v.I = __ldg(p + P_I);
    v.s_min = __ldg(p + P_SMIN); v.s_max = __ldg(p + P_SMAX); v.sv_min = __ldg(p + P_SVMIN);
    v.sv_max = __ldg(p + P_SVMAX); v.v_switch = __ldg(p + P_VSWITCH); v.a_max = __ldg(p + P_AMAX);
    v.v_min = __ldg(p + P_VMIN); v.v_max = __ldg(p + P_VMAX);
    return v;
}

// accl_constraints, dynamic_models.py:29-60
__device__ __forceinline__ double accl_constraints(double vel, double accl, const VehParams& p) {
    double pos_limit = (vel > p.v_switch) ? p.a_max * p.v_switch / vel : p.a_max;
    if ((vel <= p.v_min && accl <= 0) || (vel >= p.v_max && accl >= 0)) accl = 0.;
    else if (accl <= -p.a_max) accl = -p.a_max;
    else if (accl >= pos_limit) accl = pos_limit;
    return accl;
}

// steering_constraint, dynamic_models.py:62-87
__device__ __forceinline__ double steering_constraint(double sa, double sv, const VehParams& p) {
    if ((sa <= p.s_min && sv <= 0) || (sa >= p.s_max && sv >= 0)) sv = 0.;
    else if (sv <= p.sv_min) sv = p.sv_min;
    else if (sv >= p.sv_max) sv = p.sv_max;
    return sv;
}

// vehicle_dynamics_st (with the embedded vehicle_dynamics_ks branch), dynamic_models.py:90-176
__device__ __forceinline__ void vehicle_dynamics_st(const double (&x)[7], double u_sv, double u_accl,
                                                    const VehParams& p, double (&f)[7]) {
    const double g = 9.81;
    const double u0 = steering_constraint(x[2], u_sv, p);
    const double u1 = accl_constraints(x[3], u_accl, p);
    if (fabs(x[3]) < 0.5) {
        // kinematic branch :152-160; vehicle_dynamics_ks re-applies the (idempotent) constraints :112
        const double lwb = p.lf + p.lr;
        const double k0 = steering_constraint(x[2], u0, p);
        const double k1 = accl_constraints(x[3], u1, p);
        double s4, c4;
        sincos(x[4], &s4, &c4);
        const double t2 = tan(x[2]);
        const double c2 = cos(x[2]);
        f[0] = x[3] * c4;
        f[1] = x[3] * s4;
        f[2] = k0;
        f[3] = k1;
        f[4] = x[3] / lwb * t2;
        f[5] = u1 / lwb * t2 + x[3] / (lwb * (c2 * c2)) * u0;
        f[6] = 0.0;
    } else {
        double sb, cb;
        sincos(x[6] + x[4], &sb, &cb);
        const double lrlf = p.lr + p.lf;
        const double glr = g * p.lr - u1 * p.h;   // (g*lr - u[1]*h)
        const double glf = g * p.lf + u1 * p.h;   // (g*lf + u[1]*h)
        f[0] = x[3] * cb;
        f[1] = x[3] * sb;
        f[2] = u0;
        f[3] = u1;
        f[4] = x[5];
        f[5] = -p.mu * p.m / (x[3] * p.I * lrlf) * (p.lf * p.lf * p.C_Sf * glr + p.lr * p.lr * p.C_Sr * glf) * x[5]
             + p.mu * p.m / (p.I * lrlf) * (p.lr * p.C_Sr * glf - p.lf * p.C_Sf * glr) * x[6]
             + p.mu * p.m / (p.I * lrlf) * p.lf * p.C_Sf * glr * x[2];
        f[6] = (p.mu / (x[3] * x[3] * lrlf) * (p.C_Sr * glf * p.lr - p.C_Sf * glr * p.lf) - 1) * x[5]
             - p.mu / (x[3] * lrlf) * (p.C_Sr * glf + p.C_Sf * glr) * x[6]
             + p.mu / (x[3] * lrlf) * (p.C_Sf * glr) * x[2];
    }
}

// pid, dynamic_models.py:178-221
__device__ __forceinline__ void pid(double speed, double steer, double cur_speed, double cur_steer,
                                    const VehParams& p, double& accl, double& sv) {
    const double steer_diff = steer - cur_steer;
    if (fabs(steer_diff) > 1e-4) sv = (steer_diff / fabs(steer_diff)) * p.sv_max;
    else sv = 0.0;
    const double vel_diff = speed - cur_speed;
    double kp;
    if (cur_speed > 0.) {
        kp = (vel_diff > 0) ? 10.0 * p.a_max / p.v_max : 10.0 * p.a_max / (-p.v_min);
    } else {
        kp = (vel_diff > 0) ? 2.0 * p.a_max / p.v_max : 2.0 * p.a_max / (-p.v_min);
    }
    accl = kp * vel_diff;
}

// ---------------------------------------------------------------- K1: dynamics

__global__ void __launch_bounds__(128) dynamics_kernel(SimConst c, SimState st, StepScratch sc, F110StepIO io) {
    cudaGridDependencySynchronize();   // PDL: the previous step's post kernel (or whatever precedes in the stream) is done
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    // Every global load this thread needs is issued before the first store or branch that depends on one: the kernel is a
    // single dependent chain per thread, and with a cold L2 each load left in program order behind a branch costs a DRAM
    // round trip of its own (the publish copy, the masks, the state, the action and the parameters were five in a row).
    //
    // (1) The launch-order history the lidar kernel recorded during the previous step (the "next" halves of the
    // buffers; its length was latched into heavy_cnt[0] by the post kernel) becomes this step's order.  A copy rather
    // than a pointer flip, so a captured CUDA graph stays valid step after step.
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned nw = sc.num_units / 4;                        // num_units is padded to a multiple of 4
    const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(sc.unit_heavy + sc.num_units);
    uint32_t* __restrict__ dst = reinterpret_cast<uint32_t*>(sc.unit_heavy);
    uint32_t v0[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const unsigned k = (unsigned)s + j * stride; v0[j] = k < nw ? src[k] : 0u; }
    const unsigned n_heavy = sc.heavy_cnt[0];
    const unsigned heavy0 = (unsigned)s < sc.front_units ? sc.heavy_list[sc.front_units + s] : 0u;   // used only if s < n_heavy

    // (2) this thread's vehicle
    const bool veh = s < c.NA;
    const int env = veh ? s / c.A : 0;
    const int a = s - env * c.A;
    uint8_t active = 1, rst_flag = 0;
    double xl[7] = {0., 0., 0., 0., 0., 0., 0.}, b0 = 0., b1 = 0., px = 0., py = 0., pth = 0.;
    int cnt = 0;
    double raw_steer = 0., speed = 0.;
    if (veh) {
        if (io.active_mask) active = io.active_mask[env];
        if (io.reset_mask) {
            rst_flag = io.reset_mask[env];
            const double* ps = io.reset_poses + (size_t)s * 3;
            px = ps[0]; py = ps[1]; pth = ps[2];
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) xl[k] = st.x[k][s];
        b0 = st.steer_buf0[s]; b1 = st.steer_buf1[s];
        cnt = st.steer_cnt[s];
        if (io.actions) {
            if (io.actions_f64) {
                const double2 v = reinterpret_cast<const double2*>(io.actions)[s];
                raw_steer = v.x; speed = v.y;
            } else {
                const float2 v = reinterpret_cast<const float2*>(io.actions)[s];
                raw_steer = (double)v.x; speed = (double)v.y;
            }
        }
    }
    const VehParams p = load_params(c.params + (veh ? a : 0) * F110_NUM_PARAMS);

    // (1, continued) the stores of the copy
#pragma unroll
    for (int j = 0; j < 8; ++j) { const unsigned k = (unsigned)s + j * stride; if (k < nw) dst[k] = v0[j]; }
    for (unsigned base = (unsigned)s + 8 * stride; base < nw; base += 8 * stride) {   // batches the grid does not cover at once
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const unsigned k = base + j * stride; v[j] = k < nw ? src[k] : 0u; }
#pragma unroll
        for (int j = 0; j < 8; ++j) { const unsigned k = base + j * stride; if (k < nw) dst[k] = v[j]; }
    }
    if ((unsigned)s < n_heavy) sc.heavy_list[s] = heavy0;
    for (unsigned k = (unsigned)s + stride; k < n_heavy; k += stride) sc.heavy_list[k] = sc.heavy_list[sc.front_units + k];

    if (!veh || !active) return;
    const bool rst = rst_flag != 0;

    double x[7];
    if (rst) {
        // F110Env.reset bookkeeping (f110_env.py:440-451) + RaceCar.reset (base_classes.py:183-204)
        x[0] = px; x[1] = py; x[2] = 0.; x[3] = 0.; x[4] = pth; x[5] = 0.; x[6] = 0.;
        b0 = b1 = 0.;
        cnt = 0;
        st.start_x[s] = px; st.start_y[s] = py; st.start_th[s] = pth;
        st.near_start[s] = 1;
        st.toggles[s] = 0;
        st.collisions[s] = 0;
        if (a == c.ego) {
            const double th = -pth;
            st.rot_c[env] = cos(th);
            st.rot_s[env] = sin(th);
        }
        if (a == 0) { st.time[env] = 0.0; st.step_count[env] = 0u; }
        raw_steer = 0.; speed = 0.;   // the reset's own step uses a zero action (f110_env.py:457-458)
    } else {
#pragma unroll
        for (int k = 0; k < 7; ++k) x[k] = xl[k];
        if (a == 0) st.step_count[env] += 1u;
    }

    // steering delay FIFO, base_classes.py:270-278
    double steer;
    if (cnt < 2) { steer = 0.; cnt += 1; }
    else steer = b1;
    b1 = b0; b0 = raw_steer;

    double accl, sv;
    pid(speed, steer, x[3], x[2], p, accl, sv);
    sv = clipd(sv, p.sv_min, p.sv_max);         // :283
    accl = clipd(accl, -p.a_max, p.a_max);      // :284

    const double dt = c.timestep;
    if (c.integrator == F110_INTEGRATOR_RK4) {  // :285-374
        // One copy of the right-hand side, run four times, instead of four inlined copies: the kernel's instruction
        // stream is fetched cold by every SM, and the stages cannot overlap anyway.  Bit-identical to the reference's
        // unrolled form: k/2 == k*0.5 and 2*k are exact, k*1.0 == k, and k1 + 2*k2 + 2*k3 + k4 is summed left to right.
        double sum[7], xs[7], k[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) { xs[i] = x[i]; sum[i] = 0.; }
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            vehicle_dynamics_st(xs, sv, accl, p, k);
            const double wgt = (stage == 1 || stage == 2) ? 2.0 : 1.0;    // weight in the final sum
            const double half = stage < 2 ? 0.5 : 1.0;                    // k/2 for the two midpoint stages, k for the last
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                sum[i] = stage == 0 ? k[i] : sum[i] + wgt * k[i];
                xs[i] = x[i] + dt * (k[i] * half);
            }
        }
        const double w = dt * (1.0 / 6.0);
#pragma unroll
        for (int i = 0; i < 7; ++i) x[i] = x[i] + w * sum[i];
    } else {                                    // Euler :376-396
        double f[7];
        vehicle_dynamics_st(x, sv, accl, p, f);
#pragma unroll
        for (int i = 0; i < 7; ++i) x[i] = x[i] + dt * f[i];
    }

    // post clamps :400-417
    x[2] = clipd(x[2], p.s_min, p.s_max);
    x[3] = clipd(x[3], p.v_min, p.v_max);
    x[4] = wrap_angle(x[4]);
    const double YAW_RATE_CAP = 10.0;
    if (x[5] != x[5]) x[5] = 0.0;
    else if (isinf(x[5])) x[5] = x[5] > 0 ? YAW_RATE_CAP : -YAW_RATE_CAP;
    x[5] = clipd(x[5], -YAW_RATE_CAP, YAW_RATE_CAP);
    const double SLIP_CAP = 60 * (F110_PI / 180.0);
    if (x[6] != x[6]) x[6] = 0.0;
    x[6] = clipd(x[6], -SLIP_CAP, SLIP_CAP);

#pragma unroll
    for (int k = 0; k < 7; ++k) st.x[k][s] = x[k];
    st.steer_buf0[s] = b0; st.steer_buf1[s] = b1;
    st.steer_cnt[s] = cnt;

    // lidar pose :420-422 and the wrapped index of beam 0 (laser_models.py:167-172)
    double sx = x[0], sy = x[1];
    if (c.lidar_dist != 0.0) {
        double sn, cs;
        sincos(x[4], &sn, &cs);
        sx = x[0] + c.lidar_dist * cs;
        sy = x[1] + c.lidar_dist * sn;
    }
    sc.scan_x[s] = sx; sc.scan_y[s] = sy; sc.pre_yaw[s] = x[4];
    double ti = (double)c.theta_dis * (x[4] - c.fov / 2.) / (2. * F110_PI);
    ti = fmod(ti, (double)c.theta_dis);
    while (ti < 0) ti += (double)c.theta_dis;
    sc.theta0[s] = ti;
    sc.ttc_hit[s] = 0;
}

// Simulator.reset alone (base_classes.py:627-643): poses only, no step, env bookkeeping untouched
__global__ void sim_reset_kernel(SimConst c, SimState st, const double* __restrict__ poses, const uint8_t* __restrict__ mask) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= c.NA) return;
    const int env = s / c.A;
    if (mask && !mask[env]) return;
    const double* ps = poses + (size_t)s * 3;
    st.x[0][s] = ps[0]; st.x[1][s] = ps[1]; st.x[2][s] = 0.; st.x[3][s] = 0.;
    st.x[4][s] = ps[2]; st.x[5][s] = 0.; st.x[6][s] = 0.;
    st.steer_buf0[s] = 0.; st.steer_buf1[s] = 0.; st.steer_cnt[s] = 0;
    if (s - env * c.A == 0) st.step_count[env] = 0u;
}

// ---------------------------------------------------------------- K2: lidar

// xy_2_rc + distance_transform, laser_models.py:55-104 -- the exact restatement (two IEEE divisions).
// Kept out of line and out of the hot loop: only the finishing loop of a ray whose lookup fell in the guard band calls it.
// (A float->int conversion of NaN does not give 0 on this hardware -- (int)NaN is negative -- hence the two-sided clamp.)
__device__ __noinline__ int cell_index_exact(double x_rot, double y_rot, double res, double wres, double hres, int W, int last) {
    int idx;
    if (x_rot < 0 || x_rot >= wres || y_rot < 0 || y_rot >= hres) {
        idx = last;   // (r, c) = (-1, -1): numba wraps the negative indices to dt[H-1][W-1]
    } else {
        const int col = (int)(x_rot / res);
        const int row = (int)(y_rot / res);
        idx = row * W + col;
        idx = idx < 0 ? 0 : (idx > last ? last : idx);   // memory safety for NaN poses / the 1-ulp wres edge
    }
    return idx;
}

// Cell index without the two fp64 divisions.  q = x_rot * (2^F / res) is the quotient in 2^-F cell units, F = m.fx_bits
// (fl(x*inv)*2^F == fl(x*(inv*2^F)): scaling by a power of two commutes with rounding).  Its error against the true
// quotient is < 2.3e-16 relative, i.e. < 1e-6 units since q < 2^32, and the reference's own rounded quotient is within
// half an ulp of the true one, so whenever the F fractional bits are at least one unit away from both cell edges,
// trunc(q) >> F is exactly the reference's int(x_rot/res) and (q < W << F) is exactly its in-map test.  Everything
// else -- the 2/2^F of lookups next to a cell edge, the map border, negative (converts to 0), huge (0xFFFFFFFF) and NaN
// (0 or 0x80000000: fraction 0 either way) coordinates -- is not decided here: fast_cell returns false and the caller takes
// the exact path.
template <bool IDENT>
__device__ __forceinline__ void map_frame(const MapView& m, double x, double y, double& x_rot, double& y_rot) {
    const double x_trans = x - m.ox;
    const double y_trans = y - m.oy;
    if (IDENT) {          // orig_c == 1, orig_s == 0: x*1 + y*0 == x and -x*0 + y*1 == y exactly
        x_rot = x_trans; y_rot = y_trans;
    } else {
        x_rot = x_trans * m.oc + y_trans * m.os;
        y_rot = -x_trans * m.os + y_trans * m.oc;
    }
}

__device__ __forceinline__ bool fast_cell(const MapView& m, double x_rot, double y_rot, int& idx) {
    const unsigned ux = __double2uint_rz(x_rot * m.inv_fx);
    const unsigned uy = __double2uint_rz(y_rot * m.inv_fx);
    idx = (int)(uy >> m.fx_bits) * m.W + (int)(ux >> m.fx_bits);
    // (f - 1) <= mask - 2  <=>  1 <= f <= mask - 1 for the fraction f
    return ((ux & m.fx_mask) - 1u) <= m.fx_mask - 2u && ((uy & m.fx_mask) - 1u) <= m.fx_mask - 2u && ux < m.w_fx && uy < m.h_fx;
}

// Philox2x32-10 (Salmon et al. 2011), counter-based: no per-ray generator state in HBM.  One call yields the
// 64 bits one Box-Muller sample needs, at half the integer multiplies of Philox4x32.
__device__ __forceinline__ uint2 philox2x32_10(uint2 ctr, uint32_t key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi = __umulhi(0xD256D193u, ctr.x), lo = 0xD256D193u * ctr.x;
        ctr = make_uint2(hi ^ key ^ ctr.y, lo);
        key += 0x9E3779B9u;
    }
    return ctr;
}

__device__ __forceinline__ float gaussian_from_bits(uint32_t a, uint32_t b) {
    const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
    const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530717958647692f * u2);
}

// n / d for a run-time d through the multiply-shift pair the host precomputed (Granlund & Montgomery 1994)
__device__ __forceinline__ unsigned fast_div(unsigned n, FastDiv d) {
    const unsigned t = __umulhi(n, d.mul);
    return (t + ((n - t) >> d.sh1)) >> d.sh2;
}

// _pack_flat_obs lidar channel, f110_env.py:557-560
__device__ __forceinline__ float obs_lidar(double range, float lm) {
    float rf = (float)range;
    if (rf != rf) rf = lm;
    else if (isinf(rf)) rf = rf > 0 ? lm : 0.0f;
    rf = rf < 0.0f ? 0.0f : rf;
    rf = rf > lm ? lm : rf;
    return rf / lm;
}

#ifndef HEAVY_ITERS
#define HEAVY_ITERS 48u
#endif
#ifndef LIDAR_MAX_THREADS
#define LIDAR_MAX_THREADS 128
#endif
#ifndef LIDAR_MIN_BLOCKS
#define LIDAR_MIN_BLOCKS 12
#endif
// COUNT : count dt lookups (roofline L-bar)            IDENT : map origin yaw == 0
// DIRECT: A == 1 (env == s, no opponent ray-cast can follow, no fp64 scratch copy of the scan)
template <bool COUNT, bool IDENT, bool DIRECT>
__global__ void __launch_bounds__(LIDAR_MAX_THREADS, LIDAR_MIN_BLOCKS) lidar_kernel(SimConst c, MapView m, SimState st, StepScratch sc, F110StepIO io) {
    cudaGridDependencySynchronize();   // PDL: everything below reads what the dynamics kernel (and the previous step) wrote
    const unsigned total = (unsigned)c.NA * (unsigned)c.B;
    // ---- work-unit selection (one unit = 32 consecutive rays = one warp).  Ray lengths are heavy-tailed (median 4
    // lookups, p99 42, max ~300 on the Shanghai map), so a long ray that starts in the last wave of CTAs leaves
    // most SMs idle while it finishes.  A ray's length changes little from one step to the next, so every warp
    // records whether its unit was long (>= HEAVY_ITERS lookups) and the next step's grid runs those units FIRST,
    // in a front region of sc.front_units warps (the post kernel latches the count, the dynamics kernel publishes the
    // list); the remaining warps walk the units in natural order and skip the ones the front region took.  Only the
    // launch order depends on this history, never a result.
    const unsigned gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31u;
    unsigned unit;
    if (gwarp < sc.front_units) {
        if (gwarp >= sc.heavy_cnt[0]) return;
        unit = sc.heavy_list[gwarp];
    } else {
        unit = gwarp - sc.front_units;
        if (unit >= sc.num_units) return;
        if (sc.unit_heavy[unit]) return;
    }
    const unsigned r = unit * 32u + lane;
    unsigned nlook = 0;
    bool live = r < total;
    unsigned s = 0, i = 0;
    unsigned env = 0;
    if (live) {
        s = fast_div(r, c.div_B);
        i = r - s * (unsigned)c.B;
        env = DIRECT ? s : fast_div(s, c.div_A);
        if (io.active_mask && !io.active_mask[env]) live = false;
    }
    if (live) {
        // beam direction: closed form of the reference's running sum theta_index += increment with wrap
        // (laser_models.py:174-184).  The running sum differs from the closed form by < 1.3e-10 after
        // 1080 adds; only when the value sits within 1e-9 of an integer can int() disagree, and then the
        // sum is replayed exactly.
        const double t0 = sc.theta0[s];
        const double td = (double)c.theta_dis;
        double t = t0 + (double)i * c.theta_inc;
        if (t >= td) t -= td;
        if (fabs(t - rint(t)) < 1e-9) {
            t = t0;
            for (unsigned k = 0; k < i; ++k) {
                t += c.theta_inc;
                while (t >= td) t -= td;
            }
        }
        int ti = (int)t;
        // memory safety only: a NaN yaw (a car poisoned by a NaN command) converts to a NEGATIVE index on this hardware
        ti = (unsigned)ti < (unsigned)c.theta_dis ? ti : c.theta_dis - 1;
        const double sn = __ldg(c.sines + ti);
        const double cs = __ldg(c.cosines + ti);

        // trace_ray, laser_models.py:129-144
        double x = sc.scan_x[s], y = sc.scan_y[s];
        const double eps = c.eps, max_range = c.max_range;
        // The hot loop holds only the guarded fixed-point lookup and no call: a lookup that lands in the guard band
        // leaves it for good and the ray is finished by the second loop, in the reference's own arithmetic (two IEEE
        // divisions per lookup).  A call inside the hot loop costs 5 % of the kernel: everything live across it has to
        // sit in callee-saved registers or be re-read, lookup after lookup, for a path 4 in 2^21 lookups take.
        double x_rot, y_rot, d = 1.0, total_d = 0.0;
        int idx;
        map_frame<IDENT>(m, x, y, x_rot, y_rot);
        bool fast = fast_cell(m, x_rot, y_rot, idx);
        while (fast) {
            d = __ldg(m.dt + idx);
            total_d += d;
            ++nlook;
            if (!(d > eps && total_d <= max_range)) break;
            x += d * cs;
            y += d * sn;
            map_frame<IDENT>(m, x, y, x_rot, y_rot);
            fast = fast_cell(m, x_rot, y_rot, idx);
        }
        if (!fast) {
            for (;;) {
                d = __ldg(m.dt + cell_index_exact(x_rot, y_rot, m.res, m.wres, m.hres, m.W, m.last));
                total_d += d;
                ++nlook;
                if (!(d > eps && total_d <= max_range)) break;
                x += d * cs;
                y += d * sn;
                map_frame<IDENT>(m, x, y, x_rot, y_rot);
            }
        }
        if (total_d > max_range) total_d = max_range;

        // scan += noise, laser_models.py:450-452
        double range = total_d;
        if (io.noise) {
            range += io.noise[r];
        } else if (c.noise_std > 0.0) {
            // counter = (ray id, steps since the env's reset): like the reference's generator, which is re-seeded by
            // reset (base_classes.py:204), the stream restarts with every episode; unlike it, every ray has its own
            const uint2 bits = philox2x32_10(make_uint2(r, st.step_count[env]), c.noise_key);
            range += c.noise_std * (double)gaussian_from_bits(bits.x, bits.y);
        }
        // The scan goes straight to the caller's buffers.  With opponents (A >= 2) the post kernel lowers the few beams
        // that hit another car afterwards, from the fp64 copy kept in scratch.
        // Streaming stores (evict-first): nothing in this kernel reads the outputs back, and at 32 768 envs the 142 MB of
        // observations would otherwise push the map out of L2 (2.5 % of the kernel there).
        if (io.scans_f64) __stcs(io.scans_f64 + r, range);
        if (io.scans_f32) __stcs(io.scans_f32 + r, (float)range);
        if (DIRECT) {
            if (io.obs) __stcs(io.obs + (size_t)s * (c.B + 8) + i, obs_lidar(range, c.lidar_max));
        } else {
            sc.scan[r] = range;
            if (io.obs && s == env * (unsigned)c.A) __stcs(io.obs + (size_t)env * (c.B + 8) + i, obs_lidar(range, c.lidar_max));
        }

        // check_ttc_jit, laser_models.py:205-213 (any-reduction; the reference's early break is irrelevant)
        const double vel = st.x[3][s];
        if (vel != 0.0) {
            const double proj_vel = vel * __ldg(c.beam_cos + i);
            const double num = range - __ldg(c.side_dist + i);
            // |ttc| > 2*thresh whenever num > 2*thresh*|proj_vel|: then ttc is either >= thresh or negative and the
            // exact quotient need not be formed (the negation keeps NaNs on the exact path)
            if (!(num > 2.0 * c.ttc_thresh * fabs(proj_vel))) {
                const double ttc = num / proj_vel;
                if ((ttc < c.ttc_thresh) && (ttc >= 0.0)) sc.ttc_hit[s] = 1;
            }
        }
    }
    // ---- history for the next step's launch order
    const unsigned wmax = __reduce_max_sync(0xffffffffu, nlook);
    if (lane == 0) {
        bool heavy = wmax >= HEAVY_ITERS;
        if (heavy) {
            const unsigned slot = atomicAdd(sc.heavy_cnt + 1, 1u);
            if (slot < sc.front_units) sc.heavy_list[sc.front_units + slot] = unit;
            else heavy = false;
        }
        sc.unit_heavy[sc.num_units + unit] = heavy ? 1 : 0;
    }
    if (COUNT) {
        const unsigned wsum = __reduce_add_sync(0xffffffffu, nlook);
        const unsigned wrays = __popc(__ballot_sync(0xffffffffu, live));
        if (lane == 0 && wrays) {
            atomicAdd(sc.lookups, (unsigned long long)wsum);
            atomicAdd(sc.lookups + 1, (unsigned long long)wrays);
            atomicMax(sc.lookups + 2, (unsigned long long)wmax);   // longest ray seen so far
        }
    }
}

// ---------------------------------------------------------------- K3: post

// the lidar kernel of this step is complete: latch how many heavy units it recorded and re-arm the counter
__device__ __forceinline__ void latch_launch_order(const StepScratch& sc) {
    const unsigned n = sc.heavy_cnt[1];
    sc.heavy_cnt[0] = n < sc.front_units ? n : sc.front_units;
    sc.heavy_cnt[1] = 0u;
}

// get_trmtx + get_vertices, collision_models.py:218-260; order rl, rr, fr, fl
__device__ __forceinline__ void get_vertices(double px, double py, double yaw, double length, double width, double* v) {
    double sn, cs;
    sincos(yaw, &sn, &cs);
    const double hx[4] = { -length / 2, -length / 2, length / 2, length / 2 };
    const double hy[4] = { width / 2, -width / 2, -width / 2, width / 2 };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[2 * k] = cs * hx[k] + (-sn) * hy[k] + px;
        v[2 * k + 1] = sn * hx[k] + cs * hy[k] + py;
    }
}

__device__ __forceinline__ int furthest_point(const double* v, double dx, double dy) {  // np.argmax: first maximum
    int best = 0;
    double bv = v[0] * dx + v[1] * dy;
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        const double d = v[2 * k] * dx + v[2 * k + 1] * dy;
        if (d > bv) { bv = d; best = k; }
    }
    return best;
}

__device__ __forceinline__ void gjk_support(const double* v1, const double* v2, double dx, double dy, double& ox, double& oy) {
    const int i = furthest_point(v1, dx, dy);
    const int j = furthest_point(v2, -dx, -dy);
    ox = v1[2 * i] - v2[2 * j];
    oy = v1[2 * i + 1] - v2[2 * j + 1];
}

// tripleProduct(a, b, c) = b*(a.c) - a*(b.c), collision_models.py:52-64
__device__ __forceinline__ void triple_product(double ax, double ay, double bx, double by, double cx, double cy,
                                               double& ox, double& oy) {
    const double ac = ax * cx + ay * cy;
    const double bc = bx * cx + by * cy;
    ox = bx * ac - ax * bc;
    oy = by * ac - ay * bc;
}

// collision (GJK), collision_models.py:113-182
__device__ bool gjk_collision(const double* v1, const double* v2) {
    double s0x, s0y, s1x = 0., s1y = 0.;   // simplex rows 0 and 1 (row 2 is always the newest point a)
    double dx = (v1[0] + v1[2] + v1[4] + v1[6]) / 4 - (v2[0] + v2[2] + v2[4] + v2[6]) / 4;
    double dy = (v1[1] + v1[3] + v1[5] + v1[7]) / 4 - (v2[1] + v2[3] + v2[5] + v2[7]) / 4;
    if (dx == 0 && dy == 0) dx = 1.0;
    double ax, ay;
    gjk_support(v1, v2, dx, dy, ax, ay);
    s0x = ax; s0y = ay;
    if (dx * ax + dy * ay <= 0) return false;
    dx = -ax; dy = -ay;
    int index = 0;
    for (int iter = 0; iter < 1000;) {
        gjk_support(v1, v2, dx, dy, ax, ay);
        index += 1;
        if (dx * ax + dy * ay <= 0) return false;
        const double aox = -ax, aoy = -ay;
        if (index < 2) {
            s1x = ax; s1y = ay;
            const double abx = s0x - ax, aby = s0y - ay;
            triple_product(abx, aby, aox, aoy, abx, aby, dx, dy);
            if (sqrt(dx * dx + dy * dy) < 1e-10) { dx = aby; dy = -1 * abx; }
            continue;   // the reference does not count this pass (:154-160)
        }
        const double abx = s1x - ax, aby = s1y - ay;
        const double acx = s0x - ax, acy = s0y - ay;
        double px, py;
        triple_product(abx, aby, acx, acy, acx, acy, px, py);      // acperp
        if (px * aox + py * aoy >= 0) {
            dx = px; dy = py;
        } else {
            triple_product(acx, acy, abx, aby, abx, aby, px, py);  // abperp
            if (px * aox + py * aoy < 0) return true;
            s0x = s1x; s0y = s1y;
            dx = px; dy = py;
        }
        s1x = ax; s1y = ay;   // simplex[1] = simplex[2]
        index -= 1;
        ++iter;
    }
    return false;
}

// index of the first minimum of |scan_angles[k] - a| over a strictly increasing table (np.argmin semantics).
// |ang[k] - a| is V-shaped over a monotone table, so a local minimum is the global one: start from the index a
// uniform table would give and walk; ties resolve to the lower index, as argmin does.
__device__ __forceinline__ int nearest_beam(const double* __restrict__ ang, int B, double a) {
    if (a != a) return 0;
    const double a0 = __ldg(ang), a1 = __ldg(ang + B - 1);
    double g = (a - a0) / (a1 - a0) * (double)(B - 1);
    g = g < 0. ? 0. : (g > (double)(B - 1) ? (double)(B - 1) : g);
    int k = (int)(g + 0.5);
    // the three candidates of a uniform table, loaded together
    const int km = k > 0 ? k - 1 : 0, kp = k + 1 < B ? k + 1 : B - 1;
    double dm = fabs(__ldg(ang + km) - a), dk = fabs(__ldg(ang + k) - a), dp = fabs(__ldg(ang + kp) - a);
    if (dp < dk) { k = kp; dk = dp; } else if (dm <= dk && km != k) { k = km; dk = dm; }
    while (k + 1 < B) {
        const double dn = fabs(__ldg(ang + k + 1) - a);
        if (dn < dk) { ++k; dk = dn; } else break;
    }
    while (k > 0) {
        const double dn = fabs(__ldg(ang + k - 1) - a);
        if (dn <= dk) { --k; dk = dn; } else break;
    }
    return k;
}

// get_range (+ are_collinear), laser_models.py:230-280; v3 = (cos, sin)(beam_theta + pi/2) hoisted by the caller
__device__ __forceinline__ double get_range(double ox, double oy, double v3x, double v3y,
                                            double vax, double vay, double vbx, double vby) {
    const double v1x = ox - vax, v1y = oy - vay;
    const double v2x = vbx - vax, v2y = vby - vay;
    const double denom = v2x * v3x + v2y * v3y;
    double distance = INFINITY;
    if (fabs(denom) > 0.0) {
        // the reference forms d1 = cross/denom and d2 = dot/denom and accepts d1 >= 0 && 0 <= d2 <= 1.  Those tests do not
        // need the quotients: a rounded quotient is >= 0 (-0.0 included) iff the numerator is zero or has the divisor's
        // sign, and it is <= 1 iff |numerator| <= |divisor| (a larger numerator is at least one ulp larger, so the quotient
        // rounds above 1).  Most edges are missed by the beam, so the one division left is rarely executed.
        const double crs = v2x * v1y - v2y * v1x;
        const double dot = v1x * v3x + v1y * v3y;
        const bool neg = denom < 0.0;
        const bool d1_ok = crs == 0.0 || ((crs < 0.0) == neg);
        const bool d2_ok = (dot == 0.0 || ((dot < 0.0) == neg)) && fabs(dot) <= fabs(denom);
        if (d1_ok && d2_ok) distance = crs / denom;
    } else {
        const double bax = vax - ox, bay = vay - oy;
        const double cax = ox - vbx, cay = oy - vby;
        if (fabs(bax * cay - bay * cax) < 1e-8) {
            const double da = sqrt((vax - ox) * (vax - ox) + (vay - oy) * (vay - oy));
            const double db = sqrt((vbx - ox) * (vbx - ox) + (vby - oy) * (vby - oy));
            distance = da < db ? da : db;
        }
    }
    return distance;
}

constexpr int POST_THREADS = 128;

// Shared memory of one env inside a post_kernel CTA (doubles first, then ints; sized by A).
struct EnvSmem {
    double (*pose)[3];   // [A]    own pose AFTER iTTC zeroing (ray-cast origin, base_classes.py:225)
    double (*pre)[3];    // [A]    Simulator.agent_poses: BEFORE iTTC zeroing (:587)
    double (*verts)[8];  // [A*A]  opponent b as seen by a: a's own length/width (:223)
    int* ind;            // [4*A*A] nearest beam of each vertex
    int* lo;             // [A*A]  blocked-view window (get_blocked_view_indices)
    int* hi;
    int* cone;           // [4*A*A] beam-index intervals [f0, f1], [b0, b1] of the forward / backward cones
    int* coll;           // [A]    GJK flags, later GJK | iTTC
    int* hit;            // [A]    iTTC flags
    int* lapdone;        // [A]
};
__host__ __device__ constexpr int post_smem_doubles(int A) { return 6 * A + 8 * A * A + (10 * A * A + 3 * A + 1) / 2; }
__device__ __forceinline__ EnvSmem env_smem(double* base, int A) {
    EnvSmem e;
    e.pose = reinterpret_cast<double (*)[3]>(base);
    e.pre = reinterpret_cast<double (*)[3]>(base + 3 * A);
    e.verts = reinterpret_cast<double (*)[8]>(base + 6 * A);
    e.ind = reinterpret_cast<int*>(base + 6 * A + 8 * A * A);
    e.lo = e.ind + 4 * A * A;
    e.hi = e.lo + A * A;
    e.cone = e.hi + A * A;
    e.coll = e.cone + 4 * A * A;
    e.hit = e.coll + A;
    e.lapdone = e.hit + A;
    return e;
}
// envs per CTA: enough of them that the 5 A^2 scalar work items of stage B fill the CTA's lanes
__host__ __device__ constexpr int post_envs_per_cta(int A) {
    return (POST_THREADS / (5 * A * A)) < 1 ? 1 : ((POST_THREADS / (5 * A * A)) > 16 ? 16 : (POST_THREADS / (5 * A * A)));
}

__global__ void __launch_bounds__(POST_THREADS) post_kernel(SimConst c, SimState st, StepScratch sc, F110StepIO io) {
    cudaGridDependencySynchronize();
    const int A = c.A, B = c.B;
    const int tid = threadIdx.x;
    const int epc = post_envs_per_cta(A);
    const int env0 = blockIdx.x * epc;
    if (blockIdx.x == 0 && tid == 0) latch_launch_order(sc);
    extern __shared__ double s_dyn[];
    const int stride = post_smem_doubles(A);
    __shared__ int s_active[16];
    if (tid < epc) {
        const int env = env0 + tid;
        s_active[tid] = env < c.N && !(io.active_mask && !io.active_mask[env]);
    }
    __syncthreads();

    // ---- stage A: iTTC consequences, RaceCar.check_ttc base_classes.py:243-252.  One thread per (env, agent).
    for (int t = tid; t < epc * A; t += POST_THREADS) {
        const int le = t / A, a = t - le * A;
        if (!s_active[le]) continue;
        const EnvSmem e = env_smem(s_dyn + le * stride, A);
        const int s = (env0 + le) * A + a;
        const int hit = sc.ttc_hit[s];
        const double px = st.x[0][s], py = st.x[1][s];
        const double yaw_pre = sc.pre_yaw[s];
        if (hit) { st.x[3][s] = 0.; st.x[4][s] = 0.; st.x[5][s] = 0.; st.x[6][s] = 0.; }
        e.pose[a][0] = px; e.pose[a][1] = py; e.pose[a][2] = hit ? 0. : yaw_pre;
        e.pre[a][0] = px; e.pre[a][1] = py; e.pre[a][2] = yaw_pre;
        e.hit[a] = hit;
        e.coll[a] = 0;
    }
    __syncthreads();

    // ---- stage B: all-pairs GJK (check_collision :549-563) and the blocked-view window of every opponent b as seen
    // from a (get_blocked_view_indices, laser_models.py:282-315): one thread per (env, a, b, vertex) and per GJK pair
    const int nvert = A * A * 4, nitem = nvert + A * A;
    for (int t = tid; t < epc * nitem; t += POST_THREADS) {
        const int le = t / nitem, u = t - le * nitem;
        if (!s_active[le]) continue;
        const EnvSmem e = env_smem(s_dyn + le * stride, A);
        if (u < nvert) {
            const int a = u / (4 * A), b = (u >> 2) - a * A, k = u & 3;
            if (a == b) continue;
            // vertex k of opponent b, a's own length/width (ray_cast_agents :223); order rl, rr, fr, fl
            const double L = __ldg(c.params + a * F110_NUM_PARAMS + P_LENGTH), Wd = __ldg(c.params + a * F110_NUM_PARAMS + P_WIDTH);
            double sb, cb;
            sincos(e.pre[b][2], &sb, &cb);
            const double hx = (k < 2) ? -L / 2 : L / 2;
            const double hy = (k == 0 || k == 3) ? Wd / 2 : -Wd / 2;
            const double vxw = cb * hx + (-sb) * hy + e.pre[b][0];
            const double vyw = sb * hx + cb * hy + e.pre[b][1];
            e.verts[a * A + b][2 * k] = vxw;
            e.verts[a * A + b][2 * k + 1] = vyw;
            // arctan2(sin(yaw), cos(yaw)) == yaw for the wrapped yaw and arctan2(v/|v|) == arctan2(v), each to an ulp; the
            // angle only selects the nearest beam, so the shorter dependent chain cannot change a window except at an
            // exact tie between two beams
            const double vx = vxw - e.pose[a][0], vy = vyw - e.pose[a][1];
            double ang = e.pose[a][2] - atan2(vy, vx);
            if (vx == 0. && vy == 0.) ang = nan("");   // the reference divides by the zero norm -> NaN -> argmin 0
            if (ang > F110_PI) ang = ang - 2 * F110_PI;
            else if (ang < -F110_PI) ang = ang + 2 * F110_PI;
            e.ind[u] = nearest_beam(c.scan_angles, B, -ang);
        } else {
            const int p = u - nvert, a = p / A, b = p - a * A;
            if (a < b) {
                double va[8], vb[8];
                const double L = __ldg(c.sim_params + P_LENGTH), Wd = __ldg(c.sim_params + P_WIDTH);
                get_vertices(e.pre[a][0], e.pre[a][1], e.pre[a][2], L, Wd, va);
                get_vertices(e.pre[b][0], e.pre[b][1], e.pre[b][2], L, Wd, vb);
                if (gjk_collision(va, vb)) { e.coll[a] = 1; e.coll[b] = 1; }
            }
        }
    }
    __syncthreads();
    for (int t = tid; t < epc * A * A; t += POST_THREADS) {
        const int le = t / (A * A), u = t - le * A * A;
        if (!s_active[le]) continue;
        const EnvSmem e = env_smem(s_dyn + le * stride, A);
        const int* q = e.ind + 4 * u;
        const int a = u / A, b = u - a * A;
        if (a == b) { e.lo[u] = B; e.hi[u] = -1; e.cone[4 * u] = B; e.cone[4 * u + 1] = -1; e.cone[4 * u + 2] = B; e.cone[4 * u + 3] = -1; continue; }
        e.lo[u] = min(min(q[0], q[1]), min(q[2], q[3]));
        e.hi[u] = max(max(q[0], q[1]), max(q[2], q[3]));
        // The reference walks every beam of the window (all 1080 when the opponent is behind the car) although only
        // beams whose LINE crosses the opponent can be lowered: a proper hit needs the ray to enter the car's bounding
        // circle, and the collinear fallback of get_range (:270-274) needs the beam's line to contain an edge, forwards
        // or backwards.  Beams further than alpha (+1e-6 rad) from both the bearing of the circle's centre and its
        // opposite are never visited -- for them get_range returns inf on all four edges.
        const double L = __ldg(c.params + a * F110_NUM_PARAMS + P_LENGTH), Wd = __ldg(c.params + a * F110_NUM_PARAMS + P_WIDTH);
        const double R = 0.5 * sqrt(L * L + Wd * Wd) * (1.0 + 1e-9) + 1e-9;
        const double dx = e.pre[b][0] - e.pose[a][0], dy = e.pre[b][1] - e.pose[a][1];
        const double dist = sqrt(dx * dx + dy * dy);
        // cone -> beam-index intervals (supersets by one beam each side), clipped to the reference window
        int f0 = e.lo[u], f1 = e.hi[u], b0 = B, b1 = -1;
        const double alpha = asin(R / dist) + 1e-6;
        if (dist > R * (1.0 + 1e-6) && alpha < 0.75) {
            // beam angles lie inside (-pi + 0.75, pi - 0.75) (fov <= 4.7 rad): a cone of half-width < 0.75 around a
            // bearing in [-pi, pi] can only meet them un-wrapped
            double phi = atan2(dy, dx) - e.pose[a][2];           // in [-2 pi, 2 pi] -> [-pi, pi]
            if (phi > F110_PI) phi -= 2 * F110_PI;
            else if (phi < -F110_PI) phi += 2 * F110_PI;
            const double back = phi > 0. ? phi - F110_PI : phi + F110_PI;
            const double amin = __ldg(c.scan_angles), amax = __ldg(c.scan_angles + B - 1);
            if (phi + alpha < amin || phi - alpha > amax) { f0 = B; f1 = -1; }
            else {
                f0 = max(f0, nearest_beam(c.scan_angles, B, phi - alpha) - 1);
                f1 = min(f1, nearest_beam(c.scan_angles, B, phi + alpha) + 1);
            }
            if (!(back + alpha < amin || back - alpha > amax)) {
                b0 = max(e.lo[u], nearest_beam(c.scan_angles, B, back - alpha) - 1);
                b1 = min(e.hi[u], nearest_beam(c.scan_angles, B, back + alpha) + 1);
                if (b0 <= f1 && f0 <= b1) { f0 = min(f0, b0); f1 = max(f1, b1); b0 = B; b1 = -1; }   // merge overlap
            }
        }   // else: overlapping cars, NaN, or a cone too wide to reason about -> the whole reference window
        e.cone[4 * u] = f0; e.cone[4 * u + 1] = f1; e.cone[4 * u + 2] = b0; e.cone[4 * u + 3] = b1;
    }
    __syncthreads();

    // ---- stage C: finish zone / laps per agent, _check_done f110_env.py:320-348
    for (int t = tid; t < epc * A; t += POST_THREADS) {
        const int le = t / A, a = t - le * A;
        if (!s_active[le]) continue;
        const EnvSmem e = env_smem(s_dyn + le * stride, A);
        const int env = env0 + le;
        const int s = env * A + a;
        const double new_time = st.time[env] + c.timestep;   // :406 (written back in stage E, after the barrier)
        const int col = (e.coll[a] | e.hit[a]) ? 1 : 0;
        const double dxp = e.pose[a][0] - st.start_x[s];
        const double dyp = e.pose[a][1] - st.start_y[s];
        const double rc = st.rot_c[env], rs = st.rot_s[env];
        // start_rot = [[cos(-t), -sin(-t)], [sin(-t), cos(-t)]]
        const double lx = rc * dxp + (-rs) * dyp;
        double ty = rs * dxp + rc * dyp;
        if (ty > 2) ty -= 2;
        else if (ty < -2) ty = -2 - ty;
        else ty = 0;
        const double dist2 = lx * lx + ty * ty;
        const bool closes = dist2 <= 0.1;
        int near = st.near_start[s];
        int tog = st.toggles[s];
        if (closes && !near) { near = 1; tog += 1; }
        else if (!closes && near) { near = 0; tog += 1; }
        st.near_start[s] = (uint8_t)near;
        st.toggles[s] = tog;
        const double lapc = (double)(tog / 2);
        st.lap_counts[s] = lapc;
        double lapt = st.lap_times[s];
        if (tog < 4) { lapt = new_time; st.lap_times[s] = lapt; }
        st.collisions[s] = (uint8_t)col;
        e.lapdone[a] = tog >= 4;
        if (io.collisions) io.collisions[s] = (uint8_t)col;
        if (io.toggles) io.toggles[s] = tog;
        if (io.lap_times) io.lap_times[s] = lapt;
        if (io.lap_counts) io.lap_counts[s] = lapc;
        if (io.state) {
            double* o = io.state + (size_t)s * 7;
#pragma unroll
            for (int k = 0; k < 7; ++k) o[k] = st.x[k][s];   // this thread wrote the zeroed entries itself in stage A
        }
    }

    // ---- stage D: opponent ray-cast (ray_cast_agents :206-227).  The lidar kernel already wrote every scan; only the
    // beams inside some opponent's blocked-view window can get shorter, so only those are re-read (fp64 scratch
    // copy), lowered and re-written.
    const float lm = c.lidar_max;
    for (int le = 0; le < epc; ++le) {
      if (!s_active[le]) continue;
      const EnvSmem e = env_smem(s_dyn + le * stride, A);
      const int env = env0 + le;
      for (int a = 0, u = 0; a < A; ++a)
      for (int b = 0; b < A; ++b, ++u) {           // nested counters: no integer divisions in this 24-trip loop
        if (a == b) continue;
        const double* v = e.verts[u];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int i0 = e.cone[4 * u + 2 * half], i1 = e.cone[4 * u + 2 * half + 1];
            // beam i always belongs to thread i % POST_THREADS, so successive opponents of the same car lower a beam
            // through the same thread, in the reference's order, via the scratch copy
            for (int i = (i0 & ~(POST_THREADS - 1)) + tid; i <= i1; i += POST_THREADS) {
                if (i < i0) continue;
                const size_t g = (size_t)(env * A + a) * B + i;
                const double range0 = sc.scan[g];
                double range = range0;
                double v3x, v3y;
                sincos(e.pose[a][2] + __ldg(c.scan_angles + i) + F110_PI / 2., &v3y, &v3x);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int j2 = (j + 1) & 3;
                    const double rr = get_range(e.pose[a][0], e.pose[a][1], v3x, v3y, v[2 * j], v[2 * j + 1], v[2 * j2], v[2 * j2 + 1]);
                    if (rr < range) range = rr;
                }
                if (range < range0) {
                    sc.scan[g] = range;
                    if (io.scans_f64) io.scans_f64[g] = range;
                    if (io.scans_f32) io.scans_f32[g] = (float)range;
                    if (a == 0 && io.obs) io.obs[(size_t)env * (B + 8) + i] = obs_lidar(range, lm);
                }
            }
        }
      }
    }
    __syncthreads();

    // ---- stage E: done, reward, time, pose part of the observation, episode statistics.  One thread per env.
    if (tid < epc && s_active[tid]) {
        const EnvSmem e = env_smem(s_dyn + tid * stride, A);
        const int env = env0 + tid;
        const double new_time = st.time[env] + c.timestep;
        bool all_laps = true;
        for (int a = 0; a < A; ++a) all_laps = all_laps && e.lapdone[a];
        const bool ego_col = (e.coll[c.ego] | e.hit[c.ego]) != 0;
        const bool done = ego_col || all_laps;   // f110_env.py:350
        st.time[env] = new_time;
        if (io.time) io.time[env] = new_time;
        if (io.reward) io.reward[env] = (float)c.timestep;
        if (io.terminated) io.terminated[env] = done ? 1 : 0;
        if (io.obs) {   // f110_env.py:563-579: slots for agents 0 and 1
            float* q = io.obs + (size_t)env * (B + 8) + B;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                q[4 * k] = (float)e.pose[k][0];
                q[4 * k + 1] = (float)e.pose[k][1];
                q[4 * k + 2] = (float)wrap_angle(e.pose[k][2]);
                q[4 * k + 3] = (e.coll[k] | e.hit[k]) ? 1.0f : 0.0f;
            }
        }
        if (done) {
            atomicAdd(sc.stats + F110_STAT_EPISODES, 1.0);
            atomicAdd(sc.stats + F110_STAT_EPISODE_STEPS, (double)st.step_count[env] + 1.0);
            atomicAdd(sc.stats + F110_STAT_EPISODE_TIME, new_time);
            if (ego_col) atomicAdd(sc.stats + F110_STAT_EGO_COLLISIONS, 1.0);
            if (all_laps) atomicAdd(sc.stats + F110_STAT_LAPS_DONE, 1.0);
        }
    }
}

// K3 for A == 1: nothing to ray-cast and no pair to test, the lidar kernel already wrote the scans -> one thread
// per env applies the iTTC consequence and the finish-zone / done bookkeeping (same statements as post_kernel).
__global__ void __launch_bounds__(128) post_single_kernel(SimConst c, SimState st, StepScratch sc, F110StepIO io) {
    cudaGridDependencySynchronize();
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env == 0) latch_launch_order(sc);
    if (env >= c.N) return;
    if (io.active_mask && !io.active_mask[env]) return;
    const int s = env;
    const int hit = sc.ttc_hit[s];
    const double px = st.x[0][s], py = st.x[1][s];
    double yaw = sc.pre_yaw[s];
    if (hit) { st.x[3][s] = 0.; st.x[4][s] = 0.; st.x[5][s] = 0.; st.x[6][s] = 0.; yaw = 0.; }
    const double new_time = st.time[env] + c.timestep;
    const double dxp = px - st.start_x[s];
    const double dyp = py - st.start_y[s];
    const double rc = st.rot_c[env], rs = st.rot_s[env];
    const double lx = rc * dxp + (-rs) * dyp;
    double ty = rs * dxp + rc * dyp;
    if (ty > 2) ty -= 2;
    else if (ty < -2) ty = -2 - ty;
    else ty = 0;
    const bool closes = (lx * lx + ty * ty) <= 0.1;
    int near = st.near_start[s];
    int tog = st.toggles[s];
    if (closes && !near) { near = 1; tog += 1; }
    else if (!closes && near) { near = 0; tog += 1; }
    st.near_start[s] = (uint8_t)near;
    st.toggles[s] = tog;
    const double lapc = (double)(tog / 2);
    st.lap_counts[s] = lapc;
    double lapt = st.lap_times[s];
    if (tog < 4) { lapt = new_time; st.lap_times[s] = lapt; }
    st.collisions[s] = (uint8_t)hit;
    st.time[env] = new_time;
    const bool done = hit || tog >= 4;
    if (io.collisions) io.collisions[s] = (uint8_t)hit;
    if (io.toggles) io.toggles[s] = tog;
    if (io.lap_times) io.lap_times[s] = lapt;
    if (io.lap_counts) io.lap_counts[s] = lapc;
    if (io.state) {
        double* o = io.state + (size_t)s * 7;
#pragma unroll
        for (int k = 0; k < 7; ++k) o[k] = st.x[k][s];
    }
    if (io.time) io.time[env] = new_time;
    if (io.reward) io.reward[env] = (float)c.timestep;
    if (io.terminated) io.terminated[env] = done ? 1 : 0;
    if (io.obs) {
        float* q = io.obs + (size_t)env * (c.B + 8) + c.B;
        q[0] = (float)px; q[1] = (float)py; q[2] = (float)wrap_angle(yaw); q[3] = hit ? 1.0f : 0.0f;
        q[4] = q[5] = q[6] = q[7] = 0.0f;
    }
    if (done) {
        atomicAdd(sc.stats + F110_STAT_EPISODES, 1.0);
        atomicAdd(sc.stats + F110_STAT_EPISODE_STEPS, (double)st.step_count[env] + 1.0);
        atomicAdd(sc.stats + F110_STAT_EPISODE_TIME, new_time);
        if (hit) atomicAdd(sc.stats + F110_STAT_EGO_COLLISIONS, 1.0);
        if (tog >= 4) atomicAdd(sc.stats + F110_STAT_LAPS_DONE, 1.0);
    }
}

}  // namespace

// Launch with programmatic stream serialisation (PDL): the grid may be scheduled while the previous kernel of the stream
// drains; the kernel itself waits in cudaGridDependencySynchronize() before it touches anything that kernel wrote.
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

void launch_dynamics(const SimConst& c, const SimState& st, const StepScratch& sc, const F110StepIO& io, cudaStream_t s) {
    const int threads = 128;
    launch_pdl(dynamics_kernel, dim3((c.NA + threads - 1) / threads), dim3(threads), 0, s, c, st, sc, io);
}

template <bool COUNT, bool IDENT>
static void launch_lidar_t(const SimConst& c, const MapView& m, const SimState& st, const StepScratch& sc,
                           const F110StepIO& io, unsigned blocks, unsigned threads, cudaStream_t s) {
    if (c.A == 1) launch_pdl(lidar_kernel<COUNT, IDENT, true>, dim3(blocks), dim3(threads), 0, s, c, m, st, sc, io);
    else launch_pdl(lidar_kernel<COUNT, IDENT, false>, dim3(blocks), dim3(threads), 0, s, c, m, st, sc, io);
}

void launch_lidar(const SimConst& c, const MapView& m, const SimState& st, const StepScratch& sc, const F110StepIO& io,
                  bool count_lookups, int threads_per_block, cudaStream_t s) {
    const unsigned threads = (unsigned)threads_per_block;
    const unsigned warps = sc.front_units + sc.num_units;
    const unsigned blocks = (warps * 32u + threads - 1) / threads;
    const bool ident = (m.oc == 1.0 && m.os == 0.0);
    if (count_lookups) {
        if (ident) launch_lidar_t<true, true>(c, m, st, sc, io, blocks, threads, s);
        else launch_lidar_t<true, false>(c, m, st, sc, io, blocks, threads, s);
    } else {
        if (ident) launch_lidar_t<false, true>(c, m, st, sc, io, blocks, threads, s);
        else launch_lidar_t<false, false>(c, m, st, sc, io, blocks, threads, s);
    }
}

void launch_post(const SimConst& c, const SimState& st, const StepScratch& sc, const F110StepIO& io, cudaStream_t s) {
    if (c.A == 1) launch_pdl(post_single_kernel, dim3((c.N + 127) / 128), dim3(128), 0, s, c, st, sc, io);
    else {
        const int epc = post_envs_per_cta(c.A);
        launch_pdl(post_kernel, dim3((c.N + epc - 1) / epc), dim3(POST_THREADS), sizeof(double) * post_smem_doubles(c.A) * epc, s, c, st, sc, io);
    }
}

void launch_sim_reset(const SimConst& c, const SimState& st, const double* poses, const uint8_t* mask, cudaStream_t s) {
    const int threads = 128;
    sim_reset_kernel<<<(c.NA + threads - 1) / threads, threads, 0, s>>>(c, st, poses, mask);
}
