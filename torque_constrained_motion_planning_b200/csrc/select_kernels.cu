// select_kernels.cu -- goal-IK post-filter fused behind the IK solver.
//
// Per target pose the reference planner sweeps the free joint (ikfast.py:153-161), solves IK for each
// value, drops solutions outside the joint limits (ikfast.py:167, franka_ik_fast.py:55-57), keeps the
// one nearest to the current configuration (closest_inverse_kinematics, ikfast.py:172-188, max-norm by
// default; select_solution, ik_utils.py:43-52) and then requires the static torque test of that grasp
// configuration (panda_primitives.py:263).  Here one warp owns one pose and its lanes the values of the sweep: each lane
// solves, filters by limits, runs the STATIC torque test (rne_core<double, false, TOOL>) on every
// surviving solution and keeps the nearest feasible one -- no solution set ever leaves the SM.
#include "ik_core.cuh"
#include "panda_model.cuh"
#include "tcmp_internal.h"

namespace tcmp {

struct JointLimits {
    double lo[7], hi[7];
};

// One warp per pose; lanes own the free-joint values of the sweep (32 per round).  Each lane solves, filters
// and torque-tests its <= 8 solutions and keeps its nearest survivor; a shuffle arg-min over (cost, free index)
// picks the pose's winner -- ties resolve to the earliest free value, then solver order, exactly as a serial scan
// in sweep order would.
template <bool TOOL>
__global__ void __launch_bounds__(128)
ik_select_kernel(int64_t n, int n_free, int free_broadcast, const double *__restrict__ rot9,
                 const double *__restrict__ trans3, const double *__restrict__ free_vals,
                 const double *__restrict__ q_ref, int ref_broadcast, JointLimits lim, int check_torque,
                 double mass, double payload_threshold, int use_max_norm, double *__restrict__ best_q,
                 double *__restrict__ best_cost, int32_t *__restrict__ n_valid) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double mp_inertial = TOOL ? 0.0 : (mass > payload_threshold ? mass : 0.0);
    const double mp_tool = TOOL ? mass : 0.0;
    for (int64_t p = warp; p < n; p += n_warps) {
        double R[9], ref[7];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = __ldg(rot9 + i * n + p);
#pragma unroll
        for (int j = 0; j < 7; ++j) ref[j] = ref_broadcast ? __ldg(q_ref + j) : __ldg(q_ref + j * n + p);
        const double tx = __ldg(trans3 + p), ty = __ldg(trans3 + n + p), tz = __ldg(trans3 + 2 * n + p);
        double bq[7] = {0, 0, 0, 0, 0, 0, 0};
        double bcost = INFINITY;
        int bf = 0x7fffffff;   // free index of the lane's best (tie-break key)
        int valid = 0;
        for (int f = lane; f < n_free; f += 32) {
            const double j6 = free_broadcast ? __ldg(free_vals + f) : __ldg(free_vals + (int64_t)f * n + p);
            ik::Pose P;
            ik::prepare_pose(R, tx, ty, tz, j6, P);
            double sols[56];
            ik::Emit out;
            out.sols = sols;
            out.count = 0;
            out.status = 0;
            ik::solve_one(P, out);
            const int cnt = out.count < 8 ? out.count : 8;
            for (int s = 0; s < cnt; ++s) {
                double q[7];
                bool inside = true;
                double c2 = 0.0, cmax = 0.0;
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    q[j] = sols[s * 7 + j];
                    inside = inside && !(q[j] < lim.lo[j]) && !(q[j] > lim.hi[j]);   // violates_limits
                    const double d = fabs(q[j] - ref[j]);
                    c2 += d * d;
                    cmax = fmax(cmax, d);
                }
                if (!inside) continue;
                if (check_torque) {
                    double tau[7];
                    const double z[7] = {0, 0, 0, 0, 0, 0, 0};
                    rne_core<double, false, TOOL>(q, z, z, mp_inertial, mp_tool, tau);
                    if (!within_limits<double>(tau)) continue;
                }
                ++valid;
                const double cost = use_max_norm ? cmax : sqrt(c2);
                if (cost < bcost) {   // strict: the earliest (f, s) wins ties within the lane
                    bcost = cost;
                    bf = f;
#pragma unroll
                    for (int j = 0; j < 7; ++j) bq[j] = q[j];
                }
            }
        }
        // warp arg-min over (cost, free index) and sum of survivors
        double wc = bcost;
        int wf = bf, wl = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double oc = __shfl_xor_sync(0xffffffffu, wc, off);
            const int of = __shfl_xor_sync(0xffffffffu, wf, off);
            const int ol = __shfl_xor_sync(0xffffffffu, wl, off);
            valid += __shfl_xor_sync(0xffffffffu, valid, off);
            if (oc < wc || (oc == wc && of < wf)) { wc = oc; wf = of; wl = ol; }
        }
        if (lane == wl) {
#pragma unroll
            for (int j = 0; j < 7; ++j) best_q[j * n + p] = bq[j];
            best_cost[p] = bcost;
            n_valid[p] = valid;
        }
    }
}

cudaError_t launch_ik_select(int64_t n, const double *rot9, const double *trans3, const double *free_vals,
                             int n_free, int free_broadcast, const double *q_ref, int ref_broadcast,
                             const double *q_lo, const double *q_hi, int mode, double mass,
                             double payload_threshold, int use_max_norm, double *best_q, double *best_cost,
                             int32_t *n_valid, cudaStream_t st) {
    JointLimits lim;
    for (int j = 0; j < 7; ++j) {
        lim.lo[j] = q_lo[j];
        lim.hi[j] = q_hi[j];
    }
    const int check = mode != TCMP_MODE_BASE;
    if (mode == TCMP_MODE_DYN) {
        const int grid = grid_for(reinterpret_cast<const void *>(ik_select_kernel<true>), 128, n * 32);
        ik_select_kernel<true><<<grid, 128, 0, st>>>(n, n_free, free_broadcast, rot9, trans3, free_vals, q_ref,
                                                     ref_broadcast, lim, check, mass, payload_threshold, use_max_norm,
                                                     best_q, best_cost, n_valid);
    } else {
        const int grid = grid_for(reinterpret_cast<const void *>(ik_select_kernel<false>), 128, n * 32);
        ik_select_kernel<false><<<grid, 128, 0, st>>>(n, n_free, free_broadcast, rot9, trans3, free_vals, q_ref,
                                                      ref_broadcast, lim, check, mass, payload_threshold, use_max_norm,
                                                      best_q, best_cost, n_valid);
    }
    return cudaGetLastError();
}

}  // namespace tcmp
