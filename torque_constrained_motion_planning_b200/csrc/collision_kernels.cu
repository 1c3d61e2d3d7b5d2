// collision_kernels.cu -- SURVEY.md 8f-1/8f-2: the synthetic-scene collision predicate on the GPU, fused IN
// FRONT of the torque test so that RRT* tree growth checks whole candidate edges in one launch.
//
// The reference's collision predicate is PyBullet's (utils.get_collision_fn, utils.py:3165-3218), out of scope
// and not installed; planner-level runs use the stand-in defined in collision.py (joint-limit test as in
// utils.py:3177-3178, then link spheres from the DH forward kinematics against axis-aligned boxes / spheres).
// These kernels implement exactly that stand-in, keep the reference's evaluation ORDER -- collision first,
// torque only for collision-free configurations (rrt_star.py:92-96) -- and generate the extend steps of
// utils.get_extend_fn / get_refine_fn (utils.py:3031-3041, 3068-3077) on the fly.
//
//   collision_kernel       1 thread = 1 configuration            -> hit[n]
//   extend_prefix_kernel   1 warp   = 1 candidate edge q1 -> q2; lanes = extend steps (32 per round);
//                          __ballot_sync / __ffs gives the safe-prefix length (safe_path_force_aware,
//                          rrt_star.py:90-98); the warp stops at the first round containing a failure.
#include "panda_model.cuh"
#include "tcmp_internal.h"

namespace tcmp {

// (sin, cos)(i pi / 512) for the table-driven sincos of the static torque tests (panda_model.cuh sincos6_table<GLOBAL>):
// read in place through the read-only cache -- this kernel's shared memory is spoken for, and its torque tests are
// too few per CTA to pay for staging 16 KB.  Own copy per translation unit (no relocatable device code in the build).
static __device__ const SinCos kExtSinCosTable[kSinCosTableSize] = {
#include "sincos_table.inc"
};

struct Scene {
    int n_obs;
    int kind[TCMP_MAX_OBSTACLES];            // 0 = box, 1 = sphere
    double c[TCMP_MAX_OBSTACLES][3];         // centre
    double h[TCMP_MAX_OBSTACLES][3];         // box half extents; sphere: h[0] = radius
    double lo[7], hi[7];                     // joint limits (limits_fn, utils.py:3177)
    double payload_radius;                   // > 0: a sphere at the grasp target (held object)
};

// Link-sphere model of collision.py (_SEGMENTS): spheres strung between consecutive DH frame origins
// 1..7 and the grasp target (8).
__device__ constexpr int kSegCount = 7;
__device__ constexpr int kSegSamples[kSegCount] = {3, 2, 4, 2, 2, 2, 3};
__device__ constexpr double kSegRadius[kSegCount] = {0.07, 0.07, 0.065, 0.06, 0.055, 0.05, 0.05};

__device__ __forceinline__ bool sphere_hits(const Scene &S, double x, double y, double z, double r) {
    bool hit = false;
    for (int o = 0; o < S.n_obs; ++o) {
        const double dx = x - S.c[o][0], dy = y - S.c[o][1], dz = z - S.c[o][2];
        if (S.kind[o] == 1) {
            hit = hit || (sqrt(dx * dx + dy * dy + dz * dz) < r + S.h[o][0]);
        } else {
            const double ex = fmax(fabs(dx) - S.h[o][0], 0.0), ey = fmax(fabs(dy) - S.h[o][1], 0.0),
                         ez = fmax(fabs(dz) - S.h[o][2], 0.0);
            hit = hit || (sqrt(ex * ex + ey * ey + ez * ez) < r);
        }
    }
    return hit;
}

// collision.py get_collision_fn.batch for one configuration
__device__ bool config_collides(const Scene &S, const double (&q)[7]) {
    bool out = false;
#pragma unroll
    for (int j = 0; j < 7; ++j) out = out || (q[j] < S.lo[j]) || (q[j] > S.hi[j]);
    if (out || S.n_obs == 0) return out;
    // origins of DH frames 0..7 and the grasp target (collision.link_frames)
    double O[9][3];
    double X[3] = {1, 0, 0}, Y[3] = {0, 1, 0}, Z[3] = {0, 0, 1}, p[3] = {0, 0, 0};
    const double A[7] = {0, 0, 0, 0.0825, -0.0825, 0, 0.088};
    const double D[7] = {0.333, 0, 0.316, 0, 0.384, 0, 0};
    const int AL[7] = {0, -1, 1, 1, -1, 1, 1};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        double s, c;
        sincos(q[k], &s, &c);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double y = AL[k] == 0 ? Y[r] : (AL[k] > 0 ? Z[r] : -Z[r]);
            const double z = AL[k] == 0 ? Z[r] : (AL[k] > 0 ? -Y[r] : Y[r]);
            p[r] += A[k] * X[r] + D[k] * z;
            const double x = X[r];
            X[r] = c * x + s * y;
            Y[r] = c * y - s * x;
            Z[r] = z;
            O[k][r] = p[r];
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        O[7][r] = p[r] + 0.107 * Z[r];
        O[8][r] = O[7][r] + 0.105 * Z[r];
    }
    bool hit = false;
#pragma unroll
    for (int g = 0; g < kSegCount; ++g) {
        const int a = g + 1, b = g + 2, m = kSegSamples[g];
        for (int i = 0; i < m; ++i) {
            const double t = (double)i / (double)(m - 1);   // np.linspace(0, 1, m)
            const double x = (1 - t) * O[a][0] + t * O[b][0], y = (1 - t) * O[a][1] + t * O[b][1],
                         z = (1 - t) * O[a][2] + t * O[b][2];
            hit = hit || sphere_hits(S, x, y, z, kSegRadius[g]);
        }
    }
    if (S.payload_radius > 0) hit = hit || sphere_hits(S, O[8][0], O[8][1], O[8][2], S.payload_radius);
    return hit;
}

__global__ void __launch_bounds__(128)
collision_kernel(int64_t n, const double *__restrict__ q, Scene S, uint8_t *__restrict__ hit_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double qs[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) qs[j] = __ldg(q + j * n + i);
        hit_out[i] = (uint8_t)config_collides(S, qs);
    }
}

template <bool TOOL, typename P>
__global__ void __launch_bounds__(128)
extend_prefix_kernel(int64_t n_edges, const double *__restrict__ q1, const double *__restrict__ q2, Scene S,
                     double r0, double r1, double r2, double r3, double r4, double r5, double r6, int check_torque,
                     double mass, double payload_threshold, int32_t *__restrict__ n_steps_out,
                     int32_t *__restrict__ prefix_out, const __grid_constant__ P prm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double res[7] = {r0, r1, r2, r3, r4, r5, r6};
    const double mp_inertial = TOOL ? 0.0 : (mass > payload_threshold ? mass : 0.0);
    const double mp_tool = TOOL ? mass : 0.0;
    for (int64_t e = warp; e < n_edges; e += n_warps) {
        double a[7], b[7];
        double nrm2 = 0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            a[j] = __ldg(q1 + j * n_edges + e);
            b[j] = __ldg(q2 + j * n_edges + e);
            const double d = __ddiv_rn(__dadd_rn(b[j], -a[j]), res[j]);   // np.divide(difference, resolutions)
            nrm2 = __dadd_rn(nrm2, __dmul_rn(d, d));
        }
        // steps = int(np.linalg.norm(., ord=2)) (utils.py:3074); the sequence has steps + 1 configurations
        // A non-finite norm (NaN / inf end point) has no step count: report 0 configurations, safe prefix 0, so the
        // host never truncates a sequence with it (the cast of a NaN to int is undefined); a huge finite norm is
        // clamped for the same reason.
        const double nrm = sqrt(nrm2);
        const bool sane = nrm == nrm && nrm < 1.0e6;
        const int N = sane ? (int)nrm + 1 : 0;
        int prefix = N;
        for (int base = 0; base < N; base += 32) {
            const int k = base + lane;
            const bool active = k < N;
            // get_refine_fn recurrence (utils.py:3036-3039): q <- (1 / (N - i)) * (q2 - q) + q, i = 0..k
            double q[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) q[j] = a[j];
            const int last = active ? k : N - 1;
            for (int i = 0; i <= last; ++i) {
                const double w = __ddiv_rn(1.0, (double)(N - i));
#pragma unroll
                for (int j = 0; j < 7; ++j) q[j] = __dadd_rn(__dmul_rn(w, __dadd_rn(b[j], -q[j])), q[j]);
            }
            bool bad = config_collides(S, q);
            if (!bad && check_torque) {   // torque only for collision-free configurations (rrt_star.py:92-96)
                double tau[7];
                const double z[7] = {0, 0, 0, 0, 0, 0, 0};
                rne_core_table<false, TOOL, P, true>(q, z, z, mp_inertial, mp_tool, tau, kExtSinCosTable, prm);   // table read in place (__ldg)
                bad = !limits_ok<double, P>(tau, prm);
            }
            const unsigned fails = __ballot_sync(0xffffffffu, active && bad);
            if (fails) {
                prefix = base + __ffs(fails) - 1;
                break;
            }
        }
        if (lane == 0) {
            n_steps_out[e] = N;
            prefix_out[e] = prefix;
        }
    }
}

static cudaError_t make_scene(int n_obs, const tcmp_obstacle *obs, const double *q_lo, const double *q_hi,
                              double payload_radius, Scene *S) {
    if (n_obs < 0 || n_obs > TCMP_MAX_OBSTACLES) return cudaErrorInvalidValue;
    S->n_obs = n_obs;
    for (int o = 0; o < TCMP_MAX_OBSTACLES; ++o) {
        S->kind[o] = o < n_obs ? obs[o].kind : 0;
        for (int r = 0; r < 3; ++r) {
            S->c[o][r] = o < n_obs ? obs[o].center[r] : 0.0;
            S->h[o][r] = o < n_obs ? obs[o].half[r] : 0.0;
        }
    }
    for (int j = 0; j < 7; ++j) {
        S->lo[j] = q_lo[j];
        S->hi[j] = q_hi[j];
    }
    S->payload_radius = payload_radius;
    return cudaSuccess;
}

cudaError_t launch_collision_batch(int64_t n, const double *q, int n_obs, const tcmp_obstacle *obs,
                                   const double *q_lo, const double *q_hi, double payload_radius, uint8_t *hit_out,
                                   cudaStream_t st) {
    Scene S;
    cudaError_t e = make_scene(n_obs, obs, q_lo, q_hi, payload_radius, &S);
    if (e != cudaSuccess) return e;
    const int grid = grid_for(reinterpret_cast<const void *>(collision_kernel), 128, n);
    collision_kernel<<<grid, 128, 0, st>>>(n, q, S, hit_out);
    return cudaGetLastError();
}

template <bool TOOL, typename P>
static cudaError_t launch_extend_p(int64_t n_edges, const double *q1, const double *q2, const Scene &S, const double *res,
                                   int check, double mass, double payload_threshold, int32_t *n_steps_out,
                                   int32_t *prefix_out, const P &prm, cudaStream_t st) {
    auto kern = extend_prefix_kernel<TOOL, P>;
    const int grid = grid_for(reinterpret_cast<const void *>(kern), 128, n_edges * 32);
    kern<<<grid, 128, 0, st>>>(n_edges, q1, q2, S, res[0], res[1], res[2], res[3], res[4], res[5], res[6], check, mass,
                               payload_threshold, n_steps_out, prefix_out, prm);
    return cudaGetLastError();
}

template <bool TOOL>
static cudaError_t launch_extend_t(int64_t n_edges, const double *q1, const double *q2, const Scene &S, const double *res,
                                   int check, double mass, double payload_threshold, int32_t *n_steps_out,
                                   int32_t *prefix_out, const tcmp_model *model, cudaStream_t st) {
    if (model)
        return launch_extend_p<TOOL>(n_edges, q1, q2, S, res, check, mass, payload_threshold, n_steps_out, prefix_out,
                                     params_from_desc<double>(*model), st);
    return launch_extend_p<TOOL>(n_edges, q1, q2, S, res, check, mass, payload_threshold, n_steps_out, prefix_out,
                                 ConstParams(), st);
}

cudaError_t launch_extend_prefix(int mode, int64_t n_edges, const double *q1, const double *q2,
                                 const double *res, int n_obs, const tcmp_obstacle *obs, const double *q_lo,
                                 const double *q_hi, double payload_radius, double mass, double payload_threshold,
                                 int32_t *n_steps_out, int32_t *prefix_out, cudaStream_t st, const tcmp_model *model) {
    Scene S;
    cudaError_t e = make_scene(n_obs, obs, q_lo, q_hi, payload_radius, &S);
    if (e != cudaSuccess) return e;
    const int check = mode != TCMP_MODE_BASE;
    if (mode == TCMP_MODE_DYN)
        return launch_extend_t<true>(n_edges, q1, q2, S, res, check, mass, payload_threshold, n_steps_out, prefix_out,
                                     model, st);
    return launch_extend_t<false>(n_edges, q1, q2, S, res, check, mass, payload_threshold, n_steps_out, prefix_out,
                                  model, st);
}

}  // namespace tcmp
