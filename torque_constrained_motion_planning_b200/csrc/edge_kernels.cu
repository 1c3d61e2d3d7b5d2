// edge_kernels.cu -- K4: min-jerk waypoint generation fused in front of the torque test.
//
//  * edge kernel: one warp per RRT* edge (qa -> qb).  Lanes own waypoints; each round of 32
//    waypoints ends in a __ballot_sync / __ffs first-failure reduction, and the warp stops at the
//    first round that contains a failure -- the prefix semantics of safe_path_force_aware
//    (rrt_star.py:90-98) and of the final check (rrt_star.py:208-210).  112 B in, 4 B out per edge:
//    purely FP64-pipe bound.
//  * trajectory kernel: one thread per sample of the piecewise quintic the smoother produced
//    (min_jerk_v2.py:144-182 via panda_primitives.py:299-316); optionally writes q/qd/qdd/tau so
//    the Conf logging pass (utils.py:3376-3378) comes from the same launch.
//
// Sample times follow np.linspace(1/n, 1, n) (min_jerk_v2.py:176): start + i*step with the last
// sample forced to 1.0; __dmul_rn/__dadd_rn keep the two roundings NumPy makes.
#include "panda_model.cuh"
#include "tcmp_internal.h"

namespace tcmp {

template <typename T> __device__ __forceinline__ T linspace_sample(int i, int n, double interval, double step) {
    const double t = (i == n - 1 && n > 1) ? 1.0 : __dadd_rn(interval, __dmul_rn((double)i, step));
    return (T)t;
}

// (sin, cos)(i pi / 512) for the table-driven sincos of the dynamic fp64 edge kernel (see rne_kernels.cu / panda_model.cuh
// sincos6_table); this translation unit carries its own read-only copy (no relocatable device code in the build).
static __device__ const SinCos kEdgeSinCosTable[kSinCosTableSize] = {
#include "sincos_table.inc"
};

// Multi-GPU form: lane 0 stores the edge's first-failure index into the gathered buffer of every rank (peer
// pointers over NVLink, see tcmp_rne_batch_scatter) instead of a local array + a separate NCCL all-gather.
struct IndexDests {
    int32_t *p[TCMP_MAX_PEERS];
    int n;          // 0 = plain local output through `first_fail`
    int64_t offset;
};

// P: ConstParams (compiled-in Panda) or RtParams<T> (caller-supplied inertial set, tcmp_*_model entry points).
template <typename T, bool DYN, bool TOOL, typename P>
__global__ void __launch_bounds__(128)
edge_kernel(int64_t n_edges, int W, double interval, double step, const T *__restrict__ qa,
            const T *__restrict__ qb, T mass, T payload_threshold, int32_t *__restrict__ first_fail,
            IndexDests dests, const __grid_constant__ P prm) {
    // fp64, dynamic: the same table-driven sincos as K1 (14 instead of 21 FP64 instr / angle)
    constexpr bool kTable = sizeof(T) == 8 && DYN;
    __shared__ SinCos tab[kTable ? kSinCosTableSize : 1];
    if constexpr (kTable) {
        for (int t = threadIdx.x; t < kSinCosTableSize; t += blockDim.x) tab[t] = kEdgeSinCosTable[t];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const T mp_inertial = TOOL ? T(0) : (mass > payload_threshold ? mass : T(0));
    const T mp_tool = TOOL ? mass : T(0);
    for (int64_t e = warp; e < n_edges; e += n_warps) {
        T a0[7], A[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            a0[j] = __ldg(qa + j * n_edges + e);
            // min_jerk_v2.py:121 with x = qa, v = a = 0, t = 1:  A = gx - x;  B = C = 0 (:110,:119)
            A[j] = __ldg(qb + j * n_edges + e) - a0[j];
        }
        int first = W;
        for (int base = 0; base < W; base += 32) {
            const int w = base + lane;
            const bool active = w < W;
            const T t = linspace_sample<T>(active ? w : W - 1, W, interval, step);
            T qs[7], vs[7], as[7], tau[7];
            // a3 = 10A, a4 = -15A, a5 = 6A (min_jerk_v2.py:128-130); x, v, a per :216-220
            const T t2 = t * t;
            const T px = t2 * t * (T(10) + t * (T(-15) + T(6) * t));
            const T pv = t2 * (T(30) + t * (T(-60) + T(30) * t));
            const T pa = t * (T(60) + t * (T(-180) + T(120) * t));
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                qs[j] = a0[j] + A[j] * px;
                if constexpr (DYN) {
                    vs[j] = A[j] * pv;
                    as[j] = A[j] * pa;
                }
            }
            if constexpr (kTable) rne_core_table<DYN, TOOL, P>(qs, vs, as, mp_inertial, mp_tool, tau, tab, prm);
            else rne_core<T, DYN, TOOL, P>(qs, vs, as, mp_inertial, mp_tool, tau, prm);
            const unsigned fails = __ballot_sync(0xffffffffu, active && !limits_ok<T, P>(tau, prm));
            if (fails) {
                first = base + __ffs(fails) - 1;
                break;
            }
        }
        if (lane == 0) {
            if (dests.n == 0) {
                first_fail[e] = first;
            } else {
#pragma unroll
                for (int d = 0; d < TCMP_MAX_PEERS; ++d)
                    if (d < dests.n) dests.p[d][dests.offset + e] = first;
            }
        }
    }
}

template <typename T, bool DYN, bool TOOL, typename P>
__global__ void __launch_bounds__(128)
traj_kernel(int n_seg, int S, double interval, double step, const double *__restrict__ coeffs, T mass,
            T payload_threshold, T *__restrict__ q_out, T *__restrict__ qd_out, T *__restrict__ qdd_out,
            T *__restrict__ tau_out, uint8_t *__restrict__ feasible_out, int32_t *__restrict__ first_fail,
            const __grid_constant__ P prm) {
    constexpr bool kTable = sizeof(T) == 8 && DYN;    // table-driven sincos, as in K1 and the edge kernel
    __shared__ SinCos tab[kTable ? kSinCosTableSize : 1];
    if constexpr (kTable) {
        for (int t = threadIdx.x; t < kSinCosTableSize; t += blockDim.x) tab[t] = kEdgeSinCosTable[t];
        __syncthreads();
    }
    const int64_t n = (int64_t)n_seg * S;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const T mp_inertial = TOOL ? T(0) : (mass > payload_threshold ? mass : T(0));
    const T mp_tool = TOOL ? mass : T(0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int seg = (int)(i / S), it = (int)(i - (int64_t)seg * S);
        const T t = linspace_sample<T>(it, S, interval, step);
        T qs[7], vs[7], as[7], tau[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const double *c = coeffs + ((int64_t)seg * 7 + j) * 6;
            const T c0 = (T)c[0], c1 = (T)c[1], c2 = (T)c[2], c3 = (T)c[3], c4 = (T)c[4], c5 = (T)c[5];
            qs[j] = c0 + t * (c1 + t * (c2 + t * (c3 + t * (c4 + t * c5))));                    // :216
            vs[j] = c1 + t * (T(2) * c2 + t * (T(3) * c3 + t * (T(4) * c4 + t * (T(5) * c5))));  // :218
            as[j] = T(2) * c2 + t * (T(6) * c3 + t * (T(12) * c4 + t * (T(20) * c5)));           // :220
        }
        if constexpr (kTable) rne_core_table<DYN, TOOL, P>(qs, vs, as, mp_inertial, mp_tool, tau, tab, prm);
        else rne_core<T, DYN, TOOL, P>(qs, vs, as, mp_inertial, mp_tool, tau, prm);
        const bool ok = limits_ok<T, P>(tau, prm);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            if (q_out) q_out[j * n + i] = qs[j];
            if (qd_out) qd_out[j * n + i] = vs[j];
            if (qdd_out) qdd_out[j * n + i] = as[j];
            if (tau_out) tau_out[j * n + i] = tau[j];
        }
        if (feasible_out) feasible_out[i] = (uint8_t)ok;
        if (first_fail && !ok) atomicMin(first_fail, (int32_t)i);
    }
}

static void linspace_params(int n, double *interval, double *step) {
    *interval = 1.0 / n;
    *step = n > 1 ? (1.0 - *interval) / (n - 1) : 0.0;
}

template <typename T, bool DYN, bool TOOL, typename P>
static cudaError_t launch_edge_p(int64_t n_edges, int W, const void *qa, const void *qb, double ps, double pt,
                                 int32_t *ff, const IndexDests &dests, const P &prm, cudaStream_t st) {
    double interval, step;
    linspace_params(W, &interval, &step);
    auto kern = edge_kernel<T, DYN, TOOL, P>;
    const int grid = grid_for(reinterpret_cast<const void *>(kern), 128, n_edges * 32, kEdgeWaves);
    kern<<<grid, 128, 0, st>>>(n_edges, W, interval, step, (const T *)qa, (const T *)qb, (T)ps, (T)pt, ff, dests, prm);
    return cudaGetLastError();
}

template <typename T, bool DYN, bool TOOL>
static cudaError_t launch_edge_t(int64_t n_edges, int W, const void *qa, const void *qb, double ps, double pt,
                                 int32_t *ff, const IndexDests &dests, const tcmp_model *model, cudaStream_t st) {
    if (model)
        return launch_edge_p<T, DYN, TOOL>(n_edges, W, qa, qb, ps, pt, ff, dests, params_from_desc<T>(*model), st);
    return launch_edge_p<T, DYN, TOOL>(n_edges, W, qa, qb, ps, pt, ff, dests, ConstParams(), st);
}

template <typename T>
static cudaError_t launch_edge_typed(int mode, int64_t n_edges, int W, const void *qa, const void *qb,
                                     double ps, double pt, int static_only, int32_t *ff, const IndexDests &dests,
                                     const tcmp_model *model, cudaStream_t st) {
    const bool dynamic = (mode != TCMP_MODE_NOV) && !static_only;
    const bool tool = (mode == TCMP_MODE_DYN);
    if (dynamic) {
        if (tool) return launch_edge_t<T, true, true>(n_edges, W, qa, qb, ps, pt, ff, dests, model, st);
        return launch_edge_t<T, true, false>(n_edges, W, qa, qb, ps, pt, ff, dests, model, st);
    }
    if (tool) return launch_edge_t<T, false, true>(n_edges, W, qa, qb, ps, pt, ff, dests, model, st);
    return launch_edge_t<T, false, false>(n_edges, W, qa, qb, ps, pt, ff, dests, model, st);
}

cudaError_t launch_edge_feasibility(int mode, int dtype, int64_t n_edges, int W, const void *qa, const void *qb,
                                    double ps, double pt, int static_only, int32_t *ff, cudaStream_t st,
                                    const tcmp_model *model) {
    IndexDests none;
    none.n = 0;
    none.offset = 0;
    for (int i = 0; i < TCMP_MAX_PEERS; ++i) none.p[i] = nullptr;
    if (mode == TCMP_MODE_BASE) return launch_fill<int32_t>(n_edges, ff, W, st);
    if (dtype == TCMP_F64)
        return launch_edge_typed<double>(mode, n_edges, W, qa, qb, ps, pt, static_only, ff, none, model, st);
    return launch_edge_typed<float>(mode, n_edges, W, qa, qb, ps, pt, static_only, ff, none, model, st);
}

cudaError_t launch_edge_feasibility_scatter(int mode, int64_t n_edges, int W, const void *qa, const void *qb,
                                            double ps, double pt, int static_only, int n_dest,
                                            void *const *dest_ff, int64_t dest_offset, cudaStream_t st) {
    IndexDests d;
    d.n = n_dest;
    d.offset = dest_offset;
    for (int i = 0; i < TCMP_MAX_PEERS; ++i) d.p[i] = i < n_dest ? (int32_t *)dest_ff[i] : nullptr;
    if (mode == TCMP_MODE_BASE) {
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < n_dest && e == cudaSuccess; ++i) e = launch_fill<int32_t>(n_edges, d.p[i] + dest_offset, W, st);
        return e;
    }
    return launch_edge_typed<double>(mode, n_edges, W, qa, qb, ps, pt, static_only, nullptr, d, nullptr, st);
}

template <typename T, bool DYN, bool TOOL, typename P>
static cudaError_t launch_traj_p(int n_seg, int S, const double *coeffs, double ps, double pt, void *q, void *qd,
                                 void *qdd, void *tau, uint8_t *mask, int32_t *ff, const P &prm, cudaStream_t st) {
    double interval, step;
    linspace_params(S, &interval, &step);
    auto kern = traj_kernel<T, DYN, TOOL, P>;
    const int grid = grid_for(reinterpret_cast<const void *>(kern), 128, (int64_t)n_seg * S);
    kern<<<grid, 128, 0, st>>>(n_seg, S, interval, step, coeffs, (T)ps, (T)pt, (T *)q, (T *)qd, (T *)qdd,
                               (T *)tau, mask, ff, prm);
    return cudaGetLastError();
}

template <typename T, bool DYN, bool TOOL>
static cudaError_t launch_traj_t(int n_seg, int S, const double *coeffs, double ps, double pt, void *q, void *qd,
                                 void *qdd, void *tau, uint8_t *mask, int32_t *ff, const tcmp_model *model,
                                 cudaStream_t st) {
    if (model)
        return launch_traj_p<T, DYN, TOOL>(n_seg, S, coeffs, ps, pt, q, qd, qdd, tau, mask, ff,
                                           params_from_desc<T>(*model), st);
    return launch_traj_p<T, DYN, TOOL>(n_seg, S, coeffs, ps, pt, q, qd, qdd, tau, mask, ff, ConstParams(), st);
}

template <typename T>
static cudaError_t launch_traj_typed(int mode, int n_seg, int S, const double *coeffs, double ps, double pt,
                                     void *q, void *qd, void *qdd, void *tau, uint8_t *mask, int32_t *ff,
                                     const tcmp_model *model, cudaStream_t st) {
    if (mode == TCMP_MODE_NOV)
        return launch_traj_t<T, false, false>(n_seg, S, coeffs, ps, pt, q, qd, qdd, tau, mask, ff, model, st);
    if (mode == TCMP_MODE_DYN)
        return launch_traj_t<T, true, true>(n_seg, S, coeffs, ps, pt, q, qd, qdd, tau, mask, ff, model, st);
    return launch_traj_t<T, true, false>(n_seg, S, coeffs, ps, pt, q, qd, qdd, tau, mask, ff, model, st);
}

cudaError_t launch_traj_feasibility(int mode, int dtype, int n_seg, int S, const double *coeffs, double ps,
                                    double pt, void *q, void *qd, void *qdd, void *tau, uint8_t *mask,
                                    int32_t *ff, cudaStream_t st, const tcmp_model *model) {
    if (mode == TCMP_MODE_BASE) {
        // Constant-true test (panda_primitives.py:13-16): the verdict is all-feasible and first_fail is left alone,
        // but the samples are still the trajectory the planner returns, and Conf logs rne torques for them whatever
        // the test mode (utils.py:3376-3378) -- so q / qd / qdd / tau are written by the rne kernel.
        if (q || qd || qdd || tau) {
            const cudaError_t e =
                dtype == TCMP_F64
                    ? launch_traj_typed<double>(TCMP_MODE_RNE, n_seg, S, coeffs, ps, pt, q, qd, qdd, tau, nullptr, nullptr,
                                                model, st)
                    : launch_traj_typed<float>(TCMP_MODE_RNE, n_seg, S, coeffs, ps, pt, q, qd, qdd, tau, nullptr, nullptr,
                                               model, st);
            if (e != cudaSuccess) return e;
        }
        return mask ? launch_fill<uint8_t>((int64_t)n_seg * S, mask, 1, st) : cudaSuccess;
    }
    if (dtype == TCMP_F64)
        return launch_traj_typed<double>(mode, n_seg, S, coeffs, ps, pt, q, qd, qdd, tau, mask, ff, model, st);
    return launch_traj_typed<float>(mode, n_seg, S, coeffs, ps, pt, q, qd, qdd, tau, mask, ff, model, st);
}

}  // namespace tcmp
