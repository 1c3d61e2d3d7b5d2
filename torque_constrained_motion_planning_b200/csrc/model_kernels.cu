// model_kernels.cu -- the batched torque test with a caller-supplied inertial set (tcmp_rne_batch_model).
//
// Replaces what a reference user does by editing rne.py's module-level tables (ms / cs / inertia_matrices,
// rne.py:102,119,138) before calling rne(): here the tables are an argument.  The host folds link8 and the hand
// onto link 7, regroups the chain into base parameters with the SAME constexpr routine the compiled-in Panda goes
// through (panda_model.cuh: regroup_model), and passes the result to the kernel by value (440 B of kernel
// parameters, read through the constant bank).  The recursion is rne_core<..., RtParams<T>>: identical structure
// to K1-K3, parameters from registers instead of immediates.  One thread per state, SoA [7][n], streaming loads
// and stores; FP64-pipe bound like K1 (no double buffering: this is the configurable path, not the headline one).
#include <math.h>

#include "panda_model.cuh"
#include "tcmp_internal.h"

namespace tcmp {

static __device__ const SinCos kModelSinCosTable[kSinCosTableSize] = {
#include "sincos_table.inc"
};

// 2 CTAs / SM (194 registers: ptxas keeps the 55 parameters in registers).  Asking for 3 like K1 (168 registers) makes
// it spill 184 B instead of re-reading the constant bank: 15.4 against 16.8 G states/s (profiles/r02/
// extras_variants.log); the table-driven sincos is worth +5 % here (16.0 -> 16.8).
#ifndef TCMP_MODEL_MIN_BLOCKS
#define TCMP_MODEL_MIN_BLOCKS 2
#endif
template <typename T, bool DYN, bool TOOL>
__global__ void __launch_bounds__(128, TCMP_MODEL_MIN_BLOCKS)
rne_model_kernel(int64_t n, const T *__restrict__ q, const T *__restrict__ qd, const T *__restrict__ qdd,
                 const T *__restrict__ payload_mass, T payload_scalar, T payload_threshold, T *__restrict__ tau_out,
                 uint8_t *__restrict__ feasible_out, const __grid_constant__ RtParams<T> P) {
    constexpr bool kTable = sizeof(T) == 8 && DYN;    // table-driven sincos, as in K1
    __shared__ SinCos tab[kTable ? kSinCosTableSize : 1];
    if constexpr (kTable) {
        for (int t = threadIdx.x; t < kSinCosTableSize; t += blockDim.x) tab[t] = kModelSinCosTable[t];
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T qs[7], vs[7], as[7], tau[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            qs[j] = __ldcs(q + j * n + i);
            if constexpr (DYN) {
                vs[j] = __ldcs(qd + j * n + i);
                as[j] = __ldcs(qdd + j * n + i);
            }
        }
        const T mass = payload_mass ? __ldcs(payload_mass + i) : payload_scalar;
        // same payload rule as K1-K3: rigid body iff mass > threshold (rne / nov), tool-point force (dyn)
        const T mp_inertial = TOOL ? T(0) : (mass > payload_threshold ? mass : T(0));
        const T mp_tool = TOOL ? mass : T(0);
        if constexpr (kTable) rne_core_table<DYN, TOOL, RtParams<T>>(qs, vs, as, mp_inertial, mp_tool, tau, tab, P);
        else rne_core<T, DYN, TOOL, RtParams<T>>(qs, vs, as, mp_inertial, mp_tool, tau, P);
        if (tau_out) {
#pragma unroll
            for (int j = 0; j < 7; ++j) __stcs(tau_out + j * n + i, tau[j]);
        }
        if (feasible_out) __stcs(feasible_out + i, (uint8_t)within_limits<T>(tau, P));
    }
}

void default_model_desc(tcmp_model *out) {
    for (int k = 0; k < 9; ++k) {
        out->mass[k] = kMass[k];
        for (int j = 0; j < 3; ++j) out->com[k][j] = kCom[k][j];
        for (int j = 0; j < 6; ++j) out->inertia[k][j] = kInertia[k][j];
    }
    out->payload_radius = kPayloadR;
    out->tool_z = kToolOffset;
    for (int i = 0; i < 7; ++i) out->torque_limit[i] = torque_limit(i);
}

// NULL when the record is usable, else what is wrong with it.
const char *model_desc_problem(const tcmp_model &d) {
    for (int k = 0; k < 9; ++k) {
        if (!(d.mass[k] >= 0.0) || !isfinite(d.mass[k])) return "mass must be finite and >= 0";
        for (int j = 0; j < 3; ++j)
            if (!isfinite(d.com[k][j])) return "com must be finite";
        for (int j = 0; j < 6; ++j)
            if (!isfinite(d.inertia[k][j])) return "inertia must be finite";
    }
    if (!isfinite(d.payload_radius) || !isfinite(d.tool_z)) return "payload_radius / tool_z must be finite";
    for (int i = 0; i < 7; ++i)
        if (!(d.torque_limit[i] > 0.0) || !isfinite(d.torque_limit[i])) return "torque_limit must be finite and > 0";
    return nullptr;
}

template <typename T, bool DYN, bool TOOL>
static cudaError_t launch_model_kernel(const tcmp_model &d, int64_t n, const void *q, const void *qd, const void *qdd,
                                       const void *pm, double ps, double pt, void *tau, uint8_t *mask,
                                       cudaStream_t st) {
    auto kern = rne_model_kernel<T, DYN, TOOL>;
    const int grid = grid_for(reinterpret_cast<const void *>(kern), 128, n);
    kern<<<grid, 128, 0, st>>>(n, (const T *)q, (const T *)qd, (const T *)qdd, (const T *)pm, (T)ps, (T)pt, (T *)tau,
                               mask, params_from_desc<T>(d));
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_model_typed(const tcmp_model &d, int mode, int64_t n, const void *q, const void *qd,
                                      const void *qdd, const void *pm, double ps, double pt, void *tau, uint8_t *mask,
                                      cudaStream_t st) {
    const bool dynamic = (mode != TCMP_MODE_NOV) && qd && qdd;   // panda_primitives.py:136-137,175-177
    const bool tool = (mode == TCMP_MODE_DYN);
    if (dynamic) {
        if (tool) return launch_model_kernel<T, true, true>(d, n, q, qd, qdd, pm, ps, pt, tau, mask, st);
        return launch_model_kernel<T, true, false>(d, n, q, qd, qdd, pm, ps, pt, tau, mask, st);
    }
    if (tool) return launch_model_kernel<T, false, true>(d, n, q, qd, qdd, pm, ps, pt, tau, mask, st);
    return launch_model_kernel<T, false, false>(d, n, q, qd, qdd, pm, ps, pt, tau, mask, st);
}

cudaError_t launch_rne_batch_model(const tcmp_model &d, int mode, int dtype, int64_t n, const void *q, const void *qd,
                                   const void *qdd, const void *pm, double ps, double pt, void *tau, uint8_t *mask,
                                   cudaStream_t st) {
    if (dtype == TCMP_F64) return launch_model_typed<double>(d, mode, n, q, qd, qdd, pm, ps, pt, tau, mask, st);
    return launch_model_typed<float>(d, mode, n, q, qd, qdd, pm, ps, pt, tau, mask, st);
}

}  // namespace tcmp
