// ik_kernels.cu -- K5 (batched analytic IK, joint 7 free) and K6 (batched FK) for the Panda
// link0 -> link8 chain.
//
// Replaces ComputeIk / ComputeFk of the reference's generated solver (ikfast_panda_arm.cpp:307-395,
// :412-3114 IKSolver::ComputeIk, :3115-12593 rotationfunction0) and the solution expansion of
// get_ik (:12885-12902, ikfast.h:167-181).  This is a hand-structured restatement of the solver's
// decision tree, not a translation of the generated code:
//
//   wrist-centre reduction -> j4 (two roots of one asin) -> j6 (two roots, atan2 + asin)
//   -> j5 (one atan2) -> residual ZYZ rotation -> j2 (two roots of one acos) -> j1, j3 (atan2)
//
// (1-based joint names; the code below uses the solver's 0-based j0..j6.)  What IS kept bit-for-bit
// are the things that decide the SOLUTION COUNT: the branch thresholds (1e-7 sincos / atan2
// magnitude, 1e-6 branch and duplicate-root tests, 1e-5 residual checks, 5e-6 special-angle
// tests; ikfast_panda_arm.cpp:109-126), the clamped asin/acos (:136-142,169-175), the truncated
// pi literals used for angle wrapping (:67-69) and the solver's numeric coefficients of the
// Panda geometry.  One thread owns one (pose, free value) solve; up to 8 solutions.
#include "ik_core.cuh"
#include "tcmp_internal.h"

namespace tcmp {
namespace ik {

// Thread mapping is solve-major (s = pose * n_free + f, the ABI's output order): the 32 solves of a warp own one
// contiguous 32 x 448 B span of sols_out.  Each lane emits its solutions into a padded shared-memory row
// (57 doubles: an odd stride keeps 64-bit accesses conflict-free), then the warp copies the span out with fully
// coalesced 256 B stores -- instead of 56 scattered 8 B stores per lane that each dirty their own 32 B sector.
constexpr int kSolRow = 57;


// Compacting kernel, CTA-wide (round 2).  History: one lane per solve for its whole life ran with 14 of 32 lanes active
// (~45 % of a sweep's solves fail the solver's first gate and return at once while their warp-mates run all 8
// branches); per-WARP survivor queues brought that to 20 and 3.3 G solves/s, but every warp then sat at its own program
// counter in ~100 KB of inlined solver code and the kernel stalled on instruction fetch (ncu: 1.1 "no instruction"
// stalls per issue, 2.1 with solution sets).  Now ONE large CTA per SM: every iteration all lanes screen one solve each
// (pose reduction + the j3 gate), survivors are appended to ONE shared-memory queue (block-level prefix over the warps'
// ballots), and whenever the queue holds a CTA's worth of entries ALL warps solve at the same time -- they walk the
// solver's code together, so one instruction fetch serves them all: 4.3 G solves/s counts only (+28 %; 5.6 G with the
// class queues below).  Solution sets
// are staged in (dynamic) shared memory, one padded row per lane, and written row by row with coalesced stores.
__device__ __forceinline__ void load_pose(int64_t s, int64_t n, int n_free, int free_broadcast,
                                          const double *__restrict__ rot9, const double *__restrict__ trans3,
                                          const double *__restrict__ free_vals, Pose &P) {
    const int64_t p = s / n_free;
    const int f = (int)(s - p * n_free);
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = __ldg(rot9 + i * n + p);
    const double j6 = free_broadcast ? __ldg(free_vals + f) : __ldg(free_vals + (int64_t)f * n + p);
    prepare_pose(R, __ldg(trans3 + p), __ldg(trans3 + n + p), __ldg(trans3 + 2 * n + p), j6, P);
}

// The queue is split by WHICH j3 roots are live (ik_core.cuh plan_roots): both, only the first, only the second.  A
// batch is taken from one class, so the solver's two-root loop runs both iterations with every lane busy, or skips the
// same dead iteration on every lane (a warp-uniform branch) -- a mixed warp would execute both iterations with half of
// its one-root lanes idle in each.  ~48 % of the surviving solves have two live roots, ~52 % one.
#ifndef TCMP_IK_CLASSES
#define TCMP_IK_CLASSES 3     // 1 = single queue (screen on the j3 gate only)
#endif
template <int THREADS, bool WRITE_SOLS>
__global__ void __launch_bounds__(THREADS, 1)
ik_kernel_cta(int64_t n, int n_free, int free_broadcast, const double *__restrict__ rot9,
              const double *__restrict__ trans3, const double *__restrict__ free_vals,
              double *__restrict__ sols_out, int32_t *__restrict__ count_out, uint8_t *__restrict__ status_out) {
    constexpr int kWarps = THREADS / 32;
    constexpr int kClasses = TCMP_IK_CLASSES;
    extern __shared__ double rows[];                        // WRITE_SOLS: [THREADS][kSolRow]
    __shared__ long long queue[kClasses][2 * THREADS];      // per class: <= THREADS - 1 left over + THREADS pushed
    __shared__ int warp_cnt[kClasses][kWarps];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t total = n * n_free;
    const int64_t per_iter = (int64_t)gridDim.x * THREADS;
    const int64_t n_iter = (total + per_iter - 1) / per_iter;     // the same trip count for every thread (barriers inside)
    int qn[kClasses];                                             // CTA-uniform queue lengths
#pragma unroll
    for (int k = 0; k < kClasses; ++k) qn[k] = 0;
    int64_t it = 0;
    for (;;) {
        // ---- pick a full class, or (once every solve has been screened) any non-empty one; else screen more ----
        int cls = -1, take = 0;
#pragma unroll
        for (int k = 0; k < kClasses; ++k)
            if (cls < 0 && qn[k] >= THREADS) { cls = k; take = THREADS; }
        if (cls < 0 && it >= n_iter) {
#pragma unroll
            for (int k = 0; k < kClasses; ++k)
                if (cls < 0 && qn[k] > 0) { cls = k; take = qn[k]; }
            if (cls < 0) break;
        }
        if (cls >= 0) {
            // ---- solve `take` entries of class `cls`, one per thread: all warps enter the solver together ----
            long long *q = queue[0];
#pragma unroll
            for (int k = 1; k < kClasses; ++k)
                if (cls == k) q = queue[k];
            if (tid < take) {
                const long long sq = q[tid];
                Pose P;
                load_pose(sq, n, n_free, free_broadcast, rot9, trans3, free_vals, P);
                Emit out;
                out.sols = WRITE_SOLS ? rows + tid * kSolRow : nullptr;
                out.count = 0;
                out.status = 0;
                solve_one_t<false>(P, out);
                // elbow singularity: ik_redo_kernel finishes this solve with the complete tree (keeping that tree out
                // of this kernel keeps its registers down)
                const bool redo = (out.status & kStatusRedo) != 0;
                if (redo) out.count = 0;
                if (WRITE_SOLS)
                    for (int k = (out.count < 8 ? out.count : 8) * 7; k < 56; ++k) rows[tid * kSolRow + k] = 0.0;
                count_out[sq] = redo ? -1 : out.count;
                if (status_out) status_out[sq] = (uint8_t)out.status;
            }
            if (WRITE_SOLS) {
                __syncthreads();
                for (int r = wib; r < take; r += kWarps) {   // row r -> its solve's 448 B slot: two coalesced stores
                    double *dst = sols_out + q[r] * 56;
                    dst[lane] = rows[r * kSolRow + lane];
                    if (lane < 24) dst[32 + lane] = rows[r * kSolRow + 32 + lane];
                }
            }
            int rest = 0;
#pragma unroll
            for (int k = 0; k < kClasses; ++k)
                if (cls == k) { rest = qn[k] - take; qn[k] = rest; }
            long long carry = 0;
            __syncthreads();
            if (tid < rest) carry = q[take + tid];
            __syncthreads();
            if (tid < rest) q[tid] = carry;
            __syncthreads();
            continue;
        }
        // ---- screen one solve per thread ----
        const long long s = (it * gridDim.x + blockIdx.x) * THREADS + tid;
        ++it;
        int my = -1;          // class of this lane's solve, -1 = finished here (no live root) or out of range
        bool dead = false;
        if (s < total) {
            Pose P;
            load_pose(s, n, n_free, free_broadcast, rot9, trans3, free_vals, P);
            unsigned status = 0;
            if (kClasses == 1) {
                const int verdict = screen_pose(P);
                my = verdict == 1 ? 0 : -1;
                status = verdict == 2 ? kStatusInvalid : 0;
            } else {
                Emit out;
                out.sols = nullptr;
                out.count = 0;
                out.status = 0;
                RootPlan pl;
                const int n_live = plan_roots(P, pl, out);
                my = n_live == 2 ? 0 : (n_live == 1 ? (pl.live[0] ? 1 : 2) : -1);
                status = out.status;
            }
            dead = my < 0;
            if (dead) {
                count_out[s] = 0;
                if (status_out) status_out[s] = (uint8_t)status;
            }
        }
        unsigned mk[kClasses];
#pragma unroll
        for (int k = 0; k < kClasses; ++k) {
            mk[k] = __ballot_sync(0xffffffffu, my == k);
            if (lane == 0) warp_cnt[k][wib] = __popc(mk[k]);
        }
        if (WRITE_SOLS) {   // zero-fill the finished solves' slots: each warp its own 32 consecutive solves, coalesced
            unsigned m_dead = __ballot_sync(0xffffffffu, dead);
            const long long s0 = __shfl_sync(0xffffffffu, s, 0);   // the warp's first solve of this iteration
            while (m_dead) {
                const int r = __ffs(m_dead) - 1;
                m_dead &= m_dead - 1;
                double *dst = sols_out + (s0 + r) * 56;
                dst[lane] = 0.0;
                if (lane < 24) dst[32 + lane] = 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kClasses; ++k) {
            int before = 0, added = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const int cw = warp_cnt[k][w];
                if (w < wib) before += cw;
                added += cw;
            }
            if (my == k) queue[k][qn[k] + before + __popc(mk[k] & lt_mask)] = s;
            qn[k] += added;
        }
        __syncthreads();
    }
}

// Second pass of tcmp_ik_batch: the solves ik_kernel_cta marked with count = -1 (elbow singularity met on the hot
// path) are redone with the complete decision tree.  One lane per solve; the scan reads 4 B per solve (25 M solves:
// 100 MB, ~20 us), the flagged solves are rare outside hand-built singular sweeps.
template <bool WRITE_SOLS>
__global__ void __launch_bounds__(128)
ik_redo_kernel(int64_t n, int n_free, int free_broadcast, const double *__restrict__ rot9,
               const double *__restrict__ trans3, const double *__restrict__ free_vals,
               double *__restrict__ sols_out, int32_t *__restrict__ count_out, uint8_t *__restrict__ status_out) {
    const int64_t total = n * n_free;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += stride) {
        if (count_out[s] >= 0) continue;
        Pose P;
        load_pose(s, n, n_free, free_broadcast, rot9, trans3, free_vals, P);
        double sols[56];
        Emit out;
        out.sols = WRITE_SOLS ? sols : nullptr;
        out.count = 0;
        out.status = 0;
        solve_one_t<true>(P, out);
        if (WRITE_SOLS) {
            const int filled = (out.count < 8 ? out.count : 8) * 7;
            for (int k = 0; k < 56; ++k) sols_out[s * 56 + k] = k < filled ? sols[k] : 0.0;
        }
        count_out[s] = out.count;
        if (status_out) status_out[s] = (uint8_t)out.status;
    }
}

// FK (ComputeFk, :307-395): T = prod_k DH_k(q_k), rows k = 0..7 of the Panda modified-DH table.
__global__ void __launch_bounds__(256)
fk_kernel(int64_t n, const double *__restrict__ q, double *__restrict__ trans3, double *__restrict__ rot9) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // columns of the accumulated rotation and the accumulated origin
        double X[3] = {1, 0, 0}, Y[3] = {0, 1, 0}, Z[3] = {0, 0, 1}, p[3] = {0, 0, 0};
        // (a, d, alpha code) per DH row; the flange row 7 is a pure +0.107 z offset
        const double A[7] = {0, 0, 0, 0.0825, -0.0825, 0, 0.088};
        const double D[7] = {0.333, 0, 0.316, 0, 0.384, 0, 0};
        const int AL[7] = {0, -1, 1, 1, -1, 1, 1};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            double s, c;
            sincos(__ldg(q + k * n + i), &s, &c);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                // Rx(alpha): (Y, Z) -> (Z, -Y) for +pi/2, (-Z, Y) for -pi/2
                const double y = AL[k] == 0 ? Y[r] : (AL[k] > 0 ? Z[r] : -Z[r]);
                const double z = AL[k] == 0 ? Z[r] : (AL[k] > 0 ? -Y[r] : Y[r]);
                p[r] += A[k] * X[r] + D[k] * z;   // Tx(a) before the twist, Tz(d) along the new z
                const double x = X[r];
                X[r] = c * x + s * y;             // Rz(theta)
                Y[r] = c * y - s * x;
                Z[r] = z;
            }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            trans3[r * n + i] = p[r] + 0.107 * Z[r];
            rot9[(r * 3 + 0) * n + i] = X[r];
            rot9[(r * 3 + 1) * n + i] = Y[r];
            rot9[(r * 3 + 2) * n + i] = Z[r];
        }
    }
}

}  // namespace ik

// counts only: 512 threads (128 registers, one CTA per SM); with solution sets the padded rows need 57 doubles of shared
// memory per lane: 384 threads = 171 KB of the SM's 227 KB
constexpr int kIkThreadsCounts = 512;
constexpr int kIkThreadsSets = 384;

cudaError_t launch_ik_batch(int64_t n, const double *rot9, const double *trans3, const double *free_vals,
                            int n_free, int free_broadcast, double *sols_out, int32_t *count_out,
                            uint8_t *status_out, cudaStream_t st) {
    if (sols_out) {
        auto kern = ik::ik_kernel_cta<kIkThreadsSets, true>;
        const size_t smem = (size_t)kIkThreadsSets * ik::kSolRow * sizeof(double);
        static bool configured[64] = {};      // per device: opt in to > 48 KB of dynamic shared memory once
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 64 && !configured[dev]) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev] = true;
        }
        kern<<<sm_count(), kIkThreadsSets, smem, st>>>(n, n_free, free_broadcast, rot9, trans3, free_vals, sols_out,
                                                       count_out, status_out);
        const int grid2 = grid_for(reinterpret_cast<const void *>(ik::ik_redo_kernel<true>), 128, n * n_free);
        ik::ik_redo_kernel<true><<<grid2, 128, 0, st>>>(n, n_free, free_broadcast, rot9, trans3, free_vals, sols_out,
                                                        count_out, status_out);
    } else {
        ik::ik_kernel_cta<kIkThreadsCounts, false><<<sm_count(), kIkThreadsCounts, 0, st>>>(
            n, n_free, free_broadcast, rot9, trans3, free_vals, sols_out, count_out, status_out);
        const int grid2 = grid_for(reinterpret_cast<const void *>(ik::ik_redo_kernel<false>), 128, n * n_free);
        ik::ik_redo_kernel<false><<<grid2, 128, 0, st>>>(n, n_free, free_broadcast, rot9, trans3, free_vals, sols_out,
                                                         count_out, status_out);
    }
    return cudaGetLastError();
}

cudaError_t launch_fk_batch(int64_t n, const double *q, double *trans3, double *rot9, cudaStream_t st) {
    const int grid = grid_for(reinterpret_cast<const void *>(ik::fk_kernel), 256, n);
    ik::fk_kernel<<<grid, 256, 0, st>>>(n, q, trans3, rot9);
    return cudaGetLastError();
}

}  // namespace tcmp
