// rne_kernels.cu -- K1 (rne), K2 (nov), K3 (dyn): batched torque test, one thread per state.
//
// Replaces, per state: rne.rne (rne.py:198-254) + add_payload/remove_payload (rne.py:181-195)
// + the limit compare of the torque-test closures (panda_primitives.py:182-188).
//
// Data layout: structure-of-arrays [7][n] so that the 32 lanes of a warp read 32 consecutive
// elements of each joint row (one 256 B request per row per warp, fully coalesced).  Loads are
// streaming (ld.global.cs) and stores are st.global.cs: every byte is touched exactly once.
// Bound: FP64 pipe, with HBM close behind: 233 B and 669 FP64 instructions per state (rne).
#include "panda_model.cuh"
#include "tcmp_internal.h"

namespace tcmp {

// SCATTER: the fused "compute + all-gather" form for multi-GPU runs.  Instead of writing its mask shard
// locally and letting a separate NCCL all-gather move it, every thread stores its mask byte straight into
// the gathered buffer of EVERY rank (peer pointers mapped over NVLink/NVSwitch, CUDA IPC) at
// dest_offset + i.  1 B/state/peer of NVLink traffic rides under an FP64-bound kernel; no extra launch,
// no host-side collective call (which costs more CPU time than this ~45 us kernel runs).
#ifndef TCMP_PDL
#define TCMP_PDL 1
#endif
#ifndef TCMP_INT_INDEX
#define TCMP_INT_INDEX 1
#endif
#ifndef TCMP_RNE_WAVES
#define TCMP_RNE_WAVES 4
#endif
#ifndef TCMP_DOUBLE_BUFFER
#define TCMP_DOUBLE_BUFFER 1   // measured +3..5 % with 3 resident CTAs (168 registers, no spills)
#endif
#if TCMP_DOUBLE_BUFFER && !defined(TCMP_RNE_MIN_BLOCKS)
#define TCMP_RNE_MIN_BLOCKS 3
#endif

#ifndef TCMP_TABLE_SINCOS
#define TCMP_TABLE_SINCOS 1   // fp64 K1/K3: table-driven sincos (panda_model.cuh sincos6_table), 54 fewer FP64 instr / state: +6.7 %
#endif
// (sin, cos)(i pi / 512): read-only, identical for every launch; each CTA stages it in shared memory (16 KB).
__device__ const SinCos kSinCosTable[kSinCosTableSize] = {
#include "sincos_table.inc"
};

struct MaskDests {
    uint8_t *p[TCMP_MAX_PEERS];
    int n;
    int64_t offset;
    uint8_t *mc;       // NVSwitch multicast address of the same gathered buffer (or nullptr): one store reaches every rank
};

// sync blocks of all ranks as one kernel argument (tcmp_peer_signal)
struct SyncDests {
    PeerSync *sync[TCMP_MAX_PEERS];
    int n_sync, rank;
};

// Scatter epilogue: the 32 lanes of a warp store 32 consecutive mask bytes into every rank's gathered buffer.
//  * unicast (peer pointers): one byte store per lane per peer -- one 32-byte sector request per peer per warp round.
//    (Packing the verdicts with a ballot into 16-byte vector stores per peer was measured at N = 8: 54.1 us per step
//    against 52.7 us for the byte stores -- the hardware already coalesces the byte lanes.)
//  * multicast (dests.mc, NVLS): the buffer is also mapped at an NVSwitch multicast address; the warp ballots its 32
//    verdicts, lanes 0 and 1 each issue ONE 16-byte multimem.st and the switch replicates it to every rank -- 1/8 of
//    the store requests and of the outgoing NVLink bytes at N = 8.  multimem.st has no byte form, so a partial warp
//    (the batch's tail) or a misaligned destination takes the unicast path.  Measured at N = 8: 53.7 us per step
//    against 54.1 us unicast (plain kernel 48.2-50.7): what the gather costs is receiving 8 MB of small writes per
//    rank per step and draining the last CTAs' remote stores, not issuing them (profiles/r02/scatter_signal_variants.log),
//    so the unicast form stays the default and this one is opt-in (PeerMaskBuffer(multicast=True)).
template <typename I>
__device__ __forceinline__ void scatter_mask(const MaskDests &dests, I at, I n, bool ok) {
    if (dests.mc) {
        const int lane = threadIdx.x & 31;
        const I base = at - (I)lane;
        if (base + 32 <= n) {                       // warp-uniform: all 32 lanes are in this round
            const unsigned b = __ballot_sync(0xffffffffu, ok);
            if (lane < 2) {
                const unsigned h16 = (b >> (lane * 16)) & 0xffffu;
                // 4 verdict bits -> 4 bytes 0 / 1
                auto spread = [](unsigned x) { return (x & 1u) | ((x & 2u) << 7) | ((x & 4u) << 14) | ((x & 8u) << 21); };
                const unsigned w0 = spread(h16 & 0xfu), w1 = spread((h16 >> 4) & 0xfu),
                               w2 = spread((h16 >> 8) & 0xfu), w3 = spread((h16 >> 12) & 0xfu);
                uint8_t *dst = dests.mc + dests.offset + base + lane * 16;
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst),
                             "f"(__uint_as_float(w0)), "f"(__uint_as_float(w1)), "f"(__uint_as_float(w2)),
                             "f"(__uint_as_float(w3))
                             : "memory");
            }
            return;
        }
    }
    const uint8_t m = (uint8_t)ok;
#pragma unroll
    for (int d = 0; d < TCMP_MAX_PEERS; ++d)   // unrolled: constant indices keep `dests` in param space
        if (d < dests.n) dests.p[d][dests.offset + at] = m;
}

// Launch bounds: 128-thread CTAs.  Single-buffered, ptxas settles on 128 registers (4 CTAs / SM, no spills); the
// double-buffered default asks for 3 CTAs / SM (168 registers, no spills).  Capping
// lower (TCMP_RNE_MIN_BLOCKS=5) trades spills for occupancy -- measured slower, see profiles/.
#ifndef TCMP_RNE_BLOCK
#define TCMP_RNE_BLOCK 128
#endif
#ifdef TCMP_RNE_MIN_BLOCKS
#define TCMP_RNE_BOUNDS TCMP_RNE_BLOCK, TCMP_RNE_MIN_BLOCKS
#else
#define TCMP_RNE_BOUNDS TCMP_RNE_BLOCK
#endif
// I: element-index type.  int when every offset j*n + i (and the grid-stride overshoot) fits 31 bits -- one
// IMAD.WIDE per address instead of 64-bit multiply-add chains -- else int64_t.
template <typename T, typename I, bool DYN, bool TOOL, bool WRITE_TAU, bool WRITE_MASK, bool SCATTER = false>
__global__ void __launch_bounds__(TCMP_RNE_BOUNDS)
rne_batch_kernel(I n, const T *__restrict__ q, const T *__restrict__ qd, const T *__restrict__ qdd,
                 const T *__restrict__ payload_mass, T payload_scalar, T payload_threshold,
                 T *__restrict__ tau_out, uint8_t *__restrict__ feasible_out, const __grid_constant__ MaskDests dests) {
    // dynamic fp64 kernels only: the static (nov) kernel is HBM-leaning and measured 1.4 % slower with the staging
    constexpr bool kTable = TCMP_TABLE_SINCOS && sizeof(T) == 8 && DYN;
    __shared__ SinCos tab[kTable ? kSinCosTableSize : 1];
    if constexpr (kTable) {
        // constant data, no dependence on earlier grids: staged BEFORE the programmatic-launch wait below, so it
        // overlaps the previous launch's tail
        // (cp.async staging overlapped with the first state's loads was measured: no change, 22.40 vs 22.45 G states/s)
        for (int t = threadIdx.x; t < kSinCosTableSize; t += blockDim.x) tab[t] = kSinCosTable[t];
        __syncthreads();
    }
#if TCMP_PDL
    // Programmatic dependent launch: let the next launch on this stream become resident while this grid's
    // last wave drains (its CTAs then sit in griddepcontrol.wait), so back-to-back batches do not pay the
    // launch + ramp-up gap (~5 % of the kernel).  Stream-order semantics are kept: nothing is read or
    // written before the wait, which returns only when every earlier grid has completed and flushed.
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
#if TCMP_DOUBLE_BUFFER
    // Register double buffering: the NEXT grid-stride state's 22 inputs are loaded before the current state's
    // recursion starts, so a warp never sits in a long-scoreboard stall with nothing to issue (ncu: 1.06 -> 0.14
    // long-scoreboard stalls per issue).  (Ping-ponging two register sets instead of copying next -> current spills
    // and is 16 % slower; prefetch.global of the next rows costs 76 registers and 9 % -- see DESIGN.md 6b.  Round 2:
    // staging the next state in shared memory with cp.async instead -- 125 registers, 4 CTAs / SM, no spills -- is 3 %
    // SLOWER (22.1 vs 22.8 G states/s burst): more resident warps do not help a kernel whose sustained rate is set by
    // the 1 kW power cap, profiles/r02/k1_variants.log.)
    const I stride = (I)(gridDim.x * blockDim.x);
    I i = (I)(blockIdx.x * blockDim.x + threadIdx.x);
    const bool any = i < n;
    auto load = [&](I at, T (&lq)[7], T (&lv)[7], T (&la)[7], T &lm) {
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            lq[j] = __ldcs(q + j * n + at);
            if constexpr (DYN) { lv[j] = __ldcs(qd + j * n + at); la[j] = __ldcs(qdd + j * n + at); }
        }
        lm = payload_mass ? __ldcs(payload_mass + at) : payload_scalar;
    };
    auto consume = [&](I at, const T (&lq)[7], const T (&lv)[7], const T (&la)[7], T lm) {
        T tau[7];
        const T mp_inertial = TOOL ? T(0) : (lm > payload_threshold ? lm : T(0));
        const T mp_tool = TOOL ? lm : T(0);
        if constexpr (kTable) rne_core_table<DYN, TOOL>(lq, lv, la, mp_inertial, mp_tool, tau, tab);
        else rne_core<T, DYN, TOOL>(lq, lv, la, mp_inertial, mp_tool, tau);
        if constexpr (WRITE_TAU) {
#pragma unroll
            for (int j = 0; j < 7; ++j) __stcs(tau_out + j * n + at, tau[j]);
        }
        if constexpr (SCATTER) {
            scatter_mask<I>(dests, at, n, within_limits<T>(tau));
        } else if constexpr (WRITE_MASK) {
            __stcs(feasible_out + at, (uint8_t)within_limits<T>(tau));
        }
    };
    if (any) {
        T qa[7], va[7], aa[7], ma, qb[7], vb[7], ab[7], mb;
        load(i, qa, va, aa, ma);
        for (;;) {
            const I nx = i + stride;
            const bool more = nx < n;
            if (more) load(nx, qb, vb, ab, mb);
            consume(i, qa, va, aa, ma);
            if (!more) break;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                qa[j] = qb[j];
                if constexpr (DYN) { va[j] = vb[j]; aa[j] = ab[j]; }
            }
            ma = mb;
            i = nx;
        }
    }
}
#else
    const I stride = (I)(gridDim.x * blockDim.x);
    for (I i = (I)(blockIdx.x * blockDim.x + threadIdx.x); i < n; i += stride) {
        T qs[7], vs[7], as[7], tau[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) qs[j] = __ldcs(q + j * n + i);
        if constexpr (DYN) {
#pragma unroll
            for (int j = 0; j < 7; ++j) vs[j] = __ldcs(qd + j * n + i);
#pragma unroll
            for (int j = 0; j < 7; ++j) as[j] = __ldcs(qdd + j * n + i);
        }
        const T mass = payload_mass ? __ldcs(payload_mass + i) : payload_scalar;
        // rne / nov: rigid payload iff mass > threshold (panda_primitives.py:139,178; rne.py:184)
        // dyn: the mass enters only as a tool-point gravity force (panda_primitives.py:101-111)
        const T mp_inertial = TOOL ? T(0) : (mass > payload_threshold ? mass : T(0));
        const T mp_tool = TOOL ? mass : T(0);
        if constexpr (kTable) rne_core_table<DYN, TOOL>(qs, vs, as, mp_inertial, mp_tool, tau, tab);
        else rne_core<T, DYN, TOOL>(qs, vs, as, mp_inertial, mp_tool, tau);
        if constexpr (WRITE_TAU) {
#pragma unroll
            for (int j = 0; j < 7; ++j) __stcs(tau_out + j * n + i, tau[j]);
        }
        if constexpr (SCATTER) {
            scatter_mask<I>(dests, i, n, within_limits<T>(tau));
        } else if constexpr (WRITE_MASK) {
            __stcs(feasible_out + i, (uint8_t)within_limits<T>(tau));
        }
    }
}
#endif

// 31-bit offsets (7 rows of n elements plus the grid-stride overshoot of at most one grid) -> int indices
static inline bool fits_int(int64_t n) { return n < (int64_t)(0x7fffffff / 7) - 148 * 32 * 2048; }

template <typename T, typename I, bool DYN, bool TOOL, bool WT, bool WM, bool SC>
static cudaError_t launch_kernel(int64_t n, const void *q, const void *qd, const void *qdd, const void *pm, double ps,
                                 double pt, void *tau, uint8_t *mask, const MaskDests &dests, cudaStream_t st) {
    auto kern = rne_batch_kernel<T, I, DYN, TOOL, WT, WM, SC>;
    // TCMP_RNE_WAVES x the resident CTA count: the warp scheduler lets some warps of an SM run ahead of others, so
    // with exactly one wave the last stretch of a launch runs with vacated warp slots (ncu: 9.5 of 12 warps active
    // on average).  Shorter CTAs are replaced as they retire; 4 waves measured best (+3.7 % rne, +6 % nov; 8 and 16
    // lose the double buffer's benefit to the unprefetched first state of every CTA).
    const int64_t want = (n + TCMP_RNE_BLOCK - 1) / TCMP_RNE_BLOCK;
    const int64_t cap = (int64_t)grid_for(reinterpret_cast<const void *>(kern), TCMP_RNE_BLOCK, n) * TCMP_RNE_WAVES;
    const int grid = (int)(want < cap ? want : cap);
#if TCMP_PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TCMP_RNE_BLOCK);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, (I)n, (const T *)q, (const T *)qd, (const T *)qdd, (const T *)pm, (T)ps,
                              (T)pt, (T *)tau, mask, dests);
#else
    kern<<<grid, TCMP_RNE_BLOCK, 0, st>>>((I)n, (const T *)q, (const T *)qd, (const T *)qdd, (const T *)pm, (T)ps,
                                          (T)pt, (T *)tau, mask, dests);
    return cudaGetLastError();
#endif
}

template <typename T, bool DYN, bool TOOL, bool WT, bool WM, bool SC>
static cudaError_t launch_indexed(int64_t n, const void *q, const void *qd, const void *qdd, const void *pm, double ps,
                                  double pt, void *tau, uint8_t *mask, const MaskDests &dests, cudaStream_t st) {
#if TCMP_INT_INDEX
    if (fits_int(n)) return launch_kernel<T, int, DYN, TOOL, WT, WM, SC>(n, q, qd, qdd, pm, ps, pt, tau, mask, dests, st);
#endif
    return launch_kernel<T, int64_t, DYN, TOOL, WT, WM, SC>(n, q, qd, qdd, pm, ps, pt, tau, mask, dests, st);
}

template <typename T, bool DYN, bool TOOL, bool WT, bool WM>
static cudaError_t launch_one(int64_t n, const void *q, const void *qd, const void *qdd, const void *pm,
                              double ps, double pt, void *tau, uint8_t *mask, cudaStream_t st) {
    MaskDests none = {};
    return launch_indexed<T, DYN, TOOL, WT, WM, false>(n, q, qd, qdd, pm, ps, pt, tau, mask, none, st);
}

template <typename T, bool DYN, bool TOOL>
static cudaError_t launch_outputs(int64_t n, const void *q, const void *qd, const void *qdd, const void *pm,
                                  double ps, double pt, void *tau, uint8_t *mask, cudaStream_t st) {
    if (tau && mask) return launch_one<T, DYN, TOOL, true, true>(n, q, qd, qdd, pm, ps, pt, tau, mask, st);
    if (tau) return launch_one<T, DYN, TOOL, true, false>(n, q, qd, qdd, pm, ps, pt, tau, mask, st);
    return launch_one<T, DYN, TOOL, false, true>(n, q, qd, qdd, pm, ps, pt, tau, mask, st);
}

template <typename T>
static cudaError_t launch_typed(int mode, int64_t n, const void *q, const void *qd, const void *qdd,
                                const void *pm, double ps, double pt, void *tau, uint8_t *mask,
                                cudaStream_t st) {
    // Without both velocity arrays the reference tests the static state (panda_primitives.py:175-177);
    // nov always does (panda_primitives.py:136-137).
    const bool dynamic = (mode != TCMP_MODE_NOV) && qd && qdd;
    const bool tool = (mode == TCMP_MODE_DYN);
    if (dynamic) {
        if (tool) return launch_outputs<T, true, true>(n, q, qd, qdd, pm, ps, pt, tau, mask, st);
        return launch_outputs<T, true, false>(n, q, qd, qdd, pm, ps, pt, tau, mask, st);
    }
    if (tool) return launch_outputs<T, false, true>(n, q, qd, qdd, pm, ps, pt, tau, mask, st);
    return launch_outputs<T, false, false>(n, q, qd, qdd, pm, ps, pt, tau, mask, st);
}

cudaError_t launch_rne_batch(int mode, int dtype, int64_t n, const void *q, const void *qd, const void *qdd,
                             const void *pm, double ps, double pt, void *tau, uint8_t *mask,
                             cudaStream_t st) {
    if (mode == TCMP_MODE_BASE) {  // panda_primitives.py:13-16: always feasible, no torques computed
        cudaError_t e = cudaSuccess;
        if (mask) e = launch_fill<uint8_t>(n, mask, 1, st);
        if (tau && e == cudaSuccess)
            e = dtype == TCMP_F64 ? launch_fill<double>(7 * n, (double *)tau, 0.0, st)
                                  : launch_fill<float>(7 * n, (float *)tau, 0.f, st);
        return e;
    }
    if (dtype == TCMP_F64) return launch_typed<double>(mode, n, q, qd, qdd, pm, ps, pt, tau, mask, st);
    return launch_typed<float>(mode, n, q, qd, qdd, pm, ps, pt, tau, mask, st);
}

template <typename T, bool DYN, bool TOOL>
static cudaError_t launch_scatter_t(int64_t n, const void *q, const void *qd, const void *qdd, const void *pm,
                                    double ps, double pt, void *tau, const MaskDests &dests, cudaStream_t st) {
    if (tau) return launch_indexed<T, DYN, TOOL, true, true, true>(n, q, qd, qdd, pm, ps, pt, tau, nullptr, dests, st);
    return launch_indexed<T, DYN, TOOL, false, true, true>(n, q, qd, qdd, pm, ps, pt, tau, nullptr, dests, st);
}

// ---- completion flags of the fused gathers -----------------------------------------------------------------------
// Publish: enqueued after a scatter kernel.  Kernel completion has drained that grid's peer stores, so one thread
// fences and stores this rank's new epoch into the `arrived[rank]` word of every rank's sync block.  (Publishing from
// inside the scatter kernel -- a system-scope fence + ticket at the end of each CTA, flags written by the CTA that
// draws the last ticket -- was built and measured: 21-27 us per launch at 4 grid waves, 6-8 us at one, because every
// retiring CTA then waits for its stores to be acknowledged by a saturated memory system; this kernel costs 3 us
// in-stream and nothing on a side stream.  profiles/r02/scatter_signal_variants.log)
__global__ void peer_signal_kernel(SyncDests d) {
    PeerSync *own = d.sync[d.rank];
    const unsigned long long e = own->epoch + 1;
    own->epoch = e;
    __threadfence_system();
    for (int r = 0; r < d.n_sync; ++r) *reinterpret_cast<volatile unsigned long long *>(&d.sync[r]->arrived[d.rank]) = e;
}

// Consumer side: lane r waits until rank r has published an epoch >= this rank's own (every rank runs the same
// sequence of scatter steps, so epochs line up); acquire loads at system scope, back-off while spinning.
// A peer that died never publishes: after ~20 s of polling the lane gives up and records it in the sync block
// (pad[0] = 1 + the rank it waited for; PeerMaskBuffer.check() raises on it) instead of hanging the GPU.
__global__ void peer_wait_kernel(PeerSync *own, int world) {
    const unsigned long long e = own->epoch;
    if ((int)threadIdx.x < world) {
        const unsigned long long *flag = &own->arrived[threadIdx.x];
        unsigned long long v;
        const long long t0 = clock64();
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if (v >= e) break;
            if (clock64() - t0 > 40000000000ll) {          // ~20 s at 2 GHz
                own->pad[0] = 1u + threadIdx.x;
                break;
            }
            __nanosleep(200);
        }
    }
}

// Stand-alone gather step: copy `bytes` of this rank's result block into every rank's gathered buffer at dest_offset
// (a few CTAs, 16-byte loads and stores when everything is 16-byte aligned).  Run on the side stream behind the
// producing kernel it rides under the NEXT step's kernel -- the alternative to the fused peer-store epilogue.
__global__ void __launch_bounds__(256)
peer_push_kernel(const uint8_t *__restrict__ src, int64_t bytes, MaskDests d, int vec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const int64_t nv = bytes / 16;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        for (int64_t i = tid; i < nv; i += stride) {
            const uint4 v = __ldg(s4 + i);
#pragma unroll
            for (int r = 0; r < TCMP_MAX_PEERS; ++r)
                if (r < d.n) reinterpret_cast<uint4 *>(d.p[r] + d.offset)[i] = v;
        }
        for (int64_t i = nv * 16 + tid; i < bytes; i += stride) {
            const uint8_t v = src[i];
#pragma unroll
            for (int r = 0; r < TCMP_MAX_PEERS; ++r)
                if (r < d.n) d.p[r][d.offset + i] = v;
        }
    } else {
        for (int64_t i = tid; i < bytes; i += stride) {
            const uint8_t v = src[i];
#pragma unroll
            for (int r = 0; r < TCMP_MAX_PEERS; ++r)
                if (r < d.n) d.p[r][d.offset + i] = v;
        }
    }
}

cudaError_t launch_peer_push(const void *src, int64_t bytes, int n_dest, void *const *dests, int64_t dest_offset,
                             cudaStream_t st) {
    MaskDests d = {};
    d.n = n_dest;
    d.offset = dest_offset;
    int vec = ((uintptr_t)src % 16 == 0);
    for (int i = 0; i < n_dest; ++i) {
        d.p[i] = (uint8_t *)dests[i];
        if (((uintptr_t)dests[i] + (uintptr_t)dest_offset) % 16 != 0) vec = 0;
    }
    const int64_t units = vec ? bytes / 16 : bytes;
    int grid = (int)((units + 255) / 256);
    if (grid > 32) grid = 32;     // a side-stream copy: leave the SMs to the torque kernel it overlaps
    if (grid < 1) grid = 1;
    peer_push_kernel<<<grid, 256, 0, st>>>((const uint8_t *)src, bytes, d, vec);
    return cudaGetLastError();
}

cudaError_t launch_peer_signal(int rank, int n_dest, void *const *dest_sync, cudaStream_t st) {
    SyncDests d = {};
    d.n_sync = n_dest;
    d.rank = rank;
    for (int i = 0; i < n_dest; ++i) d.sync[i] = (PeerSync *)dest_sync[i];
    peer_signal_kernel<<<1, 1, 0, st>>>(d);
    return cudaGetLastError();
}

cudaError_t launch_peer_wait(void *own_sync, int world, cudaStream_t st) {
    peer_wait_kernel<<<1, 32, 0, st>>>((PeerSync *)own_sync, world);
    return cudaGetLastError();
}

cudaError_t launch_rne_batch_scatter(int mode, int dtype, int64_t n, const void *q, const void *qd, const void *qdd,
                                     const void *pm, double ps, double pt, void *tau, int n_dest,
                                     void *const *dest_masks, int64_t dest_offset, cudaStream_t st, void *mc_masks) {
    MaskDests d = {};
    d.n = n_dest;
    d.offset = dest_offset;
    for (int i = 0; i < TCMP_MAX_PEERS; ++i) d.p[i] = i < n_dest ? (uint8_t *)dest_masks[i] : nullptr;
    // the 16-byte multicast stores need a 16-byte aligned destination (the warp's base index is a multiple of 32)
    d.mc = (mc_masks && ((uintptr_t)mc_masks + (uintptr_t)dest_offset) % 16 == 0) ? (uint8_t *)mc_masks : nullptr;
    if (dtype != TCMP_F64) return cudaErrorNotSupported;
    const bool dynamic = (mode != TCMP_MODE_NOV) && qd && qdd;
    const bool tool = (mode == TCMP_MODE_DYN);
    if (mode == TCMP_MODE_BASE) {
        cudaError_t e = cudaSuccess;
        for (int i = 0; i < n_dest && e == cudaSuccess; ++i) e = launch_fill<uint8_t>(n, d.p[i] + dest_offset, 1, st);
        return e;
    }
    if (dynamic) {
        if (tool) return launch_scatter_t<double, true, true>(n, q, qd, qdd, pm, ps, pt, tau, d, st);
        return launch_scatter_t<double, true, false>(n, q, qd, qdd, pm, ps, pt, tau, d, st);
    }
    if (tool) return launch_scatter_t<double, false, true>(n, q, qd, qdd, pm, ps, pt, tau, d, st);
    return launch_scatter_t<double, false, false>(n, q, qd, qdd, pm, ps, pt, tau, d, st);
}

}  // namespace tcmp
