// tests/native/ik_host.cpp -- TEST HARNESS: compiles the product's IK solver core (csrc/ik_core.cuh,
// the same source the CUDA kernel uses) for the HOST so the decision tree can be exercised against the
// compiled reference on a CPU-only box.  Never linked into libtcmp.so; not a fallback path.
#define _GNU_SOURCE 1
#include <cstdint>
#include "../../torque_constrained_motion_planning_b200/csrc/ik_core.cuh"

extern "C" void host_ik_batch(int64_t n, const double *rot9, const double *trans3, const double *free_vals,
                              int n_free, int free_broadcast, double *sols_out, int32_t *count_out,
                              uint8_t *status_out) {
    using namespace tcmp::ik;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p)
        for (int f = 0; f < n_free; ++f) {
            double R[9];
            for (int i = 0; i < 9; ++i) R[i] = rot9[i * n + p];
            Pose P;
            prepare_pose(R, trans3[p], trans3[n + p], trans3[2 * n + p],
                         free_broadcast ? free_vals[f] : free_vals[(int64_t)f * n + p], P);
            const int64_t o = p * n_free + f;
            Emit out;
            out.sols = sols_out ? sols_out + o * 56 : nullptr;
            out.count = 0;
            out.status = 0;
            solve_one(P, out);
            if (out.sols)
                for (int k = (out.count < 8 ? out.count : 8) * 7; k < 56; ++k) out.sols[k] = 0.0;
            count_out[o] = out.count;
            if (status_out) status_out[o] = (uint8_t)out.status;
        }
}
