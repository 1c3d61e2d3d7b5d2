/*
 * oracle/ikfast_ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin plain-C driver around the UNMODIFIED reference IKFast solver.  It does not
 * copy any reference source: the Makefile compiles
 * /root/reference/src/ikfast_panda_arm.cpp where it lies (via -include of this
 * file's REF_IKFAST_CPP macro) into oracle/_ref/libikfast_ref.so, which is
 * git-ignored.  The reference .cpp unconditionally includes Python.h and defines
 * CPython wrappers (ikfast_panda_arm.cpp:25,12839-12993), so the .so links
 * libpython even though the entry points below never touch it.
 *
 * Entry points mirror the reference C API (ikfast_panda_arm.cpp:307 ComputeFk,
 * :12770 ComputeIk) and the list marshalling of get_ik (:12885-12902):
 * every solution is expanded with IkSolution::GetSolution (ikfast.h:167-181)
 * using zero-filled free values, exactly as get_ik does.
 */
#define IKFAST_NO_MAIN
#include REF_IKFAST_CPP

#include <cstdint>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

/* One solve; sols_out is [8][7]; returns the solution count (may exceed 8 in
 * principle -- only the first max_sols are written). */
int ref_ik_one(const double *eerot9, const double *eetrans3, double free_val,
               double *sols_out, int max_sols) {
    IkSolutionList<IkReal> solutions;
    IkReal pfree[1] = {free_val};
    bool ok = false;
    try {
        ok = ComputeIk(eetrans3, eerot9, pfree, solutions);
    } catch (const std::exception &) {
        return -1; /* IKFAST_ASSERT fired (ikfast_panda_arm.cpp:57) */
    }
    if (!ok) return 0;
    int n = (int)solutions.GetNumSolutions();
    for (int i = 0; i < n && i < max_sols; ++i) {
        const IkSolutionBase<IkReal> &sol = solutions.GetSolution(i);
        std::vector<IkReal> vsolfree(sol.GetFree().size());
        IkReal vals[7];
        sol.GetSolution(vals, vsolfree.size() > 0 ? &vsolfree[0] : NULL);
        for (int j = 0; j < 7; ++j) sols_out[i * 7 + j] = vals[j];
    }
    return n;
}

/* Batch: rot9 SoA [9][n], trans3 SoA [3][n], free [n_free][n] (or [n_free] when
 * free_broadcast != 0); outputs sols [n*n_free][8][7], counts [n*n_free], solve
 * index = pose * n_free + f.  OpenMP over solves. */
void ref_ik_batch(int64_t n, const double *rot9, const double *trans3, const double *free_vals,
                  int n_free, int free_broadcast, double *sols_out, int32_t *count_out,
                  int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (int64_t s = 0; s < n * (int64_t)n_free; ++s) {
        int64_t p = s / n_free;
        int f = (int)(s % n_free);
        double R[9], t[3], sols[8 * 7];
        for (int i = 0; i < 9; ++i) R[i] = rot9[i * n + p];
        for (int i = 0; i < 3; ++i) t[i] = trans3[i * n + p];
        double fv = free_broadcast ? free_vals[f] : free_vals[(int64_t)f * n + p];
        for (int i = 0; i < 56; ++i) sols[i] = 0.0;
        int c = ref_ik_one(R, t, fv, sols, 8);
        count_out[s] = c;
        if (sols_out)
            for (int i = 0; i < 56; ++i) sols_out[s * 56 + i] = sols[i];
    }
}

/* q SoA [7][n] -> trans3 SoA [3][n], rot9 SoA [9][n] (row-major rotation). */
void ref_fk_batch(int64_t n, const double *q, double *trans3, double *rot9) {
    for (int64_t s = 0; s < n; ++s) {
        double j[7], t[3], R[9];
        for (int i = 0; i < 7; ++i) j[i] = q[i * n + s];
        ComputeFk(j, t, R);
        for (int i = 0; i < 3; ++i) trans3[i * n + s] = t[i];
        for (int i = 0; i < 9; ++i) rot9[i * n + s] = R[i];
    }
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
}
