// tests/native/rne_host.cpp -- TEST HARNESS: compiles the product's RNE recursion (csrc/panda_model.cuh, the same
// source the CUDA kernels instantiate) for the HOST, so the customised recursion, the compile-time regrouping and
// the run-time model folding can be checked against the oracle on a CPU-only box.  Never linked into libtcmp.so;
// not a fallback path.
#include <cstdint>

#include "../../torque_constrained_motion_planning_b200/csrc/panda_model.cuh"

using namespace tcmp;

template <bool DYN, bool TOOL, typename P>
static void run(int64_t n, const double *q, const double *qd, const double *qdd, const double *pm, double ps,
                double pt, double *tau_out, uint8_t *ok_out, const P &p) {
    for (int64_t i = 0; i < n; ++i) {
        double qs[7], vs[7] = {0}, as[7] = {0}, tau[7];
        for (int j = 0; j < 7; ++j) {
            qs[j] = q[j * n + i];
            if (DYN) { vs[j] = qd[j * n + i]; as[j] = qdd[j * n + i]; }
        }
        const double mass = pm ? pm[i] : ps;
        const double mi = TOOL ? 0.0 : (mass > pt ? mass : 0.0), mt = TOOL ? mass : 0.0;   // rne_kernels.cu's payload rule
        rne_core<double, DYN, TOOL, P>(qs, vs, as, mi, mt, tau, p);
        for (int j = 0; j < 7; ++j) tau_out[j * n + i] = tau[j];
        if constexpr (kIsConst<P>) ok_out[i] = within_limits<double>(tau);
        else ok_out[i] = within_limits<double>(tau, p);
    }
}

template <typename P>
static void dispatch(int mode, int64_t n, const double *q, const double *qd, const double *qdd, const double *pm,
                     double ps, double pt, double *tau, uint8_t *ok, const P &p) {
    const bool dyn = mode != 1 && qd && qdd, tool = mode == 2;
    if (dyn) {
        if (tool) run<true, true>(n, q, qd, qdd, pm, ps, pt, tau, ok, p);
        else run<true, false>(n, q, qd, qdd, pm, ps, pt, tau, ok, p);
    } else {
        if (tool) run<false, true>(n, q, qd, qdd, pm, ps, pt, tau, ok, p);
        else run<false, false>(n, q, qd, qdd, pm, ps, pt, tau, ok, p);
    }
}

// model == NULL: the compiled-in Panda (ConstParams), else the record folded and regrouped at run time.
extern "C" void host_rne_batch(const tcmp_model *model, int mode, int64_t n, const double *q, const double *qd,
                               const double *qdd, const double *pm, double ps, double pt, double *tau, uint8_t *ok) {
    if (model) dispatch(mode, n, q, qd, qdd, pm, ps, pt, tau, ok, params_from_desc<double>(*model));
    else dispatch(mode, n, q, qd, qdd, pm, ps, pt, tau, ok, ConstParams());
}

// ---- K1's table-driven sincos --------------------------------------------------------------------------------
static const SinCos kTable[kSinCosTableSize] = {
#include "../../torque_constrained_motion_planning_b200/csrc/sincos_table.inc"
};

// sin / cos of n angles through the table path (angle j of a state = x[i], the other five zero); ok_out[i] = 0
// where the fast path declined (|x| >= 4096 or non-finite).
extern "C" void host_sincos_table(int64_t n, const double *x, double *s_out, double *c_out, uint8_t *ok_out) {
    for (int64_t i = 0; i < n; ++i) {
        double q[7] = {0, x[i], 0, 0, 0, 0, 0}, s[7], c[7];
        ok_out[i] = sincos6_table(q, s, c, kTable);
        s_out[i] = ok_out[i] ? s[1] : 0.0;
        c_out[i] = ok_out[i] ? c[1] : 0.0;
    }
}

// The compiled-in Panda through rne_core_table (what K1 instantiates), dynamic rne mode.
extern "C" void host_rne_batch_table(int64_t n, const double *q, const double *qd, const double *qdd, const double *pm,
                                     double pt, double *tau, uint8_t *ok) {
    for (int64_t i = 0; i < n; ++i) {
        double qs[7], vs[7], as[7], t[7];
        for (int j = 0; j < 7; ++j) { qs[j] = q[j * n + i]; vs[j] = qd[j * n + i]; as[j] = qdd[j * n + i]; }
        rne_core_table<true, false>(qs, vs, as, pm[i] > pt ? pm[i] : 0.0, 0.0, t, kTable);
        for (int j = 0; j < 7; ++j) tau[j * n + i] = t[j];
        ok[i] = within_limits<double>(t);
    }
}
