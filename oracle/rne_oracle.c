/*
 * oracle/rne_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU, fp64 restatement of the reference's torque-feasibility path,
 * used only as the checker in tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  Nothing under
 * torque_constrained_motion_planning_b200/ may call, link or import it.
 *
 * It follows the reference's arithmetic literally (dense 4x4 homogeneous
 * transforms, dense 6x6 spatial matrices, same evaluation order) rather than
 * an optimised form, so that it agrees with the NumPy reference to ~1e-13 N.m
 * and can serve as a fast stand-in for it on 1M-state parity runs.
 *
 * Parity status: PINNED for rne / nov / min-jerk -- oracle/make_golden.py
 * imports the unmodified reference modules (/root/reference/src/rne.py,
 * min_jerk_v2.py) and writes tests/golden/*.npz; tests/test_oracle_golden.py
 * checks this file against those vectors and against SURVEY.md Appendix B.
 * `dyn` is a DEFINED oracle ("parity unpinned"): the reference's arithmetic
 * for it lives in a module (`panda_dynamics_model`) that is absent from the
 * reference tree and in PyBullet (see oracle_dyn below).
 *
 * Reference citations are /root/reference/src/<file>:<line>.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define NLINK 10 /* 7 arm links + link8 + hand + optional payload link (rne.py:200-203) */

/* rne.py:47-54 -- modified-DH rows (a, d, alpha); theta = q[k] (0 for row 7). */
static const double DH_A[8] = {0, 0, 0, 0.0825, -0.0825, 0, 0.088, 0};
static const double DH_D[8] = {0.333, 0, 0.316, 0, 0.384, 0, 0.0, 0.107};
#define PI_ 3.141592653589793 /* np.pi */
static const double DH_ALPHA[8] = {0, -PI_ / 2, PI_ / 2, PI_ / 2, -PI_ / 2, PI_ / 2, PI_ / 2, 0};

/*
 * The inertial tables as one record, so the same restatement can be run with OTHER inertial parameters (a
 * different hand, a recalibrated link): the reference keeps them in the module-level lists ms / cs /
 * inertia_matrices (rne.py:102,119,138), which oracle/make_golden.py overwrites in place to pin this form.  Field
 * order and sizes are those of `tcmp_model` (include/tcmp.h); the DH geometry is not part of it.
 */
typedef struct oracle_model {
    double mass[9];         /* rne.py:125-136: panda_link1..7, panda_link8, panda_hand */
    double com[9][3];       /* rne.py:106-117; the payload link's COM stays (0,0,0): add_payload ignores r (:181-188) */
    double inertia[9][6];   /* rne.py:65-75: ixx ixy ixz iyy iyz izz about the COM */
    double payload_radius;  /* rne.py:182,186: hand_width + 0.025 */
    double tool_z;          /* panda_mod.urdf:87-91: grasp target past the flange (dyn mode) */
    double torque_limit[7]; /* panda_primitives.py:162-166 + utils.py:1558 + panda_mod.urdf:127..283 (effort) */
} oracle_model;

static const oracle_model PANDA = {
    {4.970684, 0.646926, 3.228604, 3.587895, 1.225946, 1.666555, 7.35522e-01, 0.0, 0.68},
    {{3.875e-03, 2.081e-03, -0.1750},       {-3.141e-03, -2.872e-02, 3.495e-03},
     {2.7518e-02, 3.9252e-02, -6.6502e-02}, {-5.317e-02, 1.04419e-01, 2.7454e-02},
     {-1.1953e-02, 4.1065e-02, -3.8437e-02}, {6.0149e-02, -1.4117e-02, -1.0517e-02},
     {1.0517e-02, -4.252e-03, 6.1597e-02},  {0, 0, 0}, {0, 0, 0}},
    {{7.0337e-01, -1.3900e-04, 6.7720e-03, 7.0661e-01, 1.9169e-02, 9.1170e-03},
     {7.9620e-03, -3.9250e-03, 1.0254e-02, 2.8110e-02, 7.0400e-04, 2.5995e-02},
     {3.7242e-02, -4.7610e-03, -1.1396e-02, 3.6155e-02, -1.2805e-02, 1.0830e-02},
     {2.5853e-02, 7.7960e-03, -1.3320e-03, 1.9552e-02, 8.6410e-03, 2.8323e-02},
     {3.5549e-02, -2.1170e-03, -4.0370e-03, 2.9474e-02, 2.2900e-04, 8.6270e-03},
     {1.9640e-03, 1.0900e-04, -1.1580e-03, 4.3540e-03, 3.4100e-04, 5.4330e-03},
     {1.2516e-02, -4.2800e-04, -1.1960e-03, 1.0027e-02, -7.4100e-04, 4.8150e-03},
     {0.001, 0.0, 0.0, 0.001, 0.0, 0.001},
     {0.1, 0.0, 0.0, 0.1, 0.0, 0.1}},
    0.14 + 0.025,
    0.105,
    {87, 87, 87, 87, 12, 12, 12},
};

void oracle_model_default(oracle_model *m) { *m = PANDA; }

/* ---- small dense helpers (row-major) ------------------------------------ */
static void skew(const double v[3], double S[3][3]) { /* rne.py:4-7 */
    S[0][0] = 0;     S[0][1] = -v[2]; S[0][2] = v[1];
    S[1][0] = v[2];  S[1][1] = 0;     S[1][2] = -v[0];
    S[2][0] = -v[1]; S[2][1] = v[0];  S[2][2] = 0;
}
static void mm3(const double A[3][3], const double B[3][3], double C[3][3]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += A[i][k] * B[k][j];
            C[i][j] = s;
        }
}
static void mv6(const double A[6][6], const double x[6], double y[6]) {
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += A[i][k] * x[k];
        y[i] = s;
    }
}
static void mm6(const double A[6][6], const double B[6][6], double C[6][6]) {
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < 6; k++) s += A[i][k] * B[k][j];
            C[i][j] = s;
        }
}

/* rne.py:32-44 get_tf_mat + rne.py:46-63 get_parent_to_child_transform(q, k, k+1):
 * T = DH_k(theta), returned inverted.  The reference inverts with
 * np.linalg.inv (LAPACK LU); a rigid transform's inverse is [R^T, -R^T p],
 * which differs from the LU result only in the last ulp. */
static void xup_matrix(int k, double theta, double X[4][4]) {
    double T[4][4];
    memset(T, 0, sizeof(T));
    if (k < 8) {
        double a = DH_A[k], d = DH_D[k], al = DH_ALPHA[k];
        double cq = cos(theta), sq = sin(theta), ca = cos(al), sa = sin(al);
        T[0][0] = cq;      T[0][1] = -sq;     T[0][2] = 0;   T[0][3] = a;
        T[1][0] = sq * ca; T[1][1] = cq * ca; T[1][2] = -sa; T[1][3] = -sa * d;
        T[2][0] = sq * sa; T[2][1] = cq * sa; T[2][2] = ca;  T[2][3] = ca * d;
        T[3][3] = 1;
    } else { /* rne.py:60-61: identity for links beyond the DH table */
        T[0][0] = T[1][1] = T[2][2] = T[3][3] = 1;
    }
    memset(X, 0, sizeof(double) * 16);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) X[i][j] = T[j][i];
    for (int i = 0; i < 3; i++) {
        double s = 0;
        for (int j = 0; j < 3; j++) s += X[i][j] * T[j][3];
        X[i][3] = -s;
    }
    X[3][3] = 1;
}

/* rne.py:9-14 adjoint(a) = [[R, skew(t) R], [0, R]] */
static void adjoint(const double X[4][4], double A[6][6]) {
    double R[3][3], S[3][3], SR[3][3], t[3];
    for (int i = 0; i < 3; i++) {
        t[i] = X[i][3];
        for (int j = 0; j < 3; j++) R[i][j] = X[i][j];
    }
    skew(t, S);
    mm3(S, R, SR);
    memset(A, 0, sizeof(double) * 36);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            A[i][j] = R[i][j];
            A[i][j + 3] = SR[i][j];
            A[i + 3][j + 3] = R[i][j];
        }
}

/* rne.py:21-24 crm(v) = [[skew(w), skew(vlin)], [0, skew(w)]], v = [lin; ang] */
static void crm(const double v[6], double M[6][6]) {
    double Sw[3][3], Sv[3][3];
    skew(v + 3, Sw);
    skew(v, Sv);
    memset(M, 0, sizeof(double) * 36);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            M[i][j] = Sw[i][j];
            M[i][j + 3] = Sv[i][j];
            M[i + 3][j + 3] = Sw[i][j];
        }
}

/* rne.py:16-19 spatial_inertia(m, c, I) = [[m 1, m C^T], [m C, I + m C C^T]] */
static void spatial_inertia(double m, const double c[3], const double I6[6], double S[6][6]) {
    double C[3][3], Ct[3][3], CCt[3][3];
    double I3[3][3] = {{I6[0], I6[1], I6[2]}, {I6[1], I6[3], I6[4]}, {I6[2], I6[4], I6[5]}}; /* rne.py:82 */
    skew(c, C);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Ct[i][j] = C[j][i];
    /* m * C @ C^T evaluates as (m*C) @ C^T in NumPy */
    double mC[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) mC[i][j] = m * C[i][j];
    mm3(mC, Ct, CCt);
    memset(S, 0, sizeof(double) * 36);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            S[i][j] = (i == j) ? m : 0.0;
            S[i][j + 3] = m * Ct[i][j];
            S[i + 3][j] = mC[i][j];
            S[i + 3][j + 3] = I3[i][j] + CCt[i][j];
        }
}

/*
 * rne.py:198-254 rne(q, qd, qdd) with the module-global payload state made an
 * explicit argument: has_payload / payload_mass are what add_payload(r, m)
 * (rne.py:181-188) would have left in the globals.
 */
static void rne_with(const oracle_model *mdl, const double q7[7], const double qd7[7], const double qdd7[7],
                     int has_payload, double payload_mass, double tau_out[7]) {
    double q[NLINK] = {0}, qd[NLINK] = {0}, qdd[NLINK] = {0}; /* rne.py:206-208 */
    for (int i = 0; i < 7; i++) {
        q[i] = q7[i];
        qd[i] = qd7[i];
        qdd[i] = qdd7[i];
    }
    const int nb = 9 + (has_payload ? 1 : 0); /* rne.py:211 */
    double v[NLINK][6], a[NLINK][6], f[NLINK][6], Xups[NLINK][4][4];
    const double neg_a_grav[6] = {0, 0, 9.81, 0, 0, 0}; /* -a_grav, rne.py:199,232 */

    for (int i = 1; i <= nb; i++) { /* forward pass, rne.py:217-241 */
        const int k = i - 1;
        double vJ[6] = {0, 0, 0, 0, 0, qd[k]};
        double aJ[6] = {0, 0, 0, 0, 0, qdd[k]};
        double X[4][4], A[6][6];
        xup_matrix(k, q[k], X);
        if (i == 7) X[2][3] = 0; /* rne.py:226-227 */
        adjoint(X, A);
        if (i - 1 == 0) { /* parent(i) == 0, rne.py:228-232 */
            memcpy(v[k], vJ, sizeof(vJ));
            mv6(A, neg_a_grav, a[k]);
            for (int r = 0; r < 6; r++) a[k][r] += aJ[r];
        } else { /* rne.py:233-236 */
            double t[6], M[6][6], c[6];
            mv6(A, v[k - 1], t);
            for (int r = 0; r < 6; r++) v[k][r] = t[r] + vJ[r];
            mv6(A, a[k - 1], t);
            crm(v[k], M);
            mv6(M, vJ, c);
            for (int r = 0; r < 6; r++) a[k][r] = t[r] + aJ[r] + c[r];
        }
        memcpy(Xups[k], X, sizeof(X));

        /* rne.py:240-241  f = I a + crf(v) @ I @ v, crf = -crm^T (rne.py:26-27) */
        double m, I6[6];
        static const double payload_com[3] = {0, 0, 0}; /* add_payload ignores r: rne.py:181-188 */
        const double *com = k < 9 ? mdl->com[k] : payload_com;
        if (k < 9) {
            m = mdl->mass[k];
            memcpy(I6, mdl->inertia[k], sizeof(I6));
        } else { /* payload link: rne.py:85-100 new_inertia([0,0,0.14+0.025], m) */
            const double r[3] = {0, 0, mdl->payload_radius};
            m = payload_mass;
            I6[0] = m * (r[1] * r[1] + r[2] * r[2]);
            I6[1] = -m * (r[0] * r[1]);
            I6[2] = -m * (r[0] * r[2]);
            I6[3] = m * (r[0] * r[0] + r[2] * r[2]);
            I6[4] = -m * (r[1] * r[2]);
            I6[5] = m * (r[0] * r[0] + r[1] * r[1]);
        }
        double S[6][6], M[6][6], F[6][6], FS[6][6], t1[6], t2[6];
        spatial_inertia(m, com, I6, S);
        crm(v[k], M);
        for (int r = 0; r < 6; r++)
            for (int c2 = 0; c2 < 6; c2++) F[r][c2] = -M[c2][r];
        mv6(S, a[k], t1);
        mm6(F, S, FS); /* NumPy evaluates crf(v) @ I @ v left to right */
        mv6(FS, v[k], t2);
        for (int r = 0; r < 6; r++) f[k][r] = t1[r] + t2[r];
    }

    double tau[NLINK];
    for (int i = nb; i >= 1; i--) { /* backward pass, rne.py:245-251 */
        const int k = i - 1;
        tau[k] = f[k][5];
        if (i - 1 != 0) {
            double A[6][6], t[6];
            adjoint(Xups[k], A);
            for (int r = 0; r < 6; r++) {
                double s = 0;
                for (int c2 = 0; c2 < 6; c2++) s += A[c2][r] * f[k][c2];
                t[r] = s;
            }
            for (int r = 0; r < 6; r++) f[k - 1][r] += t[r];
        }
    }
    for (int i = 0; i < 7; i++) tau_out[i] = tau[i]; /* rne.py:253 */
}

void oracle_rne(const double q7[7], const double qd7[7], const double qdd7[7],
                int has_payload, double payload_mass, double tau_out[7]) {
    rne_with(&PANDA, q7, qd7, qdd7, has_payload, payload_mass, tau_out);
}

/* panda_primitives.py:182-188: infeasible iff any |tau_i| >= limit_i for i in 0..5
 * (range(len(max_limits)-1): joint 7 is never tested; EPS = 1, :165). */
static int within_limits_of(const oracle_model *mdl, const double tau[7]) {
    for (int i = 0; i < 6; i++)
        if (fabs(tau[i]) >= mdl->torque_limit[i] * 1) return 0;
    return 1;
}
int oracle_within_limits(const double tau[7]) { return within_limits_of(&PANDA, tau); }

/* ---- forward kinematics of the DH chain (used by the defined `dyn` oracle) --- */
static void mm4(const double A[4][4], const double B[4][4], double C[4][4]) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += A[i][k] * B[k][j];
            C[i][j] = s;
        }
}
static void dh_matrix(int k, double theta, double T[4][4]) {
    double a = DH_A[k], d = DH_D[k], al = DH_ALPHA[k];
    double cq = cos(theta), sq = sin(theta), ca = cos(al), sa = sin(al);
    memset(T, 0, sizeof(double) * 16);
    T[0][0] = cq;      T[0][1] = -sq;     T[0][3] = a;
    T[1][0] = sq * ca; T[1][1] = cq * ca; T[1][2] = -sa; T[1][3] = -sa * d;
    T[2][0] = sq * sa; T[2][1] = cq * sa; T[2][2] = ca;  T[2][3] = ca * d;
    T[3][3] = 1;
}

/*
 * DEFINED oracle for `dyn` (panda_primitives.py:60-116) -- PARITY UNPINNED.
 *   tau = M(q) qdd + C(q,qd) qd + g(q) + J(q)^T [0,0,m*9.81,0,0,0]        (:85-111)
 * M, C, g come from `panda_dynamics_model`, which is not in the reference tree;
 * they are defined here as the same rigid-body model the reference's own rne.py
 * carries (no payload link), so M qdd + C qd + g == rne(q, qd, qdd) without payload.
 * J is the geometric Jacobian (linear rows) of the panda_grasptarget origin:
 * link7 frame -> Tz(0.107) [link8, DH row 7] -> hand (Rz(-pi/4), no offset,
 * panda_mod.urdf:7-11) -> Tz(0.105) (panda_mod.urdf:87-91); the unmodified DH
 * chain is used for J (no zeroing of the joint-7 offset: that quirk is rne.py's).
 * J^T F with F = (0,0,m g) is  tau_i = m g * (z_i x (p_tool - p_i)).z.
 * The reference applies payload_mass unconditionally here (no 0.01 threshold, :71-76).
 */
static void dyn_with(const oracle_model *mdl, const double q[7], const double qd[7], const double qdd[7],
                     double payload_mass, double tau_out[7]) {
    double tau[7];
    rne_with(mdl, q, qd, qdd, 0, 0.0, tau);
    double T[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    double z[7][3], p[7][3];
    for (int k = 0; k < 8; k++) {
        double D[4][4], N[4][4];
        dh_matrix(k, k < 7 ? q[k] : 0.0, D);
        mm4(T, D, N);
        memcpy(T, N, sizeof(N));
        if (k < 7)
            for (int r = 0; r < 3; r++) {
                z[k][r] = T[r][2];
                p[k][r] = T[r][3];
            }
    }
    /* tool origin: the hand's Rz(-pi/4) does not move a point on its z axis */
    double pt[3];
    for (int r = 0; r < 3; r++) pt[r] = T[r][3] + T[r][2] * mdl->tool_z;
    const double force = payload_mass * 9.81; /* panda_primitives.py:101 */
    for (int i = 0; i < 7; i++) {
        double dx = pt[0] - p[i][0], dy = pt[1] - p[i][1];
        double jz = z[i][0] * dy - z[i][1] * dx; /* (z_i x d).z */
        tau_out[i] = jz * force + tau[i];
    }
}
void oracle_dyn(const double q[7], const double qd[7], const double qdd[7],
                double payload_mass, double tau_out[7]) {
    dyn_with(&PANDA, q, qd, qdd, payload_mass, tau_out);
}

/*
 * One torque test. mode: 0 = rne (panda_primitives.py:171-191), 1 = nov (:130-151),
 * 2 = dyn (:66-115, defined oracle), 3 = base (:13-16).  Writes tau (7) if non-NULL.
 * payload_threshold is 0.01 for the reference's closures (:139,:178); the raw
 * rne.add_payload() rule is `m > 0` (rne.py:184), i.e. threshold 0.
 */
static int torque_test_with(const oracle_model *mdl, int mode, const double q[7], const double qd[7],
                            const double qdd[7], double payload_mass, double payload_threshold,
                            double tau_out[7]) {
    static const double Z[7] = {0, 0, 0, 0, 0, 0, 0};
    double tau[7] = {0};
    if (mode == 3) {
        if (tau_out) memcpy(tau_out, tau, sizeof(tau));
        return 1;
    }
    const double *v = (mode == 1 || !qd) ? Z : qd;
    const double *a = (mode == 1 || !qdd) ? Z : qdd;
    if (mode == 2) {
        dyn_with(mdl, q, v, a, payload_mass, tau);
    } else {
        int has = payload_mass > payload_threshold;
        rne_with(mdl, q, v, a, has, has ? payload_mass : 0.0, tau);
    }
    if (tau_out) memcpy(tau_out, tau, sizeof(tau));
    return within_limits_of(mdl, tau);
}
int oracle_torque_test(int mode, const double q[7], const double qd[7], const double qdd[7],
                       double payload_mass, double payload_threshold, double tau_out[7]) {
    return torque_test_with(&PANDA, mode, q, qd, qdd, payload_mass, payload_threshold, tau_out);
}

/* Batched, SoA [7][n] like the C-ABI; OpenMP over states when built with -fopenmp.
 * payload_mass may be NULL (payload_scalar is used for every state). */
void oracle_torque_test_batch_model(const oracle_model *mdl, int mode, int64_t n, const double *q,
                                    const double *qd, const double *qdd, const double *payload_mass,
                                    double payload_scalar, double payload_threshold, double *tau_out,
                                    uint8_t *feasible_out, int nthreads) {
    if (!mdl) mdl = &PANDA;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (int64_t s = 0; s < n; s++) {
        double qs[7], vs[7], as[7], tau[7];
        for (int j = 0; j < 7; j++) {
            qs[j] = q[j * n + s];
            vs[j] = qd ? qd[j * n + s] : 0.0;
            as[j] = qdd ? qdd[j * n + s] : 0.0;
        }
        double m = payload_mass ? payload_mass[s] : payload_scalar;
        int ok = torque_test_with(mdl, mode, qs, vs, as, m, payload_threshold, tau);
        if (tau_out)
            for (int j = 0; j < 7; j++) tau_out[j * n + s] = tau[j];
        if (feasible_out) feasible_out[s] = (uint8_t)ok;
    }
}
void oracle_torque_test_batch(int mode, int64_t n, const double *q, const double *qd,
                              const double *qdd, const double *payload_mass,
                              double payload_scalar, double payload_threshold,
                              double *tau_out, uint8_t *feasible_out, int nthreads) {
    oracle_torque_test_batch_model(&PANDA, mode, n, q, qd, qdd, payload_mass, payload_scalar,
                                   payload_threshold, tau_out, feasible_out, nthreads);
}

/* ---- min-jerk (min_jerk_v2.py) ------------------------------------------- */

/* min_jerk_v2.py:80-142 minjerk_coefficients(points[L][7]) with duration_array=None
 * (unit segment durations, :102-103).  coeffs layout [L-1][7][6] = a0..a5. */
void oracle_minjerk_coefficients(int L, const double *points /* [L][7] */, double *coeffs) {
    double x[7], v[7], a[7];
    const int N = L - 1;
    for (int j = 0; j < 7; j++) {
        x[j] = points[j];
        v[j] = 0;
        a[j] = 0; /* never updated in the loop: every segment starts with zero acceleration */
    }
    for (int i = 0; i < N; i++) {
        const double t = 1.0;
        for (int j = 0; j < 7; j++) {
            double gx = points[(i + 1) * 7 + j], gv, ga = 0.0;
            if (i == N - 1) {
                gv = 0.0;
            } else { /* :111-118 */
                double t0 = t, t1 = 1.0;
                double d0 = points[(i + 1) * 7 + j] - points[i * 7 + j];
                double d1 = points[(i + 2) * 7 + j] - points[(i + 1) * 7 + j];
                double v0 = d0 / t0, v1 = d1 / t1;
                gv = (v0 * v1 >= 1e-10) ? 0.5 * (v0 + v1) : 0.0;
            }
            double A = (gx - (x[j] + v[j] * t + (a[j] / 2.0) * t * t)) / (t * t * t); /* :121 */
            double B = (gv - (v[j] + a[j] * t)) / (t * t);                            /* :122 */
            double C = (ga - a[j]) / t;                                               /* :123 */
            double *c = coeffs + ((size_t)i * 7 + j) * 6;
            c[0] = x[j];
            c[1] = v[j];
            c[2] = a[j] / 2.0;
            c[3] = 10 * A - 4 * B + 0.5 * C;
            c[4] = (-15 * A + 7 * B - C) / t;
            c[5] = (6 * A - 3 * B + 0.5 * C) / (t * t);
            x[j] = gx; /* :132-133 (a is not carried) */
            v[j] = gv;
        }
    }
}

/* min_jerk_v2.py:204-222 _minjerk_trajectory_point (tm = 1) */
static void minjerk_point(const double c[6], double t, double *x, double *v, double *a) {
    *x = c[0] + c[1] * t + c[2] * pow(t, 2) + c[3] * pow(t, 3) + c[4] * pow(t, 4) + c[5] * pow(t, 5);
    *v = c[1] + 2 * c[2] * t + 3 * c[3] * pow(t, 2) + 4 * c[4] * pow(t, 3) + 5 * c[5] * pow(t, 4);
    *a = 2 * c[2] + 6 * c[3] * t + 12 * c[4] * pow(t, 2) + 20 * c[5] * pow(t, 3);
}

/* np.linspace(1/n, 1, n)[i] (min_jerk_v2.py:176-177): start + i*step, endpoint forced. */
double oracle_linspace_sample(int n, int i) {
    double interval = 1.0 / n;
    if (n == 1) return interval;
    double step = (1.0 - interval) / (n - 1);
    return (i == n - 1) ? 1.0 : interval + i * step;
}

/* min_jerk_v2.py:144-182 minjerk_trajectory: samples laid out [(L-1)*n][7] each. */
void oracle_minjerk_trajectory(int L, const double *coeffs, int num_intervals,
                               double *xs, double *vs, double *as) {
    for (int seg = 0; seg < L - 1; seg++)
        for (int it = 0; it < num_intervals; it++) {
            double t = oracle_linspace_sample(num_intervals, it);
            size_t row = (size_t)seg * num_intervals + it;
            for (int j = 0; j < 7; j++)
                minjerk_point(coeffs + ((size_t)seg * 7 + j) * 6, t, &xs[row * 7 + j],
                              &vs[row * 7 + j], &as[row * 7 + j]);
        }
}

/*
 * Config-4 edge check: 2-point min-jerk from qa to qb, W samples
 * t = linspace(1/W, 1, W), each tested with torque test `mode`; first_fail = index
 * of the first infeasible waypoint (rrt_star.py:208-210 stops at the first failure),
 * or W when the whole edge is feasible.  qa/qb SoA [7][n_edges].
 */
void oracle_edge_feasibility_model(const oracle_model *mdl, int mode, int64_t n_edges, int W, const double *qa,
                                   const double *qb, double payload_mass, double payload_threshold,
                                   int32_t *first_fail, int nthreads) {
    if (!mdl) mdl = &PANDA;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 64)
#endif
    for (int64_t e = 0; e < n_edges; e++) {
        double pts[14], coeffs[7 * 6];
        for (int j = 0; j < 7; j++) {
            pts[j] = qa[j * n_edges + e];
            pts[7 + j] = qb[j * n_edges + e];
        }
        oracle_minjerk_coefficients(2, pts, coeffs);
        int ff = W;
        for (int w = 0; w < W; w++) {
            double t = oracle_linspace_sample(W, w), x[7], v[7], a[7];
            for (int j = 0; j < 7; j++) minjerk_point(coeffs + j * 6, t, &x[j], &v[j], &a[j]);
            if (!torque_test_with(mdl, mode, x, v, a, payload_mass, payload_threshold, 0)) {
                ff = w;
                break;
            }
        }
        first_fail[e] = ff;
    }
}
void oracle_edge_feasibility(int mode, int64_t n_edges, int W, const double *qa, const double *qb,
                             double payload_mass, double payload_threshold,
                             int32_t *first_fail, int nthreads) {
    oracle_edge_feasibility_model(&PANDA, mode, n_edges, W, qa, qb, payload_mass, payload_threshold, first_fail,
                                  nthreads);
}

/* Final-trajectory check (rrt_star.py:203-210): samples of the multi-segment min-jerk
 * through path[L][7]; mask per sample and the first failing index (n_samples if none). */
void oracle_traj_feasibility_model(const oracle_model *mdl, int mode, int L, const double *path, int num_intervals,
                                   double payload_mass, double payload_threshold,
                                   double *tau_out /* [n_samples][7] or NULL */,
                                   uint8_t *mask_out, int32_t *first_fail) {
    if (!mdl) mdl = &PANDA;
    const int ns = (L - 1) * num_intervals;
    double coeffs[(L - 1) * 7 * 6];
    oracle_minjerk_coefficients(L, path, coeffs);
    int ff = ns;
    for (int seg = 0; seg < L - 1; seg++)
        for (int it = 0; it < num_intervals; it++) {
            int row = seg * num_intervals + it;
            double t = oracle_linspace_sample(num_intervals, it), x[7], v[7], a[7], tau[7];
            for (int j = 0; j < 7; j++)
                minjerk_point(coeffs + ((size_t)seg * 7 + j) * 6, t, &x[j], &v[j], &a[j]);
            int ok = torque_test_with(mdl, mode, x, v, a, payload_mass, payload_threshold, tau);
            if (tau_out)
                for (int j = 0; j < 7; j++) tau_out[(size_t)row * 7 + j] = tau[j];
            if (mask_out) mask_out[row] = (uint8_t)ok;
            if (!ok && row < ff) ff = row;
        }
    if (first_fail) *first_fail = ff;
}
void oracle_traj_feasibility(int mode, int L, const double *path, int num_intervals,
                             double payload_mass, double payload_threshold,
                             double *tau_out, uint8_t *mask_out, int32_t *first_fail) {
    oracle_traj_feasibility_model(&PANDA, mode, L, path, num_intervals, payload_mass, payload_threshold, tau_out,
                                  mask_out, first_fail);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
