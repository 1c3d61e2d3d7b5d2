// ik_core.cuh -- solver core shared by the CUDA kernels (ik_kernels.cu) and a host-compiled
// debugging harness (tests/native/ik_host.cu; never part of libtcmp.so).
//
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
#pragma once
#include <math.h>
#include <stdint.h>
#ifdef __CUDACC__
#define TCMP_HD __host__ __device__
// Measured: making make_root / atan2_checked / solve_shoulder real (noinline) functions and not unrolling the root
// loops halves the SASS (140 KB -> 72 KB) but is 15 % SLOWER (call ABI + 432 B of local stack), so everything
// stays inlined; define TCMP_OUTLINE as __noinline__ to reproduce.
#define TCMP_OUTLINE
#ifndef TCMP_IK_UNROLL
#define TCMP_IK_UNROLL 1
#endif
#if TCMP_IK_UNROLL
#define TCMP_ROOT_LOOP
#else
#define TCMP_ROOT_LOOP _Pragma("unroll 1")
#endif
#else
#define TCMP_HD
#define TCMP_OUTLINE
#define TCMP_ROOT_LOOP
#endif

// TCMP_IK_TABLE_SINCOS = 1 (default 0, candidate for the next round): sin / cos of the solved joint angles from the
// 1024-entry table K1 uses (csrc/sincos_table.inc; x = k pi/512 + r, two short polynomials in r, abs error 2.3e-16)
// instead of libm's sincos -- a fraction of the inlined code of a kernel that stalls on instruction fetch.  The
// solution COUNT must not move: tests/test_ik_core_host.py soaks the host build of this variant against the
// compiled reference solver.  The table is read through the read-only cache on the device.
#ifndef TCMP_IK_TABLE_SINCOS
#define TCMP_IK_TABLE_SINCOS 0
#endif
#include <string.h>

namespace tcmp {
namespace ik {

#if TCMP_IK_TABLE_SINCOS
struct alignas(16) SinCosEntry { double s, c; };
#ifdef __CUDACC__
static __device__ const SinCosEntry kSinCosDev[1024] = {
#include "sincos_table.inc"
};
#endif
static const SinCosEntry kSinCosHost[1024] = {
#include "sincos_table.inc"
};
#endif

// sin and cos of one angle: the table path while |x| < 4096 rad (every angle the solver produces), libm otherwise.
TCMP_HD inline void sincos_ik(double x, double *s, double *c) {
#if TCMP_IK_TABLE_SINCOS
    int64_t bits;
    memcpy(&bits, &x, 8);
    if (((int)(bits >> 32) & 0x7fffffff) < 0x40b00000) {
        const double kt = fma(x, 162.97466172610082 /* 512 / pi */, 6755399441055744.0);   // rint via 1.5 * 2^52
        int64_t kb;
        memcpy(&kb, &kt, 8);
        const int idx = (int)(kb & 1023);
#ifdef __CUDA_ARCH__
        const double2 e2 = __ldg(reinterpret_cast<const double2 *>(kSinCosDev) + idx);
        const double es = e2.x, ec = e2.y;
#else
        const double es = kSinCosHost[idx].s, ec = kSinCosHost[idx].c;
#endif
        const double kd = kt - 6755399441055744.0;
        double r = fma(-kd, 1.57079632673412561417e+00 / 256, x);
        r = fma(-kd, 6.07710050630396597660e-11 / 256, r);
        const double z = r * r;
        const double sl = fma(r * z, fma(z, 1.0 / 120, -1.0 / 6), r);
        const double cm = z * fma(z, 1.0 / 24, -0.5);
        *s = fma(ec, sl, fma(es, cm, es));
        *c = fma(-es, sl, fma(ec, cm, ec));
        return;
    }
#endif
    sincos(x, s, c);
}

// Products and sums that must round exactly like the reference's compiled expressions (g++ -O2 on x86-64: one rounding
// per operation, no contraction): nvcc fuses a * b + c into an FMA unless told otherwise, and next to a double root of
// the solver (asin / acos argument within ~1e-12 of +-1) a 1e-16 difference in that argument moves the root by 1e-8 and
// can flip a duplicate-root or branch test -- i.e. the SOLUTION COUNT.  Everything on the path to such a test (wrist
// centre, the asin arguments of j3 and j5, U / W / K of j4, the residual rotation M) is written with xmul / xadd in the
// reference's own association order; the residual checks (compared with 1e-5) are left to the compiler.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
#else
inline double xmul(double a, double b) { return a * b; }   // host harness: built with -ffp-contract=off
inline double xadd(double a, double b) { return a + b; }
#endif
TCMP_HD inline double xsub(double a, double b) { return xadd(a, -b); }

constexpr double kPi = 3.14159265358979;      // IKPI   (ikfast_panda_arm.cpp:68) -- truncated on purpose
constexpr double k2Pi = 6.28318530717959;     // IK2PI  (:67)
constexpr double kPi2 = 1.57079632679490;     // IKPI_2 (:69)
constexpr double kHalfPiLit = 1.5707963267949;  // literal the generated formulas use for pi/2
constexpr double kSinCosThresh = 1e-7;        // IKFAST_SINCOS_THRESH   (:110)
constexpr double kAtan2Thresh = 1e-7;         // IKFAST_ATAN2_MAGTHRESH (:115)
constexpr double kSolutionThresh = 1e-6;      // IKFAST_SOLUTION_THRESH (:120)
constexpr double kEvalThresh = 1e-5;          // IKFAST_EVALCOND_THRESH (:125)
constexpr double kBranchThresh = 1e-6;        // "< 0.0000010000000000" tests on j*eval
constexpr double kAngleThresh = 5e-6;         // "< 0.0000050000000000" tests on special angles

TCMP_HD inline double clamp_asin(double f) {  // IKasin (:136-142)
    if (f <= -1) return -kPi2;
    if (f >= 1) return kPi2;
    return asin(f);
}
TCMP_HD inline double clamp_acos(double f) {  // IKacos (:169-175)
    if (f <= -1) return kPi;
    if (f >= 1) return 0.0;
    return acos(f);
}
TCMP_HD inline bool in_unit(double f) {  // the range guard in front of IKasin / IKacos
    return !(f < -1 - kSinCosThresh || f > 1 + kSinCosThresh);
}
TCMP_HD inline double sign_of(double f) { return f > 0 ? 1.0 : (f < 0 ? -1.0 : 0.0); }  // IKsign
TCMP_HD inline double pos_fmod(double x, double y) {  // IKfmod (:154-160): remainder in [0, y)
    while (x < 0) x += y;
    return fmod(x, y);
}
TCMP_HD inline double wrap_pi(double a) {  // the "> IKPI -= IK2PI / < -IKPI += IK2PI" idiom
    if (a > kPi) return a - k2Pi;
    if (a < -kPi) return a + k2Pi;
    return a;
}
// IKatan2WithCheck (:219-231): valid iff neither is NaN and |y| >= 1e-7 or |x| > 1e-7.
TCMP_HD TCMP_OUTLINE inline bool atan2_checked(double y, double x, double *out) {
    if (isnan(y) || isnan(x)) return false;
    if (!(fabs(y) >= kAtan2Thresh || fabs(x) > kAtan2Thresh)) return false;
    *out = atan2(y, x);
    return true;
}

struct Root {
    double a, s, c;  // wrapped angle; sin/cos of the unwrapped angle, as the solver computes them
};
TCMP_HD TCMP_OUTLINE inline Root make_root(double angle) {
    Root r;
    sincos_ik(angle, &r.s, &r.c);
    r.a = wrap_pi(angle);
    return r;
}
TCMP_HD inline bool same_root(const Root &x, const Root &y) {  // duplicate-root test (:495)
    return fabs(x.c - y.c) < kSolutionThresh && fabs(x.s - y.s) < kSolutionThresh;
}
// A threshold test within 1 % of its threshold.  The values tested next to a double root are sqrt-amplified rounding
// residue (two roots 1e-6 apart come from an asin / acos argument 1.2e-13 from +-1, where one ulp of the argument moves
// them by 3e-10), so there the reference's own decision depends on its libm and compiler: the count is ill-conditioned.
TCMP_HD inline bool near_threshold(double value, double thresh) {
    return fabs(value - thresh) < 0.01 * thresh;
}
TCMP_HD inline bool same_root_borderline(const Root &x, const Root &y) {
    return near_threshold(fmax(fabs(x.c - y.c), fabs(x.s - y.s)), kSolutionThresh);
}

struct Pose {
    double r[3][3];
    double px, py, pz;        // wrist centre: eetrans - 0.107 R[:,2] - (0,0,0.333)   (:431-439)
    double pp, npx, npy, npz; // |p|^2 and R^T p                                      (:444-447)
    double j6, s6, c6;
};

struct Emit {
    double *sols;   // [8][7] or nullptr
    int count;
    unsigned status;
};

enum : unsigned {
    kStatusDegenerate = 1u,  // a singular branch of the decision tree was entered (and resolved like the reference if bit 1 is clear)
    kStatusUnresolved = 2u,  // ... a branch of the generated solver this restatement does not implement: solutions dropped
    kStatusInvalid = 4u,     // non-finite target: the reference trips IKFAST_ASSERT (:57,:138) and throws; here 0 solutions
    kStatusIllConditioned = 8u,  // a duplicate-root or singular-branch test came within 1 % of its threshold: the value
                                 // tested is amplified rounding residue and the reference's own count depends on its libm
    kStatusRedo = 128u       // internal (never leaves the kernels): solve_one_t<false> met the elbow singularity and wants
                             // the complete tree (solve_one_t<true>, a separate cold region of the kernel) to redo the solve
};

TCMP_HD inline void emit_solution(Emit &out, double j0, double j1, double j2, double j3, double j4,
                                              double j5, double j6) {
    if (out.sols && out.count < 8) {
        double *s = out.sols + out.count * 7;
        s[0] = j0; s[1] = j1; s[2] = j2; s[3] = j3; s[4] = j4; s[5] = j5; s[6] = j6;
    }
    ++out.count;
}

// Residual ZYZ problem (rotationfunction0, :3115): with j3,j4,j5,j6 fixed, M = R_{3..6}^T R must
// equal Rz(j0) Ry(j1) Rz(j2).
TCMP_HD TCMP_OUTLINE inline void solve_shoulder(const Pose &P, const Root &j3, const Root &j4, const Root &j5, Emit &out) {
    // M = (Rz(j3') ...)^T R, written as three successive frame changes of the columns of R
    // (:3122-3147): first about the tool axis by j6, then j5, j4, j3.
    double M[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double ri0 = P.r[i][0], ri1 = P.r[i][1], ri2 = P.r[i][2];
        const double a = xsub(xmul(P.c6, ri0), xmul(P.s6, ri1));      // x122..x124
        const double b = xsub(xmul(-P.s6, ri0), xmul(P.c6, ri1));     // x126..x128
        const double d = xadd(xmul(ri2, j5.s), xmul(j5.c, a));        // x129..x131
        const double e = xsub(xmul(j5.s, a), xmul(ri2, j5.c));        // x125 - r*cj5, x134, x132
        const double f = xsub(xmul(j4.c, d), xmul(j4.s, b));          // x133 + x120*x126, x135, x136
        M[i][0] = xsub(xmul(j3.c, f), xmul(j3.s, e));
        M[i][1] = xadd(xmul(j4.s, d), xmul(j4.c, b));
        M[i][2] = xadd(xmul(j3.c, e), xmul(j3.s, f));
    }
    // j1 = +-acos(M22)  (:3149-3167)
    Root j1r[2];
    bool j1ok[2] = {false, false};
    const double cj1 = M[2][2];
    if (cj1 >= -1 - kSinCosThresh && cj1 <= 1 + kSinCosThresh) {
        const double a = clamp_acos(cj1);
#if TCMP_IK_TABLE_SINCOS
        double s, c_unused;
        sincos_ik(a, &s, &c_unused);
#else
        const double s = sin(a);
#endif
        j1r[0] = {a, s, cj1};
        j1r[1] = {-a, -s, cj1};
        j1ok[0] = j1ok[1] = true;
    } else if (isnan(cj1)) {
        j1r[0] = {0.0, 0.0, 1.0};
        j1ok[0] = true;
    }
    if (j1ok[0] && j1ok[1]) {
        if (same_root(j1r[0], j1r[1])) j1ok[1] = false;
        if (same_root_borderline(j1r[0], j1r[1])) out.status |= kStatusIllConditioned;
    }

TCMP_ROOT_LOOP
    for (int i1 = 0; i1 < 2; ++i1) {
        if (!j1ok[i1]) continue;
        const double j1 = j1r[i1].a, s1 = j1r[i1].s, c1 = j1r[i1].c;
        const double sg = sign_of(s1);
        // general branch requires sin(j1) away from 0 and a usable (M12, M02) pair (:3185-3189)
        if (near_threshold(fabs(s1), kBranchThresh) || near_threshold(fabs(M[1][2]) + fabs(M[0][2]), kBranchThresh))
            out.status |= kStatusIllConditioned;
        if (fabs(s1) < kBranchThresh || fabs(M[1][2]) + fabs(M[0][2]) < kBranchThresh || fabs(sg) < kBranchThresh) {
            // Shoulder singularity: joint axes 0 and 2 are collinear, only j0 +- j2 is determined.  The solver
            // resolves it by pinning j2 = 0 (it reports j2 as a free parameter, which get_ik expands with 0,
            // :3217-3262 / :3280-3325) when M is a pure rotation about z (j1 ~ 0) or about z after a flip (j1 ~ pi).
            out.status |= kStatusDegenerate;
            const bool zish = fabs(M[2][0]) < kAngleThresh && fabs(M[0][2]) < kAngleThresh &&
                              fabs(M[1][2]) < kAngleThresh && fabs(M[2][1]) < kAngleThresh;
            // the solver only reaches these cases through two more guards (:3196, :3202) that hold whenever
            // sin(j1) is below the branch threshold; anything else is outside the implemented tree
            const bool guards = (fabs(s1) < kBranchThresh || fabs(sg) < kBranchThresh ||
                                 fabs(M[2][0]) + fabs(M[2][1]) < kBranchThresh) &&
                                (fabs(M[1][2]) < kBranchThresh || fabs(s1) < kBranchThresh);
            if (!guards) { out.status |= kStatusUnresolved; continue; }
            const double near0 = -3.14159265358979 + pos_fmod(3.14159265358979 + fabs(j1), 6.28318530717959);
            const double nearpi = -3.14159265358979 +
                                  pos_fmod(3.14159265358979 + fabs(-3.14159265358979 + j1), 6.28318530717959);
            double y, x;
            if (fabs(near0) < kAngleThresh && zish) { y = -M[0][1]; x = M[0][0]; }
            else if (fabs(nearpi) < kAngleThresh && zish) { y = -M[0][1]; x = -M[0][0]; }
            else { out.status |= kStatusUnresolved; continue; }
            if (fabs(y) < kAtan2Thresh && fabs(x) < kAtan2Thresh && fabs(y * y + x * x - 1) <= kSinCosThresh) continue;
            const double j0 = isnan(y) ? kPi2 : (isnan(x) ? 0.0 : atan2(y, x));   // IKatan2 (:200-209)
            // free-parameter solutions go through the +-pi wrap of IkSolution::GetSolution (ikfast.h:172-178)
            emit_solution(out, wrap_pi(j0), j1, 0.0, j3.a, j4.a, j5.a, P.j6);
            continue;
        }
        double at;
        if (!atan2_checked(M[1][2], M[0][2], &at)) continue;
        const Root j0 = make_root(-kHalfPiLit + kHalfPiLit * (1.0 / sg) + at);  // (:9544-9552)
        {
            // 8 residuals (:9586-9593): column 2 of M against Rz(j0) Ry(j1)
            const double a = M[0][2], b = M[1][2];
            const double x02 = a * j0.c + b * j0.s;              // (Rz^T M)_02
            const double x00 = M[0][0] * j0.c + M[1][0] * j0.s;  // (Rz^T M)_00
            const double x01 = M[0][1] * j0.c + M[1][1] * j0.s;  // (Rz^T M)_01
            const double e0 = a - j0.c * s1, e1 = b - j0.s * s1, e2 = b * j0.c - a * j0.s, e3 = x02 - s1;
            const double e4 = c1 * x02 - s1 * M[2][2];
            const double e5 = -(c1 * M[2][0] + s1 * x00), e6 = -(c1 * M[2][1] + s1 * x01);
            const double e7 = 1.0 - c1 * M[2][2] - s1 * x02;
            if (fabs(e0) > kEvalThresh || fabs(e1) > kEvalThresh || fabs(e2) > kEvalThresh || fabs(e3) > kEvalThresh ||
                fabs(e4) > kEvalThresh || fabs(e5) > kEvalThresh || fabs(e6) > kEvalThresh || fabs(e7) > kEvalThresh)
                continue;
        }
        if (fabs(M[2][0]) + fabs(M[2][1]) < kBranchThresh) {  // (:9601-9605); sin(j1) already tested
            // M20^2 + M21^2 = sin^2 j1 for a rotation matrix, so this needs a non-orthonormal input: the generated
            // special cases behind it (:9607-12470) are not implemented
            out.status |= kStatusDegenerate | kStatusUnresolved;
            continue;
        }
        if (!atan2_checked(M[2][1], -M[2][0], &at)) continue;
        const Root j2 = make_root(-kHalfPiLit + kHalfPiLit * (1.0 / sg) + at);  // (:12473-12481)
        {
            // 12 residuals (:12520-12531): M against Rz(j0) Ry(j1) Rz(j2), in three frames
            const double c0 = j0.c, s0 = j0.s, c2 = j2.c, s2 = j2.s;
            const double x00 = c0 * M[0][0] + s0 * M[1][0], x01 = c0 * M[0][1] + s0 * M[1][1];
            const double y10 = c0 * M[1][0] - s0 * M[0][0], y11 = c0 * M[1][1] - s0 * M[0][1];
            const double e[12] = {
                s1 * c2 + M[2][0],
                M[2][1] - s2 * s1,
                x01 + c1 * s2,
                y10 - s2,
                y11 - c2,
                s0 * c2 + M[0][1] + s2 * c0 * c1,
                x00 - c1 * c2,
                s0 * s2 + M[0][0] - c0 * c1 * c2,
                s0 * c1 * s2 - c0 * c2 + M[1][1],
                M[1][0] - c0 * s2 - c1 * c2 * s0,
                c1 * x01 - M[2][1] * s1 + s2,
                c1 * x00 - M[2][0] * s1 - c2,
            };
            bool bad = false;
#pragma unroll
            for (int t = 0; t < 12; ++t) bad = bad || (fabs(e[t]) > kEvalThresh);
            if (bad) continue;
        }
        emit_solution(out, j0.a, j1, j2.a, j3.a, j4.a, j5.a, P.j6);
    }
}

// Wrist-centre reduction and invariants of one target (:416-447).
TCMP_HD inline void prepare_pose(const double R[9], double tx, double ty, double tz, double j6, Pose &P) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) P.r[i][j] = R[i * 3 + j];
    P.j6 = j6;
    sincos_ik(j6, &P.s6, &P.c6);
    P.px = xadd(tx, xmul(-0.107, P.r[0][2]));
    P.py = xadd(xmul(-0.107, P.r[1][2]), ty);
    P.pz = xadd(xadd(-0.333, tz), xmul(-0.107, P.r[2][2]));
    P.pp = xadd(xadd(xmul(P.px, P.px), xmul(P.py, P.py)), xmul(P.pz, P.pz));
    P.npx = xadd(xadd(xmul(P.px, P.r[0][0]), xmul(P.py, P.r[1][0])), xmul(P.pz, P.r[2][0]));
    P.npy = xadd(xadd(xmul(P.px, P.r[0][1]), xmul(P.py, P.r[1][1])), xmul(P.pz, P.r[2][1]));
    P.npz = xadd(xadd(xmul(P.px, P.r[0][2]), xmul(P.py, P.r[1][2])), xmul(P.pz, P.r[2][2]));
}

// asin argument of j3 (:461-463), in the reference's association order: c0 + c1 pp + (c2 cj6) npx + (c3 npy) sj6
TCMP_HD inline double j3_asin_argument(const Pose &P) {
    double t = xadd(0.986881610513004, xmul(-3.89793688895078, P.pp));
    t = xadd(t, xmul(xmul(0.686036892455338, P.c6), P.npx));
    return xadd(t, xmul(xmul(-0.686036892455338, P.npy), P.s6));
}

// The solver's first gate (:461-462): the asin argument of j3.  0 = no solution for this (pose, free value),
// 1 = continue with solve_one, 2 = non-finite input.  ~45 % of the solves of a free-joint sweep stop here.
TCMP_HD inline int screen_pose(const Pose &P) {
    const double arg3 = j3_asin_argument(P);
    if (!(arg3 == arg3)) return 2;
    return in_unit(arg3) ? 1 : 0;
}

// Elbow singularity (:2409-2850): K = -0.0825 + 0.0825 cos j3 + 0.316 sin j3 ~ 0 at j3 ~ 0 (forearm and upper-arm axes
// collinear) and at j3 ~ 2 atan(0.316 / 0.0825) = 2.63084142381503 (shoulder centre on the forearm axis).  The solver
// falls through two more guard levels -- products of q0 with U and W, which are below the threshold whenever the first
// level tripped on a consistent pose -- and then matches j3 against the two special angles.
// Returns the number of j4 candidates written to j4c (0..2); *j3 is overwritten when the solver pins it.
TCMP_HD TCMP_OUTLINE inline int solve_elbow_singular(const Pose &P, Root *j3p, const Root &j5, double q0, double U,
                                                     double W, Root j4c[2], Emit &out) {
    const Root j3 = *j3p;
    // second / third guard level (:2415-2417, :2427-2429): j4eval[1] = U q0 and -11.36.. W q0 (expanded in the
    // generated code); their else-branches (:2853-3030, alternative atan2 forms of j4) need K == 0 with q0 != 0.
    if (!(fabs(q0) < kBranchThresh || fabs(U * q0) < kBranchThresh) ||
        !(fabs(q0) < kBranchThresh || fabs(11.3636363636364 * W * q0) < kBranchThresh)) {
        out.status |= kStatusUnresolved;
        return 0;
    }
    // -W in the association order of :2445: ((-npz sj5) + ((cj5 npy) sj6) + (0.088 cj5)) + ((-cj5 cj6) npx)
    const double Wn = xadd(xadd(xadd(xmul(-P.npz, j5.s), xmul(xmul(j5.c, P.npy), P.s6)), xmul(0.088, j5.c)),
                           xmul(xmul(-j5.c, P.c6), P.npx));
    const double near0 = -3.14159265358979 + pos_fmod(3.14159265358979 + fabs(j3.a), 6.28318530717959);
    const double nearS = -3.14159265358979 +
                         pos_fmod(3.14159265358979 + fabs(-2.63084142381503 + j3.a), 6.28318530717959);
    if (fabs(near0) < kAngleThresh) {
        // j3 ~ 0 (:2436-2598): the solver pins j3 = 0 (sin 0, cos 1).  A consistent pose has U = W = 0 here and the
        // solver gives up ("no branches": j4 and j2 turn about the same line); only a pose whose (U, W) residue lies
        // between 1e-6 and 1e-5 gets the two opposite j4 roots of atan2(U, -W).
        if (fabs(U) + fabs(Wn) < kBranchThresh) return 0;
        double at;
        if (!atan2_checked(U, Wn, &at)) return 0;
        *j3p = Root{0.0, 0.0, 1.0};
        const Root j4r[2] = {make_root(-at), make_root(3.14159265358979 - at)};
        const bool j4ok[2] = {true, !same_root(j4r[0], j4r[1])};
        if (near_threshold(fabs(U) + fabs(Wn), kBranchThresh)) out.status |= kStatusIllConditioned;
        int n4 = 0;
        for (int i4 = 0; i4 < 2; ++i4) {
            if (!j4ok[i4]) continue;
            if (fabs(j4r[i4].c * Wn - j4r[i4].s * U) > kEvalThresh) continue;   // (:2583)
            j4c[n4++] = j4r[i4];
        }
        return n4;
    }
    if (fabs(nearS) < kAngleThresh) {
        // j3 ~ 2.63084 (:2774-2835): every j4 solves the position equations and the spherical shoulder absorbs the
        // rotation about the forearm axis -- a one-parameter family.  The solver evaluates its general formula with
        // K frozen at -3.85e-10 (the value of its rounded sin / cos literals), i.e. j4 = atan2(-U, -W) of whatever
        // residue the pose carries, and keeps that single member of the family.
        // (:2782-2786) with x1017 = 2597402597.4026 sj6, x1018 = 2597402597.4026 cj6 -- same association order, so the
        // host build reproduces the reference's j4 (a quotient of rounding residues when K is 0 to the last bit)
        const double x1017 = xmul(2597402597.4026, P.s6), x1018 = xmul(2597402597.4026, P.c6);
        const double y = xadd(xmul(-P.npy, x1018), xmul(-P.npx, x1017));
        const double x = xadd(xadd(xadd(xmul(228571428.571429, j5.c), xmul(xmul(-j5.c, P.npx), x1018)),
                                   xmul(xmul(-2597402597.4026, P.npz), j5.s)),
                              xmul(xmul(j5.c, P.npy), x1017));
        if (fabs(y) < kAtan2Thresh && fabs(x) < kAtan2Thresh && fabs(y * y + x * x - 1) <= kSinCosThresh) return 0;
        const Root j4 = make_root(isnan(y) ? kPi2 : (isnan(x) ? 0.0 : atan2(y, x)));
        const double e0 = -U - 3.85e-10 * j4.s, e1 = Wn - 3.85e-10 * j4.c, e2 = j4.s * Wn + j4.c * U,
                     e3 = -3.85e-10 + j4.c * Wn - j4.s * U;
        if (fabs(e0) > kEvalThresh || fabs(e1) > kEvalThresh || fabs(e2) > kEvalThresh || fabs(e3) > kEvalThresh) return 0;
        j4c[0] = j4;
        return 1;
    }
    return 0;   // "branch miss [j4]" (:2849): K ~ 0 away from both special angles cannot happen for real j3
}

// ---- one solve (IKSolver::ComputeIk, :412), split so that a kernel can re-balance its lanes between the stages -----
//
//   plan_roots   everything that precedes the per-root work: j3 from |p|^2 (two roots of one asin, duplicate test,
//                :461-501), the root-independent guards and quantities of the j5 formula (:503-508, :2351-2360), and for
//                each j3 root whether it survives to the j5 roots (asin argument in range, :2361-2363).
//   solve_root   one j3 root: j5 (two roots), j4, the residual shoulder problem -- up to 4 solutions.
//
// A solve is plan_roots + solve_root for every live root, in root order (solve_one_t).  ~45 % of a sweep's solves have no
// live root, ~29 % one, ~27 % two: a warp that runs 32 solves through one loop idles a third of its lanes, so the batch
// kernel sorts planned solves into a one-root and a two-root queue and runs solve_root on homogeneous batches.
struct RootPlan {
    Root j3[2];
    bool live[2];      // the root reaches the j5 roots
    double cn, sn;     // x78 = cj6 npx, x79 = npy sj6
    double x975, at5;  // 0.088 - cn + sn;  atan2(npz, x975)
    double hinv;       // 1 / |(x975, npz)|
};

// Returns the number of live roots (0: the solve is finished, out.status says why when it is not a plain "no solution").
TCMP_HD inline int plan_roots(const Pose &P, RootPlan &pl, Emit &out) {
    pl.live[0] = pl.live[1] = false;
    pl.cn = xmul(P.c6, P.npx);
    pl.sn = xmul(P.npy, P.s6);
    // j3 from |p|^2 (:461-485)
    const double arg3 = j3_asin_argument(P);
    if (!(arg3 == arg3)) {   // NaN anywhere in the pose / free value ends up here
        out.status |= kStatusInvalid;
        return 0;
    }
    if (!in_unit(arg3)) return 0;
    const double a3 = clamp_asin(arg3);
    pl.j3[0] = make_root(1.10379390314189 + a3);
    pl.j3[1] = make_root(4.24538655673168 - a3);
    const bool second = !same_root(pl.j3[0], pl.j3[1]);
    if (same_root_borderline(pl.j3[0], pl.j3[1])) out.status |= kStatusIllConditioned;
    // branch guards for the j5 formula (:503-508); none of this depends on the j3 root
    const double cn = pl.cn, sn = pl.sn;
    const double g0 = 1.0 + 129.132231404959 * (cn * cn) + 22.7272727272727 * sn + 129.132231404959 * (sn * sn) +
                      (-258.264462809917) * cn * sn + 129.132231404959 * (P.npz * P.npz) + (-22.7272727272727) * cn;
    pl.x975 = xadd(xsub(0.088, cn), sn);
    const double g1 = fabs(pl.x975) + fabs(P.npz);
    if (fabs(g0) < kBranchThresh || fabs(g1) < kBranchThresh) {
        // Shoulder centre within 8.8e-5 m of the joint-6 axis (g0 = 129.13 h^2, h the distance).  The generated
        // sub-tree (:509-2346) solves j4 from asin(U / K) first and j5 from a quotient by h^2, but each of its j5
        // formulas is guarded by the same g0 (times 1, sin j4 or cos j4) and ends in "no branches", and its doubly
        // singular part (:516-1575, j3 ~ 2.63084 as well) needs |2.6e9 U| <= 1 where |U| = 0.068 on this axis.
        // The geometry agrees: 0.384 + 0.316 cos j3 - 0.0825 sin j3 >= 0.057 > h, no configuration puts the
        // shoulder there.  So: 0 solutions, resolved (the compiled reference returns 0 on every such pose of
        // tests/ik_families.py:wrist_axis_family).
        out.status |= kStatusDegenerate;
        return 0;
    }
    // j5: atan2 + asin (:2347-2364)
    if (!atan2_checked(P.npz, pl.x975, &pl.at5)) return 0;
    const double h2 = xadd(xmul(pl.x975, pl.x975), xmul(P.npz, P.npz));
    if (h2 < -0.00001) return 0;
    const double h = fabs(h2 <= 0.0 ? 0.0 : sqrt(h2));   // IKabs(IKsqrt(.)) (:183)
    if (h == 0.0) return 0;                              // IKPowWithIntegerCheck(.,-1) (:269)
    pl.hinv = 1.0 / h;
    int n_live = 0;
    for (int r = 0; r < 2; ++r) {
        if (r == 1 && !second) break;
        const double arg5 = xmul(pl.hinv, xadd(xadd(0.384, xmul(-0.0825, pl.j3[r].s)), xmul(0.316, pl.j3[r].c)));
        pl.live[r] = in_unit(arg5);
        n_live += pl.live[r];
    }
    return n_live;
}

// One live j3 root (:2365-3097 + rotationfunction0).  WITH_ELBOW = false is the kernels' hot path: it carries everything
// but the elbow-singularity branches (1 solve in ~10^5 of a random sweep, every solve of a pose with joint 4 at 2.63084
// or 0) and bails out with kStatusRedo when it meets one; inlining those branches costs the hot path 36 registers.
template <bool WITH_ELBOW>
TCMP_HD inline void solve_root(const Pose &P, const RootPlan &pl, const Root &j3, Emit &out) {
    const double arg5 = xmul(pl.hinv, xadd(xadd(0.384, xmul(-0.0825, j3.s)), xmul(0.316, j3.c)));
    const double a5 = clamp_asin(arg5);
    const double at5 = pl.at5;
    Root j5r[2] = {make_root(-a5 - at5), make_root(3.14159265358979 + a5 - at5)};
    bool j5ok[2] = {true, !same_root(j5r[0], j5r[1])};
    if (same_root_borderline(j5r[0], j5r[1])) out.status |= kStatusIllConditioned;

TCMP_ROOT_LOOP
    for (int i5 = 0; i5 < 2; ++i5) {
        if (!j5ok[i5]) continue;
        const Root &j5 = j5r[i5];
        // j4: one root (:2404-2408 guards, :3037-3045 formula)
        const double K = xadd(xadd(-0.0825, xmul(0.0825, j3.c)), xmul(0.316, j3.s));
        const double U = xadd(xmul(P.c6, P.npy), xmul(P.npx, P.s6));
        // (-0.088 cj5) + ((-cj5 npy) sj6) + (npz sj5) + ((cj5 cj6) npx)   (:2406, :3037)
        const double W = xadd(xadd(xadd(xmul(-0.088, j5.c), xmul(xmul(-j5.c, P.npy), P.s6)), xmul(P.npz, j5.s)),
                              xmul(xmul(j5.c, P.c6), P.npx));
        const double q0 = xadd(xadd(-1.0, j3.c), xmul(3.83030303030303, j3.s));
        const double sK = sign_of(K);
        Root j3u = j3, j4c[2];
        int n4;
        // Next to the elbow singularity j4 = atan2(U, W) with |(U, W)| = |K| -> 0 inherits the rounding residue of j5
        // (itself next to a double root there: 1 - arg5 ~ K^2 / 2C^2) amplified by C^2 / K^2: below |K| = 2e-5 (joint 4
        // within 6e-5 rad of 2.63084 or 0) that reaches 1e-8 rad and more, enough to move the shoulder's duplicate-root
        // and singular-branch tests -- seen on the GPU as one count mismatch in 12 M structured solves.
        if (near_threshold(fabs(q0), kBranchThresh) || near_threshold(fabs(U) + fabs(W), kBranchThresh) || fabs(K) < 2e-5)
            out.status |= kStatusIllConditioned;
        if (fabs(q0) < kBranchThresh || fabs(U) + fabs(W) < kBranchThresh || fabs(sK) < kBranchThresh) {
            // Elbow singularity: K = 0, the shoulder centre lies on the forearm (joint-5) axis, so the position
            // equations say nothing about j4 (U^2 + W^2 = K^2).
            out.status |= kStatusDegenerate;
            if (!WITH_ELBOW) {
                out.status |= kStatusRedo;
                return;
            }
            n4 = solve_elbow_singular(P, &j3u, j5, q0, U, W, j4c, out);
        } else {
            double at4;
            if (!atan2_checked(U, W, &at4)) continue;
            const Root j4 = make_root(-kHalfPiLit + at4 + kHalfPiLit * (1.0 / sK));
            // 4 residuals (:3087-3090)
            const double e0 = j4.s * K - U, e1 = j4.c * K - W, e2 = j4.c * U - j4.s * W,
                         e3 = K - j4.c * W - j4.s * U;
            if (fabs(e0) > kEvalThresh || fabs(e1) > kEvalThresh || fabs(e2) > kEvalThresh || fabs(e3) > kEvalThresh)
                continue;
            j4c[0] = j4;
            n4 = 1;
        }
        // one call site: solve_shoulder is by far the largest inlined body of the kernel
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
        for (int i4 = 0; i4 < n4; ++i4) solve_shoulder(P, j3u, j4c[i4], j5, out);
    }
}

// The whole solve on one lane: plan, then the live roots in root order.
template <bool WITH_ELBOW>
TCMP_HD inline void solve_one_t(const Pose &P, Emit &out) {
    RootPlan pl;
    if (plan_roots(P, pl, out) == 0) return;
TCMP_ROOT_LOOP
    for (int i3 = 0; i3 < 2; ++i3) {
        if (!pl.live[i3]) continue;
        solve_root<WITH_ELBOW>(P, pl, pl.j3[i3], out);
        if (!WITH_ELBOW && (out.status & kStatusRedo)) return;
    }
}

TCMP_HD inline void solve_one(const Pose &P, Emit &out) { solve_one_t<true>(P, out); }

}  // namespace ik
}  // namespace tcmp
