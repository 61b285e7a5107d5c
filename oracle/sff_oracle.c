/*
 * sff_oracle.c -- CPU ORACLE for the collision + neighbour hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (space_filling_forest_star_b200/csrc) never links, imports or calls anything in oracle/.
 *
 * PARITY STATUS
 *   collision  : "parity unpinned".  The reference delegates all collision arithmetic to RAPID 2.01
 *                (UNC GAMMA group), which is NOT vendored (lib/rapid-2.01/ holds only a README) and the
 *                reference ships no golden verdicts.  This file restates RAPID's published algorithm
 *                (Gottschalk/Lin/Manocha, "OBBTree", SIGGRAPH'96) behind the reference's own call
 *                contract (src/environment.h:269-276, src/primitives.h:252-262).  What IS pinned,
 *                independently of any restatement, is the meaning of the verdict: on integer-coordinate
 *                triangles (all quantities exact) orc_tri_contact equals exact closed-triangle
 *                intersection computed by a different algorithm (tests/test_oracle_exact_geometry.py).
 *   k-NN/radius: pinned against vendored FLANN 1.9.1 LinearIndex compiled from /root/reference
 *                (oracle/_ref, see oracle/Makefile + oracle/ref_flann.cpp) by tests/test_oracle_knn.py.
 *
 * All arithmetic is IEEE double (collision) / float (k-NN metric) with contraction disabled
 * (-ffp-contract=off) so that results are a pure function of the inputs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * 1. Pose -> rotation.  Follows Point<T>::FillRotationMatrix, src/primitives.h:252-262
 *    (R = Rz(yaw) * Ry(pitch) * Rx(roll), row-major m[r][c]).  pose = x y z yaw pitch roll.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_rotation(const double pose[6], double m[3][3]) {
    const double yaw = pose[3], pitch = pose[4], roll = pose[5];
    m[0][0] = cos(yaw) * cos(pitch);
    m[0][1] = cos(yaw) * sin(pitch) * sin(roll) - sin(yaw) * cos(roll);
    m[0][2] = cos(yaw) * sin(pitch) * cos(roll) + sin(yaw) * sin(roll);
    m[1][0] = sin(yaw) * cos(pitch);
    m[1][1] = sin(yaw) * sin(pitch) * sin(roll) + cos(yaw) * cos(roll);
    m[1][2] = sin(yaw) * sin(pitch) * cos(roll) - cos(yaw) * sin(roll);
    m[2][0] = -sin(pitch);
    m[2][1] = cos(pitch) * sin(roll);
    m[2][2] = cos(pitch) * cos(roll);
}

/* ------------------------------------------------------------------------------------------------
 * 2. Triangle/triangle contact: 17-axis separating-axis test.
 *    Strict comparison: disjoint on an axis iff mn1 > mx2 or mn2 > mx1; touching == contact.
 * ---------------------------------------------------------------------------------------------- */
static inline void v_sub(double r[3], const double a[3], const double b[3]) {
    r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2];
}
static inline void v_cross(double r[3], const double a[3], const double b[3]) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double v_dot(const double a[3], const double b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline double max3(double a, double b, double c) { double t = a; if (b > t) t = b; if (c > t) t = c; return t; }
static inline double min3(double a, double b, double c) { double t = a; if (b < t) t = b; if (c < t) t = c; return t; }

/* returns 1 when the two projected intervals overlap (axis does NOT separate) */
static inline int axis_overlaps(const double ax[3], const double p1[3], const double p2[3], const double p3[3],
                                const double q1[3], const double q2[3], const double q3[3]) {
    double P1 = v_dot(ax, p1), P2 = v_dot(ax, p2), P3 = v_dot(ax, p3);
    double Q1 = v_dot(ax, q1), Q2 = v_dot(ax, q2), Q3 = v_dot(ax, q3);
    double mx1 = max3(P1, P2, P3), mn1 = min3(P1, P2, P3);
    double mx2 = max3(Q1, Q2, Q3), mn2 = min3(Q1, Q2, Q3);
    if (mn1 > mx2) return 0;
    if (mn2 > mx1) return 0;
    return 1;
}

/* P* = first triangle already expressed in the frame of the second triangle Q*. */
ORC_API int orc_tri_contact(const double P1[3], const double P2[3], const double P3[3],
                            const double Q1[3], const double Q2[3], const double Q3[3]) {
    double p1[3], p2[3], p3[3], q1[3], q2[3], q3[3];
    double e1[3], e2[3], e3[3], f1[3], f2[3], f3[3];
    double n1[3], m1[3], ax[3];
    /* everything relative to P1 */
    v_sub(p1, P1, P1); v_sub(p2, P2, P1); v_sub(p3, P3, P1);
    v_sub(q1, Q1, P1); v_sub(q2, Q2, P1); v_sub(q3, Q3, P1);
    v_sub(e1, p2, p1); v_sub(e2, p3, p2); v_sub(e3, p1, p3);
    v_sub(f1, q2, q1); v_sub(f2, q3, q2); v_sub(f3, q1, q3);
    v_cross(n1, e1, e2);
    v_cross(m1, f1, f2);
    /* the two face normals */
    if (!axis_overlaps(n1, p1, p2, p3, q1, q2, q3)) return 0;
    if (!axis_overlaps(m1, p1, p2, p3, q1, q2, q3)) return 0;
    /* nine edge x edge axes */
    const double *E[3] = {e1, e2, e3};
    const double *F[3] = {f1, f2, f3};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            v_cross(ax, E[i], F[j]);
            if (!axis_overlaps(ax, p1, p2, p3, q1, q2, q3)) return 0;
        }
    /* in-plane edge normals of both triangles */
    for (int i = 0; i < 3; ++i) {
        v_cross(ax, E[i], n1);
        if (!axis_overlaps(ax, p1, p2, p3, q1, q2, q3)) return 0;
    }
    for (int j = 0; j < 3; ++j) {
        v_cross(ax, F[j], m1);
        if (!axis_overlaps(ax, p1, p2, p3, q1, q2, q3)) return 0;
    }
    return 1;
}

/* Same 17 axes, but returns the largest normalised gap (>0: separated by at least that distance
 * along some axis; <=0: no axis separates).  Diagnostic used to enumerate near-contact poses. */
static double tri_pair_margin(const double P1[3], const double P2[3], const double P3[3],
                              const double Q1[3], const double Q2[3], const double Q3[3]) {
    double p1[3], p2[3], p3[3], q1[3], q2[3], q3[3];
    double e[3][3], f[3][3], n1[3], m1[3], ax[17][3];
    v_sub(p1, P1, P1); v_sub(p2, P2, P1); v_sub(p3, P3, P1);
    v_sub(q1, Q1, P1); v_sub(q2, Q2, P1); v_sub(q3, Q3, P1);
    v_sub(e[0], p2, p1); v_sub(e[1], p3, p2); v_sub(e[2], p1, p3);
    v_sub(f[0], q2, q1); v_sub(f[1], q3, q2); v_sub(f[2], q1, q3);
    v_cross(n1, e[0], e[1]); v_cross(m1, f[0], f[1]);
    memcpy(ax[0], n1, sizeof n1); memcpy(ax[1], m1, sizeof m1);
    int k = 2;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v_cross(ax[k++], e[i], f[j]);
    for (int i = 0; i < 3; ++i) v_cross(ax[k++], e[i], n1);
    for (int j = 0; j < 3; ++j) v_cross(ax[k++], f[j], m1);
    double best = -INFINITY;
    for (k = 0; k < 17; ++k) {
        double len = sqrt(v_dot(ax[k], ax[k]));
        if (len == 0.0) continue;
        double P[3] = {v_dot(ax[k], p1), v_dot(ax[k], p2), v_dot(ax[k], p3)};
        double Q[3] = {v_dot(ax[k], q1), v_dot(ax[k], q2), v_dot(ax[k], q3)};
        double g1 = min3(P[0], P[1], P[2]) - max3(Q[0], Q[1], Q[2]);
        double g2 = min3(Q[0], Q[1], Q[2]) - max3(P[0], P[1], P[2]);
        double g = (g1 > g2 ? g1 : g2) / len;
        if (g > best) best = g;
    }
    return best;
}

/* ------------------------------------------------------------------------------------------------
 * 3. Model-1 (obstacle, identity placement) -> model-2 (robot at pose) transform, A.2 step 1:
 *    mR = R2^T * R1 (R1 = I), mT = R2^T * (T1 - T2) (T1 = 0); vertex i = mR*p + mT.
 *    Call contract: Obstacle<T>::Collide(object, robot, robPos), src/environment.h:269-276.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { double mR[3][3]; double mT[3]; double R2[3][3]; double T2[3]; } orc_xform;

static void finish_xform(orc_xform *x);
static void make_xform(const double pose[6], orc_xform *x) {
    orc_rotation(pose, x->R2);
    x->T2[0] = pose[0]; x->T2[1] = pose[1]; x->T2[2] = pose[2];
    finish_xform(x);
}
/* x->R2, x->T2 given -> mR, mT */
static void finish_xform(orc_xform *x) {
    static const double R1[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    const double T1[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            x->mR[i][j] = x->R2[0][i] * R1[0][j] + x->R2[1][i] * R1[1][j] + x->R2[2][i] * R1[2][j];
    double u[3]; v_sub(u, T1, x->T2);
    for (int i = 0; i < 3; ++i)
        x->mT[i] = x->R2[0][i] * u[0] + x->R2[1][i] * u[1] + x->R2[2][i] * u[2];
}
static inline void xform_point(const orc_xform *x, const double p[3], double r[3]) {
    for (int i = 0; i < 3; ++i)
        r[i] = 1.0 * (x->mR[i][0] * p[0] + x->mR[i][1] * p[1] + x->mR[i][2] * p[2]) + x->mT[i];
}

/* Ground truth verdict: OR over all (obstacle triangle, robot triangle) pairs (A.2 step 5).
 * tris are [n][9] doubles (p1 p2 p3). */
ORC_API int orc_collide_brute(const double *obst, int nT, const double *robot, int nR, const double pose[6]) {
    orc_xform x; make_xform(pose, &x);
    for (int t = 0; t < nT; ++t) {
        double i1[3], i2[3], i3[3];
        xform_point(&x, obst + 9 * t, i1);
        xform_point(&x, obst + 9 * t + 3, i2);
        xform_point(&x, obst + 9 * t + 6, i3);
        for (int r = 0; r < nR; ++r)
            if (orc_tri_contact(i1, i2, i3, robot + 9 * r, robot + 9 * r + 3, robot + 9 * r + 6)) return 1;
    }
    return 0;
}

/* min over pairs of the per-pair margin: > 0 free with at least that clearance along an axis; <= 0 contact */
ORC_API double orc_pose_margin(const double *obst, int nT, const double *robot, int nR, const double pose[6]) {
    orc_xform x; make_xform(pose, &x);
    double worst = INFINITY;
    for (int t = 0; t < nT; ++t) {
        double i1[3], i2[3], i3[3];
        xform_point(&x, obst + 9 * t, i1);
        xform_point(&x, obst + 9 * t + 3, i2);
        xform_point(&x, obst + 9 * t + 6, i3);
        for (int r = 0; r < nR; ++r) {
            double m = tri_pair_margin(i1, i2, i3, robot + 9 * r, robot + 9 * r + 3, robot + 9 * r + 6);
            if (m < worst) worst = m;
        }
    }
    return worst;
}

ORC_API void orc_collide_brute_batch(const double *obst, int nT, const double *robot, int nR,
                                     const double *poses, int64_t n, uint8_t *out, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) out[i] = (uint8_t)orc_collide_brute(obst, nT, robot, nR, poses + 6 * i);
}

ORC_API void orc_pose_margin_batch(const double *obst, int nT, const double *robot, int nR,
                                   const double *poses, int64_t n, double *out, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) out[i] = orc_pose_margin(obst, nT, robot, nR, poses + 6 * i);
}

/* ------------------------------------------------------------------------------------------------
 * 4. OBB-tree model (RAPID restatement, Appendix A.3-A.5).  Used for CPU *timing* and for the
 *    n_box / n_tri work counters; verdicts must equal orc_collide_brute (checked by tests).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    double pR[3][3];   /* orientation relative to parent box (root: model frame) */
    double pT[3];      /* centre relative to parent box */
    double d[3];       /* half extents */
    int P, N;          /* children (index into boxes), -1 for leaf */
    int tri;           /* triangle index for a leaf, else -1 */
} orc_box;

typedef struct { double A; double m[3]; double s[3][3]; } orc_moment;
typedef struct { double A; double m[3]; double s[3][3]; } orc_accum;

typedef struct {
    int n_tris;
    double *tris;      /* [n][9] */
    orc_box *boxes;    /* 2n-1 used */
    int n_boxes;
    orc_moment *mom;   /* build scratch */
} orc_model;

static void m_identity(double m[3][3]) { memset(m, 0, 9 * sizeof(double)); m[0][0] = m[1][1] = m[2][2] = 1.0; }
static void m_copy(double d[3][3], const double s[3][3]) { memcpy(d, s, 9 * sizeof(double)); }
/* r = a^T * b */
static void mt_x_m(double r[3][3], const double a[3][3], const double b[3][3]) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        r[i][j] = a[0][i] * b[0][j] + a[1][i] * b[1][j] + a[2][i] * b[2][j];
}
/* r = a * b */
static void m_x_m(double r[3][3], const double a[3][3], const double b[3][3]) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        r[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
}
/* r = a^T * v */
static void mt_x_v(double r[3], const double a[3][3], const double v[3]) {
    for (int i = 0; i < 3; ++i) r[i] = a[0][i] * v[0] + a[1][i] * v[1] + a[2][i] * v[2];
}
/* r = a * v + t */
static void m_x_v_p_v(double r[3], const double a[3][3], const double v[3], const double t[3]) {
    for (int i = 0; i < 3; ++i) r[i] = (a[i][0] * v[0] + a[i][1] * v[1] + a[i][2] * v[2]) + t[i];
}

static void moment_of_tri(orc_moment *M, const double *p, const double *q, const double *r) {
    double u[3], v[3], w[3];
    v_sub(u, q, p); v_sub(v, r, p); v_cross(w, u, v);
    M->A = 0.5 * sqrt(v_dot(w, w));
    for (int i = 0; i < 3; ++i) M->m[i] = (p[i] + q[i] + r[i]) / 3;
    if (M->A == 0.0) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
            M->s[i][j] = p[i] * p[j] + q[i] * q[j] + r[i] * r[j];
        return;
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        M->s[i][j] = M->A * (9 * M->m[i] * M->m[j] + p[i] * p[j] + q[i] * q[j] + r[i] * r[j]) / 12;
}
static void accum_clear(orc_accum *a) { memset(a, 0, sizeof *a); }
static void accum_add(orc_accum *a, const orc_moment *b) {
    for (int i = 0; i < 3; ++i) a->m[i] += b->m[i] * b->A;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a->s[i][j] += b->s[i][j];
    a->A += b->A;
}
static void accum_mean(double m[3], const orc_accum *a) { for (int i = 0; i < 3; ++i) m[i] = a->m[i] / a->A; }
static void accum_cov(double C[3][3], const orc_accum *a) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i][j] = a->s[i][j] - a->m[i] * a->m[j] / a->A;
}

/* cyclic Jacobi eigen-solver for a symmetric 3x3; returns the sweep count (50 = did not converge) */
static int jacobi3(double vout[3][3], double dout[3], double a_in[3][3]) {
    double a[3][3], v[3][3], b[3], z[3], d[3];
    m_copy(a, a_in); m_identity(v);
    for (int ip = 0; ip < 3; ++ip) { b[ip] = d[ip] = a[ip][ip]; z[ip] = 0.0; }
    for (int sweep = 0; sweep < 50; ++sweep) {
        double sm = 0.0;
        for (int ip = 0; ip < 3; ++ip) for (int iq = ip + 1; iq < 3; ++iq) sm += fabs(a[ip][iq]);
        if (sm == 0.0) { m_copy(vout, v); memcpy(dout, d, sizeof d); return sweep; }
        double tresh = (sweep < 3) ? 0.2 * sm / 9.0 : 0.0;
        for (int ip = 0; ip < 3; ++ip) for (int iq = ip + 1; iq < 3; ++iq) {
            double g = 100.0 * fabs(a[ip][iq]);
            if (sweep > 3 && fabs(d[ip]) + g == fabs(d[ip]) && fabs(d[iq]) + g == fabs(d[iq])) {
                a[ip][iq] = 0.0;
            } else if (fabs(a[ip][iq]) > tresh) {
                double h = d[iq] - d[ip], t;
                if (fabs(h) + g == fabs(h)) t = a[ip][iq] / h;
                else {
                    double theta = 0.5 * h / a[ip][iq];
                    t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
                    if (theta < 0.0) t = -t;
                }
                double c = 1.0 / sqrt(1 + t * t), s = t * c, tau = s / (1.0 + c);
                h = t * a[ip][iq];
                z[ip] -= h; z[iq] += h; d[ip] -= h; d[iq] += h; a[ip][iq] = 0.0;
#define ORC_ROT(M, i, j, k, l) { double g_ = M[i][j], h_ = M[k][l]; M[i][j] = g_ - s * (h_ + g_ * tau); M[k][l] = h_ + s * (g_ - h_ * tau); }
                for (int j = 0; j < ip; ++j) ORC_ROT(a, j, ip, j, iq)
                for (int j = ip + 1; j < iq; ++j) ORC_ROT(a, ip, j, j, iq)
                for (int j = iq + 1; j < 3; ++j) ORC_ROT(a, ip, j, iq, j)
                for (int j = 0; j < 3; ++j) ORC_ROT(v, j, ip, j, iq)
#undef ORC_ROT
            }
        }
        for (int ip = 0; ip < 3; ++ip) { b[ip] += z[ip]; d[ip] = b[ip]; z[ip] = 0.0; }
    }
    m_copy(vout, v); memcpy(dout, d, sizeof d);
    return 50;
}
/* eigenvectors as columns, only the largest-eigenvalue column moved to column 0 (A.5) */
static int eigen_largest_first(double evecs[3][3], double cov[3][3]) {
    double evals[3];
    int n = jacobi3(evecs, evals, cov);
    int big = 0;
    if (evals[2] > evals[0]) big = (evals[2] > evals[1]) ? 2 : 1;
    else big = (evals[0] > evals[1]) ? 0 : 1;
    if (big != 0) {
        double t = evals[big]; evals[big] = evals[0]; evals[0] = t;
        for (int r = 0; r < 3; ++r) { t = evecs[r][big]; evecs[r][big] = evecs[r][0]; evecs[r][0] = t; }
    }
    return n;
}

static inline void minmax_upd(double mn[3], double mx[3], const double c[3]) {
    for (int k = 0; k < 3; ++k) { if (c[k] < mn[k]) mn[k] = c[k]; else if (c[k] > mx[k]) mx[k] = c[k]; }
}

static void build_leaf(orc_model *mdl, int bi, int tri) {
    orc_box *b = &mdl->boxes[bi];
    const double *p1 = mdl->tris + 9 * tri, *p2 = p1 + 3, *p3 = p1 + 6;
    double u12[3], u23[3], u31[3], a0[3], a1[3], a2[3];
    b->P = b->N = -1; b->tri = tri;
    v_sub(u12, p1, p2); v_sub(u23, p2, p3); v_sub(u31, p3, p1);
    double d12 = v_dot(u12, u12), d23 = v_dot(u23, u23), d31 = v_dot(u31, u31), l;
    const double *longest;
    if (d12 > d23) { if (d12 > d31) { l = 1.0 / sqrt(d12); longest = u12; } else { l = 1.0 / sqrt(d31); longest = u31; } }
    else           { if (d23 > d31) { l = 1.0 / sqrt(d23); longest = u23; } else { l = 1.0 / sqrt(d31); longest = u31; } }
    for (int k = 0; k < 3; ++k) a0[k] = longest[k] * l;
    v_cross(a2, u12, u23);
    l = 1.0 / sqrt(v_dot(a2, a2));
    for (int k = 0; k < 3; ++k) a2[k] *= l;
    v_cross(a1, a2, a0);
    for (int k = 0; k < 3; ++k) { b->pR[k][0] = a0[k]; b->pR[k][1] = a1[k]; b->pR[k][2] = a2[k]; }
    double mn[3], mx[3], c[3];
    mt_x_v(c, b->pR, p1); memcpy(mn, c, sizeof c); memcpy(mx, c, sizeof c);
    mt_x_v(c, b->pR, p2); minmax_upd(mn, mx, c);
    mt_x_v(c, b->pR, p3); minmax_upd(mn, mx, c);
    for (int k = 0; k < 3; ++k) c[k] = (mn[k] + mx[k]) * 0.5;
    for (int k = 0; k < 3; ++k) b->pT[k] = c[0] * b->pR[k][0] + c[1] * b->pR[k][1] + c[2] * b->pR[k][2];
    for (int k = 0; k < 3; ++k) b->d[k] = (mx[k] - mn[k]) * 0.5;
}

/* box bi has pR (model frame) and pT (= mean point) preset; t[0..n) are its triangles */
static void build_split(orc_model *mdl, int bi, int *t, int n) {
    if (n == 1) { build_leaf(mdl, bi, t[0]); return; }
    orc_box *b = &mdl->boxes[bi];
    orc_accum M1, M2; double C[3][3], c[3], mn[3], mx[3];
    int n1 = 0;
    b->tri = -1;
    double axdmp = b->pR[0][0] * b->pT[0] + b->pR[1][0] * b->pT[1] + b->pR[2][0] * b->pT[2];
    accum_clear(&M1); accum_clear(&M2);
    mt_x_v(c, b->pR, mdl->tris + 9 * t[0]);
    memcpy(mn, c, sizeof c); memcpy(mx, c, sizeof c);
    for (int i = 0; i < n; ++i) {
        int in = t[i];
        const double *p = mdl->tris + 9 * in;
        mt_x_v(c, b->pR, p);     minmax_upd(mn, mx, c);
        mt_x_v(c, b->pR, p + 3); minmax_upd(mn, mx, c);
        mt_x_v(c, b->pR, p + 6); minmax_upd(mn, mx, c);
        const double *mm = mdl->mom[in].m;
        double proj = b->pR[0][0] * mm[0] + b->pR[1][0] * mm[1] + b->pR[2][0] * mm[2];
        if (((proj < axdmp) && (n != 2)) || ((n == 2) && (i == 0))) {
            accum_add(&M1, &mdl->mom[in]);
            int tmp = t[i]; t[i] = t[n1]; t[n1] = tmp; ++n1;
        } else {
            accum_add(&M2, &mdl->mom[in]);
        }
    }
    if (n1 == 0 || n1 == n) {   /* degenerate partition: split in the middle and re-accumulate */
        n1 = n / 2;
        accum_clear(&M1); for (int i = 0; i < n1; ++i) accum_add(&M1, &mdl->mom[t[i]]);
        accum_clear(&M2); for (int i = n1; i < n; ++i) accum_add(&M2, &mdl->mom[t[i]]);
    }
    for (int k = 0; k < 3; ++k) c[k] = (mn[k] + mx[k]) * 0.5;
    for (int k = 0; k < 3; ++k) b->pT[k] = c[0] * b->pR[k][0] + c[1] * b->pR[k][1] + c[2] * b->pR[k][2];
    for (int k = 0; k < 3; ++k) b->d[k] = (mx[k] - mn[k]) * 0.5;

    int Pi = mdl->n_boxes++, Ni = mdl->n_boxes++;
    b->P = Pi; b->N = Ni;
    double tR[3][3];
    for (int side = 0; side < 2; ++side) {
        int ci = side == 0 ? Pi : Ni;
        int cn = side == 0 ? n1 : n - n1;
        int *ct = side == 0 ? t : t + n1;
        orc_accum *M = side == 0 ? &M1 : &M2;
        orc_box *cb = &mdl->boxes[ci];
        if (cn > 1) {
            accum_mean(cb->pT, M);
            accum_cov(C, M);
            if (eigen_largest_first(tR, C) > 30) m_identity(tR);
            m_copy(cb->pR, tR);
            build_split(mdl, ci, ct, cn);
        } else {
            build_leaf(mdl, ci, ct[0]);
        }
        /* re-express the child relative to this (parent) box */
        b = &mdl->boxes[bi]; cb = &mdl->boxes[ci];
        m_copy(C, cb->pR); mt_x_m(cb->pR, b->pR, C);
        v_sub(c, cb->pT, b->pT); mt_x_v(cb->pT, b->pR, c);
    }
}

ORC_API orc_model *orc_model_build(const double *tris, int n) {
    orc_model *mdl = (orc_model *)calloc(1, sizeof *mdl);
    mdl->n_tris = n;
    mdl->tris = (double *)malloc(sizeof(double) * 9 * (size_t)n);
    memcpy(mdl->tris, tris, sizeof(double) * 9 * (size_t)n);
    mdl->boxes = (orc_box *)calloc((size_t)2 * n, sizeof(orc_box));
    mdl->mom = (orc_moment *)calloc((size_t)n, sizeof(orc_moment));
    double Amin = 0.0; int zero = 0;
    for (int i = 0; i < n; ++i) {
        moment_of_tri(&mdl->mom[i], tris + 9 * i, tris + 9 * i + 3, tris + 9 * i + 6);
        if (mdl->mom[i].A == 0.0) zero = 1;
        else if (Amin == 0.0 || mdl->mom[i].A < Amin) Amin = mdl->mom[i].A;
    }
    if (zero) { if (Amin == 0.0) Amin = 1.0; for (int i = 0; i < n; ++i) if (mdl->mom[i].A == 0.0) mdl->mom[i].A = Amin; }
    orc_accum M; double C[3][3]; accum_clear(&M);
    for (int i = 0; i < n; ++i) accum_add(&M, &mdl->mom[i]);
    accum_mean(mdl->boxes[0].pT, &M);
    accum_cov(C, &M);
    eigen_largest_first(mdl->boxes[0].pR, C);
    int *t = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) t[i] = i;
    mdl->n_boxes = 1;
    build_split(mdl, 0, t, n);
    free(t); free(mdl->mom); mdl->mom = NULL;
    return mdl;
}
ORC_API void orc_model_free(orc_model *m) { if (!m) return; free(m->tris); free(m->boxes); free(m); }
ORC_API int orc_model_num_boxes(const orc_model *m) { return m->n_boxes; }

/* A.3: 15-axis OBB/OBB separating axis test with the 1e-6 inflation of |B|. returns nonzero if disjoint */
static int obb_disjoint(const double B[3][3], const double T[3], const double a[3], const double b[3]) {
    double Bf[3][3], t, s;
    const double reps = 1e-6;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Bf[i][j] = fabs(B[i][j]) + reps;
    t = fabs(T[0]); if (!(t <= (a[0] + b[0] * Bf[0][0] + b[1] * Bf[0][1] + b[2] * Bf[0][2]))) return 1;
    s = T[0] * B[0][0] + T[1] * B[1][0] + T[2] * B[2][0]; t = fabs(s);
    if (!(t <= (b[0] + a[0] * Bf[0][0] + a[1] * Bf[1][0] + a[2] * Bf[2][0]))) return 2;
    t = fabs(T[1]); if (!(t <= (a[1] + b[0] * Bf[1][0] + b[1] * Bf[1][1] + b[2] * Bf[1][2]))) return 3;
    t = fabs(T[2]); if (!(t <= (a[2] + b[0] * Bf[2][0] + b[1] * Bf[2][1] + b[2] * Bf[2][2]))) return 4;
    s = T[0] * B[0][1] + T[1] * B[1][1] + T[2] * B[2][1]; t = fabs(s);
    if (!(t <= (b[1] + a[0] * Bf[0][1] + a[1] * Bf[1][1] + a[2] * Bf[2][1]))) return 5;
    s = T[0] * B[0][2] + T[1] * B[1][2] + T[2] * B[2][2]; t = fabs(s);
    if (!(t <= (b[2] + a[0] * Bf[0][2] + a[1] * Bf[1][2] + a[2] * Bf[2][2]))) return 6;
    s = T[2] * B[1][0] - T[1] * B[2][0]; t = fabs(s);
    if (!(t <= (a[1] * Bf[2][0] + a[2] * Bf[1][0] + b[1] * Bf[0][2] + b[2] * Bf[0][1]))) return 7;
    s = T[2] * B[1][1] - T[1] * B[2][1]; t = fabs(s);
    if (!(t <= (a[1] * Bf[2][1] + a[2] * Bf[1][1] + b[0] * Bf[0][2] + b[2] * Bf[0][0]))) return 8;
    s = T[2] * B[1][2] - T[1] * B[2][2]; t = fabs(s);
    if (!(t <= (a[1] * Bf[2][2] + a[2] * Bf[1][2] + b[0] * Bf[0][1] + b[1] * Bf[0][0]))) return 9;
    s = T[0] * B[2][0] - T[2] * B[0][0]; t = fabs(s);
    if (!(t <= (a[0] * Bf[2][0] + a[2] * Bf[0][0] + b[1] * Bf[1][2] + b[2] * Bf[1][1]))) return 10;
    s = T[0] * B[2][1] - T[2] * B[0][1]; t = fabs(s);
    if (!(t <= (a[0] * Bf[2][1] + a[2] * Bf[0][1] + b[0] * Bf[1][2] + b[2] * Bf[1][0]))) return 11;
    s = T[0] * B[2][2] - T[2] * B[0][2]; t = fabs(s);
    if (!(t <= (a[0] * Bf[2][2] + a[2] * Bf[0][2] + b[0] * Bf[1][1] + b[1] * Bf[1][0]))) return 12;
    s = T[1] * B[0][0] - T[0] * B[1][0]; t = fabs(s);
    if (!(t <= (a[0] * Bf[1][0] + a[1] * Bf[0][0] + b[1] * Bf[2][2] + b[2] * Bf[2][1]))) return 13;
    s = T[1] * B[0][1] - T[0] * B[1][1]; t = fabs(s);
    if (!(t <= (a[0] * Bf[1][1] + a[1] * Bf[0][1] + b[0] * Bf[2][2] + b[2] * Bf[2][0]))) return 14;
    s = T[1] * B[0][2] - T[0] * B[1][2]; t = fabs(s);
    if (!(t <= (a[0] * Bf[1][2] + a[1] * Bf[0][2] + b[0] * Bf[2][1] + b[1] * Bf[2][0]))) return 15;
    return 0;
}

typedef struct {
    const orc_model *m1, *m2;
    const orc_xform *x;
    int first_contact;
    int64_t n_box, n_tri, n_desc, n_contacts;
} orc_query;

static void collide_rec(orc_query *q, int b1i, int b2i, const double R[3][3], const double T[3]) {
    if (q->first_contact && q->n_contacts > 0) return;
    const orc_box *b1 = &q->m1->boxes[b1i], *b2 = &q->m2->boxes[b2i];
    q->n_box++;
    if (obb_disjoint(R, T, b1->d, b2->d)) return;
    int l1 = b1->P < 0, l2 = b2->P < 0;
    if (l1 && l2) {
        double i1[3], i2[3], i3[3];
        const double *p = q->m1->tris + 9 * b1->tri, *r = q->m2->tris + 9 * b2->tri;
        xform_point(q->x, p, i1); xform_point(q->x, p + 3, i2); xform_point(q->x, p + 6, i3);
        q->n_tri++;
        if (orc_tri_contact(i1, i2, i3, r, r + 3, r + 6)) q->n_contacts++;
        return;
    }
    double cR[3][3], cT[3], U[3];
    if (l2 || (!l1 && (b1->d[0] > b2->d[0]))) {
        int kids[2] = {b1->N, b1->P};
        for (int k = 0; k < 2; ++k) {
            const orc_box *c = &q->m1->boxes[kids[k]];
            mt_x_m(cR, c->pR, R); v_sub(U, T, c->pT); mt_x_v(cT, c->pR, U);
            q->n_desc++;
            collide_rec(q, kids[k], b2i, cR, cT);
        }
    } else {
        int kids[2] = {b2->N, b2->P};
        for (int k = 0; k < 2; ++k) {
            const orc_box *c = &q->m2->boxes[kids[k]];
            m_x_m(cR, R, c->pR); m_x_v_p_v(cT, R, c->pT, T);
            q->n_desc++;
            collide_rec(q, b1i, kids[k], cR, cT);
        }
    }
}

static int collide_obbtree_x(const orc_model *obst, const orc_model *robot, const orc_xform *xp, int first_contact,
                             int64_t counters[4]);
/* RAPID_Collide(I, 0, obstacle, R(pose), T(pose), robot, flag).  counters[4] += {box, tri, desc, contacts} */
ORC_API int orc_collide_obbtree(const orc_model *obst, const orc_model *robot, const double pose[6],
                                int first_contact, int64_t counters[4]) {
    orc_xform x; make_xform(pose, &x);
    return collide_obbtree_x(obst, robot, &x, first_contact, counters);
}
/* same, but with RAPID_Collide's own arguments (R2 row-major, T2): what the CPU RAPID.H stand-in forwards */
ORC_API int orc_collide_obbtree_rt(const orc_model *obst, const orc_model *robot, const double R2[3][3], const double T2[3],
                                   int first_contact, int64_t counters[4]) {
    orc_xform x;
    memcpy(x.R2, R2, sizeof x.R2); memcpy(x.T2, T2, sizeof x.T2);
    finish_xform(&x);
    return collide_obbtree_x(obst, robot, &x, first_contact, counters);
}
static int collide_obbtree_x(const orc_model *obst, const orc_model *robot, const orc_xform *xp, int first_contact,
                             int64_t counters[4]) {
    orc_xform x = *xp;
    static const double R1[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    const double T1[3] = {0, 0, 0};
    double tR1[3][3], tR2[3][3], tT1[3], tT2[3], R[3][3], T[3], u[3];
    const orc_box *r1 = &obst->boxes[0], *r2 = &robot->boxes[0];
    m_x_m(tR1, R1, r1->pR); m_x_v_p_v(tT1, R1, r1->pT, T1);
    m_x_m(tR2, x.R2, r2->pR); m_x_v_p_v(tT2, x.R2, r2->pT, x.T2);
    mt_x_m(R, tR1, tR2); v_sub(u, tT2, tT1); mt_x_v(T, tR1, u);
    orc_query q = {obst, robot, &x, first_contact, 0, 0, 0, 0};
    collide_rec(&q, 0, 0, R, T);
    if (counters) { counters[0] += q.n_box; counters[1] += q.n_tri; counters[2] += q.n_desc; counters[3] += q.n_contacts; }
    return q.n_contacts != 0;
}

ORC_API void orc_collide_obbtree_batch(const orc_model *obst, const orc_model *robot, const double *poses, int64_t n,
                                       int first_contact, uint8_t *out, int64_t counters[4], int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    int64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : c0, c1, c2, c3)
    for (int64_t i = 0; i < n; ++i) {
        int64_t c[4] = {0, 0, 0, 0};
        int v = orc_collide_obbtree(obst, robot, poses + 6 * i, first_contact, c);
        if (out) out[i] = (uint8_t)v;
        c0 += c[0]; c1 += c[1]; c2 += c[2]; c3 += c[3];
    }
    if (counters) { counters[0] += c0; counters[1] += c1; counters[2] += c2; counters[3] += c3; }
}

/* ------------------------------------------------------------------------------------------------
 * 5. Metric + local planner.
 *    Point<T>::distance            src/primitives.h:224-235  (wrap: :277-292)
 *    Solver<T,R>::isPathFree       src/problemStruct.h:154-168, collisionSampleSize :121
 * ---------------------------------------------------------------------------------------------- */
static inline double wrap_angle(double a) {
    if (a < -M_PI) return a + 2 * M_PI;
    else if (a >= M_PI) return a - 2 * M_PI;
    return a;
}
ORC_API double orc_distance6(const double a[6], const double b[6]) {
    double sum = 0;
    for (int i = 0; i < 3; ++i) { double d = a[i] - b[i]; sum += d * d; }
    for (int i = 3; i < 6; ++i) { double d = wrap_angle(b[i] - a[i]); sum += d * d; }
    return sqrt(sum);
}

/* number of interior samples the reference loop visits when nothing collides */
ORC_API int64_t orc_edge_num_samples(const double start[6], const double finish[6], double sample) {
    double parts = orc_distance6(start, finish) / sample;
    int64_t cnt = 0;
    for (unsigned int index = 1; index < parts; ++index) ++cnt;   /* same unsigned<double compare as the reference */
    return cnt;
}

/* rot_mode 0 = reference (interior samples carry yaw=pitch=roll=0)
 * rot_mode 1 = interpolate angles along the wrapped difference (opt-in "fixed" mode)
 * use_tree: 0 brute force verdicts, 1 OBB-tree verdicts (models must be given)
 * returns 1 if free; *first_hit = index of first colliding sample (1-based) or 0 */
ORC_API int orc_edge_free(const double *obst, int nT, const double *robot, int nR,
                          const orc_model *mo, const orc_model *mr,
                          const double start[6], const double finish[6], double sample, int rot_mode,
                          int32_t *first_hit, int64_t *samples_tested) {
    double total = orc_distance6(start, finish);
    double parts = total / sample;
    double dir[3] = {finish[0] - start[0], finish[1] - start[1], finish[2] - start[2]};
    double adir[3] = {wrap_angle(finish[3] - start[3]), wrap_angle(finish[4] - start[4]), wrap_angle(finish[5] - start[5])};
    int is_free = 1;
    if (first_hit) *first_hit = 0;
    for (unsigned int index = 1; index < parts && is_free; ++index) {
        double pos[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 3; ++i) pos[i] = start[i] + index * dir[i] / parts;
        if (rot_mode == 1) for (int i = 0; i < 3; ++i) pos[3 + i] = start[3 + i] + index * adir[i] / parts;
        int hit = mo ? orc_collide_obbtree(mo, mr, pos, 1, NULL) : orc_collide_brute(obst, nT, robot, nR, pos);
        if (samples_tested) ++*samples_tested;
        if (hit) { is_free = 0; if (first_hit) *first_hit = (int32_t)index; }
    }
    return is_free;
}

ORC_API void orc_edge_free_batch(const double *obst, int nT, const double *robot, int nR,
                                 const orc_model *mo, const orc_model *mr,
                                 const double *starts, const double *ends, int64_t m, double sample, int rot_mode,
                                 uint8_t *free_out, int32_t *first_hit_out, int64_t *samples_tested, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    int64_t tot = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : tot)
    for (int64_t i = 0; i < m; ++i) {
        int32_t fh = 0; int64_t st = 0;
        free_out[i] = (uint8_t)orc_edge_free(obst, nT, robot, nR, mo, mr, starts + 6 * i, ends + 6 * i, sample, rot_mode, &fh, &st);
        if (first_hit_out) first_hit_out[i] = fh;
        tot += st;
    }
    if (samples_tested) *samples_tested += tot;
}

/* ------------------------------------------------------------------------------------------------
 * 6. Exact k-NN / radius in float, the *intended* metric of D6Distance (src/primitives.h:404-438 with
 *    `+=` at :418/:423 and size-bounded loops), evaluated as FLANN's LinearIndex does
 *    (lib/flann/src/cpp/flann/algorithms/linear_index.h:132-147): ascending-id scan; result order is
 *    (d2, id) (result_set.h:72-75, :151-171, :475-496).  a = stored point, b = query.
 * ---------------------------------------------------------------------------------------------- */
static inline float wrap_angle_f(float angle) {
    /* NormalizeAngle<float>: comparisons and the +-2*pi in double, result narrowed to float */
    if (angle < -M_PI) return (float)(angle + 2 * M_PI);
    else if (angle >= M_PI) return (float)(angle - 2 * M_PI);
    return angle;
}
ORC_API float orc_d6_float(const float *a, const float *b, int dim) {
    float result = 0.0f, diff;
    int nlin = dim < 3 ? dim : 3;
    for (int i = 0; i < nlin; ++i) { diff = a[i] - b[i]; result += diff * diff; }
    for (int i = 3; i < dim; ++i) { diff = wrap_angle_f(b[i] - a[i]); result += diff * diff; }
    return result;
}

typedef struct { float d; int32_t id; } orc_di;
static inline int di_less(orc_di a, orc_di b) { return a.d < b.d || (a.d == b.d && a.id < b.id); }

ORC_API void orc_knn_linear(const float *nodes, int64_t n, int dim, const float *queries, int64_t nq, int k,
                            int32_t *ids_out, float *d2_out, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
    {
        orc_di *best = (orc_di *)malloc(sizeof(orc_di) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(dynamic, 4)
        for (int64_t qi = 0; qi < nq; ++qi) {
            const float *q = queries + qi * dim;
            int cnt = 0;
            for (int64_t i = 0; i < n; ++i) {
                orc_di c = {orc_d6_float(nodes + i * dim, q, dim), (int32_t)i};
                if (cnt == k) { if (k == 0 || !di_less(c, best[k - 1])) continue; }
                else ++cnt;
                int j = cnt - 1;
                while (j > 0 && di_less(c, best[j - 1])) { best[j] = best[j - 1]; --j; }
                best[j] = c;
            }
            for (int j = 0; j < k; ++j) {
                ids_out[qi * k + j] = j < cnt ? best[j].id : -1;
                d2_out[qi * k + j] = j < cnt ? best[j].d : INFINITY;
            }
        }
        free(best);
    }
}

static int di_cmp(const void *a, const void *b) {
    orc_di x = *(const orc_di *)a, y = *(const orc_di *)b;
    return di_less(x, y) ? -1 : (di_less(y, x) ? 1 : 0);
}
/* two-call protocol: counts[nq] always written; if ids_out != NULL rows are written at offsets[qi]
 * (offsets = exclusive scan of counts, supplied by the caller). strict d2 < r2, sorted by (d2,id). */
ORC_API void orc_radius_linear(const float *nodes, int64_t n, int dim, const float *queries, int64_t nq, float r2,
                               int32_t *counts, const int64_t *offsets, int32_t *ids_out, float *d2_out, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t qi = 0; qi < nq; ++qi) {
        const float *q = queries + qi * dim;
        int32_t cnt = 0;
        for (int64_t i = 0; i < n; ++i) if (orc_d6_float(nodes + i * dim, q, dim) < r2) ++cnt;
        counts[qi] = cnt;
        if (ids_out && cnt > 0) {
            orc_di *buf = (orc_di *)malloc(sizeof(orc_di) * (size_t)cnt);
            int32_t c = 0;
            for (int64_t i = 0; i < n; ++i) {
                float d = orc_d6_float(nodes + i * dim, q, dim);
                if (d < r2) { buf[c].d = d; buf[c].id = (int32_t)i; ++c; }
            }
            qsort(buf, (size_t)cnt, sizeof(orc_di), di_cmp);
            for (int32_t j = 0; j < cnt; ++j) { ids_out[offsets[qi] + j] = buf[j].id; d2_out[offsets[qi] + j] = buf[j].d; }
            free(buf);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * 7. Synthetic pose stream: Philox4x32-10, key = seed, counter = (index, stream).
 *    Distribution mirrors RandGen<T>::randomPointInSpace, src/randGen.h:123-146.  Only + * / sqrt are
 *    used (no libm transcendental) so that the CUDA generator is bit-identical.
 * ---------------------------------------------------------------------------------------------- */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
ORC_API void orc_philox4x32_10(uint64_t seed, uint64_t index, uint32_t stream, uint32_t out[4]) {
    uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), stream, 0};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}
static inline float u01(uint32_t bits) { return (float)(bits >> 8) * 5.9604644775390625e-08f; /* 2^-24 */ }

/* acos(x) for x in [-1,1]: sqrt(1-|x|) * poly(|x|) (Abramowitz & Stegun 4.4.46), reflected for x<0 */
static inline float acos_poly(float x) {
    float ax = fabsf(x);
    float p = -0.0012624911f;
    p = p * ax + 0.0066700901f;
    p = p * ax + -0.0170881256f;
    p = p * ax + 0.0308918810f;
    p = p * ax + -0.0501743046f;
    p = p * ax + 0.0889789874f;
    p = p * ax + -0.2145988016f;
    p = p * ax + 1.5707963050f;
    float r = sqrtf(1.0f - ax) * p;
    return x < 0.0f ? 3.14159274f - r : r;
}
/* range = {minx,maxx,miny,maxy,minz,maxz}; out float32 [n][6] for pose indices first..first+n-1 */
ORC_API void orc_gen_poses(uint64_t seed, uint64_t first, int64_t n, const float range[6], float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint32_t a[4], b[4];
        orc_philox4x32_10(seed, first + (uint64_t)i, 0, a);
        orc_philox4x32_10(seed, first + (uint64_t)i, 1, b);
        float *o = out + 6 * i;
        o[0] = range[0] + u01(a[0]) * (range[1] - range[0]);
        o[1] = range[2] + u01(a[1]) * (range[3] - range[2]);
        o[2] = range[4] + u01(a[2]) * (range[5] - range[4]);
        o[3] = -3.14159274f + u01(a[3]) * 6.28318548f;
        float phi = acos_poly(1.0f - 2.0f * u01(b[0])) + 1.57079637f;
        if (u01(b[1]) < 0.5f) phi -= 3.14159274f;
        o[4] = phi;
        o[5] = -3.14159274f + u01(b[2]) * 6.28318548f;
    }
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
