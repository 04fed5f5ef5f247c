/* orbx CPU oracle — TEST INFRASTRUCTURE ONLY (see orbx_oracle.h).
 * SE3 / quaternion helpers shared by the local-BA and pose-optimisation oracles, restating g2o's
 *   SE3Quat::exp, operator*, map, normalizeRotation    Thirdparty/g2o/g2o/types/se3quat.h:188-285
 *   Eigen::Quaterniond(Matrix3d), toRotationMatrix
 * and a dense Cholesky solve standing in for the reference's linear solvers (same solution up to rounding). */
#ifndef ORBX_SE3_ORACLE_H
#define ORBX_SE3_ORACLE_H
#include <math.h>

typedef struct { double q[4]; double t[3]; } se3;   /* q = (x,y,z,w) like Eigen::Quaterniond::coeffs() */

static inline void quat_to_R(const double q[4], double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
/* Eigen::Quaterniond(Matrix3d) */
static inline void R_to_quat(const double R[9], double q[4]) {
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
}
static inline void quat_normalize(double q[4]) {   /* SE3Quat::normalizeRotation */
    if (q[3] < 0) for (int i = 0; i < 4; i++) q[i] = -q[i];
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) q[i] /= n;
}
static inline void quat_mul(const double a[4], const double b[4], double o[4]) {
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    o[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
/* Eigen's Quaternion * Vector3 (QuaternionBase::_transformVector): uv = 2 (q.vec x v); v + w uv + q.vec x uv */
static inline void quat_rotate(const double q[4], const double v[3], double o[3]) {
    double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    for (int i = 0; i < 3; i++) uv[i] += uv[i];
    const double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
    for (int i = 0; i < 3; i++) o[i] = (v[i] + q[3] * uv[i]) + c[i];
}
/* SE3Quat::map: _r*xyz + _t (se3quat.h:217-220) */
static inline void se3_map(const se3 *T, const double X[3], double o[3]) {
    double r[3];
    quat_rotate(T->q, X, r);
    for (int i = 0; i < 3; i++) o[i] = r[i] + T->t[i];
}
/* T <- exp(update) * T, VertexSE3Expmap::oplusImpl */
static inline void se3_oplus(se3 *T, const double u[6]) {
    const double wx = u[0], wy = u[1], wz = u[2];
    const double theta = sqrt(wx * wx + wy * wy + wz * wz);
    const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double O2[9], R[9], V[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += O[3 * r + k] * O[3 * k + c];
        O2[3 * r + c] = s;
    }
    if (theta < 0.00001) {
        for (int i = 0; i < 9; i++) { R[i] = (i % 4 == 0) + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / pow(theta, 3);
        for (int i = 0; i < 9; i++) { R[i] = (i % 4 == 0) + a * O[i] + b * O2[i]; V[i] = (i % 4 == 0) + b * O[i] + c * O2[i]; }
    }
    se3 E;
    R_to_quat(R, E.q);
    quat_normalize(E.q);
    for (int r = 0; r < 3; r++) E.t[r] = V[3 * r] * u[3] + V[3 * r + 1] * u[4] + V[3 * r + 2] * u[5];
    /* SE3Quat::operator*: r = r1*r2, t = t1 + r1*t2, normalize */
    se3 N;
    quat_mul(E.q, T->q, N.q);
    double rt[3];
    quat_rotate(E.q, T->t, rt);
    for (int r = 0; r < 3; r++) N.t[r] = E.t[r] + rt[r];
    quat_normalize(N.q);
    *T = N;
}

static inline int chol_solve(double *Ain, const double *b, double *x, int n) {   /* dense SPD solve, A destroyed */
    for (int j = 0; j < n; j++) {
        double d = Ain[j * n + j];
        for (int k = 0; k < j; k++) d -= Ain[j * n + k] * Ain[j * n + k];
        if (!(d > 0)) return 0;
        d = sqrt(d);
        Ain[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = Ain[i * n + j];
            for (int k = 0; k < j; k++) s -= Ain[i * n + k] * Ain[j * n + k];
            Ain[i * n + j] = s / d;
        }
    }
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= Ain[i * n + k] * x[k];
        x[i] = s / Ain[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= Ain[k * n + i] * x[k];
        x[i] = s / Ain[i * n + i];
    }
    return 1;
}


#endif
