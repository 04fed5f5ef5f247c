/* orbx CPU oracle, local bundle adjustment part — TEST INFRASTRUCTURE ONLY (see orbx_oracle.h).
 *
 * Restates in dependency-free C (double precision, like g2o) what Optimizer::LocalBundleAdjustment
 * (reference src/Optimizer.cc:454-779) asks g2o to do, following
 *   EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ        Thirdparty/g2o/g2o/types/types_six_dof_expmap.{h,cpp}
 *   SE3Quat::exp, operator*, map, normalizeRotation    types/se3quat.h:188-285
 *   BaseBinaryEdge::constructQuadraticForm             core/base_binary_edge.hpp:55-120 (rho'' term dropped, base_edge.h:96-102)
 *   RobustKernelHuber::robustify                       core/robust_kernel_impl.cpp:78-91
 *   BlockSolver::buildSystem / setLambda / solve       core/block_solver.hpp:143-295, 354-486, 502-589
 *   OptimizationAlgorithmLevenberg::solve              core/optimization_algorithm_levenberg.cpp:61-189
 *   SparseOptimizer::optimize / push / pop / update    core/sparse_optimizer.cpp:354-435, 600-613
 * The reduced camera system is solved by a dense Cholesky instead of Eigen's SimplicialLDLT with AMD ordering
 * (solvers/linear_solver_eigen.h:94-124): same solution up to rounding.
 * PARITY PINNING: pinned against the reference's own code.  The reference has no tests or vectors for this path, but its
 * src/Optimizer.cc, src/Converter.cc and the whole vendored g2o compile unmodified against the Eigen stand-in oracle/eigenmini
 * (oracle/_ref/liboptimizer_ref.so, `make -C oracle ref_opt`).  tests/test_oracle_ref_optimizer.py: the edge / exp-map / Huber /
 * Converter leaves of this file return the same bits as the reference's classes on 10^4 inputs, and
 * Optimizer::LocalBundleAdjustment run on KeyFrame / MapPoint graphs built by the reference's constructors gives the same erased
 * observations and bad points, poses equal to float rounding and points within 1e-5 relative.  tests/test_lba_oracle.py
 * additionally checks this file against an independent numpy/scipy implementation of the same LM schedule.
 */
#include "orbx_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static double g_t_schur = 0, g_t_chol = 0;   /* split of schur_solve: Schur complement build / dense solve + back-substitution */

#include "se3_oracle.h"

typedef struct {
    const orbo_lba_problem *P;
    se3 *kf;          /* n_kf */
    double *pt;       /* n_pts x 3 */
    uint8_t *level1;  /* per edge: excluded from the second round */
    int robust;
    /* active sets */
    int *kf_idx, *pt_idx;   /* index in the reduced system or -1 */
    int np, nl;
    /* system */
    double *Hpp, *bp, *Hll, *bl, *Hpl;   /* np x 36, np x 6, nl x 9, nl x 3, n_edges x 18 (6x3 row-major) */
    double *err, *chi2;                  /* n_edges x 3, n_edges */
} lba;

static double huber_delta(int stereo) { return stereo ? (double)(float)sqrt(7.815) : (double)(float)sqrt(5.991); }   /* const float thHuber*, Optimizer.cc:569-570 */

static int edge_active(const lba *S, int e) { return !S->level1[e]; }

/* computeError + chi2; returns depth */
static double edge_error(const lba *S, int e, double err[3], double *chi2) {
    const orbo_lba_problem *P = S->P;
    double Xc[3];
    se3_map(&S->kf[P->e_kf[e]], S->pt + 3 * P->e_pt[e], Xc);
    const double info = (double)P->e_inv_sigma2[e];
    if (!P->e_stereo[e]) {
        const double u = Xc[0] / Xc[2] * P->fx + P->cx, v = Xc[1] / Xc[2] * P->fy + P->cy;
        err[0] = P->e_obs[3 * e] - u; err[1] = P->e_obs[3 * e + 1] - v; err[2] = 0;
    } else {
        /* cam_project keeps invz (and bf) in float, types_six_dof_expmap.cpp:150-157; `1.0f/trans_xyz[2]` has a double
         * divisor, so the division is done in double and narrowed once */
        const float invz = (float)(1.0 / Xc[2]);
        const double u = Xc[0] * invz * P->fx + P->cx, v = Xc[1] * invz * P->fy + P->cy;
        const double ur = u - (double)((float)P->bf * invz);
        err[0] = P->e_obs[3 * e] - u; err[1] = P->e_obs[3 * e + 1] - v; err[2] = P->e_obs[3 * e + 2] - ur;
    }
    /* chi2() = _error.dot(information() * _error) with information = I * invSigma2 (base_edge.h:58-61) */
    *chi2 = (err[0] * (info * err[0]) + err[1] * (info * err[1])) + err[2] * (info * err[2]);
    return Xc[2];
}

static double compute_errors(lba *S) {   /* computeActiveErrors + activeRobustChi2 */
    double total = 0;
    for (int e = 0; e < S->P->n_edges; e++) {
        if (!edge_active(S, e)) continue;
        edge_error(S, e, S->err + 3 * e, &S->chi2[e]);
        double c = S->chi2[e];
        if (S->robust) {
            const double d = huber_delta(S->P->e_stereo[e]), dsqr = (double)(float)(d * d);   /* `float dsqr`, robust_kernel_impl.h:84 */
            if (c > dsqr) c = 2 * sqrt(c) * d - dsqr;
        }
        total += c;
    }
    return total;
}

/* linearizeOplus of EdgeSE3ProjectXYZ (types_six_dof_expmap.cpp:103-139) and EdgeStereoSE3ProjectXYZ (:188-234):
 * A = dE/dX (D x 3), B = dE/dxi (D x 6), row-major */
static void edge_jacobians(const se3 *T, const double *X, int D, double fx, double fy, double bf, double A[9], double B[18]) {
    double R[9], Xc[3];
    memset(A, 0, sizeof(double) * 9); memset(B, 0, sizeof(double) * 18);
    quat_to_R(T->q, R);
    se3_map(T, X, Xc);
    const double x = Xc[0], y = Xc[1], z = Xc[2], z2 = z * z;
    if (D == 2) {   /* _jacobianOplusXi = -1./z * tmp * R with tmp = [fx 0 -x/z*fx; 0 fy -y/z*fy] (.cpp:116-126): ((-1/z) tmp) R, k ascending */
        const double s = -1. / z, t00 = s * fx, t01 = s * 0., t02 = s * (-x / z * fx), t10 = s * 0., t11 = s * fy, t12 = s * (-y / z * fy);
        for (int c = 0; c < 3; c++) {
            A[c] = (t00 * R[c] + t01 * R[3 + c]) + t02 * R[6 + c];
            A[3 + c] = (t10 * R[c] + t11 * R[3 + c]) + t12 * R[6 + c];
        }
    } else {        /* written out per coefficient in the stereo edge (.cpp:203-213) */
        for (int c = 0; c < 3; c++) {
            A[c] = -fx * R[c] / z + fx * x * R[6 + c] / z2;
            A[3 + c] = -fy * R[3 + c] / z + fy * y * R[6 + c] / z2;
            A[6 + c] = A[c] - bf * R[6 + c] / z2;
        }
    }
    B[0] = x * y / z2 * fx; B[1] = -(1 + (x * x / z2)) * fx; B[2] = y / z * fx; B[3] = -1. / z * fx; B[4] = 0; B[5] = x / z2 * fx;
    B[6] = (1 + y * y / z2) * fy; B[7] = -x * y / z2 * fy; B[8] = -x / z * fy; B[9] = 0; B[10] = -1. / z * fy; B[11] = y / z2 * fy;
    if (D == 3) { B[12] = B[0] - bf * y / z2; B[13] = B[1] + bf * x / z2; B[14] = B[2]; B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf / z2; }
}

static void build_system(lba *S) {   /* BlockSolver::buildSystem */
    const orbo_lba_problem *P = S->P;
    memset(S->Hpp, 0, sizeof(double) * 36 * S->np); memset(S->bp, 0, sizeof(double) * 6 * S->np);
    memset(S->Hll, 0, sizeof(double) * 9 * S->nl); memset(S->bl, 0, sizeof(double) * 3 * S->nl);
    for (int e = 0; e < P->n_edges; e++) {
        if (!edge_active(S, e)) continue;
        const int D = P->e_stereo[e] ? 3 : 2;
        double A[9], B[18];   /* A = dE/dX (D x 3), B = dE/dxi (D x 6) */
        edge_jacobians(&S->kf[P->e_kf[e]], S->pt + 3 * P->e_pt[e], D, P->fx, P->fy, P->bf, A, B);
        const double info = (double)P->e_inv_sigma2[e];
        double rho1 = 1.0;
        if (S->robust) {
            const double d = huber_delta(P->e_stereo[e]);
            if (S->chi2[e] > (double)(float)(d * d)) rho1 = d / sqrt(S->chi2[e]);
        }
        const double w = rho1 * info;
        const double *er = S->err + 3 * e;
        const int ip = S->kf_idx[P->e_kf[e]], il = S->pt_idx[P->e_pt[e]];
        if (il >= 0) {
            for (int a = 0; a < 3; a++) {
                double s = 0;
                for (int d = 0; d < D; d++) s += A[3 * d + a] * (-info * er[d]) * rho1;
                S->bl[3 * il + a] += s;
                for (int b = 0; b < 3; b++) {
                    double h = 0;
                    for (int d = 0; d < D; d++) h += A[3 * d + a] * w * A[3 * d + b];
                    S->Hll[9 * il + 3 * a + b] += h;
                }
            }
        }
        if (ip >= 0) {
            for (int a = 0; a < 6; a++) {
                double s = 0;
                for (int d = 0; d < D; d++) s += B[6 * d + a] * (-info * er[d]) * rho1;
                S->bp[6 * ip + a] += s;
                for (int b = 0; b < 6; b++) {
                    double h = 0;
                    for (int d = 0; d < D; d++) h += B[6 * d + a] * w * B[6 * d + b];
                    S->Hpp[36 * ip + 6 * a + b] += h;
                }
                if (il >= 0) for (int b = 0; b < 3; b++) {
                    double h = 0;
                    for (int d = 0; d < D; d++) h += B[6 * d + a] * w * A[3 * d + b];
                    S->Hpl[18 * e + 3 * a + b] = h;      /* one edge per (pose, landmark) pair */
                }
            }
        }
    }
}

static void inv3(const double *M, double *I) {
    const double a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5], g = M[6], h = M[7], i = M[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g), id = 1.0 / det;
    I[0] = (e * i - f * h) * id; I[1] = (c * h - b * i) * id; I[2] = (b * f - c * e) * id;
    I[3] = (f * g - d * i) * id; I[4] = (a * i - c * g) * id; I[5] = (c * d - a * f) * id;
    I[6] = (d * h - e * g) * id; I[7] = (b * g - a * h) * id; I[8] = (a * e - b * d) * id;
}

/* BlockSolver::solve with lambda on both diagonals; Hs (6np x 6np, full symmetric), bs out for inspection */
static int schur_solve(lba *S, double lambda, double *xp, double *xl, double *Hs, double *bs) {
    const orbo_lba_problem *P = S->P;
    const int np = S->np, nl = S->nl, n = 6 * np;
    const double t_begin = now_s();
    memset(Hs, 0, sizeof(double) * n * n);
    for (int p = 0; p < np; p++)
        for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++)
            Hs[(6 * p + a) * n + 6 * p + b] = S->Hpp[36 * p + 6 * a + b] + (a == b ? lambda : 0);
    memcpy(bs, S->bp, sizeof(double) * n);
    double *Dinv = (double *)malloc(sizeof(double) * 9 * (nl ? nl : 1));
    for (int l = 0; l < nl; l++) {
        double D[9];
        memcpy(D, S->Hll + 9 * l, sizeof(D));
        D[0] += lambda; D[4] += lambda; D[8] += lambda;
        inv3(D, Dinv + 9 * l);
    }
    /* edges grouped by landmark */
    int *head = (int *)malloc(sizeof(int) * (nl + 1)), *list = (int *)malloc(sizeof(int) * (P->n_edges ? P->n_edges : 1));
    memset(head, 0, sizeof(int) * (nl + 1));
    for (int e = 0; e < P->n_edges; e++) if (edge_active(S, e) && S->pt_idx[P->e_pt[e]] >= 0 && S->kf_idx[P->e_kf[e]] >= 0) head[S->pt_idx[P->e_pt[e]] + 1]++;
    for (int l = 0; l < nl; l++) head[l + 1] += head[l];
    int *cur = (int *)malloc(sizeof(int) * (nl ? nl : 1));
    memcpy(cur, head, sizeof(int) * nl);
    for (int e = 0; e < P->n_edges; e++) if (edge_active(S, e) && S->pt_idx[P->e_pt[e]] >= 0 && S->kf_idx[P->e_kf[e]] >= 0) list[cur[S->pt_idx[P->e_pt[e]]]++] = e;
    for (int l = 0; l < nl; l++) {
        const double *Di = Dinv + 9 * l;
        double db[3];
        for (int a = 0; a < 3; a++) db[a] = Di[3 * a] * S->bl[3 * l] + Di[3 * a + 1] * S->bl[3 * l + 1] + Di[3 * a + 2] * S->bl[3 * l + 2];
        for (int i = head[l]; i < head[l + 1]; i++) {
            const int e1 = list[i], p1 = S->kf_idx[P->e_kf[e1]];
            const double *B1 = S->Hpl + 18 * e1;
            double BD[18];
            for (int a = 0; a < 6; a++) for (int b = 0; b < 3; b++)
                BD[3 * a + b] = B1[3 * a] * Di[b] + B1[3 * a + 1] * Di[3 + b] + B1[3 * a + 2] * Di[6 + b];
            for (int a = 0; a < 6; a++) bs[6 * p1 + a] -= B1[3 * a] * db[0] + B1[3 * a + 1] * db[1] + B1[3 * a + 2] * db[2];
            for (int j = head[l]; j < head[l + 1]; j++) {
                const int e2 = list[j], p2 = S->kf_idx[P->e_kf[e2]];
                const double *B2 = S->Hpl + 18 * e2;
                for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++)
                    Hs[(6 * p1 + a) * n + 6 * p2 + b] -= BD[3 * a] * B2[3 * b] + BD[3 * a + 1] * B2[3 * b + 1] + BD[3 * a + 2] * B2[3 * b + 2];
            }
        }
    }
    const double t_mid = now_s();
    double *Hc = (double *)malloc(sizeof(double) * (n ? n * n : 1));
    memcpy(Hc, Hs, sizeof(double) * n * n);
    const int ok = n == 0 ? 1 : chol_solve(Hc, bs, xp, n);
    if (!ok) {   /* failed factorisation: the trial is rejected whatever x holds; define x = 0 (the CUDA path does the same) */
        memset(xp, 0, sizeof(double) * n);
        memset(xl, 0, sizeof(double) * 3 * nl);
    }
    if (ok) {
        for (int l = 0; l < nl; l++) {
            double c[3] = {S->bl[3 * l], S->bl[3 * l + 1], S->bl[3 * l + 2]};
            for (int i = head[l]; i < head[l + 1]; i++) {
                const int e = list[i], p = S->kf_idx[P->e_kf[e]];
                const double *B = S->Hpl + 18 * e;
                for (int b = 0; b < 3; b++) for (int a = 0; a < 6; a++) c[b] -= B[3 * a + b] * xp[6 * p + a];
            }
            const double *Di = Dinv + 9 * l;
            for (int a = 0; a < 3; a++) xl[3 * l + a] = Di[3 * a] * c[0] + Di[3 * a + 1] * c[1] + Di[3 * a + 2] * c[2];
        }
    }
    free(Hc); free(Dinv); free(head); free(list); free(cur);
    g_t_schur += t_mid - t_begin; g_t_chol += now_s() - t_mid;
    return ok;
}

static int stop_requested(const orbo_lba_problem *P) { return P->stop_flag && *P->stop_flag; }

/* one optimizer.initializeOptimization(level 0) + optimize(iterations); returns iterations done */
static int optimize(lba *S, int iterations, orbo_lba_trace *tr) {
    const orbo_lba_problem *P = S->P;
    /* active vertices: those touched by an active edge (sparse_optimizer.cpp:166-267), free poses first, then landmarks */
    uint8_t *kf_used = (uint8_t *)calloc(P->n_kf ? P->n_kf : 1, 1), *pt_used = (uint8_t *)calloc(P->n_pts ? P->n_pts : 1, 1);
    for (int e = 0; e < P->n_edges; e++) if (edge_active(S, e)) { kf_used[P->e_kf[e]] = 1; pt_used[P->e_pt[e]] = 1; }
    S->np = S->nl = 0;
    for (int k = 0; k < P->n_kf; k++) S->kf_idx[k] = (kf_used[k] && !P->kf_fixed[k]) ? S->np++ : -1;
    for (int l = 0; l < P->n_pts; l++) S->pt_idx[l] = pt_used[l] ? S->nl++ : -1;
    free(kf_used); free(pt_used);
    const int np = S->np, nl = S->nl, n = 6 * np;
    double *xp = (double *)calloc(n ? n : 1, sizeof(double)), *xl = (double *)calloc(nl ? 3 * nl : 1, sizeof(double));
    double *Hs = (double *)malloc(sizeof(double) * (n ? n * n : 1)), *bs = (double *)malloc(sizeof(double) * (n ? n : 1));
    se3 *kf_bak = (se3 *)malloc(sizeof(se3) * (P->n_kf ? P->n_kf : 1));
    double *pt_bak = (double *)malloc(sizeof(double) * 3 * (P->n_pts ? P->n_pts : 1));
    double lambda = 0, ni = 2;
    int nBad = 0, done = 0;
    for (int it = 0; it < iterations && !stop_requested(P); it++) {
        double tb = now_s();
        double currentChi = compute_errors(S);
        const double iniChi = currentChi;
        double tempChi = currentChi;
        build_system(S);
        if (tr) { tr->t_build += now_s() - tb; tr->n_builds++; }
        if (it == 0) {   /* computeLambdaInit: tau * max diagonal over every active vertex */
            double mx = 0;
            for (int p = 0; p < np; p++) for (int a = 0; a < 6; a++) mx = fmax(mx, fabs(S->Hpp[36 * p + 7 * a]));
            for (int l = 0; l < nl; l++) for (int a = 0; a < 3; a++) mx = fmax(mx, fabs(S->Hll[9 * l + 4 * a]));
            lambda = 1e-5 * mx; ni = 2; nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            memcpy(kf_bak, S->kf, sizeof(se3) * P->n_kf); memcpy(pt_bak, S->pt, sizeof(double) * 3 * P->n_pts);   /* push */
            g_t_schur = g_t_chol = 0;
            const int ok2 = schur_solve(S, lambda, xp, xl, Hs, bs);
            if (tr) { tr->t_schur += g_t_schur; tr->t_solve += g_t_chol; }
            if (tr && tr->n_trials == 0 && tr->Hschur) {   /* first trial of the call: keep the reduced system */
                memcpy(tr->Hschur, Hs, sizeof(double) * n * n); memcpy(tr->bschur, bs, sizeof(double) * n);
                tr->dim = n; tr->lambda0 = lambda;
                if (tr->xp) memcpy(tr->xp, xp, sizeof(double) * n);
            }
            if (ok2) {   /* SparseOptimizer::update */
                for (int k = 0; k < P->n_kf; k++) if (S->kf_idx[k] >= 0) se3_oplus(&S->kf[k], xp + 6 * S->kf_idx[k]);
                for (int l = 0; l < P->n_pts; l++) if (S->pt_idx[l] >= 0) for (int a = 0; a < 3; a++) S->pt[3 * l + a] += xl[3 * S->pt_idx[l] + a];
            }
            tempChi = compute_errors(S);
            if (!ok2) tempChi = 1.7976931348623157e308;
            rho = currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < n; j++) scale += xp[j] * (lambda * xp[j] + S->bp[j]);
            for (int j = 0; j < 3 * nl; j++) scale += xl[j] * (lambda * xl[j] + S->bl[j]);
            scale += 1e-3;
            rho /= scale;
            if (tr) { if (tr->n_trials < ORBO_LBA_MAX_TRACE) { tr->chi2[tr->n_trials] = tempChi; tr->lambda[tr->n_trials] = lambda; } tr->n_trials++; }
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni; ni *= 2;
                memcpy(S->kf, kf_bak, sizeof(se3) * P->n_kf); memcpy(S->pt, pt_bak, sizeof(double) * 3 * P->n_pts);   /* pop */
            }
            qmax++;
        } while (rho < 0 && qmax < 10 && !stop_requested(P));
        done++;
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
        if (nBad >= 3) break;
    }
    free(xp); free(xl); free(Hs); free(bs); free(kf_bak); free(pt_bak);
    return done;
}

/* Optimizer::LocalBundleAdjustment from `optimizer.initializeOptimization()` (Optimizer.cc:659) to the erase list (:735) */
int orbo_lba_solve(const orbo_lba_problem *P, int its1, int its2, double *kf_out, double *pt_out, double *chi2_out,
                   uint8_t *erase_out, orbo_lba_trace *tr) {
    lba S;
    memset(&S, 0, sizeof(S));
    S.P = P;
    S.kf = (se3 *)malloc(sizeof(se3) * (P->n_kf ? P->n_kf : 1));
    S.pt = (double *)malloc(sizeof(double) * 3 * (P->n_pts ? P->n_pts : 1));
    for (int k = 0; k < P->n_kf; k++) { memcpy(S.kf[k].q, P->kf_pose + 7 * k, sizeof(double) * 4); memcpy(S.kf[k].t, P->kf_pose + 7 * k + 4, sizeof(double) * 3); }
    memcpy(S.pt, P->pts, sizeof(double) * 3 * P->n_pts);
    S.level1 = (uint8_t *)calloc(P->n_edges ? P->n_edges : 1, 1);
    S.kf_idx = (int *)malloc(sizeof(int) * (P->n_kf ? P->n_kf : 1)); S.pt_idx = (int *)malloc(sizeof(int) * (P->n_pts ? P->n_pts : 1));
    S.Hpp = (double *)malloc(sizeof(double) * 36 * (P->n_kf ? P->n_kf : 1)); S.bp = (double *)malloc(sizeof(double) * 6 * (P->n_kf ? P->n_kf : 1));
    S.Hll = (double *)malloc(sizeof(double) * 9 * (P->n_pts ? P->n_pts : 1)); S.bl = (double *)malloc(sizeof(double) * 3 * (P->n_pts ? P->n_pts : 1));
    S.Hpl = (double *)calloc(18 * (size_t)(P->n_edges ? P->n_edges : 1), sizeof(double));
    S.err = (double *)calloc(3 * (size_t)(P->n_edges ? P->n_edges : 1), sizeof(double)); S.chi2 = (double *)calloc(P->n_edges ? P->n_edges : 1, sizeof(double));
    if (tr) { tr->n_trials = 0; tr->t_build = tr->t_schur = tr->t_solve = 0; tr->n_builds = 0; }
    int rc = 0;
    if (!stop_requested(P)) {
        S.robust = 1;
        optimize(&S, its1, tr);
        if (!stop_requested(P) && its2 > 0) {
            for (int e = 0; e < P->n_edges; e++) {   /* Optimizer.cc:671-703 */
                /* e->chi2() reads the error stored by the last computeActiveErrors (not refreshed after a rejected
                 * trial's pop); isDepthPositive() re-maps the point with the current estimates */
                double er[3], c;
                const double z = edge_error(&S, e, er, &c);
                if (S.chi2[e] > (P->e_stereo[e] ? 7.815 : 5.991) || !(z > 0)) S.level1[e] = 1;
            }
            S.robust = 0;
            optimize(&S, its2, tr);
        }
        for (int e = 0; e < P->n_edges; e++) {       /* Optimizer.cc:709-735; chi2() is the edge's stored error */
            /* chi2() is the stored error: edges left out of round 2 keep round 1's last value */
            double er[3], c;
            const double z = edge_error(&S, e, er, &c);
            c = S.chi2[e];
            if (chi2_out) chi2_out[e] = c;
            if (erase_out) erase_out[e] = (c > (P->e_stereo[e] ? 7.815 : 5.991) || !(z > 0)) ? 1 : 0;
        }
    } else rc = 1;
    for (int k = 0; k < P->n_kf; k++) { memcpy(kf_out + 7 * k, S.kf[k].q, sizeof(double) * 4); memcpy(kf_out + 7 * k + 4, S.kf[k].t, sizeof(double) * 3); }
    memcpy(pt_out, S.pt, sizeof(double) * 3 * P->n_pts);
    free(S.kf); free(S.pt); free(S.level1); free(S.kf_idx); free(S.pt_idx); free(S.Hpp); free(S.bp); free(S.Hll); free(S.bl);
    free(S.Hpl); free(S.err); free(S.chi2);
    return rc;
}

/* ---- leaf entry points: the same static functions as above on one edge / one vertex, for tests/test_oracle_ref_optimizer.py,
 * which compares them with the reference's g2o classes compiled from /root/reference (oracle/_ref/liboptimizer_ref.so) ---- */
void orbo_lba_edge_eval(int stereo, const double pose[7], const double X[3], const double obs[3], const double K[5], float inv_sigma2,
                        double *err, double *chi2, int *depth_positive, double *A, double *B) {
    orbo_lba_problem P;
    memset(&P, 0, sizeof(P));
    const int32_t zero = 0;
    const uint8_t st = (uint8_t)(stereo != 0), fixed = 0;
    P.n_kf = 1; P.kf_pose = pose; P.kf_fixed = &fixed; P.n_pts = 1; P.pts = X; P.n_edges = 1; P.e_kf = &zero; P.e_pt = &zero;
    P.e_obs = obs; P.e_inv_sigma2 = &inv_sigma2; P.e_stereo = &st; P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3]; P.bf = K[4];
    lba S;
    memset(&S, 0, sizeof(S));
    se3 T;
    memcpy(T.q, pose, sizeof(double) * 4); memcpy(T.t, pose + 4, sizeof(double) * 3);
    double pt[3] = {X[0], X[1], X[2]}, e3[3];
    S.P = &P; S.kf = &T; S.pt = pt;
    const double z = edge_error(&S, 0, e3, chi2);
    for (int i = 0; i < (stereo ? 3 : 2); i++) err[i] = e3[i];
    *depth_positive = z > 0;
    double A9[9], B18[18];
    edge_jacobians(&T, pt, stereo ? 3 : 2, P.fx, P.fy, P.bf, A9, B18);
    memcpy(A, A9, sizeof(double) * 3 * (stereo ? 3 : 2)); memcpy(B, B18, sizeof(double) * 6 * (stereo ? 3 : 2));
}
void orbo_se3_oplus(const double pose[7], const double update[6], double out[7]) {
    se3 T;
    memcpy(T.q, pose, sizeof(double) * 4); memcpy(T.t, pose + 4, sizeof(double) * 3);
    se3_oplus(&T, update);
    memcpy(out, T.q, sizeof(double) * 4); memcpy(out + 4, T.t, sizeof(double) * 3);
}
void orbo_se3_map(const double pose[7], const double X[3], double out[3]) {
    se3 T;
    memcpy(T.q, pose, sizeof(double) * 4); memcpy(T.t, pose + 4, sizeof(double) * 3);
    se3_map(&T, X, out);
}
/* RobustKernelHuber::robustify as the two optimiser oracles use it: rho[0] enters the robust chi2 (compute_errors), rho[1] scales
 * the information and the right-hand side (build_system); rho[2] is dropped by g2o (base_edge.h:96-102) */
void orbo_huber(double e2, double delta, double rho[3]) {
    const double dsqr = (double)(float)(delta * delta);   /* `float dsqr`, robust_kernel_impl.h:84 */
    if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.; rho[2] = 0.; }
    else { const double sqrte = sqrt(e2); rho[0] = 2 * sqrte * delta - dsqr; rho[1] = delta / sqrte; rho[2] = -0.5 * rho[1] / e2; }
}
/* Converter::toSE3Quat(cv::Mat) (src/Converter.cc:41-51): float entries widened, Quaterniond(R), normalizeRotation */
void orbo_to_se3quat(const float Tcw[16], double pose[7]) {
    double R[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[3 * r + c] = (double)Tcw[4 * r + c];
    R_to_quat(R, pose);
    quat_normalize(pose);
    for (int r = 0; r < 3; r++) pose[4 + r] = (double)Tcw[4 * r + 3];
}
/* Converter::toCvMat(SE3Quat) (src/Converter.cc:53-57, 67-75): to_homogeneous_matrix narrowed to float */
void orbo_to_cvmat(const double pose[7], float Tcw[16]) {
    double R[9];
    quat_to_R(pose, R);
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) Tcw[4 * r + c] = (float)R[3 * r + c]; Tcw[4 * r + 3] = (float)pose[4 + r]; }
    Tcw[12] = Tcw[13] = Tcw[14] = 0.f; Tcw[15] = 1.f;
}
