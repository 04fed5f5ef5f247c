/* orbx CPU oracle, pose-only optimisation — TEST INFRASTRUCTURE ONLY (see orbx_oracle.h).
 *
 * Restates what Optimizer::PoseOptimization (reference src/Optimizer.cc:239-452) asks g2o to do:
 *   EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose   Thirdparty/g2o/g2o/types/types_six_dof_expmap.h:143-202,
 *                                                                  .cpp:266-364 (stereo cam_project keeps 1/z in float)
 *   BaseUnaryEdge::constructQuadraticForm                          core/base_unary_edge.hpp:46-74
 *   RobustKernelHuber::robustify                                   core/robust_kernel_impl.cpp:78-91
 *   BlockSolver::buildSystem / setLambda / solve (no Schur: the graph has no marginalised vertex)   core/block_solver.hpp
 *   LinearSolverDense::solve (6x6)                                 solvers/linear_solver_dense.h:63-115
 *   OptimizationAlgorithmLevenberg::solve                          core/optimization_algorithm_levenberg.cpp:61-189
 * Four rounds of optimize(10); each restarts from the frame's pose (:366), excludes the previous round's outliers
 * (setLevel(1)), classifies every edge with FLOAT chi2 against 5.991f / 7.815f (:377-379, :401-403), drops the Huber kernel
 * after the third round (:391, :416) and stops early when the graph has fewer than 10 edges (:419).
 * The 6x6 system is solved by Cholesky (LL^T) where the reference uses Eigen's LDLT: same solution up to rounding.
 * PARITY PINNING: pinned against the reference's own code: src/Optimizer.cc + g2o compiled unmodified against the Eigen stand-in
 * oracle/eigenmini (oracle/_ref/liboptimizer_ref.so).  tests/test_oracle_ref_optimizer.py: the OnlyPose edges of this file return
 * the same bits as the reference's classes, and Optimizer::PoseOptimization run on a reference-built Frame gives the same
 * mvbOutlier flags, return value and nBadPoseOpt and a float pose identical to this file's.  tests/test_pose_oracle.py also
 * checks it against an independent numpy implementation of the same schedule.
 */
#include "orbx_oracle.h"
#include "se3_oracle.h"
#include <float.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const orbo_pose_problem *P;
    se3 T;
    uint8_t *level1;
    int robust;
    double *err, *chi2;    /* stored _error / chi2 of the last computeActiveErrors per edge */
    double H[36], b[6];
} pose_opt;

static int is_stereo(const orbo_pose_problem *P, int e) { return !(P->obs[3 * e + 2] < 0); }   /* mvuRight[i] < 0 -> monocular, :281 */
static double delta_of(int stereo) { return stereo ? (double)(float)sqrt(7.815) : (double)(float)sqrt(5.991); }   /* const float deltaMono/Stereo, :270-271 */

static void edge_error(const pose_opt *S, int e, double err[3], double *chi2, double Xc[3]) {
    const orbo_pose_problem *P = S->P;
    se3_map(&S->T, P->Xw + 3 * e, Xc);
    const double info = (double)P->inv_sigma2[e];
    if (!is_stereo(P, e)) {
        err[0] = P->obs[3 * e] - (Xc[0] / Xc[2] * P->fx + P->cx);
        err[1] = P->obs[3 * e + 1] - (Xc[1] / Xc[2] * P->fy + P->cy);
        err[2] = 0;
    } else {
        const float invz = (float)(1.0 / Xc[2]);       /* `1.0f/trans_xyz[2]`: double division, narrowed once (.cpp:303) */
        const double u = Xc[0] * invz * P->fx + P->cx, v = Xc[1] * invz * P->fy + P->cy;
        err[0] = P->obs[3 * e] - u; err[1] = P->obs[3 * e + 1] - v;
        /* here bf is the edge's DOUBLE member (set from the float mbf, Optimizer.cc:337), so bf*invz is a double product --
         * unlike EdgeStereoSE3ProjectXYZ::cam_project, whose bf parameter is a float */
        err[2] = P->obs[3 * e + 2] - (u - (double)(float)P->bf * (double)invz);
    }
    *chi2 = (err[0] * (info * err[0]) + err[1] * (info * err[1])) + err[2] * (info * err[2]);   /* _error.dot(information() * _error) */
}

static double compute_errors(pose_opt *S) {
    double total = 0;
    for (int e = 0; e < S->P->n; e++) {
        if (S->level1[e]) continue;
        double Xc[3];
        edge_error(S, e, S->err + 3 * e, &S->chi2[e], Xc);
        double c = S->chi2[e];
        if (S->robust) {
            const double d = delta_of(is_stereo(S->P, e)), dsqr = (double)(float)(d * d);   /* `float dsqr`, robust_kernel_impl.h:84 */
            if (c > dsqr) c = 2 * sqrt(c) * d - dsqr;
        }
        total += c;
    }
    return total;
}

/* linearizeOplus of EdgeSE3ProjectXYZOnlyPose (types_six_dof_expmap.cpp:266-288) and EdgeStereoSE3ProjectXYZOnlyPose (:335-364):
 * J = dE/dxi (D x 6), row-major, from the point in camera coordinates */
static void pose_edge_jacobian(const double Xc[3], int st, double fx, double fy, double bf, double J[18]) {
    const double x = Xc[0], y = Xc[1], invz = 1.0 / Xc[2], invz_2 = invz * invz;
    memset(J, 0, sizeof(double) * 18);
    J[0] = x * y * invz_2 * fx; J[1] = -(1 + (x * x * invz_2)) * fx; J[2] = y * invz * fx; J[3] = -invz * fx; J[4] = 0; J[5] = x * invz_2 * fx;
    J[6] = (1 + y * y * invz_2) * fy; J[7] = -x * y * invz_2 * fy; J[8] = -x * invz * fy; J[9] = 0; J[10] = -invz * fy; J[11] = y * invz_2 * fy;
    if (st) { J[12] = J[0] - bf * y * invz_2; J[13] = J[1] + bf * x * invz_2; J[14] = J[2]; J[15] = J[3]; J[16] = 0; J[17] = J[5] - bf * invz_2; }
}

static void build_system(pose_opt *S) {
    const orbo_pose_problem *P = S->P;
    memset(S->H, 0, sizeof(S->H)); memset(S->b, 0, sizeof(S->b));
    for (int e = 0; e < P->n; e++) {
        if (S->level1[e]) continue;
        const int st = is_stereo(P, e), D = st ? 3 : 2;
        double Xc[3];
        se3_map(&S->T, P->Xw + 3 * e, Xc);
        double J[18];
        pose_edge_jacobian(Xc, st, P->fx, P->fy, (double)(float)P->bf, J);
        const double info = (double)P->inv_sigma2[e];
        double rho1 = 1.0;
        if (S->robust) {
            const double d = delta_of(st);
            if (S->chi2[e] > (double)(float)(d * d)) rho1 = d / sqrt(S->chi2[e]);
        }
        const double *er = S->err + 3 * e;
        for (int a = 0; a < 6; a++) {
            double s = 0;
            for (int d = 0; d < D; d++) s += J[6 * d + a] * info * er[d];
            S->b[a] -= rho1 * s;
            for (int c = 0; c < 6; c++) {
                double h = 0;
                for (int d = 0; d < D; d++) h += J[6 * d + a] * (rho1 * info) * J[6 * d + c];
                S->H[6 * a + c] += h;
            }
        }
    }
}

/* initializeOptimization(0) + optimize(iterations); returns LM trials */
static int optimize(pose_opt *S, int iterations) {
    int n_active = 0, trials = 0;
    for (int e = 0; e < S->P->n; e++) n_active += !S->level1[e];
    if (n_active == 0) return 0;                       /* "0 vertices to optimize": optimize() returns at once */
    double lambda = 0, ni = 2, x[6] = {0};
    int nBad = 0;
    for (int it = 0; it < iterations; it++) {
        double currentChi = compute_errors(S);
        const double iniChi = currentChi;
        double tempChi = currentChi;
        build_system(S);
        if (it == 0) {
            double mx = 0;
            for (int a = 0; a < 6; a++) mx = fmax(mx, fabs(S->H[7 * a]));
            lambda = 1e-5 * mx; ni = 2; nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            const se3 bak = S->T;                       /* push */
            double A[36];
            memcpy(A, S->H, sizeof(A));
            for (int a = 0; a < 6; a++) A[7 * a] += lambda;
            const int ok = chol_solve(A, S->b, x, 6);
            if (!ok) memset(x, 0, sizeof(x));
            if (ok) se3_oplus(&S->T, x);
            tempChi = compute_errors(S);
            if (!ok) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + S->b[j]);
            scale += 1e-3;
            rho /= scale;
            trials++;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni; ni *= 2;
                S->T = bak;                             /* pop */
            }
            qmax++;
        } while (rho < 0 && qmax < 10);
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
        if (nBad >= 3) break;
    }
    return trials;
}

int orbo_pose_optimize(const orbo_pose_problem *P, double pose_out[7], uint8_t *outlier, int32_t *n_bad, int32_t *lm_trials) {
    const int n = P->n;
    memcpy(pose_out, P->pose, sizeof(double) * 7);
    if (n_bad) *n_bad = 0;
    if (lm_trials) *lm_trials = 0;
    for (int e = 0; e < n; e++) outlier[e] = 0;         /* mvbOutlier[i] = false, :284, :319 */
    if (n < 3) return 0;                                /* :355-356 */
    pose_opt S;
    memset(&S, 0, sizeof(S));
    S.P = P;
    S.level1 = (uint8_t *)calloc((size_t)n, 1);
    S.err = (double *)calloc(3 * (size_t)n, sizeof(double)); S.chi2 = (double *)calloc((size_t)n, sizeof(double));
    S.robust = 1;
    const float chi2Mono = 5.991, chi2Stereo = 7.815;
    int nBad = 0, trials = 0;
    for (int it = 0; it < 4; it++) {
        memcpy(S.T.q, P->pose, sizeof(double) * 4); memcpy(S.T.t, P->pose + 4, sizeof(double) * 3);   /* setEstimate(mTcw), :366 */
        trials += optimize(&S, 10);
        nBad = 0;
        for (int e = 0; e < n; e++) {                   /* :371-417 (mono and stereo loops do the same per edge) */
            if (outlier[e]) {                           /* was left out of this round: refresh its error */
                double Xc[3];
                edge_error(&S, e, S.err + 3 * e, &S.chi2[e], Xc);
            }
            const float chi2 = (float)S.chi2[e];
            if (chi2 > (is_stereo(P, e) ? chi2Stereo : chi2Mono)) { outlier[e] = 1; S.level1[e] = 1; nBad++; }
            else { outlier[e] = 0; S.level1[e] = 0; }
        }
        if (it == 2) S.robust = 0;
        if (n < 10) break;                              /* optimizer.edges().size() < 10, :419 */
    }
    memcpy(pose_out, S.T.q, sizeof(double) * 4); memcpy(pose_out + 4, S.T.t, sizeof(double) * 3);
    if (n_bad) *n_bad = nBad;
    if (lm_trials) *lm_trials = trials;
    free(S.level1); free(S.err); free(S.chi2);
    return n - nBad;
}

/* leaf entry point: the same static functions on one edge (compared with the reference's g2o classes in
 * tests/test_oracle_ref_optimizer.py); obs[2] < 0 selects the monocular edge like mvuRight */
void orbo_pose_edge_eval(const double pose[7], const double X[3], const double obs[3], const double K[5], float inv_sigma2,
                         double *err, double *chi2, int *depth_positive, double *J) {
    orbo_pose_problem P;
    memset(&P, 0, sizeof(P));
    P.n = 1; P.Xw = X; P.obs = obs; P.inv_sigma2 = &inv_sigma2;
    memcpy(P.pose, pose, sizeof(double) * 7);
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3]; P.bf = K[4];
    pose_opt S;
    memset(&S, 0, sizeof(S));
    S.P = &P;
    memcpy(S.T.q, pose, sizeof(double) * 4); memcpy(S.T.t, pose + 4, sizeof(double) * 3);
    double e3[3], Xc[3], J18[18];
    edge_error(&S, 0, e3, chi2, Xc);
    const int st = is_stereo(&P, 0);
    for (int i = 0; i < (st ? 3 : 2); i++) err[i] = e3[i];
    *depth_positive = Xc[2] > 0;
    pose_edge_jacobian(Xc, st, P.fx, P.fy, (double)(float)P.bf, J18);
    memcpy(J, J18, sizeof(double) * 6 * (st ? 3 : 2));
}
