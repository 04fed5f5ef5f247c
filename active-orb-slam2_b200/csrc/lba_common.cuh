// Device helpers shared by the local-BA kernels (lba.cu: one kernel per stage, host-driven Levenberg loop;
// lba_fused.cu: the whole optimize() call in one thread-block cluster).
#pragma once
#include <math.h>
#include "orbx_internal.cuh"

#define LBA_THREADS 256
#define LBA_SMEM_KF 64          // H_pp pre-reduction in shared memory up to this many free keyframes
#define LBA_SCHUR_CTAS 32
#define LBA_SCHUR_THREADS 512
#define LBA_SOLVE_THREADS 1024

struct LbaDev {
    int n_kf, n_pts, n_edges, np, n;         // np free keyframes, n = 6 np
    double *kf;          // [n_kf][7] quaternion (x,y,z,w), translation
    const int *kfidx;    // [n_kf] index in the reduced system or -1 (fixed)
    double *pt;          // [n_pts][3]
    const int *ptstart;  // [n_pts + 1] edges are sorted by landmark
    const int *ekf, *ept;
    const double *obs;   // [E][3]
    const double *info;  // [E]
    const uint8_t *stereo;
    uint8_t *level1;     // [E] excluded from the second round
    double *err, *chi2;  // [E][3], [E]  (the edge's stored _error / chi2())
    double *Hpl;         // [E][18]  6x3 row-major
    double *Hpp;         // [np][27]  21 upper-triangular entries of the 6x6 block, then b_p
    double *Hll;         // [n_pts][9]  6 upper-triangular entries of the 3x3 block, then b_l
    double *Hs, *bs, *xp, *xl;   // [n][n] (upper block triangle filled), [n], [n], [n_pts][3]
    double *scal;        // 0 chi2, 1 scale, 2 max diagonal, 3 solve ok
    double fx, fy, cx, cy, bf;
    float bf_f;
    double d_mono, d_stereo;     // Huber deltas (float sqrt(5.991), sqrt(7.815), Optimizer.cc:569-570)
    // work lists of the cluster kernel (lba_fused.cu), built on the host per window
    const int4 *kfe;             // edges ordered by (free) keyframe: edge, keyframe, landmark, stereo
    const int4 *kchunk;          // chunks of kfe: keyframe, first entry, entries
    const int *kf_cstart;        // [np + 1] chunks of a keyframe
    int n_kchunks;
    const int4 *pairs;           // (edge, edge, landmark) with pose_i <= pose_j, ordered by upper block
    const int4 *pchunk;          // chunks of pairs: block, first entry, entries, diagonal flag
    const int *blk_cstart;       // [nblk + 1] chunks of a block
    int n_pchunks;
    double *hppart, *part;       // [n_kchunks][27], [n_pchunks][42] partial sums
    double *dinv;                // [n_pts][10]  D^-1 (6) and D^-1 b_l (3) of the current trial, 16-byte aligned rows
};

__device__ __forceinline__ void quat_to_R(const double *q, double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

__device__ __forceinline__ double block_sum(double v, double *tmp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) tmp[w] = v;
    __syncthreads();
    double s = 0;
    if (w == 0) {
        s = lane < (int)(blockDim.x >> 5) ? tmp[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    __syncthreads();
    return s;   // valid in thread 0
}

// residual of edge e at the current estimates; returns the depth
__device__ __forceinline__ double edge_residual(const LbaDev &D, int e, const double R[9], const double *t, double Xc[3], double er[3]) {
    const double *X = D.pt + 3 * D.ept[e];
    for (int r = 0; r < 3; r++) Xc[r] = R[3 * r] * X[0] + R[3 * r + 1] * X[1] + R[3 * r + 2] * X[2] + t[r];
    const double *o = D.obs + 3 * e;
    if (!D.stereo[e]) {
        er[0] = o[0] - (Xc[0] / Xc[2] * D.fx + D.cx);
        er[1] = o[1] - (Xc[1] / Xc[2] * D.fy + D.cy);
        er[2] = 0;
    } else {   // cam_project keeps 1/z and bf in float (types_six_dof_expmap.cpp:150-157); `1.0f/trans_xyz[2]` divides in
               // double (the divisor is a double) and narrows once
        const float invz = __double2float_rn(1.0 / Xc[2]);
        const double u = Xc[0] * (double)invz * D.fx + D.cx;
        er[0] = o[0] - u;
        er[1] = o[1] - (Xc[1] * (double)invz * D.fy + D.cy);
        er[2] = o[2] - (u - (double)__fmul_rn(D.bf_f, invz));
    }
    return Xc[2];
}

// symmetric 3x3 inverse of (H_ll + lambda I); h = (xx, xy, xz, yy, yz, zz)
__device__ __forceinline__ void dinv3(const double *h, double lambda, double I[6]) {
    const double a = h[0] + lambda, b = h[1], c = h[2], e = h[3] + lambda, f = h[4], i = h[5] + lambda;
    const double c00 = e * i - f * f, c01 = c * f - b * i, c02 = b * f - c * e;
    const double id = 1.0 / (a * c00 + b * c01 + c * c02);
    I[0] = c00 * id; I[1] = c01 * id; I[2] = c02 * id;
    I[3] = (a * i - c * c) * id; I[4] = (b * c - a * f) * id; I[5] = (a * e - b * b) * id;
}

__device__ __forceinline__ int upper_block(int p1, int p2, int np) { return p1 * np - p1 * (p1 - 1) / 2 + (p2 - p1); }

// Eigen::Quaterniond(Matrix3d); static indexing only, so that everything stays in registers
__device__ __forceinline__ void R_to_quat(const double R[9], double q[4]) {
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else if (R[0] >= R[4] && R[0] >= R[8]) {             // i = 0, j = 1, k = 2 (Eigen: i = 0; if (m11 > m00) i = 1; if (m22 > m(i,i)) i = 2)
        t = sqrt(R[0] - R[4] - R[8] + 1.0);
        q[0] = 0.5 * t; t = 0.5 / t;
        q[3] = (R[7] - R[5]) * t; q[1] = (R[3] + R[1]) * t; q[2] = (R[6] + R[2]) * t;
    } else if (R[4] > R[0] && R[4] >= R[8]) {              // i = 1, j = 2, k = 0
        t = sqrt(R[4] - R[8] - R[0] + 1.0);
        q[1] = 0.5 * t; t = 0.5 / t;
        q[3] = (R[2] - R[6]) * t; q[2] = (R[7] + R[5]) * t; q[0] = (R[1] + R[3]) * t;
    } else {                                               // i = 2, j = 0, k = 1
        t = sqrt(R[8] - R[0] - R[4] + 1.0);
        q[2] = 0.5 * t; t = 0.5 / t;
        q[3] = (R[3] - R[1]) * t; q[0] = (R[2] + R[6]) * t; q[1] = (R[5] + R[7]) * t;
    }
}
__device__ __forceinline__ void quat_normalize(double q[4]) {
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= nrm; q[1] /= nrm; q[2] /= nrm; q[3] /= nrm;
}

// q / |q| with w >= 0: one rsqrt and four products instead of a square root and four divisions (this sits on the serial chain of every
// Levenberg trial)
__device__ __forceinline__ void quat_normalize_fast(double q[4]) {
    const double rn = rsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), sg = q[3] < 0 ? -rn : rn;
    q[0] *= sg; q[1] *= sg; q[2] *= sg; q[3] *= sg;
}

// T <- exp(u) * T  (VertexSE3Expmap::oplusImpl, SE3Quat::exp, SE3Quat::operator*); every loop unrolled, registers only.  Same
// formulas as g2o; reciprocals are taken once (1 / theta from one rsqrt, the quaternion norms likewise), so results agree with the
// oracle to rounding, not to the bit (the optimisers are held to 1e-4 relative).
__device__ __forceinline__ void se3_oplus(double *T, const double *u) {
    const double wx = u[0], wy = u[1], wz = u[2];
    const double t2 = wx * wx + wy * wy + wz * wz;
    const double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double O2[9], R[9], V[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) O2[3 * r + c] = O[3 * r] * O[c] + O[3 * r + 1] * O[3 + c] + O[3 * r + 2] * O[6 + c];
    double a = 1.0, b = 1.0, cc = 1.0;
    const bool small = !(t2 >= 0.00001 * 0.00001);         // SE3Quat::exp: R = V = I + Omega + Omega^2 below theta = 1e-5
    if (!small) {
        const double it = rsqrt(t2), theta = t2 * it, it2 = it * it;
        double sn, cs;
        sincos(theta, &sn, &cs);
        a = sn * it; b = (1 - cs) * it2; cc = (theta - sn) * (it2 * it);
    }
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        R[i] = id + a * O[i] + b * O2[i];
        V[i] = small ? R[i] : id + b * O[i] + cc * O2[i];
    }
    double eq[4], et[3];
    {   // Eigen::Quaterniond(R); the trace branch with 1 / sqrt taken once
        const double t = R[0] + R[4] + R[8];
        if (t > 0) {
            const double rs = rsqrt(t + 1.0), h = 0.5 * rs;
            eq[3] = 0.5 * (t + 1.0) * rs;
            eq[0] = (R[7] - R[5]) * h; eq[1] = (R[2] - R[6]) * h; eq[2] = (R[3] - R[1]) * h;
        } else R_to_quat(R, eq);
    }
    quat_normalize_fast(eq);
#pragma unroll
    for (int r = 0; r < 3; r++) et[r] = V[3 * r] * u[3] + V[3 * r + 1] * u[4] + V[3 * r + 2] * u[5];
    double nq[4];
    nq[3] = eq[3] * T[3] - eq[0] * T[0] - eq[1] * T[1] - eq[2] * T[2];
    nq[0] = eq[3] * T[0] + eq[0] * T[3] + eq[1] * T[2] - eq[2] * T[1];
    nq[1] = eq[3] * T[1] + eq[1] * T[3] + eq[2] * T[0] - eq[0] * T[2];
    nq[2] = eq[3] * T[2] + eq[2] * T[3] + eq[0] * T[1] - eq[1] * T[0];
    double Rq[9], nt[3];
    quat_to_R(eq, Rq);
#pragma unroll
    for (int r = 0; r < 3; r++) nt[r] = et[r] + Rq[3 * r] * T[4] + Rq[3 * r + 1] * T[5] + Rq[3 * r + 2] * T[6];
    quat_normalize_fast(nq);
    T[0] = nq[0]; T[1] = nq[1]; T[2] = nq[2]; T[3] = nq[3]; T[4] = nt[0]; T[5] = nt[1]; T[6] = nt[2];
}
