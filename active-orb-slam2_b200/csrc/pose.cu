// Pose-only optimisation: replaces Optimizer::PoseOptimization (reference src/Optimizer.cc:239-452) and the g2o pieces it
// drives (EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose, types_six_dof_expmap.cpp:266-364;
// BaseUnaryEdge::constructQuadraticForm, base_unary_edge.hpp:46-74; RobustKernelHuber; LinearSolverDense on the 6x6
// system; OptimizationAlgorithmLevenberg::solve, optimization_algorithm_levenberg.cpp:61-189).
//
// One CTA per frame, one launch per batch of frames: the four rounds, their Levenberg loops and the outlier
// classification all run on the device.  A pass over the observations is a grid-stride loop of the CTA's threads with the
// 21 + 6 sums of H and b (and the robust chi2) kept in registers, reduced by shuffles and one shared-memory step in a
// fixed order; thread 0 factorises the 6x6 system, applies exp(x) to the pose and takes the Levenberg decision that the
// other threads read back from shared memory.  All f64, like g2o.
#include "lba_common.cuh"
#include <float.h>

#define PO_THREADS 256
#define PO_WARPS (PO_THREADS / 32)
#define PO_NACC 28        // 21 upper-triangular entries of H, 6 of b, robust chi2

struct PoseProbDev {
    int off, n;
    double pose[7];
    double fx, fy, cx, cy, bf;
};
struct PoseOutDev {
    double pose[7];
    int n_inliers, n_bad, trials, pad;
};

struct orbx_pose {
    int device, max_obs, max_frames;
    double *d_Xw, *d_obs, *d_chi2; float *d_info; uint8_t *d_outlier;
    PoseProbDev *d_prob; PoseOutDev *d_out;
    int32_t *d_index;          // [max_obs] keypoint of every listed observation (device-resident entry point)
    // pinned staging, same layout
    uint8_t *h_arena, *d_arena; size_t arena_bytes;     // host entry point: one upload and one download per call, same layout both sides
    cudaStream_t stream;
    int last_launches;
};

struct PoseShared {
    double Rt[12];              // rotation (row-major) and translation of the current estimate
    double T[7], Tbak[7];
    double red[PO_WARPS][PO_NACC];
    double sum[PO_NACC];
    double H[21], b[6], x[6];
    double lambda, ni, currentChi, iniChi, rho;
    int again, stop, nbad_lm, trials, qmax;
};

__device__ __forceinline__ void po_set_pose(PoseShared &S) {
    quat_to_R(S.T, S.Rt);
    S.Rt[9] = S.T[4]; S.Rt[10] = S.T[5]; S.Rt[11] = S.T[6];
}

// reduce acc[0..cnt) over the CTA into S.sum (visible to every thread after the call)
template <int CNT>
__device__ __forceinline__ void po_reduce(PoseShared &S, double *acc) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < CNT; i++) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) S.red[w][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < CNT) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < PO_WARPS; k++) s += S.red[k][threadIdx.x];
        S.sum[threadIdx.x] = s;
    }
    __syncthreads();
}

// error of one observation at the estimate in S.Rt: e (3), chi2, camera-frame point
__device__ __forceinline__ double po_error(const PoseShared &S, const PoseProbDev &P, const double *X, const double *o, bool stereo,
                                           double info, double er[3], double Xc[3]) {
    const double *R = S.Rt;
    for (int r = 0; r < 3; r++) Xc[r] = R[3 * r] * X[0] + R[3 * r + 1] * X[1] + R[3 * r + 2] * X[2] + R[9 + r];
    if (!stereo) {
        er[0] = o[0] - (Xc[0] / Xc[2] * P.fx + P.cx);
        er[1] = o[1] - (Xc[1] / Xc[2] * P.fy + P.cy);
        er[2] = 0;
    } else {   // cam_project: `const float invz = 1.0f/trans_xyz[2]` (double quotient narrowed once), bf is the edge's double member
        const double invz = (double)__double2float_rn(1.0 / Xc[2]);
        const double u = Xc[0] * invz * P.fx + P.cx;
        er[0] = o[0] - u;
        er[1] = o[1] - (Xc[1] * invz * P.fy + P.cy);
        er[2] = o[2] - (u - P.bf * invz);
    }
    return info * (er[0] * er[0] + er[1] * er[1] + er[2] * er[2]);
}

// computeActiveErrors (+ linearizeOplus + constructQuadraticForm when BUILD): sums land in S.sum
template <bool BUILD>
__device__ __forceinline__ void po_pass(PoseShared &S, const PoseProbDev &P, const double *__restrict__ Xw, const double *__restrict__ obs,
                                        const float *__restrict__ info_f, const uint8_t *__restrict__ outlier, double *__restrict__ chi2,
                                        bool robust, double d_mono, double d_stereo) {
    double acc[PO_NACC];
#pragma unroll
    for (int i = 0; i < PO_NACC; i++) acc[i] = 0;
    for (int e = threadIdx.x; e < P.n; e += PO_THREADS) {
        if (outlier[e]) continue;                          // level 1: not part of this round
        const double *X = Xw + 3 * (size_t)e, *o = obs + 3 * (size_t)e;
        const bool st = !(o[2] < 0);                      // mvuRight[i] < 0 -> monocular edge (Optimizer.cc:281)
        const double info = (double)info_f[e];
        double er[3], Xc[3];
        const double c = po_error(S, P, X, o, st, info, er, Xc);
        chi2[e] = c;
        double rho1 = 1.0, cr = c;
        if (robust) {
            const double d = st ? d_stereo : d_mono, dsqr = (double)(float)(d * d);   // RobustKernelHuber keeps dsqr in a float member (robust_kernel_impl.h:84)
            if (c > dsqr) { const double sq = sqrt(c); cr = 2 * sq * d - dsqr; rho1 = d / sq; }
        }
        acc[27] += cr;
        if (!BUILD) continue;
        const double x = Xc[0], y = Xc[1], invz = 1.0 / Xc[2], invz_2 = invz * invz, fx = P.fx, fy = P.fy, bf = P.bf;
        double J[18];
        J[0] = x * y * invz_2 * fx; J[1] = -(1 + (x * x * invz_2)) * fx; J[2] = y * invz * fx; J[3] = -invz * fx; J[4] = 0; J[5] = x * invz_2 * fx;
        J[6] = (1 + y * y * invz_2) * fy; J[7] = -x * y * invz_2 * fy; J[8] = -x * invz * fy; J[9] = 0; J[10] = -invz * fy; J[11] = y * invz_2 * fy;
        if (st) { J[12] = J[0] - bf * y * invz_2; J[13] = J[1] + bf * x * invz_2; J[14] = J[2]; J[15] = J[3]; J[16] = 0; J[17] = J[5] - bf * invz_2; }
        else {
#pragma unroll
            for (int i = 12; i < 18; i++) J[i] = 0;
        }
        const double w = rho1 * info;
        const double w0 = info * er[0] * rho1, w1 = info * er[1] * rho1, w2 = info * er[2] * rho1;
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = a; b < 6; b++) acc[k++] += w * (J[a] * J[b] + J[6 + a] * J[6 + b] + J[12 + a] * J[12 + b]);
#pragma unroll
        for (int a = 0; a < 6; a++) acc[21 + a] -= J[a] * w0 + J[6 + a] * w1 + J[12 + a] * w2;
    }
    if (BUILD) po_reduce<PO_NACC>(S, acc);
    else {
        double one[1] = {acc[27]};
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        double v = one[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) S.red[w][27] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0;
            for (int k = 0; k < PO_WARPS; k++) s += S.red[k][27];
            S.sum[27] = s;
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int po2_solve6(const double (&Hu)[21], const double (&b)[6], double lambda, double (&x)[6]);

__global__ void __launch_bounds__(PO_THREADS)
k_pose_optimize(const PoseProbDev *__restrict__ probs, const double *__restrict__ Xw_all, const double *__restrict__ obs_all,
                const float *__restrict__ info_all, uint8_t *__restrict__ outlier_all, double *__restrict__ chi2_all,
                PoseOutDev *__restrict__ out, int iterations) {
    __shared__ PoseShared S;
    __shared__ PoseProbDev P;
    const int tid = threadIdx.x;
    if (tid == 0) P = probs[blockIdx.x];
    __syncthreads();
    const double *Xw = Xw_all + 3 * (size_t)P.off, *obs = obs_all + 3 * (size_t)P.off;
    const float *info = info_all + P.off;
    uint8_t *outlier = outlier_all + P.off;
    double *chi2 = chi2_all + P.off;
    const int n = P.n;
    const double d_mono = (double)(float)sqrt(5.991), d_stereo = (double)(float)sqrt(7.815);   // const float deltaMono / deltaStereo
    for (int e = tid; e < n; e += PO_THREADS) outlier[e] = 0;
    if (tid == 0) { for (int i = 0; i < 7; i++) S.T[i] = P.pose[i]; S.trials = 0; }
    __syncthreads();
    int n_bad = 0;
    if (n >= 3) {                                                              // :355-356
        bool robust = true;
        for (int round = 0; round < 4; round++) {
            if (tid == 0) { for (int i = 0; i < 7; i++) S.T[i] = P.pose[i]; po_set_pose(S); }   // setEstimate(mTcw), :366
            __syncthreads();
            if (n - n_bad > 0) {                                               // otherwise "0 vertices to optimize"
                // ---- OptimizationAlgorithmLevenberg, `iterations` iterations ----
                if (tid == 0) { S.stop = 0; S.nbad_lm = 0; }
                for (int it = 0; it < iterations; it++) {
                    po_pass<true>(S, P, Xw, obs, info, outlier, chi2, robust, d_mono, d_stereo);
                    if (tid == 0) {
                        for (int i = 0; i < 21; i++) S.H[i] = S.sum[i];
                        for (int i = 0; i < 6; i++) S.b[i] = S.sum[21 + i];
                        S.currentChi = S.iniChi = S.sum[27];
                        if (it == 0) {
                            double mx = 0;
                            int k = 0;
                            for (int a = 0; a < 6; a++) { mx = fmax(mx, fabs(S.H[k])); k += 6 - a; }
                            S.lambda = 1e-5 * mx; S.ni = 2; S.nbad_lm = 0;
                        }
                        S.rho = 0; S.qmax = 0;
                    }
                    __syncthreads();
                    int again;
                    do {
                        if (tid == 0) {
                            for (int i = 0; i < 7; i++) S.Tbak[i] = S.T[i];   // push
                            const int ok = po2_solve6(S.H, S.b, S.lambda, S.x);
                            if (!ok) for (int i = 0; i < 6; i++) S.x[i] = 0;
                            if (ok) { se3_oplus(S.T, S.x); po_set_pose(S); }
                            S.again = ok;                                     // reused as "solve ok" until the decision below
                        }
                        __syncthreads();
                        po_pass<false>(S, P, Xw, obs, info, outlier, chi2, robust, d_mono, d_stereo);
                        if (tid == 0) {
                            double tempChi = S.sum[27];
                            if (!S.again) tempChi = DBL_MAX;
                            double rho = S.currentChi - tempChi, scale = 0;
                            for (int j = 0; j < 6; j++) scale += S.x[j] * (S.lambda * S.x[j] + S.b[j]);
                            scale += 1e-3;
                            rho /= scale;
                            S.trials++;
                            if (rho > 0 && isfinite(tempChi)) {
                                const double t = 2 * rho - 1;
                                double alpha = 1. - t * t * t;
                                alpha = fmin(alpha, 2. / 3.);
                                S.lambda *= fmax(1. / 3., alpha);
                                S.ni = 2;
                                S.currentChi = tempChi;
                            } else {
                                S.lambda *= S.ni; S.ni *= 2;
                                for (int i = 0; i < 7; i++) S.T[i] = S.Tbak[i];   // pop
                                po_set_pose(S);
                            }
                            S.rho = rho;
                            S.qmax++;
                            S.again = rho < 0 && S.qmax < 10;
                        }
                        __syncthreads();
                        again = S.again;
                        __syncthreads();                                      // thread 0 reuses S.again in the next trial
                    } while (again);
                    if (tid == 0) {
                        int stop = 0;
                        if (S.qmax == 10 || S.rho == 0) stop = 1;
                        else {
                            if ((S.iniChi - S.currentChi) * 1e3 < S.iniChi) S.nbad_lm++; else S.nbad_lm = 0;
                            if (S.nbad_lm >= 3) stop = 1;
                        }
                        S.stop = stop;
                    }
                    __syncthreads();
                    const int stop = S.stop;
                    __syncthreads();
                    if (stop) break;
                }
            }
            // ---- classification, :371-417 ----
            const float th_mono = 5.991f, th_stereo = 7.815f;
            double cnt[1] = {0};
            for (int e = tid; e < n; e += PO_THREADS) {
                const double *o = obs + 3 * (size_t)e;
                const bool st = !(o[2] < 0);
                double c = chi2[e];
                if (outlier[e]) {                                             // left out of this round: e->computeError()
                    double er[3], Xc[3];
                    c = po_error(S, P, Xw + 3 * (size_t)e, o, st, (double)info[e], er, Xc);
                    chi2[e] = c;
                }
                const bool bad = (float)c > (st ? th_stereo : th_mono);
                outlier[e] = bad ? 1 : 0;
                cnt[0] += bad ? 1.0 : 0.0;
            }
            po_reduce<1>(S, cnt);
            n_bad = (int)S.sum[0];
            __syncthreads();
            if (round == 2) robust = false;                                   // setRobustKernel(0), :391, :416
            if (n < 10) break;                                                // optimizer.edges().size() < 10, :419
        }
    }
    if (tid == 0) {
        PoseOutDev o;
        for (int i = 0; i < 7; i++) o.pose[i] = S.T[i];
        o.n_bad = n_bad; o.n_inliers = n >= 3 ? n - n_bad : 0; o.trials = S.trials; o.pad = 0;
        out[blockIdx.x] = o;
    }
}

// ---- the cluster form (the one that normally runs; the kernel above takes frames too large to stage in shared memory) -----------------
// Shared pieces first: a transposing butterfly for the 28 sums of a pass (28 64-bit shuffles per warp instead of 140), the error of one
// observation with the pose in registers, and the 6x6 solve.
// every lane ends with ONE of the 28 warp totals: lane l holds total number po2_slot(l) (lanes with slot 28 hold padding)
__device__ __forceinline__ double po2_bfly28(const double (&v)[PO_NACC], int lane) {
    double a[14], b[7], c[4], d[2];
    bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 14; i++) {
        const double send = up ? v[i] : v[14 + i], keep = up ? v[14 + i] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    up = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        const double send = up ? a[i] : a[7 + i], keep = up ? a[7 + i] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    up = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double lo = b[i], hi = i < 3 ? b[4 + (i < 3 ? i : 0)] : 0.0;
        const double send = up ? lo : hi, keep = up ? hi : lo;
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    up = (lane & 2) != 0;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double send = up ? c[i] : c[2 + i], keep = up ? c[2 + i] : c[i];
        d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    up = (lane & 1) != 0;
    const double send = up ? d[0] : d[1], keep = up ? d[1] : d[0];
    return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}
__device__ __forceinline__ int po2_slot(int lane) {
    if ((lane & 7) == 7) return PO_NACC;                   // the padded fourth entry of the upper 3-element half
    return ((lane & 16) ? 14 : 0) + ((lane & 8) ? 7 : 0) + ((lane & 4) ? 4 : 0) + ((lane & 2) ? 2 : 0) + (lane & 1);
}

struct Po2Cam { double fx, fy, cx, cy, bf; };

__device__ __forceinline__ double po2_error(const double (&R)[12], const Po2Cam &K, const double *X, const double *o, bool stereo, double info,
                                            double er[3], double Xc[3]) {
#pragma unroll
    for (int r = 0; r < 3; r++) Xc[r] = R[3 * r] * X[0] + R[3 * r + 1] * X[1] + R[3 * r + 2] * X[2] + R[9 + r];
    if (!stereo) {
        er[0] = o[0] - (Xc[0] / Xc[2] * K.fx + K.cx);
        er[1] = o[1] - (Xc[1] / Xc[2] * K.fy + K.cy);
        er[2] = 0;
    } else {   // cam_project: `const float invz = 1.0f/trans_xyz[2]` (double quotient narrowed once), bf is the edge's double member
        const double invz = (double)__double2float_rn(1.0 / Xc[2]);
        const double u = Xc[0] * invz * K.fx + K.cx;
        er[0] = o[0] - u;
        er[1] = o[1] - (Xc[1] * invz * K.fy + K.cy);
        er[2] = o[2] - (u - K.bf * invz);
    }
    return info * (er[0] * er[0] + er[1] * er[1] + er[2] * er[2]);
}

// (H + lambda I) x = b, H as its 21 upper-triangular entries; L L^T with 1 / L_jj kept from one rsqrt per column; 0 if not positive definite
__device__ __forceinline__ int po2_solve6(const double (&Hu)[21], const double (&b)[6], double lambda, double (&x)[6]) {
    // every loop runs 0..5 with compile-time guards: loops whose bounds depend on an outer unrolled index are only partially unrolled by
    // nvcc, which would put A into local memory on the one chain every thread is waiting for
    double A[36], inv[6];
    {
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int c = 0; c < 6; c++)
                if (c >= a) { const double v = Hu[k] + (a == c ? lambda : 0.0); A[6 * a + c] = v; A[6 * c + a] = v; k++; }
    }
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = A[j * 6 + j];
#pragma unroll
        for (int q = 0; q < 6; q++)
            if (q < j) d -= A[j * 6 + q] * A[j * 6 + q];
        ok = ok && (d > 0);
        inv[j] = rsqrt(d);
#pragma unroll
        for (int i = 0; i < 6; i++)
            if (i > j) {
                double s = A[i * 6 + j];
#pragma unroll
                for (int q = 0; q < 6; q++)
                    if (q < j) s -= A[i * 6 + q] * A[j * 6 + q];
                A[i * 6 + j] = s * inv[j];
            }
    }
    if (!ok) return 0;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = b[i];
#pragma unroll
        for (int q = 0; q < 6; q++)
            if (q < i) s -= A[i * 6 + q] * y[q];
        y[i] = s * inv[i];
    }
#pragma unroll
    for (int ii = 0; ii < 6; ii++) {
        const int i = 5 - ii;
        double s = y[i];
#pragma unroll
        for (int q = 0; q < 6; q++)
            if (q > i) s -= A[q * 6 + i] * y[q];
        y[i] = s * inv[i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = y[i];
    return 1;
}

__device__ __forceinline__ void po2_set_pose(const double (&T)[7], double (&R)[12]) {
    quat_to_R(T, R);
    R[9] = T[4]; R[10] = T[5]; R[11] = T[6];
}

// ---- a thread-block cluster per frame ------------------------------------------------------------------------------------------------
// ncu on the one-block form (profiles/r2_af_pose.txt): one frame of 400 observations runs 59 Levenberg trials in 650 k cycles on ONE SM,
// issue-active 26 %, no memory stalls worth naming; 82 k instructions per warp, all on one dependent f64 chain: a pass over the
// observations (the SM's 64 FP64 lanes are the resource), then thread 0's solve / exp map / decision with everybody else at a barrier.
// Here a cluster of C CTAs (C = 8, 4, 2 or 1: the largest that still gives every frame of the batch its own SMs and at least one
// observation per thread) shares a frame: CTA r stages observations [n r / C, n (r + 1) / C) in its shared memory once (nothing is read
// from global memory inside the Levenberg loop) and runs the passes over its share; the partial sums are exchanged through distributed
// shared memory (each CTA stores its 28 sums into every peer's buffer, one cluster barrier) and added in rank order by everybody; and
// EVERY thread carries the pose, the 6x6 system and the Levenberg state in registers and repeats the (identical) solve / exp map /
// decision, so all CTAs take the same decisions from the same numbers without a broadcast and a trial needs only the barriers of its
// reductions.  One warp per scheduler (128 threads) makes that redundancy free: the FP64 lanes it uses would idle.
// Measured and dropped: building the next iteration's quadratic form inside the pass that evaluates a trial (one pass per accepted
// iteration instead of two, but every rejected trial then pays the 28-sum exchange: 0.189 -> 0.200 ms at 400 observations / 58 trials,
// 0.160 -> 0.151 ms at 1500 / 29, 0.367 -> 0.375 ms for 64 frames x 500; profiles/r2_ah_pose_ab.txt, kernels 4 / 5).
#define P3_THREADS 128
#define P3_WARPS (P3_THREADS / 32)
#define P3_MAXC 8
#define P3_CAP_MAX 3072            // observations a CTA can stage: 61 bytes each

struct Po3Shared {
    double red[P3_WARPS][PO_NACC];
    double part[2][P3_MAXC][PO_NACC];      // [buffer][rank]: written by every CTA of the cluster (its own copy included)
    double sum[PO_NACC];
    double red1[P3_WARPS];
    double part1[2][P3_MAXC];
};

__device__ __forceinline__ void po3_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned po3_cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned po3_cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// store a double into the same shared-memory variable of CTA `rank` of the cluster
__device__ __forceinline__ void po3_store_remote(double *local, unsigned rank, double v) {
    unsigned a = (unsigned)__cvta_generic_to_shared(local), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}

struct Po3Obs {                    // this CTA's observations, structure of arrays in shared memory
    const double *X, *O; double *chi2; const float *info; uint8_t *outl; int cap, n;
};

template <bool BUILD>
__device__ __forceinline__ void po3_pass(Po3Shared &S, int &b28, int &b1, unsigned rank, unsigned C, const double (&R)[12], const Po2Cam &K,
                                         const Po3Obs &Q, bool robust, double d_mono, double d_stereo, double (&out)[PO_NACC]) {
    double acc[PO_NACC];
#pragma unroll
    for (int i = 0; i < PO_NACC; i++) acc[i] = 0;
    for (int j = threadIdx.x; j < Q.n; j += P3_THREADS) {
        if (Q.outl[j]) continue;                           // level 1: not part of this round
        const double X[3] = {Q.X[j], Q.X[Q.cap + j], Q.X[2 * Q.cap + j]}, o[3] = {Q.O[j], Q.O[Q.cap + j], Q.O[2 * Q.cap + j]};
        const bool st = !(o[2] < 0);                      // mvuRight[i] < 0 -> monocular edge (Optimizer.cc:281)
        const double info = (double)Q.info[j];
        double er[3], Xc[3];
        const double c = po2_error(R, K, X, o, st, info, er, Xc);
        Q.chi2[j] = c;
        double rho1 = 1.0, cr = c;
        if (robust) {
            const double d = st ? d_stereo : d_mono, dsqr = (double)(float)(d * d);   // RobustKernelHuber keeps dsqr in a float member
            if (c > dsqr) { const double rs = rsqrt(c), sq = c * rs; cr = 2 * sq * d - dsqr; rho1 = d * rs; }
        }
        acc[27] += cr;
        if (!BUILD) continue;
        const double x = Xc[0], y = Xc[1], invz = 1.0 / Xc[2], invz_2 = invz * invz, fx = K.fx, fy = K.fy, bf = K.bf;
        double J[18];
        J[0] = x * y * invz_2 * fx; J[1] = -(1 + (x * x * invz_2)) * fx; J[2] = y * invz * fx; J[3] = -invz * fx; J[4] = 0; J[5] = x * invz_2 * fx;
        J[6] = (1 + y * y * invz_2) * fy; J[7] = -x * y * invz_2 * fy; J[8] = -x * invz * fy; J[9] = 0; J[10] = -invz * fy; J[11] = y * invz_2 * fy;
        if (st) { J[12] = J[0] - bf * y * invz_2; J[13] = J[1] + bf * x * invz_2; J[14] = J[2]; J[15] = J[3]; J[16] = 0; J[17] = J[5] - bf * invz_2; }
        else {
#pragma unroll
            for (int i = 12; i < 18; i++) J[i] = 0;
        }
        const double w = rho1 * info;
        const double w0 = info * er[0] * rho1, w1 = info * er[1] * rho1, w2 = info * er[2] * rho1;
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = a; b < 6; b++) acc[k++] += w * (J[a] * J[b] + J[6 + a] * J[6 + b] + J[12 + a] * J[12 + b]);
#pragma unroll
        for (int a = 0; a < 6; a++) acc[21 + a] -= J[a] * w0 + J[6 + a] * w1 + J[12 + a] * w2;
    }
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (BUILD) {
        const double mine = po2_bfly28(acc, lane);
        const int slot = po2_slot(lane);
        if (slot < PO_NACC) S.red[w][slot] = mine;
        __syncthreads();
        if (tid < PO_NACC) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < P3_WARPS; k++) s += S.red[k][tid];
            for (unsigned r = 0; r < C; r++) po3_store_remote(&S.part[b28][rank][tid], r, s);
        }
        po3_cluster_sync();
        if (tid < PO_NACC) {
            double s = 0;
            for (unsigned r = 0; r < C; r++) s += S.part[b28][r][tid];        // rank order: the same sum in every CTA
            S.sum[tid] = s;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PO_NACC; i++) out[i] = S.sum[i];
        b28 ^= 1;
    } else {
        double v = acc[27];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) S.red1[w] = v;
        __syncthreads();
        if (tid < (int)C) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < P3_WARPS; k++) s += S.red1[k];
            po3_store_remote(&S.part1[b1][rank], (unsigned)tid, s);
        }
        po3_cluster_sync();
        double s = 0;
        for (unsigned r = 0; r < C; r++) s += S.part1[b1][r];
        out[27] = s;
        b1 ^= 1;
    }
}

__global__ void __launch_bounds__(P3_THREADS, 1)
k_pose_optimize3(const PoseProbDev *__restrict__ probs, const double *__restrict__ Xw_all, const double *__restrict__ obs_all,
                 const float *__restrict__ info_all, uint8_t *__restrict__ outlier_all, PoseOutDev *__restrict__ out, int iterations, int cap) {
    extern __shared__ __align__(16) uint8_t po3_dyn[];
    __shared__ Po3Shared S;
    __shared__ PoseProbDev P;
    const int tid = threadIdx.x;
    const unsigned rank = po3_cluster_rank(), C = po3_cluster_size();
    const int frame = blockIdx.x / C;
    if (tid == 0) P = probs[frame];
    __syncthreads();
    const int n = P.n;
    const int e0 = (int)((long long)n * rank / C), e1 = (int)((long long)n * (rank + 1) / C);
    double *sX = reinterpret_cast<double *>(po3_dyn), *sO = sX + 3 * cap, *sChi = sO + 3 * cap;
    float *sInfo = reinterpret_cast<float *>(sChi + cap);
    uint8_t *sOut = reinterpret_cast<uint8_t *>(sInfo + cap);
    Po3Obs Q = {sX, sO, sChi, sInfo, sOut, cap, e1 - e0};
    {
        const double *Xw = Xw_all + 3 * ((size_t)P.off + e0), *obs = obs_all + 3 * ((size_t)P.off + e0);
        const float *info = info_all + P.off + e0;
        for (int i = tid; i < 3 * Q.n; i += P3_THREADS) {
            const int j = i / 3, c = i - 3 * j;
            sX[c * cap + j] = Xw[i];
            sO[c * cap + j] = obs[i];
        }
        for (int j = tid; j < Q.n; j += P3_THREADS) { sInfo[j] = info[j]; sOut[j] = 0; sChi[j] = 0; }
    }
    __syncthreads();
    po3_cluster_sync();              // every CTA of the cluster is running before anybody stores into a peer's shared memory
    const Po2Cam K = {P.fx, P.fy, P.cx, P.cy, P.bf};
    const double d_mono = (double)(float)sqrt(5.991), d_stereo = (double)(float)sqrt(7.815);   // const float deltaMono / deltaStereo
    double T[7], R[12], sums[PO_NACC];
#pragma unroll
    for (int i = 0; i < 7; i++) T[i] = P.pose[i];
    int n_bad = 0, trials = 0, b28 = 0, b1 = 0;
    if (n >= 3) {                                                              // :355-356
        bool robust = true;
        for (int round = 0; round < 4; round++) {
#pragma unroll
            for (int i = 0; i < 7; i++) T[i] = P.pose[i];                      // setEstimate(mTcw), :366
            po2_set_pose(T, R);
            if (n - n_bad > 0) {                                               // otherwise "0 vertices to optimize"
                // ---- OptimizationAlgorithmLevenberg, `iterations` iterations ----
                double H[21], b[6], x[6], chiB = 0, lambda = 0, ni = 2;
                int nbad_lm = 0;
                for (int it = 0; it < iterations; it++) {
                    po3_pass<true>(S, b28, b1, rank, C, R, K, Q, robust, d_mono, d_stereo, sums);
#pragma unroll
                    for (int i = 0; i < 21; i++) H[i] = sums[i];
#pragma unroll
                    for (int i = 0; i < 6; i++) b[i] = sums[21 + i];
                    chiB = sums[27];
                    double currentChi = chiB;
                    const double iniChi = chiB;
                    if (it == 0) {
                        const double mx = fmax(fmax(fmax(fabs(H[0]), fabs(H[6])), fmax(fabs(H[11]), fabs(H[15]))), fmax(fabs(H[18]), fabs(H[20])));
                        lambda = 1e-5 * mx; ni = 2; nbad_lm = 0;
                    }
                    double rho = 0;
                    int qmax = 0;
                    do {
                        double Tbak[7];
#pragma unroll
                        for (int i = 0; i < 7; i++) Tbak[i] = T[i];            // push
                        const int ok = po2_solve6(H, b, lambda, x);
                        if (!ok) {
#pragma unroll
                            for (int i = 0; i < 6; i++) x[i] = 0;
                        } else { se3_oplus(T, x); po2_set_pose(T, R); }
                        po3_pass<false>(S, b28, b1, rank, C, R, K, Q, robust, d_mono, d_stereo, sums);
                        double tempChi = sums[27];
                        if (!ok) tempChi = DBL_MAX;
                        double scale = 0;
#pragma unroll
                        for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
                        scale += 1e-3;
                        rho = (currentChi - tempChi) / scale;
                        trials++;
                        if (rho > 0 && isfinite(tempChi)) {
                            const double t = 2 * rho - 1;
                            double alpha = 1. - t * t * t;
                            alpha = fmin(alpha, 2. / 3.);
                            lambda *= fmax(1. / 3., alpha);
                            ni = 2;
                            currentChi = tempChi;
                        } else {
                            lambda *= ni; ni *= 2;
#pragma unroll
                            for (int i = 0; i < 7; i++) T[i] = Tbak[i];        // pop
                            po2_set_pose(T, R);
                        }
                        qmax++;
                    } while (rho < 0 && qmax < 10);
                    bool stop = false;
                    if (qmax == 10 || rho == 0) stop = true;
                    else {
                        if ((iniChi - currentChi) * 1e3 < iniChi) nbad_lm++; else nbad_lm = 0;
                        if (nbad_lm >= 3) stop = true;
                    }
                    if (stop) break;
                }
            }
            // ---- classification, :371-417 ----
            const float th_mono = 5.991f, th_stereo = 7.815f;
            int cnt = 0;
            for (int j = tid; j < Q.n; j += P3_THREADS) {
                const double X[3] = {Q.X[j], Q.X[cap + j], Q.X[2 * cap + j]}, o[3] = {Q.O[j], Q.O[cap + j], Q.O[2 * cap + j]};
                const bool st = !(o[2] < 0);
                double c = sChi[j];
                if (sOut[j]) {                                                // left out of this round: e->computeError()
                    double er[3], Xc[3];
                    c = po2_error(R, K, X, o, st, (double)sInfo[j], er, Xc);
                    sChi[j] = c;
                }
                const bool bad = (float)c > (st ? th_stereo : th_mono);
                sOut[j] = bad ? 1 : 0;
                cnt += bad ? 1 : 0;
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if ((tid & 31) == 0) S.red1[tid >> 5] = (double)cnt;
            __syncthreads();
            if (tid < (int)C) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < P3_WARPS; k++) s += S.red1[k];
                po3_store_remote(&S.part1[b1][rank], (unsigned)tid, s);
            }
            po3_cluster_sync();
            n_bad = 0;
            for (unsigned r = 0; r < C; r++) n_bad += (int)S.part1[b1][r];
            b1 ^= 1;
            if (round == 2) robust = false;                                   // setRobustKernel(0), :391, :416
            if (n < 10) break;                                                // optimizer.edges().size() < 10, :419
        }
    }
    uint8_t *outlier = outlier_all + P.off + e0;
    for (int j = tid; j < Q.n; j += P3_THREADS) outlier[j] = sOut[j];
    if (tid == 0 && rank == 0) {
        PoseOutDev o;
        for (int i = 0; i < 7; i++) o.pose[i] = T[i];
        o.n_bad = n_bad; o.n_inliers = n >= 3 ? n - n_bad : 0; o.trials = trials; o.pad = 0;
        out[frame] = o;
    }
    po3_cluster_sync();                                                       // no CTA leaves while a peer may still store into its shared memory
}

// ORBX_POSE_KERNEL = 1 forces the one-block form (A/B runs).  n_max = upper bound of a frame's observations: the cluster form stages a
// frame's share in shared memory, larger frames take the one-block form.
static cudaError_t po_launch(int n_frames, int n_max, cudaStream_t s, const PoseProbDev *probs, const double *Xw, const double *obs,
                             const float *info, uint8_t *outlier, double *chi2, PoseOutDev *out) {
    static int one_block = -1;
    if (one_block < 0) { const char *e = getenv("ORBX_POSE_KERNEL"); one_block = e && atoi(e) == 1; }
    int C = 1;
    while (C < P3_MAXC && n_frames * 2 * C <= 148) C *= 2;          // every frame of the batch keeps its own SMs
    while (C > 1 && n_max <= P3_THREADS * (C / 2)) C /= 2;             // no more CTAs than one observation per thread asks for
    const int cap = ((n_max + C - 1) / C + 1 + 31) & ~31;
    if (one_block || cap > P3_CAP_MAX) {
        k_pose_optimize<<<n_frames, PO_THREADS, 0, s>>>(probs, Xw, obs, info, outlier, chi2, out, 10);
        return cudaGetLastError();
    }
    {   // the shared-memory limit of a kernel is a per-device attribute
        static bool raised[64] = {};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64 || !raised[dev]) {
            if ((e = ORBX_RAISE_SMEM(k_pose_optimize3)) != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) raised[dev] = true;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_frames * C)); cfg.blockDim = dim3(P3_THREADS);
    cfg.dynamicSmemBytes = (size_t)cap * 61 + 64; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_pose_optimize3, probs, Xw, obs, info, outlier, out, 10, cap);
}

// ---- host side ----------------------------------------------------------------------------------------------------------
extern "C" void orbx_pose_destroy(orbx_pose *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_Xw); cudaFree(h->d_obs); cudaFree(h->d_chi2); cudaFree(h->d_info); cudaFree(h->d_outlier);
    cudaFree(h->d_prob); cudaFree(h->d_out); cudaFree(h->d_index); cudaFree(h->d_arena);
    if (h->h_arena) cudaFreeHost(h->h_arena);
    if (h->stream) cudaStreamDestroy(h->stream);
    free(h);
}

extern "C" orbx_status orbx_pose_create(orbx_pose **out, int max_observations, int max_frames, int device) {
    if (!out) return ORBX_ERR_INVALID;
    *out = nullptr;
    if (max_observations < 1 || max_frames < 1) {
        orbx_set_error("orbx_pose_create: bad argument");
        return ORBX_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        orbx_set_error("no CUDA device %d (%d visible)", device, ndev);
        return ORBX_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    ORBX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        orbx_set_error("device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
        return ORBX_ERR_NO_DEVICE;
    }
    ORBX_CUDA(cudaSetDevice(device));
    orbx_pose *h = (orbx_pose *)calloc(1, sizeof(orbx_pose));
    if (!h) return ORBX_ERR_NOMEM;
    h->device = device; h->max_obs = max_observations; h->max_frames = max_frames;
    const size_t no = (size_t)max_observations, nf = (size_t)max_frames;
    h->arena_bytes = no * (24 + 24 + 4 + 1) + nf * (sizeof(PoseProbDev) + sizeof(PoseOutDev)) + 256;
    cudaError_t ce = cudaSuccess;
#define TRY(x) if (ce == cudaSuccess) ce = (x)
    TRY(cudaMalloc((void **)&h->d_Xw, 24 * no));
    TRY(cudaMalloc((void **)&h->d_obs, 24 * no));
    TRY(cudaMalloc((void **)&h->d_chi2, 8 * no));
    TRY(cudaMalloc((void **)&h->d_info, 4 * no));
    TRY(cudaMalloc((void **)&h->d_outlier, no));
    TRY(cudaMalloc((void **)&h->d_prob, sizeof(PoseProbDev) * nf));
    TRY(cudaMalloc((void **)&h->d_out, sizeof(PoseOutDev) * nf));
    TRY(cudaMalloc((void **)&h->d_index, sizeof(int32_t) * no));
    TRY(cudaMallocHost((void **)&h->h_arena, h->arena_bytes));
    TRY(cudaMalloc((void **)&h->d_arena, h->arena_bytes));
    TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
#undef TRY
    if (ce != cudaSuccess) {
        orbx_set_error("orbx_pose_create: %s", cudaGetErrorString(ce));
        orbx_pose_destroy(h);
        return ORBX_ERR_CUDA;
    }
    *out = h;
    return ORBX_OK;
}

extern "C" orbx_status orbx_pose_optimize_host(orbx_pose *h, const orbx_pose_problem *probs, int n_frames, orbx_pose_result *res) {
    if (!h || n_frames < 0 || (n_frames && (!probs || !res))) return ORBX_ERR_INVALID;
    h->last_launches = 0;
    if (n_frames == 0) return ORBX_OK;
    if (n_frames > h->max_frames) {
        orbx_set_error("orbx_pose: %d frames, handle was created for %d", n_frames, h->max_frames);
        return ORBX_ERR_CAPACITY;
    }
    size_t total = 0;
    for (int f = 0; f < n_frames; f++) {
        const orbx_pose_problem &p = probs[f];
        if (p.n < 0 || (p.n && (!p.Xw || !p.obs || !p.inv_sigma2 || !res[f].outlier))) return ORBX_ERR_INVALID;
        total += (size_t)p.n;
    }
    if (total > (size_t)h->max_obs) {
        orbx_set_error("orbx_pose: %zu observations in the batch, handle was created for %d", total, h->max_obs);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaSetDevice(h->device));
    // pinned arena: Xw | obs | info | problems
    const size_t nt = total ? total : 1;
    double *aX = reinterpret_cast<double *>(h->h_arena), *aO = aX + 3 * nt;
    float *aI = reinterpret_cast<float *>(aO + 3 * nt);
    PoseProbDev *aP = reinterpret_cast<PoseProbDev *>(reinterpret_cast<uint8_t *>(aI) + ((4 * nt + 15) & ~(size_t)15));
    size_t off = 0;
    for (int f = 0; f < n_frames; f++) {
        const orbx_pose_problem &p = probs[f];
        memcpy(aX + 3 * off, p.Xw, sizeof(double) * 3 * p.n);
        memcpy(aO + 3 * off, p.obs, sizeof(double) * 3 * p.n);
        memcpy(aI + off, p.inv_sigma2, sizeof(float) * p.n);
        PoseProbDev &d = aP[f];
        d.off = (int)off; d.n = p.n;
        for (int i = 0; i < 7; i++) d.pose[i] = p.pose[i];
        d.fx = p.fx; d.fy = p.fy; d.cx = p.cx; d.cy = p.cy; d.bf = p.bf;
        off += (size_t)p.n;
    }
    cudaStream_t s = h->stream;
    // results come back through the same arena (after the problems)
    PoseOutDev *aR = reinterpret_cast<PoseOutDev *>(aP + n_frames);
    uint8_t *aB = reinterpret_cast<uint8_t *>(aR + n_frames);
    const size_t in_bytes = (size_t)(reinterpret_cast<uint8_t *>(aR) - h->h_arena), out_bytes = sizeof(PoseOutDev) * n_frames + nt;
    uint8_t *dA = h->d_arena;
    auto dev = [&](const void *host) { return dA + (reinterpret_cast<const uint8_t *>(host) - h->h_arena); };
    ORBX_CUDA(cudaMemcpyAsync(dA, h->h_arena, in_bytes, cudaMemcpyHostToDevice, s));
    int n_max = 0;
    for (int f = 0; f < n_frames; f++) n_max = probs[f].n > n_max ? probs[f].n : n_max;
    ORBX_CUDA(po_launch(n_frames, n_max, s, reinterpret_cast<const PoseProbDev *>(dev(aP)), reinterpret_cast<const double *>(dev(aX)),
                        reinterpret_cast<const double *>(dev(aO)), reinterpret_cast<const float *>(dev(aI)), dev(aB), h->d_chi2,
                        reinterpret_cast<PoseOutDev *>(dev(aR))));
    h->last_launches = 1;
    ORBX_CUDA(cudaMemcpyAsync(aR, dev(aR), out_bytes, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    off = 0;
    for (int f = 0; f < n_frames; f++) {
        for (int i = 0; i < 7; i++) res[f].pose[i] = aR[f].pose[i];
        res[f].n_inliers = aR[f].n_inliers; res[f].n_bad = aR[f].n_bad; res[f].lm_trials = aR[f].trials;
        if (probs[f].n) memcpy(res[f].outlier, aB + off, (size_t)probs[f].n);
        off += (size_t)probs[f].n;
    }
    return ORBX_OK;
}

// ---- device-resident: PoseOptimization right after SearchByProjection(Cur, Last), nothing leaves HBM -------------------------
// One CTA per frame lists the keypoints that received a map point (ascending keypoint index = the order in which the
// reference adds its edges, Optimizer.cc:275-350): Xw = the matched last-frame point widened to double, obs = (kpUn.pt.x,
// kpUn.pt.y, mvuRight[k] or -1), inv_sigma2 of the keypoint's octave; the initial estimate is Converter::toSE3Quat of the job's
// float pose (Eigen::Quaterniond(R) + normalizeRotation).
#define PG_THREADS 256
__global__ void __launch_bounds__(PG_THREADS)
k_pose_gather(const orbx_frame_match_job *__restrict__ jobs, const float *__restrict__ inv_sigma2, int nlevels, int pitch,
              double fx, double fy, double cx, double cy, double bf, double *__restrict__ Xw_all, double *__restrict__ obs_all,
              float *__restrict__ info_all, int32_t *__restrict__ index_all, PoseProbDev *__restrict__ probs, int kp_pitch) {
    __shared__ int warp_cnt[PG_THREADS / 32];
    __shared__ int base;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const orbx_frame_match_job &J = jobs[f];
    const orbx_frame_view &F = J.cur;
    // the caller's outlier row has kp_pitch entries (<= pitch, checked by the host): a count beyond it is a caller error, never an
    // out-of-bounds write
    const int n = min(F.n_dev ? *F.n_dev : F.n, kp_pitch);
    double *Xw = Xw_all + 3 * (size_t)f * pitch, *obs = obs_all + 3 * (size_t)f * pitch;
    float *info = info_all + (size_t)f * pitch;
    int32_t *index = index_all + (size_t)f * pitch;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int k0 = 0; k0 < n; k0 += PG_THREADS) {
        const int k = k0 + tid;
        const int m = k < n ? J.match[k] : -1;
        const bool has = m >= 0 && m < J.n_last;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; w++) off += warp_cnt[w];
        off += __popc(bal & ((1u << lane) - 1u));
        if (has && off < pitch) {
            const orbx_last_point p = J.pts[m];
            const orbx_keypoint kp = F.keys_un[k];
            Xw[3 * off] = (double)p.x; Xw[3 * off + 1] = (double)p.y; Xw[3 * off + 2] = (double)p.z;
            obs[3 * off] = (double)kp.x; obs[3 * off + 1] = (double)kp.y;
            obs[3 * off + 2] = F.u_right ? (double)F.u_right[k] : -1.0;
            const int oc = min(max(kp.octave, 0), nlevels - 1);
            info[off] = inv_sigma2[oc];
            index[off] = k;
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < PG_THREADS / 32; w++) t += warp_cnt[w]; base += t; }
        __syncthreads();
    }
    if (tid == 0) {
        PoseProbDev P;
        P.off = f * pitch; P.n = min(base, pitch);
        double R[9], q[4];
        for (int i = 0; i < 9; i++) R[i] = (double)J.Rcw[i];
        R_to_quat(R, q);
        quat_normalize(q);
        P.pose[0] = q[0]; P.pose[1] = q[1]; P.pose[2] = q[2]; P.pose[3] = q[3];
        P.pose[4] = (double)J.tcw[0]; P.pose[5] = (double)J.tcw[1]; P.pose[6] = (double)J.tcw[2];
        P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy; P.bf = bf;
        probs[f] = P;
    }
}

// results back in the caller's terms: pose (7 doubles), counts, and mvbOutlier per KEYPOINT of the frame
__global__ void __launch_bounds__(PG_THREADS)
k_pose_scatter(const PoseProbDev *__restrict__ probs, const PoseOutDev *__restrict__ outs, const uint8_t *__restrict__ outlier_all,
               const int32_t *__restrict__ index_all, double *__restrict__ pose_out, int32_t *__restrict__ n_inliers,
               uint8_t *__restrict__ outlier_kp, int kp_pitch) {
    const int f = blockIdx.x;
    const PoseProbDev P = probs[f];
    for (int j = threadIdx.x; j < P.n; j += PG_THREADS)
        outlier_kp[(size_t)f * kp_pitch + index_all[P.off + j]] = outlier_all[P.off + j];
    if (threadIdx.x < 7) pose_out[7 * f + threadIdx.x] = outs[f].pose[threadIdx.x];
    if (threadIdx.x == 0 && n_inliers) n_inliers[f] = outs[f].n_inliers;
}

extern "C" orbx_status orbx_pose_from_matches_device(orbx_pose *h, const orbx_frame_match_job *d_jobs, int n_frames,
                                                     const float *d_inv_sigma2, int nlevels, double fx, double fy, double cx, double cy,
                                                     double bf, double *d_pose_out, int32_t *d_n_inliers, uint8_t *d_outlier_kp,
                                                     int kp_pitch, void *stream) {
    if (!h || n_frames < 0 || (n_frames && (!d_jobs || !d_inv_sigma2 || !d_pose_out || !d_outlier_kp)) || nlevels < 1 || kp_pitch < 1)
        return ORBX_ERR_INVALID;
    h->last_launches = 0;
    if (n_frames == 0) return ORBX_OK;
    if (n_frames > h->max_frames) {
        orbx_set_error("orbx_pose: %d frames, handle was created for %d", n_frames, h->max_frames);
        return ORBX_ERR_CAPACITY;
    }
    const int pitch = h->max_obs / n_frames;           // observations a frame may list
    if (pitch < kp_pitch) {
        orbx_set_error("orbx_pose: %d observations for %d frames of up to %d keypoints", h->max_obs, n_frames, kp_pitch);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    ORBX_CUDA(cudaMemsetAsync(d_outlier_kp, 0, (size_t)n_frames * kp_pitch, s));
    k_pose_gather<<<n_frames, PG_THREADS, 0, s>>>(d_jobs, d_inv_sigma2, nlevels, pitch, fx, fy, cx, cy, bf, h->d_Xw, h->d_obs, h->d_info,
                                                  h->d_index, h->d_prob, kp_pitch);
    ORBX_CUDA(po_launch(n_frames, kp_pitch, s, h->d_prob, h->d_Xw, h->d_obs, h->d_info, h->d_outlier, h->d_chi2, h->d_out));
    k_pose_scatter<<<n_frames, PG_THREADS, 0, s>>>(h->d_prob, h->d_out, h->d_outlier, h->d_index, d_pose_out, d_n_inliers, d_outlier_kp,
                                                   kp_pitch);
    ORBX_CUDA(cudaGetLastError());
    h->last_launches = 3;
    return ORBX_OK;
}

extern "C" int orbx_pose_last_launches(const orbx_pose *h) { return h ? h->last_launches : 0; }
