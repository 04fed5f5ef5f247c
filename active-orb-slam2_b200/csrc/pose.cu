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
    uint8_t *h_arena; size_t arena_bytes;
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

// (H + lambda I) x = b for the 6x6 system, upper-triangular H (21 entries); returns 0 if not positive definite
__device__ __forceinline__ int po_solve6(const double *Hu, const double *b, double lambda, double *x) {
    // every loop has compile-time bounds and is unrolled, so that A lives in registers (no local-memory round trips on the one
    // thread everybody else is waiting for)
    double A[36];
    {
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int c = a; c < 6; c++) { A[6 * a + c] = Hu[k]; A[6 * c + a] = Hu[k]; k++; }
    }
#pragma unroll
    for (int a = 0; a < 6; a++) A[7 * a] += lambda;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = A[j * 6 + j];
#pragma unroll
        for (int q = 0; q < j; q++) d -= A[j * 6 + q] * A[j * 6 + q];
        ok = ok && (d > 0);
        d = sqrt(d);
        A[j * 6 + j] = d;
        const double inv = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double s = A[i * 6 + j];
#pragma unroll
            for (int q = 0; q < j; q++) s -= A[i * 6 + q] * A[j * 6 + q];
            A[i * 6 + j] = s * inv;
        }
    }
    if (!ok) return 0;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = b[i];
#pragma unroll
        for (int q = 0; q < i; q++) s -= A[i * 6 + q] * y[q];
        y[i] = s / A[i * 6 + i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double s = y[i];
#pragma unroll
        for (int q = i + 1; q < 6; q++) s -= A[q * 6 + i] * y[q];
        y[i] = s / A[i * 6 + i];
    }
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = y[i];
    return 1;
}

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
                            const int ok = po_solve6(S.H, S.b, S.lambda, S.x);
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

// ---- host side ----------------------------------------------------------------------------------------------------------
extern "C" void orbx_pose_destroy(orbx_pose *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_Xw); cudaFree(h->d_obs); cudaFree(h->d_chi2); cudaFree(h->d_info); cudaFree(h->d_outlier);
    cudaFree(h->d_prob); cudaFree(h->d_out); cudaFree(h->d_index);
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
    ORBX_CUDA(cudaMemcpyAsync(h->d_Xw, aX, 24 * nt, cudaMemcpyHostToDevice, s));
    ORBX_CUDA(cudaMemcpyAsync(h->d_obs, aO, 24 * nt, cudaMemcpyHostToDevice, s));
    ORBX_CUDA(cudaMemcpyAsync(h->d_info, aI, 4 * nt, cudaMemcpyHostToDevice, s));
    ORBX_CUDA(cudaMemcpyAsync(h->d_prob, aP, sizeof(PoseProbDev) * n_frames, cudaMemcpyHostToDevice, s));
    k_pose_optimize<<<n_frames, PO_THREADS, 0, s>>>(h->d_prob, h->d_Xw, h->d_obs, h->d_info, h->d_outlier, h->d_chi2, h->d_out, 10);
    ORBX_CUDA(cudaGetLastError());
    h->last_launches = 1;
    // results come back through the same arena (after the problems)
    PoseOutDev *aR = reinterpret_cast<PoseOutDev *>(aP + n_frames);
    uint8_t *aB = reinterpret_cast<uint8_t *>(aR + n_frames);
    ORBX_CUDA(cudaMemcpyAsync(aR, h->d_out, sizeof(PoseOutDev) * n_frames, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(aB, h->d_outlier, nt, cudaMemcpyDeviceToHost, s));
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
    k_pose_optimize<<<n_frames, PO_THREADS, 0, s>>>(h->d_prob, h->d_Xw, h->d_obs, h->d_info, h->d_outlier, h->d_chi2, h->d_out, 10);
    k_pose_scatter<<<n_frames, PG_THREADS, 0, s>>>(h->d_prob, h->d_out, h->d_outlier, h->d_index, d_pose_out, d_n_inliers, d_outlier_kp,
                                                   kp_pitch);
    ORBX_CUDA(cudaGetLastError());
    h->last_launches = 3;
    return ORBX_OK;
}

extern "C" int orbx_pose_last_launches(const orbx_pose *h) { return h ? h->last_launches : 0; }
