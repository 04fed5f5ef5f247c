// Local bundle adjustment: replaces the g2o work behind Optimizer::LocalBundleAdjustment (reference
// src/Optimizer.cc:454-779): EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ computeError + linearizeOplus
// (Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:103-157, :188-234), BaseBinaryEdge::constructQuadraticForm
// with the Huber kernel (core/base_binary_edge.hpp:55-120, core/robust_kernel_impl.cpp:78-91),
// BlockSolver::buildSystem / setLambda / solve incl. the Schur complement (core/block_solver.hpp:354-486),
// the reduced solve (solvers/linear_solver_eigen.h:94-124, here a dense Cholesky), VertexSE3Expmap::oplusImpl /
// SE3Quat::exp (types/se3quat.h:223-257) and the Levenberg schedule
// (core/optimization_algorithm_levenberg.cpp:61-189) with the two-round outlier policy of Optimizer.cc:659-735.
//
// Everything is double precision like g2o (tolerance vs the oracle: 1e-4 relative on poses / points, identical
// outlier sets); none of it is shaped as a dense GEMM: the Schur product is block-sparse (about 30x fewer flops
// than the dense 6P x 3L x 6P contraction) and needs f64, so the tensor pipes are not used.
// This file is the host side (window upload, Levenberg control flow of the host-driven path, C ABI) plus
//   k2_* (lba_chunk.cu)  residuals / Jacobians / quadratic form / Schur complement as atomic-free chunked sums
//   k_lba_fused (lba_fused.cu)  a whole optimize() call in one thread-block cluster (windows up to 36 free keyframes)
//   k_lba_solve    one CTA: Cholesky + two triangular solves of the reduced system in shared memory
//   k_lba_update   landmark back-substitution, point and pose updates, Levenberg's scale term
// The Levenberg control flow runs on the host and reads four doubles back per trial.
#include <math.h>
#include <algorithm>
#include <cmath>
#include <vector>
#include "orbx_internal.cuh"

#include <chrono>
#include "lba_common.cuh"

// lba_chunk.cu
orbx_status orbx_lba_chunk_linearize(const LbaDev &D, int robust, int build, int want_hpp, cudaStream_t s, int *launches);
orbx_status orbx_lba_chunk_schur(const LbaDev &D, double lambda, cudaStream_t s, int *launches);
orbx_status orbx_lba_chunk_init();
// lba_fused.cu
size_t orbx_lba_fused_smem(int np);
bool orbx_lba_fused_fits(int n_kf, int np);
orbx_status orbx_lba_fused_init();
orbx_status orbx_lba_grid_init();
bool orbx_lba_grid_available(int n_kf, int np);
size_t orbx_lba_grid_scratch(int n_pts, int n_kf);
orbx_status orbx_lba_grid_launch(const LbaDev &D, double *kf_bak, double *pt_bak, int iterations, int robust, int capture, double *cap_Hs,
                                 double *cap_bs, double *cap_xp, double *out, double *slots, int max_pts, int max_kf, cudaStream_t s);
orbx_status orbx_lba_fused_launch(const LbaDev &D, double *kf_bak, double *pt_bak, int iterations, int robust, int capture,
                                  double *cap_Hs, double *cap_bs, double *cap_xp, double *out, cudaStream_t s, int wide);

// computeLambdaInit: max |diagonal| over every active vertex (optimization_algorithm_levenberg.cpp:166-180)
__global__ void __launch_bounds__(LBA_THREADS) k_lba_maxdiag(LbaDev D) {
    __shared__ double tmp[32];
    double m = 0;
    const int diag6[6] = {0, 6, 11, 15, 18, 20}, diag3[3] = {0, 3, 5};
    for (int i = threadIdx.x; i < D.np * 6; i += LBA_THREADS) m = fmax(m, fabs(D.Hpp[27 * (i / 6) + diag6[i % 6]]));
    for (int i = threadIdx.x; i < D.n_pts * 3; i += LBA_THREADS) m = fmax(m, fabs(D.Hll[9 * (i / 3) + diag3[i % 3]]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) tmp[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < LBA_THREADS / 32; w++) m = fmax(m, tmp[w]);
        D.scal[2] = m;
    }
}

// dense Cholesky solve of the reduced camera system; A = upper block triangle of H_schur (row-major n x n).
// W is the n x n workspace (shared memory when it fits, else the lower triangle of A itself).
__global__ void __launch_bounds__(LBA_SOLVE_THREADS) k_lba_solve(LbaDev D, int use_smem) {
    extern __shared__ __align__(16) double sw[];
    __shared__ int ok;
    const int n = D.n, tid = threadIdx.x;
    double *W = use_smem ? sw : D.Hs;
    const int ld = n;
    // symmetric fill: element (r,c) with r >= c comes from the stored upper entry (c,r)
    for (int i = tid; i < n * n; i += LBA_SOLVE_THREADS) {
        const int r = i / n, c = i - r * n;
        if (r >= c) {
            const int pr = r / 6, pc = c / 6;
            const double v = pr == pc ? D.Hs[(size_t)r * n + c] : D.Hs[(size_t)c * n + r];   // diagonal blocks are stored in full
            if (use_smem) W[r * ld + c] = v;
            else if (pr != pc) W[r * ld + c] = v;                 // in place: write the mirror of an off-diagonal block
        }
    }
    if (tid == 0) ok = 1;
    __syncthreads();
    for (int j = 0; j < n; j++) {
        if (tid == 0) {
            const double d = W[j * ld + j];
            if (!(d > 0)) ok = 0; else W[j * ld + j] = sqrt(d);
        }
        __syncthreads();
        if (!ok) break;
        const double dj = W[j * ld + j];
        for (int i = j + 1 + tid; i < n; i += LBA_SOLVE_THREADS) W[i * ld + j] /= dj;
        __syncthreads();
        const int m = n - j - 1;                         // trailing update of the lower triangle
        for (int t = tid; t < m * m; t += LBA_SOLVE_THREADS) {
            const int r = j + 1 + t / m, c = j + 1 + t % m;
            if (c <= r) W[r * ld + c] -= W[r * ld + j] * W[c * ld + j];
        }
        __syncthreads();
    }
    if (ok) {
        // L y = b, then L^T x = y; one warp, lanes split the dot products
        if (tid < 32) {
            for (int i = 0; i < n; i++) {
                double s = 0;
                for (int k = tid; k < i; k += 32) s += W[i * ld + k] * D.xp[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (tid == 0) D.xp[i] = (D.bs[i] - s) / W[i * ld + i];
                __syncwarp();
            }
            for (int i = n - 1; i >= 0; i--) {
                double s = 0;
                for (int k = i + 1 + tid; k < n; k += 32) s += W[k * ld + i] * D.xp[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (tid == 0) D.xp[i] = (D.xp[i] - s) / W[i * ld + i];
                __syncwarp();
            }
        }
    } else {
        for (int i = tid; i < n; i += LBA_SOLVE_THREADS) D.xp[i] = 0;   // failed factorisation: no step (the trial is rejected)
    }
    if (tid == 0) D.scal[3] = ok ? 1.0 : 0.0;
}

// x_l = D^-1 (b_l - B^T x_p) (block_solver.hpp:461-481), SparseOptimizer::update, computeScale
__global__ void __launch_bounds__(LBA_THREADS) k_lba_update(LbaDev D, double lambda) {
    __shared__ double tmp[32];
    const int i = blockIdx.x * LBA_THREADS + threadIdx.x;
    double sc = 0;
    if (i < D.n_pts) {
        const double *hl = D.Hll + 9 * i;
        double c0 = hl[6], c1 = hl[7], c2 = hl[8];
        for (int e = D.ptstart[i]; e < D.ptstart[i + 1]; e++) {
            const int p = D.kfidx[D.ekf[e]];
            if (p < 0 || D.level1[e]) continue;
            const double *B = D.Hpl + 18 * (size_t)e, *x = D.xp + 6 * p;
            for (int a = 0; a < 6; a++) { c0 -= B[3 * a] * x[a]; c1 -= B[3 * a + 1] * x[a]; c2 -= B[3 * a + 2] * x[a]; }
        }
        double Di[6];
        dinv3(hl, lambda, Di);
        const double x0 = Di[0] * c0 + Di[1] * c1 + Di[2] * c2, x1 = Di[1] * c0 + Di[3] * c1 + Di[4] * c2, x2 = Di[2] * c0 + Di[4] * c1 + Di[5] * c2;
        D.xl[3 * i] = x0; D.xl[3 * i + 1] = x1; D.xl[3 * i + 2] = x2;
        sc = x0 * (lambda * x0 + hl[6]) + x1 * (lambda * x1 + hl[7]) + x2 * (lambda * x2 + hl[8]);
        D.pt[3 * i] += x0; D.pt[3 * i + 1] += x1; D.pt[3 * i + 2] += x2;
    } else if (i - D.n_pts < D.n_kf) {
        const int k = i - D.n_pts, p = D.kfidx[k];
        if (p >= 0) {
            const double *x = D.xp + 6 * p, *b = D.Hpp + 27 * p + 21;
            for (int a = 0; a < 6; a++) sc += x[a] * (lambda * x[a] + b[a]);
            se3_oplus(D.kf + 7 * k, x);
        }
    }
    const double s = block_sum(sc, tmp);
    if (threadIdx.x == 0 && s != 0) atomicAdd(&D.scal[1], s);
}

// Optimizer.cc:671-703 / :709-735: stored chi2 over the threshold or depth not positive
__global__ void __launch_bounds__(LBA_THREADS) k_lba_classify(LbaDev D, uint8_t *flag) {
    const int e = blockIdx.x * LBA_THREADS + threadIdx.x;
    if (e >= D.n_edges) return;
    double R[9];
    const double *T = D.kf + 7 * D.ekf[e], *X = D.pt + 3 * D.ept[e];
    quat_to_R(T, R);
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + T[6];
    flag[e] = (D.chi2[e] > (D.stereo[e] ? 7.815 : 5.991) || !(z > 0)) ? 1 : 0;
}

// ---- host ----------------------------------------------------------------------------------------------------
struct orbx_lba {
    int device, max_kf, max_pts, max_edges;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    LbaDev D;
    double *d_kf_bak, *d_pt_bak;
    uint8_t *d_flag;
    uint8_t *arena_h, *arena_d; size_t arena_cap, arena_used;   // the window as uploaded (see lba_load)
    std::vector<int> v_start, v_kfidx, v_kcount, v_bcount, v_cur;
    std::vector<int4> v_plist;
    std::vector<int> v_kp;
    double *h_scal;
    std::vector<int> perm;      // sorted position -> caller's edge index
    const volatile uint8_t *stop;
    int launches;
    int loaded;
    int use_fused;      // 0 = always the multi-kernel path (ORBX_LBA_MULTIKERNEL=1, for tests and large windows)
    // work lists of the cluster kernel
    int pending;        // orbx_lba_solve_begin issued, orbx_lba_solve_end not yet
    double *h_kf, *h_pt, *h_chi; uint8_t *h_flag;   // pinned result staging of the asynchronous form
    double *d_hppart, *d_part, *d_dinv;
    size_t cap_hppart, cap_part;
    double *d_slots;    // block-level partial sums of the whole-GPU kernel (lba_chunk.cu, k_lba_grid)
    int use_grid;       // a single window (orbx_lba_solve_host) takes the whole-GPU kernel; ORBX_LBA_GRID=0 keeps it on the cluster kernel
};

extern "C" void orbx_lba_destroy(orbx_lba *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    cudaFree(h->D.level1); cudaFree(h->D.err); cudaFree(h->D.chi2); cudaFree(h->D.Hpl);
    cudaFree(h->D.Hpp); cudaFree(h->D.Hll); cudaFree(h->D.Hs); cudaFree(h->D.bs); cudaFree(h->D.xp); cudaFree(h->D.xl);
    cudaFree(h->D.scal); cudaFree(h->d_kf_bak); cudaFree(h->d_pt_bak); cudaFree(h->d_flag); cudaFree(h->arena_d);
    if (h->arena_h) cudaFreeHost(h->arena_h);
    cudaFree(h->d_hppart); cudaFree(h->d_part); cudaFree(h->d_dinv); cudaFree(h->d_slots);
    if (h->h_scal) cudaFreeHost(h->h_scal);
    if (h->h_kf) cudaFreeHost(h->h_kf);
    if (h->h_pt) cudaFreeHost(h->h_pt);
    if (h->h_chi) cudaFreeHost(h->h_chi);
    if (h->h_flag) cudaFreeHost(h->h_flag);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" orbx_status orbx_lba_create(orbx_lba **out, int max_keyframes, int max_points, int max_edges, int device) {
    if (!out) return ORBX_ERR_INVALID;
    *out = nullptr;
    if (max_keyframes < 1 || max_points < 1 || max_edges < 1) {
        orbx_set_error("orbx_lba_create: bad argument");
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
    orbx_lba *h = new orbx_lba();
    memset(&h->D, 0, sizeof(h->D));
    h->device = device; h->max_kf = max_keyframes; h->max_pts = max_points; h->max_edges = max_edges;
    h->d_kf_bak = h->d_pt_bak = nullptr; h->d_flag = nullptr; h->h_scal = nullptr;
    h->arena_h = h->arena_d = nullptr; h->arena_cap = h->arena_used = 0;
    h->stream = nullptr; h->ev0 = h->ev1 = nullptr; h->stop = nullptr; h->launches = 0; h->loaded = 0;
    { const char *mk = getenv("ORBX_LBA_MULTIKERNEL"); h->use_fused = !(mk && mk[0] == '1'); }
    h->d_hppart = h->d_part = h->d_dinv = h->d_slots = nullptr;
    { const char *g = getenv("ORBX_LBA_GRID"); h->use_grid = !(g && g[0] == '0'); }
    h->pending = 0; h->h_kf = h->h_pt = h->h_chi = nullptr; h->h_flag = nullptr;
    h->cap_hppart = h->cap_part = 0;
    const size_t K = max_keyframes, L = max_points, E = max_edges, N = 6 * K;
    cudaError_t ce = cudaSuccess;
#define TRY(x) if (ce == cudaSuccess) ce = (x)
    TRY(cudaMalloc((void **)&h->d_kf_bak, sizeof(double) * 7 * K));
    TRY(cudaMalloc((void **)&h->d_dinv, sizeof(double) * 10 * (L + 1)));
    TRY(cudaMalloc((void **)&h->d_pt_bak, sizeof(double) * 3 * L));
    TRY(cudaMalloc((void **)&h->D.level1, E));
    TRY(cudaMalloc((void **)&h->d_flag, E));
    TRY(cudaMalloc((void **)&h->D.err, sizeof(double) * 3 * E));
    TRY(cudaMalloc((void **)&h->D.chi2, sizeof(double) * E));
    TRY(cudaMalloc((void **)&h->D.Hpl, sizeof(double) * 18 * E));
    TRY(cudaMalloc((void **)&h->D.Hpp, sizeof(double) * 27 * K));
    TRY(cudaMalloc((void **)&h->D.Hll, sizeof(double) * 9 * L));
    TRY(cudaMalloc((void **)&h->D.Hs, sizeof(double) * N * N));
    TRY(cudaMalloc((void **)&h->D.bs, sizeof(double) * N));
    TRY(cudaMalloc((void **)&h->D.xp, sizeof(double) * N));
    TRY(cudaMalloc((void **)&h->D.xl, sizeof(double) * 3 * L));
    TRY(cudaMalloc((void **)&h->D.scal, sizeof(double) * 16));
    TRY(cudaMalloc((void **)&h->d_slots, sizeof(double) * orbx_lba_grid_scratch(max_points, max_keyframes)));
    TRY(cudaMemset(h->d_slots, 0, sizeof(double) * orbx_lba_grid_scratch(max_points, max_keyframes)));     // the arrival counters start at zero
    TRY(cudaMallocHost((void **)&h->h_scal, sizeof(double) * 16));
    TRY(cudaMallocHost((void **)&h->h_kf, sizeof(double) * 7 * K));
    TRY(cudaMallocHost((void **)&h->h_pt, sizeof(double) * 3 * L));
    TRY(cudaMallocHost((void **)&h->h_chi, sizeof(double) * E));
    TRY(cudaMallocHost((void **)&h->h_flag, E));
    TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    TRY(cudaEventCreate(&h->ev0));
    TRY(cudaEventCreate(&h->ev1));
    TRY(ORBX_RAISE_SMEM(k_lba_solve));
    if (ce == cudaSuccess && orbx_lba_fused_init() != ORBX_OK) ce = cudaErrorUnknown;
    if (ce == cudaSuccess && orbx_lba_chunk_init() != ORBX_OK) ce = cudaErrorUnknown;
    if (ce == cudaSuccess && orbx_lba_grid_init() != ORBX_OK) ce = cudaErrorUnknown;
#undef TRY
    if (ce != cudaSuccess) {
        orbx_set_error("orbx_lba_create: %s", cudaGetErrorString(ce));
        orbx_lba_destroy(h);
        return ORBX_ERR_CUDA;
    }
    *out = h;
    return ORBX_OK;
}

template <typename T>
static orbx_status lba_grow(T **p, size_t *cap, size_t need) {
    if (need <= *cap && *p) return ORBX_OK;
    if (*p) ORBX_CUDA(cudaFree(*p));
    *p = nullptr;
    need += need / 4 + 64;
    ORBX_CUDA(cudaMalloc((void **)p, sizeof(T) * need));
    *cap = need;
    return ORBX_OK;
}

// One pinned host arena mirrored by one device arena: a window (estimates, sorted edges, work lists) is assembled in
// place on the host and goes to the device in a single asynchronous copy.
static orbx_status arena_reserve(orbx_lba *h, size_t bytes) {
    if (bytes <= h->arena_cap && h->arena_h) return ORBX_OK;
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    if (h->arena_h) ORBX_CUDA(cudaFreeHost(h->arena_h));
    if (h->arena_d) ORBX_CUDA(cudaFree(h->arena_d));
    h->arena_h = h->arena_d = nullptr;
    bytes += bytes / 4 + 4096;
    ORBX_CUDA(cudaMallocHost((void **)&h->arena_h, bytes));
    ORBX_CUDA(cudaMalloc((void **)&h->arena_d, bytes));
    h->arena_cap = bytes;
    return ORBX_OK;
}
template <typename T>
static T *arena_take(orbx_lba *h, size_t n, T **dev) {
    h->arena_used = (h->arena_used + 15) & ~(size_t)15;
    T *p = reinterpret_cast<T *>(h->arena_h + h->arena_used);
    *dev = reinterpret_cast<T *>(h->arena_d + h->arena_used);
    h->arena_used += sizeof(T) * n;
    return p;
}

// list chunk: the cluster kernel has 64 warps for the whole window and one CTA finishing the blocks, so it wants few, long
// chunks; the multi-kernel path spreads chunks over the whole GPU and wants them short
#define LBA_CHUNK_FUSED 128
#define LBA_CHUNK_WIDE 64
// upload a problem: estimates, edges sorted by landmark, reduced indices, and (for the cluster kernel) the work lists:
// edges by keyframe and (edge, edge) pairs of a landmark by pose-pair block, both cut into chunks
static orbx_status lba_load(orbx_lba *h, const orbx_lba_problem *P, bool wide = false) {
    if (!P || P->n_kf < 0 || P->n_pts < 0 || P->n_edges < 0) return ORBX_ERR_INVALID;
    if (P->n_kf > h->max_kf || P->n_pts > h->max_pts || P->n_edges > h->max_edges) {
        orbx_set_error("problem (%d keyframes, %d points, %d edges) exceeds the handle (%d, %d, %d)", P->n_kf, P->n_pts, P->n_edges,
                       h->max_kf, h->max_pts, h->max_edges);
        return ORBX_ERR_CAPACITY;
    }
    if ((P->n_kf && (!P->kf_pose || !P->kf_fixed)) || (P->n_pts && !P->pts) ||
        (P->n_edges && (!P->e_kf || !P->e_pt || !P->e_obs || !P->e_inv_sigma2 || !P->e_stereo)))
        return ORBX_ERR_INVALID;
    const int E = P->n_edges, L = P->n_pts, K = P->n_kf;
    static const bool ltrace = getenv("ORBX_LBA_TRACE") != nullptr;
    const auto lt0 = std::chrono::steady_clock::now();
    auto lsince = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - lt0).count(); };
    double lt[6] = {0, 0, 0, 0, 0, 0};
    for (int e = 0; e < E; e++)
        if (P->e_kf[e] < 0 || P->e_kf[e] >= K || P->e_pt[e] < 0 || P->e_pt[e] >= L) {
            orbx_set_error("edge %d refers to vertex (%d, %d) outside the problem", e, P->e_kf[e], P->e_pt[e]);
            return ORBX_ERR_INVALID;
        }
    // ---- counts first, so that the arena can be sized -----------------------------------------------------------------
    std::vector<int> &start = h->v_start, &kfidx = h->v_kfidx, &kcount = h->v_kcount, &bcount = h->v_bcount, &cur = h->v_cur;
    start.assign(L + 1, 0);
    for (int e = 0; e < E; e++) start[P->e_pt[e] + 1]++;
    for (int l = 0; l < L; l++) start[l + 1] += start[l];
    h->perm.assign(E, 0);
    cur.assign(start.begin(), start.end() - 1);
    for (int e = 0; e < E; e++) h->perm[cur[P->e_pt[e]]++] = e;       // stable: caller's order inside a landmark
    lt[0] = lsince();                                   // validation + sort by landmark
    kfidx.assign(K, -1);
    int np = 0;
    for (int k = 0; k < K; k++) kfidx[k] = P->kf_fixed[k] ? -1 : np++;
    const int nblk = np * (np + 1) / 2;
    auto ub = [np](int p1, int p2) { return p1 * np - p1 * (p1 - 1) / 2 + (p2 - p1); };
    size_t npairs = 0, n_kchunks = 0, n_pchunks = 0;
    const int LBA_CHUNK = (!wide && h->use_fused && orbx_lba_fused_fits(K, np)) ? LBA_CHUNK_FUSED : LBA_CHUNK_WIDE;
    if (true) {
        kcount.assign(np + 1, 0);
        bcount.assign(nblk + 1, 0);
        for (int e = 0; e < E; e++) { const int p = kfidx[P->e_kf[e]]; if (p >= 0) kcount[p + 1]++; }
        // every unordered pair of a landmark's edges once (a landmark has one edge per keyframe, so two different edges lie in different
        // keyframes and the pair belongs to exactly one block of the upper triangle), and every edge with itself; the pairs are listed
        // here with their block and scattered into block order below (a counting sort: within a block they stay in landmark order)
        std::vector<int4> &plist = h->v_plist;
        std::vector<int> &kp = h->v_kp;                  // reduced keyframe index of every edge in landmark order
        plist.clear();
        plist.reserve((size_t)E * 4);
        kp.resize(E);
        for (int s = 0; s < E; s++) kp[s] = kfidx[P->e_kf[h->perm[s]]];
        for (int l = 0; l < L; l++)
            for (int i = start[l]; i < start[l + 1]; i++) {
                const int p1 = kp[i];
                if (p1 < 0) continue;
                int b = ub(p1, p1);
                bcount[b + 1]++;
                plist.push_back(make_int4(i, i, l, b));
                for (int j = i + 1; j < start[l + 1]; j++) {
                    const int p2 = kp[j];
                    if (p2 < 0 || p2 == p1) continue;
                    if (p1 < p2) { b = ub(p1, p2); plist.push_back(make_int4(i, j, l, b)); }
                    else { b = ub(p2, p1); plist.push_back(make_int4(j, i, l, b)); }
                    bcount[b + 1]++;
                }
            }
        // every keyframe and every block gets at least one chunk (an empty one if it has no entries): the warp that owns a chunk also
        // reports it done, and whoever completes a keyframe / block writes its sum (lba_chunk.cu)
        for (int p = 0; p < np; p++) { n_kchunks += std::max(1, (kcount[p + 1] + LBA_CHUNK - 1) / LBA_CHUNK); kcount[p + 1] += kcount[p]; }
        for (int b = 0; b < nblk; b++) { n_pchunks += std::max(1, (bcount[b + 1] + LBA_CHUNK - 1) / LBA_CHUNK); bcount[b + 1] += bcount[b]; }
        npairs = (size_t)bcount[nblk];
    }
    lt[1] = lsince();                                   // + pair enumeration and counts
    const size_t bytes = 8 * (size_t)(7 * K + 3 * L + 3 * E + E) + 4 * (size_t)(K + L + 1 + 2 * E) + E +
                         16 * ((size_t)E + npairs + n_kchunks + n_pchunks) + 4 * (size_t)(np + 1 + nblk + 1) + 16 * 20;
    orbx_status st = arena_reserve(h, bytes);
    if (st) return st;
    h->arena_used = 0;
    LbaDev &D = h->D;
    double *kf = arena_take(h, 7 * (size_t)K, &D.kf), *pt = arena_take(h, 3 * (size_t)L, &D.pt);
    int *d_i; double *d_d; uint8_t *d_u; int4 *d_4;
    int *a_kfidx = arena_take(h, K, &d_i); D.kfidx = d_i;
    int *a_start = arena_take(h, (size_t)L + 1, &d_i); D.ptstart = d_i;
    int *ekf = arena_take(h, E, &d_i); D.ekf = d_i;
    int *ept = arena_take(h, E, &d_i); D.ept = d_i;
    double *obs = arena_take(h, 3 * (size_t)E, &d_d); D.obs = d_d;
    double *info = arena_take(h, E, &d_d); D.info = d_d;
    uint8_t *st8 = arena_take(h, E, &d_u); D.stereo = d_u;
    memcpy(kf, P->kf_pose, sizeof(double) * 7 * K);
    memcpy(pt, P->pts, sizeof(double) * 3 * L);
    memcpy(a_kfidx, kfidx.data(), sizeof(int) * K);
    memcpy(a_start, start.data(), sizeof(int) * (L + 1));
    for (int s = 0; s < E; s++) {
        const int e = h->perm[s];
        ekf[s] = P->e_kf[e]; ept[s] = P->e_pt[e]; st8[s] = P->e_stereo[e] ? 1 : 0;
        obs[3 * s] = P->e_obs[3 * e]; obs[3 * s + 1] = P->e_obs[3 * e + 1]; obs[3 * s + 2] = P->e_obs[3 * e + 2];
        info[s] = (double)P->e_inv_sigma2[e];
    }
    lt[2] = lsince();                                   // + arena, estimates, edge arrays
    D.n_kchunks = D.n_pchunks = 0;
    if (true) {
        int4 *kfe = arena_take(h, E > 0 ? E : 1, &d_4); D.kfe = d_4;
        int4 *kchunk = arena_take(h, n_kchunks + 1, &d_4); D.kchunk = d_4;
        int *kfc = arena_take(h, (size_t)np + 1, &d_i); D.kf_cstart = d_i;
        int4 *pairs = arena_take(h, npairs + 1, &d_4); D.pairs = d_4;
        int4 *pchunk = arena_take(h, n_pchunks + 1, &d_4); D.pchunk = d_4;
        int *blkc = arena_take(h, (size_t)nblk + 1, &d_i); D.blk_cstart = d_i;
        cur.assign(kcount.begin(), kcount.end() - 1);
        for (int s = 0; s < E; s++) { const int p = kfidx[ekf[s]]; if (p >= 0) kfe[cur[p]++] = make_int4(s, ekf[s], ept[s], st8[s]); }
        int nc = 0;
        for (int p = 0; p < np; p++) {
            kfc[p] = nc;
            if (kcount[p] == kcount[p + 1]) kchunk[nc++] = make_int4(p, kcount[p], 0, 0);
            for (int b = kcount[p]; b < kcount[p + 1]; b += LBA_CHUNK) kchunk[nc++] = make_int4(p, b, std::min(LBA_CHUNK, kcount[p + 1] - b), 0);
        }
        kfc[np] = nc;
        D.n_kchunks = nc;
        cur.assign(bcount.begin(), bcount.end() - 1);
        for (const int4 &q : h->v_plist) pairs[cur[q.w]++] = make_int4(q.x, q.y, q.z, 0);
        nc = 0;
        int blk = 0;
        for (int p1 = 0; p1 < np; p1++)
            for (int p2 = p1; p2 < np; p2++, blk++) {
                blkc[blk] = nc;
                if (bcount[blk] == bcount[blk + 1]) pchunk[nc++] = make_int4(blk, bcount[blk], 0, p1 == p2);
                for (int b = bcount[blk]; b < bcount[blk + 1]; b += LBA_CHUNK)
                    pchunk[nc++] = make_int4(blk, b, std::min(LBA_CHUNK, bcount[blk + 1] - b), p1 == p2);
            }
        blkc[nblk] = nc;
        D.n_pchunks = nc;
        if ((st = lba_grow(&h->d_hppart, &h->cap_hppart, 27 * ((size_t)D.n_kchunks + 1)))) return st;
        if ((st = lba_grow(&h->d_part, &h->cap_part, 42 * ((size_t)D.n_pchunks + 1)))) return st;
        D.hppart = h->d_hppart; D.part = h->d_part; D.dinv = h->d_dinv;
    }
    lt[3] = lsince();                                   // + work lists
    if (h->arena_used > h->arena_cap) {
        orbx_set_error("internal: arena sized %zu, used %zu", h->arena_cap, h->arena_used);
        return ORBX_ERR_NOMEM;
    }
    cudaStream_t s = h->stream;
    ORBX_CUDA(cudaMemcpyAsync(h->arena_d, h->arena_h, h->arena_used, cudaMemcpyHostToDevice, s));
    if (E) {
        ORBX_CUDA(cudaMemsetAsync(D.level1, 0, E, s));
        ORBX_CUDA(cudaMemsetAsync(D.chi2, 0, sizeof(double) * E, s));
        ORBX_CUDA(cudaMemsetAsync(D.err, 0, sizeof(double) * 3 * E, s));
    }
    D.n_kf = K; D.n_pts = L; D.n_edges = E; D.np = np; D.n = 6 * np;
    D.fx = P->fx; D.fy = P->fy; D.cx = P->cx; D.cy = P->cy; D.bf = P->bf; D.bf_f = (float)P->bf;
    D.d_mono = (double)(float)sqrt(5.991); D.d_stereo = (double)(float)sqrt(7.815);
    h->stop = P->stop_flag;
    h->loaded = 1;
    if (ltrace) fprintf(stderr, "lba_load: sort %.0f, pairs %.0f, edge arrays %.0f, lists %.0f, enqueue %.0f us (cumulative), %zu bytes\n", lt[0], lt[1], lt[2], lt[3], lsince(), h->arena_used);
    return ORBX_OK;
}

static inline int blocks_for(int n) { return n > 0 ? (n + LBA_THREADS - 1) / LBA_THREADS : 1; }

static orbx_status lba_errors(orbx_lba *h, int robust) {      // computeActiveErrors + activeRobustChi2 -> scal[0]
    return orbx_lba_chunk_linearize(h->D, robust, 0, 0, h->stream, &h->launches);
}

static orbx_status lba_build(orbx_lba *h, int robust, int want_hpp) {       // computeActiveErrors + BlockSolver::buildSystem
    return orbx_lba_chunk_linearize(h->D, robust, 1, want_hpp, h->stream, &h->launches);
}

static orbx_status lba_schur(orbx_lba *h, double lambda) {    // BlockSolver::solve up to the linear solve
    return orbx_lba_chunk_schur(h->D, lambda, h->stream, &h->launches);
}

static orbx_status lba_solve_update(orbx_lba *h, double lambda) {
    const LbaDev &D = h->D;
    const size_t sm = sizeof(double) * (size_t)D.n * D.n;
    const int use_smem = sm <= 200 * 1024;
    k_lba_solve<<<1, LBA_SOLVE_THREADS, use_smem ? sm : 0, h->stream>>>(D, use_smem);
    ORBX_CUDA(cudaMemsetAsync(D.scal + 1, 0, sizeof(double), h->stream));
    k_lba_update<<<blocks_for(D.n_pts + D.n_kf), LBA_THREADS, 0, h->stream>>>(D, lambda);
    h->launches += 2;
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

static orbx_status read_scal(orbx_lba *h) {
    ORBX_CUDA(cudaMemcpyAsync(h->h_scal, h->D.scal, sizeof(double) * 4, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

static inline bool stop_requested(const orbx_lba *h) { return h->stop && *h->stop; }

// optimizer.initializeOptimization(0) + optimizer.optimize(iterations)
static orbx_status lba_optimize(orbx_lba *h, int iterations, int robust, orbx_lba_result *res, bool *first) {
    LbaDev &D = h->D;
    orbx_status st;
    if (h->use_fused && iterations > 0 && orbx_lba_fused_fits(D.n_kf, D.np)) {
        // whole optimize() call in one cluster kernel; the stop flag is honoured between the two rounds only (a round
        // is a few hundred microseconds here, less than one iteration of the reference)
        const int capture = *first ? 1 : 0;
        const bool want = capture && (res->first_Hschur || res->first_bschur || res->first_xp);
        ORBX_CUDA(cudaMemsetAsync(D.scal + 4, 0, sizeof(double) * 2, h->stream));   // scal[6..11]: phase timers, cleared per solve
        if (h->use_grid && orbx_lba_grid_available(D.n_kf, D.np)) {     // one window at a time: the whole GPU
            if ((st = orbx_lba_grid_launch(D, h->d_kf_bak, h->d_pt_bak, iterations, robust, capture, want ? D.Hs : nullptr, D.bs, D.xp,
                                           D.scal + 4, h->d_slots, h->max_pts, h->max_kf, h->stream)))
                return st;
        } else if ((st = orbx_lba_fused_launch(D, h->d_kf_bak, h->d_pt_bak, iterations, robust, capture, want ? D.Hs : nullptr, D.bs, D.xp,
                                               D.scal + 4, h->stream, 1)))     // the 16-CTA cluster
            return st;
        h->launches++;
        ORBX_CUDA(cudaMemcpyAsync(h->h_scal + 4, D.scal + 4, sizeof(double) * 8, cudaMemcpyDeviceToHost, h->stream));
        if (want) {
            if (res->first_Hschur) ORBX_CUDA(cudaMemcpyAsync(res->first_Hschur, D.Hs, sizeof(double) * D.n * D.n, cudaMemcpyDeviceToHost, h->stream));
            if (res->first_bschur) ORBX_CUDA(cudaMemcpyAsync(res->first_bschur, D.bs, sizeof(double) * D.n, cudaMemcpyDeviceToHost, h->stream));
            if (res->first_xp) ORBX_CUDA(cudaMemcpyAsync(res->first_xp, D.xp, sizeof(double) * D.n, cudaMemcpyDeviceToHost, h->stream));
        }
        ORBX_CUDA(cudaStreamSynchronize(h->stream));
        res->lm_trials += (int)h->h_scal[4];     // scal[4] was cleared before this launch
        if (capture) { res->first_lambda = h->h_scal[5]; *first = false; }
        return ORBX_OK;
    }
    double lambda = 0, ni = 2;
    int nBad = 0;
    for (int it = 0; it < iterations && !stop_requested(h); it++) {
        if ((st = lba_build(h, robust, it == 0))) return st;      // computeLambdaInit reads the H_pp diagonals
        if (it == 0) { k_lba_maxdiag<<<1, LBA_THREADS, 0, h->stream>>>(D); h->launches++; }
        if ((st = read_scal(h))) return st;
        double currentChi = h->h_scal[0];
        const double iniChi = currentChi;
        if (it == 0) { lambda = 1e-5 * h->h_scal[2]; ni = 2; nBad = 0; }
        double rho = 0;
        int qmax = 0;
        do {
            ORBX_CUDA(cudaMemcpyAsync(h->d_kf_bak, D.kf, sizeof(double) * 7 * D.n_kf, cudaMemcpyDeviceToDevice, h->stream));   // push
            ORBX_CUDA(cudaMemcpyAsync(h->d_pt_bak, D.pt, sizeof(double) * 3 * D.n_pts, cudaMemcpyDeviceToDevice, h->stream));
            if ((st = lba_schur(h, lambda))) return st;
            const bool capture = *first;
            if (capture) {
                *first = false;
                res->first_lambda = lambda;
                if (res->first_Hschur) {
                    std::vector<double> Hs((size_t)D.n * D.n);
                    ORBX_CUDA(cudaMemcpyAsync(Hs.data(), D.Hs, sizeof(double) * Hs.size(), cudaMemcpyDeviceToHost, h->stream));
                    ORBX_CUDA(cudaStreamSynchronize(h->stream));
                    for (int r = 0; r < D.n; r++)
                        for (int c = 0; c < D.n; c++) {
                            const bool up = r / 6 <= c / 6;
                            res->first_Hschur[(size_t)r * D.n + c] = up ? Hs[(size_t)r * D.n + c] : Hs[(size_t)c * D.n + r];
                        }
                }
                if (res->first_bschur) ORBX_CUDA(cudaMemcpyAsync(res->first_bschur, D.bs, sizeof(double) * D.n, cudaMemcpyDeviceToHost, h->stream));
            }
            if ((st = lba_solve_update(h, lambda))) return st;
            if (capture && res->first_xp) ORBX_CUDA(cudaMemcpyAsync(res->first_xp, D.xp, sizeof(double) * D.n, cudaMemcpyDeviceToHost, h->stream));
            if ((st = lba_errors(h, robust))) return st;
            if ((st = read_scal(h))) return st;
            double tempChi = h->h_scal[0];
            const bool ok2 = h->h_scal[3] != 0.0;
            if (!ok2) tempChi = 1.7976931348623157e308;
            rho = currentChi - tempChi;
            const double scale = h->h_scal[1] + 1e-3;
            rho /= scale;
            res->lm_trials++;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                lambda *= std::max(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni; ni *= 2;
                ORBX_CUDA(cudaMemcpyAsync(D.kf, h->d_kf_bak, sizeof(double) * 7 * D.n_kf, cudaMemcpyDeviceToDevice, h->stream));   // pop
                ORBX_CUDA(cudaMemcpyAsync(D.pt, h->d_pt_bak, sizeof(double) * 3 * D.n_pts, cudaMemcpyDeviceToDevice, h->stream));
            }
            qmax++;
        } while (rho < 0 && qmax < 10 && !stop_requested(h));
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
        if (nBad >= 3) break;
    }
    return ORBX_OK;
}

extern "C" orbx_status orbx_lba_solve_host(orbx_lba *h, const orbx_lba_problem *prob, int its1, int its2, orbx_lba_result *res) {
    if (!h || !prob || !res || !res->kf_pose || !res->pts || its1 < 0 || its2 < 0) return ORBX_ERR_INVALID;
    if (h->pending) {
        orbx_set_error("orbx_lba_solve_host: a window submitted with orbx_lba_solve_begin is still in flight on this handle");
        return ORBX_ERR_INVALID;
    }
    ORBX_CUDA(cudaSetDevice(h->device));
    h->launches = 0;
    res->lm_trials = 0; res->stopped = 0; res->first_lambda = 0;
    static const bool trace = getenv("ORBX_LBA_TRACE") != nullptr;     // host-side timing of the call's stages to stderr
    const auto t_0 = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count(); };
    orbx_status st = lba_load(h, prob, h->use_grid != 0);      // short chunks: the whole GPU works on this window
    if (st) return st;
    const double us_load = since(t_0);
    LbaDev &D = h->D;
    const int E = D.n_edges;
    ORBX_CUDA(cudaMemsetAsync(D.scal + 6, 0, sizeof(double) * 6, h->stream));
    if (stop_requested(h)) {       // Optimizer.cc:656-658
        res->stopped = 1;
        memcpy(res->kf_pose, prob->kf_pose, sizeof(double) * 7 * D.n_kf);
        memcpy(res->pts, prob->pts, sizeof(double) * 3 * D.n_pts);
        if (res->chi2) memset(res->chi2, 0, sizeof(double) * E);
        if (res->erase) memset(res->erase, 0, E);
        return ORBX_OK;
    }
    bool first = true;
    if ((st = lba_optimize(h, its1, 1, res, &first))) return st;
    if (!stop_requested(h) && its2 > 0) {
        k_lba_classify<<<blocks_for(E), LBA_THREADS, 0, h->stream>>>(D, D.level1);    // e->setLevel(1), Optimizer.cc:680-683
        h->launches++;
        if ((st = lba_optimize(h, its2, 0, res, &first))) return st;
    }
    k_lba_classify<<<blocks_for(E), LBA_THREADS, 0, h->stream>>>(D, h->d_flag);        // vToErase, Optimizer.cc:709-735
    h->launches++;
    ORBX_CUDA(cudaGetLastError());
    const double us_opt = since(t_0) - us_load;
    // through the handle's pinned staging: a copy into the caller's (pageable) arrays would block the host once per array
    ORBX_CUDA(cudaMemcpyAsync(h->h_kf, D.kf, sizeof(double) * 7 * D.n_kf, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(h->h_pt, D.pt, sizeof(double) * 3 * D.n_pts, cudaMemcpyDeviceToHost, h->stream));
    if (E) {
        ORBX_CUDA(cudaMemcpyAsync(h->h_chi, D.chi2, sizeof(double) * E, cudaMemcpyDeviceToHost, h->stream));
        ORBX_CUDA(cudaMemcpyAsync(h->h_flag, h->d_flag, E, cudaMemcpyDeviceToHost, h->stream));
    }
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(res->kf_pose, h->h_kf, sizeof(double) * 7 * D.n_kf);
    memcpy(res->pts, h->h_pt, sizeof(double) * 3 * D.n_pts);
    for (int s = 0; s < E; s++) {
        if (res->chi2) res->chi2[h->perm[s]] = h->h_chi[s];
        if (res->erase) res->erase[h->perm[s]] = h->h_flag[s];
    }
    if (trace) fprintf(stderr, "orbx_lba_solve_host: load %.0f us, optimize %.0f us, download %.0f us\n", us_load, us_opt, since(t_0) - us_load - us_opt);
    return ORBX_OK;
}

// ---- asynchronous form: many windows in flight, one handle (and stream) per window -----------------------------------
extern "C" orbx_status orbx_lba_solve_begin(orbx_lba *h, const orbx_lba_problem *prob, int its1, int its2) {
    if (!h || !prob || its1 < 1 || its2 < 0) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->device));
    h->launches = 0;
    h->pending = 0;
    orbx_status st = lba_load(h, prob);
    if (st) return st;
    LbaDev &D = h->D;
    if (!h->use_fused || !orbx_lba_fused_fits(D.n_kf, D.np)) {
        orbx_set_error("orbx_lba_solve_begin: %d free keyframes do not fit the single-kernel path; use orbx_lba_solve_host", D.np);
        return ORBX_ERR_UNSUPPORTED;
    }
    const int E = D.n_edges;
    h->pending = stop_requested(h) ? 2 : 1;          // 2 = stopped before the start (Optimizer.cc:656-658)
    if (h->pending == 2) return ORBX_OK;
    cudaStream_t s = h->stream;
    ORBX_CUDA(cudaMemsetAsync(D.scal + 4, 0, sizeof(double) * 8, s));
    // many windows in flight: the portable 8-CTA cluster, so that 16 of them fit the device side by side
    if ((st = orbx_lba_fused_launch(D, h->d_kf_bak, h->d_pt_bak, its1, 1, 0, nullptr, nullptr, nullptr, D.scal + 4, s, 0))) return st;
    if (its2 > 0) {
        k_lba_classify<<<blocks_for(E), LBA_THREADS, 0, s>>>(D, D.level1);
        if ((st = orbx_lba_fused_launch(D, h->d_kf_bak, h->d_pt_bak, its2, 0, 0, nullptr, nullptr, nullptr, D.scal + 4, s, 0))) return st;
    }
    k_lba_classify<<<blocks_for(E), LBA_THREADS, 0, s>>>(D, h->d_flag);
    h->launches = its2 > 0 ? 4 : 2;
    ORBX_CUDA(cudaGetLastError());
    ORBX_CUDA(cudaMemcpyAsync(h->h_kf, D.kf, sizeof(double) * 7 * D.n_kf, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(h->h_pt, D.pt, sizeof(double) * 3 * D.n_pts, cudaMemcpyDeviceToHost, s));
    if (E) {
        ORBX_CUDA(cudaMemcpyAsync(h->h_chi, D.chi2, sizeof(double) * E, cudaMemcpyDeviceToHost, s));
        ORBX_CUDA(cudaMemcpyAsync(h->h_flag, h->d_flag, E, cudaMemcpyDeviceToHost, s));
    }
    ORBX_CUDA(cudaMemcpyAsync(h->h_scal + 4, D.scal + 4, sizeof(double) * 8, cudaMemcpyDeviceToHost, s));
    return ORBX_OK;
}

extern "C" orbx_status orbx_lba_solve_end(orbx_lba *h, const orbx_lba_problem *prob, orbx_lba_result *res) {
    if (!h || !prob || !res || !res->kf_pose || !res->pts || !h->pending) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->device));
    const LbaDev &D = h->D;
    const int E = D.n_edges;
    res->lm_trials = 0; res->stopped = 0; res->first_lambda = 0;
    if (h->pending == 2) {
        res->stopped = 1;
        memcpy(res->kf_pose, prob->kf_pose, sizeof(double) * 7 * D.n_kf);
        memcpy(res->pts, prob->pts, sizeof(double) * 3 * D.n_pts);
        if (res->chi2) memset(res->chi2, 0, sizeof(double) * E);
        if (res->erase) memset(res->erase, 0, E);
        h->pending = 0;
        return ORBX_OK;
    }
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(res->kf_pose, h->h_kf, sizeof(double) * 7 * D.n_kf);
    memcpy(res->pts, h->h_pt, sizeof(double) * 3 * D.n_pts);
    for (int s = 0; s < E; s++) {
        if (res->chi2) res->chi2[h->perm[s]] = h->h_chi[s];
        if (res->erase) res->erase[h->perm[s]] = h->h_flag[s];
    }
    res->lm_trials = (int)h->h_scal[4];
    h->pending = 0;
    return ORBX_OK;
}

extern "C" orbx_status orbx_lba_build_schur_timed(orbx_lba *h, const orbx_lba_problem *prob, double lambda, int reps, float *ms,
                                                  double *Hschur, double *bschur) {
    if (!h || reps < 1) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->device));
    orbx_status st;
    if (prob && (st = lba_load(h, prob, true))) return st;
    if (!h->loaded) {
        orbx_set_error("orbx_lba_build_schur_timed: no problem loaded");
        return ORBX_ERR_INVALID;
    }
    h->launches = 0;
    LbaDev &D = h->D;
    ORBX_CUDA(cudaEventRecord(h->ev0, h->stream));
    for (int r = 0; r < reps; r++) {
        if ((st = lba_build(h, 1, 0))) return st;
        if ((st = lba_schur(h, lambda))) return st;
    }
    ORBX_CUDA(cudaEventRecord(h->ev1, h->stream));
    ORBX_CUDA(cudaEventSynchronize(h->ev1));
    if (ms) ORBX_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    if (Hschur) {
        std::vector<double> Hs((size_t)D.n * D.n);
        ORBX_CUDA(cudaMemcpy(Hs.data(), D.Hs, sizeof(double) * Hs.size(), cudaMemcpyDeviceToHost));
        for (int r = 0; r < D.n; r++)
            for (int c = 0; c < D.n; c++) Hschur[(size_t)r * D.n + c] = r / 6 <= c / 6 ? Hs[(size_t)r * D.n + c] : Hs[(size_t)c * D.n + r];
    }
    if (bschur) ORBX_CUDA(cudaMemcpy(bschur, D.bs, sizeof(double) * D.n, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

extern "C" int orbx_lba_last_launches(const orbx_lba *h) { return h ? h->launches : 0; }

/* diagnostics: nanoseconds the cluster kernel spent per phase in the last solve (build, schur, reduce, solve, update, err) */
extern "C" orbx_status orbx_lba_phase_ns(const orbx_lba *h, double out[6]) {
    if (!h || !out) return ORBX_ERR_INVALID;
    for (int i = 0; i < 6; i++) out[i] = h->h_scal[6 + i];
    return ORBX_OK;
}
