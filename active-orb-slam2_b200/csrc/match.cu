// Window + Hamming matchers: replaces ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)
// (reference src/ORBmatcher.cc:45-129), ORBmatcher::SearchByProjection(Cur, Last, th, mono) (:1328-1470),
// DescriptorDistance (:1647-1663), ComputeThreeMaxima (:1601-1642) and the Frame grid they search
// (src/Frame.cc:259-274 AssignFeaturesToGrid, :411-421 PosInGrid, :356-409 GetFeaturesInArea).
//
// One CTA per (frame, point set) job.  The reference loop is sequential because a keypoint claimed by an earlier
// point is skipped by later ones (`mvpMapPoints[idx]->Observations()>0`).  Here every point evaluates its window
// in parallel against "claimed by a point with a smaller index" (minclaim[k] < i) from the previous sweep, and
// sweeps repeat until no choice changes.  Point 0 is final after sweep 1, point j once all i<j are final, so the
// fixed point equals the sequential result; conflicts are rare and 2-3 sweeps are typical.
// Float arithmetic that decides a comparison is written with explicit round-to-nearest intrinsics in the
// reference's operation order (cv::Mat products are sequential float multiply-adds without FMA).
#include "orbx_internal.cuh"
#include "block_scan.cuh"

#define M_THREADS 256
#define GRID_COLS 64   // FRAME_GRID_COLS, Frame.h:38
#define GRID_ROWS 48   // FRAME_GRID_ROWS, Frame.h:37
#define NCELL (GRID_COLS * GRID_ROWS)
#define TH_HIGH 100    // ORBmatcher.cc:36
#define HISTO_LENGTH 30

struct orbx_matcher {
    int device, max_kp, max_pts, max_jobs;
    int *d_gidx;      // [jobs][max_kp]  keypoint indices in grid-cell order
    int *d_choice;    // [jobs][2][max_pts]
    int *d_minclaim;  // [jobs][2][max_kp]
    int *d_owner;     // [jobs][max_kp]
    // staging of the _host entry points (one job)
    orbx_keypoint *d_keys; uint8_t *d_desc; float *d_uright; uint8_t *d_claimed; float *d_scale;
    void *d_pts; uint8_t *d_ptdesc; int32_t *d_match; int32_t *d_nm; orbx_frame_match_job *d_job;
    cudaStream_t stream;
    int last_launches;
};

struct MatchShared {
    int start[NCELL + 1];
    int cur[NCELL];
    int warp_tmp[34];
    int hist[HISTO_LENGTH];
    int keep[3];
    int nacc, nrej;
};

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint8_t *b) {
    const uint4 b0 = *reinterpret_cast<const uint4 *>(b), b1 = *reinterpret_cast<const uint4 *>(b + 16);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// Frame::AssignFeaturesToGrid: cell = ix*GRID_ROWS + iy, indices ascending inside a cell (push_back order)
__device__ void grid_build(const orbx_frame_view &F, int n, MatchShared &sh, int *gidx) {
    const int tid = threadIdx.x;
    for (int c = tid; c <= NCELL; c += M_THREADS) sh.start[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += M_THREADS) {
        const int px = (int)roundf(__fmul_rn(__fsub_rn(F.keys_un[i].x, F.min_x), F.grid_w_inv));   // PosInGrid
        const int py = (int)roundf(__fmul_rn(__fsub_rn(F.keys_un[i].y, F.min_y), F.grid_h_inv));
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
        atomicAdd(&sh.start[px * GRID_ROWS + py], 1);
    }
    block_excl_scan(sh.start, NCELL + 1, sh.warp_tmp);
    for (int c = tid; c < NCELL; c += M_THREADS) sh.cur[c] = sh.start[c];
    __syncthreads();
    for (int i = tid; i < n; i += M_THREADS) {
        const int px = (int)roundf(__fmul_rn(__fsub_rn(F.keys_un[i].x, F.min_x), F.grid_w_inv));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(F.keys_un[i].y, F.min_y), F.grid_h_inv));
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
        gidx[atomicAdd(&sh.cur[px * GRID_ROWS + py], 1)] = i;
    }
    __syncthreads();
    for (int c = tid; c < NCELL; c += M_THREADS) {
        const int b = sh.start[c], e = sh.start[c + 1];
        for (int i = b + 1; i < e; i++) {
            const int k = gidx[i];
            int j = i - 1;
            while (j >= b && gidx[j] > k) { gidx[j + 1] = gidx[j]; j--; }
            gidx[j + 1] = k;
        }
    }
    __syncthreads();
}

// Frame::GetFeaturesInArea: calls fn(idx) for every keypoint of the window, in the reference's order
template <class Fn>
__device__ __forceinline__ void features_in_area(const orbx_frame_view &F, const MatchShared &sh, const int *gidx, float x,
                                                 float y, float r, int minLevel, int maxLevel, Fn fn) {
    const float dx0 = __fsub_rn(x, F.min_x), dy0 = __fsub_rn(y, F.min_y);
    int x0 = (int)floorf(__fmul_rn(__fsub_rn(dx0, r), F.grid_w_inv)); x0 = max(0, x0);
    if (x0 >= GRID_COLS) return;
    int x1 = (int)ceilf(__fmul_rn(__fadd_rn(dx0, r), F.grid_w_inv)); x1 = min(GRID_COLS - 1, x1);
    if (x1 < 0) return;
    int y0 = (int)floorf(__fmul_rn(__fsub_rn(dy0, r), F.grid_h_inv)); y0 = max(0, y0);
    if (y0 >= GRID_ROWS) return;
    int y1 = (int)ceilf(__fmul_rn(__fadd_rn(dy0, r), F.grid_h_inv)); y1 = min(GRID_ROWS - 1, y1);
    if (y1 < 0) return;
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = x0; ix <= x1; ix++)
        for (int iy = y0; iy <= y1; iy++) {
            const int c = ix * GRID_ROWS + iy;
            for (int j = sh.start[c]; j < sh.start[c + 1]; j++) {
                const int idx = gidx[j];
                const orbx_keypoint &kp = F.keys_un[idx];
                if (check) {
                    const int oct = kp.octave;
                    if (oct < minLevel) continue;
                    if (maxLevel >= 0 && oct > maxLevel) continue;
                }
                if (fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r) fn(idx);
            }
        }
}

// sweeps of "choose among the keypoints not claimed by a smaller index" until nothing changes
template <class Eval>
__device__ void resolve_claims(int n_kp, int n_pts, const uint8_t *claimed, int *choice2, int *minclaim2, int max_pts,
                               int max_kp, Eval eval, int &final_buf) {
    const int tid = threadIdx.x;
    int *mc[2] = {minclaim2, minclaim2 + max_kp};
    int *ch[2] = {choice2, choice2 + max_pts};
    for (int k = tid; k < n_kp; k += M_THREADS) mc[0][k] = (claimed && claimed[k]) ? -1 : 0x7fffffff;
    for (int i = tid; i < n_pts; i += M_THREADS) ch[0][i] = -2;
    __syncthreads();
    int cur = 0;
    for (int sweep = 0; sweep <= n_pts + 1; sweep++) {
        const int nxt = cur ^ 1;
        int changed = 0;
        for (int k = tid; k < n_kp; k += M_THREADS) mc[nxt][k] = (claimed && claimed[k]) ? -1 : 0x7fffffff;
        for (int i = tid; i < n_pts; i += M_THREADS) {
            const int c = eval(i, mc[cur]);
            ch[nxt][i] = c;
            changed |= c != ch[cur][i];
        }
        __syncthreads();
        for (int i = tid; i < n_pts; i += M_THREADS) {
            const int c = ch[nxt][i];
            if (c >= 0) atomicMin(&mc[nxt][c], eval.blocks(i) ? i : 0x7fffffff);
        }
        cur = nxt;
        if (!__syncthreads_or(changed)) break;
    }
    final_buf = cur;
}

// ---- SearchByProjection(CurrentFrame, LastFrame, th, bMono) ------------------------------------------------------
struct FrameEval {
    const orbx_frame_match_job &J;
    const orbx_frame_view &F;
    const MatchShared &sh;
    const int *gidx;
    __device__ bool blocks(int i) const { return J.pts[i].blocks != 0; }
    __device__ int operator()(int i, const int *minclaim) const {
        const orbx_last_point p = J.pts[i];
        if (!p.valid) return -1;
        const float *R = J.Rcw, *t = J.tcw;
        const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], p.x), __fmul_rn(R[1], p.y)), __fmul_rn(R[2], p.z)), t[0]);
        const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], p.x), __fmul_rn(R[4], p.y)), __fmul_rn(R[5], p.z)), t[1]);
        const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], p.x), __fmul_rn(R[7], p.y)), __fmul_rn(R[8], p.z)), t[2]);
        const float invzc = __double2float_rn(__ddiv_rn(1.0, (double)zc));     // const float invzc = 1.0/x3Dc.at<float>(2)
        if (invzc < 0) return -1;
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(F.fx, xc), invzc), F.cx);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(F.fy, yc), invzc), F.cy);
        if (u < F.min_x || u > F.max_x) return -1;
        if (v < F.min_y || v > F.max_y) return -1;
        const int oct = p.octave;
        const float radius = __fmul_rn(J.th, F.scale_factors[oct]);
        int minL, maxL;
        if (J.forward) { minL = oct; maxL = -1; }
        else if (J.backward) { minL = 0; maxL = oct; }
        else { minL = oct - 1; maxL = oct + 1; }
        const uint8_t *d = J.last_desc + (size_t)32 * i;
        const uint4 d0 = *reinterpret_cast<const uint4 *>(d), d1 = *reinterpret_cast<const uint4 *>(d + 16);
        const float ur = __fsub_rn(u, __fmul_rn(F.bf, invzc));
        int bestDist = 256, bestIdx = -1;
        features_in_area(F, sh, gidx, u, v, radius, minL, maxL, [&](int i2) {
            if (minclaim[i2] < i) return;
            if (F.u_right) {
                const float urk = F.u_right[i2];
                if (urk > 0 && fabsf(__fsub_rn(ur, urk)) > radius) return;
            }
            const int dist = hamming256(d0, d1, F.desc + (size_t)32 * i2);
            if (dist < bestDist) { bestDist = dist; bestIdx = i2; }
        });
        return bestDist <= TH_HIGH ? bestIdx : -1;
    }
};

__global__ void __launch_bounds__(M_THREADS)
k_match_frame(const orbx_frame_match_job *__restrict__ jobs, int *__restrict__ gidx_all, int *__restrict__ choice_all,
              int *__restrict__ minclaim_all, int *__restrict__ owner_all, int max_kp, int max_pts) {
    __shared__ MatchShared sh;
    __shared__ orbx_frame_match_job J;
    const int tid = threadIdx.x, job = blockIdx.x;
    if (tid == 0) J = jobs[job];
    __syncthreads();
    const orbx_frame_view &F = J.cur;
    const int n = min(F.n_dev ? *F.n_dev : F.n, max_kp), n_pts = min(J.n_last, max_pts);
    int *gidx = gidx_all + (size_t)job * max_kp, *choice2 = choice_all + (size_t)job * 2 * max_pts;
    int *minclaim2 = minclaim_all + (size_t)job * 2 * max_kp, *owner = owner_all + (size_t)job * max_kp;
    grid_build(F, n, sh, gidx);
    FrameEval ev{J, F, sh, gidx};
    int fb;
    resolve_claims(n, n_pts, F.claimed, choice2, minclaim2, max_pts, max_kp, ev, fb);
    const int *choice = choice2 + (size_t)fb * max_pts;
    // owner = last point that wrote mvpMapPoints[k]; rotation histogram over every accepted point (:1431-1446)
    for (int k = tid; k < n; k += M_THREADS) owner[k] = -1;
    if (tid < HISTO_LENGTH) sh.hist[tid] = 0;
    if (tid == 0) { sh.nacc = 0; sh.nrej = 0; }
    __syncthreads();
    const float factor = 1.0f / HISTO_LENGTH;
    int nacc = 0;
    for (int i = tid; i < n_pts; i += M_THREADS) {
        const int c = choice[i];
        if (c < 0) continue;
        nacc++;
        atomicMax(&owner[c], i);
        if (J.check_ori) {
            float rot = __fsub_rn(J.pts[i].angle, F.keys_un[c].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == HISTO_LENGTH) bin = 0;
            if (bin >= 0 && bin < HISTO_LENGTH) atomicAdd(&sh.hist[bin], 1);
        }
    }
    if (nacc) atomicAdd(&sh.nacc, nacc);
    __syncthreads();
    if (tid == 0) {   // ComputeThreeMaxima, ORBmatcher.cc:1601-1642
        int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < HISTO_LENGTH; i++) {
            const int s = sh.hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; i3 = i2; i2 = i1; i1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; i3 = i2; i2 = i; }
            else if (s > max3) { max3 = s; i3 = i; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { i3 = -1; }
        sh.keep[0] = i1; sh.keep[1] = i2; sh.keep[2] = i3;
    }
    __syncthreads();
    for (int k = tid; k < n; k += M_THREADS) if (owner[k] >= 0) J.match[k] = owner[k];
    __syncthreads();
    if (J.check_ori) {
        int nrej = 0;
        for (int i = tid; i < n_pts; i += M_THREADS) {
            const int c = choice[i];
            if (c < 0) continue;
            float rot = __fsub_rn(J.pts[i].angle, F.keys_un[c].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == HISTO_LENGTH) bin = 0;
            if (bin != sh.keep[0] && bin != sh.keep[1] && bin != sh.keep[2]) { J.match[c] = -1; nrej++; }
        }
        if (nrej) atomicAdd(&sh.nrej, nrej);
    }
    __syncthreads();
    if (tid == 0) *J.nmatches = sh.nacc - sh.nrej;
}

// ---- SearchByProjection(Frame &F, vpMapPoints, th) ------------------------------------------------------------
struct PointsJob {
    orbx_frame_view F;
    int n_pts;
    const orbx_track_point *pts;
    const uint8_t *pt_desc;
    float th, nnratio;
    int32_t *match, *nmatches;
};

struct PointsEval {
    const PointsJob &J;
    const orbx_frame_view &F;
    const MatchShared &sh;
    const int *gidx;
    __device__ bool blocks(int i) const { return J.pts[i].blocks != 0; }
    __device__ int operator()(int i, const int *minclaim) const {
        const orbx_track_point p = J.pts[i];
        if (!p.in_view) return -1;
        float r = p.view_cos > 0.998f ? 2.5f : 4.0f;             // RadiusByViewingCos
        if (J.th != 1.0f) r = __fmul_rn(r, J.th);
        const float rs = __fmul_rn(r, F.scale_factors[p.level]);
        const uint8_t *d = J.pt_desc + (size_t)32 * i;
        const uint4 d0 = *reinterpret_cast<const uint4 *>(d), d1 = *reinterpret_cast<const uint4 *>(d + 16);
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        features_in_area(F, sh, gidx, p.proj_x, p.proj_y, rs, p.level - 1, p.level, [&](int idx) {
            if (minclaim[idx] < i) return;
            if (F.u_right) {
                const float urk = F.u_right[idx];
                if (urk > 0 && fabsf(__fsub_rn(p.proj_xr, urk)) > rs) return;
            }
            const int dist = hamming256(d0, d1, F.desc + (size_t)32 * idx);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F.keys_un[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F.keys_un[idx].octave; bestDist2 = dist;
            }
        });
        if (bestDist > TH_HIGH) return -1;
        if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(J.nnratio, (float)bestDist2)) return -1;
        return bestIdx;
    }
};

__global__ void __launch_bounds__(M_THREADS)
k_match_points(const PointsJob *__restrict__ jobs, int *__restrict__ gidx_all, int *__restrict__ choice_all,
               int *__restrict__ minclaim_all, int *__restrict__ owner_all, int max_kp, int max_pts) {
    __shared__ MatchShared sh;
    __shared__ PointsJob J;
    const int tid = threadIdx.x, job = blockIdx.x;
    if (tid == 0) J = jobs[job];
    __syncthreads();
    const orbx_frame_view &F = J.F;
    const int n = min(F.n_dev ? *F.n_dev : F.n, max_kp), n_pts = min(J.n_pts, max_pts);
    int *gidx = gidx_all + (size_t)job * max_kp, *choice2 = choice_all + (size_t)job * 2 * max_pts;
    int *minclaim2 = minclaim_all + (size_t)job * 2 * max_kp, *owner = owner_all + (size_t)job * max_kp;
    grid_build(F, n, sh, gidx);
    PointsEval ev{J, F, sh, gidx};
    int fb;
    resolve_claims(n, n_pts, F.claimed, choice2, minclaim2, max_pts, max_kp, ev, fb);
    const int *choice = choice2 + (size_t)fb * max_pts;
    for (int k = tid; k < n; k += M_THREADS) owner[k] = -1;
    if (tid == 0) sh.nacc = 0;
    __syncthreads();
    int nacc = 0;
    for (int i = tid; i < n_pts; i += M_THREADS) {
        const int c = choice[i];
        if (c < 0) continue;
        nacc++;
        atomicMax(&owner[c], i);
    }
    if (nacc) atomicAdd(&sh.nacc, nacc);
    __syncthreads();
    for (int k = tid; k < n; k += M_THREADS) if (owner[k] >= 0) J.match[k] = owner[k];
    if (tid == 0) *J.nmatches = sh.nacc;
}

// ---- host side -------------------------------------------------------------------------------------------------
extern "C" int orbx_hamming256(const uint8_t a[32], const uint8_t b[32]) {
    int d = 0;
    for (int i = 0; i < 32; i += 8) {
        uint64_t x, y;
        memcpy(&x, a + i, 8);
        memcpy(&y, b + i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

extern "C" void orbx_matcher_destroy(orbx_matcher *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    cudaFree(m->d_gidx); cudaFree(m->d_choice); cudaFree(m->d_minclaim); cudaFree(m->d_owner); cudaFree(m->d_keys);
    cudaFree(m->d_desc); cudaFree(m->d_uright); cudaFree(m->d_claimed); cudaFree(m->d_scale); cudaFree(m->d_pts);
    cudaFree(m->d_ptdesc); cudaFree(m->d_match); cudaFree(m->d_nm); cudaFree(m->d_job);
    if (m->stream) cudaStreamDestroy(m->stream);
    free(m);
}

extern "C" orbx_status orbx_matcher_create(orbx_matcher **out, int max_keypoints, int max_points, int max_jobs, int device) {
    if (!out) return ORBX_ERR_INVALID;
    *out = nullptr;
    if (max_keypoints < 1 || max_points < 1 || max_jobs < 1) {
        orbx_set_error("orbx_matcher_create: bad argument");
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
    orbx_matcher *m = (orbx_matcher *)calloc(1, sizeof(orbx_matcher));
    if (!m) return ORBX_ERR_NOMEM;
    m->device = device; m->max_kp = max_keypoints; m->max_pts = max_points; m->max_jobs = max_jobs;
    const size_t kp = (size_t)max_keypoints, pt = (size_t)max_points, jb = (size_t)max_jobs;
    const size_t ptsz = sizeof(orbx_track_point) > sizeof(orbx_last_point) ? sizeof(orbx_track_point) : sizeof(orbx_last_point);
    cudaError_t ce = cudaSuccess;
#define TRY(x) if (ce == cudaSuccess) ce = (x)
    TRY(cudaMalloc((void **)&m->d_gidx, sizeof(int) * kp * jb));
    TRY(cudaMalloc((void **)&m->d_choice, sizeof(int) * 2 * pt * jb));
    TRY(cudaMalloc((void **)&m->d_minclaim, sizeof(int) * 2 * kp * jb));
    TRY(cudaMalloc((void **)&m->d_owner, sizeof(int) * kp * jb));
    TRY(cudaMalloc((void **)&m->d_keys, sizeof(orbx_keypoint) * kp));
    TRY(cudaMalloc((void **)&m->d_desc, 32 * kp));
    TRY(cudaMalloc((void **)&m->d_uright, sizeof(float) * kp));
    TRY(cudaMalloc((void **)&m->d_claimed, kp));
    TRY(cudaMalloc((void **)&m->d_scale, sizeof(float) * 64));
    TRY(cudaMalloc((void **)&m->d_pts, ptsz * pt));
    TRY(cudaMalloc((void **)&m->d_ptdesc, 32 * pt));
    TRY(cudaMalloc((void **)&m->d_match, sizeof(int32_t) * kp));
    TRY(cudaMalloc((void **)&m->d_nm, sizeof(int32_t)));
    TRY(cudaMalloc((void **)&m->d_job, sizeof(orbx_frame_match_job) > sizeof(PointsJob) ? sizeof(orbx_frame_match_job) : sizeof(PointsJob)));
    TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
#undef TRY
    if (ce != cudaSuccess) {
        orbx_set_error("orbx_matcher_create: %s", cudaGetErrorString(ce));
        orbx_matcher_destroy(m);
        return ORBX_ERR_CUDA;
    }
    *out = m;
    return ORBX_OK;
}

// copy the host frame view into the staging buffers and return the device-side view
static orbx_status stage_frame(orbx_matcher *m, const orbx_frame_view *F, orbx_frame_view *D, cudaStream_t s) {
    if (!F || F->n < 0 || (F->n > 0 && (!F->keys_un || !F->desc)) || !F->scale_factors || F->nlevels < 1 || F->nlevels > 64)
        return ORBX_ERR_INVALID;
    if (F->n > m->max_kp) {
        orbx_set_error("frame has %d keypoints, matcher was created for %d", F->n, m->max_kp);
        return ORBX_ERR_CAPACITY;
    }
    *D = *F;
    D->n_dev = nullptr;
    const size_t n = (size_t)F->n;
    if (n) {
        ORBX_CUDA(cudaMemcpyAsync(m->d_keys, F->keys_un, sizeof(orbx_keypoint) * n, cudaMemcpyHostToDevice, s));
        ORBX_CUDA(cudaMemcpyAsync(m->d_desc, F->desc, 32 * n, cudaMemcpyHostToDevice, s));
        if (F->u_right) ORBX_CUDA(cudaMemcpyAsync(m->d_uright, F->u_right, sizeof(float) * n, cudaMemcpyHostToDevice, s));
        if (F->claimed) ORBX_CUDA(cudaMemcpyAsync(m->d_claimed, F->claimed, n, cudaMemcpyHostToDevice, s));
    }
    ORBX_CUDA(cudaMemcpyAsync(m->d_scale, F->scale_factors, sizeof(float) * F->nlevels, cudaMemcpyHostToDevice, s));
    D->keys_un = m->d_keys;
    D->desc = m->d_desc;
    D->u_right = F->u_right ? m->d_uright : nullptr;
    D->claimed = F->claimed ? m->d_claimed : nullptr;
    D->scale_factors = m->d_scale;
    return ORBX_OK;
}

extern "C" orbx_status orbx_match_projection_frame_device(orbx_matcher *m, const orbx_frame_match_job *d_jobs, int n_jobs,
                                                          void *stream) {
    if (!m || n_jobs < 0 || (n_jobs && !d_jobs)) return ORBX_ERR_INVALID;
    if (n_jobs > m->max_jobs) {
        orbx_set_error("%d jobs, matcher was created for %d", n_jobs, m->max_jobs);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaSetDevice(m->device));
    m->last_launches = 0;
    if (n_jobs == 0) return ORBX_OK;
    k_match_frame<<<n_jobs, M_THREADS, 0, (cudaStream_t)stream>>>(d_jobs, m->d_gidx, m->d_choice, m->d_minclaim, m->d_owner,
                                                                  m->max_kp, m->max_pts);
    m->last_launches = 1;
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

extern "C" orbx_status orbx_match_projection_frame_host(orbx_matcher *m, const orbx_frame_view *cur, int n_last,
                                                        const orbx_last_point *pts, const uint8_t *last_desc, const float Rcw[9],
                                                        const float tcw[3], int forward, int backward, float th, int check_ori,
                                                        int32_t *match, int32_t *nmatches) {
    if (!m || !cur || n_last < 0 || !match || !nmatches || !Rcw || !tcw || (n_last && (!pts || !last_desc))) return ORBX_ERR_INVALID;
    if (n_last > m->max_pts) {
        orbx_set_error("%d points, matcher was created for %d", n_last, m->max_pts);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    orbx_frame_match_job J;
    memset(&J, 0, sizeof(J));
    orbx_status st = stage_frame(m, cur, &J.cur, s);
    if (st) return st;
    for (int l = 0; l < n_last; l++)
        if (pts[l].valid && (pts[l].octave < 0 || pts[l].octave >= cur->nlevels)) return ORBX_ERR_INVALID;
    if (n_last) {
        ORBX_CUDA(cudaMemcpyAsync(m->d_pts, pts, sizeof(orbx_last_point) * n_last, cudaMemcpyHostToDevice, s));
        ORBX_CUDA(cudaMemcpyAsync(m->d_ptdesc, last_desc, (size_t)32 * n_last, cudaMemcpyHostToDevice, s));
    }
    if (cur->n) ORBX_CUDA(cudaMemcpyAsync(m->d_match, match, sizeof(int32_t) * cur->n, cudaMemcpyHostToDevice, s));
    J.n_last = n_last;
    J.pts = (const orbx_last_point *)m->d_pts;
    J.last_desc = m->d_ptdesc;
    memcpy(J.Rcw, Rcw, sizeof(J.Rcw));
    memcpy(J.tcw, tcw, sizeof(J.tcw));
    J.forward = forward; J.backward = backward; J.th = th; J.check_ori = check_ori;
    J.match = m->d_match; J.nmatches = m->d_nm;
    ORBX_CUDA(cudaMemcpyAsync(m->d_job, &J, sizeof(J), cudaMemcpyHostToDevice, s));
    st = orbx_match_projection_frame_device(m, m->d_job, 1, s);
    if (st) return st;
    if (cur->n) ORBX_CUDA(cudaMemcpyAsync(match, m->d_match, sizeof(int32_t) * cur->n, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(nmatches, m->d_nm, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" orbx_status orbx_match_projection_points_host(orbx_matcher *m, const orbx_frame_view *F, int n_pts,
                                                         const orbx_track_point *pts, const uint8_t *pt_desc, float th,
                                                         float nnratio, int32_t *match, int32_t *nmatches) {
    if (!m || !F || n_pts < 0 || !match || !nmatches || (n_pts && (!pts || !pt_desc))) return ORBX_ERR_INVALID;
    if (n_pts > m->max_pts) {
        orbx_set_error("%d points, matcher was created for %d", n_pts, m->max_pts);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    PointsJob J;
    memset(&J, 0, sizeof(J));
    orbx_status st = stage_frame(m, F, &J.F, s);
    if (st) return st;
    for (int l = 0; l < n_pts; l++)
        if (pts[l].in_view && (pts[l].level < 0 || pts[l].level >= F->nlevels)) return ORBX_ERR_INVALID;
    if (n_pts) {
        ORBX_CUDA(cudaMemcpyAsync(m->d_pts, pts, sizeof(orbx_track_point) * n_pts, cudaMemcpyHostToDevice, s));
        ORBX_CUDA(cudaMemcpyAsync(m->d_ptdesc, pt_desc, (size_t)32 * n_pts, cudaMemcpyHostToDevice, s));
    }
    if (F->n) ORBX_CUDA(cudaMemcpyAsync(m->d_match, match, sizeof(int32_t) * F->n, cudaMemcpyHostToDevice, s));
    J.n_pts = n_pts;
    J.pts = (const orbx_track_point *)m->d_pts;
    J.pt_desc = m->d_ptdesc;
    J.th = th; J.nnratio = nnratio;
    J.match = m->d_match; J.nmatches = m->d_nm;
    ORBX_CUDA(cudaMemcpyAsync(m->d_job, &J, sizeof(J), cudaMemcpyHostToDevice, s));
    k_match_points<<<1, M_THREADS, 0, s>>>((const PointsJob *)m->d_job, m->d_gidx, m->d_choice, m->d_minclaim, m->d_owner,
                                           m->max_kp, m->max_pts);
    m->last_launches = 1;
    ORBX_CUDA(cudaGetLastError());
    if (F->n) ORBX_CUDA(cudaMemcpyAsync(match, m->d_match, sizeof(int32_t) * F->n, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(nmatches, m->d_nm, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" int orbx_matcher_last_launches(const orbx_matcher *m) { return m ? m->last_launches : 0; }
