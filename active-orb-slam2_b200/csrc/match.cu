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
#include <cooperative_groups.h>
#include "orbx_internal.cuh"
#include "block_scan.cuh"
namespace cg = cooperative_groups;

#define M_THREADS 1024
#define M_WARPS (M_THREADS / 32)
#define GRID_COLS 64   // FRAME_GRID_COLS, Frame.h:38
#define GRID_ROWS 48   // FRAME_GRID_ROWS, Frame.h:37
#define NCELL (GRID_COLS * GRID_ROWS)
#define TH_HIGH 100    // ORBmatcher.cc:36
#define HISTO_LENGTH 30
#define M_MAX_KP 8192  // keypoints per frame the shared-memory grid can hold

struct orbx_matcher {
    int device, max_kp, max_pts, max_jobs;
    int smem;         // dynamic shared memory of the match kernels
    int *d_choice;    // [jobs][2][max_pts]
    int *d_minclaim;  // [jobs][2][max_kp]
    int *d_owner;     // [jobs][max_kp]
    int *d_sweeps;    // [jobs] sweeps the claim resolution took (diagnostics)
    int *d_novf;      // [jobs] overflow-list counters of the clustered frame kernel (zero between launches)
    int sm_count;
    int2 *d_cand;     // [jobs][M_CAND][max_pts] candidate (key, pack) per point, entry-major
    int *d_lcount;    // [jobs][2][max_pts] candidates per point; overflow list
    // staging of the _host entry points (one job)
    orbx_keypoint *d_keys; uint8_t *d_desc; float *d_uright; uint8_t *d_claimed; float *d_scale;
    void *d_pts; uint8_t *d_ptdesc; int32_t *d_match; int32_t *d_nm; orbx_frame_match_job *d_job;
    int *d_items;     // [jobs][2][max_pts] bucket matchers: work item -> (feature of a, node of b)
    uint8_t *d_arena; size_t arena_cap, arena_used;   // staging of the bucket matchers' host inputs
    cudaStream_t stream;
    int last_launches;
};

// static part of the shared state; the dynamic part holds the grid (see MatchGrid)
struct MatchShared {
    int warp_tmp[34];
    int hist[HISTO_LENGTH];
    int keep[3];
    int nacc, nrej, changed, novf;
};

// The frame's keypoints re-ordered by grid cell (cell = ix*GRID_ROWS + iy, ascending index inside a cell, which
// is the push_back order of Frame::AssignFeaturesToGrid), resident in shared memory.  Because cells of one grid
// column are adjacent, the part of a search window that lies in column ix is ONE contiguous range.
struct MatchGrid {
    int *start;     // NCELL + 1
    int *cur;       // NCELL (build only)
    float *x, *y;   // mvKeysUn[idx].pt
    int *oct;       // mvKeysUn[idx].octave
    int *idx;       // original keypoint index
    __device__ void carve(int *base, int max_kp) {
        start = base; cur = base + NCELL + 1;
        x = reinterpret_cast<float *>(cur + NCELL); y = x + max_kp;
        oct = reinterpret_cast<int *>(y + max_kp); idx = oct + max_kp;
    }
};
static size_t match_smem_bytes(int max_kp) { return sizeof(int) * (2 * NCELL + 1) + (size_t)16 * max_kp; }

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint8_t *b) {
    const uint4 b0 = __ldg(reinterpret_cast<const uint4 *>(b)), b1 = __ldg(reinterpret_cast<const uint4 *>(b + 16));
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__device__ __forceinline__ int grid_cell(const orbx_frame_view &F, float kx, float ky) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(kx, F.min_x), F.grid_w_inv));   // Frame::PosInGrid
    const int py = (int)roundf(__fmul_rn(__fsub_rn(ky, F.min_y), F.grid_h_inv));
    if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) return -1;
    return px * GRID_ROWS + py;
}

// Frame::AssignFeaturesToGrid
__device__ void grid_build(const orbx_frame_view &F, int n, MatchShared &sh, MatchGrid &g) {
    const int tid = threadIdx.x;
    for (int c = tid; c <= NCELL; c += M_THREADS) g.start[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += M_THREADS) {
        const int c = grid_cell(F, F.keys_un[i].x, F.keys_un[i].y);
        if (c >= 0) atomicAdd(&g.start[c], 1);
    }
    block_excl_scan(g.start, NCELL + 1, sh.warp_tmp);
    for (int c = tid; c < NCELL; c += M_THREADS) g.cur[c] = g.start[c];
    __syncthreads();
    for (int i = tid; i < n; i += M_THREADS) {
        const int c = grid_cell(F, F.keys_un[i].x, F.keys_un[i].y);
        if (c >= 0) g.idx[atomicAdd(&g.cur[c], 1)] = i;
    }
    __syncthreads();
    for (int c = tid; c < NCELL; c += M_THREADS) {
        const int b = g.start[c], e = g.start[c + 1];
        for (int i = b + 1; i < e; i++) {
            const int k = g.idx[i];
            int j = i - 1;
            while (j >= b && g.idx[j] > k) { g.idx[j + 1] = g.idx[j]; j--; }
            g.idx[j + 1] = k;
        }
    }
    __syncthreads();
    for (int j = tid; j < g.start[NCELL]; j += M_THREADS) {
        const orbx_keypoint &kp = F.keys_un[g.idx[j]];
        g.x[j] = kp.x; g.y[j] = kp.y; g.oct[j] = kp.octave;
    }
    __syncthreads();
}

// Frame::GetFeaturesInArea, warp-cooperative, in two phases so that the global loads of a window are issued
// together instead of one grid column at a time:
//   1. lanes scan the window column by column in shared memory and compact the keypoints that pass the level and
//      |dx|,|dy| < r tests into a per-warp list of (grid position, seq);
//   2. whenever 32 entries are waiting (and at the end) every lane takes one entry and ALL lanes call
//      fn(valid, keypoint index, seq).
// seq increases in the reference's visiting order (column by column, cell by cell, push order), so the reference's
// "first strict minimum" is the minimum of (dist, seq).
#define M_LIST 64
template <class Fn>
__device__ __forceinline__ void features_in_area(const orbx_frame_view &F, const MatchGrid &g, int2 *list, float x, float y,
                                                 float r, int minLevel, int maxLevel, Fn fn) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
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
    int seq0 = 0, count = 0;
    for (int ix = x0; ix <= x1; ix++) {
        const int s = g.start[ix * GRID_ROWS + y0], e = g.start[ix * GRID_ROWS + y1 + 1];
        for (int jb = s; jb < e; jb += 32) {
            const int j = jb + lane;
            bool pass = j < e;
            if (pass && check) {
                const int oct = g.oct[j];
                pass = oct >= minLevel && !(maxLevel >= 0 && oct > maxLevel);
            }
            if (pass) pass = fabsf(__fsub_rn(g.x[j], x)) < r && fabsf(__fsub_rn(g.y[j], y)) < r;
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if (pass) list[count + __popc(m & lt)] = make_int2(j, seq0 + (j - s));
            count += __popc(m);
            __syncwarp();
            if (count >= 32) {
                const int2 c = list[lane];
                fn(true, g.idx[c.x], c.y);
                __syncwarp();
                if (lane < count - 32) list[lane] = list[lane + 32];   // count - 32 < 32: no overlap
                count -= 32;
                __syncwarp();
            }
        }
        seq0 += e - s;
    }
    if (count > 0) {
        const bool valid = lane < count;
        const int2 c = valid ? list[lane] : make_int2(0, 0);
        fn(valid, valid ? g.idx[c.x] : 0, c.y);
    }
    __syncwarp();
}

__device__ __forceinline__ unsigned warp_min(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// A candidate is (key, pack): key = dist << 22 | seq (unique per point, ordered like the reference's scan),
// pack = keypoint index | octave << 16.  An Eval provides
//   candidates(i, emit)  warp-cooperative enumeration of the candidates of point i that pass every test that does
//                        not depend on claims; ALL lanes call emit(valid, key, pack) once per round of <= 32
//   decide(k1,p1,k2,p2)  the reference's accept rule from the two smallest unclaimed candidates
//   blocks(i)            Observations() > 0 of point i
struct Best2 {
    unsigned k1 = 0xffffffffu, k2 = 0xffffffffu;
    int p1 = -1, p2 = -1;
    __device__ __forceinline__ void add(unsigned key, int pack) {
        if (key < k1) { k2 = k1; p2 = p1; k1 = key; p1 = pack; }
        else if (key < k2) { k2 = key; p2 = pack; }
    }
};

// claim-dependent choice of point i by a whole warp (used for points whose candidate list overflowed)
template <class Eval>
__device__ int eval_warp(const Eval &ev, int i, const int *minclaim) {
    Best2 b;
    ev.candidates(i, [&](bool valid, unsigned key, int pack) {
        if (valid && minclaim[pack & 0xffff] >= i) b.add(key, pack);
    });
    const unsigned w1 = warp_min(b.k1);
    if (w1 == 0xffffffffu) return -1;
    const int src = __ffs(__ballot_sync(0xffffffffu, b.k1 == w1)) - 1;
    const int p1 = __shfl_sync(0xffffffffu, b.p1, src);
    const bool me = (threadIdx.x & 31) == src;
    const unsigned cand = me ? b.k2 : b.k1;
    const int candp = me ? b.p2 : b.p1;
    const unsigned w2 = warp_min(cand);
    int p2 = -1;
    if (w2 != 0xffffffffu) p2 = __shfl_sync(0xffffffffu, candp, __ffs(__ballot_sync(0xffffffffu, cand == w2)) - 1);
    return ev.decide(w1, p1, w2, p2);
}

#define M_CAND 64   // candidates kept per point; points with more fall back to eval_warp in every sweep

// The reference's sequential claim semantics as a fixed point.  Candidate lists are built once (they do not depend
// on claims); every sweep then lets each point pick among the candidates not claimed by a smaller index in the
// previous sweep, until no choice changes.
// Candidate lists of the points part, part + nparts, ... (warp per point).  With nparts > 1 the CTAs of a thread-block cluster
// share the work of one job: the lists, their lengths and the overflow list live in global memory, the overflow counter is
// *novf_global (zero on entry, reset by the caller after use).
template <class Eval>
__device__ void build_candidates(int n_pts, int2 *cand, int *lcount, int *ovf, int max_pts, MatchShared &sh, const Eval &ev,
                                 int part, int nparts, int *novf_global) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    if (tid == 0) sh.novf = 0;
    __syncthreads();
    for (int i = warp + part * M_WARPS; i < n_pts; i += M_WARPS * nparts) {
        int cnt = 0;
        ev.candidates(i, [&](bool valid, unsigned key, int pack) {
            const unsigned m = __ballot_sync(0xffffffffu, valid);
            const int pos = cnt + __popc(m & lt);
            if (valid && pos < M_CAND) cand[(size_t)pos * max_pts + i] = make_int2((int)key, pack);
            cnt += __popc(m);
        });
        if (lane == 0) {
            lcount[i] = cnt;
            if (cnt > M_CAND) ovf[atomicAdd(novf_global ? novf_global : &sh.novf, 1)] = i;
        }
    }
    __syncthreads();
}

template <class Eval>
__device__ void run_sweeps(int n_kp, int n_pts, const uint8_t *claimed, int *choice2, int *minclaim2, int2 *cand,
                           int *lcount, int *ovf, int max_pts, int max_kp, MatchShared &sh, const Eval &ev, int &final_buf, int novf);

template <class Eval>
__device__ void resolve_claims(int n_kp, int n_pts, const uint8_t *claimed, int *choice2, int *minclaim2, int2 *cand,
                               int *lcount, int *ovf, int max_pts, int max_kp, MatchShared &sh, const Eval &ev, int &final_buf) {
    build_candidates(n_pts, cand, lcount, ovf, max_pts, sh, ev, 0, 1, nullptr);
    run_sweeps(n_kp, n_pts, claimed, choice2, minclaim2, cand, lcount, ovf, max_pts, max_kp, sh, ev, final_buf, sh.novf);
}

template <class Eval>
__device__ void run_sweeps(int n_kp, int n_pts, const uint8_t *claimed, int *choice2, int *minclaim2, int2 *cand,
                           int *lcount, int *ovf, int max_pts, int max_kp, MatchShared &sh, const Eval &ev, int &final_buf, int novf) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int *mc[2] = {minclaim2, minclaim2 + max_kp};
    int *ch[2] = {choice2, choice2 + max_pts};
    for (int k = tid; k < n_kp; k += M_THREADS) mc[0][k] = (claimed && claimed[k]) ? -1 : 0x7fffffff;
    for (int i = tid; i < n_pts; i += M_THREADS) ch[0][i] = -2;
    __syncthreads();
    int cur = 0, sweeps = 0;
    for (int sweep = 0; sweep <= n_pts + 1; sweep++) {
        const int nxt = cur ^ 1;
        int changed = 0;
        const int *mcc = mc[cur];
        for (int k = tid; k < n_kp; k += M_THREADS) mc[nxt][k] = (claimed && claimed[k]) ? -1 : 0x7fffffff;
        for (int i = tid; i < n_pts; i += M_THREADS) {
            const int cnt = lcount[i];
            if (cnt > M_CAND) continue;
            Best2 b;
            for (int e = 0; e < cnt; e++) {
                const int2 c = cand[(size_t)e * max_pts + i];
                if (mcc[c.y & 0xffff] >= i) b.add((unsigned)c.x, c.y);
            }
            const int c = b.k1 == 0xffffffffu ? -1 : ev.decide(b.k1, b.p1, b.k2, b.p2);
            changed |= c != ch[cur][i];
            ch[nxt][i] = c;
        }
        for (int o = warp; o < novf; o += M_WARPS) {
            const int i = ovf[o];
            const int c = eval_warp(ev, i, mcc);
            if (lane == 0) {
                changed |= c != ch[cur][i];
                ch[nxt][i] = c;
            }
        }
        __syncthreads();
        for (int i = tid; i < n_pts; i += M_THREADS) {
            const int c = ch[nxt][i];
            if (c >= 0 && ev.blocks(i)) atomicMin(&mc[nxt][c], i);
        }
        cur = nxt;
        sweeps++;
        if (!__syncthreads_or(changed)) break;
    }
    final_buf = cur;
    if (tid == 0) sh.changed = sweeps;
}

// ---- SearchByProjection(CurrentFrame, LastFrame, th, bMono) ------------------------------------------------------
struct FrameEval {
    const orbx_frame_match_job &J;
    const orbx_frame_view &F;
    const MatchGrid &g;
    int2 *list;      // this warp's compaction list (M_LIST entries)
    __device__ bool blocks(int i) const { return J.variant == 1 || J.pts[i].blocks != 0; }
    __device__ int decide(unsigned k1, int p1, unsigned, int) const {
        const int lim = J.max_dist > 0 ? J.max_dist : TH_HIGH;     // ORBmatcher.cc:1421 (TH_HIGH) / :1553 (ORBdist)
        return (int)(k1 >> 22) <= lim ? (p1 & 0xffff) : -1;
    }
    template <class Emit>
    __device__ void candidates(int i, Emit emit) const {
        const orbx_last_point p = J.pts[i];
        if (!p.valid) return;
        const float *R = J.Rcw, *t = J.tcw;
        const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], p.x), __fmul_rn(R[1], p.y)), __fmul_rn(R[2], p.z)), t[0]);
        const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], p.x), __fmul_rn(R[4], p.y)), __fmul_rn(R[5], p.z)), t[1]);
        const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], p.x), __fmul_rn(R[7], p.y)), __fmul_rn(R[8], p.z)), t[2]);
        const float invzc = __double2float_rn(__ddiv_rn(1.0, (double)zc));     // const float invzc = 1.0/x3Dc.at<float>(2)
        if (J.variant == 0 && invzc < 0) return;                 // the KeyFrame overload has no depth-sign test
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(F.fx, xc), invzc), F.cx);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(F.fy, yc), invzc), F.cy);
        if (u < F.min_x || u > F.max_x) return;
        if (v < F.min_y || v > F.max_y) return;
        const int oct = p.octave;
        const float radius = __fmul_rn(J.th, F.scale_factors[oct]);
        int minL, maxL;
        if (J.variant == 0 && J.forward) { minL = oct; maxL = -1; }
        else if (J.variant == 0 && J.backward) { minL = 0; maxL = oct; }
        else { minL = oct - 1; maxL = oct + 1; }
        const uint8_t *d = J.last_desc + (size_t)32 * i;
        const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(d)), d1 = __ldg(reinterpret_cast<const uint4 *>(d + 16));
        const float ur = __fsub_rn(u, __fmul_rn(F.bf, invzc));
        features_in_area(F, g, list, u, v, radius, minL, maxL, [&](bool valid, int i2, int seq) {
            unsigned key = 0;
            if (valid) {
                const float urk = (F.u_right && J.variant == 0) ? F.u_right[i2] : -1.f;
                key = ((unsigned)hamming256(d0, d1, F.desc + (size_t)32 * i2) << 22) | (unsigned)seq;
                if (urk > 0 && fabsf(__fsub_rn(ur, urk)) > radius) valid = false;      // ORBmatcher.cc:1405-1411
            }
            emit(valid, key, i2);
        });
    }
};

__global__ void __launch_bounds__(M_THREADS)
k_match_frame(const orbx_frame_match_job *__restrict__ jobs, int *__restrict__ choice_all, int *__restrict__ minclaim_all,
              int *__restrict__ owner_all, int *__restrict__ sweeps_all, int2 *__restrict__ cand_all, int *__restrict__ lcount_all,
              int *__restrict__ novf_all, int max_kp, int max_pts) {
    extern __shared__ __align__(16) int dyn[];
    __shared__ MatchShared sh;
    __shared__ orbx_frame_match_job J;
    __shared__ MatchGrid g;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
    const int tid = threadIdx.x, job = blockIdx.x / csize;
    if (tid == 0) { J = jobs[job]; g.carve(dyn, max_kp); }
    __syncthreads();
    const orbx_frame_view &F = J.cur;
    const int n = min(F.n_dev ? *F.n_dev : F.n, max_kp), n_pts = min(J.n_last, max_pts);
    int *choice2 = choice_all + (size_t)job * 2 * max_pts;
    int *minclaim2 = minclaim_all + (size_t)job * 2 * max_kp, *owner = owner_all + (size_t)job * max_kp;
    grid_build(F, n, sh, g);
    __shared__ int2 lists[M_WARPS][M_LIST];
    FrameEval ev{J, F, g, lists[tid >> 5]};
    int fb;
    // the CTAs of the cluster each build the grid and a share of the candidate lists; CTA 0 then resolves the claims alone
    int2 *cand = cand_all + (size_t)job * M_CAND * max_pts;
    int *lcount = lcount_all + (size_t)job * 2 * max_pts, *ovf = lcount + max_pts;
    build_candidates(n_pts, cand, lcount, ovf, max_pts, sh, ev, crank, csize, &novf_all[job]);
    cluster.sync();
    if (crank != 0) return;
    const int novf = novf_all[job];
    __syncthreads();
    if (tid == 0) novf_all[job] = 0;
    run_sweeps(n, n_pts, F.claimed, choice2, minclaim2, cand, lcount, ovf, max_pts, max_kp, sh, ev, fb, novf);
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
    if (tid == 0) { *J.nmatches = sh.nacc - sh.nrej; sweeps_all[job] = sh.changed; }
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
    const MatchGrid &g;
    int2 *list;
    __device__ bool blocks(int i) const { return J.pts[i].blocks != 0; }
    // The reference keeps (best, second) in visiting order (ORBmatcher.cc:101-114); that equals: best = first
    // occurrence of the minimum, second = first occurrence of the minimum among the rest, i.e. the two smallest keys.
    __device__ int decide(unsigned k1, int p1, unsigned k2, int p2) const {
        const int bestDist = (int)(k1 >> 22);
        if (bestDist > TH_HIGH) return -1;
        const int bestLevel = p1 >> 16;
        const int bestDist2 = k2 == 0xffffffffu ? 256 : (int)(k2 >> 22), bestLevel2 = k2 == 0xffffffffu ? -1 : (p2 >> 16);
        if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(J.nnratio, (float)bestDist2)) return -1;   // :118-121
        return p1 & 0xffff;
    }
    template <class Emit>
    __device__ void candidates(int i, Emit emit) const {
        const orbx_track_point p = J.pts[i];
        if (!p.in_view) return;
        float r = p.view_cos > 0.998f ? 2.5f : 4.0f;             // RadiusByViewingCos
        if (J.th != 1.0f) r = __fmul_rn(r, J.th);
        const float rs = __fmul_rn(r, F.scale_factors[p.level]);
        const uint8_t *d = J.pt_desc + (size_t)32 * i;
        const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(d)), d1 = __ldg(reinterpret_cast<const uint4 *>(d + 16));
        features_in_area(F, g, list, p.proj_x, p.proj_y, rs, p.level - 1, p.level, [&](bool valid, int idx, int seq) {
            unsigned key = 0;
            int pack = 0;
            if (valid) {
                const float urk = F.u_right ? F.u_right[idx] : -1.f;
                pack = idx | (F.keys_un[idx].octave << 16);
                key = ((unsigned)hamming256(d0, d1, F.desc + (size_t)32 * idx) << 22) | (unsigned)seq;
                if (urk > 0 && fabsf(__fsub_rn(p.proj_xr, urk)) > rs) valid = false;   // ORBmatcher.cc:91-96
            }
            emit(valid, key, pack);
        });
    }
};

__global__ void __launch_bounds__(M_THREADS)
k_match_points(const PointsJob *__restrict__ jobs, int *__restrict__ choice_all, int *__restrict__ minclaim_all,
               int *__restrict__ owner_all, int *__restrict__ sweeps_all, int2 *__restrict__ cand_all, int *__restrict__ lcount_all,
              int max_kp, int max_pts) {
    extern __shared__ __align__(16) int dyn[];
    __shared__ MatchShared sh;
    __shared__ PointsJob J;
    __shared__ MatchGrid g;
    const int tid = threadIdx.x, job = blockIdx.x;
    if (tid == 0) { J = jobs[job]; g.carve(dyn, max_kp); }
    __syncthreads();
    const orbx_frame_view &F = J.F;
    const int n = min(F.n_dev ? *F.n_dev : F.n, max_kp), n_pts = min(J.n_pts, max_pts);
    int *choice2 = choice_all + (size_t)job * 2 * max_pts;
    int *minclaim2 = minclaim_all + (size_t)job * 2 * max_kp, *owner = owner_all + (size_t)job * max_kp;
    grid_build(F, n, sh, g);
    __shared__ int2 lists[M_WARPS][M_LIST];
    PointsEval ev{J, F, g, lists[tid >> 5]};
    int fb;
    resolve_claims(n, n_pts, F.claimed, choice2, minclaim2, cand_all + (size_t)job * M_CAND * max_pts,
                   lcount_all + (size_t)job * 2 * max_pts, lcount_all + (size_t)job * 2 * max_pts + max_pts, max_pts, max_kp, sh, ev, fb);
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
    if (tid == 0) { *J.nmatches = sh.nacc; sweeps_all[job] = sh.changed; }
}

// ---- SearchForInitialization (ORBmatcher.cc:405-520) ----------------------------------------------------------------------
struct InitJob {
    orbx_frame_view F2;
    int n1;
    const orbx_keypoint *keys1;
    const uint8_t *desc1;
    const float *prev_xy;
    int window, check_ori;
    float nnratio;
    int32_t *match12, *nmatches;
    int *matched_dist, *match21;     // [F2.n] scratch
};

__global__ void __launch_bounds__(M_THREADS) k_match_init(const InitJob *__restrict__ jobs, int max_kp) {
    extern __shared__ __align__(16) int dyn[];
    __shared__ MatchShared sh;
    __shared__ InitJob J;
    __shared__ MatchGrid g;
    __shared__ int2 lists[M_WARPS][M_LIST];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) { J = jobs[blockIdx.x]; g.carve(dyn, max_kp); }
    __syncthreads();
    const orbx_frame_view &F2 = J.F2;
    const int n2 = min(F2.n, max_kp), n1 = J.n1;
    grid_build(F2, n2, sh, g);
    for (int k = tid; k < n2; k += M_THREADS) { J.matched_dist[k] = 0x7fffffff; J.match21[k] = -1; }
    for (int i = tid; i < n1; i += M_THREADS) J.match12[i] = -1;
    if (tid < HISTO_LENGTH) sh.hist[tid] = 0;
    if (tid == 0) sh.nacc = 0;
    __syncthreads();
    const float factor = 1.0f / HISTO_LENGTH;
    if (tid < 32) {
        // strictly sequential over F1's keypoints: a closer match may steal a keypoint from an earlier one (:470-474)
        int nmatches = 0;
        for (int i1 = 0; i1 < n1; i1++) {
            const orbx_keypoint kp1 = J.keys1[i1];
            if (kp1.octave > 0) continue;
            const uint8_t *d = J.desc1 + (size_t)32 * i1;
            const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(d)), d1 = __ldg(reinterpret_cast<const uint4 *>(d + 16));
            Best2 b;
            features_in_area(F2, g, lists[0], J.prev_xy[2 * i1], J.prev_xy[2 * i1 + 1], (float)J.window, kp1.octave, kp1.octave,
                             [&](bool valid, int i2, int seq) {
                                 if (!valid) return;
                                 const int dist = hamming256(d0, d1, F2.desc + (size_t)32 * i2);
                                 if (J.matched_dist[i2] <= dist) return;
                                 b.add(((unsigned)dist << 22) | (unsigned)seq, i2);
                             });
            const unsigned w1 = warp_min(b.k1);
            if (w1 == 0xffffffffu) continue;
            const int src = __ffs(__ballot_sync(0xffffffffu, b.k1 == w1)) - 1;
            const int bestIdx2 = __shfl_sync(0xffffffffu, b.p1, src);
            const unsigned w2 = warp_min(lane == src ? b.k2 : b.k1);
            const int bestDist = (int)(w1 >> 22);
            const float bestDist2 = w2 == 0xffffffffu ? 2147483647.0f : (float)(int)(w2 >> 22);      // (float)INT_MAX
            if (bestDist <= 50 && (float)bestDist < __fmul_rn(bestDist2, J.nnratio)) {
                if (lane == 0) {
                    const int old = J.match21[bestIdx2];
                    if (old >= 0) { J.match12[old] = -1; nmatches--; }
                    J.match12[i1] = bestIdx2;
                    J.match21[bestIdx2] = i1;
                    J.matched_dist[bestIdx2] = bestDist;
                    nmatches++;
                    if (J.check_ori) {
                        float rot = __fsub_rn(kp1.angle, F2.keys_un[bestIdx2].angle);
                        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                        int bin = (int)roundf(__fmul_rn(rot, factor));
                        if (bin == HISTO_LENGTH) bin = 0;
                        sh.hist[bin]++;
                        J.match21[bestIdx2] = i1;
                    }
                }
                __threadfence_block();
                __syncwarp();
            }
        }
        if (lane == 0) sh.nacc = nmatches;
    }
    __syncthreads();
    // rotation histogram entries are the accepted i1 in order, including ones whose match was stolen later: the reference
    // only removes an entry's match if it still has one (:503-507).  An accepted i1 is recognised by its histogram bin being
    // recomputable only while it still holds a match; stolen ones have match12 == -1 and are skipped there as well.  The bin
    // COUNTS however include stolen entries, which is why they were accumulated above at acceptance time.
    if (J.check_ori) {
        if (tid == 0) {
            int max1 = 0, max2 = 0, max3 = 0, a = -1, b = -1, c = -1;
            for (int i = 0; i < HISTO_LENGTH; i++) {
                const int s = sh.hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; c = b; b = a; a = i; }
                else if (s > max2) { max3 = max2; max2 = s; c = b; b = i; }
                else if (s > max3) { max3 = s; c = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { b = -1; c = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { c = -1; }
            sh.keep[0] = a; sh.keep[1] = b; sh.keep[2] = c;
            sh.nrej = 0;
        }
        __syncthreads();
        int nrej = 0;
        for (int i1 = tid; i1 < n1; i1 += M_THREADS) {
            const int i2 = J.match12[i1];
            if (i2 < 0) continue;
            float rot = __fsub_rn(J.keys1[i1].angle, F2.keys_un[i2].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == HISTO_LENGTH) bin = 0;
            if (bin != sh.keep[0] && bin != sh.keep[1] && bin != sh.keep[2]) { J.match12[i1] = -1; nrej++; }
        }
        if (nrej) atomicAdd(&sh.nrej, nrej);
        __syncthreads();
        if (tid == 0) sh.nacc -= sh.nrej;
    }
    __syncthreads();
    if (tid == 0) *J.nmatches = sh.nacc;
}

// ---- window + Hamming core of SearchByProjection(KF, Scw, ...), Fuse x2, SearchBySim3 ---------------------------------------
struct WindowJob {
    orbx_frame_view F;
    int n_pts;
    const orbx_window_point *pts;
    const uint8_t *pt_desc;
    int flags, max_dist;
    const float *inv_sigma2;
    int32_t *best_idx, *best_dist, *nacc;
};

struct WindowEval {
    const WindowJob &J;
    const orbx_frame_view &F;
    const MatchGrid &g;
    int2 *list;
    __device__ bool blocks(int) const { return (J.flags & 2) != 0; }
    __device__ int decide(unsigned k1, int p1, unsigned, int) const {
        return (int)(k1 >> 22) <= J.max_dist ? (p1 & 0xffff) : -1;
    }
    template <class Emit>
    __device__ void candidates(int i, Emit emit) const {
        const orbx_window_point p = J.pts[i];
        if (!p.valid) return;
        const uint8_t *d = J.pt_desc + (size_t)32 * i;
        const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(d)), d1 = __ldg(reinterpret_cast<const uint4 *>(d + 16));
        // level window applied like the reference does after KeyFrame::GetFeaturesInArea: oct < min || oct > max -> skip
        features_in_area(F, g, list, p.u, p.v, p.radius, -1, -1, [&](bool valid, int idx, int seq) {
            unsigned key = 0;
            if (valid) {
                const orbx_keypoint kp = F.keys_un[idx];
                if (kp.octave < p.min_level || kp.octave > p.max_level) valid = false;
                if (valid && (J.flags & 1)) {       // Fuse: reprojection error gate, ORBmatcher.cc:903-927
                    const float ex = __fsub_rn(p.u, kp.x), ey = __fsub_rn(p.v, kp.y);
                    const float urk = F.u_right ? F.u_right[idx] : -1.f;
                    float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                    double lim = 5.99;
                    if (urk >= 0) { const float er = __fsub_rn(p.ur, urk); e2 = __fadd_rn(e2, __fmul_rn(er, er)); lim = 7.8; }
                    if ((double)__fmul_rn(e2, J.inv_sigma2[kp.octave]) > lim) valid = false;
                }
                if (valid) key = ((unsigned)hamming256(d0, d1, F.desc + (size_t)32 * idx) << 22) | (unsigned)seq;
            }
            emit(valid, key, idx);
        });
    }
};

__global__ void __launch_bounds__(M_THREADS)
k_match_window(const WindowJob *__restrict__ jobs, int *__restrict__ choice_all, int *__restrict__ minclaim_all,
               int *__restrict__ sweeps_all, int2 *__restrict__ cand_all, int *__restrict__ lcount_all, int max_kp, int max_pts) {
    extern __shared__ __align__(16) int dyn[];
    __shared__ MatchShared sh;
    __shared__ WindowJob J;
    __shared__ MatchGrid g;
    __shared__ int2 lists[M_WARPS][M_LIST];
    const int tid = threadIdx.x, job = blockIdx.x;
    if (tid == 0) { J = jobs[job]; g.carve(dyn, max_kp); }
    __syncthreads();
    const orbx_frame_view &F = J.F;
    const int n = min(F.n_dev ? *F.n_dev : F.n, max_kp), n_pts = min(J.n_pts, max_pts);
    int *choice2 = choice_all + (size_t)job * 2 * max_pts, *minclaim2 = minclaim_all + (size_t)job * 2 * max_kp;
    int2 *cand = cand_all + (size_t)job * M_CAND * max_pts;
    int *lcount = lcount_all + (size_t)job * 2 * max_pts;
    grid_build(F, n, sh, g);
    WindowEval ev{J, F, g, lists[tid >> 5]};
    int fb;
    resolve_claims(n, n_pts, (J.flags & 2) ? F.claimed : nullptr, choice2, minclaim2, cand, lcount, lcount + max_pts, max_pts, max_kp, sh, ev, fb);
    const int *choice = choice2 + (size_t)fb * max_pts;
    if (tid == 0) sh.nacc = 0;
    __syncthreads();
    // the distance of the chosen candidate: look it up in the point's list (overflowed lists recompute it)
    int nacc = 0;
    for (int i = tid; i < n_pts; i += M_THREADS) {
        const int c = choice[i];
        int dist = 256;
        if (c >= 0) {
            nacc++;
            const uint8_t *d = J.pt_desc + (size_t)32 * i;
            const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(d)), d1 = __ldg(reinterpret_cast<const uint4 *>(d + 16));
            dist = hamming256(d0, d1, F.desc + (size_t)32 * c);
        }
        J.best_idx[i] = c;
        J.best_dist[i] = dist;
    }
    if (nacc) atomicAdd(&sh.nacc, nacc);
    __syncthreads();
    if (tid == 0) { *J.nacc = sh.nacc; sweeps_all[job] = sh.changed; }
}

// ---- SearchByBoW (KF, F) / (KF, KF) and SearchForTriangulation --------------------------------------------------
struct BucketDev {
    orbx_bucket_job J;          // all pointers are device pointers
    int32_t *match_a, *nmatches;
};

// work item = one feature of set a that lives in a node shared with set b, in the reference's processing order
struct BucketEval {
    const BucketDev &Jd;
    const int *item_a, *item_bnode;    // per item: feature index in a, node index in b
    __device__ bool blocks(int) const { return Jd.J.mode != 2; }        // SearchForTriangulation never sets vbMatched2
    __device__ int decide(unsigned k1, int p1, unsigned k2, int) const {
        const orbx_bucket_job &J = Jd.J;
        const int d1 = (int)(k1 >> 22);
        if (J.mode == 2) return p1 & 0xffff;                              // candidates are already <= TH_LOW and gated
        const int d2 = k2 == 0xffffffffu ? 256 : (int)(k2 >> 22);
        const bool pass = J.mode == 0 ? d1 <= 50 : d1 < 50;               // TH_LOW; ORBmatcher.cc:233 vs :598
        if (!pass || !((float)d1 < __fmul_rn(J.nnratio, (float)d2))) return -1;
        return p1 & 0xffff;
    }
    template <class Emit>
    __device__ void candidates(int i, Emit emit) const {
        const orbx_bucket_job &J = Jd.J;
        const int lane = threadIdx.x & 31;
        const int idx1 = item_a[i], nb = item_bnode[i];
        if (!J.a.valid[idx1]) return;
        const bool st1 = J.a.u_right && J.a.u_right[idx1] >= 0;
        if (J.mode == 2 && J.only_stereo && !st1) return;
        const uint8_t *d = J.a.desc + (size_t)32 * idx1;
        const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(d)), d1 = __ldg(reinterpret_cast<const uint4 *>(d + 16));
        const int s = J.b.node_start[nb], e = J.b.node_start[nb + 1];
        float la = 0, lb = 0, lc = 0, den = 0;
        if (J.mode == 2) {     // epipolar line of kp1 in image 2, CheckDistEpipolarLine (ORBmatcher.cc:140-157)
            const float x = J.a.keys_un[idx1].x, y = J.a.keys_un[idx1].y;
            la = __fadd_rn(__fadd_rn(__fmul_rn(x, J.F12[0]), __fmul_rn(y, J.F12[3])), J.F12[6]);
            lb = __fadd_rn(__fadd_rn(__fmul_rn(x, J.F12[1]), __fmul_rn(y, J.F12[4])), J.F12[7]);
            lc = __fadd_rn(__fadd_rn(__fmul_rn(x, J.F12[2]), __fmul_rn(y, J.F12[5])), J.F12[8]);
            den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
        }
        for (int kb = s; kb < e; kb += 32) {
            const int k = kb + lane;
            bool valid = k < e;
            unsigned key = 0;
            int idx2 = 0;
            if (valid) {
                idx2 = J.b.node_feat[k];
                if (J.mode != 0 && !J.b.valid[idx2]) valid = false;
            }
            if (valid) {
                const int dist = hamming256(d0, d1, J.b.desc + (size_t)32 * idx2);
                const int seq = k - s;
                if (J.mode == 2) {
                    // the reference keeps a candidate when dist <= running best: the LAST of equal distances wins
                    key = ((unsigned)dist << 22) | (unsigned)(0x3fffff - seq);
                    const bool st2 = J.b.u_right && J.b.u_right[idx2] >= 0;
                    if (J.only_stereo && !st2) valid = false;
                    if (dist > 50) valid = false;
                    const orbx_keypoint kp2 = J.b.keys_un[idx2];
                    if (valid && !st1 && !st2) {
                        const float dx = __fsub_rn(J.ex, kp2.x), dy = __fsub_rn(J.ey, kp2.y);
                        if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, J.scale_b[kp2.octave])) valid = false;
                    }
                    if (valid) {
                        const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, kp2.x), __fmul_rn(lb, kp2.y)), lc);
                        if (den == 0) valid = false;
                        else {
                            const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
                            valid = (double)dsqr < __dmul_rn(3.84, (double)J.sigma2_b[kp2.octave]);
                        }
                    }
                } else {
                    key = ((unsigned)dist << 22) | (unsigned)seq;
                }
            }
            emit(valid, key, idx2);
        }
    }
};

__global__ void __launch_bounds__(M_THREADS)
k_match_buckets(const BucketDev *__restrict__ jobs, int *__restrict__ choice_all, int *__restrict__ minclaim_all,
                int *__restrict__ items_all, int *__restrict__ sweeps_all, int2 *__restrict__ cand_all, int *__restrict__ lcount_all,
                int max_kp, int max_pts) {
    extern __shared__ __align__(16) int dyn[];      // node scan: [n_nodes_a + 1]
    __shared__ MatchShared sh;
    __shared__ BucketDev Jd;
    const int tid = threadIdx.x, job = blockIdx.x;
    if (tid == 0) Jd = jobs[job];
    __syncthreads();
    const orbx_bucket_job &J = Jd.J;
    const int nA = min(J.a.n, max_pts), nB = min(J.b.n, max_kp);
    int *choice2 = choice_all + (size_t)job * 2 * max_pts, *minclaim2 = minclaim_all + (size_t)job * 2 * max_kp;
    int *item_a = items_all + (size_t)job * 2 * max_pts, *item_bnode = item_a + max_pts;
    // merge-join of the two FeatureVectors (ORBmatcher.cc:176-258): every node of a looks its id up in b
    int *off = dyn;
    const int nna = J.a.n_nodes;
    for (int ia = tid; ia <= nna; ia += M_THREADS) {
        int cnt = 0;
        if (ia < nna) {
            const uint32_t id = J.a.node_id[ia];
            int lo = 0, hi = J.b.n_nodes;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (J.b.node_id[mid] < id) lo = mid + 1; else hi = mid; }
            if (lo < J.b.n_nodes && J.b.node_id[lo] == id) cnt = J.a.node_start[ia + 1] - J.a.node_start[ia];
        }
        off[ia] = cnt;
    }
    const int n_items = min(block_excl_scan(off, nna + 1, sh.warp_tmp), max_pts);
    for (int ia = tid; ia < nna; ia += M_THREADS) {
        const int cnt = (ia + 1 <= nna ? off[ia + 1] : n_items) - off[ia];
        if (cnt <= 0) continue;
        const uint32_t id = J.a.node_id[ia];
        int lo = 0, hi = J.b.n_nodes;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (J.b.node_id[mid] < id) lo = mid + 1; else hi = mid; }
        for (int k = 0; k < cnt && off[ia] + k < max_pts; k++) {
            item_a[off[ia] + k] = J.a.node_feat[J.a.node_start[ia] + k];
            item_bnode[off[ia] + k] = lo;
        }
    }
    __syncthreads();
    BucketEval ev{Jd, item_a, item_bnode};
    int fb;
    resolve_claims(nB, n_items, nullptr, choice2, minclaim2, cand_all + (size_t)job * M_CAND * max_pts,
                   lcount_all + (size_t)job * 2 * max_pts, lcount_all + (size_t)job * 2 * max_pts + max_pts, max_pts, max_kp, sh, ev, fb);
    const int *choice = choice2 + (size_t)fb * max_pts;
    for (int i = tid; i < nA; i += M_THREADS) Jd.match_a[i] = -1;
    if (tid < HISTO_LENGTH) sh.hist[tid] = 0;
    if (tid == 0) { sh.nacc = 0; sh.nrej = 0; }
    __syncthreads();
    const float factor = 1.0f / HISTO_LENGTH;
    int nacc = 0;
    for (int i = tid; i < n_items; i += M_THREADS) {
        const int c = choice[i];
        if (c < 0) continue;
        nacc++;
        const int idx1 = item_a[i];
        Jd.match_a[idx1] = c;                 // a feature appears in exactly one node: no write conflict
        if (J.check_ori) {
            float rot = __fsub_rn(J.a.keys_un[idx1].angle, J.b.keys_un[c].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == HISTO_LENGTH) bin = 0;
            if (bin >= 0 && bin < HISTO_LENGTH) atomicAdd(&sh.hist[bin], 1);
        }
    }
    if (nacc) atomicAdd(&sh.nacc, nacc);
    __syncthreads();
    if (tid == 0) {   // ComputeThreeMaxima
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
    if (J.check_ori) {
        int nrej = 0;
        for (int i = tid; i < n_items; i += M_THREADS) {
            const int c = choice[i];
            if (c < 0) continue;
            const int idx1 = item_a[i];
            float rot = __fsub_rn(J.a.keys_un[idx1].angle, J.b.keys_un[c].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == HISTO_LENGTH) bin = 0;
            if (bin != sh.keep[0] && bin != sh.keep[1] && bin != sh.keep[2]) { Jd.match_a[idx1] = -1; nrej++; }
        }
        if (nrej) atomicAdd(&sh.nrej, nrej);
    }
    __syncthreads();
    if (tid == 0) { *Jd.nmatches = sh.nacc - sh.nrej; sweeps_all[job] = sh.changed; }
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
    cudaFree(m->d_choice); cudaFree(m->d_minclaim); cudaFree(m->d_owner); cudaFree(m->d_sweeps); cudaFree(m->d_novf); cudaFree(m->d_items); cudaFree(m->d_arena); cudaFree(m->d_cand); cudaFree(m->d_lcount); cudaFree(m->d_keys);
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
    if (max_keypoints > M_MAX_KP) {
        orbx_set_error("orbx_matcher_create: at most %d keypoints per frame fit the shared-memory grid", M_MAX_KP);
        return ORBX_ERR_UNSUPPORTED;
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
    TRY(cudaMalloc((void **)&m->d_choice, sizeof(int) * 2 * pt * jb));
    TRY(cudaMalloc((void **)&m->d_minclaim, sizeof(int) * 2 * kp * jb));
    TRY(cudaMalloc((void **)&m->d_owner, sizeof(int) * kp * jb));
    TRY(cudaMalloc((void **)&m->d_sweeps, sizeof(int) * jb));
    TRY(cudaMalloc((void **)&m->d_novf, sizeof(int) * jb));
    TRY(cudaMemset(m->d_novf, 0, sizeof(int) * jb));
    m->sm_count = prop.multiProcessorCount;
    TRY(cudaMalloc((void **)&m->d_items, sizeof(int) * 2 * pt * jb));
    m->arena_cap = 2 * (kp + pt) * (28 + 32 + 4 + 1 + 4 + 8) + 4096;
    TRY(cudaMalloc((void **)&m->d_arena, m->arena_cap));
    TRY(cudaMalloc((void **)&m->d_cand, sizeof(int2) * M_CAND * pt * jb));
    TRY(cudaMalloc((void **)&m->d_lcount, sizeof(int) * 2 * pt * jb));
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
    m->smem = (int)match_smem_bytes(max_keypoints);
    TRY(ORBX_RAISE_SMEM(k_match_frame));
    TRY(ORBX_RAISE_SMEM(k_match_points));
    TRY(ORBX_RAISE_SMEM(k_match_buckets));
    TRY(ORBX_RAISE_SMEM(k_match_window));
    TRY(ORBX_RAISE_SMEM(k_match_init));
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
    // a thread-block cluster per job: as many CTAs as the SMs allow (1, 2, 4 or 8) share the candidate lists of one frame
    int csize = 1;
    // (8 only for a handful of frames, where the kernel is a latency chain on otherwise idle SMs: with 16 frames per launch and other
    //  sub-batches' kernels next to it, 8 CTAs per frame cost the pipelined step 6 %)
    const int cmax = n_jobs <= 4 ? 8 : 4;
    while (csize < cmax && 2 * csize * n_jobs <= m->sm_count) csize *= 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_jobs * csize));
    cfg.blockDim = dim3(M_THREADS);
    cfg.dynamicSmemBytes = (size_t)m->smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ORBX_CUDA(cudaLaunchKernelEx(&cfg, k_match_frame, d_jobs, m->d_choice, m->d_minclaim, m->d_owner, m->d_sweeps, m->d_cand, m->d_lcount,
                                 m->d_novf, m->max_kp, m->max_pts));
    m->last_launches = 1;
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

static orbx_status match_frame_host(orbx_matcher *m, const orbx_frame_view *cur, int n_last, const orbx_last_point *pts,
                                    const uint8_t *last_desc, const float Rcw[9], const float tcw[3], int forward, int backward,
                                    float th, int check_ori, int max_dist, int variant, int32_t *match, int32_t *nmatches) {
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
    J.max_dist = max_dist; J.variant = variant;
    J.match = m->d_match; J.nmatches = m->d_nm;
    ORBX_CUDA(cudaMemcpyAsync(m->d_job, &J, sizeof(J), cudaMemcpyHostToDevice, s));
    st = orbx_match_projection_frame_device(m, m->d_job, 1, s);
    if (st) return st;
    if (cur->n) ORBX_CUDA(cudaMemcpyAsync(match, m->d_match, sizeof(int32_t) * cur->n, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(nmatches, m->d_nm, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" orbx_status orbx_match_projection_frame_host(orbx_matcher *m, const orbx_frame_view *cur, int n_last,
                                                        const orbx_last_point *pts, const uint8_t *last_desc, const float Rcw[9],
                                                        const float tcw[3], int forward, int backward, float th, int check_ori,
                                                        int32_t *match, int32_t *nmatches) {
    return match_frame_host(m, cur, n_last, pts, last_desc, Rcw, tcw, forward, backward, th, check_ori, TH_HIGH, 0, match, nmatches);
}

extern "C" orbx_status orbx_match_projection_keyframe_host(orbx_matcher *m, const orbx_frame_view *cur, int n_pts,
                                                           const orbx_last_point *pts, const uint8_t *pt_desc, const float Rcw[9],
                                                           const float tcw[3], float th, int orb_dist, int check_ori, int32_t *match,
                                                           int32_t *nmatches) {
    if (orb_dist < 1 || orb_dist > 256) return ORBX_ERR_INVALID;
    return match_frame_host(m, cur, n_pts, pts, pt_desc, Rcw, tcw, 0, 0, th, check_ori, orb_dist, 1, match, nmatches);
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
    k_match_points<<<1, M_THREADS, m->smem, s>>>((const PointsJob *)m->d_job, m->d_choice, m->d_minclaim, m->d_owner, m->d_sweeps, m->d_cand, m->d_lcount, m->max_kp,
                                                 m->max_pts);
    m->last_launches = 1;
    ORBX_CUDA(cudaGetLastError());
    if (F->n) ORBX_CUDA(cudaMemcpyAsync(match, m->d_match, sizeof(int32_t) * F->n, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(nmatches, m->d_nm, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" int orbx_matcher_last_launches(const orbx_matcher *m) { return m ? m->last_launches : 0; }

extern "C" orbx_status orbx_matcher_last_sweeps(orbx_matcher *m, int32_t *out, int n_jobs) {
    if (!m || !out || n_jobs < 0 || n_jobs > m->max_jobs) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(m->device));
    ORBX_CUDA(cudaDeviceSynchronize());
    ORBX_CUDA(cudaMemcpy(out, m->d_sweeps, sizeof(int32_t) * n_jobs, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

// bump allocator over the staging arena: copies `bytes` from the host and returns the device address
static orbx_status arena_put(orbx_matcher *m, const void *src, size_t bytes, const void **dev, cudaStream_t s) {
    *dev = nullptr;
    if (!src || bytes == 0) return ORBX_OK;
    const size_t off = (m->arena_used + 15) & ~(size_t)15;
    if (off + bytes > m->arena_cap) {
        orbx_set_error("bucket matcher inputs exceed the staging arena (%zu bytes)", m->arena_cap);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaMemcpyAsync(m->d_arena + off, src, bytes, cudaMemcpyHostToDevice, s));
    *dev = m->d_arena + off;
    m->arena_used = off + bytes;
    return ORBX_OK;
}

static orbx_status stage_bow_set(orbx_matcher *m, const orbx_bow_set *H, orbx_bow_set *D, int cap, cudaStream_t s) {
    if (H->n < 0 || H->n_nodes < 0 || (H->n && (!H->keys_un || !H->desc || !H->valid)) ||
        (H->n_nodes && (!H->node_id || !H->node_start || !H->node_feat)))
        return ORBX_ERR_INVALID;
    if (H->n > cap) {
        orbx_set_error("%d features, matcher was created for %d", H->n, cap);
        return ORBX_ERR_CAPACITY;
    }
    const int nf = H->n_nodes ? H->node_start[H->n_nodes] : 0;
    if (nf < 0 || nf > H->n) return ORBX_ERR_INVALID;
    for (int i = 0; i < H->n_nodes; i++) {
        if (H->node_start[i] > H->node_start[i + 1] || (i && H->node_id[i - 1] >= H->node_id[i])) return ORBX_ERR_INVALID;
    }
    for (int k = 0; k < nf; k++) if (H->node_feat[k] < 0 || H->node_feat[k] >= H->n) return ORBX_ERR_INVALID;
    *D = *H;
    orbx_status st;
    const void *p;
    if ((st = arena_put(m, H->keys_un, sizeof(orbx_keypoint) * H->n, &p, s))) return st; D->keys_un = (const orbx_keypoint *)p;
    if ((st = arena_put(m, H->desc, (size_t)32 * H->n, &p, s))) return st; D->desc = (const uint8_t *)p;
    if ((st = arena_put(m, H->u_right, H->u_right ? sizeof(float) * H->n : 0, &p, s))) return st; D->u_right = (const float *)p;
    if ((st = arena_put(m, H->valid, H->n, &p, s))) return st; D->valid = (const uint8_t *)p;
    if ((st = arena_put(m, H->node_id, sizeof(uint32_t) * H->n_nodes, &p, s))) return st; D->node_id = (const uint32_t *)p;
    if ((st = arena_put(m, H->node_start, sizeof(int32_t) * (H->n_nodes + 1), &p, s))) return st; D->node_start = (const int32_t *)p;
    if ((st = arena_put(m, H->node_feat, sizeof(int32_t) * nf, &p, s))) return st; D->node_feat = (const int32_t *)p;
    return ORBX_OK;
}

extern "C" orbx_status orbx_match_buckets_host(orbx_matcher *m, const orbx_bucket_job *job, int32_t *match_a, int32_t *nmatches) {
    if (!m || !job || !match_a || !nmatches || job->mode < 0 || job->mode > 2) return ORBX_ERR_INVALID;
    if (job->mode == 2 && (!job->sigma2_b || !job->scale_b || job->nlevels < 1 || job->nlevels > 64)) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    m->arena_used = 0;
    BucketDev D;
    memset(&D, 0, sizeof(D));
    D.J = *job;
    orbx_status st;
    if ((st = stage_bow_set(m, &job->a, &D.J.a, m->max_pts, s))) return st;
    if ((st = stage_bow_set(m, &job->b, &D.J.b, m->max_kp, s))) return st;
    if (job->mode == 2) {
        for (int i = 0; i < job->b.n; i++)
            if (job->b.keys_un[i].octave < 0 || job->b.keys_un[i].octave >= job->nlevels) return ORBX_ERR_INVALID;
        const void *p;
        if ((st = arena_put(m, job->sigma2_b, sizeof(float) * job->nlevels, &p, s))) return st; D.J.sigma2_b = (const float *)p;
        if ((st = arena_put(m, job->scale_b, sizeof(float) * job->nlevels, &p, s))) return st; D.J.scale_b = (const float *)p;
    }
    D.match_a = m->d_match;       // max_kp entries; set a is bounded by max_pts: use the larger of the two buffers
    if (job->a.n > m->max_kp) {
        orbx_set_error("%d features in set a, matcher was created for %d keypoints", job->a.n, m->max_kp);
        return ORBX_ERR_CAPACITY;
    }
    D.nmatches = m->d_nm;
    const void *djob;
    if ((st = arena_put(m, &D, sizeof(D), &djob, s))) return st;
    const size_t smem = sizeof(int) * ((size_t)job->a.n_nodes + 2);
    k_match_buckets<<<1, M_THREADS, smem, s>>>((const BucketDev *)djob, m->d_choice, m->d_minclaim, m->d_items, m->d_sweeps, m->d_cand,
                                               m->d_lcount, m->max_kp, m->max_pts);
    m->last_launches = 1;
    ORBX_CUDA(cudaGetLastError());
    if (job->a.n) ORBX_CUDA(cudaMemcpyAsync(match_a, m->d_match, sizeof(int32_t) * job->a.n, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(nmatches, m->d_nm, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" orbx_status orbx_match_window_host(orbx_matcher *m, const orbx_frame_view *F, int n_pts, const orbx_window_point *pts,
                                              const uint8_t *pt_desc, int flags, const float *inv_sigma2, int max_dist,
                                              int32_t *best_idx, int32_t *best_dist, int32_t *n_accepted) {
    if (!m || !F || n_pts < 0 || !best_idx || !best_dist || (n_pts && (!pts || !pt_desc)) || max_dist < 0 || max_dist > 256 ||
        ((flags & 1) && !inv_sigma2))
        return ORBX_ERR_INVALID;
    if (n_pts > m->max_pts) {
        orbx_set_error("%d points, matcher was created for %d", n_pts, m->max_pts);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    WindowJob J;
    memset(&J, 0, sizeof(J));
    orbx_status st = stage_frame(m, F, &J.F, s);
    if (st) return st;
    m->arena_used = 0;
    const void *p;
    if ((st = arena_put(m, pts, sizeof(orbx_window_point) * n_pts, &p, s))) return st; J.pts = (const orbx_window_point *)p;
    if ((st = arena_put(m, pt_desc, (size_t)32 * n_pts, &p, s))) return st; J.pt_desc = (const uint8_t *)p;
    if (flags & 1) { if ((st = arena_put(m, inv_sigma2, sizeof(float) * F->nlevels, &p, s))) return st; J.inv_sigma2 = (const float *)p; }
    J.n_pts = n_pts; J.flags = flags; J.max_dist = max_dist;
    // outputs live in the scratch that the other matchers use for `owner` (max_kp ints) -- sized for points here
    if ((st = arena_put(m, best_idx, sizeof(int32_t) * (n_pts ? n_pts : 1), &p, s))) return st; J.best_idx = (int32_t *)p;
    if ((st = arena_put(m, best_dist, sizeof(int32_t) * (n_pts ? n_pts : 1), &p, s))) return st; J.best_dist = (int32_t *)p;
    J.nacc = m->d_nm;
    const void *djob;
    if ((st = arena_put(m, &J, sizeof(J), &djob, s))) return st;
    k_match_window<<<1, M_THREADS, m->smem, s>>>((const WindowJob *)djob, m->d_choice, m->d_minclaim, m->d_sweeps, m->d_cand, m->d_lcount,
                                                 m->max_kp, m->max_pts);
    m->last_launches = 1;
    ORBX_CUDA(cudaGetLastError());
    if (n_pts) {
        ORBX_CUDA(cudaMemcpyAsync(best_idx, J.best_idx, sizeof(int32_t) * n_pts, cudaMemcpyDeviceToHost, s));
        ORBX_CUDA(cudaMemcpyAsync(best_dist, J.best_dist, sizeof(int32_t) * n_pts, cudaMemcpyDeviceToHost, s));
    }
    int32_t nacc = 0;
    ORBX_CUDA(cudaMemcpyAsync(&nacc, m->d_nm, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    if (n_accepted) *n_accepted = nacc;
    return ORBX_OK;
}

extern "C" orbx_status orbx_match_initialization_host(orbx_matcher *m, const orbx_frame_view *F1, const orbx_frame_view *F2,
                                                      const float *prev_xy, int window_size, float nnratio, int check_ori,
                                                      int32_t *match12, int32_t *nmatches) {
    if (!m || !F1 || !F2 || !match12 || !nmatches || F1->n < 0 || (F1->n && (!F1->keys_un || !F1->desc || !prev_xy)) || window_size < 0)
        return ORBX_ERR_INVALID;
    if (F1->n > m->max_pts) {
        orbx_set_error("%d keypoints in F1, matcher was created for %d points", F1->n, m->max_pts);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    InitJob J;
    memset(&J, 0, sizeof(J));
    orbx_status st = stage_frame(m, F2, &J.F2, s);
    if (st) return st;
    m->arena_used = 0;
    const void *p;
    const size_t n1 = (size_t)F1->n;
    if ((st = arena_put(m, F1->keys_un, sizeof(orbx_keypoint) * n1, &p, s))) return st; J.keys1 = (const orbx_keypoint *)p;
    if ((st = arena_put(m, F1->desc, 32 * n1, &p, s))) return st; J.desc1 = (const uint8_t *)p;
    if ((st = arena_put(m, prev_xy, sizeof(float) * 2 * n1, &p, s))) return st; J.prev_xy = (const float *)p;
    J.n1 = F1->n; J.window = window_size; J.check_ori = check_ori; J.nnratio = nnratio;
    J.match12 = m->d_choice;                 // scratch of the claim resolution, unused by this kernel: >= max_pts ints
    J.matched_dist = m->d_minclaim;          // >= 2 * max_kp ints
    J.match21 = m->d_minclaim + m->max_kp;
    J.nmatches = m->d_nm;
    const void *djob;
    if ((st = arena_put(m, &J, sizeof(J), &djob, s))) return st;
    k_match_init<<<1, M_THREADS, m->smem, s>>>((const InitJob *)djob, m->max_kp);
    m->last_launches = 1;
    ORBX_CUDA(cudaGetLastError());
    if (n1) ORBX_CUDA(cudaMemcpyAsync(match12, m->d_choice, sizeof(int32_t) * n1, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(nmatches, m->d_nm, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    return ORBX_OK;
}
