// Batched many-sequence mode (BASELINE.json config 5): n independent sequences advance in lockstep, one new frame
// (or rectified pair) per sequence per step, through the per-frame chain Tracking runs (Tracking.cc:857-880 with a stereo / RGB-D
// last frame, Tracking.cc:900-965):
//     ORBextractor::operator() on the new image(s)                       -> orbx_extractor_run_device
//     Frame::ComputeStereoMatches (stereo only)                           -> orbx_stereo_matches_device
//     ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono)  -> orbx_match_projection_frame_device
// where the last frame's map points are its own keypoints unprojected with their depth, Frame::UnprojectStereo (Frame.cc:695-709),
// as Tracking::UpdateLastFrame creates them (Tracking.cc:941-947).  Everything between the upload of the new images and the
// download of the step's results stays on the device: the previous frame's keypoints, descriptors and depths are where the
// previous step left them.  One call = host buffers in, host buffers out (the C-ABI form of the step that bench.py's end-to-end
// figure goes through); sequences are independent, so several handles (on one GPU or on several) shard them with no exchange.
#include "orbx_internal.cuh"

#define SEQ_SLOTS 4
#define SEQ_MAX_SUBS 8

// The sequences of a handle are cut into sub-batches that run the per-step chain on streams of their own and meet again at the end
// of the step: the under-filled kernels of one sub-batch (the quadtree kernel has one CTA per frame and level, the matcher one small
// cluster per frame) and every kernel's tail run under the other sub-batches' kernels, a sub-batch's pyramid is still in L2 when its
// FAST / blur / descriptor kernels read it, and with host buffers the upload of one sub-batch overlaps the kernels of another.
struct SeqSub {
    orbx_extractor *ex;
    orbx_matcher *mt;
    orbx_stereo *st;
    orbx_pose *pz;
    cudaStream_t stream;
    cudaStream_t copy;         // keypoints / descriptors go back to the host here while stereo, search and pose optimisation still run
    cudaEvent_t done, ex_done, copy_done;
    int q0, nq;                // sequences [q0, q0 + nq)
};

struct orbx_sequences {
    orbx_sequences_config c;
    int n_img;                 // images per step: n_sequences x (stereo ? 2 : 1), order L0 R0 L1 R1 ...
    int cap;                   // keypoint slots per frame
    int n_sub;
    SeqSub sub[SEQ_MAX_SUBS];
    cudaStream_t stream;
    cudaEvent_t ev_fork;
    uint8_t *d_img;            // [n_img][height][width]
    // two generations of extractor outputs: cur = gen, last = gen ^ 1
    orbx_keypoint *d_kps[2];   // [n_img][cap]
    uint8_t *d_desc[2];        // [n_img][cap][32]
    int32_t *d_cnt[2];         // [n_img]
    float *d_ur, *d_depth[2];  // [n_seq][cap] mvuRight of the current frame; mvDepth per generation
    int32_t *d_kept;
    orbx_last_point *d_pts;    // [n_seq][cap] last frame's keypoints as map points
    int32_t *d_match, *d_nm;   // [n_seq][cap] followed by [n_seq] (one allocation, one memset)
    float *d_sf, *d_is2;       // mvScaleFactors, mvInvLevelSigma2
    double *d_pose;            // [n_seq][7] optimised SE3Quat (config pose)
    int32_t *d_inl;            // [n_seq]
    uint8_t *d_outkp;          // [n_seq][cap] mvbOutlier
    // per-step staging of the poses, [n_seq][24] floats (Tcw of the new frame, Tcw of the last): mapped pinned memory that the
    // preparation kernel reads directly; a ring, so that steps can be enqueued ahead of the device (a slot is rewritten only after
    // the kernel that reads it has run)
    float *h_pose_ring, *d_pose_ring;
    orbx_frame_match_job *d_jobs;   // [n_seq], filled by the preparation kernel
    float *h_pose_last;        // [n_seq][12] Tcw of the previous step
    int *h_status;             // pinned, [n_img]
    cudaEvent_t slot_ev[SEQ_SLOTS][SEQ_MAX_SUBS];
    int slot;
    int gen, steps, in_flight;
    int last_launches;
    // results for callers whose buffers are ordinary (pageable) memory: a device-to-host copy into pageable memory goes through the
    // driver's own staging and blocks the host once per copy (0.1 ms per step of one stereo frame, measured).  Such callers get the
    // step's results by DMA into this pinned block (allocated on first use) and a host copy of the valid rows in _step_end.
    uint8_t *h_stage;
    orbx_sequences_outputs stage;          // views into h_stage, same shapes as the caller's arrays
    orbx_sequences_outputs checked;        // the last caller struct whose pointers were classified ...
    int checked_valid, checked_pageable;   // ... and what they were
    orbx_sequences_outputs pending;        // caller buffers the staged step still has to be copied into
    int pending_valid, pending_pose;
};

// 1 if any of the caller's output buffers is not page-locked (cudaHostAlloc / cudaHostRegister) memory
static int outputs_pageable(orbx_sequences *h, const orbx_sequences_outputs *o) {
    if (h->checked_valid && !memcmp(&h->checked, o, sizeof(*o))) return h->checked_pageable;
    const void *ptrs[] = {o->kps, o->desc, o->counts, o->match, o->nmatches, o->u_right, o->depth, o->pose, o->n_inliers, o->outlier};
    int pageable = 0;
    for (const void *q : ptrs) {
        if (!q) continue;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, q) != cudaSuccess) { cudaGetLastError(); pageable = 1; break; }
        if (a.type != cudaMemoryTypeHost && a.type != cudaMemoryTypeManaged) { pageable = 1; break; }
    }
    h->checked = *o; h->checked_valid = 1; h->checked_pageable = pageable;
    return pageable;
}

static cudaError_t stage_alloc(orbx_sequences *h) {
    if (h->h_stage) return cudaSuccess;
    const size_t cap = (size_t)h->cap, ni = (size_t)h->n_img, ns = (size_t)h->c.n_sequences;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t at = off; off += (bytes + 63) & ~(size_t)63; return at; };
    const size_t o_kps = take(sizeof(orbx_keypoint) * cap * ni), o_desc = take(32 * cap * ni), o_cnt = take(4 * ni), o_match = take(4 * cap * ns),
                 o_nm = take(4 * ns), o_ur = take(4 * cap * ns), o_dp = take(4 * cap * ns), o_pose = take(56 * ns), o_inl = take(4 * ns),
                 o_out = take(cap * ns);
    cudaError_t e = cudaMallocHost((void **)&h->h_stage, off);
    if (e != cudaSuccess) return e;
    uint8_t *b = h->h_stage;
    h->stage.kps = reinterpret_cast<orbx_keypoint *>(b + o_kps); h->stage.desc = b + o_desc; h->stage.counts = reinterpret_cast<int32_t *>(b + o_cnt);
    h->stage.match = reinterpret_cast<int32_t *>(b + o_match); h->stage.nmatches = reinterpret_cast<int32_t *>(b + o_nm);
    h->stage.u_right = reinterpret_cast<float *>(b + o_ur); h->stage.depth = reinterpret_cast<float *>(b + o_dp);
    h->stage.pose = reinterpret_cast<double *>(b + o_pose); h->stage.n_inliers = reinterpret_cast<int32_t *>(b + o_inl); h->stage.outlier = b + o_out;
    return cudaSuccess;
}

// the staged step's results into the caller's arrays (after the streams have been synchronised): keypoints / descriptors up to each
// image's count, the per-sequence rows whole
static void unstage(orbx_sequences *h) {
    if (!h->pending_valid) return;
    h->pending_valid = 0;
    const orbx_sequences_outputs &o = h->pending, &t = h->stage;
    const size_t cap = (size_t)h->cap, ns = (size_t)h->c.n_sequences;
    for (int i = 0; i < h->n_img; i++) {
        const int32_t c = t.counts[i];
        const size_t n = c < 0 ? 0 : ((size_t)c > cap ? cap : (size_t)c);
        o.counts[i] = c;
        if (o.kps) memcpy(o.kps + cap * i, t.kps + cap * i, sizeof(orbx_keypoint) * n);
        if (o.desc) memcpy(o.desc + 32 * cap * i, t.desc + 32 * cap * i, 32 * n);
    }
    memcpy(o.match, t.match, 4 * cap * ns);
    memcpy(o.nmatches, t.nmatches, 4 * ns);
    if (h->c.stereo && o.u_right) memcpy(o.u_right, t.u_right, 4 * cap * ns);
    if (h->c.stereo && o.depth) memcpy(o.depth, t.depth, 4 * cap * ns);
    if (h->pending_pose) {
        if (o.pose) memcpy(o.pose, t.pose, 56 * ns);
        if (o.n_inliers) memcpy(o.n_inliers, t.n_inliers, 4 * ns);
        if (o.outlier) memcpy(o.outlier, t.outlier, cap * ns);
    }
}

// Everything a step needs before its kernels run, in one launch and without a copy: per sequence (blockIdx.y)
//   * the last frame's keypoints as map points, Frame::UnprojectStereo (Frame.cc:695-709): x = (u-cx)*z*invfx, y = (v-cy)*z*invfy,
//     X = Rwc*(x,y,z) + Ow with mRwc = mRcw.t(), mOw = -mRcw.t()*mtcw (Frame::UpdatePoseMatrices, Frame.cc:290-296), in float, each
//     product-sum left to right like cv::Mat's CV_32F gemm; keypoints without depth get no map point (valid = 0);
//   * CurrentFrame.mvpMapPoints all NULL (match = -1), nmatches = 0;
//   * the job of the projection search: the new frame's pose and bForward / bBackward (ORBmatcher.cc:1340-1351: twc = -Rcw.t()*tcw,
//     tlc = Rlw*twc + tlw).
// The poses are read straight from the mapped pinned staging slot of the step ([n_seq][24]: Tcw of the new frame, Tcw of the last).
struct SeqPrep {
    const orbx_keypoint *kps_last, *kps_cur;  const int32_t *cnt_last, *cnt_cur;  const uint8_t *desc_last, *desc_cur;
    const float *depth_last, *ur_cur, *sf;
    orbx_last_point *pts;  int32_t *match, *nm;  orbx_frame_match_job *jobs;
    int cap, per, nlevels, have_last, mono, check_ori;
    float const_depth, fx, fy, cx, cy, bf, b, width, height, th;
};
__global__ void k_prep_step(const SeqPrep P, const float *__restrict__ poses, int q0) {
    const int s = q0 + blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const float *T = poses + 24 * s, *Tl = T + 12;         // rows of [Rcw | tcw] of the new and of the last frame
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.nm[s] = 0;
        orbx_frame_match_job J;
        memset(&J, 0, sizeof(J));
        J.cur.n = 0;
        J.cur.n_dev = P.cnt_cur + P.per * s;
        J.cur.keys_un = P.kps_cur + (size_t)P.per * P.cap * s;
        J.cur.desc = P.desc_cur + (size_t)32 * P.per * P.cap * s;
        J.cur.u_right = P.ur_cur ? P.ur_cur + (size_t)P.cap * s : nullptr;
        J.cur.claimed = nullptr;
        J.cur.min_x = 0.f; J.cur.min_y = 0.f; J.cur.max_x = P.width; J.cur.max_y = P.height;   // undistorted input (Frame.cc:423-437)
        J.cur.grid_w_inv = __fdiv_rn(64.f, P.width);                                           // Frame.cc:127-128
        J.cur.grid_h_inv = __fdiv_rn(48.f, P.height);
        J.cur.fx = P.fx; J.cur.fy = P.fy; J.cur.cx = P.cx; J.cur.cy = P.cy; J.cur.bf = P.bf; J.cur.b = P.b;
        J.cur.scale_factors = P.sf;
        J.cur.nlevels = P.nlevels;
        J.n_last = P.have_last ? P.cnt_last[P.per * s] : 0;
        J.pts = P.pts + (size_t)P.cap * s;
        J.last_desc = P.desc_last + (size_t)32 * P.per * P.cap * s;
        for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) J.Rcw[3 * r + k] = T[4 * r + k]; J.tcw[r] = T[4 * r + 3]; }
        float twc[3];
        for (int r = 0; r < 3; r++) twc[r] = -__fadd_rn(__fadd_rn(__fmul_rn(T[r], T[3]), __fmul_rn(T[4 + r], T[7])), __fmul_rn(T[8 + r], T[11]));
        const float tlc2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tl[8], twc[0]), __fmul_rn(Tl[9], twc[1])), __fmul_rn(Tl[10], twc[2])), Tl[11]);
        J.forward = tlc2 > P.b && !P.mono;
        J.backward = -tlc2 > P.b && !P.mono;
        J.th = P.th;
        J.check_ori = P.check_ori;
        J.match = P.match + (size_t)P.cap * s;
        J.nmatches = P.nm + s;
        J.max_dist = 0;
        J.variant = 0;
        P.jobs[s] = J;
    }
    if (i >= P.cap) return;
    P.match[(size_t)s * P.cap + i] = -1;
    orbx_last_point p;
    p.x = p.y = p.z = p.angle = 0.f;
    p.octave = 0;
    p.valid = p.blocks = 0;
    p.pad[0] = p.pad[1] = 0;
    if (P.have_last && i < P.cnt_last[P.per * s]) {
        const orbx_keypoint k = P.kps_last[(size_t)s * P.per * P.cap + i];
        const float z = P.depth_last ? P.depth_last[(size_t)s * P.cap + i] : P.const_depth;
        if (z > 0) {
            const float x = __fmul_rn(__fmul_rn(__fsub_rn(k.x, P.cx), z), __fdiv_rn(1.0f, P.fx));
            const float y = __fmul_rn(__fmul_rn(__fsub_rn(k.y, P.cy), z), __fdiv_rn(1.0f, P.fy));
            float X[3];
            for (int r = 0; r < 3; r++) {
                const float Ow = -__fadd_rn(__fadd_rn(__fmul_rn(Tl[r], Tl[3]), __fmul_rn(Tl[4 + r], Tl[7])), __fmul_rn(Tl[8 + r], Tl[11]));
                X[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tl[r], x), __fmul_rn(Tl[4 + r], y)), __fmul_rn(Tl[8 + r], z)), Ow);   // row r of Rwc = column r of Rcw
            }
            p.x = X[0]; p.y = X[1]; p.z = X[2];
            p.angle = k.angle;
            p.octave = k.octave;
            p.valid = 1;
            p.blocks = 1;
        }
    }
    P.pts[(size_t)s * P.cap + i] = p;
}

extern "C" void orbx_sequences_destroy(orbx_sequences *h) {
    if (!h) return;
    cudaSetDevice(h->c.device);
    cudaDeviceSynchronize();
    for (int k = 0; k < h->n_sub; k++) {
        SeqSub &S = h->sub[k];
        if (S.ex) orbx_extractor_destroy(S.ex);
        if (S.mt) orbx_matcher_destroy(S.mt);
        if (S.st) orbx_stereo_destroy(S.st);
        if (S.pz) orbx_pose_destroy(S.pz);
        if (S.stream) cudaStreamDestroy(S.stream);
        if (S.copy) cudaStreamDestroy(S.copy);
        if (S.done) cudaEventDestroy(S.done);
        if (S.ex_done) cudaEventDestroy(S.ex_done);
        if (S.copy_done) cudaEventDestroy(S.copy_done);
    }
    cudaFree(h->d_img);
    for (int g = 0; g < 2; g++) { cudaFree(h->d_kps[g]); cudaFree(h->d_desc[g]); cudaFree(h->d_cnt[g]); cudaFree(h->d_depth[g]); }
    cudaFree(h->d_ur); cudaFree(h->d_kept); cudaFree(h->d_pts); cudaFree(h->d_match); cudaFree(h->d_sf); cudaFree(h->d_is2); cudaFree(h->d_pose); cudaFree(h->d_inl); cudaFree(h->d_outkp); cudaFree(h->d_jobs);
    cudaFreeHost(h->h_pose_ring); cudaFreeHost(h->h_status); if (h->h_stage) cudaFreeHost(h->h_stage);
    for (int i = 0; i < SEQ_SLOTS; i++) for (int k = 0; k < SEQ_MAX_SUBS; k++) if (h->slot_ev[i][k]) cudaEventDestroy(h->slot_ev[i][k]);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    free(h->h_pose_last);
    if (h->stream) cudaStreamDestroy(h->stream);
    free(h);
}

extern "C" orbx_status orbx_sequences_create(orbx_sequences **out, const orbx_sequences_config *cfg) {
    if (!out || !cfg || cfg->n_sequences <= 0 || cfg->width <= 0 || cfg->height <= 0) return ORBX_ERR_INVALID;
    if (!cfg->stereo && !(cfg->const_depth > 0)) {
        orbx_set_error("orbx_sequences_create: monocular input needs const_depth > 0 (the last frame's keypoints get their depth from it)");
        return ORBX_ERR_INVALID;
    }
    orbx_sequences *h = (orbx_sequences *)calloc(1, sizeof(orbx_sequences));
    if (!h) return ORBX_ERR_NOMEM;
    h->c = *cfg;
    const int ns = cfg->n_sequences, per = cfg->stereo ? 2 : 1;
    h->n_img = ns * per;
    // one stream unless asked otherwise (config n_sub > 0 or the environment variable ORBX_SEQ_SUBS).  A first version joined the
    // sub-batches at the end of every step: on the B200 at 64 VGA sequences that cost more than the overlap returned (91.8 k frames/s
    // with 1, 89.8 k with 2, 90.3 k with 4, 85.9 k with 8 sub-batches; profiles/r2_f_subbatch_sweep.txt), so they now run unjoined
    int n_sub = cfg->n_sub > 0 ? cfg->n_sub : (getenv("ORBX_SEQ_SUBS") ? atoi(getenv("ORBX_SEQ_SUBS")) : 1);
    n_sub = n_sub < 1 ? 1 : (n_sub > SEQ_MAX_SUBS ? SEQ_MAX_SUBS : n_sub);
    if (n_sub > ns) n_sub = ns;
    h->n_sub = n_sub;
    orbx_status st = ORBX_OK;
#define TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { orbx_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); orbx_sequences_destroy(h); return ORBX_ERR_CUDA; } } while (0)
    TRY(cudaSetDevice(cfg->device));
    for (int k = 0; k < n_sub; k++) {
        SeqSub &S = h->sub[k];
        S.q0 = (int)((long long)ns * k / n_sub);
        S.nq = (int)((long long)ns * (k + 1) / n_sub) - S.q0;
        if ((st = orbx_extractor_create(&S.ex, cfg->nfeatures, cfg->scale_factor, cfg->nlevels, cfg->ini_th, cfg->min_th, cfg->width, cfg->height,
                                        S.nq * per, cfg->device))) { orbx_sequences_destroy(h); return st; }
        h->cap = orbx_extractor_capacity(S.ex);
        if ((st = orbx_matcher_create(&S.mt, h->cap, h->cap, S.nq, cfg->device))) { orbx_sequences_destroy(h); return st; }
        if (cfg->stereo && (st = orbx_stereo_create(&S.st, h->cap, S.nq, cfg->device))) { orbx_sequences_destroy(h); return st; }
        if (cfg->pose && (st = orbx_pose_create(&S.pz, h->cap * S.nq, S.nq, cfg->device))) { orbx_sequences_destroy(h); return st; }
        TRY(cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking));
        TRY(cudaStreamCreateWithFlags(&S.copy, cudaStreamNonBlocking));
        TRY(cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming));
        TRY(cudaEventCreateWithFlags(&S.ex_done, cudaEventDisableTiming));
        TRY(cudaEventCreateWithFlags(&S.copy_done, cudaEventDisableTiming));
    }
    TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    const size_t cap = (size_t)h->cap;
    TRY(cudaMalloc(&h->d_img, (size_t)h->n_img * cfg->width * cfg->height));
    for (int g = 0; g < 2; g++) {
        TRY(cudaMalloc(&h->d_kps[g], sizeof(orbx_keypoint) * cap * h->n_img));
        TRY(cudaMalloc(&h->d_desc[g], 32 * cap * h->n_img));
        TRY(cudaMalloc(&h->d_cnt[g], sizeof(int32_t) * h->n_img));
        TRY(cudaMemset(h->d_cnt[g], 0, sizeof(int32_t) * h->n_img));
        TRY(cudaMalloc(&h->d_depth[g], sizeof(float) * cap * ns));
    }
    TRY(cudaMalloc(&h->d_ur, sizeof(float) * cap * ns));
    TRY(cudaMalloc(&h->d_kept, sizeof(int32_t) * ns));
    TRY(cudaMalloc(&h->d_pts, sizeof(orbx_last_point) * cap * ns));
    TRY(cudaMalloc(&h->d_match, sizeof(int32_t) * (cap + 1) * ns));
    h->d_nm = h->d_match + cap * ns;
    TRY(cudaMalloc(&h->d_sf, sizeof(float) * ORBX_MAX_LEVELS));
    TRY(cudaMalloc(&h->d_is2, sizeof(float) * ORBX_MAX_LEVELS));
    if (cfg->pose) {
        TRY(cudaMalloc(&h->d_pose, sizeof(double) * 7 * ns));
        TRY(cudaMalloc(&h->d_inl, sizeof(int32_t) * ns));
        TRY(cudaMalloc(&h->d_outkp, cap * ns));
    }
    TRY(cudaMalloc(&h->d_jobs, sizeof(orbx_frame_match_job) * ns));
    TRY(cudaHostAlloc(&h->h_pose_ring, sizeof(float) * 24 * ns * SEQ_SLOTS, cudaHostAllocMapped));
    TRY(cudaHostGetDevicePointer(&h->d_pose_ring, h->h_pose_ring, 0));
    TRY(cudaMallocHost(&h->h_status, sizeof(int) * h->n_img));
    for (int i = 0; i < SEQ_SLOTS; i++) for (int k = 0; k < n_sub; k++) TRY(cudaEventCreateWithFlags(&h->slot_ev[i][k], cudaEventDisableTiming));
    h->h_pose_last = (float *)calloc((size_t)12 * ns, sizeof(float));
    if (!h->h_pose_last) { orbx_sequences_destroy(h); return ORBX_ERR_NOMEM; }
    float sf[ORBX_MAX_LEVELS] = {0}, is2[ORBX_MAX_LEVELS] = {0};
    if ((st = orbx_extractor_tables(h->sub[0].ex, sf, nullptr, nullptr, is2, nullptr))) { orbx_sequences_destroy(h); return st; }
    TRY(cudaMemcpy(h->d_sf, sf, sizeof(sf), cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(h->d_is2, is2, sizeof(is2), cudaMemcpyHostToDevice));
#undef TRY
    *out = h;
    return ORBX_OK;
}

extern "C" int orbx_sequences_capacity(const orbx_sequences *h) { return h ? h->cap : 0; }
extern "C" int orbx_sequences_last_launches(const orbx_sequences *h) { return h ? h->last_launches : 0; }

extern "C" orbx_status orbx_sequences_reset(orbx_sequences *h) {
    if (!h) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->c.device));
    ORBX_CUDA(cudaDeviceSynchronize());
    h->steps = 0;
    h->in_flight = 0;
    return ORBX_OK;
}

// One step on stream s.  images: device pointer (host_images = false) or host pointer (true: every sub-batch uploads its own images
// on its own stream); o: host buffers every sub-batch downloads its results into (NULL: results stay on the device).
static orbx_status run_step(orbx_sequences *h, const uint8_t *images, bool host_images, size_t image_pitch, int stride, const float *Tcw,
                            const orbx_sequences_outputs *o, cudaStream_t s) {
    const orbx_sequences_config &c = h->c;
    const int ns = c.n_sequences, per = c.stereo ? 2 : 1, cap = h->cap;
    const int g = h->gen, gl = g ^ 1;
    const float b = c.bf / c.fx;
    h->last_launches = 0;
    const int slot = h->slot;
    h->slot = (slot + 1) % SEQ_SLOTS;
    for (int k = 0; k < h->n_sub; k++) ORBX_CUDA(cudaEventSynchronize(h->slot_ev[slot][k]));   // the kernels of the step that used this staging slot have run
    float *h_pose = h->h_pose_ring + (size_t)24 * ns * slot;
    for (int q = 0; q < ns; q++) {
        memcpy(h_pose + 24 * q, Tcw + 12 * q, sizeof(float) * 12);
        memcpy(h_pose + 24 * q + 12, h->h_pose_last + 12 * q, sizeof(float) * 12);
    }
    const bool have_last = h->steps > 0;
    orbx_sequences_outputs tgt;                          // where the copies of this step land: the caller's arrays, or the pinned block
    if (o) {
        tgt = *o;
        if (outputs_pageable(h, o)) {
            ORBX_CUDA(stage_alloc(h));
            const orbx_sequences_outputs &t = h->stage;
            tgt.counts = t.counts; tgt.match = t.match; tgt.nmatches = t.nmatches;
            if (o->kps) tgt.kps = t.kps;
            if (o->desc) tgt.desc = t.desc;
            if (o->u_right) tgt.u_right = t.u_right;
            if (o->depth) tgt.depth = t.depth;
            if (o->pose) tgt.pose = t.pose;
            if (o->n_inliers) tgt.n_inliers = t.n_inliers;
            if (o->outlier) tgt.outlier = t.outlier;
            h->pending = *o; h->pending_valid = 1; h->pending_pose = c.pose && have_last;
        }
        o = &tgt;
    }
    SeqPrep P;
    P.kps_last = h->d_kps[gl]; P.kps_cur = h->d_kps[g]; P.cnt_last = h->d_cnt[gl]; P.cnt_cur = h->d_cnt[g];
    P.desc_last = h->d_desc[gl]; P.desc_cur = h->d_desc[g];
    P.depth_last = c.stereo ? h->d_depth[gl] : nullptr; P.ur_cur = c.stereo ? h->d_ur : nullptr; P.sf = h->d_sf;
    P.pts = h->d_pts; P.match = h->d_match; P.nm = h->d_nm; P.jobs = h->d_jobs;
    P.cap = cap; P.per = per; P.nlevels = c.nlevels; P.have_last = have_last; P.mono = c.mono; P.check_ori = c.check_ori;
    P.const_depth = c.const_depth; P.fx = c.fx; P.fy = c.fy; P.cx = c.cx; P.cy = c.cy; P.bf = c.bf; P.b = b;
    P.width = (float)c.width; P.height = (float)c.height; P.th = c.th;
    orbx_frame_match_job *d_jobs = h->d_jobs;
    // n_sub > 1: the sub-batches are independent pipelines on streams of their own.  Each waits for what the caller's stream holds
    // at this point (the new images), then runs its whole step; nothing joins them until orbx_sequences_join / _step_end, so a
    // sub-batch's step i + 1 starts as soon as its own step i is done.
    const bool fork = h->n_sub > 1;
    if (fork) ORBX_CUDA(cudaEventRecord(h->ev_fork, s));
    const size_t img_bytes = (size_t)c.width * c.height;
    for (int k = 0; k < h->n_sub; k++) {
        SeqSub &S = h->sub[k];
        cudaStream_t ss = fork ? S.stream : s;
        if (fork) ORBX_CUDA(cudaStreamWaitEvent(ss, h->ev_fork, 0));
        {   // preparation: last frame's map points, match reset, jobs (reads the mapped pinned poses of this step's slot)
            dim3 grid((cap + 255) / 256, S.nq);
            k_prep_step<<<grid, 256, 0, ss>>>(P, h->d_pose_ring + (size_t)24 * ns * slot, S.q0);
            ORBX_CUDA(cudaGetLastError());
            h->last_launches++;
            ORBX_CUDA(cudaEventRecord(h->slot_ev[slot][k], ss));
        }
        const int i0 = per * S.q0, ni = per * S.nq;
        const uint8_t *d_in = images + (size_t)i0 * image_pitch;
        size_t in_pitch = image_pitch;
        int in_stride = stride;
        if (host_images) {                                   // this sub-batch's images: rows `stride` apart on the host, packed on the device
            uint8_t *dst = h->d_img + img_bytes * i0;
            if (image_pitch == (size_t)stride * c.height)
                ORBX_CUDA(cudaMemcpy2DAsync(dst, c.width, d_in, stride, c.width, (size_t)c.height * ni, cudaMemcpyHostToDevice, ss));
            else
                for (int i = 0; i < ni; i++)
                    ORBX_CUDA(cudaMemcpy2DAsync(dst + img_bytes * i, c.width, d_in + i * image_pitch, stride, c.width, c.height, cudaMemcpyHostToDevice, ss));
            d_in = dst; in_pitch = img_bytes; in_stride = c.width;
        }
        // ORBextractor::operator() on every new image of the sub-batch
        orbx_status st = orbx_extractor_run_device(S.ex, d_in, in_pitch, ni, c.width, c.height, in_stride, h->d_kps[g] + (size_t)cap * i0,
                                                   h->d_desc[g] + (size_t)32 * cap * i0, h->d_cnt[g] + i0, ss);
        if (st) return st;
        h->last_launches += orbx_extractor_last_launches(S.ex);
        if (o) {     // the extractor's results start their way back now, on the copy stream; the rest of the step does not wait for them
            ORBX_CUDA(cudaEventRecord(S.ex_done, ss));
            ORBX_CUDA(cudaStreamWaitEvent(S.copy, S.ex_done, 0));
            ORBX_CUDA(cudaMemcpyAsync(o->counts + i0, h->d_cnt[g] + i0, sizeof(int32_t) * ni, cudaMemcpyDeviceToHost, S.copy));
            if (o->kps) ORBX_CUDA(cudaMemcpyAsync(o->kps + (size_t)cap * i0, h->d_kps[g] + (size_t)cap * i0, sizeof(orbx_keypoint) * cap * ni, cudaMemcpyDeviceToHost, S.copy));
            if (o->desc) ORBX_CUDA(cudaMemcpyAsync(o->desc + (size_t)32 * cap * i0, h->d_desc[g] + (size_t)32 * cap * i0, (size_t)32 * cap * ni, cudaMemcpyDeviceToHost, S.copy));
            ORBX_CUDA(cudaMemcpyAsync(h->h_status + i0, S.ex->d_status, sizeof(int) * ni, cudaMemcpyDeviceToHost, S.copy));
            ORBX_CUDA(cudaEventRecord(S.copy_done, S.copy));
        }
        // Frame::ComputeStereoMatches
        if (c.stereo) {
            orbx_stereo_side L = {h->d_kps[g] + (size_t)cap * i0, h->d_desc[g] + (size_t)32 * cap * i0, h->d_cnt[g] + i0, 2 * cap, 2, S.ex, 0, 2, cap};
            orbx_stereo_side R = {L.keys + cap, L.desc + (size_t)32 * cap, L.counts + 1, 2 * cap, 2, S.ex, 1, 2, cap};
            if ((st = orbx_stereo_matches_device(S.st, &L, &R, S.nq, c.bf, b, h->d_ur + (size_t)cap * S.q0, h->d_depth[g] + (size_t)cap * S.q0, cap,
                                                 h->d_kept + S.q0, ss))) return st;
            h->last_launches += orbx_stereo_last_launches(S.st);
        }
        // SearchByProjection(CurrentFrame, LastFrame, th, bMono)
        if (have_last) {
            if ((st = orbx_match_projection_frame_device(S.mt, d_jobs + S.q0, S.nq, ss))) return st;
            h->last_launches += orbx_matcher_last_launches(S.mt);
        }
        // Optimizer::PoseOptimization from the match array, starting at the pose the search projected with (Tracking.cc:868-870)
        if (c.pose && have_last) {
            if ((st = orbx_pose_from_matches_device(S.pz, d_jobs + S.q0, S.nq, h->d_is2, c.nlevels, c.fx, c.fy, c.cx, c.cy, c.bf, h->d_pose + 7 * S.q0,
                                                    h->d_inl + S.q0, h->d_outkp + (size_t)cap * S.q0, cap, ss))) return st;
            h->last_launches += orbx_pose_last_launches(S.pz);
        }
        if (o) {                                             // the sub-batch's results back to the caller's buffers
            ORBX_CUDA(cudaMemcpyAsync(o->match + (size_t)cap * S.q0, h->d_match + (size_t)cap * S.q0, sizeof(int32_t) * cap * S.nq, cudaMemcpyDeviceToHost, ss));
            ORBX_CUDA(cudaMemcpyAsync(o->nmatches + S.q0, h->d_nm + S.q0, sizeof(int32_t) * S.nq, cudaMemcpyDeviceToHost, ss));
            if (c.stereo && o->u_right) ORBX_CUDA(cudaMemcpyAsync(o->u_right + (size_t)cap * S.q0, h->d_ur + (size_t)cap * S.q0, sizeof(float) * cap * S.nq, cudaMemcpyDeviceToHost, ss));
            if (c.stereo && o->depth) ORBX_CUDA(cudaMemcpyAsync(o->depth + (size_t)cap * S.q0, h->d_depth[g] + (size_t)cap * S.q0, sizeof(float) * cap * S.nq, cudaMemcpyDeviceToHost, ss));
            if (c.pose && have_last) {
                if (o->pose) ORBX_CUDA(cudaMemcpyAsync(o->pose + 7 * S.q0, h->d_pose + 7 * S.q0, sizeof(double) * 7 * S.nq, cudaMemcpyDeviceToHost, ss));
                if (o->n_inliers) ORBX_CUDA(cudaMemcpyAsync(o->n_inliers + S.q0, h->d_inl + S.q0, sizeof(int32_t) * S.nq, cudaMemcpyDeviceToHost, ss));
                if (o->outlier) ORBX_CUDA(cudaMemcpyAsync(o->outlier + (size_t)cap * S.q0, h->d_outkp + (size_t)cap * S.q0, (size_t)cap * S.nq, cudaMemcpyDeviceToHost, ss));
            }
            ORBX_CUDA(cudaStreamWaitEvent(ss, S.copy_done, 0));               // whoever waits for this stream waits for the copy stream too
        }
        if (fork) ORBX_CUDA(cudaEventRecord(S.done, ss));
    }
    memcpy(h->h_pose_last, Tcw, sizeof(float) * 12 * ns);
    h->gen ^= 1;
    h->steps++;
    return ORBX_OK;
}

extern "C" orbx_status orbx_sequences_step_begin(orbx_sequences *h, const uint8_t *images, size_t image_pitch, int stride,
                                                 const float *Tcw, const orbx_sequences_outputs *o) {
    if (!h || !images || !Tcw || !o || !o->counts || !o->nmatches || !o->match) return ORBX_ERR_INVALID;
    if (stride < h->c.width) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->c.device));
    if (h->in_flight) {                                                // one step in flight per handle
        ORBX_CUDA(cudaStreamSynchronize(h->stream));
        if (h->n_sub > 1) for (int k = 0; k < h->n_sub; k++) ORBX_CUDA(cudaStreamSynchronize(h->sub[k].stream));
        unstage(h);
    }
    const orbx_status st = run_step(h, images, true, image_pitch, stride, Tcw, o, h->stream);
    if (st) return st;
    h->in_flight = 1;
    return ORBX_OK;
}

extern "C" orbx_status orbx_sequences_step_device(orbx_sequences *h, const uint8_t *d_images, size_t frame_pitch, int stride,
                                                  const float *Tcw, void *stream) {
    if (!h || !d_images || !Tcw || stride < h->c.width) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->c.device));
    return run_step(h, d_images, false, frame_pitch, stride, Tcw, nullptr, (cudaStream_t)stream);
}

extern "C" orbx_status orbx_sequences_set_last_poses(orbx_sequences *h, const float *Tcw_last) {
    if (!h || !Tcw_last) return ORBX_ERR_INVALID;
    memcpy(h->h_pose_last, Tcw_last, sizeof(float) * 12 * h->c.n_sequences);
    return ORBX_OK;
}

extern "C" orbx_status orbx_sequences_join(orbx_sequences *h, void *stream) {
    if (!h) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->c.device));
    if (h->n_sub > 1 && h->steps > 0)
        for (int k = 0; k < h->n_sub; k++) ORBX_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, h->sub[k].done, 0));
    return ORBX_OK;
}

extern "C" orbx_status orbx_sequences_device_view(const orbx_sequences *h, orbx_sequences_device *v) {
    if (!h || !v) return ORBX_ERR_INVALID;
    const int g = h->gen ^ 1;                                   // the generation the last step filled
    v->extractor = h->sub[0].ex;
    v->kps = h->d_kps[g]; v->desc = h->d_desc[g]; v->counts = h->d_cnt[g];
    v->match = h->d_match; v->nmatches = h->d_nm; v->u_right = h->d_ur; v->depth = h->d_depth[g];
    v->jobs = h->d_jobs;
    v->pose = h->d_pose; v->n_inliers = h->d_inl; v->outlier = h->d_outkp;
    v->stream = h->stream;
    return ORBX_OK;
}

extern "C" orbx_status orbx_sequences_step_end(orbx_sequences *h) {
    if (!h) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->c.device));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    if (h->n_sub > 1) for (int k = 0; k < h->n_sub; k++) ORBX_CUDA(cudaStreamSynchronize(h->sub[k].stream));
    unstage(h);
    if (h->in_flight) {
        h->in_flight = 0;
        for (int i = 0; i < h->n_img; i++)
            if (h->h_status[i]) {
                orbx_set_error("image %d: device status 0x%x (1 = candidate list overflow, 2 = quadtree depth, 4 = node overflow)", i, h->h_status[i]);
                return ORBX_ERR_CAPACITY;
            }
    }
    return ORBX_OK;
}

extern "C" orbx_status orbx_sequences_step_host(orbx_sequences *h, const uint8_t *images, size_t image_pitch, int stride, const float *Tcw,
                                                const orbx_sequences_outputs *o) {
    const orbx_status st = orbx_sequences_step_begin(h, images, image_pitch, stride, Tcw, o);
    return st ? st : orbx_sequences_step_end(h);
}
