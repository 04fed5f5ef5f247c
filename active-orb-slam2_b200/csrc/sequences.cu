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

struct orbx_sequences {
    orbx_sequences_config c;
    int n_img;                 // images per step: n_sequences x (stereo ? 2 : 1), order L0 R0 L1 R1 ...
    int cap;                   // keypoint slots per frame
    orbx_extractor *ex;
    orbx_matcher *mt;
    orbx_stereo *st;
    cudaStream_t stream;
    uint8_t *d_img;            // [n_img][height][width]
    // two generations of extractor outputs: cur = gen, last = gen ^ 1
    orbx_keypoint *d_kps[2];   // [n_img][cap]
    uint8_t *d_desc[2];        // [n_img][cap][32]
    int32_t *d_cnt[2];         // [n_img]
    float *d_ur, *d_depth[2];  // [n_seq][cap] mvuRight of the current frame; mvDepth per generation
    int32_t *d_kept;
    orbx_last_point *d_pts;    // [n_seq][cap] last frame's keypoints as map points
    int32_t *d_match, *d_nm;   // [n_seq][cap], [n_seq]
    float *d_sf;               // mvScaleFactors
    // pinned staging of the per-step job array and last-frame poses: a ring, so that steps can be enqueued ahead of the device
    // (a slot is rewritten only after the copies that read it have run)
    orbx_frame_match_job *d_jobs, *h_jobs_ring;   // [n_seq], [SEQ_SLOTS][n_seq]
    float *h_pose_last;        // [n_seq][12] Tcw of the previous step
    float *d_Twc, *h_Twc_ring; // [n_seq][12] Rwc | Ow of the last frame, for the unprojection; [SEQ_SLOTS][n_seq][12]
    cudaEvent_t slot_ev[4];
    int slot;
    int gen, steps, in_flight;
    int last_launches;
};

// Frame::UnprojectStereo for every keypoint of the last frame of every sequence (Frame.cc:695-709): x = (u-cx)*z*invfx,
// y = (v-cy)*z*invfy, X = Rwc*(x,y,z) + Ow in float, each product-sum left to right like cv::Mat's CV_32F gemm.  Keypoints
// without depth get no map point (valid = 0).  One thread per keypoint slot.
__global__ void k_unproject_last(const orbx_keypoint *__restrict__ kps, const int32_t *__restrict__ cnt, int kp_pitch, int cnt_step,
                                 const float *__restrict__ depth, int depth_pitch, float const_depth, const float *__restrict__ Twc,
                                 float cx, float cy, float invfx, float invfy, orbx_last_point *__restrict__ out, int cap) {
    const int s = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    orbx_last_point p;
    p.x = p.y = p.z = p.angle = 0.f;
    p.octave = 0;
    p.valid = p.blocks = 0;
    p.pad[0] = p.pad[1] = 0;
    if (i < cnt[s * cnt_step]) {
        const orbx_keypoint k = kps[(size_t)s * kp_pitch + i];
        const float z = depth ? depth[(size_t)s * depth_pitch + i] : const_depth;
        if (z > 0) {
            const float x = __fmul_rn(__fmul_rn(__fsub_rn(k.x, cx), z), invfx);
            const float y = __fmul_rn(__fmul_rn(__fsub_rn(k.y, cy), z), invfy);
            const float *T = Twc + 12 * s;
            p.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[0], x), __fmul_rn(T[1], y)), __fmul_rn(T[2], z)), T[9]);
            p.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[3], x), __fmul_rn(T[4], y)), __fmul_rn(T[5], z)), T[10]);
            p.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T[6], x), __fmul_rn(T[7], y)), __fmul_rn(T[8], z)), T[11]);
            p.angle = k.angle;
            p.octave = k.octave;
            p.valid = 1;
            p.blocks = 1;
        }
    }
    out[(size_t)s * cap + i] = p;
}

extern "C" void orbx_sequences_destroy(orbx_sequences *h) {
    if (!h) return;
    cudaSetDevice(h->c.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->ex) orbx_extractor_destroy(h->ex);
    if (h->mt) orbx_matcher_destroy(h->mt);
    if (h->st) orbx_stereo_destroy(h->st);
    cudaFree(h->d_img);
    for (int g = 0; g < 2; g++) { cudaFree(h->d_kps[g]); cudaFree(h->d_desc[g]); cudaFree(h->d_cnt[g]); cudaFree(h->d_depth[g]); }
    cudaFree(h->d_ur); cudaFree(h->d_kept); cudaFree(h->d_pts); cudaFree(h->d_match); cudaFree(h->d_nm); cudaFree(h->d_sf);
    cudaFree(h->d_jobs); cudaFree(h->d_Twc);
    cudaFreeHost(h->h_jobs_ring); cudaFreeHost(h->h_Twc_ring);
    for (int i = 0; i < SEQ_SLOTS; i++) if (h->slot_ev[i]) cudaEventDestroy(h->slot_ev[i]);
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
    const int ns = cfg->n_sequences;
    h->n_img = ns * (cfg->stereo ? 2 : 1);
    orbx_status st = orbx_extractor_create(&h->ex, cfg->nfeatures, cfg->scale_factor, cfg->nlevels, cfg->ini_th, cfg->min_th, cfg->width,
                                           cfg->height, h->n_img, cfg->device);
    if (st) { orbx_sequences_destroy(h); return st; }
    h->cap = orbx_extractor_capacity(h->ex);
    if ((st = orbx_matcher_create(&h->mt, h->cap, h->cap, ns, cfg->device))) { orbx_sequences_destroy(h); return st; }
    if (cfg->stereo && (st = orbx_stereo_create(&h->st, h->cap, ns, cfg->device))) { orbx_sequences_destroy(h); return st; }
#define TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { orbx_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); orbx_sequences_destroy(h); return ORBX_ERR_CUDA; } } while (0)
    TRY(cudaSetDevice(cfg->device));
    TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
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
    TRY(cudaMalloc(&h->d_match, sizeof(int32_t) * cap * ns));
    TRY(cudaMalloc(&h->d_nm, sizeof(int32_t) * ns));
    TRY(cudaMalloc(&h->d_sf, sizeof(float) * ORBX_MAX_LEVELS));
    TRY(cudaMalloc(&h->d_jobs, sizeof(orbx_frame_match_job) * ns));
    TRY(cudaMalloc(&h->d_Twc, sizeof(float) * 12 * ns));
    TRY(cudaMallocHost(&h->h_jobs_ring, sizeof(orbx_frame_match_job) * ns * SEQ_SLOTS));
    TRY(cudaMallocHost(&h->h_Twc_ring, sizeof(float) * 12 * ns * SEQ_SLOTS));
    for (int i = 0; i < SEQ_SLOTS; i++) TRY(cudaEventCreateWithFlags(&h->slot_ev[i], cudaEventDisableTiming));
    h->h_pose_last = (float *)calloc((size_t)12 * ns, sizeof(float));
    if (!h->h_pose_last) { orbx_sequences_destroy(h); return ORBX_ERR_NOMEM; }
    float sf[ORBX_MAX_LEVELS] = {0};
    if ((st = orbx_extractor_tables(h->ex, sf, nullptr, nullptr, nullptr, nullptr))) { orbx_sequences_destroy(h); return st; }
    TRY(cudaMemcpy(h->d_sf, sf, sizeof(sf), cudaMemcpyHostToDevice));
#undef TRY
    *out = h;
    return ORBX_OK;
}

extern "C" int orbx_sequences_capacity(const orbx_sequences *h) { return h ? h->cap : 0; }
extern "C" int orbx_sequences_last_launches(const orbx_sequences *h) { return h ? h->last_launches : 0; }

extern "C" orbx_status orbx_sequences_reset(orbx_sequences *h) {
    if (!h) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->c.device));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    h->steps = 0;
    return ORBX_OK;
}

// the device part of one step on stream s: unprojection of the last frame, extraction, stereo association, projection search
static orbx_status step_core(orbx_sequences *h, const uint8_t *d_images, size_t frame_pitch, int stride, const float *Tcw, cudaStream_t s) {
    const orbx_sequences_config &c = h->c;
    const int ns = c.n_sequences, per = c.stereo ? 2 : 1, cap = h->cap;
    const int g = h->gen, gl = g ^ 1;
    h->last_launches = 0;
    const int slot = h->slot;
    h->slot = (slot + 1) % SEQ_SLOTS;
    ORBX_CUDA(cudaEventSynchronize(h->slot_ev[slot]));          // the copies of the step that used this staging slot have run
    orbx_frame_match_job *h_jobs = h->h_jobs_ring + (size_t)slot * ns;
    float *h_Twc = h->h_Twc_ring + (size_t)12 * ns * slot;
    // ---- the last frame's map points: Frame::UnprojectStereo with the last frame's pose (Frame.cc:290-296, 695-709) ----
    const bool have_last = h->steps > 0;
    if (have_last) {
        for (int q = 0; q < ns; q++) {
            const float *T = h->h_pose_last + 12 * q;      // rows of [Rcw | tcw]
            float *W = h_Twc + 12 * q;                  // Rwc (row-major 3x3), then Ow
            for (int r = 0; r < 3; r++)
                for (int k = 0; k < 3; k++) W[3 * r + k] = T[4 * k + r];                          // mRwc = mRcw.t()
            for (int r = 0; r < 3; r++)                                                            // mOw = -mRcw.t()*mtcw
                W[9 + r] = -((T[r] * T[3] + T[4 + r] * T[7]) + T[8 + r] * T[11]);
        }
        ORBX_CUDA(cudaMemcpyAsync(h->d_Twc, h_Twc, sizeof(float) * 12 * ns, cudaMemcpyHostToDevice, s));
        dim3 grid((cap + 255) / 256, ns);
        k_unproject_last<<<grid, 256, 0, s>>>(h->d_kps[gl], h->d_cnt[gl], per * cap, per, c.stereo ? h->d_depth[gl] : nullptr, cap,
                                              c.const_depth, h->d_Twc, c.cx, c.cy, 1.0f / c.fx, 1.0f / c.fy, h->d_pts, cap);
        ORBX_CUDA(cudaGetLastError());
        h->last_launches++;
    }
    // ---- ORBextractor::operator() on every new image ------------------------------------------------------------------------
    orbx_status st = orbx_extractor_run_device(h->ex, d_images, frame_pitch, h->n_img, c.width, c.height, stride,
                                               h->d_kps[g], h->d_desc[g], h->d_cnt[g], s);
    if (st) return st;
    h->last_launches += orbx_extractor_last_launches(h->ex);
    // ---- Frame::ComputeStereoMatches ------------------------------------------------------------------------------------------
    const float b = c.bf / c.fx;
    if (c.stereo) {
        orbx_stereo_side L = {h->d_kps[g], h->d_desc[g], h->d_cnt[g], 2 * cap, 2, h->ex, 0, 2, cap};
        orbx_stereo_side R = {h->d_kps[g] + cap, h->d_desc[g] + (size_t)32 * cap, h->d_cnt[g] + 1, 2 * cap, 2, h->ex, 1, 2, cap};
        if ((st = orbx_stereo_matches_device(h->st, &L, &R, ns, c.bf, b, h->d_ur, h->d_depth[g], cap, h->d_kept, s))) return st;
        h->last_launches += orbx_stereo_last_launches(h->st);
    }
    // ---- SearchByProjection(CurrentFrame, LastFrame, th, bMono) -------------------------------------------------------------
    ORBX_CUDA(cudaMemsetAsync(h->d_match, 0xff, sizeof(int32_t) * cap * ns, s));     // mvpMapPoints all NULL
    ORBX_CUDA(cudaMemsetAsync(h->d_nm, 0, sizeof(int32_t) * ns, s));
    if (have_last) {
        for (int q = 0; q < ns; q++) {
            orbx_frame_match_job &J = h_jobs[q];
            memset(&J, 0, sizeof(J));
            J.cur.n = 0;
            J.cur.n_dev = h->d_cnt[g] + per * q;
            J.cur.keys_un = h->d_kps[g] + (size_t)per * cap * q;
            J.cur.desc = h->d_desc[g] + (size_t)32 * per * cap * q;
            J.cur.u_right = c.stereo ? h->d_ur + (size_t)cap * q : nullptr;
            J.cur.claimed = nullptr;
            J.cur.min_x = 0.f; J.cur.min_y = 0.f; J.cur.max_x = (float)c.width; J.cur.max_y = (float)c.height;   // undistorted input (Frame.cc:423-437)
            J.cur.grid_w_inv = 64.f / (J.cur.max_x - J.cur.min_x);                                                // Frame.cc:127-128
            J.cur.grid_h_inv = 48.f / (J.cur.max_y - J.cur.min_y);
            J.cur.fx = c.fx; J.cur.fy = c.fy; J.cur.cx = c.cx; J.cur.cy = c.cy; J.cur.bf = c.bf; J.cur.b = b;
            J.cur.scale_factors = h->d_sf;
            J.cur.nlevels = c.nlevels;
            J.n_last = cap;                                 // slots past the last frame's count are written invalid by k_unproject_last
            J.pts = h->d_pts + (size_t)cap * q;
            J.last_desc = h->d_desc[gl] + (size_t)32 * per * cap * q;
            const float *T = Tcw + 12 * q, *Tl = h->h_pose_last + 12 * q;
            for (int r = 0; r < 3; r++) { for (int k = 0; k < 3; k++) J.Rcw[3 * r + k] = T[4 * r + k]; J.tcw[r] = T[4 * r + 3]; }
            // bForward / bBackward (ORBmatcher.cc:1340-1351): twc = -Rcw.t()*tcw; tlc = Rlw*twc + tlw, in float left to right
            float twc[3], tlc2;
            for (int r = 0; r < 3; r++) twc[r] = -((T[r] * T[3] + T[4 + r] * T[7]) + T[8 + r] * T[11]);
            tlc2 = ((Tl[8] * twc[0] + Tl[9] * twc[1]) + Tl[10] * twc[2]) + Tl[11];
            J.forward = tlc2 > b && !c.mono;
            J.backward = -tlc2 > b && !c.mono;
            J.th = c.th;
            J.check_ori = c.check_ori;
            J.match = h->d_match + (size_t)cap * q;
            J.nmatches = h->d_nm + q;
            J.max_dist = 0;
            J.variant = 0;
        }
        ORBX_CUDA(cudaMemcpyAsync(h->d_jobs, h_jobs, sizeof(orbx_frame_match_job) * ns, cudaMemcpyHostToDevice, s));
        if ((st = orbx_match_projection_frame_device(h->mt, h->d_jobs, ns, s))) return st;
        h->last_launches += orbx_matcher_last_launches(h->mt);
    }
    ORBX_CUDA(cudaEventRecord(h->slot_ev[slot], s));
    memcpy(h->h_pose_last, Tcw, sizeof(float) * 12 * ns);
    h->gen ^= 1;
    h->steps++;
    return ORBX_OK;
}

extern "C" orbx_status orbx_sequences_step_begin(orbx_sequences *h, const uint8_t *images, size_t image_pitch, int stride,
                                                 const float *Tcw, const orbx_sequences_outputs *o) {
    if (!h || !images || !Tcw || !o || !o->counts || !o->nmatches || !o->match) return ORBX_ERR_INVALID;
    const orbx_sequences_config &c = h->c;
    if (stride < c.width) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(c.device));
    cudaStream_t s = h->stream;
    const int ns = c.n_sequences, cap = h->cap;
    if (h->in_flight) ORBX_CUDA(cudaStreamSynchronize(s));     // one step in flight per handle: the pinned job / pose staging is reused
    // ---- upload the new images (one 2-D copy: rows `stride` apart on the host, packed on the device) -------------------
    if (image_pitch == (size_t)stride * c.height)
        ORBX_CUDA(cudaMemcpy2DAsync(h->d_img, c.width, images, stride, c.width, (size_t)c.height * h->n_img, cudaMemcpyHostToDevice, s));
    else
        for (int i = 0; i < h->n_img; i++)
            ORBX_CUDA(cudaMemcpy2DAsync(h->d_img + (size_t)i * c.width * c.height, c.width, images + i * image_pitch, stride, c.width,
                                        c.height, cudaMemcpyHostToDevice, s));
    const orbx_status st = step_core(h, h->d_img, (size_t)c.width * c.height, c.width, Tcw, s);
    if (st) return st;
    const int g = h->gen ^ 1;                                   // the generation step_core just filled
    // ---- the step's results back to the caller's buffers -------------------------------------------------------------------
    ORBX_CUDA(cudaMemcpyAsync(o->counts, h->d_cnt[g], sizeof(int32_t) * h->n_img, cudaMemcpyDeviceToHost, s));
    if (o->kps) ORBX_CUDA(cudaMemcpyAsync(o->kps, h->d_kps[g], sizeof(orbx_keypoint) * (size_t)cap * h->n_img, cudaMemcpyDeviceToHost, s));
    if (o->desc) ORBX_CUDA(cudaMemcpyAsync(o->desc, h->d_desc[g], (size_t)32 * cap * h->n_img, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(o->match, h->d_match, sizeof(int32_t) * (size_t)cap * ns, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(o->nmatches, h->d_nm, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, s));
    if (c.stereo && o->u_right) ORBX_CUDA(cudaMemcpyAsync(o->u_right, h->d_ur, sizeof(float) * (size_t)cap * ns, cudaMemcpyDeviceToHost, s));
    if (c.stereo && o->depth) ORBX_CUDA(cudaMemcpyAsync(o->depth, h->d_depth[g], sizeof(float) * (size_t)cap * ns, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(h->ex->h_status, h->ex->d_status, sizeof(int) * h->n_img, cudaMemcpyDeviceToHost, s));
    h->in_flight = 1;
    return ORBX_OK;
}

extern "C" orbx_status orbx_sequences_step_device(orbx_sequences *h, const uint8_t *d_images, size_t frame_pitch, int stride,
                                                  const float *Tcw, void *stream) {
    if (!h || !d_images || !Tcw || stride < h->c.width) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->c.device));
    return step_core(h, d_images, frame_pitch, stride, Tcw, (cudaStream_t)stream);
}

extern "C" orbx_status orbx_sequences_device_view(const orbx_sequences *h, orbx_sequences_device *v) {
    if (!h || !v) return ORBX_ERR_INVALID;
    const int g = h->gen ^ 1;                                   // the generation the last step filled
    v->extractor = h->ex;
    v->kps = h->d_kps[g]; v->desc = h->d_desc[g]; v->counts = h->d_cnt[g];
    v->match = h->d_match; v->nmatches = h->d_nm; v->u_right = h->d_ur; v->depth = h->d_depth[g];
    v->jobs = h->d_jobs;
    v->stream = h->stream;
    return ORBX_OK;
}

extern "C" orbx_status orbx_sequences_step_end(orbx_sequences *h) {
    if (!h) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->c.device));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    if (h->in_flight) {
        h->in_flight = 0;
        for (int i = 0; i < h->n_img; i++)
            if (h->ex->h_status[i]) {
                orbx_set_error("image %d: device status 0x%x (1 = candidate list overflow, 2 = quadtree depth, 4 = node overflow)", i, h->ex->h_status[i]);
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
