// Stereo association: replaces Frame::ComputeStereoMatches (reference src/Frame.cc:495-669).
//
// Two launches per batch of rectified pairs, everything read where the extractor left it in HBM (keypoints,
// descriptors, both pyramids = mvImagePyramid of the left / right ORBextractor):
//   k_stereo_search  grid (splits, pairs): a CTA stages the right keypoints' row bands in shared memory (the reference's
//                    vRowIndices table, :505-522, is "minr <= row <= maxr" per right keypoint) and gives each warp a
//                    share of the left keypoints: lanes scan the right keypoints in ascending index (ballot skips
//                    empty chunks), Hamming distance only for those on the row / octave / disparity band, warp-min of
//                    (dist, iR) = the reference's first strict minimum (:573-577); then the 11 SADs of the 11x11
//                    window (:590-623) from a per-warp shared-memory copy of the two patches, parabola fit and the
//                    disparity gates (:630-652) in the reference's float operation order.
//   k_stereo_median  one CTA per pair: the median of the accepted SADs by a two-pass radix select (:656-657 sorts
//                    pairs, but only the value at rank size/2 is used), then the cut at 1.5f*1.4f*median (:658-668).
// All SAD arithmetic is integer (u8 differences); cv::norm(IL, IR, NORM_L1) on CV_32F is exact on these values.
#include "orbx_internal.cuh"

#define ST_THREADS 256
#define ST_WARPS (ST_THREADS / 32)
#define ST_PER_CTA 128          // left keypoints per CTA
#define ST_TH_HIGH 100          // ORBmatcher::TH_HIGH
#define ST_TH_ORB 75            // (TH_HIGH + TH_LOW) / 2, Frame.cc:500
#define ST_W 5
#define ST_L 5
#define ST_MED_THREADS 1024

struct orbx_stereo {
    int device, max_kp, max_pairs;
    int32_t *d_sad;            // [max_pairs][max_kp] SAD pushed into vDistIdx, -1 = none
    // staging of the _host entry point (one pair)
    orbx_keypoint *d_keys[2]; uint8_t *d_desc[2]; float *d_out[2]; int32_t *d_kept;
    cudaStream_t stream;
    int last_launches;
};

struct StereoSideDev {
    const orbx_keypoint *keys; const uint8_t *desc; const int32_t *counts;
    int pitch, count_step, n_fixed;     // n_fixed >= 0: that many keypoints in every pair (host entry point)
    const uint8_t *pyr; size_t pyr_frame; int first_slot, slot_step;
};
struct StereoTables {
    float scale[ORBX_MAX_LEVELS], inv_scale[ORBX_MAX_LEVELS];
    int w[ORBX_MAX_LEVELS], h[ORBX_MAX_LEVELS], pitch[ORBX_MAX_LEVELS];
    unsigned long long off[ORBX_MAX_LEVELS];   // byte offset of the first INTERIOR pixel of the level inside a frame's block
    int nlevels;
};

__device__ __forceinline__ int st_hamming(const uint4 a0, const uint4 a1, const uint8_t *b) {
    const uint4 b0 = __ldg(reinterpret_cast<const uint4 *>(b)), b1 = __ldg(reinterpret_cast<const uint4 *>(b + 16));
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(ST_THREADS)
k_stereo_search(const StereoSideDev Ls, const StereoSideDev Rs, const __grid_constant__ StereoTables T, float bf, float b,
                float *__restrict__ u_right, float *__restrict__ depth, int out_pitch, int32_t *__restrict__ sad_out, int sad_pitch) {
    extern __shared__ __align__(16) uint8_t st_smem[];
    __shared__ uint8_t patch[ST_WARPS][11 * 11 + 11 * 21 + 4];
    const int pair = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nl = Ls.n_fixed >= 0 ? Ls.n_fixed : Ls.counts[(size_t)pair * Ls.count_step];
    const int nr = Rs.n_fixed >= 0 ? Rs.n_fixed : Rs.counts[(size_t)pair * Rs.count_step];
    const int i0 = blockIdx.x * ST_PER_CTA;
    if (i0 >= nl) return;       // (the grid covers max_count keypoints; a larger count is the caller's contract violation)
    const orbx_keypoint *kl = Ls.keys + (size_t)pair * Ls.pitch, *kr = Rs.keys + (size_t)pair * Rs.pitch;
    const uint8_t *dl = Ls.desc + (size_t)pair * Ls.pitch * 32, *dr = Rs.desc + (size_t)pair * Rs.pitch * 32;
    const uint8_t *pl = Ls.pyr + (size_t)(Ls.first_slot + pair * Ls.slot_step) * Ls.pyr_frame;
    const uint8_t *pr = Rs.pyr + (size_t)(Rs.first_slot + pair * Rs.slot_step) * Rs.pyr_frame;
    float *ur_o = u_right + (size_t)pair * out_pitch, *dp_o = depth + (size_t)pair * out_pitch;
    int32_t *sad_o = sad_out + (size_t)pair * sad_pitch;

    // right keypoints: x and (minr | maxr << 12 | octave << 24), Frame.cc:512-521
    const int nr_pad = (nr + 31) & ~31;
    float *rx = reinterpret_cast<float *>(st_smem);
    uint32_t *rband = reinterpret_cast<uint32_t *>(rx + nr_pad);
    const int n_rows = T.h[0];
    for (int j = tid; j < nr_pad; j += ST_THREADS) {
        float x = 0.f; uint32_t band = 0xfffu;       // minr = 4095 > maxr = 0: never on a row
        if (j < nr) {
            const orbx_keypoint k = kr[j];
            const int oc = min(max(k.octave, 0), T.nlevels - 1);
            const float r = __fmul_rn(2.0f, T.scale[oc]);
            int maxr = (int)ceilf(__fadd_rn(k.y, r)), minr = (int)floorf(__fsub_rn(k.y, r));
            minr = max(minr, 0); maxr = min(maxr, n_rows - 1);     // rows outside the image do not exist (reference: UB)
            if (minr <= maxr) band = (uint32_t)minr | ((uint32_t)maxr << 12) | ((uint32_t)(k.octave & 0xff) << 24);
            x = k.x;
        }
        rx[j] = x; rband[j] = band;
    }
    __syncthreads();

    const float minZ = b, minD = 0.0f, maxD = __fdiv_rn(bf, minZ);                     // :525-527
    const int i1 = min(i0 + ST_PER_CTA, nl);
    for (int iL = i0 + warp; iL < i1; iL += ST_WARPS) {
        const orbx_keypoint k = kl[iL];
        const int levelL = k.octave;
        const float vL = k.y, uL = k.x;
        float o_ur = -1.0f, o_dp = -1.0f; int o_sad = -1;
        const int rowL = (int)vL;
        const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);
        bool alive = rowL >= 0 && rowL < n_rows && !(maxU < 0) && levelL >= 0 && levelL < T.nlevels;
        unsigned best = 0xffffffffu;
        if (alive) {
            const uint4 a0 = __ldg(reinterpret_cast<const uint4 *>(dl + (size_t)iL * 32));
            const uint4 a1 = __ldg(reinterpret_cast<const uint4 *>(dl + (size_t)iL * 32 + 16));
            for (int jb = 0; jb < nr_pad; jb += 32) {
                const int j = jb + lane;
                const uint32_t band = rband[j];
                const int minr = band & 0xfff, maxr = (band >> 12) & 0xfff, oc = (int)(band >> 24);
                const float uR = rx[j];
                const bool ok = rowL >= minr && rowL <= maxr && oc >= levelL - 1 && oc <= levelL + 1 && uR >= minU && uR <= maxU;
                if (ok) {
                    const unsigned d = (unsigned)st_hamming(a0, a1, dr + (size_t)j * 32);
                    best = min(best, (d << 16) | (unsigned)j);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        const int bestDist = (int)(best >> 16), bestIdxR = (int)(best & 0xffff);
        // bestDist starts at TH_HIGH and only a strictly smaller distance replaces it (:552, :573)
        alive = alive && best != 0xffffffffu && bestDist < ST_TH_HIGH && bestDist < ST_TH_ORB;
        if (alive) {                                                                    // warp-uniform from here on
            const float uR0 = rx[bestIdxR];
            const float sf = T.inv_scale[levelL];
            const float scaleduL = roundf(__fmul_rn(k.x, sf)), scaledvL = roundf(__fmul_rn(k.y, sf));
            const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
            const int lw = T.w[levelL], lh = T.h[levelL], lp = T.pitch[levelL];
            const int yl = (int)scaledvL, xl = (int)scaleduL, xr0 = (int)scaleduR0;
            const float iniu = __fadd_rn(scaleduR0, (float)(ST_L - ST_W)), endu = __fadd_rn(scaleduR0, (float)(ST_L + ST_W + 1));
            bool ok = !(yl - ST_W < 0 || yl + ST_W + 1 > lh || xl - ST_W < 0 || xl + ST_W + 1 > lw);
            ok = ok && !(iniu < 0 || endu >= (float)lw) && !(xr0 - ST_L - ST_W < 0);
            if (ok) {
                // stage IL (11x11) and the IR strip (11x21) of this level
                uint8_t *pp = patch[warp];
                const uint8_t *gl = pl + T.off[levelL] + (size_t)(yl - ST_W) * lp + (xl - ST_W);
                const uint8_t *gr = pr + T.off[levelL] + (size_t)(yl - ST_W) * lp + (xr0 - ST_L - ST_W);
                __syncwarp();
                for (int t = lane; t < 121 + 231; t += 32) {
                    if (t < 121) { const int r = t / 11, c = t - r * 11; pp[t] = __ldg(gl + (size_t)r * lp + c); }
                    else { const int q = t - 121, r = q / 21, c = q - r * 21; pp[t] = __ldg(gr + (size_t)r * lp + c); }
                }
                __syncwarp();
                // lane = incR + 5 + 11 * part; part 0: rows 0..5, part 1: rows 6..10
                int acc = 0;
                if (lane < 22) {
                    const int part = lane >= 11, inc = lane - 11 * part;            // inc = incR + L in 0..10
                    const int cL = pp[5 * 11 + 5], cR = pp[121 + 5 * 21 + inc + 5];
                    const int r0 = part ? 6 : 0, r1 = part ? 11 : 6;
                    for (int r = r0; r < r1; r++) {
                        const uint8_t *a = pp + r * 11, *q = pp + 121 + r * 21 + inc;
#pragma unroll
                        for (int c = 0; c < 11; c++) acc += abs(((int)a[c] - cL) - ((int)q[c] - cR));
                    }
                }
                acc += __shfl_down_sync(0xffffffffu, acc, 11);                        // lanes 0..10: the 11 distances
                unsigned key = lane < 11 ? ((unsigned)acc << 4) | (unsigned)lane : 0xffffffffu;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
                const int bestinc = (int)(key & 15), bestSad = (int)(key >> 4);       // first strict minimum, :617-621
                const int i1s = min(max(bestinc - 1, 0), 10), i3s = min(bestinc + 1, 10);
                const float dist1 = (float)__shfl_sync(0xffffffffu, acc, i1s);
                const float dist3 = (float)__shfl_sync(0xffffffffu, acc, i3s);
                const float dist2 = (float)bestSad;
                if (bestinc != 0 && bestinc != 2 * ST_L) {                            // :625-626
                    const float num = __fsub_rn(dist1, dist3);
                    const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2)));
                    const float deltaR = __fdiv_rn(num, den);                         // :633 (0/0 = NaN passes the next test, like the reference)
                    if (!(deltaR < -1 || deltaR > 1)) {
                        float bestuR = __fmul_rn(T.scale[levelL], __fadd_rn(__fadd_rn(scaleduR0, (float)(bestinc - ST_L)), deltaR));
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= minD && disparity < maxD) {                  // :643-652
                            if (disparity <= 0) {
                                disparity = 0.01f;
                                bestuR = (float)((double)uL - 0.01);
                            }
                            o_dp = __fdiv_rn(bf, disparity);
                            o_ur = bestuR;
                            o_sad = bestSad;
                        }
                    }
                }
            }
        }
        if (lane == 0) { ur_o[iL] = o_ur; dp_o[iL] = o_dp; sad_o[iL] = o_sad; }
    }
}

// median cut, Frame.cc:656-668
__global__ void __launch_bounds__(ST_MED_THREADS)
k_stereo_median(const StereoSideDev Ls, float *__restrict__ u_right, float *__restrict__ depth, int out_pitch,
                const int32_t *__restrict__ sad_in, int sad_pitch, int32_t *__restrict__ kept_out) {
    __shared__ int hist[256];
    __shared__ int sel_bin, sel_rank, n_acc, n_cut;
    const int pair = blockIdx.x, tid = threadIdx.x;
    const int nl = Ls.n_fixed >= 0 ? Ls.n_fixed : Ls.counts[(size_t)pair * Ls.count_step];
    const int32_t *sad = sad_in + (size_t)pair * sad_pitch;
    float *ur_o = u_right + (size_t)pair * out_pitch, *dp_o = depth + (size_t)pair * out_pitch;
    if (tid < 256) hist[tid] = 0;
    if (tid == 0) { n_acc = 0; n_cut = 0; }
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < nl; i += ST_MED_THREADS) {
        const int s = sad[i];
        if (s >= 0) { atomicAdd(&hist[min(s >> 8, 255)], 1); mine++; }
    }
    if (mine) atomicAdd(&n_acc, mine);
    __syncthreads();
    const int n = n_acc;
    if (n == 0) { if (tid == 0 && kept_out) kept_out[pair] = 0; return; }
    if (tid == 0) {                                     // bin that holds rank n/2 of the ascending order
        int k = n / 2, bsel = 0;
        for (; bsel < 256; bsel++) { if (k < hist[bsel]) break; k -= hist[bsel]; }
        sel_bin = bsel; sel_rank = k;
    }
    __syncthreads();
    const int bsel = sel_bin;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < nl; i += ST_MED_THREADS) {
        const int s = sad[i];
        if (s >= 0 && min(s >> 8, 255) == bsel) atomicAdd(&hist[s & 255], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int k = sel_rank, lo = 0;
        for (; lo < 256; lo++) { if (k < hist[lo]) break; k -= hist[lo]; }
        sel_rank = (bsel << 8) | lo;
    }
    __syncthreads();
    const float median = (float)sel_rank;
    const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), median);
    int cut = 0;
    for (int i = tid; i < nl; i += ST_MED_THREADS) {
        const int s = sad[i];
        if (s >= 0 && !((float)s < thDist)) { ur_o[i] = -1.0f; dp_o[i] = -1.0f; cut++; }
    }
    if (cut) atomicAdd(&n_cut, cut);
    __syncthreads();
    if (tid == 0 && kept_out) kept_out[pair] = n - n_cut;
}

// ---- host side ----------------------------------------------------------------------------------------------------------
extern "C" void orbx_stereo_destroy(orbx_stereo *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_sad);
    for (int s = 0; s < 2; s++) { cudaFree(h->d_keys[s]); cudaFree(h->d_desc[s]); cudaFree(h->d_out[s]); }
    cudaFree(h->d_kept);
    if (h->stream) cudaStreamDestroy(h->stream);
    free(h);
}

extern "C" orbx_status orbx_stereo_create(orbx_stereo **out, int max_keypoints, int max_pairs, int device) {
    if (!out) return ORBX_ERR_INVALID;
    *out = nullptr;
    if (max_keypoints < 1 || max_pairs < 1 || max_keypoints > 65535) {
        orbx_set_error("orbx_stereo_create: bad argument (1..65535 keypoints per image)");
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
    orbx_stereo *h = (orbx_stereo *)calloc(1, sizeof(orbx_stereo));
    if (!h) return ORBX_ERR_NOMEM;
    h->device = device; h->max_kp = max_keypoints; h->max_pairs = max_pairs;
    const size_t kp = (size_t)max_keypoints;
    cudaError_t ce = cudaSuccess;
#define TRY(x) if (ce == cudaSuccess) ce = (x)
    TRY(cudaMalloc((void **)&h->d_sad, sizeof(int32_t) * kp * max_pairs));
    for (int s = 0; s < 2; s++) {
        TRY(cudaMalloc((void **)&h->d_keys[s], sizeof(orbx_keypoint) * kp));
        TRY(cudaMalloc((void **)&h->d_desc[s], 32 * kp));
        TRY(cudaMalloc((void **)&h->d_out[s], sizeof(float) * kp));
    }
    TRY(cudaMalloc((void **)&h->d_kept, sizeof(int32_t)));
    TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    TRY(ORBX_RAISE_SMEM(k_stereo_search));
#undef TRY
    if (ce != cudaSuccess) {
        orbx_set_error("orbx_stereo_create: %s", cudaGetErrorString(ce));
        orbx_stereo_destroy(h);
        return ORBX_ERR_CUDA;
    }
    // k_stereo_search keeps 8 bytes per right keypoint in shared memory: refuse here what its launch could not get
    cudaFuncAttributes fa;
    ORBX_CUDA(cudaFuncGetAttributes(&fa, (const void *)k_stereo_search));
    if ((size_t)8 * ((kp + 31) & ~(size_t)31) > (size_t)fa.maxDynamicSharedSizeBytes) {
        orbx_set_error("orbx_stereo_create: %d keypoints per image need %zu bytes of shared memory, %d available (at most %d keypoints)",
                       max_keypoints, (size_t)8 * ((kp + 31) & ~(size_t)31), fa.maxDynamicSharedSizeBytes,
                       (fa.maxDynamicSharedSizeBytes / 8) & ~31);
        orbx_stereo_destroy(h);
        return ORBX_ERR_CAPACITY;
    }
    *out = h;
    return ORBX_OK;
}

static orbx_status stereo_side(const orbx_stereo_side *S, StereoSideDev *D, int n_pairs) {
    const orbx_extractor *e = S->extractor;
    if (!e || e->cur_w == 0 || S->first_slot < 0 || S->first_slot + (long long)(n_pairs - 1) * S->slot_step >= e->max_batch ||
        S->first_slot + (long long)(n_pairs - 1) * S->slot_step < 0) {
        orbx_set_error("orbx_stereo: pyramid slots outside the extractor's batch (or the extractor has not run)");
        return ORBX_ERR_INVALID;
    }
    D->keys = S->keys; D->desc = S->desc; D->counts = S->counts; D->pitch = S->pitch; D->count_step = S->count_step;
    D->n_fixed = -1;
    D->pyr = e->d_pyr; D->pyr_frame = e->pyr_frame_cap; D->first_slot = S->first_slot; D->slot_step = S->slot_step;
    return ORBX_OK;
}

static orbx_status stereo_launch(orbx_stereo *h, const StereoSideDev &L, const StereoSideDev &R, const orbx_extractor *el,
                                 const orbx_extractor *er, int n_pairs, int max_left, int max_right, float bf, float b,
                                 float *d_u_right, float *d_depth, int out_pitch, int32_t *d_kept, cudaStream_t s) {
    if (el->cur_w != er->cur_w || el->cur_h != er->cur_h || el->nlevels != er->nlevels || el->scale_factor != er->scale_factor) {
        orbx_set_error("orbx_stereo: left and right extractors differ in image size or pyramid");
        return ORBX_ERR_INVALID;
    }
    if (el->device != h->device || er->device != h->device) return ORBX_ERR_INVALID;
    if (max_left > h->max_kp || max_right > h->max_kp || n_pairs > h->max_pairs) {
        orbx_set_error("orbx_stereo: %d pairs of up to %d / %d keypoints, handle was created for %d pairs of %d", n_pairs, max_left,
                       max_right, h->max_pairs, h->max_kp);
        return ORBX_ERR_CAPACITY;
    }
    StereoTables T;
    memset(&T, 0, sizeof(T));
    T.nlevels = el->nlevels;
    for (int l = 0; l < el->nlevels; l++) {
        const OrbxLevel &v = el->lv[l];
        if (er->lv[l].pitch != v.pitch || er->lv[l].off != v.off) return ORBX_ERR_INVALID;
        T.scale[l] = el->scale[l]; T.inv_scale[l] = el->inv_scale[l];
        T.w[l] = v.w; T.h[l] = v.h; T.pitch[l] = v.pitch;
        T.off[l] = v.off + (size_t)ORBX_EDGE * v.pitch + ORBX_EDGE;
    }
    h->last_launches = 0;
    if (n_pairs == 0 || max_left == 0) return ORBX_OK;
    const size_t smem = (size_t)8 * ((max_right + 31) & ~31);
    dim3 grid((max_left + ST_PER_CTA - 1) / ST_PER_CTA, n_pairs);
    k_stereo_search<<<grid, ST_THREADS, smem, s>>>(L, R, T, bf, b, d_u_right, d_depth, out_pitch, h->d_sad, h->max_kp);
    ORBX_CUDA(cudaGetLastError());
    k_stereo_median<<<n_pairs, ST_MED_THREADS, 0, s>>>(L, d_u_right, d_depth, out_pitch, h->d_sad, h->max_kp, d_kept);
    ORBX_CUDA(cudaGetLastError());
    h->last_launches = 2;
    return ORBX_OK;
}

extern "C" orbx_status orbx_stereo_matches_device(orbx_stereo *h, const orbx_stereo_side *left, const orbx_stereo_side *right,
                                                  int n_pairs, float bf, float b, float *d_u_right, float *d_depth, int out_pitch,
                                                  int32_t *d_kept, void *stream) {
    if (!h || !left || !right || n_pairs < 0 || !d_u_right || !d_depth) return ORBX_ERR_INVALID;
    if (n_pairs == 0) { h->last_launches = 0; return ORBX_OK; }
    if (!left->keys || !left->desc || !left->counts || !right->keys || !right->desc || !right->counts) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(h->device));
    StereoSideDev L, R;
    orbx_status st;
    if ((st = stereo_side(left, &L, n_pairs)) != ORBX_OK || (st = stereo_side(right, &R, n_pairs)) != ORBX_OK) return st;
    // counts live on the device: size the grid for the arrays' pitch (CTAs past a pair's count return at once)
    const int max_l = left->max_count > 0 ? left->max_count : left->pitch, max_r = right->max_count > 0 ? right->max_count : right->pitch;
    if (out_pitch < max_l) return ORBX_ERR_INVALID;
    return stereo_launch(h, L, R, left->extractor, right->extractor, n_pairs, max_l, max_r, bf, b, d_u_right, d_depth, out_pitch, d_kept,
                         (cudaStream_t)stream);
}

extern "C" orbx_status orbx_stereo_matches_host(orbx_stereo *h, const orbx_extractor *left, int left_slot,
                                                const orbx_extractor *right, int right_slot, const orbx_keypoint *keys_l,
                                                const uint8_t *desc_l, int n_left, const orbx_keypoint *keys_r,
                                                const uint8_t *desc_r, int n_right, float bf, float b, float *u_right, float *depth,
                                                int32_t *n_kept) {
    if (!h || !left || !right || n_left < 0 || n_right < 0 || (n_left && (!keys_l || !desc_l || !u_right || !depth)) ||
        (n_right && (!keys_r || !desc_r)))
        return ORBX_ERR_INVALID;
    if (n_kept) *n_kept = 0;
    if (n_left == 0) return ORBX_OK;
    if (n_left > h->max_kp || n_right > h->max_kp) {
        orbx_set_error("orbx_stereo: %d / %d keypoints, handle was created for %d", n_left, n_right, h->max_kp);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    orbx_stereo_side sl = {h->d_keys[0], h->d_desc[0], nullptr, h->max_kp, 0, left, left_slot, 0, 0};
    orbx_stereo_side sr = {h->d_keys[1], h->d_desc[1], nullptr, h->max_kp, 0, right, right_slot, 0, 0};
    StereoSideDev L, R;
    orbx_status st;
    if ((st = stereo_side(&sl, &L, 1)) != ORBX_OK || (st = stereo_side(&sr, &R, 1)) != ORBX_OK) return st;
    L.n_fixed = n_left; R.n_fixed = n_right;
    // the pyramids were written on whatever stream ran the extractors: make them visible first
    ORBX_CUDA(cudaDeviceSynchronize());
    ORBX_CUDA(cudaMemcpyAsync(h->d_keys[0], keys_l, sizeof(orbx_keypoint) * n_left, cudaMemcpyHostToDevice, s));
    ORBX_CUDA(cudaMemcpyAsync(h->d_desc[0], desc_l, (size_t)32 * n_left, cudaMemcpyHostToDevice, s));
    if (n_right) {
        ORBX_CUDA(cudaMemcpyAsync(h->d_keys[1], keys_r, sizeof(orbx_keypoint) * n_right, cudaMemcpyHostToDevice, s));
        ORBX_CUDA(cudaMemcpyAsync(h->d_desc[1], desc_r, (size_t)32 * n_right, cudaMemcpyHostToDevice, s));
    }
    st = stereo_launch(h, L, R, left, right, 1, n_left, n_right, bf, b, h->d_out[0], h->d_out[1], h->max_kp, h->d_kept, s);
    if (st != ORBX_OK) return st;
    ORBX_CUDA(cudaMemcpyAsync(u_right, h->d_out[0], sizeof(float) * n_left, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(depth, h->d_out[1], sizeof(float) * n_left, cudaMemcpyDeviceToHost, s));
    int32_t kept = 0;
    ORBX_CUDA(cudaMemcpyAsync(&kept, h->d_kept, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    if (n_kept) *n_kept = kept;
    return ORBX_OK;
}

// the same for the pair the two extractors' _host entry points just processed: keypoints, descriptors and counts are read from
// the extractors' own device output buffers (mvKeys / mDescriptors are exactly what operator() returned), nothing is uploaded
extern "C" orbx_status orbx_stereo_matches_extractors_host(orbx_stereo *h, const orbx_extractor *left, int left_slot,
                                                           const orbx_extractor *right, int right_slot, int n_left, float bf, float b,
                                                           float *u_right, float *depth, int32_t *n_kept) {
    if (!h || !left || !right || n_left < 0 || (n_left && (!u_right || !depth))) return ORBX_ERR_INVALID;
    if (n_kept) *n_kept = 0;
    if (n_left == 0) return ORBX_OK;
    if (left_slot < 0 || left_slot >= left->last_batch || right_slot < 0 || right_slot >= right->last_batch) {
        orbx_set_error("orbx_stereo: slot outside the extractor's last batch");
        return ORBX_ERR_INVALID;
    }
    if (left->capacity > h->max_kp || right->capacity > h->max_kp || n_left > left->capacity) {
        orbx_set_error("orbx_stereo: extractor capacity %d / %d, handle was created for %d keypoints", left->capacity, right->capacity, h->max_kp);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    orbx_stereo_side sl = {left->d_kps + (size_t)left_slot * left->capacity, left->d_desc + (size_t)32 * left_slot * left->capacity,
                           left->d_counts + left_slot, left->capacity, 0, left, left_slot, 0, 0};
    orbx_stereo_side sr = {right->d_kps + (size_t)right_slot * right->capacity, right->d_desc + (size_t)32 * right_slot * right->capacity,
                           right->d_counts + right_slot, right->capacity, 0, right, right_slot, 0, 0};
    StereoSideDev L, R;
    orbx_status st;
    if ((st = stereo_side(&sl, &L, 1)) != ORBX_OK || (st = stereo_side(&sr, &R, 1)) != ORBX_OK) return st;
    // the extractors' _host calls have returned, so their outputs and pyramids are complete; this handle's stream starts after them
    st = stereo_launch(h, L, R, left, right, 1, left->capacity, right->capacity, bf, b, h->d_out[0], h->d_out[1], h->max_kp, h->d_kept, s);
    if (st != ORBX_OK) return st;
    ORBX_CUDA(cudaMemcpyAsync(u_right, h->d_out[0], sizeof(float) * n_left, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(depth, h->d_out[1], sizeof(float) * n_left, cudaMemcpyDeviceToHost, s));
    int32_t kept = 0;
    ORBX_CUDA(cudaMemcpyAsync(&kept, h->d_kept, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    if (n_kept) *n_kept = kept;
    return ORBX_OK;
}

extern "C" int orbx_stereo_last_launches(const orbx_stereo *h) { return h ? h->last_launches : 0; }
