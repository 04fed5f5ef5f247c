// Host side of the extractor C ABI (include/orbx.h): parameter tables of ORBextractor::ORBextractor (reference
// src/ORBextractor.cc:410-470), per-image-size geometry (level sizes :1112, FAST cell grid :773-806, quadtree
// split paths :481-537/:543-559, OpenCV resize coefficient tables), device buffers, and the launch sequence
// that replaces ORBextractor::operator() (:1043-1105).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "orbx_internal.cuh"

static thread_local char g_err[512] = "";

void orbx_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

cudaError_t orbx_raise_smem(const void *kernel) {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kernel);
    if (e != cudaSuccess) return e;
    int dev = 0, optin = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)a.sharedSizeBytes);
}

extern "C" const char *orbx_last_error(void) { return g_err; }
extern "C" int orbx_version(void) { return 100; }

orbx_status orbx_octree_init(int smem_bytes);
orbx_status orbx_fast_init(size_t smem_bytes);

static inline int cv_round_f(float v) { return (int)lrintf(v); }   // cvRound: SSE cvtss2si, round-half-even
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename T>
static orbx_status grow(T **p, size_t *cap, size_t need) {
    if (need <= *cap && *p) return ORBX_OK;
    if (*p) ORBX_CUDA(cudaFree(*p));
    *p = nullptr;
    ORBX_CUDA(cudaMalloc((void **)p, sizeof(T) * need));
    *cap = need;
    return ORBX_OK;
}

// ---- per-size geometry ----------------------------------------------------------------------------------
struct Geometry {
    OrbxLevel lv[ORBX_MAX_LEVELS];
    size_t pyr_bytes, blur_bytes, cand_words;
    std::vector<uint32_t> rxt;
    std::vector<int2> ryt;
    std::vector<uint32_t> lut;
    std::vector<OrbxFastChunk> chunks;
    std::vector<OrbxBlurTile> btiles;
    int fast_tp, fast_th;
};

// split paths of one axis: for every coordinate v in [0, n] the sequence of "v < mid" decisions that
// ExtractorNode::DivideNode (ORBextractor.cc:481-525) takes from the initial node downwards, 15 levels deep
static void axis_paths(int n, int n_ini, float hx, bool is_x, std::vector<int> &bucket, std::vector<int> &path) {
    bucket.resize(n + 1);
    path.resize(n + 1);
    for (int v = 0; v <= n; v++) {
        int b = 0, lo = 0, hi = n;
        if (is_x) {
            b = (int)((float)v / hx);                       // vpIniNodes[kp.pt.x/hX], ORBextractor.cc:568
            if (b >= n_ini) b = n_ini - 1;
            lo = (int)(hx * (float)b);                      // ni.UL.x, ORBextractor.cc:555
            hi = (int)(hx * (float)(b + 1));
        }
        int bits = 0;
        for (int d = 0; d < 15; d++) {
            const int mid = lo + (int)ceilf((float)(hi - lo) / 2);
            if ((float)v < (float)mid) { bits <<= 1; hi = mid; }
            else { bits = (bits << 1) | 1; lo = mid; }
        }
        bucket[v] = b;
        path[v] = bits;
    }
}

static bool paths_distinct(const std::vector<int> &bucket, const std::vector<int> &path, int D) {
    // coordinates are monotone along a bucket, and so are their paths: distinct iff strictly increasing
    for (size_t v = 1; v < path.size(); v++)
        if (bucket[v] == bucket[v - 1] && (path[v] >> (15 - D)) == (path[v - 1] >> (15 - D))) return false;
    return true;
}

static uint32_t spread_bits(uint32_t v) {
    uint32_t r = 0;
    for (int i = 0; i < 16; i++) r |= ((v >> i) & 1u) << (2 * i);
    return r;
}

static orbx_status build_geometry(const orbx_extractor *e, int w, int h, Geometry &g) {
    size_t pyr = 0, blur = 0, cand = 0;
    int kp_off = 0;
    g.fast_tp = g.fast_th = 0;
    for (int l = 0; l < e->nlevels; l++) {
        OrbxLevel &L = g.lv[l];
        memset(&L, 0, sizeof(L));
        L.w = cv_round_f((float)w * e->inv_scale[l]);
        L.h = cv_round_f((float)h * e->inv_scale[l]);
        if (L.w <= 2 * ORBX_EDGE || L.h <= 2 * ORBX_EDGE) {
            orbx_set_error("level %d is %dx%d: too small for the 19-pixel border", l, L.w, L.h);
            return ORBX_ERR_UNSUPPORTED;
        }
        L.pitch = (int)align_up(L.w + 2 * ORBX_EDGE, 16);
        L.ph = L.h + 2 * ORBX_EDGE;
        L.off = pyr;
        pyr += align_up((size_t)L.pitch * L.ph, 256);
        L.bpitch = (int)align_up(L.w, 16);
        L.boff = blur;
        blur += align_up((size_t)L.bpitch * L.h, 256);
        L.scale = e->scale[l];
        L.kp_size = (float)(int)(31 * e->scale[l]);
        L.quota = e->quota[l];
        L.kp_cap = L.quota + 3 > 4 * ORBX_NINI_MAX ? L.quota + 3 : 4 * ORBX_NINI_MAX;
        L.kp_off = kp_off;
        kp_off += L.kp_cap;
        // FAST grid, ORBextractor.cc:773-783
        L.bw = L.w - 2 * ORBX_BORDER;
        L.bh = L.h - 2 * ORBX_BORDER;
        if (L.bw >= ORBX_MAX_DIM || L.bh >= ORBX_MAX_DIM) {
            orbx_set_error("level %d is %dx%d: larger than %d", l, L.w, L.h, ORBX_MAX_DIM);
            return ORBX_ERR_UNSUPPORTED;
        }
        const float width = (float)L.bw, height = (float)L.bh;
        L.ncols = (int)(width / 30.f);
        L.nrows = (int)(height / 30.f);
        if (L.ncols > 0 && L.nrows > 0) {
            L.wcell = (int)ceilf(width / L.ncols);
            L.hcell = (int)ceilf(height / L.nrows);
        } else {
            L.ncols = L.nrows = 0;
            L.wcell = L.hcell = 1;
        }
        L.cand_cap = (L.bw / 2 + L.ncols + 1) * (L.bh / 2 + L.nrows + 1);
        L.cand_off = cand;
        cand += align_up((size_t)L.cand_cap, 64);
        // quadtree, ORBextractor.cc:543-559
        L.n_ini = (int)roundf(width / height);
        if (L.n_ini < 1 || L.n_ini > ORBX_NINI_MAX) {
            orbx_set_error("level %d: %d initial quadtree nodes (aspect ratio outside 1:2 .. %d:1)", l, L.n_ini, ORBX_NINI_MAX);
            return ORBX_ERR_UNSUPPORTED;
        }
        const float hx = width / L.n_ini;
        std::vector<int> bx, px, by, py;
        axis_paths(L.bw, L.n_ini, hx, true, bx, px);
        axis_paths(L.bh, 1, 0.f, false, by, py);
        int D = 1;
        while (D < 15 && !(paths_distinct(bx, px, D) && paths_distinct(by, py, D))) D++;
        int ini_bits = 0;
        while ((1 << ini_bits) < L.n_ini) ini_bits++;
        if (!(paths_distinct(bx, px, D) && paths_distinct(by, py, D)) || ini_bits + 2 * D > 32) {
            orbx_set_error("level %d: quadtree key does not fit 32 bits", l);
            return ORBX_ERR_UNSUPPORTED;
        }
        L.depth = D;
        int d0 = 0;
        while (d0 < D && ((size_t)L.n_ini << (2 * (d0 + 1))) <= ORBX_OCT_CELLS) d0++;
        L.ncells = L.n_ini << (2 * d0);
        L.cshift = 2 * (D - d0);
        L.lutx_off = (int)g.lut.size();
        for (int x = 0; x <= L.bw; x++)
            g.lut.push_back(((uint32_t)bx[x] << (2 * D)) | spread_bits((uint32_t)(px[x] >> (15 - D))));
        L.luty_off = (int)g.lut.size();
        for (int y = 0; y <= L.bh; y++) g.lut.push_back(spread_bits((uint32_t)(py[y] >> (15 - D))) << 1);
        // cv::resize tables (OpenCV resize.cpp, INTER_LINEAR 8U: 11-bit coefficients), indexed by PADDED output
        // coordinates so that the REFLECT_101 border of ORBextractor.cc:1122 costs the kernel nothing:
        //   x entry (u32, one per padded column, 16-byte aligned groups of 4): sx | a1 << 16 | group_flag << 31, a0 = 2048 - a1
        //   y entry (int2, one per padded row): sy0 | sy1 << 16,  b0 | b1 << 16
        if (l > 0) {
            const OrbxLevel &S = g.lv[l - 1];
            const double sx_ = 1. / ((double)L.w / S.w), sy_ = 1. / ((double)L.h / S.h);
            std::vector<uint32_t> xe(L.w);
            for (int dx = 0; dx < L.w; dx++) {
                float fx = (float)((dx + 0.5) * sx_ - 0.5);
                int sx = (int)floorf(fx);
                fx -= sx;
                if (sx < 0) { fx = 0; sx = 0; }
                if (sx >= S.w - 1) { fx = 0; sx = S.w - 1; }
                const int a0 = cv_round_f((1.f - fx) * 2048), a1 = cv_round_f(fx * 2048);
                if (a0 + a1 != 2048 || a1 < 0 || a1 > 2048) {     // cannot happen: (1 - fx) and fx * 2048 are exact in float
                    orbx_set_error("level %d: resize coefficients of column %d do not sum to 2048", l, dx);
                    return ORBX_ERR_UNSUPPORTED;
                }
                xe[dx] = (uint32_t)sx | ((uint32_t)a1 << 16);
            }
            while (g.rxt.size() % 4) g.rxt.push_back(0);
            L.rx_off = (int)g.rxt.size();
            for (int px = 0; px < L.pitch; px++) {
                // the columns behind the padded image (up to the 16-byte pitch) are never looked at; they repeat the last real column, so
                // that the group they share with real columns keeps its sources in one window (a zero entry sent every row's last
                // group -- and with it the whole warp -- down the byte-by-byte path: 12 % of the kernel's instructions)
                const int pc = px < L.w + 2 * ORBX_EDGE ? px : L.w + 2 * ORBX_EDGE - 1;
                int ix = pc - ORBX_EDGE;
                if (ix < 0) ix = -ix;
                if (ix >= L.w) ix = 2 * L.w - 2 - ix;
                g.rxt.push_back(xe[ix]);
            }
            for (int px = 0; px < L.pitch; px += 4) {   // group flag: 4 columns whose sources lie in one 12-byte window
                uint32_t *q = &g.rxt[L.rx_off + px];
                int lo = 0xffff, hi = 0;
                for (int b = 0; b < 4; b++) {
                    const int sx = (int)(q[b] & 0xffff);
                    lo = sx < lo ? sx : lo; hi = sx > hi ? sx : hi;
                }
                // bytes d, d+1 (d = sx - smallest sx of the group) must lie in an 8-byte window (pyramid.cu); order is free, so the
                // groups that straddle a reflection point qualify too
                if (hi - lo <= 6) q[0] |= 0x80000000u;
            }
            L.ry_off = (int)g.ryt.size();
            for (int py = 0; py < L.ph; py++) {
                int dy = py - ORBX_EDGE;
                if (dy < 0) dy = -dy;
                if (dy >= L.h) dy = 2 * L.h - 2 - dy;
                float fy = (float)((dy + 0.5) * sy_ - 0.5);
                const int sy = (int)floorf(fy);
                fy -= sy;
                const int b0 = cv_round_f((1.f - fy) * 2048), b1 = cv_round_f(fy * 2048);
                const int sy0 = sy < 0 ? 0 : (sy >= S.h ? S.h - 1 : sy);
                const int sy1 = sy + 1 < 0 ? 0 : (sy + 1 >= S.h ? S.h - 1 : sy + 1);
                g.ryt.push_back(make_int2(sy0 | (sy1 << 16), (b0 & 0xffff) | (b1 << 16)));
            }
        }
        // FAST tiles: runs of cells of one cell row (ORBextractor.cc:785-806 decides which cells exist)
        const int maxBX = L.w - ORBX_BORDER, maxBY = L.h - ORBX_BORDER;
        // up to ~200 detection columns per tile (the kernel packs tile x in 8 bits), cells spread evenly over the tiles of a row
        const int max_per_chunk = L.wcell >= 200 ? 1 : (200 / L.wcell > ORBX_FAST_CELLS ? ORBX_FAST_CELLS : 200 / L.wcell);
        for (int i = 0; i < L.nrows; i++) {
            const int iniY = ORBX_BORDER + i * L.hcell;
            if (iniY >= maxBY - 3) continue;
            int maxY = iniY + L.hcell + 6;
            if (maxY > maxBY) maxY = maxBY;
            const int ch = maxY - iniY;
            if (ch < 7) continue;                          // cv::FAST finds nothing in fewer than 7 rows
            int nvalid = 0, last_cw = 0;
            for (int j = 0; j < L.ncols; j++) {
                const int iniX = ORBX_BORDER + j * L.wcell;
                if (iniX >= maxBX - 6) break;
                int maxX = iniX + L.wcell + 6;
                if (maxX > maxBX) maxX = maxBX;
                if (maxX - iniX < 7) break;
                nvalid = j + 1;
                last_cw = maxX - iniX;
            }
            const int row_chunks = (nvalid + max_per_chunk - 1) / max_per_chunk;
            const int per_chunk = row_chunks ? (nvalid + row_chunks - 1) / row_chunks : 1;
            for (int j0 = 0; j0 < nvalid; j0 += per_chunk) {
                const int j1 = j0 + per_chunk < nvalid ? j0 + per_chunk : nvalid;
                OrbxFastChunk c;
                c.level = (int16_t)l;
                c.ncells = (int16_t)(j1 - j0);
                c.x0 = (int16_t)(ORBX_BORDER + j0 * L.wcell);
                c.y0 = (int16_t)iniY;
                c.wcell = (int16_t)L.wcell;
                c.last_cw = (int16_t)(j1 == nvalid ? last_cw : L.wcell + 6);
                c.tw = (int16_t)((j1 - 1 - j0) * L.wcell + c.last_cw);
                c.th = (int16_t)ch;
                if (c.tw - 6 > 256 || c.th - 6 > 255) {      // the FAST kernel packs tile coordinates in 8 + 8 bits
                    orbx_set_error("level %d: FAST tile %dx%d too large", l, c.tw, c.th);
                    return ORBX_ERR_UNSUPPORTED;
                }
                g.chunks.push_back(c);
                const int tp = (int)align_up(c.tw + 15, 16);     // TMA box: starts at a 16-byte aligned column, width multiple of 16
                if (tp > g.fast_tp) g.fast_tp = tp;
                if (ch > g.fast_th) g.fast_th = ch;
            }
        }
        for (int y0 = 0; y0 < L.h; y0 += ORBX_BLUR_TH)
            for (int x0 = 0; x0 < L.w; x0 += ORBX_BLUR_TW) {
                OrbxBlurTile t = {(int16_t)l, (int16_t)x0, (int16_t)y0, 0};
                g.btiles.push_back(t);
            }
    }
    g.pyr_bytes = pyr;
    g.blur_bytes = blur;
    g.cand_words = cand;
    return ORBX_OK;
}

// TMA descriptors: level l of every frame as a (pitch, padded rows, frames) u8 tensor; one box = one FAST tile.  The
// driver entry point is fetched through the runtime so that the library does not link libcuda.
static orbx_status build_tensor_maps(orbx_extractor *e) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        ORBX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess) {
            orbx_set_error("cuTensorMapEncodeTiled is not available in this driver");
            return ORBX_ERR_UNSUPPORTED;
        }
        encode = (encode_fn)fn;
    }
    if (e->n_chunks == 0) return ORBX_OK;
    if (e->fast_tp > 256 || e->fast_th > 256 || e->fast_tp % 16) {
        orbx_set_error("FAST tile %dx%d does not fit a TMA box", e->fast_tp, e->fast_th);
        return ORBX_ERR_UNSUPPORTED;
    }
    for (int l = 0; l < e->nlevels; l++) {
        const OrbxLevel &L = e->lv[l];
        const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.ph, (cuuint64_t)e->max_batch};
        const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)e->pyr_frame_cap};
        const cuuint32_t box[3] = {(cuuint32_t)e->fast_tp, (cuuint32_t)e->fast_th, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = encode(&e->tmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, e->d_pyr + L.off, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            orbx_set_error("cuTensorMapEncodeTiled failed for level %d (CUresult %d)", l, (int)r);
            return ORBX_ERR_CUDA;
        }
    }
    return ORBX_OK;
}

// make `w x h` the current geometry: tables to the device, buffers grown if this size needs more
static orbx_status configure(orbx_extractor *e, int w, int h) {
    if (w == e->cur_w && h == e->cur_h) return ORBX_OK;
    Geometry g;
    orbx_status st = build_geometry(e, w, h, g);
    if (st != ORBX_OK) return st;
    ORBX_CUDA(cudaDeviceSynchronize());   // nothing may still be reading the old tables
    if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }   // it holds the old geometry and tensor maps
    const size_t mb = (size_t)e->max_batch;
    if (g.pyr_bytes > e->pyr_frame_cap || !e->d_pyr) {
        size_t cap = 0;
        st = grow(&e->d_pyr, &cap, g.pyr_bytes * mb);
        if (st) return st;
        e->pyr_frame_cap = g.pyr_bytes;
    }
    if (g.blur_bytes > e->blur_frame_cap || !e->d_blur) {
        size_t cap = 0;
        st = grow(&e->d_blur, &cap, g.blur_bytes * mb);
        if (st) return st;
        e->blur_frame_cap = g.blur_bytes;
    }
    if (g.cand_words > e->cand_frame_cap || !e->d_cand) {
        size_t cap = 0;
        if ((st = grow(&e->d_cand, &cap, g.cand_words * mb))) return st;
        cap = 0;
        if ((st = grow(&e->d_skey, &cap, g.cand_words * mb))) return st;
        cap = 0;
        if ((st = grow(&e->d_scand, &cap, g.cand_words * mb))) return st;
        e->cand_frame_cap = g.cand_words;
    }
    if ((st = grow(&e->d_rxt, &e->rxt_cap, g.rxt.size() + 4))) return st;
    if ((st = grow(&e->d_ryt, &e->ryt_cap, g.ryt.size() + 1))) return st;
    if ((st = grow(&e->d_lut, &e->lut_cap, g.lut.size() + 1))) return st;
    size_t cap = (size_t)e->chunks_cap;
    if ((st = grow(&e->d_chunks, &cap, g.chunks.size() + 1))) return st;
    e->chunks_cap = (int)cap;
    cap = (size_t)e->btiles_cap;
    if ((st = grow(&e->d_btiles, &cap, g.btiles.size() + 1))) return st;
    e->btiles_cap = (int)cap;
    if (!g.rxt.empty()) ORBX_CUDA(cudaMemcpy(e->d_rxt, g.rxt.data(), sizeof(uint32_t) * g.rxt.size(), cudaMemcpyHostToDevice));
    if (!g.ryt.empty()) ORBX_CUDA(cudaMemcpy(e->d_ryt, g.ryt.data(), sizeof(int2) * g.ryt.size(), cudaMemcpyHostToDevice));
    ORBX_CUDA(cudaMemcpy(e->d_lut, g.lut.data(), sizeof(uint32_t) * g.lut.size(), cudaMemcpyHostToDevice));
    if (!g.chunks.empty())
        ORBX_CUDA(cudaMemcpy(e->d_chunks, g.chunks.data(), sizeof(OrbxFastChunk) * g.chunks.size(), cudaMemcpyHostToDevice));
    ORBX_CUDA(cudaMemcpy(e->d_btiles, g.btiles.data(), sizeof(OrbxBlurTile) * g.btiles.size(), cudaMemcpyHostToDevice));
    memcpy(e->lv, g.lv, sizeof(g.lv));
    ORBX_CUDA(cudaMemcpy(e->d_lv, e->lv, sizeof(OrbxLevel) * ORBX_MAX_LEVELS, cudaMemcpyHostToDevice));
    e->n_chunks = (int)g.chunks.size();
    e->n_btiles = (int)g.btiles.size();
    e->fast_tp = g.fast_tp;
    e->fast_th = g.fast_th;
    if ((st = orbx_fast_init(orbx_fast_smem_bytes(e->fast_tp, e->fast_th)))) return st;
    if ((st = build_tensor_maps(e))) return st;
    e->cur_w = w;
    e->cur_h = h;
    return ORBX_OK;
}

extern "C" orbx_status orbx_extractor_create(orbx_extractor **out, int nfeatures, float scale_factor, int nlevels,
                                             int ini_th_fast, int min_th_fast, int max_width, int max_height, int max_batch,
                                             int device) {
    if (!out) return ORBX_ERR_INVALID;
    *out = nullptr;
    if (nfeatures < 0 || nlevels < 1 || nlevels > ORBX_MAX_LEVELS || !(scale_factor > 1.0f) || ini_th_fast < 1 ||
        ini_th_fast > 254 || min_th_fast < 1 || min_th_fast > 254 || max_width < 1 || max_height < 1 || max_batch < 1) {
        orbx_set_error("orbx_extractor_create: bad argument");
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
    orbx_extractor *e = (orbx_extractor *)calloc(1, sizeof(orbx_extractor));
    if (!e) return ORBX_ERR_NOMEM;
    e->device = device;
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->ini_th = ini_th_fast; e->min_th = min_th_fast;
    e->max_w = max_width; e->max_h = max_height; e->max_batch = max_batch;
    // ORBextractor.cc:414-446 (scaleFactor is a double member initialised from the float argument)
    e->scale_factor = (double)scale_factor;
    e->scale[0] = 1.0f; e->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        e->scale[i] = (float)(e->scale[i - 1] * e->scale_factor);
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    for (int i = 0; i < nlevels; i++) {
        e->inv_scale[i] = 1.0f / e->scale[i];
        e->inv_sigma2[i] = 1.0f / e->sigma2[i];
    }
    const float factor = (float)(1.0f / e->scale_factor);
    float n_desired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        e->quota[l] = cv_round_f(n_desired);
        sum += e->quota[l];
        n_desired *= factor;
    }
    e->quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    e->capacity = 0;
    e->node_cap = 0;
    for (int l = 0; l < nlevels; l++) {
        const int c = e->quota[l] + 3 > 4 * ORBX_NINI_MAX ? e->quota[l] + 3 : 4 * ORBX_NINI_MAX;
        e->capacity += c;
        if (c > e->node_cap) e->node_cap = c;
    }
    e->oct_smem = orbx_octree_smem_bytes(e->node_cap);
    orbx_status st = ORBX_OK;
    do {
        if (e->oct_smem > 227 * 1024) {
            orbx_set_error("nfeatures=%d needs %d bytes of shared memory per quadtree CTA", nfeatures, e->oct_smem);
            st = ORBX_ERR_UNSUPPORTED;
            break;
        }
        if ((st = orbx_octree_init(e->oct_smem))) break;
        const size_t mb = (size_t)max_batch;
        cudaError_t ce;
#define TRY(x) if ((ce = (x)) != cudaSuccess) { orbx_set_error("%s -> %s", #x, cudaGetErrorString(ce)); st = ORBX_ERR_CUDA; break; }
        TRY(cudaMalloc((void **)&e->d_lv, sizeof(OrbxLevel) * ORBX_MAX_LEVELS));
        TRY(cudaMalloc((void **)&e->d_ncand, sizeof(int) * ORBX_MAX_LEVELS * mb));
        TRY(cudaMalloc((void **)&e->d_lvl_cnt, sizeof(int) * ORBX_MAX_LEVELS * mb));
        TRY(cudaMalloc((void **)&e->d_lvl_kp, sizeof(uint32_t) * e->capacity * mb));
        TRY(cudaMalloc((void **)&e->d_status, sizeof(int) * mb));
        TRY(cudaMalloc((void **)&e->d_img, (size_t)max_width * max_height * mb));
        TRY(cudaMalloc((void **)&e->d_kps, sizeof(orbx_keypoint) * e->capacity * mb));
        TRY(cudaMalloc((void **)&e->d_desc, (size_t)32 * e->capacity * mb));
        TRY(cudaMalloc((void **)&e->d_counts, sizeof(int32_t) * mb));
        TRY(cudaMallocHost((void **)&e->h_status, sizeof(int) * mb));
        TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        TRY(cudaStreamCreateWithFlags(&e->aux, cudaStreamNonBlocking));
        TRY(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
        TRY(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
#undef TRY
        e->img_cap = (size_t)max_width * max_height;
        st = configure(e, max_width, max_height);
    } while (0);
    if (st != ORBX_OK) {
        orbx_extractor_destroy(e);
        return st;
    }
    *out = e;
    return ORBX_OK;
}

extern "C" void orbx_extractor_destroy(orbx_extractor *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    cudaFree(e->d_lv); cudaFree(e->d_pyr); cudaFree(e->d_blur); cudaFree(e->d_rxt); cudaFree(e->d_ryt); cudaFree(e->d_lut);
    cudaFree(e->d_chunks); cudaFree(e->d_btiles); cudaFree(e->d_cand); cudaFree(e->d_skey); cudaFree(e->d_scand);
    cudaFree(e->d_ncand); cudaFree(e->d_lvl_kp); cudaFree(e->d_lvl_cnt); cudaFree(e->d_status); cudaFree(e->d_img);
    cudaFree(e->d_kps); cudaFree(e->d_desc); cudaFree(e->d_counts);
    for (int i = 0; e->prof_ev && i < e->prof_slots * (ORBX_STAGES + 1); i++) cudaEventDestroy(e->prof_ev[i]);
    free(e->prof_ev);
    if (e->h_status) cudaFreeHost(e->h_status);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->graph) cudaGraphExecDestroy(e->graph);
    if (e->aux) cudaStreamDestroy(e->aux);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    free(e);
}

extern "C" int orbx_extractor_capacity(const orbx_extractor *e) { return e ? e->capacity : 0; }

extern "C" orbx_status orbx_extractor_tables(const orbx_extractor *e, float *scale, float *inv_scale, float *sigma2,
                                             float *inv_sigma2, int32_t *features_per_level) {
    if (!e) return ORBX_ERR_INVALID;
    for (int l = 0; l < e->nlevels; l++) {
        if (scale) scale[l] = e->scale[l];
        if (inv_scale) inv_scale[l] = e->inv_scale[l];
        if (sigma2) sigma2[l] = e->sigma2[l];
        if (inv_sigma2) inv_sigma2[l] = e->inv_sigma2[l];
        if (features_per_level) features_per_level[l] = e->quota[l];
    }
    return ORBX_OK;
}

extern "C" orbx_status orbx_extractor_run_device(orbx_extractor *e, const uint8_t *d_images, size_t frame_pitch, int batch,
                                                 int width, int height, int stride, orbx_keypoint *d_kps, uint8_t *d_desc,
                                                 int32_t *d_counts, void *stream) {
    if (!e || batch < 0 || !d_counts) return ORBX_ERR_INVALID;
    if (batch > e->max_batch || width > e->max_w || height > e->max_h) {
        orbx_set_error("batch %d of %dx%d exceeds the handle's %d of %dx%d", batch, width, height, e->max_batch, e->max_w, e->max_h);
        return ORBX_ERR_CAPACITY;
    }
    cudaStream_t s = (cudaStream_t)stream;
    ORBX_CUDA(cudaSetDevice(e->device));
    e->last_launches = 0;
    e->last_batch = batch;
    if (batch == 0) return ORBX_OK;
    if (width <= 0 || height <= 0 || !d_images) {   // `if(_image.empty()) return;` ORBextractor.cc:1046
        ORBX_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * batch, s));
        e->last_batch = 0;
        return ORBX_OK;
    }
    if (stride < width || !d_kps || !d_desc) return ORBX_ERR_INVALID;
    orbx_status st = configure(e, width, height);
    if (st) return st;
    ORBX_CUDA(cudaMemsetAsync(e->d_status, 0, sizeof(int) * batch, s));
    cudaEvent_t *ev = e->prof_ev ? e->prof_ev + (size_t)(e->prof_runs % e->prof_slots) * (ORBX_STAGES + 1) : nullptr;
#define MARK(i) if (ev) ORBX_CUDA(cudaEventRecord(ev[i], s))
    MARK(0);
    if ((st = orbx_launch_pyramid(e, d_images, frame_pitch, batch, stride, s))) return st;
    MARK(1);
    if ((st = orbx_launch_fast(e, batch, s))) return st;
    MARK(2);
    if (ev) {                       // per-stage timing asked for: one stream, stages back to back
        if ((st = orbx_launch_octree(e, batch, s))) return st;
        MARK(3);
        if ((st = orbx_launch_blur(e, batch, s))) return st;
    } else {
        // The quadtree kernel has one CTA per (frame, level) and leaves most SMs idle; the blur only needs the pyramid.
        // Fork after FAST: quadtree on the caller's stream, blur on the side stream, join before the descriptors.
        ORBX_CUDA(cudaEventRecord(e->ev_fork, s));
        ORBX_CUDA(cudaStreamWaitEvent(e->aux, e->ev_fork, 0));
        if ((st = orbx_launch_octree(e, batch, s))) return st;
        if ((st = orbx_launch_blur(e, batch, e->aux))) return st;
        ORBX_CUDA(cudaEventRecord(e->ev_join, e->aux));
        ORBX_CUDA(cudaStreamWaitEvent(s, e->ev_join, 0));
    }
    MARK(4);
    if ((st = orbx_launch_describe(e, batch, d_kps, d_desc, d_counts, s))) return st;
    MARK(5);
#undef MARK
    if (ev) e->prof_runs++;
    return ORBX_OK;
}

static orbx_status check_status(orbx_extractor *e, int batch, cudaStream_t s) {
    ORBX_CUDA(cudaMemcpyAsync(e->h_status, e->d_status, sizeof(int) * batch, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    for (int b = 0; b < batch; b++)
        if (e->h_status[b]) {
            orbx_set_error("frame %d: device status 0x%x (1 = candidate list overflow, 2 = quadtree depth, 4 = node overflow)", b,
                           e->h_status[b]);
            return ORBX_ERR_CAPACITY;
        }
    return ORBX_OK;
}

extern "C" orbx_status orbx_extractor_run_host(orbx_extractor *e, const uint8_t *const *images, int batch, int width,
                                               int height, int stride, orbx_keypoint *kps, uint8_t *desc, int32_t *counts) {
    if (!e || batch < 0 || !counts) return ORBX_ERR_INVALID;
    if (batch == 0) return ORBX_OK;
    if (width <= 0 || height <= 0 || !images) {
        for (int b = 0; b < batch; b++) counts[b] = 0;
        e->last_batch = 0;
        e->last_launches = 0;
        return ORBX_OK;
    }
    if (batch > e->max_batch || width > e->max_w || height > e->max_h) {
        orbx_set_error("batch %d of %dx%d exceeds the handle's %d of %dx%d", batch, width, height, e->max_batch, e->max_w, e->max_h);
        return ORBX_ERR_CAPACITY;
    }
    if (stride < width || !kps || !desc) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(e->device));
    cudaStream_t s = e->stream;
    const size_t fp = (size_t)width * height;
    for (int b = 0; b < batch; b++) {
        if (!images[b]) return ORBX_ERR_INVALID;
        ORBX_CUDA(cudaMemcpy2DAsync(e->d_img + fp * b, width, images[b], stride, width, height, cudaMemcpyHostToDevice, s));
    }
    orbx_status st = ORBX_OK;
    static const bool use_graph = getenv("ORBX_NO_GRAPH") == nullptr;
    if (use_graph && !e->prof_ev) {
        // replay the captured stage sequence (same buffers, same geometry); capture it on first use
        if ((st = configure(e, width, height))) return st;
        if (!e->graph || e->graph_w != width || e->graph_h != height || e->graph_batch != batch) {
            if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
            cudaGraph_t g = nullptr;
            ORBX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            st = orbx_extractor_run_device(e, e->d_img, fp, batch, width, height, width, e->d_kps, e->d_desc, e->d_counts, s);
            const cudaError_t ce = cudaStreamEndCapture(s, &g);
            if (st) { if (g) cudaGraphDestroy(g); return st; }
            if (ce != cudaSuccess) { orbx_set_error("stream capture of the extractor failed: %s", cudaGetErrorString(ce)); return ORBX_ERR_CUDA; }
            const cudaError_t ci = cudaGraphInstantiate(&e->graph, g, 0);
            cudaGraphDestroy(g);
            if (ci != cudaSuccess) { e->graph = nullptr; orbx_set_error("cudaGraphInstantiate: %s", cudaGetErrorString(ci)); return ORBX_ERR_CUDA; }
            e->graph_w = width; e->graph_h = height; e->graph_batch = batch;
        }
        const int launches = e->last_launches;           // what the capture counted
        ORBX_CUDA(cudaGraphLaunch(e->graph, s));
        e->last_launches = launches;
        e->last_batch = batch;
    } else {
        st = orbx_extractor_run_device(e, e->d_img, fp, batch, width, height, width, e->d_kps, e->d_desc, e->d_counts, s);
        if (st) return st;
    }
    const size_t cap = (size_t)e->capacity;
    ORBX_CUDA(cudaMemcpyAsync(kps, e->d_kps, sizeof(orbx_keypoint) * cap * batch, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(desc, e->d_desc, 32 * cap * batch, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(counts, e->d_counts, sizeof(int32_t) * batch, cudaMemcpyDeviceToHost, s));
    return check_status(e, batch, s);
}

extern "C" orbx_status orbx_extractor_pyramid(const orbx_extractor *e, int batch_idx, int level, const uint8_t **d_ptr,
                                              int *width, int *height, int *pitch) {
    if (!e || level < 0 || level >= e->nlevels || batch_idx < 0 || batch_idx >= e->max_batch || e->cur_w == 0) return ORBX_ERR_INVALID;
    const OrbxLevel &L = e->lv[level];
    if (d_ptr) *d_ptr = e->d_pyr + (size_t)batch_idx * e->pyr_frame_cap + L.off + (size_t)ORBX_EDGE * L.pitch + ORBX_EDGE;
    if (width) *width = L.w;
    if (height) *height = L.h;
    if (pitch) *pitch = L.pitch;
    return ORBX_OK;
}

extern "C" orbx_status orbx_extractor_pyramid_host(const orbx_extractor *e, int batch_idx, int level, int with_border,
                                                   uint8_t *dst, int dst_stride) {
    if (!e || !dst || level < 0 || level >= e->nlevels || batch_idx < 0 || batch_idx >= e->max_batch || e->cur_w == 0) return ORBX_ERR_INVALID;
    const OrbxLevel &L = e->lv[level];
    const int pad = with_border ? ORBX_EDGE : 0;
    const int w = L.w + 2 * pad, h = L.h + 2 * pad;
    if (dst_stride < w) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(e->device));
    ORBX_CUDA(cudaDeviceSynchronize());
    const uint8_t *src = e->d_pyr + (size_t)batch_idx * e->pyr_frame_cap + L.off + (size_t)(ORBX_EDGE - pad) * L.pitch + (ORBX_EDGE - pad);
    ORBX_CUDA(cudaMemcpy2D(dst, dst_stride, src, L.pitch, w, h, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

extern "C" orbx_status orbx_extractor_blurred_host(const orbx_extractor *e, int batch_idx, int level, uint8_t *dst, int dst_stride) {
    if (!e || !dst || level < 0 || level >= e->nlevels || batch_idx < 0 || batch_idx >= e->max_batch || e->cur_w == 0) return ORBX_ERR_INVALID;
    const OrbxLevel &L = e->lv[level];
    if (dst_stride < L.w) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(e->device));
    ORBX_CUDA(cudaDeviceSynchronize());
    ORBX_CUDA(cudaMemcpy2D(dst, dst_stride, e->d_blur + (size_t)batch_idx * e->blur_frame_cap + L.boff, L.bpitch, L.w, L.h,
                           cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

extern "C" orbx_status orbx_extractor_candidates_host(const orbx_extractor *e, int batch_idx, int level, uint32_t *dst, int cap, int *n) {
    if (!e || !n || level < 0 || level >= e->nlevels || batch_idx < 0 || batch_idx >= e->max_batch || e->cur_w == 0) return ORBX_ERR_INVALID;
    const OrbxLevel &L = e->lv[level];
    ORBX_CUDA(cudaSetDevice(e->device));
    ORBX_CUDA(cudaDeviceSynchronize());
    int cnt = 0;
    ORBX_CUDA(cudaMemcpy(&cnt, e->d_ncand + batch_idx * ORBX_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    int m = cnt < cap ? cnt : cap;
    if (m > L.cand_cap) m = L.cand_cap;
    if (dst && m > 0)
        ORBX_CUDA(cudaMemcpy(dst, e->d_cand + (size_t)batch_idx * e->cand_frame_cap + L.cand_off, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

extern "C" orbx_status orbx_extractor_level_keypoints_host(const orbx_extractor *e, int batch_idx, int level, uint32_t *dst, int cap, int *n) {
    if (!e || !n || level < 0 || level >= e->nlevels || batch_idx < 0 || batch_idx >= e->max_batch || e->cur_w == 0) return ORBX_ERR_INVALID;
    const OrbxLevel &L = e->lv[level];
    ORBX_CUDA(cudaSetDevice(e->device));
    ORBX_CUDA(cudaDeviceSynchronize());
    int cnt = 0;
    ORBX_CUDA(cudaMemcpy(&cnt, e->d_lvl_cnt + batch_idx * ORBX_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    const int m = cnt < cap ? cnt : cap;
    if (dst && m > 0)
        ORBX_CUDA(cudaMemcpy(dst, e->d_lvl_kp + (size_t)batch_idx * e->capacity + L.kp_off, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

extern "C" int orbx_extractor_last_launches(const orbx_extractor *e) { return e ? e->last_launches : 0; }

extern "C" orbx_status orbx_extractor_profile(orbx_extractor *e, int slots) {
    if (!e || slots < 0) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(e->device));
    ORBX_CUDA(cudaDeviceSynchronize());
    for (int i = 0; e->prof_ev && i < e->prof_slots * (ORBX_STAGES + 1); i++) cudaEventDestroy(e->prof_ev[i]);
    free(e->prof_ev);
    e->prof_ev = nullptr;
    e->prof_slots = e->prof_runs = 0;
    if (slots == 0) return ORBX_OK;
    e->prof_ev = (cudaEvent_t *)calloc((size_t)slots * (ORBX_STAGES + 1), sizeof(cudaEvent_t));
    if (!e->prof_ev) return ORBX_ERR_NOMEM;
    e->prof_slots = slots;
    for (int i = 0; i < slots * (ORBX_STAGES + 1); i++) ORBX_CUDA(cudaEventCreate(&e->prof_ev[i]));
    return ORBX_OK;
}

extern "C" orbx_status orbx_extractor_stage_ms(orbx_extractor *e, int *runs, float *ms) {
    if (!e || !runs || !ms || !e->prof_ev) return ORBX_ERR_INVALID;
    ORBX_CUDA(cudaSetDevice(e->device));
    const int n = e->prof_runs < e->prof_slots ? e->prof_runs : e->prof_slots;
    for (int k = 0; k < ORBX_STAGES; k++) ms[k] = 0.f;
    for (int r = 0; r < n; r++) {
        cudaEvent_t *ev = e->prof_ev + (size_t)r * (ORBX_STAGES + 1);
        ORBX_CUDA(cudaEventSynchronize(ev[ORBX_STAGES]));
        for (int k = 0; k < ORBX_STAGES; k++) {
            float t = 0.f;
            ORBX_CUDA(cudaEventElapsedTime(&t, ev[k], ev[k + 1]));
            ms[k] += t;
        }
    }
    *runs = n;
    e->prof_runs = 0;
    return ORBX_OK;
}
