// Internal declarations shared by the CUDA translation units of liborbx.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/orbx.h"

#define ORBX_MAX_LEVELS 16
#define ORBX_EDGE 19          // EDGE_THRESHOLD, ORBextractor.cc:73
#define ORBX_BORDER 16        // minBorderX = EDGE_THRESHOLD-3, ORBextractor.cc:773
#define ORBX_HALF_PATCH 15
#define ORBX_MAX_DIM 4096     // 12-bit packed candidate coordinates
#define ORBX_NINI_MAX 16      // initial quadtree nodes per level (image aspect up to 16:1)
#define ORBX_OCT_CELLS 4096   // counting-sort cells of the quadtree kernel
#define ORBX_FAST_CELLS 8     // max cells per FAST tile
// opt-in dynamic shared memory ceiling of sm_100 (227 KB).  The attribute is per kernel, not per handle, so it is
// always raised to the ceiling: a second handle with smaller needs must not lower it under a live one.
#define ORBX_SMEM_OPTIN (227 * 1024)
// raise the kernel's dynamic shared memory limit to everything the device allows next to its static part
#define ORBX_RAISE_SMEM(kernel) orbx_raise_smem((const void *)(kernel))
cudaError_t orbx_raise_smem(const void *kernel);

void orbx_set_error(const char *fmt, ...);

#define ORBX_CUDA(call)                                                                           \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            orbx_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return ORBX_ERR_CUDA;                                                                 \
        }                                                                                         \
    } while (0)

// device-side status bits (d_status[frame])
#define ORBX_ST_CAND_OVERFLOW 1
#define ORBX_ST_OCT_DEPTH 2
#define ORBX_ST_NODE_OVERFLOW 4

// ---- per-level geometry, resident in device memory (one array of nlevels entries) -----------------
struct OrbxLevel {
    int w, h;            // interior size (cvRound(cols*invScale), ORBextractor.cc:1112)
    int pitch;           // bytes between rows of the padded buffer (multiple of 16)
    int ph;              // padded height = h + 2*19
    size_t off;          // byte offset of the padded buffer inside one frame's pyramid block (256-aligned)
    int bpitch;          // row pitch of the blurred copy (interior only, multiple of 16)
    size_t boff;         // byte offset of the blurred copy inside one frame's blur block
    float scale;         // mvScaleFactor[level]
    float kp_size;       // (float)(int)(31*scale), ORBextractor.cc:837
    int quota;           // mnFeaturesPerLevel[level]
    int kp_cap;          // capacity of this level's keypoint slots: max(quota+3, 4*n_ini)
    int kp_off;          // first slot of this level inside a frame's slot array
    // FAST cell grid (ORBextractor.cc:773-806)
    int bw, bh;          // maxBorderX-minBorderX, maxBorderY-minBorderY
    int ncols, nrows, wcell, hcell;
    int cand_cap;        // worst-case number of FAST candidates of this level
    size_t cand_off;     // offset (in words) inside one frame's candidate block
    // quadtree (DistributeOctTree) tables
    int n_ini;           // nIni
    int depth;           // splits per axis after which every coordinate has its own path
    int cshift;          // key >> cshift == counting-sort cell
    int ncells;          // n_ini << 2*d0, <= ORBX_OCT_CELLS
    int lutx_off, luty_off;  // offsets into the u32 path LUT
    // resize tables (offsets into the u32 column table and the int2 row table)
    int rx_off, ry_off;
};

// One CTA of the FAST kernel: a run of cells of one cell-row of one level
struct OrbxFastChunk {
    int16_t level;
    int16_t ncells;      // cells in this chunk
    int16_t x0, y0;      // interior coordinates of the tile origin (16 + j0*wCell, 16 + i*hCell)
    int16_t tw, th;      // tile size (detection region = tile minus 3 on every side)
    int16_t wcell;       // cell pitch
    int16_t last_cw;     // width of the last cell of the chunk (the others are wcell+6)
};

// One CTA of the blur kernel: a BLUR_TW x BLUR_TH output tile of one level
struct OrbxBlurTile {
    int16_t level, x0, y0, pad;
};
#define ORBX_BLUR_TW 64
#define ORBX_BLUR_TH 32

// TMA descriptors of the pyramid levels: 3-D u8 tensors (padded column, padded row, frame), one box = one FAST tile
struct OrbxTmaps {
    CUtensorMap m[ORBX_MAX_LEVELS];
};

struct orbx_extractor {
    int device;
    int nfeatures, nlevels, ini_th, min_th;
    double scale_factor;
    float scale[ORBX_MAX_LEVELS], inv_scale[ORBX_MAX_LEVELS], sigma2[ORBX_MAX_LEVELS], inv_sigma2[ORBX_MAX_LEVELS];
    int quota[ORBX_MAX_LEVELS];
    int capacity;             // keypoint slots per frame (sum of kp_cap)
    int max_w, max_h, max_batch;
    // geometry of the currently configured image size
    int cur_w, cur_h;
    OrbxLevel lv[ORBX_MAX_LEVELS];
    OrbxLevel *d_lv;
    size_t pyr_frame_cap;     // allocated bytes per frame (for max_w x max_h)
    uint8_t *d_pyr;           // [max_batch][pyr_frame_cap]
    size_t blur_frame_cap;
    uint8_t *d_blur;          // [max_batch][blur_frame_cap]
    uint32_t *d_rxt;  size_t rxt_cap;   // resize tables per padded output column / row (see extractor.cu)
    int2 *d_ryt;  size_t ryt_cap;
    uint32_t *d_lut;  size_t lut_cap;    // quadtree path LUTs
    OrbxFastChunk *d_chunks;  int n_chunks, chunks_cap;
    int fast_tp, fast_th;     // smem tile pitch / rows of the FAST kernel (max over chunks) = TMA box
    OrbxTmaps tmaps;          // rebuilt by configure()
    OrbxBlurTile *d_btiles;   int n_btiles, btiles_cap;
    uint32_t *d_cand;         // [max_batch][cand_frame_cap]  unsorted candidates (x | y<<12 | score<<24)
    uint32_t *d_skey;         // quadtree keys, sorted
    uint32_t *d_scand;        // candidates in key order
    size_t cand_frame_cap;
    int *d_ncand;             // [max_batch][ORBX_MAX_LEVELS]
    uint32_t *d_lvl_kp;       // [max_batch][capacity]  selected keypoints per level, list order
    int *d_lvl_cnt;           // [max_batch][ORBX_MAX_LEVELS]
    int node_cap;             // nodes per quadtree CTA
    int oct_smem;
    int *d_status;            // device-side error flags [max_batch]
    // host staging for the _host entry point (pinned)
    uint8_t *h_img;  uint8_t *d_img;  size_t img_cap;
    orbx_keypoint *h_kps, *d_kps;
    uint8_t *h_desc, *d_desc;
    int32_t *h_counts, *d_counts;
    int *h_status;
    cudaStream_t stream;      // private stream of the _host entry point
    cudaStream_t aux;         // side stream: the blur runs next to the (under-filled) quadtree kernel
    cudaEvent_t ev_fork, ev_join;
    // CUDA graph of the whole stage sequence on the handle's own buffers (the _host entry point: a dozen small launches per
    // frame are launch-bound at batch 1); valid for one (width, height, batch), rebuilt when any of them changes
    cudaGraphExec_t graph;
    int graph_w, graph_h, graph_batch;
    cudaEvent_t ev_h2d[2];
    int last_launches;
    int last_batch;
    // optional per-stage timing (orbx_extractor_profile): events [prof_slots][ORBX_STAGES + 1] on the launching stream
    cudaEvent_t *prof_ev;
    int prof_slots, prof_runs;
};
#define ORBX_STAGES 5   // pyramid, fast, quadtree, blur, describe

// ---- kernel launchers (one per stage) ---------------------------------------------------------------
orbx_status orbx_launch_pyramid(orbx_extractor *e, const uint8_t *d_images, size_t frame_pitch, int batch,
                                int stride, cudaStream_t s);
orbx_status orbx_launch_fast(orbx_extractor *e, int batch, cudaStream_t s);
orbx_status orbx_launch_octree(orbx_extractor *e, int batch, cudaStream_t s);
orbx_status orbx_launch_blur(orbx_extractor *e, int batch, cudaStream_t s);
orbx_status orbx_launch_describe(orbx_extractor *e, int batch, orbx_keypoint *d_kps, uint8_t *d_desc,
                                 int32_t *d_counts, cudaStream_t s);
int orbx_octree_smem_bytes(int node_cap);
size_t orbx_fast_smem_bytes(int tp_max, int th_max);
orbx_status orbx_kernels_init(orbx_extractor *e);  // cudaFuncSetAttribute for the dynamic-smem kernels
