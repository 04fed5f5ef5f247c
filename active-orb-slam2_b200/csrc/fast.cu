// Per-cell FAST-9/16 with 3x3 non-max suppression: replaces the cv::FAST loop of
// ORBextractor::ComputeKeyPointsOctTree (reference src/ORBextractor.cc:765-829; OpenCV fast.cpp/fast_score.cpp).
//
// One CTA = a run of up to ORBX_FAST_CELLS cells of one cell-row of one level of one frame.  The tile
// (cells + 3-pixel ring) is staged in shared memory once; scores, suppression and the reference's
// "no corner at iniThFAST -> retry the cell at minThFAST" rule (ORBextractor.cc:809-815) all run from there.
//   score(p)  = max over the 16 arcs of 9 contiguous ring pixels of min|I(ring)-I(p)| (one sign) - 1
//   corner    iff score >= threshold;  kept iff score is strictly greater than its 8 neighbours, where
//               neighbours outside the cell's detection region [3,cw-3)x[3,ch-3) count as 0.
// Survivors are appended to the (frame, level) candidate list as x | y<<12 | score<<24 in the coordinates of
// ORBextractor.cc:822-823 (relative to minBorderX/Y).  Order inside the list is unspecified; the quadtree
// kernel never depends on it.
#include "orbx_internal.cuh"

#define FAST_THREADS 256

// Shared-memory budget beyond the two tile planes.  A tile row pitch tp = align16(tw + 15), so the detection region is at most
// (tp - 21) x (th - 6) pixels.  Survivors of the 3x3 suppression: no two of them are neighbours inside a cell, i.e. at most
// ceil(w/2) x ceil(h/2) per cell region -> (vw/2 + cells) x ceil(vh/2) per tile (the second pass only adds to cells that had none).
// The list of pre-test survivors holds at most every detection pixel.  Sized tightly because this decides how many CTAs fit an SM
// (47.8 KB -> 4 CTAs; 40.6 KB -> 5).
__host__ __device__ inline int orbx_fast_out_words(int tp_max, int th_max) {
    return ((tp_max - 20) / 2 + ORBX_FAST_CELLS) * ((th_max - 5) / 2);
}
__host__ __device__ inline int orbx_fast_list_entries(int tp_max, int th_max) {
    return (tp_max - 21) * (th_max - 6);
}
// the score plane holds detection pixels only: pitch tp - 16 (a multiple of 16, at least 5 more than the widest detection row), th - 6
// rows, plus one zero row above and one below: with the zero spare columns at the end of every row, all eight neighbours of a detection
// pixel can be read without a bounds test
__host__ __device__ inline int orbx_fast_score_bytes(int tp_max, int th_max) {
    return (tp_max - 16) * (th_max - 6 + 2);
}

// The ring is OpenCV's 16-pixel Bresenham circle, (dx,dy) = (0,3),(1,3),(2,2),(3,1),(3,0),(3,-1),(2,-2),(1,-3),
// (0,-3),(-1,-3),(-2,-2),(-3,-1),(-3,0),(-3,1),(-2,2),(-1,3).

// score of the pixel at p (row pitch tp) or 0 if it is not a corner at `th`
__device__ __forceinline__ int fast_score(const uint8_t *p, int tp, int th) {
    const int v = p[0];
    int r[16];
    r[0] = p[3 * tp]; r[4] = p[3]; r[8] = p[-3 * tp]; r[12] = p[-3];
    r[1] = p[3 * tp + 1]; r[2] = p[2 * tp + 2]; r[3] = p[tp + 3];
    r[5] = p[-tp + 3]; r[6] = p[-2 * tp + 2]; r[7] = p[-3 * tp + 1];
    r[9] = p[-3 * tp - 1]; r[10] = p[-2 * tp - 2]; r[11] = p[-tp - 3];
    r[13] = p[tp - 3]; r[14] = p[2 * tp - 2]; r[15] = p[3 * tp - 1];
    // sliding min / max over windows of 9 on the circular ring: 2, 4, 8 by doubling, then +1
    int mn2[16], mx2[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn2[k] = min(r[k], r[(k + 1) & 15]);
        mx2[k] = max(r[k], r[(k + 1) & 15]);
    }
    int mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn4[k] = min(mn2[k], mn2[(k + 2) & 15]);
        mx4[k] = max(mx2[k], mx2[(k + 2) & 15]);
    }
    int lo_of_max = 255, hi_of_min = 0;  // min over arcs of the arc maximum, max over arcs of the arc minimum
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), r[(k + 8) & 15]);
        const int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), r[(k + 8) & 15]);
        hi_of_min = max(hi_of_min, mn9);
        lo_of_max = min(lo_of_max, mx9);
    }
    // darker ring: min(v - ring) = v - arc max; brighter ring: min(ring - v) = arc min - v
    const int best = max(v - lo_of_max, hi_of_min - v);
    return best > th ? best - 1 : 0;
}

// Two pixels at once: their ring values side by side in the 16-bit halves of a word, so that every min / max of the sliding
// 9-window is one VIMNMX.U16x2 (or VIMNMX3.U16x2) for both.  Same arithmetic as fast_score(); returns the scores in sa / sb.
__device__ __forceinline__ void fast_score_pair(const uint8_t *pa, const uint8_t *pb, int tp, int th, int &sa, int &sb) {
    uint32_t r[16];
#define ORBX_RING(k, off) r[k] = __byte_perm((uint32_t)pa[off], (uint32_t)pb[off], 0x5410)
    ORBX_RING(0, 3 * tp); ORBX_RING(1, 3 * tp + 1); ORBX_RING(2, 2 * tp + 2); ORBX_RING(3, tp + 3);
    ORBX_RING(4, 3); ORBX_RING(5, -tp + 3); ORBX_RING(6, -2 * tp + 2); ORBX_RING(7, -3 * tp + 1);
    ORBX_RING(8, -3 * tp); ORBX_RING(9, -3 * tp - 1); ORBX_RING(10, -2 * tp - 2); ORBX_RING(11, -tp - 3);
    ORBX_RING(12, -3); ORBX_RING(13, tp - 3); ORBX_RING(14, 2 * tp - 2); ORBX_RING(15, 3 * tp - 1);
#undef ORBX_RING
    uint32_t mn2[16], mx2[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn2[k] = __vminu2(r[k], r[(k + 1) & 15]);
        mx2[k] = __vmaxu2(r[k], r[(k + 1) & 15]);
    }
    uint32_t mn4[16], mx4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn4[k] = __vminu2(mn2[k], mn2[(k + 2) & 15]);
        mx4[k] = __vmaxu2(mx2[k], mx2[(k + 2) & 15]);
    }
    uint32_t lo_of_max = 0x00ff00ffu, hi_of_min = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const uint32_t mn9a = __vimin3_u16x2(mn4[k], mn4[(k + 4) & 15], r[(k + 8) & 15]);
        const uint32_t mn9b = __vimin3_u16x2(mn4[k + 1], mn4[(k + 5) & 15], r[(k + 9) & 15]);
        const uint32_t mx9a = __vimax3_u16x2(mx4[k], mx4[(k + 4) & 15], r[(k + 8) & 15]);
        const uint32_t mx9b = __vimax3_u16x2(mx4[k + 1], mx4[(k + 5) & 15], r[(k + 9) & 15]);
        hi_of_min = __vimax3_u16x2(hi_of_min, mn9a, mn9b);
        lo_of_max = __vimin3_u16x2(lo_of_max, mx9a, mx9b);
    }
    const int va = pa[0], vb = pb[0];
    const int ba = max(va - (int)(lo_of_max & 0xffffu), (int)(hi_of_min & 0xffffu) - va);
    const int bb = max(vb - (int)(lo_of_max >> 16), (int)(hi_of_min >> 16) - vb);
    sa = ba > th ? ba - 1 : 0;
    sb = bb > th ? bb - 1 : 0;
}

struct FastShared {
    int cnt[ORBX_FAST_CELLS];     // survivors per cell in the current pass
    int empty[ORBX_FAST_CELLS];
    int n_out;                    // entries in the output list
    int n_list;                   // pixels that passed the compass test in the current pass
    int any_empty;
    int base;                     // reserved start in the global candidate list
    uint8_t cellof[256];          // detection column -> cell of the chunk
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// dynamic smem layout: [tile bytes th x tp][score bytes (th - 6) x (tp - 16)][out words][list u16]
// The tile arrives by TMA: one cp.async.bulk.tensor of the (tp x th_max) box of the level's tensor map whose first column
// is (19 + x0) rounded down to 16 (TMA wants 16-byte aligned row starts) and first row 19 + y0, completion on an mbarrier;
// rows / columns past the level read as zero and are never looked at.
// Three phases per pass, so that only the few pixels that can be corners pay for the 16-arc score and the 3x3
// suppression: (1) packed pre-test on every pixel (four per thread), survivors compacted into a list; (2) score of the
// listed pixels; (3) suppression of the listed pixels with a non-zero score.
__global__ void __launch_bounds__(FAST_THREADS, 5)
k_fast(const uint8_t *__restrict__ pyr, size_t pyr_frame, const OrbxLevel *__restrict__ lv,
       const OrbxFastChunk *__restrict__ chunks, uint32_t *__restrict__ cand, size_t cand_frame, int *__restrict__ ncand,
       int *__restrict__ status, int ini_th, int min_th, int tp_max, int th_max, const __grid_constant__ OrbxTmaps maps) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ FastShared sh;
    __shared__ __align__(8) uint64_t mbar;
    const OrbxFastChunk ck = chunks[blockIdx.x];
    const int frame = blockIdx.y;
    const OrbxLevel &L = lv[ck.level];
    const int tw = ck.tw, th = ck.th;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- stage the tile by TMA ---------------------------------------------------------------------------
    const int tp = tp_max;                             // smem row pitch = box width (multiple of 16)
    const int gx0 = ORBX_EDGE + ck.x0;                 // padded-buffer column of the tile origin
    const int shift = gx0 & 15;                        // the box starts at the aligned column gx0 - shift
    uint8_t *tile = smem;
    uint8_t *score = smem + (size_t)tp * th_max;
    const int sp = tp - 16;                            // pitch of the score plane
    uint32_t *out = reinterpret_cast<uint32_t *>(score + orbx_fast_score_bytes(tp_max, th_max));
    uint16_t *list = reinterpret_cast<uint16_t *>(out + orbx_fast_out_words(tp_max, th_max));
    const uint32_t mbar_a = smem_u32(&mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)(tp * th_max);
        const uint64_t tmap = reinterpret_cast<uint64_t>(&maps.m[ck.level]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(tile)), "l"(tmap), "r"(gx0 - shift), "r"(ORBX_EDGE + (int)ck.y0), "r"(frame), "r"(mbar_a)
                     : "memory");
    }
    // the score plane starts at zero (16 bytes per store; the plane is 16-byte aligned and a multiple of 4 bytes long, the few bytes
    // a rounded-up last store reaches into the survivor words behind it are written later)
    for (int i = tid; i < (sp * (th - 6 + 2) + 15) >> 4; i += FAST_THREADS) reinterpret_cast<uint4 *>(score)[i] = make_uint4(0u, 0u, 0u, 0u);
    const int vw = tw - 6, vh = th - 6;                // detection region
    const int wcell = ck.wcell;
    {   // detection column -> cell: x / wcell by a 16.16 reciprocal (x < 256, wcell <= 255: exact), instead of an integer division
        const uint32_t cmagic = (65536u + (uint32_t)wcell - 1u) / (uint32_t)wcell;
        for (int x = tid; x < vw; x += FAST_THREADS) sh.cellof[x] = (uint8_t)(((uint32_t)x * cmagic) >> 16);
    }
    if (tid < ORBX_FAST_CELLS) { sh.cnt[tid] = 0; sh.empty[tid] = 1; }
    if (tid == 0) { sh.n_out = 0; sh.any_empty = 0; sh.n_list = 0; }
    __syncthreads();

    // everything above overlapped the copy; now wait for the tile (phase 0 of the barrier)
    // (one warp polls the barrier, the others sleep at the block barrier behind it: the polls of eight warps were 4 % of the kernel's
    //  instructions)
    if (warp == 0)
        asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                     ::"r"(mbar_a) : "memory");
    __syncthreads();
    const uint8_t *t0 = tile + shift + 3 * tp + 3;     // detection pixel (x,y) = t0[y*tp + x]
    uint8_t *s0 = score + sp;                          // score of detection pixel (x,y) = s0[y*sp + x]; rows -1 and vh stay zero

    // words of a tile row that hold detection pixels, and the magic number that divides an item index by their count
    const int w_first = (shift + 3) >> 2, n_words = ((shift + 3 + vw - 1) >> 2) - w_first + 1;
    const int n_items = vw > 0 && vh > 0 ? n_words * vh : 0;
    const uint32_t w_magic = (uint32_t)((0x100000000ull + n_words - 1) / n_words);

    for (int pass = 0; pass < 2; pass++) {
        const int thr = pass == 0 ? ini_th : min_th;
        // (1) compass pre-test on four pixels per thread (one aligned word of the tile row), survivors -> list (x | y << 8).
        //     |ring - centre| > thr, polarity ignored, as packed bytes: d = VABSDIFF4, then bit 7 of ((d & 0x7f) + 127 - thr) | d.
        //     A 9-arc holds two neighbouring pixels of {0,4,8,12} and two of {2,6,10,14}; opposite pixels OR-ed, the two
        //     axes AND-ed.  It only has to be necessary: fast_score() decides.
        const uint32_t kadd = thr <= 126 ? (uint32_t)(127 - thr) * 0x01010101u : 0u;
        for (int it = warp * 32; it < n_items; it += FAST_THREADS) {
            const int t = it + lane;
            uint32_t m = 0;
            int y = 0, x0 = 0;
            if (t < n_items) {
                y = n_words == 1 ? t : (int)__umulhi((uint32_t)t, w_magic);
                const int wi = w_first + (t - y * n_words);
                x0 = 4 * wi - (shift + 3);                      // detection column of byte 0 of the word
                bool go = true;
                if (pass == 1) go = (sh.empty[sh.cellof[max(x0, 0)]] | sh.empty[sh.cellof[min(x0 + 3, vw - 1)]]) != 0;
                if (go) {
                    const uint32_t *rw = reinterpret_cast<const uint32_t *>(tile + (y + 3) * tp) + wi;
                    const int tpw = tp >> 2;
                    const uint32_t v = rw[0], wl = rw[-1], wr = rw[1];
                    const uint32_t d0 = __vabsdiffu4(rw[3 * tpw], v), d8 = __vabsdiffu4(rw[-3 * tpw], v);
                    const uint32_t d4 = __vabsdiffu4(__byte_perm(v, wr, 0x6543), v), d12 = __vabsdiffu4(__byte_perm(wl, v, 0x4321), v);
                    const uint32_t a = ((d0 & 0x7f7f7f7fu) + kadd) | d0 | ((d8 & 0x7f7f7f7fu) + kadd) | d8;
                    const uint32_t c = ((d4 & 0x7f7f7f7fu) + kadd) | d4 | ((d12 & 0x7f7f7f7fu) + kadd) | d12;
                    m = a & c & 0x80808080u;
                    if (m) {
                        const uint32_t *up = rw + 2 * tpw, *dn = rw - 2 * tpw;
                        const uint32_t e2 = __vabsdiffu4(__byte_perm(up[0], up[1], 0x5432), v);     // ring 2  (+2, +2)
                        const uint32_t e14 = __vabsdiffu4(__byte_perm(up[-1], up[0], 0x5432), v);   // ring 14 (-2, +2)
                        const uint32_t e6 = __vabsdiffu4(__byte_perm(dn[0], dn[1], 0x5432), v);     // ring 6  (+2, -2)
                        const uint32_t e10 = __vabsdiffu4(__byte_perm(dn[-1], dn[0], 0x5432), v);   // ring 10 (-2, -2)
                        const uint32_t f = ((e2 & 0x7f7f7f7fu) + kadd) | e2 | ((e10 & 0x7f7f7f7fu) + kadd) | e10;
                        const uint32_t g = ((e6 & 0x7f7f7f7fu) + kadd) | e6 | ((e14 & 0x7f7f7f7fu) + kadd) | e14;
                        m &= f & g;
                    }
                }
            }
            // keep the bytes that are detection pixels (and, in the second pass, lie in a cell that found nothing)
            if (m) {
                const int lo = max(-x0, 0), hi = min(vw - x0, 4);           // valid bytes of the word: [lo, hi)
                m &= (0xffffffffu << (8 * lo)) & (0xffffffffu >> (8 * (4 - hi)));
                if (pass == 1) {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (((m >> (8 * k + 7)) & 1u) && !sh.empty[sh.cellof[x0 + k]]) m &= ~(0x80u << (8 * k));
                }
            }
            if (__ballot_sync(0xffffffffu, m != 0) == 0) continue;
            const int c = __popc(m);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int base = 0;
            if (lane == 31) base = atomicAdd(&sh.n_list, incl);
            base = __shfl_sync(0xffffffffu, base, 31) + incl - c;
            const int yy = y << 8;
            while (m) {
                const int k = (__ffs(m) - 1) >> 3;
                list[base++] = (uint16_t)((x0 + k) | yy);
                m &= m - 1;
            }
        }
        __syncthreads();
        const int n_list = sh.n_list;
        // (2) scores of the listed pixels, two per thread (their rings packed as u16x2)
        for (int i = 2 * tid; i < n_list; i += 2 * FAST_THREADS) {
            const int pa = list[i], pb = list[min(i + 1, n_list - 1)];
            const int oa = (pa >> 8) * tp + (pa & 0xff), ob = (pb >> 8) * tp + (pb & 0xff);
            int sa, sb;
            fast_score_pair(t0 + oa, t0 + ob, tp, thr, sa, sb);
            s0[(pa >> 8) * sp + (pa & 0xff)] = (uint8_t)sa;
            s0[(pb >> 8) * sp + (pb & 0xff)] = (uint8_t)sb;
        }
        __syncthreads();
        // (3) 3x3 non-max suppression inside the cell's detection region
        for (int i = tid; i < n_list; i += FAST_THREADS) {
            const int pos = list[i], x = pos & 0xff, y = pos >> 8;
            const uint8_t *sq = s0 + y * sp + x;
            const int s = sq[0];
            if (s == 0) continue;
            const int cell = sh.cellof[x];
            const int cx0 = cell * wcell;                                   // first detection column of the cell
            const int cx1 = cell == ck.ncells - 1 ? vw : cx0 + wcell;      // one past the last
            // neighbours outside the cell's detection region count as 0: above / below the plane has zero rows, past the row's end
            // zero spare columns (which also serve x = 0 of the next row); only the cell's own left / right edge needs a test
            const bool l = x > cx0, r = x + 1 < cx1;
            const int lm = max(max((int)sq[-1], (int)sq[-sp - 1]), (int)sq[sp - 1]);
            const int rm = max(max((int)sq[1], (int)sq[-sp + 1]), (int)sq[sp + 1]);
            const int m = max(max((int)sq[-sp], (int)sq[sp]), max(l ? lm : 0, r ? rm : 0));
            const bool keep = s > m;
            if (keep) {
                atomicAdd(&sh.cnt[cell], 1);
                const int o = atomicAdd(&sh.n_out, 1);
                // coordinates relative to minBorder: tile origin x0-16 plus the pixel's tile position
                out[o] = (uint32_t)(ck.x0 - ORBX_BORDER + x + 3) | ((uint32_t)(ck.y0 - ORBX_BORDER + y + 3) << 12) |
                         ((uint32_t)s << 24);
            }
        }
        __syncthreads();
        if (pass == 0) {
            if (tid < ck.ncells) {
                const int em = sh.cnt[tid] == 0;
                sh.empty[tid] = em;
                if (em) sh.any_empty = 1;
            }
            if (tid == 0) sh.n_list = 0;
            __syncthreads();
            if (!sh.any_empty) break;
        }
    }

    // ---- append to the (frame, level) candidate list ------------------------------------------------
    const int n_out = sh.n_out;
    if (n_out == 0) return;
    if (tid == 0) sh.base = atomicAdd(&ncand[frame * ORBX_MAX_LEVELS + ck.level], n_out);
    __syncthreads();
    const int base = sh.base;
    uint32_t *dst = cand + (size_t)frame * cand_frame + L.cand_off;
    for (int i = tid; i < n_out; i += FAST_THREADS) {
        if (base + i < L.cand_cap) dst[base + i] = out[i];
        else atomicOr(&status[frame], ORBX_ST_CAND_OVERFLOW);
    }
}

// tile + score planes, the survivor words and the list of pre-test survivors (see orbx_fast_out_words)
size_t orbx_fast_smem_bytes(int tp_max, int th_max) {
    return (size_t)tp_max * th_max + orbx_fast_score_bytes(tp_max, th_max) + sizeof(uint32_t) * orbx_fast_out_words(tp_max, th_max) +
           sizeof(uint16_t) * orbx_fast_list_entries(tp_max, th_max);
}

orbx_status orbx_launch_fast(orbx_extractor *e, int batch, cudaStream_t s) {
    ORBX_CUDA(cudaMemsetAsync(e->d_ncand, 0, sizeof(int) * ORBX_MAX_LEVELS * batch, s));
    if (e->n_chunks == 0) return ORBX_OK;
    const int tp_max = e->fast_tp, th_max = e->fast_th;
    const size_t smem = orbx_fast_smem_bytes(tp_max, th_max);
    dim3 grid(e->n_chunks, batch);
    k_fast<<<grid, FAST_THREADS, smem, s>>>(e->d_pyr, e->pyr_frame_cap, e->d_lv, e->d_chunks, e->d_cand, e->cand_frame_cap,
                                            e->d_ncand, e->d_status, e->ini_th, e->min_th, tp_max, th_max, e->tmaps);
    e->last_launches++;
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

orbx_status orbx_fast_init(size_t smem_bytes) {
    if (smem_bytes > 227 * 1024) {
        orbx_set_error("FAST tile needs %zu bytes of shared memory", smem_bytes);
        return ORBX_ERR_UNSUPPORTED;
    }
    ORBX_CUDA(ORBX_RAISE_SMEM(k_fast));
    return ORBX_OK;
}
