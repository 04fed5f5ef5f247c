// 7x7 Gaussian blur of every pyramid level: replaces GaussianBlur(workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101)
// at reference src/ORBextractor.cc:1085-1086 (OpenCV's 8-bit fixed-point path).
//   taps Q8.8 = {18,34,48,56,48,34,18}; rows first (16-bit sums), then columns; out = (acc + 2^15) >> 16.
// The halo is read from the 19-pixel pad of the pyramid buffer, which already is the REFLECT_101 image of the
// interior (ORBextractor.cc:1122-1128), so no border logic is needed.  One CTA = one 64x32 output tile of one
// level of one frame (tile table built on the host); the result goes to an un-padded copy of the level.
#include "orbx_internal.cuh"

#define BLUR_THREADS 256
#define BLUR_IN_W (ORBX_BLUR_TW + 8)   // 70 bytes needed, padded to 72 (18 words)
#define BLUR_IN_H (ORBX_BLUR_TH + 6)

__global__ void __launch_bounds__(BLUR_THREADS)
k_blur(const uint8_t *__restrict__ pyr, size_t pyr_frame, uint8_t *__restrict__ blur, size_t blur_frame,
       const OrbxLevel *__restrict__ lv, const OrbxBlurTile *__restrict__ tiles) {
    __shared__ __align__(16) uint8_t in[BLUR_IN_H][BLUR_IN_W];
    __shared__ __align__(16) uint16_t hb[BLUR_IN_H][ORBX_BLUR_TW];
    const OrbxBlurTile t = tiles[blockIdx.x];
    const OrbxLevel &L = lv[t.level];
    const int frame = blockIdx.y, tid = threadIdx.x;
    const int pitch = L.pitch, ph = L.ph;
    // padded-buffer position of in[0][0]: column 19 + x0 - 3 (16-byte aligned because x0 % 64 == 0), row 19 + y0 - 3
    const uint8_t *src = pyr + (size_t)frame * pyr_frame + L.off;
    const int gx = ORBX_EDGE + t.x0 - 3, gy = ORBX_EDGE + t.y0 - 3;
    for (int i = tid; i < BLUR_IN_H * (BLUR_IN_W / 4); i += BLUR_THREADS) {
        const int r = i / (BLUR_IN_W / 4), c = i - r * (BLUR_IN_W / 4);
        uint32_t v = 0;
        if (gy + r < ph && gx + 4 * c + 3 < pitch)
            v = __ldg(reinterpret_cast<const uint32_t *>(src + (size_t)(gy + r) * pitch + gx) + c);
        reinterpret_cast<uint32_t *>(&in[r][0])[c] = v;
    }
    __syncthreads();
    // rows: 4 consecutive outputs per thread
    for (int i = tid; i < BLUR_IN_H * (ORBX_BLUR_TW / 4); i += BLUR_THREADS) {
        const int r = i / (ORBX_BLUR_TW / 4), c4 = (i - r * (ORBX_BLUR_TW / 4)) * 4;
        const uint8_t *p = &in[r][c4];
        int v[10];
#pragma unroll
        for (int k = 0; k < 10; k++) v[k] = p[k];
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int acc = 18 * (v[o] + v[o + 6]) + 34 * (v[o + 1] + v[o + 5]) + 48 * (v[o + 2] + v[o + 4]) + 56 * v[o + 3];
            hb[r][c4 + o] = (uint16_t)acc;
        }
    }
    __syncthreads();
    // columns: 4 consecutive outputs per thread, one 32-bit store
    uint8_t *dst = blur + (size_t)frame * blur_frame + L.boff;
    for (int i = tid; i < ORBX_BLUR_TH * (ORBX_BLUR_TW / 4); i += BLUR_THREADS) {
        const int y = i / (ORBX_BLUR_TW / 4), c4 = (i - y * (ORBX_BLUR_TW / 4)) * 4;
        const int oy = t.y0 + y, ox = t.x0 + c4;
        if (oy >= L.h || ox >= L.bpitch) continue;
        uint32_t packed = 0;
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const uint32_t acc = 18u * (hb[y][c4 + o] + hb[y + 6][c4 + o]) + 34u * (hb[y + 1][c4 + o] + hb[y + 5][c4 + o]) +
                                 48u * (hb[y + 2][c4 + o] + hb[y + 4][c4 + o]) + 56u * hb[y + 3][c4 + o];
            packed |= ((acc + 32768u) >> 16) << (8 * o);
        }
        *reinterpret_cast<uint32_t *>(dst + (size_t)oy * L.bpitch + ox) = packed;
    }
}

orbx_status orbx_launch_blur(orbx_extractor *e, int batch, cudaStream_t s) {
    if (e->n_btiles == 0) return ORBX_OK;
    dim3 grid(e->n_btiles, batch);
    k_blur<<<grid, BLUR_THREADS, 0, s>>>(e->d_pyr, e->pyr_frame_cap, e->d_blur, e->blur_frame_cap, e->d_lv, e->d_btiles);
    e->last_launches++;
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}
