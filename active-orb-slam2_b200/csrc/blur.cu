// 7x7 Gaussian blur of every pyramid level: replaces GaussianBlur(workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101)
// at reference src/ORBextractor.cc:1085-1086 (OpenCV's 8-bit fixed-point path).
//   taps Q8.8 = {18,34,48,56,48,34,18}; rows first (16-bit sums), then columns; out = (acc + 2^15) >> 16.
// The halo is read from the 19-pixel pad of the pyramid buffer, which already is the REFLECT_101 image of the
// interior (ORBextractor.cc:1122-1128), so no border logic is needed.  One CTA = one 64x32 output tile of one
// level of one frame (tile table built on the host); the result goes to an un-padded copy of the level.
//
// Both passes are integer dot products: a row sum is two DP4A (4 + 3 taps on bytes picked from the row's words by funnel
// shifts); the row sums are stored TRANSPOSED as u16, so that a column sum is four DP2A on (u16, u16) pairs of vertical
// neighbours, and two vertically adjacent outputs share their four words.  Shared-memory pitches (19 and 21 words) are odd,
// which keeps every access of the three phases free of bank conflicts.
#include "orbx_internal.cuh"

#define BLUR_THREADS 256
#define BLUR_IN_H (ORBX_BLUR_TH + 6)            // 38 input rows
#define BLUR_IN_WORDS 18                        // 70 bytes needed per row, 72 loaded
#define BLUR_IN_PITCH 19                        // words
#define BLUR_HT_PITCH 42                        // u16 per column of the transposed row sums (38 used)

__global__ void __launch_bounds__(BLUR_THREADS)
k_blur(const uint8_t *__restrict__ pyr, size_t pyr_frame, uint8_t *__restrict__ blur, size_t blur_frame,
       const OrbxLevel *__restrict__ lv, const OrbxBlurTile *__restrict__ tiles) {
    __shared__ __align__(16) uint32_t in[BLUR_IN_H * BLUR_IN_PITCH];
    __shared__ __align__(16) uint16_t ht[ORBX_BLUR_TW * BLUR_HT_PITCH];
    __shared__ __align__(16) uint8_t ob[ORBX_BLUR_TH][ORBX_BLUR_TW];
    const OrbxBlurTile t = tiles[blockIdx.x];
    const OrbxLevel &L = lv[t.level];
    const int frame = blockIdx.y, tid = threadIdx.x;
    const int pitch = L.pitch, ph = L.ph;
    // padded-buffer position of the first input byte: column 19 + x0 - 3 (16-byte aligned because x0 % 64 == 0), row 19 + y0 - 3
    const uint8_t *src = pyr + (size_t)frame * pyr_frame + L.off;
    const int gx = ORBX_EDGE + t.x0 - 3, gy = ORBX_EDGE + t.y0 - 3;
    // 38 rows x 72 bytes as 16-byte loads: the first input column 19 + x0 - 3 is 16-byte aligned (x0 % 64 == 0), the row pitch a
    // multiple of 16, so a 16-byte group that starts inside a row lies inside it; 4.5 groups per row -> 5 (the last one half used)
    if (tid < BLUR_IN_H * 5) {
        const int r = tid / 5, c = tid - r * 5;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (gy + r < ph && gx + 16 * c < pitch) v = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)(gy + r) * pitch + gx) + c);
        uint32_t *d = in + r * BLUR_IN_PITCH + 4 * c;
        d[0] = v.x; d[1] = v.y;
        if (c < 4) { d[2] = v.z; d[3] = v.w; }
    }
    __syncthreads();
    // rows: thread = (input row r, quarter qx of the 64 columns), 16 sums; lanes run over r
    if (tid < BLUR_IN_H * 4) {
        const int qx = tid / BLUR_IN_H, r = tid - qx * BLUR_IN_H;
        const uint32_t *p = in + r * BLUR_IN_PITCH + 4 * qx;
        uint32_t w[6];
#pragma unroll
        for (int k = 0; k < 6; k++) w[k] = p[k];
        uint16_t *dst = ht + (16 * qx) * BLUR_HT_PITCH + r;
        const uint32_t T0 = 0x38302212u, T1 = 0x00122230u;      // bytes {18,34,48,56}, {48,34,18,0}
#pragma unroll
        for (int o = 0; o < 16; o++) {
            const int wq = o >> 2, sh = 8 * (o & 3);
            const uint32_t lo4 = __funnelshift_r(w[wq], w[wq + 1], sh), hi4 = __funnelshift_r(w[wq + 1], w[wq + 2], sh);
            dst[o * BLUR_HT_PITCH] = (uint16_t)__dp4a(lo4, T0, __dp4a(hi4, T1, 0u));
        }
    }
    __syncthreads();
    // columns: thread = (column x, quarter yq of the 32 rows), 8 outputs from 7 words of the transposed sums; lanes run over x
    {
        const int yq = tid >> 6, x = tid & 63;
        const uint32_t *p = reinterpret_cast<const uint32_t *>(ht + x * BLUR_HT_PITCH + 8 * yq);
        uint32_t w[7];
#pragma unroll
        for (int k = 0; k < 7; k++) w[k] = p[k];
        const uint32_t T0 = 0x38302212u, T1 = 0x00122230u;      // even row of a pair: (18,34 | 48,56), (48,34 | 18,0)
        const uint32_t C0 = 0x30221200u, C1 = 0x12223038u;      // odd row of a pair:  (0,18 | 34,48), (56,48 | 34,18)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t a = __dp2a_lo(w[j], T0, __dp2a_hi(w[j + 1], T0, __dp2a_lo(w[j + 2], T1, __dp2a_hi(w[j + 3], T1, 0u))));
            const uint32_t b = __dp2a_lo(w[j], C0, __dp2a_hi(w[j + 1], C0, __dp2a_lo(w[j + 2], C1, __dp2a_hi(w[j + 3], C1, 0u))));
            ob[8 * yq + 2 * j][x] = (uint8_t)((a + 32768u) >> 16);
            ob[8 * yq + 2 * j + 1][x] = (uint8_t)((b + 32768u) >> 16);
        }
    }
    __syncthreads();
    uint8_t *dst = blur + (size_t)frame * blur_frame + L.boff;
    // 32 rows x 64 bytes as 16-byte stores (the blurred level's pitch is a multiple of 16, x0 of 64)
    if (tid < ORBX_BLUR_TH * (ORBX_BLUR_TW / 16)) {
        const int y = tid >> 2, c16 = (tid & 3) * 16;
        const int oy = t.y0 + y, ox = t.x0 + c16;
        if (oy < L.h && ox < L.bpitch)
            *reinterpret_cast<uint4 *>(dst + (size_t)oy * L.bpitch + ox) = *reinterpret_cast<const uint4 *>(&ob[y][c16]);
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
