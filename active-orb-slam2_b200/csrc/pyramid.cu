// Scale pyramid with REFLECT_101 pad: replaces ORBextractor::ComputePyramid (reference
// src/ORBextractor.cc:1107-1132: cv::resize INTER_LINEAR + cv::copyMakeBorder).
//
// One launch per level (level l reads the interior of level l-1, so levels are chained), all frames of the
// batch in one grid.  Every thread produces 16 consecutive bytes of one padded row (one 128-bit store);
// pad pixels are produced by evaluating the interior pixel they mirror, so the border needs no second pass.
// Arithmetic is OpenCV's 11-bit fixed-point bilinear (tables built on the host, orbx_extractor.cu):
//   S  = a0*p[sx] + a1*p[sx+1]                      (int32, x2048)
//   out= (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2
#include "orbx_internal.cuh"

__device__ __forceinline__ int reflect101(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2 * n - 2 - i : i;
}

// level 0: copy with reflection from the caller's image
__global__ void __launch_bounds__(256) k_pyr_level0(const uint8_t *__restrict__ src, size_t frame_pitch, int stride,
                                                    uint8_t *__restrict__ pyr, size_t pyr_frame, OrbxLevel L) {
    const int gx = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int py = blockIdx.y;
    if (gx >= L.pitch) return;
    const uint8_t *img = src + (size_t)blockIdx.z * frame_pitch;
    uint8_t *dst = pyr + (size_t)blockIdx.z * pyr_frame + L.off + (size_t)py * L.pitch + gx;
    const int iy = reflect101(py - ORBX_EDGE, L.h);
    const uint8_t *row = img + (size_t)iy * stride;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int px = gx + q * 4 + b;
            uint32_t p = 0;
            if (px < L.w + 2 * ORBX_EDGE) p = __ldg(row + reflect101(px - ORBX_EDGE, L.w));
            v |= p << (8 * b);
        }
        w[q] = v;
    }
    *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}

// level l >= 1 from the interior of level l-1
__global__ void __launch_bounds__(256) k_pyr_resize(uint8_t *__restrict__ pyr, size_t pyr_frame, OrbxLevel S, OrbxLevel L,
                                                    const int2 *__restrict__ rx, const int2 *__restrict__ ry) {
    const int gx = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int py = blockIdx.y;
    if (gx >= L.pitch) return;
    uint8_t *frame = pyr + (size_t)blockIdx.z * pyr_frame;
    const uint8_t *sint = frame + S.off + (size_t)ORBX_EDGE * S.pitch + ORBX_EDGE;  // interior origin of the source
    uint8_t *dst = frame + L.off + (size_t)py * L.pitch + gx;
    const int iy = reflect101(py - ORBX_EDGE, L.h);
    const int2 yy = __ldg(ry + iy);
    const int sy0 = yy.x & 0xffff, sy1 = yy.x >> 16;
    const int b0 = (int)(short)(yy.y & 0xffff), b1 = yy.y >> 16;
    const uint8_t *r0 = sint + (size_t)sy0 * S.pitch;
    const uint8_t *r1 = sint + (size_t)sy1 * S.pitch;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int px = gx + q * 4 + b;
            uint32_t p = 0;
            if (px < L.w + 2 * ORBX_EDGE) {
                const int ix = reflect101(px - ORBX_EDGE, L.w);
                const int2 xx = __ldg(rx + ix);
                const int sx = xx.x;
                const int a0 = (int)(short)(xx.y & 0xffff), a1 = xx.y >> 16;
                const int s0 = a0 * (int)r0[sx] + a1 * (int)r0[sx + 1];   // sx+1 may touch the pad: a1 == 0 there
                const int s1 = a0 * (int)r1[sx] + a1 * (int)r1[sx + 1];
                p = (uint32_t)((((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2);
            }
            v |= p << (8 * b);
        }
        w[q] = v;
    }
    *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}

orbx_status orbx_launch_pyramid(orbx_extractor *e, const uint8_t *d_images, size_t frame_pitch, int batch, int stride,
                                cudaStream_t s) {
    for (int l = 0; l < e->nlevels; l++) {
        const OrbxLevel &L = e->lv[l];
        dim3 block(64);
        dim3 grid((L.pitch / 16 + block.x - 1) / block.x, L.ph, batch);
        if (l == 0)
            k_pyr_level0<<<grid, block, 0, s>>>(d_images, frame_pitch, stride, e->d_pyr, e->pyr_frame_cap, L);
        else
            k_pyr_resize<<<grid, block, 0, s>>>(e->d_pyr, e->pyr_frame_cap, e->lv[l - 1], L,
                                                e->d_rtab + L.rx_off, e->d_rtab + L.ry_off);
        e->last_launches++;
    }
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}
