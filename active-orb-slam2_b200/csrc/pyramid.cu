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

// level 0: copy with reflection from the caller's image.  One item = 16 consecutive bytes of one padded row (one 128-bit store).
// The first two and the last three items of a row touch the reflected border and assemble their bytes one by one (~300 instructions);
// the others are five aligned loads and four funnel shifts.  With a warp over consecutive items of a row every warp contained border
// items and paid for both paths (406 instructions per warp, ncu r2_ab); so a block takes L0_ROWS rows, its first warps take the interior
// items and ONE warp takes all the border items of those rows.
#define L0_ROWS 6
#define L0_EDGE_ITEMS 5                      // per row: items 0, 1 and the last three
__device__ __forceinline__ void pyr_level0_item(const uint8_t *__restrict__ img, int stride, uint8_t *__restrict__ lvl, const OrbxLevel &L, int py,
                                                int item) {
    const int gx = item * 16;
    uint8_t *dst = lvl + (size_t)py * L.pitch + gx;
    const int iy = reflect101(py - ORBX_EDGE, L.h);
    const uint8_t *row = img + (size_t)iy * stride;
    uint32_t w[4];
    const int sx = gx - ORBX_EDGE;                     // source column of the first of the 16 bytes
    const int mis = sx & 3;
    if (sx >= 4 && sx - mis + 20 <= L.w && (reinterpret_cast<uintptr_t>(row) & 3) == 0) {
        // interior: five aligned words cover the 16 source bytes (no reflection inside), one funnel shift per output word
        const uint32_t *p = reinterpret_cast<const uint32_t *>(row + sx - mis);
        const uint32_t u0 = __ldg(p), u1 = __ldg(p + 1), u2 = __ldg(p + 2), u3 = __ldg(p + 3), u4 = __ldg(p + 4);
        const int sh = 8 * mis;
        w[0] = __funnelshift_r(u0, u1, sh); w[1] = __funnelshift_r(u1, u2, sh);
        w[2] = __funnelshift_r(u2, u3, sh); w[3] = __funnelshift_r(u3, u4, sh);
        *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
        return;
    }
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

// grid (row groups, 1, frames); block = (interior warps + 1) x 32 threads
__global__ void __launch_bounds__(1024) k_pyr_level0(const uint8_t *__restrict__ src, size_t frame_pitch, int stride,
                                                     uint8_t *__restrict__ pyr, size_t pyr_frame, OrbxLevel L) {
    const int n_items = L.pitch / 16, n_mid = n_items - L0_EDGE_ITEMS;          // items per row, interior items per row
    const int row0 = blockIdx.x * L0_ROWS;
    const uint8_t *img = src + (size_t)blockIdx.z * frame_pitch;
    uint8_t *lvl = pyr + (size_t)blockIdx.z * pyr_frame + L.off;
    const int edge_warp = blockDim.x / 32 - 1, warp = threadIdx.x >> 5;
    if (warp < edge_warp) {
        for (int t = threadIdx.x; t < L0_ROWS * n_mid; t += edge_warp * 32) {
            const int r = t / n_mid, py = row0 + r;
            if (py < L.ph) pyr_level0_item(img, stride, lvl, L, py, 2 + (t - r * n_mid));
        }
    } else {
        const int lane = threadIdx.x & 31;
        if (lane < L0_ROWS * L0_EDGE_ITEMS) {
            const int r = lane / L0_EDGE_ITEMS, k = lane - r * L0_EDGE_ITEMS, py = row0 + r;
            const int item = k < 2 ? k : n_items - L0_EDGE_ITEMS + k;
            if (py < L.ph && item < n_items && (k < 2 || item >= 2)) pyr_level0_item(img, stride, lvl, L, py, item);
        }
    }
}

// OpenCV's fixed-point bilinear for one pixel: t0 / t1 hold the two horizontal neighbours of the upper / lower source row
// in their low 16 bits, coef = a0 | a1 << 16 (a0 + a1 = 2048).  The horizontal step a0*p[sx] + a1*p[sx+1] is one DP2A.
// b0s / b1s = the vertical coefficients shifted left by 16 (0 .. 2048 << 16), so that (b * (s >> 4)) >> 16 is one multiply-high
__device__ __forceinline__ uint32_t resize_px(uint32_t coef, uint32_t b0s, uint32_t b1s, uint32_t t0, uint32_t t1) {
    const uint32_t s0 = __dp2a_lo(coef, t0, 0u);
    const uint32_t s1 = __dp2a_lo(coef, t1, 0u);
    return (__umulhi(b0s, s0 >> 4) + __umulhi(b1s, s1 >> 4) + 2u) >> 2;
}

// level l >= 1 from the interior of level l-1.  One thread = 16 consecutive bytes of one padded output row (one 128-bit
// store), as four groups of four pixels.  Tables are indexed by padded output coordinates (the reflected border is resolved
// on the host).  When the four source positions of a group span at most 6 bytes (flag bit of the group; always true at
// ORB-SLAM's scale factors, also across the reflection points, because the window starts at the smallest position) each
// source row costs three aligned 32-bit loads, two funnel shifts that bring bytes base .. base+7 into a register pair, and
// one byte permute per pixel; otherwise bytes are loaded one by one.
// one item = 16 consecutive bytes of padded row py of level L of one frame, from the interior of level S (see k_pyr_resize)
__device__ __forceinline__ void pyr_resize_item(uint8_t *__restrict__ frame, const OrbxLevel &S, const OrbxLevel &L, const uint4 *__restrict__ rx,
                                                const int2 *__restrict__ ry, int py, int g16) {
    const uint8_t *sint = frame + S.off + (size_t)ORBX_EDGE * S.pitch + ORBX_EDGE;  // interior origin of the source
    const int2 yy = __ldg(ry + py);
    const int sy0 = yy.x & 0xffff, sy1 = yy.x >> 16;
    const uint32_t b0 = (uint32_t)(yy.y & 0xffff) << 16, b1 = (uint32_t)(yy.y >> 16) << 16;       // 0 .. 2048, pre-shifted for resize_px
    const uint8_t *r0 = sint + (size_t)sy0 * S.pitch;
    const ptrdiff_t r10 = ((ptrdiff_t)sy1 - sy0) * S.pitch;      // lower source row relative to the upper one (multiple of 16)
    uint32_t out[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint4 e4 = __ldg(rx + 4 * g16 + g);
        const uint32_t e[4] = {e4.x, e4.y, e4.z, e4.w};
        uint32_t v = 0;
        if (e[0] >> 31) {
            const int base = (int)min(min(e[0] & 0xffff, e[1] & 0xffff), min(e[2] & 0xffff, e[3] & 0xffff));
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(r0 + base);
            const int mis = (int)(a0 & 3);
            const uint32_t *p0 = reinterpret_cast<const uint32_t *>(a0 - mis);
            const uint32_t *p1 = reinterpret_cast<const uint32_t *>(a0 - mis + r10);
            const uint32_t u0 = p0[0], u1 = p0[1], u2 = p0[2], v0 = p1[0], v1 = p1[1], v2 = p1[2];
            const int sh = 8 * mis;
            const uint32_t qa0 = __funnelshift_r(u0, u1, sh), qa1 = __funnelshift_r(u1, u2, sh);   // bytes base .. base+7, upper row
            const uint32_t qb0 = __funnelshift_r(v0, v1, sh), qb1 = __funnelshift_r(v1, v2, sh);   // ... lower row
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t d = (e[b] & 0xffff) - (uint32_t)base;      // 0 .. 6
                const uint32_t sel = d * 0x11u + 0x10u;                   // PRMT selector: bytes d, d+1 of the 8-byte window
                const uint32_t a1 = (e[b] >> 16) & 0x7fff;
                const uint32_t coef = (2048u - a1) | (a1 << 16);
                v |= resize_px(coef, b0, b1, __byte_perm(qa0, qa1, sel), __byte_perm(qb0, qb1, sel)) << (8 * b);
            }
        } else {
            const uint8_t *r1 = r0 + r10;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int sx = (int)(e[b] & 0xffff);          // sx+1 may touch the pad: a1 == 0 there
                const uint32_t t0 = (uint32_t)r0[sx] | ((uint32_t)r0[sx + 1] << 8), t1 = (uint32_t)r1[sx] | ((uint32_t)r1[sx + 1] << 8);
                const uint32_t a1 = (e[b] >> 16) & 0x7fff;
                v |= resize_px((2048u - a1) | (a1 << 16), b0, b1, t0, t1) << (8 * b);
            }
        }
        out[g] = v;
    }
    *reinterpret_cast<uint4 *>(frame + L.off + (size_t)py * L.pitch + 16 * g16) = make_uint4(out[0], out[1], out[2], out[3]);
}

__global__ void __launch_bounds__(256) k_pyr_resize(uint8_t *__restrict__ pyr, size_t pyr_frame, OrbxLevel S, OrbxLevel L,
                                                    const uint4 *__restrict__ rx, const int2 *__restrict__ ry, int cols16) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = id / cols16, g16 = id - py * cols16;
    if (py >= L.ph) return;
    pyr_resize_item(pyr + (size_t)blockIdx.z * pyr_frame, S, L, rx, ry, py, g16);
}

// The small levels in one launch: a thread-block cluster of PYR_TAIL_CTAS CTAs owns one frame and walks levels first .. nlevels-1 one
// after the other, every level spread over all threads of the cluster, with a cluster barrier (release / acquire: the level just
// written to global memory is visible to the whole cluster) before the next one reads it.  Level by level launches of these levels
// are latency- and tail-bound (8-15 us each for a few hundred thousand pixels); chained here they cost one launch.  Kept for the
// measurement (see orbx_launch_pyramid): it did not pay.
#define PYR_TAIL_CTAS 8
__global__ void __launch_bounds__(256) k_pyr_tail(uint8_t *__restrict__ pyr, size_t pyr_frame, const OrbxLevel *__restrict__ lv, int first, int nlevels,
                                                  const uint32_t *__restrict__ rxt, const int2 *__restrict__ ryt) {
    uint8_t *frame = pyr + (size_t)blockIdx.z * pyr_frame;
    const int t = blockIdx.x * 256 + threadIdx.x, nt = PYR_TAIL_CTAS * 256;
    for (int l = first; l < nlevels; l++) {
        const OrbxLevel S = lv[l - 1], L = lv[l];
        const int cols16 = L.pitch / 16, total = cols16 * L.ph;
        const uint4 *rx = reinterpret_cast<const uint4 *>(rxt + L.rx_off);
        const int2 *ry = ryt + L.ry_off;
        for (int id = t; id < total; id += nt) {
            const int py = id / cols16;
            pyr_resize_item(frame, S, L, rx, ry, py, id - py * cols16);
        }
        if (l + 1 < nlevels) {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    }
}

orbx_status orbx_launch_pyramid(orbx_extractor *e, const uint8_t *d_images, size_t frame_pitch, int batch, int stride,
                                cudaStream_t s) {
    // levels from `tail` on would go through the chained cluster kernel; off by default (ORBX_PYR_TAIL=<first level> switches it on):
    // measured on the B200 it is no faster than the level-by-level launches (64 VGA frames: pyramid 0.134 ms with levels 3-7 chained
    // against 0.128 ms; one frame: 0.044 against 0.036 ms) -- 2048 threads per frame leave the chain latency-bound
    static const int tail_env = getenv("ORBX_PYR_TAIL") ? atoi(getenv("ORBX_PYR_TAIL")) : ORBX_MAX_LEVELS;
    const int tail = tail_env < 1 ? 1 : tail_env;
    for (int l = 0; l < e->nlevels && l < tail; l++) {
        const OrbxLevel &L = e->lv[l];
        if (l == 0) {
            // interior warps for L0_ROWS rows of (items - 5) interior items each, plus the warp of the border items
            const int n_mid = L.pitch / 16 - L0_EDGE_ITEMS;
            int warps = (L0_ROWS * (n_mid > 0 ? n_mid : 0) + 31) / 32;
            warps = warps < 1 ? 1 : (warps > 31 ? 31 : warps);
            dim3 block((warps + 1) * 32);
            dim3 grid((L.ph + L0_ROWS - 1) / L0_ROWS, 1, batch);
            k_pyr_level0<<<grid, block, 0, s>>>(d_images, frame_pitch, stride, e->d_pyr, e->pyr_frame_cap, L);
        } else {
            const int cols16 = L.pitch / 16, total = cols16 * L.ph;
            dim3 grid((total + 255) / 256, 1, batch);
            k_pyr_resize<<<grid, 256, 0, s>>>(e->d_pyr, e->pyr_frame_cap, e->lv[l - 1], L,
                                              reinterpret_cast<const uint4 *>(e->d_rxt + L.rx_off), e->d_ryt + L.ry_off, cols16);
        }
        e->last_launches++;
    }
    if (tail < e->nlevels) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(PYR_TAIL_CTAS, 1, batch);
        cfg.blockDim = dim3(256);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = PYR_TAIL_CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        ORBX_CUDA(cudaLaunchKernelEx(&cfg, k_pyr_tail, e->d_pyr, e->pyr_frame_cap, (const OrbxLevel *)e->d_lv, tail, e->nlevels,
                                     (const uint32_t *)e->d_rxt, (const int2 *)e->d_ryt));
        e->last_launches++;
    }
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}
