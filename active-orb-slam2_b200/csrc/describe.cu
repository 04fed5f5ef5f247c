// Orientation + 256-bit rBRIEF descriptor + final cv::KeyPoint assembly: replaces computeOrientation / IC_Angle
// (reference src/ORBextractor.cc:472-479, :77-104), computeDescriptors / computeOrbDescriptor (:1034-1041,
// :108-147) and the tail of operator() (:1075-1104: level-major concatenation, pt *= scale[level]).
//
// One warp per keypoint.  IC_Angle: lane u+15 owns column u of the r=15 circular patch (integer moments, exact
// in any order, warp-reduced).  Descriptor: lane j produces byte j (8 steered comparisons on the blurred level).
// Every float operation that reaches a rounding decision uses an explicit round-to-nearest intrinsic
// (never contracted to FMA), mirroring the oracle built with -ffp-contract=off.
#include <float.h>
#include "orbx_internal.cuh"

#define DESC_THREADS 256
#define DESC_WARPS (DESC_THREADS / 32)

// 256 test pairs x (x0,y0,x1,y1); lane j reads its own 32 bytes, so this lives in global memory (L1-cached,
// two 128-bit loads per lane) rather than in the constant bank, whose divergent reads serialise 32-way
__device__ __align__(16) int8_t g_pattern[1024] = {
#include "orb_pattern.inc"
};
// umax of the r = 15 circular patch, ORBextractor.cc:454-469 (compile-time, so that the row loop unrolls into predicated loads)
__device__ constexpr int k_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// cv::fastAtan2 (OpenCV mathfuncs_core atan_f32), degrees
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float sc = (float)(180 / 3.1415926535897932384626433832795);
    const float p1 = __fmul_rn(0.9997878412794807f, sc), p3 = __fmul_rn(-0.3258083974640975f, sc);
    const float p5 = __fmul_rn(0.1555786518463281f, sc), p7 = __fmul_rn(-0.04432655554792128f, sc);
    const float eps = (float)DBL_EPSILON;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// the oracle's cosf/sinf stand-in (double Cody-Waite + polynomial, one rounding to float); ORBextractor.cc:112-113
__device__ __forceinline__ void sincos_rn(float xf, float &s, float &c) {
    const double x = (double)xf;
    const double kf = rint(__dmul_rn(x, 6.36619772367581382433e-01));
    double r = __dsub_rn(x, __dmul_rn(kf, 1.57079632673412561417e+00));
    r = __dsub_rn(r, __dmul_rn(kf, 6.07710050650619224932e-11));
    const double z = __dmul_rn(r, r);
    double ps = 1.58969099521155010221e-10;
    ps = __dadd_rn(__dmul_rn(ps, z), -2.50507602534068634195e-08);
    ps = __dadd_rn(__dmul_rn(ps, z), 2.75573137070700676789e-06);
    ps = __dadd_rn(__dmul_rn(ps, z), -1.98412698298579493134e-04);
    ps = __dadd_rn(__dmul_rn(ps, z), 8.33333333332248946124e-03);
    ps = __dadd_rn(__dmul_rn(ps, z), -1.66666666666666324348e-01);
    const double sn = __dadd_rn(r, __dmul_rn(__dmul_rn(r, z), ps));
    double pc = -1.13596475577881948265e-11;
    pc = __dadd_rn(__dmul_rn(pc, z), 2.08757232129817482790e-09);
    pc = __dadd_rn(__dmul_rn(pc, z), -2.75573143513906633035e-07);
    pc = __dadd_rn(__dmul_rn(pc, z), 2.48015872894767294178e-05);
    pc = __dadd_rn(__dmul_rn(pc, z), -1.38888888888741095749e-03);
    pc = __dadd_rn(__dmul_rn(pc, z), 4.16666666666666019037e-02);
    const double cs = __dadd_rn(__dsub_rn(1.0, __dmul_rn(0.5, z)), __dmul_rn(__dmul_rn(z, z), pc));
    const long long k = (long long)kf & 3;
    double so, co;
    switch (k) {
    case 0: so = sn; co = cs; break;
    case 1: so = cs; co = -sn; break;
    case 2: so = -sn; co = -cs; break;
    default: so = -cs; co = sn; break;
    }
    s = (float)so;
    c = (float)co;
}

__global__ void __launch_bounds__(DESC_THREADS)
k_describe(const uint8_t *__restrict__ pyr, size_t pyr_frame, const uint8_t *__restrict__ blur, size_t blur_frame,
           const OrbxLevel *__restrict__ lv, int nlevels, const uint32_t *__restrict__ lvl_kp, int kp_frame,
           const int *__restrict__ lvl_cnt, orbx_keypoint *__restrict__ out_kps, uint8_t *__restrict__ out_desc,
           int32_t *__restrict__ out_counts) {
    const int frame = blockIdx.y, lane = threadIdx.x & 31;
    const int g = blockIdx.x * DESC_WARPS + (threadIdx.x >> 5);   // index in the level-major output
    int total = 0, level = -1, idx = 0;
    for (int l = 0; l < nlevels; l++) {
        const int c = lvl_cnt[frame * ORBX_MAX_LEVELS + l];
        if (level < 0 && g < total + c) { level = l; idx = g - total; }
        total += c;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out_counts[frame] = total;
    if (level < 0) return;
    const OrbxLevel &L = lv[level];
    const uint32_t w = lvl_kp[(size_t)frame * kp_frame + L.kp_off + idx];
    const int x = (int)(w & 0xfff) + ORBX_BORDER, y = (int)((w >> 12) & 0xfff) + ORBX_BORDER;   // ORBextractor.cc:843-844

    // ---- IC_Angle on the un-blurred level -----------------------------------------------------------
    const int pitch = L.pitch;
    const uint8_t *center = pyr + (size_t)frame * pyr_frame + L.off + (size_t)(ORBX_EDGE + y) * pitch + ORBX_EDGE + x;
    const int u = lane - ORBX_HALF_PATCH;
    const int au = u < 0 ? -u : u;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        // column u of the circular patch: rows +v and -v together (IC_Angle's own symmetry, ORBextractor.cc:80-102), two running row
        // pointers instead of a 64-bit v * pitch per row: the sums are integers, so the order does not matter (half the instructions of
        // the row-by-row loop, which was half of this kernel: profiles/r2_av_lines.txt)
        const uint8_t *up = center + u, *dn = up;
        int colsum = *up;                                  // v = 0: umax[0] = 15 covers every lane < 31
#pragma unroll
        for (int v = 1; v <= ORBX_HALF_PATCH; v++) {
            up += pitch; dn -= pitch;
            if (au <= k_umax[v]) {
                const int a = *up, b = *dn;
                colsum += a + b;
                m01 += v * (a - b);
            }
        }
        m10 = u * colsum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- steered BRIEF on the blurred level ---------------------------------------------------------
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
    float sn, cs;
    sincos_rn(__fmul_rn(angle, factorPI), sn, cs);
    const int bp = L.bpitch;
    const uint8_t *bc = blur + (size_t)frame * blur_frame + L.boff + (size_t)y * bp + x;
    union { int4 v[2]; int8_t b[32]; } pu;
    pu.v[0] = __ldg(reinterpret_cast<const int4 *>(g_pattern) + 2 * lane);
    pu.v[1] = __ldg(reinterpret_cast<const int4 *>(g_pattern) + 2 * lane + 1);
    const int8_t *pat = pu.b;
    unsigned val = 0;
#pragma unroll
    for (int bit = 0; bit < 8; bit++) {
        const float x0 = (float)pat[4 * bit], y0 = (float)pat[4 * bit + 1];
        const float x1 = (float)pat[4 * bit + 2], y1 = (float)pat[4 * bit + 3];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, sn), __fmul_rn(y0, cs)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, cs), __fmul_rn(y0, sn)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, sn), __fmul_rn(y1, cs)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, cs), __fmul_rn(y1, sn)));
        const int t0 = bc[r0 * bp + c0], t1 = bc[r1 * bp + c1];
        val |= (unsigned)(t0 < t1) << bit;
    }
    out_desc[((size_t)frame * kp_frame + g) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        orbx_keypoint k;
        k.x = (float)x;
        k.y = (float)y;
        if (level != 0) {   // ORBextractor.cc:1095-1101
            k.x = __fmul_rn(k.x, L.scale);
            k.y = __fmul_rn(k.y, L.scale);
        }
        k.size = L.kp_size;
        k.angle = angle;
        k.response = (float)(w >> 24);
        k.octave = level;
        k.class_id = -1;
        out_kps[(size_t)frame * kp_frame + g] = k;
    }
}

orbx_status orbx_launch_describe(orbx_extractor *e, int batch, orbx_keypoint *d_kps, uint8_t *d_desc, int32_t *d_counts,
                                 cudaStream_t s) {
    dim3 grid((e->capacity + DESC_WARPS - 1) / DESC_WARPS, batch);
    k_describe<<<grid, DESC_THREADS, 0, s>>>(e->d_pyr, e->pyr_frame_cap, e->d_blur, e->blur_frame_cap, e->d_lv, e->nlevels,
                                             e->d_lvl_kp, e->capacity, e->d_lvl_cnt, d_kps, d_desc, d_counts);
    e->last_launches++;
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}
