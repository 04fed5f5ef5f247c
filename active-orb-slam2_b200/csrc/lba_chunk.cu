// Atomic-free system build for the host-driven local-BA path (lba.cu): the same chunked sums as the cluster kernel
// (lba_fused.cu), but spread over the whole GPU, one kernel per stage.
//   k2_build   two block roles in one launch (neither reads what the other writes):
//                four lanes per landmark: residuals, chi2, Jacobians, H_ll / b_l in registers, H_pl per edge;
//                warp per chunk of the keyframe-ordered edge list: partial H_pp / b_p
//   k2_pairs   warp per chunk of the block-ordered (edge, edge) list: partial B_i D^-1 B_j^T (+ coefficients), D^-1 from
//              H_ll + lambda I on the fly; the trailing blocks sum the H_pp / b_p partials in chunk order
//   k2_final   thread per entry of the upper block triangle: H_schur = H_pp + lambda I - sum, b_schur = b_p - sum
// Replaces BlockSolver::buildSystem and the Schur part of BlockSolver::solve (block_solver.hpp:371-439).
#include <algorithm>
#include "lba_common.cuh"

#define K2_THREADS 256

// chi_out / max_out (grid kernel): this block's robust chi2 sum and largest |H_ll diagonal| go to [block] instead of an atomic
__device__ __forceinline__ void k2_linearize_body(const LbaDev &D, int robust, int build, int block, double *kfRt, double *tmp,
                                                  double *chi_out = nullptr, double *max_out = nullptr) {
    for (int k = threadIdx.x; k < D.n_kf; k += K2_THREADS) {
        double R[9];
        quat_to_R(D.kf + 7 * k, R);
        for (int i = 0; i < 9; i++) kfRt[12 * k + i] = R[i];
        kfRt[12 * k + 9] = D.kf[7 * k + 4]; kfRt[12 * k + 10] = D.kf[7 * k + 5]; kfRt[12 * k + 11] = D.kf[7 * k + 6];
    }
    __syncthreads();
    // four lanes share a landmark: edge i of the landmark goes to lane i mod 4, the partial blocks meet by shuffle
    const int l = (block * K2_THREADS + threadIdx.x) >> 2, sub = threadIdx.x & 3;
    double chi = 0;
    double hl[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (l < D.n_pts) {
        for (int e = D.ptstart[l] + sub; e < D.ptstart[l + 1]; e += 4) {
            if (D.level1[e]) continue;
            const int kf = D.ekf[e];
            const double *R = kfRt + 12 * kf;
            double Xc[3], er[3];
            edge_residual(D, e, R, R + 9, Xc, er);
            D.err[3 * e] = er[0]; D.err[3 * e + 1] = er[1]; D.err[3 * e + 2] = er[2];
            const double info = D.info[e];
            const double c = info * (er[0] * er[0] + er[1] * er[1] + er[2] * er[2]);
            D.chi2[e] = c;
            double rho1 = 1.0, cr = c;
            if (robust) {
                const double d = D.stereo[e] ? D.d_stereo : D.d_mono, dsqr = (double)(float)(d * d);   // RobustKernelHuber keeps dsqr in a float member (robust_kernel_impl.h:84)
                if (c > dsqr) { const double sq = sqrt(c); cr = 2 * sq * d - dsqr; rho1 = d / sq; }
            }
            chi += cr;
            if (!build) continue;
            const int dim = D.stereo[e] ? 3 : 2;
            const double x = Xc[0], y = Xc[1], iz = 1.0 / Xc[2], iz2 = iz * iz, fx = D.fx, fy = D.fy, bf = D.bf;
            const double xz = x * iz, yz = y * iz;
            double A[9], B[18];
            for (int q = 0; q < 3; q++) {
                A[q] = -fx * R[q] * iz + fx * xz * R[6 + q] * iz;
                A[3 + q] = -fy * R[3 + q] * iz + fy * yz * R[6 + q] * iz;
                A[6 + q] = dim == 3 ? A[q] - bf * R[6 + q] * iz2 : 0.0;
            }
            B[0] = xz * yz * fx; B[1] = -(1 + xz * xz) * fx; B[2] = yz * fx; B[3] = -iz * fx; B[4] = 0; B[5] = xz * iz * fx;
            B[6] = (1 + yz * yz) * fy; B[7] = -xz * yz * fy; B[8] = -xz * fy; B[9] = 0; B[10] = -iz * fy; B[11] = yz * iz * fy;
            if (dim == 3) { B[12] = B[0] - bf * y * iz2; B[13] = B[1] + bf * x * iz2; B[14] = B[2]; B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf * iz2; }
            else { for (int i = 12; i < 18; i++) B[i] = 0; }
            const double w = rho1 * info;
            double wr[3];
            for (int d = 0; d < 3; d++) wr[d] = -info * er[d] * rho1;
            int k = 0;
            for (int a = 0; a < 3; a++)
                for (int b = a; b < 3; b++) hl[k++] += w * (A[a] * A[b] + A[3 + a] * A[3 + b] + A[6 + a] * A[6 + b]);
            for (int a = 0; a < 3; a++) hl[6 + a] += A[a] * wr[0] + A[3 + a] * wr[1] + A[6 + a] * wr[2];
            if (D.kfidx[kf] >= 0) {
                double *hpl = D.Hpl + 18 * (size_t)e;
                for (int a = 0; a < 6; a++)
                    for (int b = 0; b < 3; b++) hpl[3 * a + b] = w * (B[a] * A[b] + B[6 + a] * A[3 + b] + B[12 + a] * A[6 + b]);
            }
        }
    }
    if (build) {
#pragma unroll
        for (int i = 0; i < 9; i++) {
            hl[i] += __shfl_xor_sync(0xffffffffu, hl[i], 1);
            hl[i] += __shfl_xor_sync(0xffffffffu, hl[i], 2);
        }
        if (l < D.n_pts && sub == 0) for (int i = 0; i < 9; i++) D.Hll[9 * l + i] = hl[i];
    }
    const double s = block_sum(chi, tmp);
    if (chi_out) { if (threadIdx.x == 0) chi_out[block] = s; }
    else if (threadIdx.x == 0 && s != 0) atomicAdd(&D.scal[0], s);
    if (build && max_out) {
        double mx = (l < D.n_pts && sub == 0) ? fmax(fabs(hl[0]), fmax(fabs(hl[3]), fabs(hl[5]))) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        __syncthreads();                                   // tmp is free again (block_sum ends with a barrier, this one orders the reuse)
        if ((threadIdx.x & 31) == 0) tmp[threadIdx.x >> 5] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            double m = 0;
            for (int w = 0; w < K2_THREADS / 32; w++) m = fmax(m, tmp[w]);
            max_out[block] = m;
        }
        __syncthreads();
    }
}

// done (grid kernel): per-keyframe arrival counters; the warp that completes a keyframe's chunks adds them up in chunk order into D.Hpp
__device__ __forceinline__ void k2_hpp_body(const LbaDev &D, int robust, int block, int *done = nullptr) {
    const int lane = threadIdx.x & 31, ch = block * (K2_THREADS / 32) + (threadIdx.x >> 5);
    if (ch >= D.n_kchunks) return;
    const int4 cd = D.kchunk[ch];
    double acc[27];
#pragma unroll
    for (int i = 0; i < 27; i++) acc[i] = 0;
    for (int q = lane; q < cd.z; q += 32) {
        const int4 ke = D.kfe[cd.y + q];
        const int e = ke.x;
        if (D.level1[e]) continue;
        double R[9];
        const double *T = D.kf + 7 * ke.y, *X = D.pt + 3 * ke.z;
        quat_to_R(T, R);
        const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + T[4], y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + T[5];
        const double iz = 1.0 / (R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + T[6]), iz2 = iz * iz, xz = x * iz, yz = y * iz;
        const double fx = D.fx, fy = D.fy, bf = D.bf;
        const bool st = ke.w != 0;
        double B[18];
        B[0] = xz * yz * fx; B[1] = -(1 + xz * xz) * fx; B[2] = yz * fx; B[3] = -iz * fx; B[4] = 0; B[5] = xz * iz * fx;
        B[6] = (1 + yz * yz) * fy; B[7] = -xz * yz * fy; B[8] = -xz * fy; B[9] = 0; B[10] = -iz * fy; B[11] = yz * iz * fy;
        if (st) { B[12] = B[0] - bf * y * iz2; B[13] = B[1] + bf * x * iz2; B[14] = B[2]; B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf * iz2; }
        else { for (int i = 12; i < 18; i++) B[i] = 0; }
        // the residual is recomputed (same arithmetic as k2_linearize) so that this pass does not wait for that one
        double Xc[3], er[3];
        edge_residual(D, e, R, T + 4, Xc, er);
        const double info = D.info[e], c = info * (er[0] * er[0] + er[1] * er[1] + er[2] * er[2]);
        double rho1 = 1.0;
        if (robust) {
            const double d = st ? D.d_stereo : D.d_mono;
            if (c > d * d) rho1 = d / sqrt(c);
        }
        const double w = rho1 * info;
        const double w0 = -info * er[0] * rho1, w1 = -info * er[1] * rho1, w2 = -info * er[2] * rho1;
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; a++)
#pragma unroll
            for (int b = a; b < 6; b++) acc[k++] += w * (B[a] * B[b] + B[6 + a] * B[6 + b] + B[12 + a] * B[12 + b]);
#pragma unroll
        for (int a = 0; a < 6; a++) acc[21 + a] += B[a] * w0 + B[6 + a] * w1 + B[12 + a] * w2;
    }
    double mine = 0;
#pragma unroll
    for (int i = 0; i < 27; i++) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == i) mine = v;
    }
    if (lane < 27) D.hppart[(size_t)ch * 27 + lane] = mine;
    if (done) {
        __threadfence();
        const int p = cd.x, c0 = D.kf_cstart[p], c1 = D.kf_cstart[p + 1];
        int last = 0;
        if (lane == 0) last = atomicAdd(&done[p], 1) == c1 - c0 - 1;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            if (lane < 27) {
                double v = 0;
                for (int c = c0; c < c1; c++) v += __ldcg(D.hppart + (size_t)c * 27 + lane);
                D.Hpp[27 * p + lane] = v;
            }
            if (lane == 0) done[p] = 0;
        }
    }
}

// residuals (+ quadratic form): the first nb_lin blocks own the landmarks, the rest the keyframe chunks
__global__ void __launch_bounds__(K2_THREADS) k2_build(LbaDev D, int robust, int build, int nb_lin) {
    extern __shared__ __align__(16) double kfRt[];     // [n_kf][12]
    __shared__ double tmp[32];
    if ((int)blockIdx.x < nb_lin) k2_linearize_body(D, robust, build, blockIdx.x, kfRt, tmp);
    else k2_hpp_body(D, robust, blockIdx.x - nb_lin);
}

// H_pp / b_p from the chunk partials; and D^-1, D^-1 b_l of every landmark for the given lambda
__global__ void __launch_bounds__(K2_THREADS) k2_hpp_final_dinv(LbaDev D, double lambda, int do_hpp, int do_dinv) {
    const int i = blockIdx.x * K2_THREADS + threadIdx.x;
    if (do_hpp && i < 27 * D.np) {
        const int p = i / 27, c = i - 27 * p;
        double v = 0;
        for (int ch = D.kf_cstart[p]; ch < D.kf_cstart[p + 1]; ch++) v += D.hppart[(size_t)ch * 27 + c];
        D.Hpp[i] = v;
    }
    if (do_dinv && i < D.n_pts) {
        const double *hl = D.Hll + 9 * i;
        double Di[6];
        dinv3(hl, lambda, Di);
        double *o = D.dinv + 10 * (size_t)i;
        for (int k = 0; k < 6; k++) o[k] = Di[k];
        o[6] = Di[0] * hl[6] + Di[1] * hl[7] + Di[2] * hl[8];
        o[7] = Di[1] * hl[6] + Di[3] * hl[7] + Di[4] * hl[8];
        o[8] = Di[2] * hl[6] + Di[4] * hl[7] + Di[5] * hl[8];
    }
}

// warp per chunk of the block-ordered (edge, edge) list: partial B_i D^-1 B_j^T (+ coefficients), D^-1 from H_ll + lambda I on the fly
// done / hs_g (grid kernel): per-block arrival counters; the warp that completes a block's chunks writes the block of
// H_schur = H_pp + lambda I - sum (and b_schur = b_p - sum for a diagonal block), partials in chunk order, into hs_g (lba_solve.cuh layout)
__device__ __forceinline__ void k2_pairs_body(const LbaDev &D, double lambda, int block, int *done = nullptr, double *hs_g = nullptr) {
    const int lane = threadIdx.x & 31, ch = block * (K2_THREADS / 32) + (threadIdx.x >> 5);
    if (ch >= D.n_pchunks) return;
    const int4 cd = D.pchunk[ch];
    double acc[42];
#pragma unroll
    for (int i = 0; i < 42; i++) acc[i] = 0;
    for (int q = lane; q < cd.z; q += 32) {
        const int4 pe = D.pairs[cd.y + q];
        const uint8_t off1 = D.level1[pe.x], off2 = D.level1[pe.y];
        const double *hlp = D.Hll + 9 * (size_t)pe.z;
        const double2 *B1 = reinterpret_cast<const double2 *>(D.Hpl + 18 * (size_t)pe.x);
        const double2 *B2 = reinterpret_cast<const double2 *>(D.Hpl + 18 * (size_t)pe.y);
        double hl[9], Di[10], b1[18], b2[18];
#pragma unroll
        for (int i = 0; i < 9; i++) hl[i] = hlp[i];
#pragma unroll
        for (int i = 0; i < 9; i++) { const double2 t = B1[i]; b1[2 * i] = t.x; b1[2 * i + 1] = t.y; }
#pragma unroll
        for (int i = 0; i < 9; i++) { const double2 t = B2[i]; b2[2 * i] = t.x; b2[2 * i + 1] = t.y; }
        if (off1 || off2) continue;
        dinv3(hl, lambda, Di);                         // D^-1 = (H_ll + lambda I)^-1, then D^-1 b_l
        Di[6] = Di[0] * hl[6] + Di[1] * hl[7] + Di[2] * hl[8];
        Di[7] = Di[1] * hl[6] + Di[3] * hl[7] + Di[4] * hl[8];
        Di[8] = Di[2] * hl[6] + Di[4] * hl[7] + Di[5] * hl[8];
#pragma unroll
        for (int a = 0; a < 6; a++) {
            const double u0 = b1[3 * a], u1 = b1[3 * a + 1], u2 = b1[3 * a + 2];
            const double bd0 = u0 * Di[0] + u1 * Di[1] + u2 * Di[2], bd1 = u0 * Di[1] + u1 * Di[3] + u2 * Di[4],
                         bd2 = u0 * Di[2] + u1 * Di[4] + u2 * Di[5];
#pragma unroll
            for (int b = 0; b < 6; b++) acc[6 * a + b] += bd0 * b2[3 * b] + bd1 * b2[3 * b + 1] + bd2 * b2[3 * b + 2];
            if (cd.w) acc[36 + a] += u0 * Di[6] + u1 * Di[7] + u2 * Di[8];
        }
    }
    double mine = 0;
#pragma unroll
    for (int i = 0; i < 42; i++) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == (i & 31)) { if (i < 32) mine = v; else D.part[(size_t)ch * 42 + i] = v; }
    }
    D.part[(size_t)ch * 42 + lane] = mine;
    if (done) {
        __threadfence();
        const int blk = cd.x, c0 = D.blk_cstart[blk], c1 = D.blk_cstart[blk + 1];
        int last = 0;
        if (lane == 0) last = atomicAdd(&done[blk], 1) == c1 - c0 - 1;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            __threadfence();
            const int np = D.np, nblk = np * (np + 1) / 2;
            int p1 = 0, rem = blk;
            while (rem >= np - p1) { rem -= np - p1; p1++; }
            for (int ab = lane; ab < 36; ab += 32) {
                double v = 0;
                for (int c = c0; c < c1; c++) v -= __ldcg(D.part + (size_t)c * 42 + ab);
                if (cd.w) {
                    int a = ab / 6, b = ab - 6 * a;
                    const bool dg = a == b;
                    if (a > b) { const int t = a; a = b; b = t; }
                    v += __ldcg(D.Hpp + 27 * p1 + a * 6 - a * (a - 1) / 2 + (b - a)) + (dg ? lambda : 0.0);
                }
                hs_g[(size_t)blk * 36 + ab] = v;
            }
            if (cd.w && lane < 6) {
                double v = __ldcg(D.Hpp + 27 * p1 + 21 + lane);
                for (int c = c0; c < c1; c++) v -= __ldcg(D.part + (size_t)c * 42 + 36 + lane);
                hs_g[(size_t)nblk * 36 + 6 * p1 + lane] = v;
            }
            if (lane == 0) done[blk] = 0;
        }
    }
}

__global__ void __launch_bounds__(K2_THREADS) k2_pairs(LbaDev D, double lambda, int nb_pairs) {
    if ((int)blockIdx.x >= nb_pairs) {      // the remaining blocks: H_pp / b_p from the keyframe chunk partials
        const int i = (blockIdx.x - nb_pairs) * K2_THREADS + threadIdx.x;
        if (i < 27 * D.np) {
            const int p = i / 27, c = i - 27 * p;
            double v = 0;
            for (int ch = D.kf_cstart[p]; ch < D.kf_cstart[p + 1]; ch++) v += D.hppart[(size_t)ch * 27 + c];
            D.Hpp[i] = v;
        }
        return;
    }
    k2_pairs_body(D, lambda, blockIdx.x);
}


// The tensor-core form of k2_pairs, kept for the measurement BASELINE.json's north star asks for (LocalBA): for a pose-pair
// block the sum over shared landmarks of (B_i D^-1) B_j^T is a 6 x K by K x 6 product with K = 3 per landmark, issued here as one
// mma.sync.aligned.m8n8k4.f64 (DMMA) per (edge, edge) pair: rows = the six pose-i coordinates (two idle), columns = the six pose-j
// coordinates, k = the landmark's three coordinates; the fourth k slot carries D^-1 b_l against a unit column, so the coefficient
// vector B_i D^-1 b_l accumulates in column 6 of the same fragment.  A warp walks its chunk pair by pair with the 8 x 8 accumulator in
// registers, so the 42 warp-shuffle reductions of the scalar kernel disappear -- but each DMMA carries 108 useful multiply-adds and
// needs seven loads per lane to feed it, and B200's f64 tensor rate equals its f64 FMA rate.  Needs D^-1 precomputed (k2_hpp_final_dinv).
__global__ void __launch_bounds__(K2_THREADS) k2_pairs_mma(LbaDev D, int nb_pairs) {
    if ((int)blockIdx.x >= nb_pairs) {      // the remaining blocks: H_pp / b_p from the keyframe chunk partials
        const int i = (blockIdx.x - nb_pairs) * K2_THREADS + threadIdx.x;
        if (i < 27 * D.np) {
            const int p = i / 27, c = i - 27 * p;
            double v = 0;
            for (int ch = D.kf_cstart[p]; ch < D.kf_cstart[p + 1]; ch++) v += D.hppart[(size_t)ch * 27 + c];
            D.Hpp[i] = v;
        }
        return;
    }
    const int lane = threadIdx.x & 31, ch = blockIdx.x * (K2_THREADS / 32) + (threadIdx.x >> 5);
    if (ch >= D.n_pchunks) return;
    const int4 cd = D.pchunk[ch];
    const int ra = lane >> 2, ka = lane & 3;               // A fragment: row ra (pose-i coordinate), k = ka
    const int kb = lane & 3, nb = lane >> 2;               // B fragment: k = kb, column nb (pose-j coordinate)
    // D^-1 is stored as its six distinct entries (00 01 02 11 12 22) followed by D^-1 b_l: entry (m, ka) of the 3 x 4 matrix [D^-1 | D^-1 b_l]
    const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    const int d0 = ka < 3 ? sym[0][ka] : 6, d1 = ka < 3 ? sym[1][ka] : 7, d2 = ka < 3 ? sym[2][ka] : 8;
    const double unit = (kb == 3 && nb == 6 && cd.w) ? 1.0 : 0.0;
    double c0 = 0, c1 = 0;
    for (int q = 0; q < cd.z; q++) {
        const int4 pe = D.pairs[cd.y + q];
        if (D.level1[pe.x] || D.level1[pe.y]) continue;
        double a = 0, b = unit;
        if (ra < 6) {
            const double *B1 = D.Hpl + 18 * (size_t)pe.x + 3 * ra;
            const double *Di = D.dinv + 10 * (size_t)pe.z;
            a = B1[0] * Di[d0] + B1[1] * Di[d1] + B1[2] * Di[d2];
        }
        if (kb < 3 && nb < 6) b = D.Hpl[18 * (size_t)pe.y + 3 * nb + kb];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    // C fragment: lane holds C[ra][2 * (lane & 3)] and the next column; columns 0..5 = the block, column 6 = the coefficients
    const int col = 2 * (lane & 3);
    double *out = D.part + (size_t)ch * 42;
    if (ra < 6) {
        if (col < 6) { out[6 * ra + col] = c0; out[6 * ra + col + 1] = c1; }
        else out[36 + ra] = c0;
    }
}

// H_schur (upper block triangle of the n x n row-major matrix, the rest zero) and b_schur
__global__ void __launch_bounds__(K2_THREADS) k2_final(LbaDev D, double lambda) {
    const int np = D.np, n = D.n, nblk = np * (np + 1) / 2;
    const int i = blockIdx.x * K2_THREADS + threadIdx.x;
    if (i < n * n) {
        const int r = i / n, c = i - r * n, p1 = r / 6, p2 = c / 6;
        double v = 0;
        if (p1 <= p2) {
            const int a = r - 6 * p1, b = c - 6 * p2, blk = upper_block(p1, p2, np), ab = 6 * a + b;
            for (int ch = D.blk_cstart[blk]; ch < D.blk_cstart[blk + 1]; ch++) v -= D.part[(size_t)ch * 42 + ab];
            if (p1 == p2) {
                int a2 = a, b2 = b;
                if (a2 > b2) { const int t = a2; a2 = b2; b2 = t; }
                v += D.Hpp[27 * p1 + a2 * 6 - a2 * (a2 - 1) / 2 + (b2 - a2)] + (a == b ? lambda : 0.0);
            }
        }
        D.Hs[i] = v;
    }
    if (i < n) {
        const int p = i / 6, a = i - 6 * p, blk = upper_block(p, p, np);
        double v = D.Hpp[27 * p + 21 + a];
        for (int ch = D.blk_cstart[blk]; ch < D.blk_cstart[blk + 1]; ch++) v -= D.part[(size_t)ch * 42 + 36 + a];
        D.bs[i] = v;
    }
    (void)nblk;
}

static inline int k2_blocks(long long n) { return n > 0 ? (int)((n + K2_THREADS - 1) / K2_THREADS) : 1; }

// ---- one window on the whole GPU: one optimize() call in one cooperative kernel -------------------------------------------------------
// The cluster kernel (lba_fused.cu) gives a window 8 or 16 SMs, which is right when many windows are in flight; a single window
// (orbx_lba_solve_host: the LocalMapping thread waits for it) then spends its time in phases that 128 warps cannot fill
// (profiles/r2_ac_bench.json: 0.41 + 0.60 ms of quadratic form / Schur accumulation per window of 15 trials against 27 us per trial for the
// same arithmetic spread over the GPU by the kernels above).  Here the stages of those kernels run as phases of ONE persistent
// cooperative kernel, one CTA per SM, separated by grid barriers instead of launches, with the Levenberg loop on the device like in the
// cluster kernel: no host round trip per trial, no atomics (every block-level partial sum lands in its own slot and is added in slot
// order by every CTA, so all CTAs take the same decisions from the same numbers), and the reduced system is assembled and factorised
// by CTA 0 in its shared memory (lba_solve.cuh).
//   per iteration:  [landmark blocks: residuals, Jacobians, H_ll, H_pl | keyframe chunks: H_pp partials]  barrier
//   per trial:      [pair chunks: B_i D^-1 B_j^T partials]  barrier  [CTA 0: H_schur into shared memory, solve]  barrier
//                   [landmarks: back-substitution + update, poses: exp map]  barrier  [residuals at the trial state]  barrier  decision
#include <cooperative_groups.h>
#include "lba_solve.cuh"
namespace cg = cooperative_groups;

struct LgParams {
    LbaDev D;
    double *kf_bak, *pt_bak;
    int iterations, robust, capture;
    double *cap_Hs, *cap_bs, *cap_xp;
    double *out;                      // [0] += trials, [1] lambda of the first trial, [2..7] += nanoseconds per phase
    double *chi_part, *scl_part, *max_part;
    double *xp_g, *hs_g;              // the pose step (CTA 0 -> everybody) and the assembled reduced system (everybody -> CTA 0)
    int *done_kf, *done_blk;          // arrival counters of the keyframes' / blocks' chunks (zero between phases)
    int nb_lin, nb_hpp, nb_pairs, nb_upd;
};

// the n partial sums (or maxima) in slot order, the same value in every thread of every CTA
template <bool MAX>
__device__ __forceinline__ double lg_total(const double *part, int n, double *tmp) {
    if (threadIdx.x < 32) {
        double v = 0;
        for (int i = threadIdx.x; i < n; i += 32) { const double q = __ldcg(part + i); v = MAX ? fmax(v, q) : v + q; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const double q = __shfl_xor_sync(0xffffffffu, v, o); v = MAX ? fmax(v, q) : v + q; }
        if (threadIdx.x == 0) tmp[0] = v;
    }
    __syncthreads();
    const double r = tmp[0];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(K2_THREADS, 1) k_lba_grid(LgParams P) {
    extern __shared__ __align__(16) double dyn[];
    __shared__ double tmp[32];
    __shared__ int ok;
    cg::grid_group grid = cg::this_grid();
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
    const LbaDev &D = P.D;
    const int np = D.np, n = D.n, nblk = np * (np + 1) / 2, n_kf = D.n_kf;
    double *kfRt = dyn;                               // [64][12]
    double *hpp = kfRt + 12 * 64;                     // [np][27]  H_pp / b_p, a copy in every CTA
    double *hs = hpp + 27 * np;                       // CTA 0: the reduced system (lba_solve.cuh layout)
    double *xp = hs + nblk * 36 + 2 * n;              // [n] the pose step, a copy in every CTA
    if (cta == 0) lba_solve_table<K2_THREADS>(hs, np);
    double lambda = 0, ni = 2;
    int nBad = 0, trials = 0;
    bool first = P.capture != 0;
    unsigned long long tph[6] = {0, 0, 0, 0, 0, 0}, tlast = 0;
    auto tick = [&](int ph) {
        if (cta == 0 && tid == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (ph >= 0) tph[ph] += now - tlast;
            tlast = now;
        }
    };
    auto barrier = [&]() { __threadfence(); grid.sync(); };
    // computeActiveErrors + buildSystem at the current estimates: residuals, chi2, H_ll / b_l, H_pl, H_pp partials
    auto build = [&]() {
        for (int vb = cta; vb < P.nb_lin + P.nb_hpp; vb += G) {
            if (vb < P.nb_lin) k2_linearize_body(D, P.robust, 1, vb, kfRt, tmp, P.chi_part, P.max_part);
            else k2_hpp_body(D, P.robust, vb - P.nb_lin, P.done_kf);
            __syncthreads();
        }
        barrier();
    };
    // The pass that evaluates a trial IS the next iteration's build when the trial is accepted (same estimates, same weights), so it
    // always builds: an accepted iteration makes one pass over the edges instead of two.  A rejected trial that is retried rebuilds
    // at the restored estimates first (rare); one that ends the call leaves the trial's chi2 in place, like g2o's stored errors.
    bool have = false;
    for (int it = 0; it < P.iterations; it++) {
        tick(-1);
        if (!have) build();
        for (int i = tid; i < 27 * np; i += K2_THREADS) hpp[i] = __ldcg(D.Hpp + i);       // H_pp / b_p, finished by the build
        __syncthreads();
        double currentChi = lg_total<false>(P.chi_part, P.nb_lin, tmp);
        if (it == 0) {                                          // computeLambdaInit: the largest diagonal entry of the whole system
            double mx = lg_total<true>(P.max_part, P.nb_lin, tmp);
            for (int i = 0; i < 6 * np; i++) { const int a = i % 6; mx = fmax(mx, fabs(hpp[27 * (i / 6) + 6 * a - a * (a - 1) / 2])); }
            lambda = 1e-5 * mx; ni = 2; nBad = 0;
        }
        tick(0);
        const double iniChi = currentChi;
        double rho = 0;
        int qmax = 0;
        bool rejected = false;
        do {
            // ---- Schur complement: pair chunks ---------------------------------------------------------------------------------
            for (int vb = cta; vb < P.nb_pairs; vb += G) k2_pairs_body(D, lambda, vb, P.done_blk, P.hs_g);
            barrier();
            tick(1);
            if (cta == 0) {
                for (int i = tid; i < nblk * 36 + n; i += K2_THREADS) hs[i] = __ldcg(P.hs_g + i);
                __syncthreads();
                if (first && P.cap_Hs) {     // parity tests: the very first reduced system, expanded to a full symmetric matrix
                    for (int i = tid; i < n * n; i += K2_THREADS) {
                        int r = i / n, c = i - r * n;
                        if (r / 6 > c / 6) { const int t = r; r = c; c = t; }
                        P.cap_Hs[i] = hs[upper_block(r / 6, c / 6, np) * 36 + 6 * (r % 6) + c % 6];
                    }
                    for (int i = tid; i < n; i += K2_THREADS) P.cap_bs[i] = hs[nblk * 36 + i];
                    __syncthreads();
                }
                tick(2);
                lba_reduced_solve<K2_THREADS>(hs, xp, np, &ok);
                for (int i = tid; i < n; i += K2_THREADS) P.xp_g[i] = xp[i];
                if (tid == 0) D.scal[3] = ok ? 1.0 : 0.0;
                if (first && P.cap_xp) for (int i = tid; i < n; i += K2_THREADS) P.cap_xp[i] = xp[i];
                if (first && tid == 0) P.out[1] = lambda;
            }
            barrier();
            tick(3);
            first = false;
            if (cta != 0) for (int i = tid; i < n; i += K2_THREADS) xp[i] = __ldcg(P.xp_g + i);
            __syncthreads();
            // ---- landmark back-substitution, update (with push), computeScale ------------------------------------------------------
            for (int vb = cta; vb <= P.nb_upd; vb += G) {
                double sc = 0;
                if (vb < P.nb_upd) {       // four lanes per landmark (edge i of the landmark on lane i mod 4), like the linearisation
                    const int l = (vb * K2_THREADS + tid) >> 2, sub = tid & 3;
                    double c0 = 0, c1 = 0, c2 = 0;
                    if (l < D.n_pts) {
                        for (int e = D.ptstart[l] + sub; e < D.ptstart[l + 1]; e += 4) {
                            const int p = D.kfidx[D.ekf[e]];
                            if (p < 0 || D.level1[e]) continue;
                            const double *B = D.Hpl + 18 * (size_t)e, *x = xp + 6 * p;
#pragma unroll
                            for (int a = 0; a < 6; a++) { c0 -= B[3 * a] * x[a]; c1 -= B[3 * a + 1] * x[a]; c2 -= B[3 * a + 2] * x[a]; }
                        }
                    }
                    c0 += __shfl_xor_sync(0xffffffffu, c0, 1); c1 += __shfl_xor_sync(0xffffffffu, c1, 1); c2 += __shfl_xor_sync(0xffffffffu, c2, 1);
                    c0 += __shfl_xor_sync(0xffffffffu, c0, 2); c1 += __shfl_xor_sync(0xffffffffu, c1, 2); c2 += __shfl_xor_sync(0xffffffffu, c2, 2);
                    if (l < D.n_pts && sub == 0) {
                        const double *hl = D.Hll + 9 * l;
                        c0 += hl[6]; c1 += hl[7]; c2 += hl[8];
                        double Di[6];
                        dinv3(hl, lambda, Di);
                        const double x0 = Di[0] * c0 + Di[1] * c1 + Di[2] * c2, x1 = Di[1] * c0 + Di[3] * c1 + Di[4] * c2,
                                     x2 = Di[2] * c0 + Di[4] * c1 + Di[5] * c2;
                        sc = x0 * (lambda * x0 + hl[6]) + x1 * (lambda * x1 + hl[7]) + x2 * (lambda * x2 + hl[8]);
                        const double q0 = D.pt[3 * l], q1 = D.pt[3 * l + 1], q2 = D.pt[3 * l + 2];
                        P.pt_bak[3 * l] = q0; P.pt_bak[3 * l + 1] = q1; P.pt_bak[3 * l + 2] = q2;              // push
                        D.pt[3 * l] = q0 + x0; D.pt[3 * l + 1] = q1 + x1; D.pt[3 * l + 2] = q2 + x2;
                    }
                } else if (tid < n_kf) {
                    double T[7];
#pragma unroll
                    for (int i = 0; i < 7; i++) { T[i] = D.kf[7 * tid + i]; P.kf_bak[7 * tid + i] = T[i]; }
                    const int p = D.kfidx[tid];
                    if (p >= 0) {
                        const double *x = xp + 6 * p, *b = hpp + 27 * p + 21;
                        double xv[6];
#pragma unroll
                        for (int a = 0; a < 6; a++) { xv[a] = x[a]; sc += x[a] * (lambda * x[a] + b[a]); }
                        se3_oplus(T, xv);
#pragma unroll
                        for (int i = 0; i < 7; i++) D.kf[7 * tid + i] = T[i];
                    }
                }
                const double s = block_sum(sc, tmp);
                if (tid == 0) P.scl_part[vb] = s;
            }
            barrier();
            // ---- computeActiveErrors at the trial state (+ the quadratic form there) ------------------------------------------
            tick(4);
            build();
            tick(5);
            double tempChi = lg_total<false>(P.chi_part, P.nb_lin, tmp);
            const double scale = lg_total<false>(P.scl_part, P.nb_upd + 1, tmp);
            const bool ok2 = __ldcg(D.scal + 3) != 0.0;
            if (!ok2) tempChi = 1.7976931348623157e308;
            rho = (currentChi - tempChi) / (scale + 1e-3);
            trials++;
            rejected = !(rho > 0 && isfinite(tempChi));
            if (!rejected) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni; ni *= 2;
                for (int vb = cta; vb <= P.nb_upd; vb += G) {                                                   // pop, by the threads that pushed
                    if (vb < P.nb_upd) {
                        const int l = (vb * K2_THREADS + tid) >> 2;
                        if (l < D.n_pts && (tid & 3) == 0) { D.pt[3 * l] = P.pt_bak[3 * l]; D.pt[3 * l + 1] = P.pt_bak[3 * l + 1]; D.pt[3 * l + 2] = P.pt_bak[3 * l + 2]; }
                    } else if (tid < n_kf) {
#pragma unroll
                        for (int i = 0; i < 7; i++) D.kf[7 * tid + i] = P.kf_bak[7 * tid + i];
                    }
                }
            }
            qmax++;
            have = !rejected;
            if (rejected && rho < 0 && qmax < 10) {      // retried: the system at the restored estimates again
                barrier();
                build();
            }
            tick(-1);
        } while (rho < 0 && qmax < 10);
        if (rejected) barrier();             // the restored estimates are read by everybody in the next iteration (or by the host)
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
        if (nBad >= 3) break;
    }
    if (cta == 0 && tid == 0) {
        P.out[0] += (double)trials;
        for (int i = 0; i < 6; i++) P.out[2 + i] += (double)tph[i];
    }
}

static int g_grid_ctas = -1;      // CTAs of the cooperative kernel (one per SM), 0 = cooperative launch not available

size_t orbx_lba_fused_smem(int np);
bool orbx_lba_fused_fits(int n_kf, int np);

orbx_status orbx_lba_grid_init() {           // called for every handle: the shared-memory limit of a kernel is a per-device attribute
    if (g_grid_ctas >= 0) {
        if (g_grid_ctas > 0) ORBX_CUDA(ORBX_RAISE_SMEM(k_lba_grid));
        return ORBX_OK;
    }
    g_grid_ctas = 0;
    int dev = 0, coop = 0, sms = 0;
    ORBX_CUDA(cudaGetDevice(&dev));
    ORBX_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    ORBX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (!coop) return ORBX_OK;
    ORBX_CUDA(ORBX_RAISE_SMEM(k_lba_grid));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lba_grid, K2_THREADS, orbx_lba_fused_smem(36)) == cudaSuccess && per_sm >= 1)
        g_grid_ctas = sms;
    cudaGetLastError();
    return ORBX_OK;
}

bool orbx_lba_grid_available(int n_kf, int np) { return g_grid_ctas > 0 && orbx_lba_fused_fits(n_kf, np); }

// doubles of scratch the kernel needs for windows of up to n_pts landmarks and n_kf keyframes: three arrays of block-level partial sums,
// the pose step and the assembled reduced system
static int lg_part_slots(int n_pts) { return k2_blocks(4LL * n_pts) + 2; }
size_t orbx_lba_grid_scratch(int n_pts, int n_kf) {
    const size_t np = n_kf, n = 6 * np, nblk = np * (np + 1) / 2;
    return 3 * (size_t)lg_part_slots(n_pts) + n + nblk * 36 + n + (np + nblk + 3) / 2 + 1;      // the last term: the arrival counters (int)
}

orbx_status orbx_lba_grid_launch(const LbaDev &D, double *kf_bak, double *pt_bak, int iterations, int robust, int capture, double *cap_Hs,
                                 double *cap_bs, double *cap_xp, double *out, double *slots, int max_pts, int max_kf, cudaStream_t s) {
    LgParams P;
    P.D = D; P.kf_bak = kf_bak; P.pt_bak = pt_bak; P.iterations = iterations; P.robust = robust; P.capture = capture;
    P.cap_Hs = cap_Hs; P.cap_bs = cap_bs; P.cap_xp = cap_xp; P.out = out;
    const int ns = lg_part_slots(max_pts);
    P.chi_part = slots; P.scl_part = slots + ns; P.max_part = slots + 2 * ns;
    P.xp_g = slots + 3 * ns; P.hs_g = P.xp_g + D.n;
    {   // the counters sit behind the largest system the handle can hold, so that they stay zero whatever window ran before
        const size_t np = max_kf, n = 6 * np, nblk = np * (np + 1) / 2;
        P.done_kf = reinterpret_cast<int *>(slots + 3 * (size_t)lg_part_slots(max_pts) + n + nblk * 36 + n);
        P.done_blk = P.done_kf + np;
    }
    P.nb_lin = k2_blocks(4LL * D.n_pts);
    P.nb_hpp = D.n_kchunks > 0 ? k2_blocks((long long)D.n_kchunks * 32) : 0;
    P.nb_pairs = D.n_pchunks > 0 ? k2_blocks((long long)D.n_pchunks * 32) : 0;
    P.nb_upd = k2_blocks(4LL * D.n_pts);
    void *args[] = {&P};
    ORBX_CUDA(cudaLaunchCooperativeKernel((const void *)k_lba_grid, dim3(g_grid_ctas), dim3(K2_THREADS), args, orbx_lba_fused_smem(D.np), s));
    return ORBX_OK;
}


// computeActiveErrors (+ buildSystem when build != 0): chi2 sum lands in D.scal[0] (cleared here).
// want_hpp: also finish H_pp / b_p now (the caller needs them before the Schur step, e.g. for computeLambdaInit).
orbx_status orbx_lba_chunk_linearize(const LbaDev &D, int robust, int build, int want_hpp, cudaStream_t s, int *launches) {
    ORBX_CUDA(cudaMemsetAsync(D.scal, 0, sizeof(double), s));
    const int nb_lin = k2_blocks(4LL * D.n_pts), nb_hpp = build ? k2_blocks((long long)D.n_kchunks * 32) : 0;
    k2_build<<<nb_lin + nb_hpp, K2_THREADS, sizeof(double) * 12 * (D.n_kf > 0 ? D.n_kf : 1), s>>>(D, robust, build, nb_lin);
    *launches += 1;
    if (build && want_hpp) {
        k2_hpp_final_dinv<<<k2_blocks((long long)27 * D.np), K2_THREADS, 0, s>>>(D, 0.0, 1, 0);
        *launches += 1;
    }
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

// BlockSolver::solve up to the linear solve (also finishes H_pp / b_p from the partials of the last build)
orbx_status orbx_lba_chunk_schur(const LbaDev &D, double lambda, cudaStream_t s, int *launches) {
    const int nb_pairs = D.n_pchunks > 0 ? k2_blocks((long long)D.n_pchunks * 32) : 0;
    static const bool use_dmma = getenv("ORBX_LBA_DMMA") != nullptr;      // measurement only: the tensor-core form is slower
    if (use_dmma) {
        k2_hpp_final_dinv<<<k2_blocks(D.n_pts), K2_THREADS, 0, s>>>(D, lambda, 0, 1);
        k2_pairs_mma<<<nb_pairs + k2_blocks((long long)27 * D.np), K2_THREADS, 0, s>>>(D, nb_pairs);
        *launches += 1;
    } else
        k2_pairs<<<nb_pairs + k2_blocks((long long)27 * D.np), K2_THREADS, 0, s>>>(D, lambda, nb_pairs);
    k2_final<<<k2_blocks((long long)D.n * D.n > D.n ? (long long)D.n * D.n : D.n), K2_THREADS, 0, s>>>(D, lambda);
    *launches += 2;
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

orbx_status orbx_lba_chunk_init() {
    ORBX_CUDA(ORBX_RAISE_SMEM(k2_build));
    return ORBX_OK;
}
