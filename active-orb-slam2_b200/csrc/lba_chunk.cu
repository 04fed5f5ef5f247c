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

__device__ __forceinline__ void k2_linearize_body(const LbaDev &D, int robust, int build, int block, double *kfRt, double *tmp) {
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
    if (threadIdx.x == 0 && s != 0) atomicAdd(&D.scal[0], s);
}

__device__ __forceinline__ void k2_hpp_body(const LbaDev &D, int robust, int block) {
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
    const int lane = threadIdx.x & 31, ch = blockIdx.x * (K2_THREADS / 32) + (threadIdx.x >> 5);
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
