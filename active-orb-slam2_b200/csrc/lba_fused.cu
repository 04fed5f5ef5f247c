// One optimizer.optimize(iterations) call of Optimizer::LocalBundleAdjustment (reference src/Optimizer.cc:659-706)
// in ONE kernel: a thread-block cluster of 8 CTAs owns the window, the Levenberg loop of
// g2o core/optimization_algorithm_levenberg.cpp:61-189 runs on the device, and the CTAs meet at cluster barriers
// instead of going back to the host between stages (the multi-kernel path of lba.cu pays ~7 launches and one host
// round trip per Levenberg trial).  Same arithmetic, same f64, same edge order per landmark as lba.cu.
//
//   landmarks (and with them the edges, which are sorted by landmark) are split into 8 contiguous ranges, one per CTA;
//   H_ll / b_l of a landmark are accumulated in registers by the thread that owns it;
//   H_pp / b_p and the Schur complement are sums over lists the host builds once per window (edges by keyframe; pairs of
//   edges of one landmark by pose-pair block), cut into chunks of 64: a warp reduces a chunk in registers and writes one
//   partial block, CTA 0 adds the partials of a block in chunk order.  No atomics anywhere, so the sums are
//   deterministic and every CTA takes the same Levenberg decision from the same numbers without a broadcast;
//   the reduced camera system lives in CTA 0's shared memory in upper block-triangular layout and is factorised there
//   (blocked 6x6 Cholesky, A = U^T U, diagonal blocks kept as their inverses, right-hand side carried along) without ever
//   touching global memory.
// Used when the upper block triangle of H_schur fits shared memory (<= 36 free keyframes); larger windows take lba.cu.
#include <cooperative_groups.h>
#include "lba_common.cuh"
#include "lba_solve.cuh"

namespace cg = cooperative_groups;

#define LF_THREADS 256
#define LF_WARPS (LF_THREADS / 32)
#define LF_CTAS 8          // portable cluster size: many windows in flight
#define LF_CTAS_WIDE 16    // opt-in (non-portable) cluster size: one window at a time, twice the warps per phase
#define LF_MAX_KF 64
#define LF_SLOTS 4

struct LfParams {
    LbaDev D;
    double *kf_bak, *pt_bak;
    int iterations, robust;
    int capture;                      // 1: store the first trial's reduced system
    double *cap_Hs, *cap_bs, *cap_xp; // n x n (full symmetric), n, n
    double *out;                      // [0] trials, [1] lambda of the first trial, [2..7] += nanoseconds per phase
};

struct LfShared {
    double red[LF_SLOTS][LF_CTAS_WIDE][4];   // cluster reductions land in CTA 0's copy
    double tmp[32];
    double bc[4];                        // values gathered by thread 0 for the whole CTA
    int ok;
};

__device__ __forceinline__ double block_max(double v, double *tmp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) tmp[w] = v;
    __syncthreads();
    double s = 0;
    if (w == 0) {
        s = lane < (int)(blockDim.x >> 5) ? tmp[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
    }
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(LF_THREADS, 1) k_lba_fused(LfParams P) {
    extern __shared__ __align__(16) double dyn[];
    __shared__ LfShared sh;
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank(), C = (int)cl.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const LbaDev &D = P.D;
    const int np = D.np, n = D.n, nblk = np * (np + 1) / 2, n_kf = D.n_kf;

    // shared-memory carve-up (identical in every CTA so that map_shared_rank addresses line up)
    double *kfRt = dyn;                               // [n_kf][12]  R (row-major) and t of every keyframe
    double *hpp = kfRt + 12 * LF_MAX_KF;              // [np][27]    H_pp / b_p (used in CTA 0)
    double *hs = hpp + 27 * np;                       // [nblk][36] + [n]  the reduced system (used in CTA 0)
    double *xp = hs + nblk * 36 + 2 * n;              // [n]   (hs is followed by b_schur [n] and the block table of the solve, np (np + 1) bytes in [n] doubles)
    LfShared *sh0 = cl.map_shared_rank(&sh, 0);
    double *hs0 = cl.map_shared_rank(hs, 0), *xp0 = cl.map_shared_rank(xp, 0);
    const int total_warps = C * LF_WARPS, gwarp = rank * LF_WARPS + warp;

    const int l0 = (int)((long long)D.n_pts * rank / C), l1 = (int)((long long)D.n_pts * (rank + 1) / C);
    int slot = 0;
    if (rank == 0) lba_solve_table<LF_THREADS>(hs, np);
    cl.sync();                       // every CTA of the cluster is running before anybody stores into CTA 0's shared memory

    // residuals (+ optionally the quadratic form) of this CTA's landmarks; returns nothing, partial sums go to CTA 0
    auto linearize = [&](bool build) {
        for (int k = tid; k < n_kf; k += LF_THREADS) {
            double R[9];
            double T[7];      // poses are written by CTA 0 on another SM: read them past the (incoherent) L1
            for (int i = 0; i < 7; i++) T[i] = __ldcg(D.kf + 7 * k + i);
            quat_to_R(T, R);
            for (int i = 0; i < 9; i++) kfRt[12 * k + i] = R[i];
            kfRt[12 * k + 9] = T[4]; kfRt[12 * k + 10] = T[5]; kfRt[12 * k + 11] = T[6];
        }
        __syncthreads();
        double chi = 0, mx = 0;
        for (int l = l0 + tid; l < l1; l += LF_THREADS) {
            double hl[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int e = D.ptstart[l]; e < D.ptstart[l + 1]; e++) {
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
                if (P.robust) {
                    const double d = D.stereo[e] ? D.d_stereo : D.d_mono, dsqr = (double)(float)(d * d);   // RobustKernelHuber keeps dsqr in a float member (robust_kernel_impl.h:84)
                    if (c > dsqr) { const double sq = sqrt(c); cr = 2 * sq * d - dsqr; rho1 = d / sq; }
                }
                chi += cr;
                if (!build) continue;
                const int dim = D.stereo[e] ? 3 : 2;
                // one reciprocal per edge instead of ~20 f64 divisions (results move by an ulp or two; tolerance 1e-4)
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
                const int ip = D.kfidx[kf];
                if (ip >= 0) {
                    double *hpl = D.Hpl + 18 * (size_t)e;
                    for (int a = 0; a < 6; a++)
                        for (int b = 0; b < 3; b++) hpl[3 * a + b] = w * (B[a] * A[b] + B[6 + a] * A[3 + b] + B[12 + a] * A[6 + b]);
                }
            }
            if (build) {
                for (int i = 0; i < 9; i++) D.Hll[9 * l + i] = hl[i];
                mx = fmax(mx, fmax(fabs(hl[0]), fmax(fabs(hl[3]), fabs(hl[5]))));
            }
        }
        const double s = block_sum(chi, sh.tmp);
        const double m = build ? block_max(mx, sh.tmp) : 0.0;
        if (tid == 0) { sh0->red[slot][rank][0] = s; sh0->red[slot][rank][1] = m; }
    };

    double lambda = 0, ni = 2;
    int nBad = 0, trials = 0;
    bool first = P.capture != 0;
    // phase timers (ns, CTA 0 thread 0): build, schur, reduce, solve, update, err  -> out[2..7]
    unsigned long long tph[6] = {0, 0, 0, 0, 0, 0}, tlast = 0;
    auto tick = [&](int ph) {
        if (rank == 0 && tid == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (ph >= 0) tph[ph] += now - tlast;
            tlast = now;
        }
    };

    for (int it = 0; it < P.iterations; it++) {
        // ---- computeActiveErrors + buildSystem ------------------------------------------------------------------
        tick(-1);
        linearize(true);
        __threadfence();
        cl.sync();
        // H_pp / b_p = sum over the edges of a keyframe of B^T W B / B^T W r: chunks of the keyframe-ordered edge list, one
        // warp per chunk; B is recomputed from the stored error (cheaper than keeping 27 doubles per edge)
        for (int ch = gwarp; ch < D.n_kchunks; ch += total_warps) {
            const int4 cd = D.kchunk[ch];                   // x = keyframe (reduced index), y = first entry, z = entries
            double acc[27];
#pragma unroll
            for (int i = 0; i < 27; i++) acc[i] = 0;
            for (int q = lane; q < cd.z; q += 32) {
                const int4 ke = D.kfe[cd.y + q];              // edge, keyframe, landmark, stereo: one round trip for the rest
                const int e = ke.x;
                const uint8_t off = D.level1[e];
                const double *R = kfRt + 12 * ke.y;
                const double *X = D.pt + 3 * ke.z;
                const double X0 = __ldcg(X), X1 = __ldcg(X + 1), X2 = __ldcg(X + 2);
                const double info = D.info[e], c = __ldcg(D.chi2 + e);
                const double e0 = __ldcg(D.err + 3 * e), e1 = __ldcg(D.err + 3 * e + 1), e2 = __ldcg(D.err + 3 * e + 2);
                if (off) continue;
                const double x = R[0] * X0 + R[1] * X1 + R[2] * X2 + R[9], y = R[3] * X0 + R[4] * X1 + R[5] * X2 + R[10];
                const double iz = 1.0 / (R[6] * X0 + R[7] * X1 + R[8] * X2 + R[11]), iz2 = iz * iz, xz = x * iz, yz = y * iz;
                const double fx = D.fx, fy = D.fy, bf = D.bf;
                const bool st = ke.w != 0;
                double B[18];
                B[0] = xz * yz * fx; B[1] = -(1 + xz * xz) * fx; B[2] = yz * fx; B[3] = -iz * fx; B[4] = 0; B[5] = xz * iz * fx;
                B[6] = (1 + yz * yz) * fy; B[7] = -xz * yz * fy; B[8] = -xz * fy; B[9] = 0; B[10] = -iz * fy; B[11] = yz * iz * fy;
                if (st) { B[12] = B[0] - bf * y * iz2; B[13] = B[1] + bf * x * iz2; B[14] = B[2]; B[15] = B[3]; B[16] = 0; B[17] = B[5] - bf * iz2; }
                else { for (int i = 12; i < 18; i++) B[i] = 0; }
                double rho1 = 1.0;
                if (P.robust) {
                    const double d = st ? D.d_stereo : D.d_mono;
                    if (c > d * d) rho1 = d / sqrt(c);
                }
                const double w = rho1 * info;
                const double w0 = -info * e0 * rho1, w1 = -info * e1 * rho1, w2 = -info * e2 * rho1;
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
        __threadfence();
        cl.sync();
        if (rank == 0) {     // partials of a keyframe in chunk order
            for (int i = tid; i < 27 * np; i += LF_THREADS) {
                const int p = i / 27, c = i - 27 * p;
                double v = 0;
                for (int ch = D.kf_cstart[p]; ch < D.kf_cstart[p + 1]; ch++) v += __ldcg(D.hppart + (size_t)ch * 27 + c);
                hpp[i] = v;
            }
            __syncthreads();
            if (it == 0) {   // computeLambdaInit also looks at the pose diagonals
                double mx = 0;
                for (int i = tid; i < 6 * np; i += LF_THREADS) {
                    const int a = i % 6;                                  // diagonal entry a of the packed upper triangle: 0, 6, 11, 15, 18, 20
                    mx = fmax(mx, fabs(hpp[27 * (i / 6) + 6 * a - a * (a - 1) / 2]));
                }
                mx = block_max(mx, sh.tmp);
                if (tid == 0) sh.red[slot][0][2] = mx;
            }
        }
        cl.sync();
        if (tid == 0) {      // every CTA sums the same numbers in the same order: identical decisions without a broadcast
            double chi = 0, mx = 0;
            for (int r = 0; r < C; r++) { chi += sh0->red[slot][r][0]; mx = fmax(mx, sh0->red[slot][r][1]); }
            if (it == 0) mx = fmax(mx, sh0->red[slot][0][2]);
            sh.bc[0] = chi; sh.bc[1] = mx;
        }
        __syncthreads();
        double currentChi = sh.bc[0];
        const double maxdiag = sh.bc[1];
        slot = (slot + 1) % LF_SLOTS;
        if (it == 0) { lambda = 1e-5 * maxdiag; ni = 2; nBad = 0; }
        tick(0);
        const double iniChi = currentChi;
        double rho = 0;
        int qmax = 0;
        do {
            // ---- push -----------------------------------------------------------------------------------------
            for (int i = 3 * l0 + tid; i < 3 * l1; i += LF_THREADS) P.pt_bak[i] = D.pt[i];
            if (rank == 0) for (int i = tid; i < 7 * n_kf; i += LF_THREADS) P.kf_bak[i] = D.kf[i];
            // ---- Schur complement (block_solver.hpp:371-439) -------------------------------------------------------------
            // (1) D^-1 = (H_ll + lambda I)^-1 and D^-1 b_l of this CTA's landmarks
            for (int l = l0 + tid; l < l1; l += LF_THREADS) {
                const double *hl = D.Hll + 9 * l;
                double Di[6];
                dinv3(hl, lambda, Di);
                double *o = D.dinv + 10 * (size_t)l;
                for (int i = 0; i < 6; i++) o[i] = Di[i];
                o[6] = Di[0] * hl[6] + Di[1] * hl[7] + Di[2] * hl[8];
                o[7] = Di[1] * hl[6] + Di[3] * hl[7] + Di[4] * hl[8];
                o[8] = Di[2] * hl[6] + Di[4] * hl[7] + Di[5] * hl[8];
            }
            __threadfence();
            cl.sync();
            // (2) every (edge, edge) pair of a landmark contributes B_i D^-1 B_j^T to block (pose_i, pose_j): chunks of the
            //     block-ordered pair list, one warp per chunk, partial block per chunk
            for (int ch = gwarp; ch < D.n_pchunks; ch += total_warps) {
                const int4 cd = D.pchunk[ch];               // x = block, y = first entry, z = entries, w = 1 for a diagonal block
                double acc[42];
#pragma unroll
                for (int i = 0; i < 42; i++) acc[i] = 0;
                for (int q = lane; q < cd.z; q += 32) {
                    const int4 pe = D.pairs[cd.y + q];          // edge i, edge j, landmark: everything below is one round trip
                    const uint8_t off1 = D.level1[pe.x], off2 = D.level1[pe.y];
                    const double2 *dv = reinterpret_cast<const double2 *>(D.dinv + 10 * (size_t)pe.z);
                    const double2 *B1 = reinterpret_cast<const double2 *>(D.Hpl + 18 * (size_t)pe.x);
                    const double2 *B2 = reinterpret_cast<const double2 *>(D.Hpl + 18 * (size_t)pe.y);
                    double Di[10], b1[18], b2[18];
#pragma unroll
                    for (int i = 0; i < 5; i++) { const double2 t = __ldcg(dv + i); Di[2 * i] = t.x; Di[2 * i + 1] = t.y; }
#pragma unroll
                    for (int i = 0; i < 9; i++) { const double2 t = __ldcg(B1 + i); b1[2 * i] = t.x; b1[2 * i + 1] = t.y; }
#pragma unroll
                    for (int i = 0; i < 9; i++) { const double2 t = __ldcg(B2 + i); b2[2 * i] = t.x; b2[2 * i + 1] = t.y; }
                    if (off1 || off2) continue;
#pragma unroll
                    for (int a = 0; a < 6; a++) {
                        const double u0 = b1[3 * a], u1 = b1[3 * a + 1], u2 = b1[3 * a + 2];
                        const double bd0 = u0 * Di[0] + u1 * Di[1] + u2 * Di[2], bd1 = u0 * Di[1] + u1 * Di[3] + u2 * Di[4],
                                     bd2 = u0 * Di[2] + u1 * Di[4] + u2 * Di[5];
#pragma unroll
                        for (int b = 0; b < 6; b++) acc[6 * a + b] += bd0 * b2[3 * b] + bd1 * b2[3 * b + 1] + bd2 * b2[3 * b + 2];
                        if (cd.w) acc[36 + a] += u0 * Di[6] + u1 * Di[7] + u2 * Di[8];     // coefficients: B_i D^-1 b_l (:413)
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
            __threadfence();
            cl.sync();
            tick(1);
            // (3) H_schur = H_pp + lambda I - sum, b_schur = b_p - sum, partials of a block in chunk order: every CTA finalises a
            //     share of the entries and stores them into CTA 0's shared memory (distributed shared memory), where the solve runs
            {
                const double *hpp0 = cl.map_shared_rank(hpp, 0);
                for (int i = rank * LF_THREADS + tid; i < nblk * 36; i += C * LF_THREADS) {
                    const int blk = i / 36, ab = i - 36 * blk;
                    double v = 0;
                    for (int ch = D.blk_cstart[blk]; ch < D.blk_cstart[blk + 1]; ch++) v -= __ldcg(D.part + (size_t)ch * 42 + ab);
                    int p1 = 0, rem = blk;
                    while (rem >= np - p1) { rem -= np - p1; p1++; }
                    if (rem == 0) {        // diagonal block
                        int a = ab / 6, b = ab - 6 * a;
                        const bool dg = a == b;
                        if (a > b) { const int t = a; a = b; b = t; }
                        v += hpp0[27 * p1 + a * 6 - a * (a - 1) / 2 + (b - a)] + (dg ? lambda : 0.0);
                    }
                    hs0[i] = v;
                }
                for (int k = rank * LF_THREADS + tid; k < n; k += C * LF_THREADS) {
                    const int p = k / 6, a = k - 6 * p, blk = upper_block(p, p, np);
                    double v = hpp0[27 * p + 21 + a];
                    for (int ch = D.blk_cstart[blk]; ch < D.blk_cstart[blk + 1]; ch++) v -= __ldcg(D.part + (size_t)ch * 42 + 36 + a);
                    hs0[nblk * 36 + k] = v;
                }
            }
            cl.sync();
            if (first && P.cap_Hs) {     // parity tests: the very first reduced system, expanded to a full symmetric matrix
                for (int i = rank * LF_THREADS + tid; i < n * n; i += C * LF_THREADS) {
                    int r = i / n, c = i - r * n;
                    if (r / 6 > c / 6) { const int t = r; r = c; c = t; }
                    P.cap_Hs[i] = hs0[upper_block(r / 6, c / 6, np) * 36 + 6 * (r % 6) + c % 6];
                }
                for (int i = rank * LF_THREADS + tid; i < n; i += C * LF_THREADS) P.cap_bs[i] = hs0[nblk * 36 + i];
                cl.sync();
            }
            tick(2);
            // ---- reduced solve in CTA 0 (lba_solve.cuh) ----------------------------------------------------------------------
            if (rank == 0) {
                lba_reduced_solve<LF_THREADS>(hs, xp, np, &sh.ok);
                __syncthreads();
                if (tid == 0) sh.red[slot][0][2] = sh.ok ? 1.0 : 0.0;
            }
            cl.sync();
            tick(3);
            if (rank != 0) for (int i = tid; i < n; i += LF_THREADS) xp[i] = xp0[i];
            __syncthreads();
            if (first && P.cap_xp) for (int i = rank * LF_THREADS + tid; i < n; i += C * LF_THREADS) P.cap_xp[i] = xp[i];
            if (first && tid == 0 && rank == 0) P.out[1] = lambda;
            first = false;
            // ---- landmark back-substitution, update, computeScale ------------------------------------------------------
            double sc = 0;
            for (int l = l0 + tid; l < l1; l += LF_THREADS) {
                const double *hl = D.Hll + 9 * l;
                double c0 = hl[6], c1 = hl[7], c2 = hl[8];
                for (int e = D.ptstart[l]; e < D.ptstart[l + 1]; e++) {
                    const int p = D.kfidx[D.ekf[e]];
                    if (p < 0 || D.level1[e]) continue;
                    const double *B = D.Hpl + 18 * (size_t)e, *x = xp + 6 * p;
                    for (int a = 0; a < 6; a++) { c0 -= B[3 * a] * x[a]; c1 -= B[3 * a + 1] * x[a]; c2 -= B[3 * a + 2] * x[a]; }
                }
                double Di[6];
                dinv3(hl, lambda, Di);
                const double x0 = Di[0] * c0 + Di[1] * c1 + Di[2] * c2, x1 = Di[1] * c0 + Di[3] * c1 + Di[4] * c2,
                             x2 = Di[2] * c0 + Di[4] * c1 + Di[5] * c2;
                sc += x0 * (lambda * x0 + hl[6]) + x1 * (lambda * x1 + hl[7]) + x2 * (lambda * x2 + hl[8]);
                D.pt[3 * l] += x0; D.pt[3 * l + 1] += x1; D.pt[3 * l + 2] += x2;
            }
            if (rank == 0 && tid < n_kf) {
                const int p = D.kfidx[tid];
                if (p >= 0) {
                    const double *x = xp + 6 * p, *b = hpp + 27 * p + 21;
                    for (int a = 0; a < 6; a++) sc += x[a] * (lambda * x[a] + b[a]);
                    se3_oplus(D.kf + 7 * tid, x);
                }
            }
            const double scs = block_sum(sc, sh.tmp);
            if (tid == 0) sh0->red[slot][rank][3] = scs;
            __threadfence();
            cl.sync();
            // ---- computeActiveErrors at the trial state ---------------------------------------------------------------------
            tick(4);
            linearize(false);
            cl.sync();
            tick(5);
            if (tid == 0) {
                double chi = 0, scl = 0;
                for (int r = 0; r < C; r++) { chi += sh0->red[slot][r][0]; scl += sh0->red[slot][r][3]; }
                sh.bc[0] = chi; sh.bc[2] = scl; sh.bc[3] = sh0->red[slot][0][2];
            }
            __syncthreads();
            double tempChi = sh.bc[0];
            const double scale = sh.bc[2];
            const bool ok2 = sh.bc[3] != 0.0;
            slot = (slot + 1) % LF_SLOTS;
            if (!ok2) tempChi = 1.7976931348623157e308;
            rho = (currentChi - tempChi) / (scale + 1e-3);
            trials++;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni; ni *= 2;
                for (int i = 3 * l0 + tid; i < 3 * l1; i += LF_THREADS) D.pt[i] = P.pt_bak[i];            // pop
                if (rank == 0) for (int i = tid; i < 7 * n_kf; i += LF_THREADS) D.kf[i] = P.kf_bak[i];
            }
            qmax++;
            __threadfence();
            cl.sync();
            tick(-1);
        } while (rho < 0 && qmax < 10);
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++; else nBad = 0;
        if (nBad >= 3) break;
    }
    if (rank == 0 && tid == 0) {
        P.out[0] += (double)trials;
        for (int i = 0; i < 6; i++) P.out[2 + i] += (double)tph[i];
    }
    cl.sync();     // no CTA may exit while others still read its shared memory
}

// ---- host glue (called from lba.cu) ---------------------------------------------------------------------------------
size_t orbx_lba_fused_smem(int np) {
    const size_t nblk = (size_t)np * (np + 1) / 2, n = 6 * (size_t)np;
    return sizeof(double) * (12 * LF_MAX_KF + 27 * (size_t)np + nblk * 36 + 3 * n);
}

bool orbx_lba_fused_fits(int n_kf, int np) { return n_kf <= LF_MAX_KF && np >= 1 && orbx_lba_fused_smem(np) <= 200 * 1024; }

static int g_wide_ok = -1;     // can this device co-schedule a 16-CTA cluster of this kernel?

orbx_status orbx_lba_fused_init() {
    ORBX_CUDA(ORBX_RAISE_SMEM(k_lba_fused));       // per-device attributes: set for every handle's device
    if (g_wide_ok > 0) cudaFuncSetAttribute(k_lba_fused, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (g_wide_ok < 0) {
        g_wide_ok = 0;
        if (cudaFuncSetAttribute(k_lba_fused, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(LF_CTAS_WIDE); cfg.blockDim = dim3(LF_THREADS); cfg.dynamicSmemBytes = orbx_lba_fused_smem(36);
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = LF_CTAS_WIDE; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, k_lba_fused, &cfg) == cudaSuccess && n >= 1) g_wide_ok = 1;
        }
        cudaGetLastError();
    }
    return ORBX_OK;
}

orbx_status orbx_lba_fused_launch(const LbaDev &D, double *kf_bak, double *pt_bak, int iterations, int robust, int capture,
                                  double *cap_Hs, double *cap_bs, double *cap_xp, double *out, cudaStream_t s, int wide) {
    static const bool no_wide = getenv("ORBX_LBA_CLUSTER8") != nullptr;
    const int ctas = (wide && g_wide_ok == 1 && !no_wide) ? LF_CTAS_WIDE : LF_CTAS;
    LfParams P;
    P.D = D; P.kf_bak = kf_bak; P.pt_bak = pt_bak; P.iterations = iterations; P.robust = robust; P.capture = capture;
    P.cap_Hs = cap_Hs; P.cap_bs = cap_bs; P.cap_xp = cap_xp; P.out = out;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(LF_THREADS);
    cfg.dynamicSmemBytes = orbx_lba_fused_smem(D.np);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ORBX_CUDA(cudaLaunchKernelEx(&cfg, k_lba_fused, P));
    return ORBX_OK;
}
