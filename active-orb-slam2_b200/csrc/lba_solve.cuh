// The reduced camera system of local BA, solved by ONE thread block in its shared memory (lba_fused.cu: CTA 0 of the cluster;
// lba_grid.cu: CTA 0 of the grid).  Replaces g2o's LinearSolverEigen / Dense on H_schur (block_solver.hpp:441-468).
//   layout: hs[nblk][36] the upper block triangle (block (i, j), i <= j, at upper_block(i, j, np), 6x6 row-major), then b_schur [n],
//           then the block table [np (np + 1) / 2] uint16 (inside the next n doubles).
//   A = U^T U on the upper block triangle with the right-hand side carried along (U^T y = b comes out of the factorisation), then
//   U x = y.  A diagonal block is replaced by W_k = U_kk^-1 as soon as it is factorised, so that everything after the 6x6 pivot is
//   products, not triangular solves: U_kj = W_k^T A_kj and y_k = W_k^T b_k (one thread per column), A_ij -= U_ki^T U_kj and
//   b_j -= U_kj^T y_k (one thread per row of a block; the blocks (i, j) of a step are a prefix of a table ordered by descending i,
//   so no thread searches for its block), x_k = W_k y_k.
#pragma once
#include "lba_common.cuh"

// the blocks (i, j), i <= j, ordered by descending i: the trailing update of pivot k touches the first T (T + 1) / 2, T = np - 1 - k
template <int THREADS>
__device__ __forceinline__ void lba_solve_table(double *hs, int np) {
    const int nblk = np * (np + 1) / 2, n = 6 * np;
    uint16_t *ptab = reinterpret_cast<uint16_t *>(hs + nblk * 36 + n);
    for (int pr = threadIdx.x; pr < nblk; pr += THREADS) {
        int r = 0;
        while ((r + 1) * (r + 2) / 2 <= pr) r++;
        const int bi = np - 1 - r, bj = bi + (pr - r * (r + 1) / 2);
        ptab[pr] = (uint16_t)(bi | (bj << 8));
    }
}

// the 6x6 pivot block by ONE thread in registers: U = chol(A) (upper, A = U^T U) and W = U^-1 written over the block's upper triangle
// (nothing reads the lower one).  Every loop has constant bounds with compile-time guards so that the arrays stay in registers.
// The chain is 6 x (rsqrt + two dependent operations) for U and 5 short levels for W: ~0.4 us, against ~0.9 us for the lane-per-column
// form with its 21 + 15 shuffles.  Returns false if the block is not positive definite.
__device__ __forceinline__ bool lba_pivot6(double *blk) {
    double A[6][6], U[6][6], W[6][6], is[6];
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = 0; b < 6; b++)
            if (b >= a) A[a][b] = blk[6 * a + b];
    bool good = true;
#pragma unroll
    for (int a = 0; a < 6; a++) {
        double d = A[a][a];
#pragma unroll
        for (int m = 0; m < 6; m++)
            if (m < a) d -= U[m][a] * U[m][a];
        good = good && (d > 0);
        is[a] = rsqrt(d);
#pragma unroll
        for (int b = 0; b < 6; b++)
            if (b > a) {
                double v = A[a][b];
#pragma unroll
                for (int m = 0; m < 6; m++)
                    if (m < a) v -= U[m][a] * U[m][b];
                U[a][b] = v * is[a];
            }
    }
#pragma unroll
    for (int b = 0; b < 6; b++) {
        W[b][b] = is[b];
#pragma unroll
        for (int aa = 0; aa < 6; aa++) {
            const int a = 5 - aa;
            if (a < b) {
                double v = 0;
#pragma unroll
                for (int m = 0; m < 6; m++)
                    if (m > a && m <= b) v -= U[a][m] * W[m][b];
                W[a][b] = v * is[a];
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = 0; b < 6; b++)
            if (b >= a) blk[6 * a + b] = W[a][b];
    return good;
}

// called by every thread of the block; *ok (shared memory) ends 0 if a pivot block was not positive definite (then xp = 0: no step, the
// trial is rejected).  The solution lands in xp [n] (shared memory of the same block).
// The pivot of step k + 1 is taken off the critical path: in the trailing update of step k warp 0 updates block (k + 1, k + 1) first and
// factorises it while the other warps do the rest of the update.
template <int THREADS>
__device__ __forceinline__ void lba_reduced_solve(double *hs, double *xp, int np, int *ok) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = np * (np + 1) / 2, n = 6 * np;
    int &sh_ok = *ok;
    {
        double *bs = hs + nblk * 36;                                  // b_schur, then y, then x
        uint16_t *ptab = reinterpret_cast<uint16_t *>(hs + nblk * 36 + n);   // [nblk] (i | j << 8), rows i = np - 1, np - 2, ...
        if (tid == 0) sh_ok = lba_pivot6(hs) ? 1 : 0;                 // block (0, 0)
        __syncthreads();
        // row a of block (bi, bj) -= U_k,bi^T U_k,bj
        auto pair_row = [&](int k, int rowk, int pr, int a) {
            const unsigned e = ptab[pr];
            const int bi = e & 0xff, bj = e >> 8;
            const double *Uki = hs + (rowk + (bi - k)) * 36 + a, *Ukj = hs + (rowk + (bj - k)) * 36;
            double *out = hs + upper_block(bi, bj, np) * 36 + 6 * a;
            double u[6], v[6];
#pragma unroll
            for (int m = 0; m < 6; m++) u[m] = Uki[6 * m];
#pragma unroll
            for (int c2 = 0; c2 < 6; c2++) v[c2] = out[c2];
#pragma unroll
            for (int m = 0; m < 6; m++)
#pragma unroll
                for (int c2 = 0; c2 < 6; c2++) v[c2] -= u[m] * Ukj[6 * m + c2];
#pragma unroll
            for (int c2 = 0; c2 < 6; c2++) out[c2] = v[c2];
        };
        for (int k = 0; k < np; k++) {
            if (!sh_ok) break;
            const int rowk = upper_block(k, k, np);
            const double *Ukk = hs + rowk * 36;                        // holds W_k by now
            const int T = np - k - 1;
            for (int t = tid; t < T * 6 + 1; t += THREADS) {    // U_kj = W^T A_kj column by column; the last item is y_k = W^T b_k
                const bool rhs = t == T * 6;
                const int j = k + 1 + t / 6, b = t % 6;
                double *col = rhs ? bs + 6 * k : hs + (rowk + (j - k)) * 36 + b;
                const int st = rhs ? 1 : 6;
                double o[6], y[6];
#pragma unroll
                for (int a = 0; a < 6; a++) o[a] = col[st * a];
#pragma unroll
                for (int a = 0; a < 6; a++) {
                    double v = 0;
#pragma unroll
                    for (int m = 0; m < 6; m++)
                        if (m <= a) v += Ukk[6 * m + a] * o[m];
                    y[a] = v;
                }
#pragma unroll
                for (int a = 0; a < 6; a++) col[st * a] = y[a];
            }
            __syncthreads();
            if (T > 0) {
                // A_ij -= U_ki^T U_kj for k < i <= j (row a of a block per item) and b_j -= U_kj^T y_k.  Block (k + 1, k + 1) is the
                // first entry of the table's last row in use
                const int npair = T * (T + 1) / 2, pr_la = T * (T - 1) / 2, nit = (npair - 1) * 6 + T;
                if (warp == 0) {
                    if (lane < 6) pair_row(k, rowk, pr_la, lane);
                    __syncwarp();
                    if (lane == 0 && !lba_pivot6(hs + upper_block(k + 1, k + 1, np) * 36)) sh_ok = 0;
                } else {
                    // (a whole block per thread -- a third of the shared-memory traffic -- was measured slower: 553 -> 608 us per C3 window)
                    for (int t = tid - 32; t < nit; t += THREADS - 32) {
                        if (t < (npair - 1) * 6) {
                            const int q = t < pr_la * 6 ? t : t + 6;    // the six items of block (k + 1, k + 1) are warp 0's
                            pair_row(k, rowk, q / 6, q % 6);
                        } else {
                            const int j = k + 1 + (t - (npair - 1) * 6);
                            const double *Ukj = hs + (rowk + (j - k)) * 36;
                            double v[6];
#pragma unroll
                            for (int c2 = 0; c2 < 6; c2++) v[c2] = bs[6 * j + c2];
#pragma unroll
                            for (int m = 0; m < 6; m++)
#pragma unroll
                                for (int c2 = 0; c2 < 6; c2++) v[c2] -= Ukj[6 * m + c2] * bs[6 * k + m];
#pragma unroll
                            for (int c2 = 0; c2 < 6; c2++) bs[6 * j + c2] = v[c2];
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (sh_ok) {
            for (int k = np - 1; k >= 0; k--) {                      // backward: x_k = W_k y_k, then y_i -= U_ik x_k for i < k
                const double *Wk = hs + upper_block(k, k, np) * 36;
                double xk[6];
#pragma unroll
                for (int a = 0; a < 6; a++) {                          // every thread: 21 products from shared memory, no barrier for x_k
                    double v = 0;
#pragma unroll
                    for (int m = 0; m < 6; m++)
                        if (m >= a) v += Wk[6 * a + m] * bs[6 * k + m];
                    xk[a] = v;
                }
                if (tid < 6) {                                         // x goes to its own array: nobody waits before y_k may be overwritten
                    double mine = xk[0];                                // static indices only: xk stays in registers
#pragma unroll
                    for (int a = 1; a < 6; a++) mine = tid == a ? xk[a] : mine;
                    xp[6 * k + tid] = mine;
                }
                for (int t = tid; t < k * 6; t += THREADS) {
                    const int i = t / 6, a = t - 6 * i;
                    const double *Uik = hs + (upper_block(i, i, np) + (k - i)) * 36 + 6 * a;
                    double v = bs[6 * i + a];
#pragma unroll
                    for (int c2 = 0; c2 < 6; c2++) v -= Uik[c2] * xk[c2];
                    bs[6 * i + a] = v;
                }
                __syncthreads();
            }
        } else {
            for (int i = tid; i < n; i += THREADS) xp[i] = 0;     // failed factorisation: no step, the trial is rejected
        }
    }
    __syncthreads();
}
