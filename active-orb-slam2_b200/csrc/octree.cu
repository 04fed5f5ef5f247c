// Quadtree keypoint distribution: replaces ORBextractor::DistributeOctTree + ExtractorNode::DivideNode
// (reference src/ORBextractor.cc:539-763, :481-537).
//
// The reference walks a std::list of nodes and re-buckets keypoint vectors at every split.  Here one CTA owns
// one (frame, level) and works on ranges of a sorted array instead (tests/octree_model.py is the executable
// model of this formulation and is checked against the oracle):
//   * node boundaries depend on the level geometry only, so every coordinate has a fixed split PATH per axis
//     (host-built LUTs); interleaving the x and y paths under the initial-node index gives a key whose
//     prefixes name the quadtree nodes.  After sorting by key every node of every depth is a contiguous range
//     and DivideNode is three binary searches.
//   * std::list::push_front keeps the list in descending creation order; each node carries a stamp
//     (round, 4*rank_of_parent + quadrant) and every decision of the reference (pass order, the (size, pointer)
//     sort of the last phase, the early break at N nodes, the output order) is a function of those stamps.
//     Node "pointer order" is defined as creation order, as in the oracle.
// The sort is a counting sort over the top key bits (shared-memory histogram) followed by an insertion sort
// of the few keys that share a cell.
#include "orbx_internal.cuh"
#include "block_scan.cuh"

#define OCT_THREADS 256
#define OCT_ARRAYS 16

struct OctShared {
    int nfin, nact, size, phase, finish, cnt_lt;
    int warp_tmp[34];
};

__device__ __forceinline__ int lower_bound_u32(const uint32_t *k, int lo, int hi, uint32_t t) {
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (k[m] < t) lo = m + 1; else hi = m;
    }
    return lo;
}

// the three inner boundaries of the children of the node [s,e) whose keys share the top bits above (sh+2)
__device__ __forceinline__ void child_cuts(const uint32_t *K, int s, int e, int sh, int &c1, int &c2, int &c3) {
    const uint32_t p4 = (K[s] >> (sh + 2)) << 2;
    c2 = lower_bound_u32(K, s, e, (p4 | 2u) << sh);
    c1 = lower_bound_u32(K, s, c2, (p4 | 1u) << sh);
    c3 = lower_bound_u32(K, c2, e, (p4 | 3u) << sh);
}

__global__ void __launch_bounds__(OCT_THREADS)
k_octree(const OrbxLevel *__restrict__ lv, const uint32_t *__restrict__ cand, uint32_t *__restrict__ skey,
         uint32_t *__restrict__ scand, size_t cand_frame, const int *__restrict__ ncand, const uint32_t *__restrict__ lut,
         uint32_t *__restrict__ lvl_kp, int kp_frame, int *__restrict__ lvl_cnt, int *__restrict__ status, int NC) {
    extern __shared__ __align__(16) int sm[];
    __shared__ OctShared sh;
    const int level = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x;
    const OrbxLevel L = lv[level];
    const int n = min(ncand[frame * ORBX_MAX_LEVELS + level], L.cand_cap);
    if (n == 0) {   // ORBextractor.cc:574-585 erases every initial node -> empty result
        if (tid == 0) lvl_cnt[frame * ORBX_MAX_LEVELS + level] = 0;
        return;
    }
    const uint32_t *C = cand + (size_t)frame * cand_frame + L.cand_off;
    uint32_t *K = skey + (size_t)frame * cand_frame + L.cand_off;
    uint32_t *S = scand + (size_t)frame * cand_frame + L.cand_off;
    const uint32_t *lx = lut + L.lutx_off, *ly = lut + L.luty_off;
    const int D = L.depth, N = L.quota;

    // ---- 1. sort candidates by quadtree key ---------------------------------------------------------
    {
        int *offs = sm, *cur = sm + ORBX_OCT_CELLS + 1;
        const int ncells = L.ncells, cshift = L.cshift;
        for (int c = tid; c <= ncells; c += OCT_THREADS) offs[c] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += OCT_THREADS) {
            const uint32_t w = C[i];
            const uint32_t key = __ldg(lx + (w & 0xfff)) | __ldg(ly + ((w >> 12) & 0xfff));
            atomicAdd(&offs[key >> cshift], 1);
        }
        block_excl_scan(offs, ncells + 1, sh.warp_tmp);
        for (int c = tid; c < ncells; c += OCT_THREADS) cur[c] = offs[c];
        __syncthreads();
        for (int i = tid; i < n; i += OCT_THREADS) {
            const uint32_t w = C[i];
            const uint32_t key = __ldg(lx + (w & 0xfff)) | __ldg(ly + ((w >> 12) & 0xfff));
            const int pos = atomicAdd(&cur[key >> cshift], 1);
            K[pos] = key;
            S[pos] = w;
        }
        __syncthreads();
        for (int c = tid; c < ncells; c += OCT_THREADS) {
            const int b = offs[c], e = offs[c + 1];
            for (int i = b + 1; i < e; i++) {
                const uint32_t k = K[i], w = S[i];
                int j = i - 1;
                while (j >= b && K[j] > k) {
                    K[j + 1] = K[j];
                    S[j + 1] = S[j];
                    j--;
                }
                K[j + 1] = k;
                S[j + 1] = w;
            }
        }
        __syncthreads();
    }

    // ---- 2. replay the node list in rounds --------------------------------------------------------------
    int *fin_s = sm, *fin_e = sm + NC, *fin_st = sm + 2 * NC;
    int *act_s = sm + 3 * NC, *act_e = sm + 4 * NC, *act_i = sm + 5 * NC;
    int *nxt_s = sm + 6 * NC, *nxt_e = sm + 7 * NC, *nxt_i = sm + 8 * NC;
    int *cut1 = sm + 9 * NC, *cut2 = sm + 10 * NC, *cut3 = sm + 11 * NC;
    int *nm = sm + 12 * NC, *ns1 = sm + 13 * NC, *gr = sm + 14 * NC, *perm = sm + 15 * NC;

    if (tid == 0) {
        // initial nodes (ORBextractor.cc:551-585); list order = ascending bucket = descending stamp index
        int nfin = 0, nact = 0;
        for (int b = 0; b < L.n_ini; b++) {
            const int s = lower_bound_u32(K, 0, n, (uint32_t)b << (2 * D));
            const int e = b == L.n_ini - 1 ? n : lower_bound_u32(K, 0, n, (uint32_t)(b + 1) << (2 * D));
            if (e - s == 1) {
                fin_s[nfin] = s; fin_e[nfin] = e; fin_st[nfin] = L.n_ini - 1 - b; nfin++;
            } else if (e - s > 1) {
                act_s[nact] = s; act_e[nact] = e; act_i[nact] = L.n_ini - 1 - b; nact++;
            }
        }
        sh.nfin = nfin; sh.nact = nact; sh.size = nfin + nact; sh.phase = 0; sh.finish = 0;
    }
    __syncthreads();

    int rnd = 0;
    while (true) {
        rnd++;
        const int A = sh.nact, nfin = sh.nfin, prev = sh.size, phase = sh.phase;
        const int shq = 2 * (D - rnd);   // key shift of the children's quadrant bits
        if (A > 0 && shq < 0) {          // cannot happen: distinct pixels always separate within D splits
            if (tid == 0) atomicOr(&status[frame], ORBX_ST_OCT_DEPTH);
            break;
        }
        if (phase == 1) {
            // ORBextractor.cc:684: sort by (size, pointer) ascending, walk from the back
            for (int a = tid; a < A; a += OCT_THREADS) {
                const int sz = act_e[a] - act_s[a];
                int r = 0;
                for (int b = 0; b < A; b++) {
                    const int sb = act_e[b] - act_s[b];
                    r += (sb > sz) || (sb == sz && b < a);
                }
                perm[r] = a;
            }
            if (tid == 0) sh.cnt_lt = 0;
            __syncthreads();
        }
        for (int rho = tid; rho < A; rho += OCT_THREADS) {
            const int a = phase == 1 ? perm[rho] : rho;
            const int s = act_s[a], e = act_e[a];
            int c1, c2, c3;
            child_cuts(K, s, e, shq, c1, c2, c3);
            cut1[rho] = c1; cut2[rho] = c2; cut3[rho] = c3;
            const int n0 = c1 - s, n1 = c2 - c1, n2 = c3 - c2, n3 = e - c3;
            nm[rho] = (n0 > 1) + (n1 > 1) + (n2 > 1) + (n3 > 1);
            ns1[rho] = (n0 == 1) + (n1 == 1) + (n2 == 1) + (n3 == 1);
            gr[rho] = (n0 > 0) + (n1 > 0) + (n2 > 0) + (n3 > 0) - 1;
        }
        int kproc = A;
        if (phase == 1) {
            // how many nodes get divided before the list reaches N (ORBextractor.cc:740-741)
            block_excl_scan(gr, A, sh.warp_tmp);
            int lt = 0;
            for (int rho = tid; rho < A; rho += OCT_THREADS)   // gr is exclusive; own growth = non-empty children - 1
                lt += (prev + gr[rho] + nm[rho] + ns1[rho] - 1) < N;
            if (lt) atomicAdd(&sh.cnt_lt, lt);
            __syncthreads();
            kproc = min(A, sh.cnt_lt + 1);
            for (int rho = kproc + tid; rho < A; rho += OCT_THREADS) { nm[rho] = 0; ns1[rho] = 0; }
        }
        const int totM = block_excl_scan(nm, A, sh.warp_tmp);
        const int totS = block_excl_scan(ns1, A, sh.warp_tmp);
        const bool overflow = nfin + totS + (A - kproc) > NC || totM > NC;
        if (overflow) {
            if (tid == 0) atomicOr(&status[frame], ORBX_ST_NODE_OVERFLOW);
            break;
        }
        for (int rho = tid; rho < A; rho += OCT_THREADS) {
            const int a = phase == 1 ? perm[rho] : rho;
            if (rho < kproc) {
                const int b[5] = {act_s[a], cut1[rho], cut2[rho], cut3[rho], act_e[a]};
                int jf = nfin + ns1[rho], jm = nm[rho];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int c = b[q + 1] - b[q];
                    if (c == 1) {
                        fin_s[jf] = b[q]; fin_e[jf] = b[q + 1]; fin_st[jf] = (rnd << 16) | (4 * rho + q); jf++;
                    } else if (c > 1) {
                        const int pos = totM - 1 - jm;   // next pass walks the list front to back = newest first
                        nxt_s[pos] = b[q]; nxt_e[pos] = b[q + 1]; nxt_i[pos] = 4 * rho + q; jm++;
                    }
                }
            } else {   // never divided: stays in the list as it is
                const int j = nfin + totS + (rho - kproc);
                fin_s[j] = act_s[a]; fin_e[j] = act_e[a]; fin_st[j] = ((rnd - 1) << 16) | act_i[a];
            }
        }
        __syncthreads();
        { int *t; t = act_s; act_s = nxt_s; nxt_s = t; t = act_e; act_e = nxt_e; nxt_e = t; t = act_i; act_i = nxt_i; nxt_i = t; }
        if (tid == 0) {
            const int nf = nfin + totS + (A - kproc), size = nf + totM;
            sh.nfin = nf; sh.nact = totM; sh.size = size;
            if (size >= N || size == prev) sh.finish = 1;                 // ORBextractor.cc:661, :744
            else if (phase == 0 && size + 3 * totM > N) sh.phase = 1;      // ORBextractor.cc:665
        }
        __syncthreads();
        if (sh.finish) break;
    }
    __syncthreads();

    // ---- 3. best response per node, in list order (ORBextractor.cc:749-768) ----------------------------
    {
        const int A = sh.nact, nfin = sh.nfin;
        const int F = min(nfin + A, NC);
        for (int a = tid; a < A && nfin + a < NC; a += OCT_THREADS) {
            fin_s[nfin + a] = act_s[a]; fin_e[nfin + a] = act_e[a]; fin_st[nfin + a] = (rnd << 16) | act_i[a];
        }
        __syncthreads();
        uint32_t *dst = lvl_kp + (size_t)frame * kp_frame + L.kp_off;
        for (int j = tid; j < F; j += OCT_THREADS) {
            const int st = fin_st[j];
            int rank = 0;
            for (int m = 0; m < F; m++) rank += fin_st[m] > st;
            uint32_t best = 0;
            uint64_t best_ord = 0;
            int best_sc = -1;
            for (int k = fin_s[j]; k < fin_e[j]; k++) {
                const uint32_t w = S[k];
                const int sc = (int)(w >> 24);
                const int x = w & 0xfff, y = (w >> 12) & 0xfff;
                // position of the candidate in the reference's vToDistributeKeys: cell row, cell column, raster
                const uint64_t ord = ((uint64_t)(((y - 3) / L.hcell) * L.ncols + (x - 3) / L.wcell) << 24) | (w & 0xffffffu);
                if (sc > best_sc || (sc == best_sc && ord < best_ord)) { best = w; best_sc = sc; best_ord = ord; }
            }
            if (rank < L.kp_cap) dst[rank] = best;
        }
        if (tid == 0) lvl_cnt[frame * ORBX_MAX_LEVELS + level] = min(F, L.kp_cap);
    }
}

int orbx_octree_smem_bytes(int node_cap) {
    const int a = (2 * ORBX_OCT_CELLS + 1) * (int)sizeof(int), b = OCT_ARRAYS * node_cap * (int)sizeof(int);
    return a > b ? a : b;
}

orbx_status orbx_launch_octree(orbx_extractor *e, int batch, cudaStream_t s) {
    dim3 grid(e->nlevels, batch);
    k_octree<<<grid, OCT_THREADS, e->oct_smem, s>>>(e->d_lv, e->d_cand, e->d_skey, e->d_scand, e->cand_frame_cap, e->d_ncand,
                                                    e->d_lut, e->d_lvl_kp, e->capacity, e->d_lvl_cnt, e->d_status, e->node_cap);
    e->last_launches++;
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

orbx_status orbx_octree_init(int smem_bytes) {
    ORBX_CUDA(ORBX_RAISE_SMEM(k_octree));
    return ORBX_OK;
}
