// Block-wide exclusive scan over an array in shared memory (used by the quadtree and matcher kernels).
#pragma once

// exclusive scan of a[0..n) in shared memory, in place; returns the total to every thread
__device__ inline int block_excl_scan(int *a, int n, int *warp_tmp) {
    __syncthreads();
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = T >> 5;
    const int per = (n + T - 1) / T;
    const int b = min(tid * per, n), e = min(b + per, n);
    int sum = 0;
    for (int i = b; i < e; i++) sum += a[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tmp[w] = incl;
    __syncthreads();
    if (w == 0) {
        const int v = lane < nw ? warp_tmp[lane] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        warp_tmp[lane] = inc - v;
        if (lane == 31) warp_tmp[32] = inc;
    }
    __syncthreads();
    int base = warp_tmp[w] + incl - sum;
    for (int i = b; i < e; i++) {
        const int t = a[i];
        a[i] = base;
        base += t;
    }
    const int total = warp_tmp[32];
    __syncthreads();
    return total;
}

