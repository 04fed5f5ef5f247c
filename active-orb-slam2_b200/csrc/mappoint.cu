// MapPoint::ComputeDistinctiveDescriptors for many map points at once (reference src/MapPoint.cc:275-340; LocalMapping
// calls it for every new or fused map point, LocalMapping.cc:195, :441, :651-667).
//
// One warp per map point, lanes over the rows i of its N x N Hamming matrix.  The reference sorts every row and takes
// vDists[0.5*(N-1)]; distances are integers in [0, 256], so the rank-k value of a row is found without storing the row:
// a 9-step bisection over the value, counting the row's distances <= mid (the other descriptors are re-read through
// L1 / L2, broadcast to the lanes).  The first row with the strictly smallest median wins (warp minimum of
// median << 16 | i), like `if(median<BestMedian)` with BestMedian starting at INT_MAX.
#include "orbx_internal.cuh"

#define MP_THREADS 128

struct orbx_mappoints {
    int device, max_points, max_desc;
    int32_t *d_start, *d_best, *d_median;
    uint8_t *d_desc;
    cudaStream_t stream;
    int last_launches;
};

__global__ void __launch_bounds__(MP_THREADS)
k_distinctive(int n_points, const int32_t *__restrict__ start, const uint8_t *__restrict__ desc, int32_t *__restrict__ best_idx,
              int32_t *__restrict__ best_median) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (MP_THREADS / 32) + (threadIdx.x >> 5);
    if (p >= n_points) return;
    const int s = start[p], N = start[p + 1] - s;
    if (N <= 0) {
        if (lane == 0) { best_idx[p] = -1; if (best_median) best_median[p] = -1; }
        return;
    }
    const uint8_t *D = desc + (size_t)32 * s;
    const int k = (N - 1) >> 1;                               // (int)(0.5*(N-1))
    unsigned best = 0xffffffffu;
    for (int i = lane; i < N; i += 32) {
        const uint4 a0 = __ldg(reinterpret_cast<const uint4 *>(D + (size_t)32 * i));
        const uint4 a1 = __ldg(reinterpret_cast<const uint4 *>(D + (size_t)32 * i + 16));
        int lo = 0, hi = 256;
        while (lo < hi) {                                      // smallest v with #{j : d_ij <= v} >= k + 1
            const int mid = (lo + hi) >> 1;
            int cnt = 0;
            for (int j = 0; j < N; j++) {
                const uint4 b0 = __ldg(reinterpret_cast<const uint4 *>(D + (size_t)32 * j));
                const uint4 b1 = __ldg(reinterpret_cast<const uint4 *>(D + (size_t)32 * j + 16));
                const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                              __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                cnt += d <= mid;
            }
            if (cnt >= k + 1) hi = mid; else lo = mid + 1;
        }
        best = min(best, ((unsigned)lo << 16) | (unsigned)i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) {
        best_idx[p] = (int)(best & 0xffff);
        if (best_median) best_median[p] = (int)(best >> 16);
    }
}

extern "C" void orbx_mappoints_destroy(orbx_mappoints *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_start); cudaFree(h->d_best); cudaFree(h->d_median); cudaFree(h->d_desc);
    if (h->stream) cudaStreamDestroy(h->stream);
    free(h);
}

extern "C" orbx_status orbx_mappoints_create(orbx_mappoints **out, int max_points, int max_descriptors, int device) {
    if (!out) return ORBX_ERR_INVALID;
    *out = nullptr;
    if (max_points < 1 || max_descriptors < 1) return ORBX_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        orbx_set_error("no CUDA device %d (%d visible)", device, ndev);
        return ORBX_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    ORBX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        orbx_set_error("device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
        return ORBX_ERR_NO_DEVICE;
    }
    ORBX_CUDA(cudaSetDevice(device));
    orbx_mappoints *h = (orbx_mappoints *)calloc(1, sizeof(orbx_mappoints));
    if (!h) return ORBX_ERR_NOMEM;
    h->device = device; h->max_points = max_points; h->max_desc = max_descriptors;
    cudaError_t ce = cudaSuccess;
#define TRY(x) if (ce == cudaSuccess) ce = (x)
    TRY(cudaMalloc((void **)&h->d_start, sizeof(int32_t) * ((size_t)max_points + 1)));
    TRY(cudaMalloc((void **)&h->d_best, sizeof(int32_t) * (size_t)max_points));
    TRY(cudaMalloc((void **)&h->d_median, sizeof(int32_t) * (size_t)max_points));
    TRY(cudaMalloc((void **)&h->d_desc, (size_t)32 * max_descriptors));
    TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
#undef TRY
    if (ce != cudaSuccess) {
        orbx_set_error("orbx_mappoints_create: %s", cudaGetErrorString(ce));
        orbx_mappoints_destroy(h);
        return ORBX_ERR_CUDA;
    }
    *out = h;
    return ORBX_OK;
}

extern "C" orbx_status orbx_mappoints_distinctive_host(orbx_mappoints *h, int n_points, const int32_t *start, const uint8_t *desc,
                                                       int32_t *best_idx, int32_t *best_median) {
    if (!h || n_points < 0 || (n_points && (!start || !best_idx))) return ORBX_ERR_INVALID;
    h->last_launches = 0;
    if (n_points == 0) return ORBX_OK;
    const int total = start[n_points];
    if (start[0] != 0 || total < 0 || (total && !desc)) return ORBX_ERR_INVALID;
    for (int p = 0; p < n_points; p++)
        if (start[p + 1] < start[p] || start[p + 1] - start[p] > 65535) return ORBX_ERR_INVALID;
    if (n_points > h->max_points || total > h->max_desc) {
        orbx_set_error("orbx_mappoints: %d points / %d descriptors, handle was created for %d / %d", n_points, total, h->max_points, h->max_desc);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    ORBX_CUDA(cudaMemcpyAsync(h->d_start, start, sizeof(int32_t) * ((size_t)n_points + 1), cudaMemcpyHostToDevice, s));
    if (total) ORBX_CUDA(cudaMemcpyAsync(h->d_desc, desc, (size_t)32 * total, cudaMemcpyHostToDevice, s));
    const int per_cta = MP_THREADS / 32;
    k_distinctive<<<(n_points + per_cta - 1) / per_cta, MP_THREADS, 0, s>>>(n_points, h->d_start, h->d_desc, h->d_best, h->d_median);
    ORBX_CUDA(cudaGetLastError());
    h->last_launches = 1;
    ORBX_CUDA(cudaMemcpyAsync(best_idx, h->d_best, sizeof(int32_t) * n_points, cudaMemcpyDeviceToHost, s));
    if (best_median) ORBX_CUDA(cudaMemcpyAsync(best_median, h->d_median, sizeof(int32_t) * n_points, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" int orbx_mappoints_last_launches(const orbx_mappoints *h) { return h ? h->last_launches : 0; }
