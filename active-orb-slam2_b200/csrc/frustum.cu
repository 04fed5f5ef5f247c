// Frame::isInFrustum for all map points of the local map at once (reference src/Frame.cc:298-354; Tracking::SearchLocalPoints
// calls it per local map point, Tracking.cc:1085-1103) -- it produces exactly the orbx_track_point records that
// SearchByProjection(Frame&, vector<MapPoint*>&, th) consumes.
//
// One thread per map point, every float step in the reference's order with explicit round-to-nearest operations:
//   Pc = mRcw*P + mtcw          cv::Mat CV_32F product: ((r0*x0 + r1*x1) + r2*x2) + t, no FMA
//   u, v, u - mbf*invz          float
//   dist = cv::norm(PO)         sqrt of the double-accumulated squares, narrowed to float
//   viewCos = PO.dot(Pn)/dist   cv::Mat::dot on 3 floats accumulates products in double; double / float, narrowed to float
//   PredictScale                ceil(logf(mfMaxDistance/dist) / mfLogScaleFactor) clamped to [0, nlevels)
// The one step that cannot be made bit-identical by construction is logf (libm's implementation on the host): the level is
// computed on the device unless the quotient lies within 1e-4 of an integer, where a last-bit difference of logf could change
// the ceiling; those (about one point in 10^4 .. 10^5) are flagged in pad[0] and the adapter evaluates PredictScale for them on
// the host, with the reference's own function.
#include "orbx_internal.cuh"

__global__ void __launch_bounds__(256)
k_frustum(const orbx_frustum_frame F, int n, const orbx_frustum_point *__restrict__ pts, orbx_track_point *__restrict__ out,
          int32_t *__restrict__ n_ambiguous) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const orbx_frustum_point p = pts[i];
    orbx_track_point t;
    t.proj_x = t.proj_y = t.proj_xr = t.view_cos = 0.f;
    t.level = 0; t.in_view = 0; t.blocks = p.blocks; t.pad[0] = t.pad[1] = 0;
    bool ok = !p.skip;                                     // isBad() / already matched: the reference does not call isInFrustum
    float PcX = 0, PcY = 0, PcZ = 0;
    if (ok) {
        const float *R = F.Rcw;
        PcX = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], p.x), __fmul_rn(R[1], p.y)), __fmul_rn(R[2], p.z)), F.tcw[0]);
        PcY = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], p.x), __fmul_rn(R[4], p.y)), __fmul_rn(R[5], p.z)), F.tcw[1]);
        PcZ = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], p.x), __fmul_rn(R[7], p.y)), __fmul_rn(R[8], p.z)), F.tcw[2]);
        ok = !(PcZ < 0.0f);                                // :312
    }
    if (ok) {
        const float invz = __fdiv_rn(1.0f, PcZ);
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(F.fx, PcX), invz), F.cx);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(F.fy, PcY), invz), F.cy);
        ok = !(u < F.min_x || u > F.max_x) && !(v < F.min_y || v > F.max_y);          // :320-323
        if (ok) {
            const float maxDistance = __fmul_rn(1.2f, p.max_distance), minDistance = __fmul_rn(0.8f, p.min_distance);   // MapPoint.cc:430-441
            const float ox = __fsub_rn(p.x, F.Ow[0]), oy = __fsub_rn(p.y, F.Ow[1]), oz = __fsub_rn(p.z, F.Ow[2]);
            const double s2 = __dadd_rn(__dadd_rn(__dmul_rn((double)ox, (double)ox), __dmul_rn((double)oy, (double)oy)), __dmul_rn((double)oz, (double)oz));
            const float dist = __double2float_rn(sqrt(s2));
            ok = !(dist < minDistance || dist > maxDistance);                          // :331-332
            if (ok) {
                const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)ox, (double)p.nx), __dmul_rn((double)oy, (double)p.ny)),
                                             __dmul_rn((double)oz, (double)p.nz));
                const float viewCos = __double2float_rn(__ddiv_rn(dot, (double)dist));
                ok = !(viewCos < F.viewing_cos_limit);                                  // :339-340
                if (ok) {
                    const float ratio = __fdiv_rn(p.max_distance, dist);                // MapPoint::PredictScale, MapPoint.cc:444-459
                    const float q = __fdiv_rn(logf(ratio), F.log_scale_factor);
                    int nScale = (int)ceilf(q);
                    if (fabsf(q - rintf(q)) < 1e-4f || !(q == q)) { t.pad[0] = 1; atomicAdd(n_ambiguous, 1); }
                    if (nScale < 0) nScale = 0;
                    else if (nScale >= F.n_levels) nScale = F.n_levels - 1;
                    t.in_view = 1;
                    t.proj_x = u;
                    t.proj_xr = __fsub_rn(u, __fmul_rn(F.bf, invz));
                    t.proj_y = v;
                    t.level = nScale;
                    t.view_cos = viewCos;
                }
            }
        }
    }
    out[i] = t;
}

extern "C" orbx_status orbx_frustum_device(const orbx_frustum_frame *F, int n, const orbx_frustum_point *d_pts, orbx_track_point *d_out,
                                           int32_t *d_n_ambiguous, void *stream) {
    if (!F || n < 0 || (n && (!d_pts || !d_out || !d_n_ambiguous)) || F->n_levels < 1) return ORBX_ERR_INVALID;
    if (n == 0) return ORBX_OK;
    ORBX_CUDA(cudaMemsetAsync(d_n_ambiguous, 0, sizeof(int32_t), (cudaStream_t)stream));
    k_frustum<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*F, n, d_pts, d_out, d_n_ambiguous);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

extern "C" orbx_status orbx_frustum_host(const orbx_frustum_frame *F, int n, const orbx_frustum_point *pts, orbx_track_point *out,
                                         int32_t *n_ambiguous, int device) {
    if (!F || n < 0 || (n && (!pts || !out)) || F->n_levels < 1) return ORBX_ERR_INVALID;
    if (n_ambiguous) *n_ambiguous = 0;
    if (n == 0) return ORBX_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        orbx_set_error("no CUDA device %d (%d visible)", device, ndev);
        return ORBX_ERR_NO_DEVICE;
    }
    ORBX_CUDA(cudaSetDevice(device));
    {   // keep what the stream-ordered allocator frees: without a release threshold the pool hands its memory back to the driver at every
        // synchronisation and each call pays a fresh physical allocation (7 ms per call measured in the config-4 replay)
        static bool pool_kept[64] = {false};
        if (device < 64 && !pool_kept[device]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_kept[device] = true;
        }
    }
    cudaStream_t s = cudaStreamPerThread;
    orbx_frustum_point *d_in = nullptr; orbx_track_point *d_out = nullptr; int32_t *d_amb = nullptr;
    ORBX_CUDA(cudaMallocAsync((void **)&d_in, sizeof(orbx_frustum_point) * n, s));
    ORBX_CUDA(cudaMallocAsync((void **)&d_out, sizeof(orbx_track_point) * n, s));
    ORBX_CUDA(cudaMallocAsync((void **)&d_amb, sizeof(int32_t), s));
    ORBX_CUDA(cudaMemcpyAsync(d_in, pts, sizeof(orbx_frustum_point) * n, cudaMemcpyHostToDevice, s));
    orbx_status st = orbx_frustum_device(F, n, d_in, d_out, d_amb, s);
    int32_t amb = 0;
    if (st == ORBX_OK) {
        ORBX_CUDA(cudaMemcpyAsync(out, d_out, sizeof(orbx_track_point) * n, cudaMemcpyDeviceToHost, s));
        ORBX_CUDA(cudaMemcpyAsync(&amb, d_amb, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    cudaFreeAsync(d_in, s); cudaFreeAsync(d_out, s); cudaFreeAsync(d_amb, s);
    ORBX_CUDA(cudaStreamSynchronize(s));
    if (n_ambiguous) *n_ambiguous = amb;
    return st;
}
