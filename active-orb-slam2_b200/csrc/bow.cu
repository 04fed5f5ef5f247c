// Bag-of-words transform: the per-feature tree descent of DBoW2's TemplatedVocabulary::transform (reference
// Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1216-1262, called for every descriptor of a frame by Frame::ComputeBoW /
// KeyFrame::ComputeBoW through :1138-1200) with FORB::distance = 256-bit Hamming distance.
//
// The vocabulary (ORBvoc: k = 10, L = 6, 1,082,073 nodes, 44 MB) stays resident in HBM and fits the 126 MB L2.  Children
// of a node are stored in consecutive SLOTS in the reference's child order, 32-byte descriptors back to back, so one level
// of the descent is one coalesced read of <= k descriptors.  A warp carries three features at once (lanes 0-9, 10-19,
// 20-29 take the children of the first, second, third feature; wider trees loop), key = distance << 8 | child position
// so that the minimum is the FIRST child among equals (`d < best_d`, :1245).
#include "orbx_internal.cuh"

#define BOW_THREADS 256
#define BOW_GROUP 10            // lanes per feature
#define BOW_PER_WARP 3

struct orbx_vocabulary {
    int device, n_nodes, L, max_children;
    int32_t *d_slot_start;      // [n_nodes + 1] first child slot of a node
    int32_t *d_slot_node;       // [n_slots] node id of a slot
    uint8_t *d_slot_desc;       // [n_slots][32]
    double *d_weight;           // [n_nodes]
    int32_t *d_word;            // [n_nodes]
    // staging of the _host entry point
    uint8_t *d_desc; int32_t *d_out_word, *d_out_node; double *d_out_weight; int stage_cap;
    cudaStream_t stream;
    int last_launches;
};

__global__ void __launch_bounds__(BOW_THREADS)
k_bow_transform(const int32_t *__restrict__ slot_start, const int32_t *__restrict__ slot_node, const uint8_t *__restrict__ slot_desc,
                const double *__restrict__ node_weight, const int32_t *__restrict__ node_word, int nid_level,
                const uint8_t *__restrict__ desc, const int32_t *__restrict__ counts, int count_step, int n_fixed, int pitch,
                int32_t *__restrict__ word, int32_t *__restrict__ node, double *__restrict__ weight) {
    const int frame = blockIdx.y;
    const int n = n_fixed >= 0 ? n_fixed : counts[(size_t)frame * count_step];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / BOW_GROUP, sub = lane - grp * BOW_GROUP;          // lanes 30, 31: grp == 3, idle
    const int i = (blockIdx.x * (BOW_THREADS / 32) + warp) * BOW_PER_WARP + grp;
    const bool live = grp < BOW_PER_WARP && i < n;
    const unsigned gmask = grp < BOW_PER_WARP ? (0x3ffu << (grp * BOW_GROUP)) : 0xc0000000u;
    const uint8_t *f = desc + ((size_t)frame * pitch + (live ? i : 0)) * 32;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (live) { a0 = __ldg(reinterpret_cast<const uint4 *>(f)); a1 = __ldg(reinterpret_cast<const uint4 *>(f + 16)); }
    int cur = 0, nid = 0, level = 0;
    bool done = !live || slot_start[1] == slot_start[0];
    const bool empty_voc = slot_start[1] == slot_start[0];
    // every lane of the warp runs the loop until all three features reached a leaf (shuffles need the whole group)
    while (__any_sync(0xffffffffu, !done)) {
        unsigned best = 0xffffffffu;
        if (!done) {
            const int c0 = slot_start[cur], c1 = slot_start[cur + 1];
            for (int c = c0 + sub; c < c1; c += BOW_GROUP) {
                const uint4 b0 = __ldg(reinterpret_cast<const uint4 *>(slot_desc + (size_t)c * 32));
                const uint4 b1 = __ldg(reinterpret_cast<const uint4 *>(slot_desc + (size_t)c * 32 + 16));
                const unsigned d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                                   __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
                best = min(best, (d << 20) | (unsigned)(c - c0));             // up to 2^20 children per node
            }
        }
        // minimum over the group's lanes (xor butterflies stay inside 0..9 only for aligned powers of two, so reduce by shuffling
        // from explicit lanes of the group)
#pragma unroll
        for (int o = 1; o < BOW_GROUP; o++) {
            const int src = grp * BOW_GROUP + (sub + o) % BOW_GROUP;
            const unsigned v = __shfl_sync(0xffffffffu, best, grp < BOW_PER_WARP ? src : lane);
            best = min(best, v);      // after all rotations every lane of the group has seen every other lane's value
        }
        if (!done) {
            const int c0 = slot_start[cur];
            cur = slot_node[c0 + (int)(best & 0xfffffu)];
            level++;
            if (level == nid_level) nid = cur;
            done = slot_start[cur + 1] == slot_start[cur];                    // isLeaf()
        }
    }
    (void)gmask;
    if (live && sub == 0) {
        const size_t o = (size_t)frame * pitch + i;
        if (empty_voc) { word[o] = -1; node[o] = 0; weight[o] = 0.0; }
        else { word[o] = node_word[cur]; node[o] = nid; weight[o] = node_weight[cur]; }
    }
}

extern "C" void orbx_vocabulary_destroy(orbx_vocabulary *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_slot_start); cudaFree(h->d_slot_node); cudaFree(h->d_slot_desc); cudaFree(h->d_weight); cudaFree(h->d_word);
    cudaFree(h->d_desc); cudaFree(h->d_out_word); cudaFree(h->d_out_node); cudaFree(h->d_out_weight);
    if (h->stream) cudaStreamDestroy(h->stream);
    free(h);
}

extern "C" orbx_status orbx_vocabulary_create(orbx_vocabulary **out, int n_nodes, const int32_t *child_start, const int32_t *children,
                                              const uint8_t *node_desc, const double *node_weight, const int32_t *node_word_id, int L,
                                              int max_features, int device) {
    if (!out) return ORBX_ERR_INVALID;
    *out = nullptr;
    if (n_nodes < 1 || !child_start || !node_desc || !node_weight || !node_word_id || L < 0 || max_features < 1) {
        orbx_set_error("orbx_vocabulary_create: bad argument");
        return ORBX_ERR_INVALID;
    }
    const int n_slots = child_start[n_nodes];
    if (child_start[0] != 0 || n_slots < 0 || (n_slots && !children)) return ORBX_ERR_INVALID;
    for (int i = 0; i < n_nodes; i++)
        if (child_start[i + 1] < child_start[i]) return ORBX_ERR_INVALID;
    for (int s = 0; s < n_slots; s++)
        if (children[s] <= 0 || children[s] >= n_nodes) {
            orbx_set_error("orbx_vocabulary_create: child id %d out of range", children[s]);
            return ORBX_ERR_INVALID;
        }
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
    orbx_vocabulary *h = (orbx_vocabulary *)calloc(1, sizeof(orbx_vocabulary));
    if (!h) return ORBX_ERR_NOMEM;
    h->device = device; h->n_nodes = n_nodes; h->L = L; h->stage_cap = max_features;
    // slot order: descriptors of a node's children back to back
    uint8_t *slot_desc = (uint8_t *)malloc((size_t)32 * (n_slots ? n_slots : 1));
    if (!slot_desc) { free(h); return ORBX_ERR_NOMEM; }
    for (int s = 0; s < n_slots; s++) memcpy(slot_desc + (size_t)32 * s, node_desc + (size_t)32 * children[s], 32);
    cudaError_t ce = cudaSuccess;
#define TRY(x) if (ce == cudaSuccess) ce = (x)
    const size_t ns = n_slots ? n_slots : 1, mf = (size_t)max_features;
    TRY(cudaMalloc((void **)&h->d_slot_start, sizeof(int32_t) * ((size_t)n_nodes + 1)));
    TRY(cudaMalloc((void **)&h->d_slot_node, sizeof(int32_t) * ns));
    TRY(cudaMalloc((void **)&h->d_slot_desc, 32 * ns));
    TRY(cudaMalloc((void **)&h->d_weight, sizeof(double) * n_nodes));
    TRY(cudaMalloc((void **)&h->d_word, sizeof(int32_t) * n_nodes));
    TRY(cudaMalloc((void **)&h->d_desc, 32 * mf));
    TRY(cudaMalloc((void **)&h->d_out_word, sizeof(int32_t) * mf));
    TRY(cudaMalloc((void **)&h->d_out_node, sizeof(int32_t) * mf));
    TRY(cudaMalloc((void **)&h->d_out_weight, sizeof(double) * mf));
    TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    TRY(cudaMemcpy(h->d_slot_start, child_start, sizeof(int32_t) * ((size_t)n_nodes + 1), cudaMemcpyHostToDevice));
    if (n_slots) {
        TRY(cudaMemcpy(h->d_slot_node, children, sizeof(int32_t) * n_slots, cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(h->d_slot_desc, slot_desc, (size_t)32 * n_slots, cudaMemcpyHostToDevice));
    }
    TRY(cudaMemcpy(h->d_weight, node_weight, sizeof(double) * n_nodes, cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(h->d_word, node_word_id, sizeof(int32_t) * n_nodes, cudaMemcpyHostToDevice));
#undef TRY
    free(slot_desc);
    if (ce != cudaSuccess) {
        orbx_set_error("orbx_vocabulary_create: %s", cudaGetErrorString(ce));
        orbx_vocabulary_destroy(h);
        return ORBX_ERR_CUDA;
    }
    *out = h;
    return ORBX_OK;
}

static void bow_launch(orbx_vocabulary *h, int levelsup, const uint8_t *d_desc, const int32_t *d_counts, int count_step, int n_fixed,
                       int pitch, int max_count, int batch, int32_t *d_word, int32_t *d_node, double *d_weight, cudaStream_t s) {
    const int per_cta = (BOW_THREADS / 32) * BOW_PER_WARP;
    dim3 grid((max_count + per_cta - 1) / per_cta, batch);
    k_bow_transform<<<grid, BOW_THREADS, 0, s>>>(h->d_slot_start, h->d_slot_node, h->d_slot_desc, h->d_weight, h->d_word, h->L - levelsup,
                                                 d_desc, d_counts, count_step, n_fixed, pitch, d_word, d_node, d_weight);
    h->last_launches = 1;
}

extern "C" orbx_status orbx_vocabulary_transform_device(orbx_vocabulary *h, int levelsup, const uint8_t *d_desc, const int32_t *d_counts,
                                                        int count_step, int pitch, int max_count, int batch, int32_t *d_word,
                                                        int32_t *d_node, double *d_weight, void *stream) {
    if (!h || batch < 0 || pitch < 1 || max_count < 0 || !d_desc || !d_counts || !d_word || !d_node || !d_weight) return ORBX_ERR_INVALID;
    h->last_launches = 0;
    if (batch == 0 || max_count == 0) return ORBX_OK;
    ORBX_CUDA(cudaSetDevice(h->device));
    bow_launch(h, levelsup, d_desc, d_counts, count_step, -1, pitch, max_count, batch, d_word, d_node, d_weight, (cudaStream_t)stream);
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

extern "C" orbx_status orbx_vocabulary_transform_host(orbx_vocabulary *h, const uint8_t *desc, int n, int levelsup, int32_t *word,
                                                      int32_t *node, double *weight) {
    if (!h || n < 0 || (n && (!desc || !word || !node || !weight))) return ORBX_ERR_INVALID;
    h->last_launches = 0;
    if (n == 0) return ORBX_OK;
    if (n > h->stage_cap) {
        orbx_set_error("orbx_vocabulary: %d features, handle was created for %d", n, h->stage_cap);
        return ORBX_ERR_CAPACITY;
    }
    ORBX_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    ORBX_CUDA(cudaMemcpyAsync(h->d_desc, desc, (size_t)32 * n, cudaMemcpyHostToDevice, s));
    bow_launch(h, levelsup, h->d_desc, nullptr, 0, n, n, n, 1, h->d_out_word, h->d_out_node, h->d_out_weight, s);
    ORBX_CUDA(cudaGetLastError());
    ORBX_CUDA(cudaMemcpyAsync(word, h->d_out_word, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(node, h->d_out_node, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaMemcpyAsync(weight, h->d_out_weight, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    ORBX_CUDA(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" int orbx_vocabulary_last_launches(const orbx_vocabulary *h) { return h ? h->last_launches : 0; }
