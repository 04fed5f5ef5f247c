/* orbx CPU oracle, map-point helpers — TEST INFRASTRUCTURE ONLY (see orbx_oracle.h).
 *
 * MapPoint::ComputeDistinctiveDescriptors (reference src/MapPoint.cc:275-340): among the descriptors of a map point's
 * observations, keep the one with the least median Hamming distance to the others.  Restated as the reference does it:
 * the full N x N distance matrix (float entries holding integers), every row sorted, median = vDists[0.5*(N-1)], the
 * first row with a strictly smaller median wins (BestMedian starts at INT_MAX).
 * PARITY PINNING: the reference holds no test.  PINNED against the reference's own MapPoint::ComputeDistinctiveDescriptors run
 * from source over observations held by its KeyFrames (oracle/_ref/liborbmatcher_ref.so, tests/test_oracle_ref_matcher.py): the
 * same descriptor is chosen.  tests/test_mappoint_oracle.py also checks it against numpy.
 */
#include "orbx_oracle.h"
#include <limits.h>
#include <stddef.h>
#include <stdlib.h>

static int cmp_int(const void *a, const void *b) { return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }

void orbo_distinctive_descriptors(int n_points, const int32_t *start, const uint8_t *desc, int32_t *best_idx, int32_t *best_median) {
    for (int p = 0; p < n_points; p++) {
        const int N = start[p + 1] - start[p];
        const uint8_t *D = desc + (size_t)32 * start[p];
        best_idx[p] = -1;
        if (best_median) best_median[p] = -1;
        if (N <= 0) continue;                                    /* `if(vDescriptors.empty()) return;` */
        float *M = (float *)malloc(sizeof(float) * (size_t)N * N);
        int *v = (int *)malloc(sizeof(int) * (size_t)N);
        for (int i = 0; i < N; i++) {
            M[(size_t)i * N + i] = 0;
            for (int j = i + 1; j < N; j++) {
                const int d = orbo_hamming256(D + (size_t)32 * i, D + (size_t)32 * j);
                M[(size_t)i * N + j] = (float)d; M[(size_t)j * N + i] = (float)d;
            }
        }
        int BestMedian = INT_MAX, BestIdx = 0;
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++) v[j] = (int)M[(size_t)i * N + j];
            qsort(v, (size_t)N, sizeof(int), cmp_int);
            const int median = v[(int)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best_idx[p] = BestIdx;
        if (best_median) best_median[p] = BestMedian;
        free(M); free(v);
    }
}
