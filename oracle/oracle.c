/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  A plain-C CPU restatement of the reference's
 * algorithms on the SEE-VCN object-completion + voxelization path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it;
 * the product (see-vcn_b200/) never does.
 *
 * Each function cites the reference lines it follows (paths relative to the
 * darrenjkt/SEE-VCN checkout).  Build: `make -C oracle` (gcc -O2 -ffp-contract=off -fopenmp).
 * -ffp-contract=off matters: every fused multiply-add below is an explicit fmaf() placed
 * where nvcc 12.9 contracts the reference's expression for sm_100a (checked in the SASS of
 * the unmodified reference .cu), everything else is separately rounded.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- points in boxes ------------------------------------------------------------------
 * ref: check_pt_in_box3d + lidar_to_local_coords,
 *      detector3d/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:16-36 (GPU, MARGIN 1e-5)
 *      detector3d/pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-140    (CPU, MARGIN 1e-2)
 * fused != 0 reproduces the GPU build's contraction:
 *      local_x = fma(sx, cosa, rn(sy * -sina)); local_y = fma(sy, cosa, rn(sx * sina))
 * Returns the in-box flag; *slack (optional) receives the smallest absolute distance of the
 * x / y comparisons from their decision boundary (1e30 when z already rejects), so tests can tell
 * decided points from ones that sit within an ulp of a face (where cosf/sinf of libm and CUDA
 * may round differently). */
static int pt_in_box(const float* pt, const float* box, float margin, int fused, double* slack) {
    const float x = pt[0], y = pt[1], z = pt[2];
    const float cx = box[0], cy = box[1], cz = box[2];
    const float dx = box[3], dy = box[4], dz = box[5], rz = box[6];
    const float az = fabsf(z - cz);
    if (slack) *slack = 1e30;   /* the z test involves no cos/sin: it is exact on both sides */
    if ((double)az > (double)dz / 2.0) return 0;
    const float cosa = cosf(-rz), sina = sinf(-rz);
    const float sx = x - cx, sy = y - cy;
    float lx, ly;
    if (fused) {
        lx = fmaf(sx, cosa, sy * (-sina));
        ly = fmaf(sy, cosa, sx * sina);
    } else {
        lx = sx * cosa + sy * (-sina);
        ly = sx * sina + sy * cosa;
    }
    const double hx = (double)dx / 2.0 + (double)margin, hy = (double)dy / 2.0 + (double)margin;
    if (slack) {
        const double s1 = fabs((double)fabsf(lx) - hx), s2 = fabs((double)fabsf(ly) - hy);
        if (s1 < *slack) *slack = s1;
        if (s2 < *slack) *slack = s2;
    }
    return ((double)fabsf(lx) < hx) & ((double)fabsf(ly) < hy);
}

/* ref: points_in_boxes_kernel, roiaware_pool3d_kernel.cu:313-336 — first (lowest) box index, else -1.
 * min_slack (B*P, optional): min over the boxes visited of the decision slack. */
void orc_points_in_boxes_gpu(int B, int T, int P, const float* boxes, const float* pts, int* out, float* min_slack) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)B * P; ++i) {
        const int b = (int)(i / P);
        int res = -1;
        double ms = 1e30;
        for (int k = 0; k < T; ++k) {
            double s;
            const int in = pt_in_box(pts + i * 3, boxes + ((long long)b * T + k) * 7, 1e-5f, 1, &s);
            if (s < ms) ms = s;
            if (in) { res = k; break; }
        }
        out[i] = res;
        if (min_slack) min_slack[i] = (float)ms;
    }
}

/* ref: points_in_boxes_cpu, roiaware_pool3d.cpp:143-168 — dense (T,P) 0/1 */
void orc_points_in_boxes_cpu(int T, int P, const float* boxes, const float* pts, int* out) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < T; ++k)
        for (int j = 0; j < P; ++j) out[(long long)k * P + j] = pt_in_box(pts + (long long)j * 3, boxes + k * 7, 1e-2f, 0, NULL);
}

/* ---- furthest point sampling -------------------------------------------------------------
 * ref: farthest_point_sampling_kernel, pointnet2_batch/src/sampling_gpu.cu:100-216;
 *      block size rule opt_n_threads, cuda_utils.h:10-14; temp pre-fill 1e10, pointnet2_utils.py:26.
 * Tie rule of the kernel: a thread keeps the first (lowest k) strict maximum of its stride; the
 * shared-memory tree (strides bs/2 ... 1, `v2 > v1 ? i2 : i1`) keeps the LOWER position on equal
 * values.  After the level with stride s position p holds the winner of the threads == p (mod s), so
 * the last level decides by thread-id bit 0, the one before by bit 1, ...: among equal maxima the
 * thread with the smallest BIT-REVERSED id wins. */
static unsigned brev32(unsigned v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
}
static int opt_n_threads(int n) {
    const int pow_2 = (int)(log((double)n) / log(2.0));
    int bs = 1 << pow_2;
    if (bs > 1024) bs = 1024;
    if (bs < 1) bs = 1;
    return bs;
}

void orc_fps(int B, int N, int M, const float* xyz, float* temp, int* idxs) {
    if (M <= 0) return;
    const int bs = opt_n_threads(N);
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        const float* d = xyz + (long long)b * N * 3;
        float* t = temp + (long long)b * N;
        int* out = idxs + (long long)b * M;
        for (int k = 0; k < N; ++k) t[k] = 1e10f;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < M; ++j) {
            const float x1 = d[old * 3], y1 = d[old * 3 + 1], z1 = d[old * 3 + 2];
            float best = -1.f; int besti = 0;
            for (int k = 0; k < N; ++k) {   /* distance update in any order */
                const float ddx = d[k * 3] - x1, ddy = d[k * 3 + 1] - y1, ddz = d[k * 3 + 2] - z1;
                const float dd = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
                t[k] = fminf(dd, t[k]);
            }
            unsigned bestkey = 0xffffffffu;
            for (int tid = 0; tid < bs; ++tid) {   /* per-thread strict max, then the tree's tie rule */
                float tb = -1.f; int ti = 0;
                for (int k = tid; k < N; k += bs) if (t[k] > tb) { tb = t[k]; ti = k; }
                const unsigned key = brev32((unsigned)tid);
                if (tb > best || (tb == best && key < bestkey)) { best = tb; besti = ti; bestkey = key; }
            }
            old = besti;
            out[j] = old;
        }
    }
}

/* ---- kNN -----------------------------------------------------------------------------------
 * ref: scipy cKDTree.query(k) as called at see/surface_completion/models/vcn/utils/sampling.py:30-34
 * (float32 inputs up-cast to float64, Euclidean, ascending); torch twin topk(k, largest=False) :59-61.
 * Brute force in float64; equal distances keep the lower index first. */
void orc_knn(int B, int R, int Q, int K, const float* ref, const float* query, float* dist, int* idx) {
#pragma omp parallel for schedule(dynamic, 16) collapse(2)
    for (int b = 0; b < B; ++b)
        for (int q = 0; q < Q; ++q) {
            double bd[64]; int bi[64];
            for (int j = 0; j < K; ++j) { bd[j] = INFINITY; bi[j] = -1; }
            const float* qp = query + ((long long)b * Q + q) * 3;
            for (int r = 0; r < R; ++r) {
                const float* rp = ref + ((long long)b * R + r) * 3;
                const double dx = (double)rp[0] - qp[0], dy = (double)rp[1] - qp[1], dz = (double)rp[2] - qp[2];
                double d = dx * dx + dy * dy + dz * dz; int ci = r;
                if (d < bd[K - 1]) {   /* stable insertion: after the first strictly larger entry, shift the rest (ties keep the lower index) */
                    int j = K - 1;
                    while (j > 0 && d < bd[j - 1]) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
                    bd[j] = d; bi[j] = ci;
                }
            }
            for (int j = 0; j < K; ++j) {
                idx[((long long)b * Q + q) * K + j] = bi[j];
                if (dist) dist[((long long)b * Q + q) * K + j] = (float)sqrt(bd[j]);
            }
        }
}

/* ref: partial_with_KDTree, sampling.py:8-41: union of the k-NN index sets of the (unique)
 * partial points, ascending, complete[S] tiled cyclically to surface_pts rows.  (Duplicate
 * partial points add nothing to a union, so np.unique at :31 is not restated.) */
void orc_knn_surface_select(int B, int NP, int R, int K, int SP, const float* partial, const float* complete, float* out,
                            int* sel_count) {
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        unsigned char* mark = (unsigned char*)calloc((size_t)R, 1);
        int* nn = (int*)malloc(sizeof(int) * (size_t)K);
        for (int q = 0; q < NP; ++q) {
            orc_knn(1, R, 1, K, complete + (long long)b * R * 3, partial + ((long long)b * NP + q) * 3, NULL, nn);
            for (int j = 0; j < K; ++j) if (nn[j] >= 0) mark[nn[j]] = 1;
        }
        int* sel = (int*)malloc(sizeof(int) * (size_t)(R > 0 ? R : 1));
        int cnt = 0;
        for (int r = 0; r < R; ++r) if (mark[r]) sel[cnt++] = r;
        sel_count[b] = cnt;
        for (int j = 0; j < SP; ++j) {
            float* o = out + ((long long)b * SP + j) * 3;
            if (cnt > 0) memcpy(o, complete + ((long long)b * R + sel[j % cnt]) * 3, 12);
            else o[0] = o[1] = o[2] = 0.f;
        }
        free(sel); free(nn); free(mark);
    }
}

/* ref: get_largest_cluster, see/surface_completion/models/vcn/utils/sampling.py:83-109 -> open3d
 * cluster_dbscan (NOT vendored; setup.py:25 pins 0.14.1) — PARITY UNPINNED.  Restated for min_points <= 2,
 * where DBSCAN = connected components of the graph {d^2 < eps^2} on float64 copies of the points
 * (isolated points are noise for min_points = 2); clusters are labelled in order of their first point,
 * np.bincount/argmax picks the largest (first on ties), members keep row order and are tiled to total_pts. */
static int uf_find(int* p, int i) { while (p[i] != i) { p[i] = p[p[i]]; i = p[i]; } return i; }
void orc_largest_cluster(int B, int N, int TP, double eps, int min_points, const float* pts, float* out, int* out_count) {
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        const float* p = pts + (long long)b * N * 3;
        int* par = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
        int* deg = (int*)calloc((size_t)(N > 0 ? N : 1), sizeof(int));
        int* sz = (int*)calloc((size_t)(N > 0 ? N : 1), sizeof(int));
        for (int i = 0; i < N; ++i) par[i] = i;
        for (int i = 0; i < N; ++i)
            for (int j = i; j < N; ++j) {
                const double dx = (double)p[i * 3] - (double)p[j * 3], dy = (double)p[i * 3 + 1] - (double)p[j * 3 + 1],
                             dz = (double)p[i * 3 + 2] - (double)p[j * 3 + 2];
                const double d2 = (dx * dx + dy * dy) + dz * dz;
                if (d2 < eps * eps) {
                    ++deg[i]; if (j != i) ++deg[j];
                    if (j != i) { const int a = uf_find(par, i), c = uf_find(par, j); if (a < c) par[c] = a; else par[a] = c; }
                }
            }
        int best = -1, bestsize = 0;
        for (int i = 0; i < N; ++i) if (deg[i] >= min_points) ++sz[uf_find(par, i)];
        for (int i = 0; i < N; ++i) if (sz[i] > bestsize) { bestsize = sz[i]; best = i; }   /* root = smallest index */
        float* o = out + (long long)b * TP * 3;
        out_count[b] = bestsize;
        if (best < 0) { memset(o, 0, sizeof(float) * 3 * (size_t)TP); }
        else {
            int* list = (int*)malloc(sizeof(int) * (size_t)bestsize);
            int m = 0;
            for (int i = 0; i < N; ++i) if (deg[i] >= min_points && uf_find(par, i) == best) list[m++] = i;
            for (int j = 0; j < TP; ++j) memcpy(o + (long long)j * 3, p + (long long)list[j % m] * 3, 12);
            free(list);
        }
        free(par); free(deg); free(sz);
    }
}

/* ---- voxelization ----------------------------------------------------------------------- */
static int voxel_coord(const float* p, const float* lo, const float* vs, const int* grid, int* c) {
    for (int j = 0; j < 3; ++j) {
        const float f = floorf((p[j] - lo[j]) / vs[j]);
        if (!(f >= 0.f && f < (float)grid[j])) return 0;
        c[j] = (int)f;
    }
    return 1;
}

typedef struct { long long key; int idx; } KeyIdx;
static int cmp_keyidx(const void* a, const void* b) {
    const KeyIdx* x = (const KeyIdx*)a; const KeyIdx* y = (const KeyIdx*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx - y->idx;
}

/* ref: DynamicMeanVFE.forward, detector3d/pcdet/models/backbones_3d/vfe/dynamic_mean_vfe.py:49-76.
 * points (N,1+C) [b,x,y,z,...]; key = b*XYZ + x*YZ + y*Z + z (64-bit here; the reference's int32
 * overflows past batch 23 on the Waymo grid); torch.unique -> ascending key order;
 * scatter_mean = fp32 sum / count; coords [b,z,y,x].  Returns M. */
int orc_dynamic_voxelize(int N, int C, const float* points, const float* range, const float* vs, const int* grid,
                         int* coords, float* feats, int* counts) {
    KeyIdx* ki = (KeyIdx*)malloc(sizeof(KeyIdx) * (size_t)(N > 0 ? N : 1));
    int n = 0;
    for (int p = 0; p < N; ++p) {
        const float* row = points + (long long)p * (1 + C);
        int c[3];
        if (!voxel_coord(row + 1, range, vs, grid, c)) continue;
        const long long b = (long long)(int)row[0];
        ki[n].key = ((b * grid[0] + c[0]) * grid[1] + c[1]) * (long long)grid[2] + c[2];
        ki[n].idx = p; ++n;
    }
    qsort(ki, (size_t)n, sizeof(KeyIdx), cmp_keyidx);
    int m = 0;
    for (int i = 0; i < n;) {
        int j = i;
        /* scatter_mean (third-party torch_scatter, absent): sum / count per voxel.  The CUDA implementation adds
         * fp32 values with atomics in arrival order, so its last bits vary from run to run; the order-free value it
         * approximates is restated here: float64 sum, one division, one rounding to fp32. */
        double sum[16] = {0};
        while (j < n && ki[j].key == ki[i].key) {
            const float* row = points + (long long)ki[j].idx * (1 + C);
            for (int f = 0; f < C; ++f) sum[f] += (double)row[1 + f];
            ++j;
        }
        const long long key = ki[i].key;
        coords[m * 4 + 0] = (int)(key / ((long long)grid[0] * grid[1] * grid[2]));
        coords[m * 4 + 3] = (int)((key / ((long long)grid[1] * grid[2])) % grid[0]);
        coords[m * 4 + 2] = (int)((key / grid[2]) % grid[1]);
        coords[m * 4 + 1] = (int)(key % grid[2]);
        for (int f = 0; f < C; ++f) feats[(long long)m * C + f] = (float)(sum[f] / (double)(j - i));
        counts[m] = j - i;
        ++m; i = j;
    }
    free(ki);
    return m;
}

/* ref: VoxelGeneratorWrapper.generate, detector3d/pcdet/datasets/processor/data_processor.py:44-60 ->
 * spconv (NOT vendored, unpinned: docker/Dockerfile:58) — PARITY UNPINNED.  Restates spconv v1's
 * points_to_voxel loop: scan points in order; c = floor((p - lo)/vs); skip if outside the grid;
 * unseen voxel -> new id unless max_voxels reached (then the point is skipped); append the point if
 * the voxel holds < max_points.  coordinates are zyx.  Returns M. */
int orc_hard_voxelize(int N, int C, const float* points, const float* range, const float* vs, const int* grid,
                      int max_points, int max_voxels, float* voxels, int* coordinates, int* num_points) {
    const long long cells = (long long)grid[0] * grid[1] * grid[2];
    int* lut = (int*)malloc(sizeof(int) * (size_t)cells);
    memset(lut, 0xff, sizeof(int) * (size_t)cells);
    memset(voxels, 0, sizeof(float) * (size_t)max_voxels * max_points * C);
    memset(num_points, 0, sizeof(int) * (size_t)max_voxels);
    int m = 0;
    for (int p = 0; p < N; ++p) {
        const float* row = points + (long long)p * C;
        int c[3];
        if (!voxel_coord(row, range, vs, grid, c)) continue;
        const long long cell = ((long long)c[0] * grid[1] + c[1]) * grid[2] + c[2];
        int v = lut[cell];
        if (v < 0) {
            if (m >= max_voxels) continue;
            v = m++;
            lut[cell] = v;
            coordinates[v * 3 + 0] = c[2]; coordinates[v * 3 + 1] = c[1]; coordinates[v * 3 + 2] = c[0];
        }
        if (num_points[v] < max_points) {
            memcpy(voxels + ((long long)v * max_points + num_points[v]) * C, row, sizeof(float) * (size_t)C);
            ++num_points[v];
        }
    }
    free(lut);
    return m;
}

/* ref: MeanVFE.forward, detector3d/pcdet/models/backbones_3d/vfe/mean_vfe.py:23-29 */
void orc_mean_vfe(int M, int T, int C, const float* voxels, const float* num, float* out) {
#pragma omp parallel for schedule(static)
    for (int v = 0; v < M; ++v)
        for (int f = 0; f < C; ++f) {
            float s = 0.f;
            for (int i = 0; i < T; ++i) s += voxels[((long long)v * T + i) * C + f];
            const float nrm = num[v] > 1.0f ? num[v] : 1.0f;
            out[(long long)v * C + f] = s / nrm;
        }
}

/* ref: chamfer_dist_kernel, see/surface_completion/models/vcn/extensions/chamfer_dist/chamfer.cu:15-145 */
void orc_chamfer(int B, int N, int M, const float* a, const float* b, float* d1) {
#pragma omp parallel for schedule(static) collapse(2)
    for (int bi = 0; bi < B; ++bi)
        for (int i = 0; i < N; ++i) {
            const float* p = a + ((long long)bi * N + i) * 3;
            double best = INFINITY;
            for (int j = 0; j < M; ++j) {
                const float* q = b + ((long long)bi * M + j) * 3;
                const double dx = (double)p[0] - q[0], dy = (double)p[1] - q[1], dz = (double)p[2] - q[2];
                const double d = dx * dx + dy * dy + dz * dz;
                if (d < best) best = d;
            }
            d1[(long long)bi * N + i] = (float)best;
        }
}

/* ref: SEE_VCN.replace_with_completed_pts, see/surface_completion/SEE_VCN.py:247-265 (demo twin
 * demo/see_vcn_dataset.py:127-135): dist = original_pcd.compute_point_cloud_distance(completed) — open3d,
 * NOT vendored (setup.py:25 pins 0.14.1): per original point the Euclidean distance to its nearest completed
 * point, evaluated on float64 copies of the points — PARITY UNPINNED.  Brute-force restatement: nearest
 * distance (double) per original point; the caller drops the points with dist < thresh. */
void orc_nearest_dist(int P, int K, const float* pts, const float* completed, double* dist) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        const float* p = pts + (long long)i * 3;
        double best = INFINITY;
        for (int j = 0; j < K; ++j) {
            const float* q = completed + (long long)j * 3;
            const double dx = (double)p[0] - (double)q[0], dy = (double)p[1] - (double)q[1], dz = (double)p[2] - (double)q[2];
            const double d2 = (dx * dx + dy * dy) + dz * dz;
            if (d2 < best) best = d2;
        }
        dist[i] = sqrt(best);
    }
}
