// extern "C" doors onto the reference's own kernel launchers (compiled from /root/reference by
// oracle/Makefile `make ref`).  TEST INFRASTRUCTURE ONLY — gives the parity tests the exact GPU
// reference for points_in_boxes / FPS / gather / group on the B200 box.
// Declarations follow:
//   roiaware_pool3d_kernel.cu:339   sampling_gpu.cu:33,218   group_points_gpu.cu:75
#include <cuda_runtime.h>

void points_in_boxes_launcher(int batch_size, int boxes_num, int pts_num, const float* boxes, const float* pts,
                              int* box_idx_of_points);
void farthest_point_sampling_kernel_launcher(int b, int n, int m, const float* dataset, float* temp, int* idxs);
void gather_points_kernel_launcher_fast(int b, int c, int n, int npoints, const float* points, const int* idx, float* out);
void group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx,
                                       float* out);

extern "C" {
int ref_points_in_boxes(int batch_size, int boxes_num, int pts_num, const float* boxes, const float* pts, int* out) {
    points_in_boxes_launcher(batch_size, boxes_num, pts_num, boxes, pts, out);
    return (int)cudaDeviceSynchronize();
}
int ref_fps(int b, int n, int m, const float* dataset, float* temp, int* idxs) {
    farthest_point_sampling_kernel_launcher(b, n, m, dataset, temp, idxs);
    return (int)cudaDeviceSynchronize();
}
int ref_gather(int b, int c, int n, int npoints, const float* points, const int* idx, float* out) {
    gather_points_kernel_launcher_fast(b, c, n, npoints, points, idx, out);
    return (int)cudaDeviceSynchronize();
}
int ref_group(int b, int c, int n, int npoints, int nsample, const float* points, const int* idx, float* out) {
    group_points_kernel_launcher_fast(b, c, n, npoints, nsample, points, idx, out);
    return (int)cudaDeviceSynchronize();
}
}
