// geom.cu -- SURVEY.md 8f ranks 2-4 on the device: the point-wise geometry of the front end and the IMU propagation, batched.
//
//   icg_geom_undistort_points / icg_geom_distort_points   Camera::undistortPoints / distortPoints (IG/tracking/camera.cc:72-104), thread / point
//   icg_geom_find_fundamental_mat_ransac                   cv::findFundamentalMat(FM_RANSAC) of Tracking::trackReferenceFrame (tracking.cc:547):
//        the subsets of the cv::RNG stream are drawn up front on the host (they depend on the data only through the collinearity
//        re-draws, never on which model wins), ALL hypotheses are solved (thread / subset: 7-point null space + cubic) and scored
//        (warp / model against the <= 300 pairs) on the device, then the host replays OpenCV's serial acceptance rule and adaptive
//        iteration bound on the inlier counts -- the serial semantics are kept, the work is parallel
//   icg_geom_triangulate_points                            Tracking::triangulatePoint (tracking.cc:796-808), thread / pair (4 x 4 one-sided Jacobi)
//   icg_geom_imu_preintegrate_batch                        PreintegrationEarth / Normal propagation (preintegration_earth.cc:205-303) of many
//        intervals at once (doReintegration, IG/ic_gvins.cc:1680-1695; throughput mode): thread / interval, sequential inside
//
// The arithmetic is the host code of camera.cu / fundamental.cu / ba.cu compiled for the device (shared __host__ __device__ cores in
// geom_core.cuh), -fmad=false: same operation order, IEEE double; only libm calls (sin, cos, acos, pow, log) differ in the last ulp.
#include <float.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "geom_core.cuh"

namespace icg {

__global__ void geom_undistort_kernel(icg_camera c, float *pts, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gc::undistort_point(c, pts + 2 * (size_t) i);
}
__global__ void geom_distort_kernel(icg_camera c, float *pts, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gc::distort_point(c, pts + 2 * (size_t) i);
}

// thread / hypothesis: 7-point models of subset `it` (indices idx[7 it .. 7 it + 6]) -> nmodels[it], F[it][3][9]
__global__ void geom_ransac_models_kernel(const float *p1, const float *p2, const int *idx, int niter, int *nmodels, double *F) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= niter) return;
    float ms1[14], ms2[14];
    for (int i = 0; i < 7; i++) {
        const int v = idx[7 * it + i];
        ms1[2 * i] = p1[2 * v], ms1[2 * i + 1] = p1[2 * v + 1], ms2[2 * i] = p2[2 * v], ms2[2 * i + 1] = p2[2 * v + 1];
    }
    double Fm[3][9];
    const int nm = gc::run_7point(ms1, ms2, Fm);
    nmodels[it] = nm;
    for (int k = 0; k < 3; k++)
        for (int e = 0; e < 9; e++) F[((size_t) it * 3 + k) * 9 + e] = (k < nm && nm > 0) ? Fm[k][e] : 0.0;
}
// warp / model: inlier count of model (it, k) over the n pairs (FMEstimatorCallback::computeError + findInliers)
__global__ void geom_ransac_score_kernel(const float *p1, const float *p2, int n, const int *nmodels, const double *F, int niter, double thresh, int *good) {
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (m >= niter * 3) return;
    const int it = m / 3, k = m - 3 * it;
    int cnt = 0;
    if (k < nmodels[it]) {
        const double *Fm = F + (size_t) m * 9;
        for (int i = lane; i < n; i += 32) cnt += gc::is_inlier(p1, p2, i, Fm, thresh) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) good[m] = cnt;
}
__global__ void geom_ransac_mask_kernel(const float *p1, const float *p2, int n, const double *Fm, double thresh, uint8_t *mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mask[i] = gc::is_inlier(p1, p2, i, Fm, thresh) ? 1 : 0;
}

__global__ void geom_triangulate_kernel(const double *Tcw0, const double *Tcw1, const double *pc0, const double *pc1, int n, double *pw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gc::triangulate_point(Tcw0 + 12 * (size_t) i, Tcw1, pc0 + 2 * (size_t) i, pc1 + 2 * (size_t) i, pw + 3 * (size_t) i);
}

__global__ void geom_preintegrate_kernel(int n_int, const double *state16, const double *iewn3, const double *gravity3, const double *noise5, const double *imu,
                                         const int *imu_off, double *blobs, double *ends) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_int) return;
    gc::preintegrate_core(state16 + 16 * (size_t) k, iewn3, gravity3, noise5, imu + 7 * (size_t) imu_off[k], imu_off[k + 1] - imu_off[k],
                          blobs + (size_t) ICG_IMU_BLOB_DOUBLES * k, ends ? ends + 10 * (size_t) k : nullptr);
}

}  // namespace icg

using namespace icg;

struct icg_geom {
    int device;
    cudaStream_t stream;
    bool own_stream;
    uint8_t *d_buf = nullptr, *h_buf = nullptr;  // one device + one pinned scratch arena, grown on demand
    size_t d_bytes = 0, h_bytes = 0;
};

static int geom_reserve(icg_geom *h, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t) 255;
    if (h->d_bytes >= bytes) return ICG_OK;
    ICG_CUDA(cudaStreamSynchronize(h->stream));
    if (h->d_buf) cudaFree(h->d_buf);
    if (h->h_buf) cudaFreeHost(h->h_buf);
    h->d_buf = h->h_buf = nullptr, h->d_bytes = h->h_bytes = 0;
    if (cudaMalloc(&h->d_buf, bytes) != cudaSuccess || cudaMallocHost(&h->h_buf, bytes) != cudaSuccess) {
        set_error("icg_geom: scratch allocation of %zu bytes failed", bytes);
        return ICG_ENOMEM;
    }
    h->d_bytes = h->h_bytes = bytes;
    return ICG_OK;
}

extern "C" {

int icg_geom_create(icg_geom **out, int device, void *stream) {
    if (!out) {
        set_error("icg_geom_create: bad arguments");
        return ICG_EINVAL;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("icg_geom_create: no CUDA device (this library has no CPU fallback)");
        return ICG_ENODEVICE;
    }
    if (device < 0 || device >= ndev) {
        set_error("icg_geom_create: device %d out of range", device);
        return ICG_EINVAL;
    }
    ICG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ICG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("icg_geom_create: device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
        return ICG_ENODEVICE;
    }
    icg_geom *h = new icg_geom();
    h->device = device, h->own_stream = stream == nullptr;
    if (stream) {
        h->stream = (cudaStream_t) stream;
    } else if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        set_error("icg_geom_create: stream creation failed");
        return ICG_ECUDA;
    }
    *out = h;
    return ICG_OK;
}

void icg_geom_destroy(icg_geom *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->d_buf) cudaFree(h->d_buf);
    if (h->h_buf) cudaFreeHost(h->h_buf);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

static int geom_points(icg_geom *h, const icg_camera *c, float *pts_xy, int n, int undistort) {
    if (!h || !c || (!pts_xy && n > 0) || n < 0 || !(c->fx != 0.0) || !(c->fy != 0.0)) {
        set_error("icg_geom_%s_points: bad arguments", undistort ? "undistort" : "distort");
        return ICG_EINVAL;
    }
    if (n == 0) return ICG_OK;
    ICG_CUDA(cudaSetDevice(h->device));
    int rc = geom_reserve(h, sizeof(float) * 2 * (size_t) n);
    if (rc != ICG_OK) return rc;
    memcpy(h->h_buf, pts_xy, sizeof(float) * 2 * (size_t) n);
    ICG_CUDA(cudaMemcpyAsync(h->d_buf, h->h_buf, sizeof(float) * 2 * (size_t) n, cudaMemcpyHostToDevice, h->stream));
    if (undistort)
        geom_undistort_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(*c, (float *) h->d_buf, n);
    else
        geom_distort_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(*c, (float *) h->d_buf, n);
    ICG_CHECK_LAUNCH();
    count_launch();
    ICG_CUDA(cudaMemcpyAsync(h->h_buf, h->d_buf, sizeof(float) * 2 * (size_t) n, cudaMemcpyDeviceToHost, h->stream));
    ICG_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(pts_xy, h->h_buf, sizeof(float) * 2 * (size_t) n);
    return ICG_OK;
}
int icg_geom_undistort_points(icg_geom *h, const icg_camera *c, float *pts_xy, int n) { return geom_points(h, c, pts_xy, n, 1); }
int icg_geom_distort_points(icg_geom *h, const icg_camera *c, float *pts_xy, int n) { return geom_points(h, c, pts_xy, n, 0); }

int icg_geom_find_fundamental_mat_ransac(icg_geom *h, const float *pts1_xy, const float *pts2_xy, int n, double threshold, double confidence, int max_iters,
                                         uint8_t *status, double *F9) {
    if (!h || !pts1_xy || !pts2_xy || !status || n < 0 || max_iters < 1) {
        set_error("icg_geom_find_fundamental_mat_ransac: bad arguments");
        return ICG_EINVAL;
    }
    if (n < 15) {
        set_error("icg_geom_find_fundamental_mat_ransac: needs at least 15 point pairs (got %d)", n);
        return ICG_EUNSUPPORTED;
    }
    if (threshold <= 0) threshold = 3;
    if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
    ICG_CUDA(cudaSetDevice(h->device));
    // ---- host: every subset the serial loop could ask for, in the order of the cv::RNG stream (RANSACPointSetRegistrator::getSubset)
    std::vector<int> idx((size_t) 7 * max_iters);
    gc::CvRng rng;
    int drawn = 0;
    for (int iter = 0; iter < max_iters; iter++) {
        if (!gc::draw_subset(rng, pts1_xy, pts2_xy, n, idx.data() + 7 * (size_t) iter)) break;
        drawn++;
    }
    if (drawn == 0) {  // the first subset could not be drawn: OpenCV returns no model
        memset(status, 0, n);
        if (F9) memset(F9, 0, sizeof(double) * 9);
        return ICG_OK;
    }
    // ---- device: all models, all inlier counts
    const size_t o_p1 = 0, o_p2 = o_p1 + sizeof(float) * 2 * (size_t) n, o_idx = (o_p2 + sizeof(float) * 2 * (size_t) n + 15) & ~(size_t) 15;
    const size_t o_nm = o_idx + sizeof(int) * 7 * (size_t) drawn, o_good = o_nm + sizeof(int) * (size_t) drawn;
    const size_t o_F = (o_good + sizeof(int) * 3 * (size_t) drawn + 15) & ~(size_t) 15, o_mask = o_F + sizeof(double) * 27 * (size_t) drawn;
    const size_t total = o_mask + (size_t) n + 16;
    int rc = geom_reserve(h, total);
    if (rc != ICG_OK) return rc;
    memcpy(h->h_buf + o_p1, pts1_xy, sizeof(float) * 2 * (size_t) n);
    memcpy(h->h_buf + o_p2, pts2_xy, sizeof(float) * 2 * (size_t) n);
    memcpy(h->h_buf + o_idx, idx.data(), sizeof(int) * 7 * (size_t) drawn);
    ICG_CUDA(cudaMemcpyAsync(h->d_buf, h->h_buf, o_nm, cudaMemcpyHostToDevice, h->stream));
    const float *d1 = (const float *) (h->d_buf + o_p1), *d2 = (const float *) (h->d_buf + o_p2);
    int *d_nm = (int *) (h->d_buf + o_nm), *d_good = (int *) (h->d_buf + o_good);
    double *d_F = (double *) (h->d_buf + o_F);
    geom_ransac_models_kernel<<<(drawn + 63) / 64, 64, 0, h->stream>>>(d1, d2, (const int *) (h->d_buf + o_idx), drawn, d_nm, d_F);
    geom_ransac_score_kernel<<<(3 * drawn * 32 + 255) / 256, 256, 0, h->stream>>>(d1, d2, n, d_nm, d_F, drawn, threshold, d_good);
    ICG_CHECK_LAUNCH();
    count_launch(2);
    ICG_CUDA(cudaMemcpyAsync(h->h_buf + o_nm, h->d_buf + o_nm, o_mask - o_nm, cudaMemcpyDeviceToHost, h->stream));
    ICG_CUDA(cudaStreamSynchronize(h->stream));
    // ---- host: replay of RANSACPointSetRegistrator::run's acceptance rule on the counts (ptsetreg.cpp): strictly better than the best so
    //      far and than 6, iteration bound updated after every improvement
    const int *nm = (const int *) (h->h_buf + o_nm), *good = (const int *) (h->h_buf + o_good);
    int niters = max_iters, max_good = 0, best = -1;
    for (int iter = 0; iter < niters && iter < drawn; iter++) {
        if (nm[iter] <= 0) continue;
        for (int k = 0; k < nm[iter]; k++) {
            const int g = good[3 * iter + k];
            if (g > std::max(max_good, 6)) {
                best = 3 * iter + k, max_good = g;
                niters = gc::update_num_iters(confidence, (double) (n - g) / n, 7, niters);
            }
        }
    }
    if (best < 0) {
        memset(status, 0, n);
        if (F9) memset(F9, 0, sizeof(double) * 9);
        return ICG_OK;
    }
    geom_ransac_mask_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(d1, d2, n, d_F + (size_t) best * 9, threshold, h->d_buf + o_mask);
    ICG_CHECK_LAUNCH();
    count_launch();
    ICG_CUDA(cudaMemcpyAsync(h->h_buf + o_mask, h->d_buf + o_mask, n, cudaMemcpyDeviceToHost, h->stream));
    ICG_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(status, h->h_buf + o_mask, n);
    if (F9) memcpy(F9, (const double *) (h->h_buf + o_F) + (size_t) best * 9, sizeof(double) * 9);
    return ICG_OK;
}

int icg_geom_triangulate_points(icg_geom *h, const double *Tcw0, const double *Tcw1, const double *pc0_xy, const double *pc1_xy, int n, double *pw_xyz) {
    if (!h || n < 0 || ((!Tcw0 || !Tcw1 || !pc0_xy || !pc1_xy || !pw_xyz) && n > 0)) {
        set_error("icg_geom_triangulate_points: bad arguments");
        return ICG_EINVAL;
    }
    if (n == 0) return ICG_OK;
    ICG_CUDA(cudaSetDevice(h->device));
    const size_t N = n, o_T0 = 0, o_T1 = o_T0 + 96 * N, o_c0 = o_T1 + 96, o_c1 = o_c0 + 16 * N, o_pw = o_c1 + 16 * N, total = o_pw + 24 * N;
    int rc = geom_reserve(h, total);
    if (rc != ICG_OK) return rc;
    memcpy(h->h_buf + o_T0, Tcw0, 96 * N), memcpy(h->h_buf + o_T1, Tcw1, 96), memcpy(h->h_buf + o_c0, pc0_xy, 16 * N), memcpy(h->h_buf + o_c1, pc1_xy, 16 * N);
    ICG_CUDA(cudaMemcpyAsync(h->d_buf, h->h_buf, o_pw, cudaMemcpyHostToDevice, h->stream));
    geom_triangulate_kernel<<<(n + 63) / 64, 64, 0, h->stream>>>((const double *) (h->d_buf + o_T0), (const double *) (h->d_buf + o_T1),
                                                                (const double *) (h->d_buf + o_c0), (const double *) (h->d_buf + o_c1), n,
                                                                (double *) (h->d_buf + o_pw));
    ICG_CHECK_LAUNCH();
    count_launch();
    ICG_CUDA(cudaMemcpyAsync(h->h_buf + o_pw, h->d_buf + o_pw, 24 * N, cudaMemcpyDeviceToHost, h->stream));
    ICG_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(pw_xyz, h->h_buf + o_pw, 24 * N);
    return ICG_OK;
}

int icg_geom_imu_preintegrate_batch(icg_geom *h, int n_intervals, const double *state16, const double *iewn3, const double *gravity3, const double *noise5,
                                    const double *imu, const int32_t *imu_off, double *blobs_out, double *end_states10) {
    if (!h || n_intervals < 1 || !state16 || !gravity3 || !noise5 || !imu || !imu_off || !blobs_out) {
        set_error("icg_geom_imu_preintegrate_batch: bad arguments");
        return ICG_EINVAL;
    }
    for (int k = 0; k < n_intervals; k++)
        if (imu_off[k + 1] - imu_off[k] < 1 || imu_off[k] < 0) {
            set_error("icg_geom_imu_preintegrate_batch: interval %d has no samples", k);
            return ICG_EINVAL;
        }
    ICG_CUDA(cudaSetDevice(h->device));
    const size_t NI = n_intervals, ns = imu_off[n_intervals];
    const size_t o_st = 0, o_iw = o_st + 128 * NI, o_g = o_iw + 24, o_nz = o_g + 24, o_imu = o_nz + 40, o_off = o_imu + 56 * ns;
    const size_t o_blob = (o_off + 4 * (NI + 1) + 15) & ~(size_t) 15, o_end = o_blob + sizeof(double) * ICG_IMU_BLOB_DOUBLES * NI, total = o_end + 80 * NI;
    int rc = geom_reserve(h, total);
    if (rc != ICG_OK) return rc;
    memcpy(h->h_buf + o_st, state16, 128 * NI);
    if (iewn3) memcpy(h->h_buf + o_iw, iewn3, 24);
    memcpy(h->h_buf + o_g, gravity3, 24), memcpy(h->h_buf + o_nz, noise5, 40), memcpy(h->h_buf + o_imu, imu, 56 * ns), memcpy(h->h_buf + o_off, imu_off, 4 * (NI + 1));
    ICG_CUDA(cudaMemcpyAsync(h->d_buf, h->h_buf, o_blob, cudaMemcpyHostToDevice, h->stream));
    geom_preintegrate_kernel<<<(n_intervals + 31) / 32, 32, 0, h->stream>>>(n_intervals, (const double *) (h->d_buf + o_st),
                                                                           iewn3 ? (const double *) (h->d_buf + o_iw) : nullptr, (const double *) (h->d_buf + o_g),
                                                                           (const double *) (h->d_buf + o_nz), (const double *) (h->d_buf + o_imu),
                                                                           (const int *) (h->d_buf + o_off), (double *) (h->d_buf + o_blob),
                                                                           (double *) (h->d_buf + o_end));
    ICG_CHECK_LAUNCH();
    count_launch();
    ICG_CUDA(cudaMemcpyAsync(h->h_buf + o_blob, h->d_buf + o_blob, total - o_blob, cudaMemcpyDeviceToHost, h->stream));
    ICG_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(blobs_out, h->h_buf + o_blob, sizeof(double) * ICG_IMU_BLOB_DOUBLES * NI);
    if (end_states10) memcpy(end_states10, h->h_buf + o_end, 80 * NI);
    return ICG_OK;
}

}  // extern "C"
