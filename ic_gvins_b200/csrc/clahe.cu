// clahe.cu -- contrast-limited adaptive histogram equalisation for sm_100a (the per-frame pre-pass of path A).
//
// Replaces `clahe_->apply(frame_cur_->image(), frame_cur_->image())` (IG/tracking/tracking.cc:141) with
// `clahe_ = cv::createCLAHE(3.0, cv::Size(21, 21))` (:62).  OpenCV is an un-vendored dependency of the reference; the algorithm is
// modules/imgproc/src/clahe.cpp (CLAHE_CalcLut_Body + CLAHE_Interpolation_Body), restated in oracle/clahe_ref.c and pinned
// bit-exactly against cv2 4.13.0 (tests/golden/clahe_golden.npz).  Compile with -fmad=false (the float sequence matters).
//
//   clahe_lut_kernel    : CTA / tile -> histogram in shared memory (integer atomics: order-independent), clip + redistribute,
//                         inclusive scan, LUT = cvRound(cdf * 255 / tileArea).  Reads the frame once (+ the reflect-101 fringe).
//   clahe_interp_kernel : thread / 4 pixels -> bilinear blend of the four neighbouring tiles' LUT entries (LUTs: 441 x 256 B = 113 KB,
//                         L1/L2 resident), one uchar4 store.  HBM-bound: W*H bytes in, W*H bytes out.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace icg {

struct ClaheArgs {
    const uint8_t *src;
    uint8_t *dst;
    uint8_t *lut;
    int W, H, spitch, dpitch;
    int tiles_x, tiles_y, tw, th;
    int clip;
    float lut_scale, inv_tw, inv_th;
    // batched frames (blockIdx.y / blockIdx.z): frame f reads src + f * src_fstride, writes dst + f * dst_fstride, uses LUT block f
    size_t src_fstride, dst_fstride;
    unsigned int *hist;  // [n_frames][256] in-image pixel counts of the RAW frame (histogram gate), or NULL
};

__device__ __forceinline__ int cl_reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

__global__ void __launch_bounds__(256) clahe_lut_kernel(ClaheArgs A) {
    __shared__ int s_hist[256];
    __shared__ int s_img[256];  // the same counts restricted to pixels INSIDE the frame (the tiles of a non-divisible grid see a reflected fringe)
    __shared__ int s_warp[8];
    const int tid = threadIdx.x, tx = blockIdx.x % A.tiles_x, ty = blockIdx.x / A.tiles_x, frame = blockIdx.y;
    A.src += (size_t) frame * A.src_fstride;
    A.lut += (size_t) frame * A.tiles_x * A.tiles_y * 256;
    s_hist[tid] = 0, s_img[tid] = 0;
    __syncthreads();
    const int area = A.tw * A.th;
    for (int p = tid; p < area; p += 256) {
        const int yy = p / A.tw, xx = p - yy * A.tw;
        const int y0 = ty * A.th + yy, x0 = tx * A.tw + xx;
        const int y = cl_reflect101(y0, A.H), x = cl_reflect101(x0, A.W);  // copyMakeBorder(BORDER_REFLECT_101)
        const int v = A.src[(size_t) y * A.spitch + x];
        atomicAdd(&s_hist[v], 1);
        if (A.hist && y0 < A.H && x0 < A.W) atomicAdd(&s_img[v], 1);
    }
    __syncthreads();
    // Tracking::calculateHistigram's cv::calcHist over the raw frame comes for free from the tile pass: integer atomics, order-independent
    if (A.hist && s_img[tid]) atomicAdd(&A.hist[(size_t) frame * 256 + tid], (unsigned int) s_img[tid]);
    int h = s_hist[tid];
    if (A.clip > 0) {
        int over = h > A.clip ? h - A.clip : 0;
        if (h > A.clip) h = A.clip;
        // clipped = sum of the excess (integers: any order gives the same value)
        for (int o = 16; o > 0; o >>= 1) over += __shfl_xor_sync(0xffffffffu, over, o);
        if ((tid & 31) == 0) s_warp[tid >> 5] = over;
        __syncthreads();
        int clipped = 0;
        for (int k = 0; k < 8; k++) clipped += s_warp[k];
        __syncthreads();
        const int batch = clipped / 256, residual = clipped - batch * 256;
        h += batch;
        if (residual != 0) {
            const int step = max(256 / residual, 1);
            if (tid % step == 0 && tid / step < residual) h++;  // for (i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++
        }
    }
    // inclusive scan over the 256 bins
    int v = h;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if ((tid & 31) >= o) v += t;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = v;
    __syncthreads();
    int base = 0;
    for (int k = 0; k < (tid >> 5); k++) base += s_warp[k];
    const int sum = v + base;
    int r = __float2int_rn((float) sum * A.lut_scale);  // saturate_cast<uchar>(sum * lutScale): cvRound
    r = r < 0 ? 0 : r > 255 ? 255 : r;
    A.lut[(size_t) blockIdx.x * 256 + tid] = (uint8_t) r;
}

// Tracking::calculateHistigram (IG/tracking/tracking.cc:88-104) from the counts: sum_k (float) hist[k] * (float) k / 256.0 accumulated in double in
// bin order, divided by cols * rows (one thread per frame: the order of the 256 additions is part of the result)
__global__ void clahe_hist_stat_kernel(const unsigned int *hist, int n_frames, int W, int H, double *out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    double acc = 0;
    for (int k = 0; k < 256; k++) {
        const float prod = (float) hist[(size_t) f * 256 + k] * (float) k;
        acc += (double) prod / 256.0;
    }
    out[f] = acc / (double) (W * H);
}

__global__ void __launch_bounds__(256) clahe_interp_kernel(ClaheArgs A) {
    const int x0 = 4 * (blockIdx.x * 64 + (threadIdx.x & 63)), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    A.src += (size_t) blockIdx.z * A.src_fstride, A.dst += (size_t) blockIdx.z * A.dst_fstride;
    A.lut += (size_t) blockIdx.z * A.tiles_x * A.tiles_y * 256;
    if (x0 >= A.W || y >= A.H) return;
    const float tyf = y * A.inv_th - 0.5f;
    int ty1 = __float2int_rd(tyf), ty2 = ty1 + 1;
    const float ya = tyf - ty1, ya1 = 1.0f - ya;
    ty1 = max(ty1, 0), ty2 = min(ty2, A.tiles_y - 1);
    const uint8_t *p1 = A.lut + (size_t) ty1 * A.tiles_x * 256, *p2 = A.lut + (size_t) ty2 * A.tiles_x * 256;
    const uint8_t *srow = A.src + (size_t) y * A.spitch;
    uint8_t out[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int x = x0 + k;
        out[k] = 0;
        if (x < A.W) {
            const float txf = x * A.inv_tw - 0.5f;
            int tx1 = __float2int_rd(txf), tx2 = tx1 + 1;
            const float xa = txf - tx1, xa1 = 1.0f - xa;
            tx1 = max(tx1, 0), tx2 = min(tx2, A.tiles_x - 1);
            const int v = srow[x];
            const int i1 = tx1 * 256 + v, i2 = tx2 * 256 + v;
            const float res = (__ldg(p1 + i1) * xa1 + __ldg(p1 + i2) * xa) * ya1 + (__ldg(p2 + i1) * xa1 + __ldg(p2 + i2) * xa) * ya;
            int r = __float2int_rn(res);
            out[k] = (uint8_t) (r < 0 ? 0 : r > 255 ? 255 : r);
        }
    }
    uint8_t *drow = A.dst + (size_t) y * A.dpitch;
    if (x0 + 3 < A.W && (((size_t) (drow + x0)) & 3) == 0) {
        *(uchar4 *) (drow + x0) = make_uchar4(out[0], out[1], out[2], out[3]);
    } else {
        for (int k = 0; k < 4; k++)
            if (x0 + k < A.W) drow[x0 + k] = out[k];
    }
}

}  // namespace icg

using namespace icg;

struct icg_clahe {
    int W, H, tiles_x, tiles_y, device;
    double clip_limit;
    cudaStream_t stream;
    bool own_stream;
    uint8_t *d_img, *d_lut;
    int pitch;
    ClaheArgs A;
    int lut_frames = 1;             // frames the LUT buffer holds (grown by the batched entry point)
    unsigned int *d_hist = nullptr; // [lut_frames][256]
    double *d_stat = nullptr, *h_stat = nullptr;
};

static int clahe_reserve(icg_clahe *h, int n_frames) {
    if (n_frames <= h->lut_frames && h->d_hist) return ICG_OK;
    ICG_CUDA(cudaStreamSynchronize(h->stream));
    const int nf = std::max(n_frames, h->lut_frames);
    if (h->d_lut) cudaFree(h->d_lut);
    if (h->d_hist) cudaFree(h->d_hist);
    if (h->d_stat) cudaFree(h->d_stat);
    if (h->h_stat) cudaFreeHost(h->h_stat);
    h->d_lut = nullptr, h->d_hist = nullptr, h->d_stat = nullptr, h->h_stat = nullptr;
    if (cudaMalloc(&h->d_lut, (size_t) nf * h->tiles_x * h->tiles_y * 256) != cudaSuccess || cudaMalloc(&h->d_hist, sizeof(unsigned int) * 256 * (size_t) nf) != cudaSuccess ||
        cudaMalloc(&h->d_stat, sizeof(double) * nf) != cudaSuccess || cudaMallocHost(&h->h_stat, sizeof(double) * nf) != cudaSuccess) {
        set_error("icg_clahe: allocation for %d frames failed", nf);
        return ICG_ENOMEM;
    }
    h->lut_frames = nf;
    h->A.lut = h->d_lut;
    return ICG_OK;
}

static int clahe_launch(icg_clahe *h, const uint8_t *dsrc, int spitch, uint8_t *ddst, int dpitch, int n_frames = 1, size_t src_fstride = 0,
                        size_t dst_fstride = 0, bool want_hist = false) {
    ClaheArgs A = h->A;
    A.src = dsrc, A.dst = ddst, A.spitch = spitch, A.dpitch = dpitch, A.src_fstride = src_fstride, A.dst_fstride = dst_fstride;
    A.hist = want_hist ? h->d_hist : nullptr;
    if (want_hist) ICG_CUDA(cudaMemsetAsync(h->d_hist, 0, sizeof(unsigned int) * 256 * (size_t) n_frames, h->stream));
    clahe_lut_kernel<<<dim3(h->tiles_x * h->tiles_y, n_frames), 256, 0, h->stream>>>(A);
    ICG_CHECK_LAUNCH();
    clahe_interp_kernel<<<dim3((h->W + 255) / 256, (h->H + 3) / 4, n_frames), 256, 0, h->stream>>>(A);
    ICG_CHECK_LAUNCH();
    count_launch(2);
    if (want_hist) {
        clahe_hist_stat_kernel<<<(n_frames + 63) / 64, 64, 0, h->stream>>>(h->d_hist, n_frames, h->W, h->H, h->d_stat);
        ICG_CHECK_LAUNCH();
        count_launch();
    }
    return ICG_OK;
}

extern "C" {

int icg_clahe_create(icg_clahe **out, int width, int height, int tiles_x, int tiles_y, double clip_limit, int device, void *stream) {
    if (!out || width < 1 || height < 1 || tiles_x < 1 || tiles_y < 1 || tiles_x > width || tiles_y > height) {
        set_error("icg_clahe_create: bad arguments");
        return ICG_EINVAL;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("icg_clahe_create: no CUDA device (this library has no CPU fallback)");
        return ICG_ENODEVICE;
    }
    if (device < 0 || device >= ndev) {
        set_error("icg_clahe_create: device %d out of range", device);
        return ICG_EINVAL;
    }
    ICG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ICG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("icg_clahe_create: device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
        return ICG_ENODEVICE;
    }
    icg_clahe *h = new icg_clahe();
    h->W = width, h->H = height, h->tiles_x = tiles_x, h->tiles_y = tiles_y, h->device = device, h->clip_limit = clip_limit;
    h->own_stream = stream == nullptr;
    if (stream)
        h->stream = (cudaStream_t) stream;
    else
        ICG_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->pitch = (width + 15) & ~15;
    ICG_CUDA(cudaMalloc(&h->d_img, (size_t) h->pitch * height));
    ICG_CUDA(cudaMalloc(&h->d_lut, (size_t) tiles_x * tiles_y * 256));
    // CLAHE_Impl::apply (clahe.cpp): pad right / bottom to a multiple of the grid, tile size from the padded image
    int Wp = width, Hp = height;
    if (width % tiles_x != 0 || height % tiles_y != 0) {
        Wp = width + (tiles_x - (width % tiles_x));
        Hp = height + (tiles_y - (height % tiles_y));
    }
    ClaheArgs &A = h->A;
    memset(&A, 0, sizeof(A));
    A.lut = h->d_lut, A.W = width, A.H = height, A.tiles_x = tiles_x, A.tiles_y = tiles_y, A.tw = Wp / tiles_x, A.th = Hp / tiles_y;
    const int area = A.tw * A.th;
    A.lut_scale = (float) 255 / area;
    A.clip = 0;
    if (clip_limit > 0.0) {
        A.clip = (int) (clip_limit * area / 256);
        if (A.clip < 1) A.clip = 1;
    }
    A.inv_tw = 1.0f / A.tw, A.inv_th = 1.0f / A.th;
    *out = h;
    return ICG_OK;
}

void icg_clahe_destroy(icg_clahe *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->d_img);
    cudaFree(h->d_lut);
    if (h->d_hist) cudaFree(h->d_hist);
    if (h->d_stat) cudaFree(h->d_stat);
    if (h->h_stat) cudaFreeHost(h->h_stat);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

int icg_clahe_apply(icg_clahe *h, const uint8_t *src, int src_stride, uint8_t *dst, int dst_stride) {
    if (!h || !src || !dst || src_stride < h->W || dst_stride < h->W) {
        set_error("icg_clahe_apply: bad arguments");
        return ICG_EINVAL;
    }
    ICG_CUDA(cudaSetDevice(h->device));
    ICG_CUDA(cudaMemcpy2DAsync(h->d_img, h->pitch, src, src_stride, h->W, h->H, cudaMemcpyHostToDevice, h->stream));
    int rc = clahe_launch(h, h->d_img, h->pitch, h->d_img, h->pitch);  // in place on the device copy
    if (rc != ICG_OK) return rc;
    ICG_CUDA(cudaMemcpy2DAsync(dst, dst_stride, h->d_img, h->pitch, h->W, h->H, cudaMemcpyDeviceToHost, h->stream));
    ICG_CUDA(cudaStreamSynchronize(h->stream));
    return ICG_OK;
}

int icg_clahe_apply_dev(icg_clahe *h, const uint8_t *dev_src, int src_pitch, uint8_t *dev_dst, int dst_pitch) {
    if (!h || !dev_src || !dev_dst || src_pitch < h->W || dst_pitch < h->W) {
        set_error("icg_clahe_apply_dev: bad arguments");
        return ICG_EINVAL;
    }
    ICG_CUDA(cudaSetDevice(h->device));
    return clahe_launch(h, dev_src, src_pitch, dev_dst, dst_pitch);
}

int icg_clahe_apply_batch_dev(icg_clahe *h, int n_frames, const uint8_t *dev_src, int src_pitch, size_t src_frame_stride, uint8_t *dev_dst, int dst_pitch,
                              size_t dst_frame_stride, double *hist_out) {
    if (!h || n_frames < 1 || !dev_src || !dev_dst || src_pitch < h->W || dst_pitch < h->W) {
        set_error("icg_clahe_apply_batch_dev: bad arguments");
        return ICG_EINVAL;
    }
    ICG_CUDA(cudaSetDevice(h->device));
    int rc = clahe_reserve(h, n_frames);
    if (rc != ICG_OK) return rc;
    rc = clahe_launch(h, dev_src, src_pitch, dev_dst, dst_pitch, n_frames, src_frame_stride, dst_frame_stride, hist_out != nullptr);
    if (rc != ICG_OK) return rc;
    if (hist_out) {  // the gate statistic is consumed by the host (Tracking::preprocessing decides whether to skip the frame): synchronise
        ICG_CUDA(cudaMemcpyAsync(h->h_stat, h->d_stat, sizeof(double) * n_frames, cudaMemcpyDeviceToHost, h->stream));
        ICG_CUDA(cudaStreamSynchronize(h->stream));
        memcpy(hist_out, h->h_stat, sizeof(double) * n_frames);
    }
    return ICG_OK;
}

int icg_clahe_sync(icg_clahe *h) {
    if (!h) return ICG_EINVAL;
    ICG_CUDA(cudaSetDevice(h->device));
    ICG_CUDA(cudaStreamSynchronize(h->stream));
    return ICG_OK;
}

}  // extern "C"
