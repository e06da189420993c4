// detect.cu -- Path A, detection leg: Shi-Tomasi block detection + sub-pixel refinement for sm_100a.
//
// Replaces the tbb::parallel_for body of Tracking::featuresDetection (IG/tracking/tracking.cc:627-656):
//     cv::goodFeaturesToTrack(block_image, out, n, 0.01, track_min_pixel_distance_, block_mask)      (:647)
//     cv::cornerSubPix(block_image, out, Size(5,5), Size(-1,-1), TermCriteria(COUNT+EPS, 20, 0.01))    (:651)
// for ALL blocks of a frame in one call.  OpenCV is an un-vendored dependency of the reference; the arithmetic restated here
// (float sequences matched bit-for-bit against cv2 4.13.0) is pinned by oracle/detect_ref.c and
// tests/golden/detect_golden.npz.  ROI semantics as in C++: the Sobel derivative of a block reads the frame's pixels beyond
// the block edge, the 3x3 covariance box filter reflects at the block edge, getRectSubPix replicates at the block edge.
//
// Kernels (all HBM/L2 streaming or tiny):  detect_eig (min-eigenvalue map + masked per-block maximum),
// detect_nms (threshold, 3x3 non-maximum suppression, candidate list), detect_select (CTA per block: bitonic sort by
// (value desc, address desc) + greedy min-distance grid), detect_subpix (warp per corner).  Compile with -fmad=false; the two
// fused multiply-adds OpenCV's AVX2 Sobel performs are written explicitly with __fmaf_rn.
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace icg {

constexpr int DET_MAX_CAND = 16384;  // candidates per block after NMS (sorted in 128 KB of dynamic shared memory)
constexpr int DET_CELL_CAP = 4;      // accepted corners per min-distance grid cell (cell = round(minDistance) -> at most 4)
constexpr int DET_MAX_CELLS = 1024;

struct DetRect {
    int x, y, w, h;
};

struct DetArgs {
    const uint8_t *img;
    const uint8_t *mask;  // may be null
    int W, H, pitch;
    int n_blocks, cap;    // cap = corner capacity per block in the outputs
    int roi_cap;          // eig plane capacity per block (floats)
    const DetRect *rois;
    const int *max_corners;
    float *eig;                 // [n_blocks][roi_cap]
    unsigned int *maxkey;       // [n_blocks] order-preserving key of the masked maximum
    unsigned long long *cand;   // [n_blocks][DET_MAX_CAND]
    int *ncand;                 // [n_blocks]
    float *out_xy;              // [n_blocks][cap][2]
    int *out_n;                 // [n_blocks]
    int *overflow;
    double quality, min_distance;
    // batched frames: block b belongs to frame b / bpf, whose image (and mask) start frame * img_stride bytes further
    size_t img_stride;
    int bpf;
};
// kernel prologue: point the by-value argument copy at the block's frame
#define DET_FRAME(A, b)                                           \
    do {                                                          \
        const size_t fo_ = (size_t) ((b) / (A).bpf) * (A).img_stride; \
        (A).img += fo_;                                           \
        if ((A).mask) (A).mask += fo_;                            \
    } while (0)

__device__ __forceinline__ int refl(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}
__device__ __forceinline__ unsigned int float_key(float v) {  // monotonic float -> uint
    unsigned int b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// cv::Sobel(..., CV_32F, ksize 3, scale = 1/(4*3*255)) at frame pixel (x, y): reflect-101 at the FRAME border only
// fma_row: OpenCV's AVX2 u8->f32 row filter covers 32 pixels per iteration with FMA; the last (roi_width % 32) columns of a row
// take the scalar loop, which rounds every product and sum separately (measured against cv2 4.13.0).
__device__ __forceinline__ void sobel_at(const DetArgs &A, int x, int y, bool fma_row, float &dx, float &dy) {
    const float s = (float) (1.0 / (4 * 3 * 255.0)), s2 = (float) (2.0 * (1.0 / (4 * 3 * 255.0)));
    const int xm = refl(x - 1, A.W), xp = refl(x + 1, A.W), ym = refl(y - 1, A.H), yp = refl(y + 1, A.H);
    const uint8_t *r0 = A.img + (size_t) ym * A.pitch, *r1 = A.img + (size_t) y * A.pitch, *r2 = A.img + (size_t) yp * A.pitch;
    const float a00 = r0[xm], a01 = r0[x], a02 = r0[xp], a10 = r1[xm], a12 = r1[xp], a20 = r2[xm], a21 = r2[x], a22 = r2[xp];
    const float top = a02 - a00, mid = a12 - a10, bot = a22 - a20;
    dx = __fmaf_rn(top + bot, s, s2 * mid);                                  // SymmColumnFilter: fma(S0 + S2, f1, f0 * S1)
    float rowm, rowp;
    if (fma_row) {
        rowm = __fmaf_rn(a02, s, __fmaf_rn(a01, s2, s * a00));  // RowVec: s*L, fma(C, 2s, .), fma(R, s, .)
        rowp = __fmaf_rn(a22, s, __fmaf_rn(a21, s2, s * a20));
    } else {
        rowm = (s * a00 + s2 * a01) + s * a02;
        rowp = (s * a20 + s2 * a21) + s * a22;
    }
    dy = rowp - rowm;
}

// ------------------------------------------------------------------------------------------------ eig map
constexpr int ET_W = 32, ET_H = 16;
__global__ void __launch_bounds__(256) detect_eig(DetArgs A) {
    __shared__ float s_xx[ET_H + 2][ET_W + 2], s_xy[ET_H + 2][ET_W + 2], s_yy[ET_H + 2][ET_W + 2];
    __shared__ unsigned int s_max;
    const int b = blockIdx.z;
    DET_FRAME(A, b);
    const DetRect R = A.rois[b];
    const int tx0 = blockIdx.x * ET_W, ty0 = blockIdx.y * ET_H;
    if (tx0 >= R.w || ty0 >= R.h) return;
    const int tid = threadIdx.x;
    if (tid == 0) s_max = 0;
    for (int e = tid; e < (ET_H + 2) * (ET_W + 2); e += 256) {
        const int cy = e / (ET_W + 2), cx = e - cy * (ET_W + 2);
        // covariance cell (tx0 + cx - 1, ty0 + cy - 1) in ROI coordinates, reflect-101 at the ROI edge (boxFilter on the ROI-sized cov Mat)
        const int rx = refl(tx0 + cx - 1, R.w), ry = refl(ty0 + cy - 1, R.h);
        float dx, dy;
        sobel_at(A, R.x + rx, R.y + ry, rx < (R.w & ~31), dx, dy);
        s_xx[cy][cx] = dx * dx, s_xy[cy][cx] = dx * dy, s_yy[cy][cx] = dy * dy;
    }
    __syncthreads();
    unsigned int lmax = 0;
    for (int e = tid; e < ET_W * ET_H; e += 256) {
        const int py = e / ET_W, px = e - py * ET_W;
        const int x = tx0 + px, y = ty0 + py;
        if (x >= R.w || y >= R.h) continue;
        double sa = 0, sb = 0, sc = 0;  // unnormalised 3x3 box sums accumulated in f64, rounded once to f32
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int i = 0; i < 3; i++) sa += (double) s_xx[py + j][px + i], sb += (double) s_xy[py + j][px + i], sc += (double) s_yy[py + j][px + i];
        const float a = (float) sa * 0.5f, bb = (float) sb, c = (float) sc * 0.5f;
        const float ev = (a + c) - sqrtf((a - c) * (a - c) + bb * bb);
        A.eig[(size_t) b * A.roi_cap + (size_t) y * R.w + x] = ev;
        if (!A.mask || A.mask[(size_t) (R.y + y) * A.pitch + R.x + x]) lmax = max(lmax, float_key(ev));
    }
    for (int o = 16; o > 0; o >>= 1) lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if ((tid & 31) == 0 && lmax) atomicMax(&s_max, lmax);
    __syncthreads();
    if (tid == 0 && s_max) atomicMax(&A.maxkey[b], s_max);
}

// ------------------------------------------------------------------------------------------------ threshold + NMS
__global__ void __launch_bounds__(256) detect_nms(DetArgs A) {
    const int b = blockIdx.z;
    DET_FRAME(A, b);
    const DetRect R = A.rois[b];
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x < 1 || y < 1 || x >= R.w - 1 || y >= R.h - 1) return;
    const unsigned int mk = A.maxkey[b];
    const double maxVal = mk ? (double) key_float(mk) : 0.0;  // minMaxLoc over the mask (0 when the mask is empty)
    const float thr = (float) (maxVal * A.quality);           // threshold(eig, eig, maxVal * qualityLevel, 0, THRESH_TOZERO)
    const float *E = A.eig + (size_t) b * A.roi_cap;
    const float val = E[(size_t) y * R.w + x];
    if (!(val > thr)) return;
    float mx = val;  // dilate 3x3 of the thresholded map: a neighbour above val is necessarily above thr
#pragma unroll
    for (int j = -1; j <= 1; j++)
#pragma unroll
        for (int i = -1; i <= 1; i++) mx = fmaxf(mx, E[(size_t) (y + j) * R.w + x + i]);
    if (val != mx) return;
    if (A.mask && !A.mask[(size_t) (R.y + y) * A.pitch + R.x + x]) return;
    const int slot = atomicAdd(&A.ncand[b], 1);
    if (slot >= DET_MAX_CAND) {
        *A.overflow = 1;
        return;
    }
    // sort key: value descending, then linear address descending (goodFeaturesToTrack's greaterThanPtr); val > 0 here or thr < 0
    A.cand[(size_t) b * DET_MAX_CAND + slot] = ((unsigned long long) float_key(val) << 32) | (unsigned int) (y * R.w + x);
}

// ------------------------------------------------------------------------------------------------ sort + greedy min-distance selection
__global__ void __launch_bounds__(1024) detect_select(DetArgs A) {
    extern __shared__ unsigned long long s_key[];  // DET_MAX_CAND keys
    __shared__ float s_gx[DET_MAX_CELLS][DET_CELL_CAP], s_gy[DET_MAX_CELLS][DET_CELL_CAP];
    __shared__ unsigned char s_gn[DET_MAX_CELLS];
    const int b = blockIdx.x, tid = threadIdx.x;
    DET_FRAME(A, b);
    const DetRect R = A.rois[b];
    int n = min(A.ncand[b], DET_MAX_CAND);
    int npow = 1;
    while (npow < n) npow <<= 1;
    for (int i = tid; i < npow; i += 1024) s_key[i] = i < n ? A.cand[(size_t) b * DET_MAX_CAND + i] : 0ull;
    __syncthreads();
    // bitonic sort, descending
    for (int k = 2; k <= npow; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow; i += 1024) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = s_key[i], c = s_key[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? a < c : a > c) s_key[i] = c, s_key[ixj] = a;
                }
            }
            __syncthreads();
        }
    // greedy acceptance with the min-distance cell grid (sequential by definition; a few hundred candidates are examined)
    const int max_corners = min(A.max_corners[b], A.cap);
    if (tid == 0) {
        int n_out = 0;
        float *out = A.out_xy + (size_t) b * A.cap * 2;
        if (max_corners > 0) {
            if (A.min_distance >= 1) {
                const int cell = (int) rint(A.min_distance);
                const int gw = (R.w + cell - 1) / cell, gh = (R.h + cell - 1) / cell;
                const double md2 = A.min_distance * A.min_distance;
                if (gw * gh > DET_MAX_CELLS) {
                    *A.overflow = 2;
                } else {
                    for (int c = 0; c < gw * gh; c++) s_gn[c] = 0;
                    for (int i = 0; i < n; i++) {
                        const int addr = (int) (s_key[i] & 0xffffffffu);
                        const int y = addr / R.w, x = addr - y * R.w;
                        const int xc = x / cell, yc = y / cell;
                        const int x1 = max(0, xc - 1), y1 = max(0, yc - 1), x2 = min(gw - 1, xc + 1), y2 = min(gh - 1, yc + 1);
                        bool good = true;
                        for (int yy = y1; yy <= y2 && good; yy++)
                            for (int xx = x1; xx <= x2 && good; xx++) {
                                const int c = yy * gw + xx;
                                for (int k = 0; k < s_gn[c]; k++) {
                                    const float dx = (float) x - s_gx[c][k], dy = (float) y - s_gy[c][k];
                                    if ((double) (dx * dx + dy * dy) < md2) {
                                        good = false;
                                        break;
                                    }
                                }
                            }
                        if (!good) continue;
                        const int c = yc * gw + xc;
                        if (s_gn[c] < DET_CELL_CAP) {
                            s_gx[c][s_gn[c]] = (float) x, s_gy[c][s_gn[c]] = (float) y, s_gn[c]++;
                        } else {
                            *A.overflow = 3;  // more than 4 corners in one cell is geometrically impossible for cell = round(minDistance)
                        }
                        out[2 * n_out] = (float) x, out[2 * n_out + 1] = (float) y;
                        if (++n_out == max_corners) break;
                    }
                }
            } else {
                for (int i = 0; i < n && n_out < max_corners; i++) {
                    const int addr = (int) (s_key[i] & 0xffffffffu);
                    out[2 * n_out] = (float) (addr % R.w), out[2 * n_out + 1] = (float) (addr / R.w);
                    n_out++;
                }
            }
        }
        A.out_n[b] = n_out;
    }
}

// ------------------------------------------------------------------------------------------------ cornerSubPix (warp per corner)
__global__ void __launch_bounds__(128) detect_subpix(DetArgs A, int half_win, int max_iter, double eps2) {
    __shared__ float s_sub[4][13 * 13];
    __shared__ float s_mask[11 * 11];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    DET_FRAME(A, b);
    const int p = blockIdx.x * 4 + warp;
    if (half_win != 5) return;  // the only window the reference uses (tracking.cc:623)
    for (int e = threadIdx.x; e < 121; e += 128) {
        const int i = e / 11, j = e - 11 * i;
        const float ty = (float) (i - 5) / 5, tx = (float) (j - 5) / 5;
        s_mask[e] = expf(-ty * ty) * expf(-tx * tx);
    }
    __syncthreads();
    if (p >= A.out_n[b]) return;
    const DetRect R = A.rois[b];
    const uint8_t *roi = A.img + (size_t) R.y * A.pitch + R.x;
    float *xy = A.out_xy + ((size_t) b * A.cap + p) * 2;
    const float cTx = xy[0], cTy = xy[1];
    float cIx = cTx, cIy = cTy;
    float *sub = s_sub[warp];
    int iter = 0;
    float err = 0.f;
    do {
        // getRectSubPix(src, Size(13, 13), cI, ..., CV_32F): bilinear, replicate at the ROI edge
        const float cx = cIx - 6.f, cy = cIy - 6.f;
        const int ipx = __float2int_rd(cx), ipy = __float2int_rd(cy);
        const float a = cx - (float) ipx, bq = cy - (float) ipy;
        const float a11 = (1.f - a) * (1.f - bq), a12 = a * (1.f - bq), a21 = (1.f - a) * bq, a22 = a * bq;
        __syncwarp();
        for (int e = lane; e < 169; e += 32) {
            const int yy = e / 13, xx = e - 13 * yy;
            const int X0 = min(max(ipx + xx, 0), R.w - 1), X1 = min(max(ipx + xx + 1, 0), R.w - 1);
            const int Y0 = min(max(ipy + yy, 0), R.h - 1), Y1 = min(max(ipy + yy + 1, 0), R.h - 1);
            sub[e] = (float) roi[(size_t) Y0 * A.pitch + X0] * a11 + (float) roi[(size_t) Y0 * A.pitch + X1] * a12 + (float) roi[(size_t) Y1 * A.pitch + X0] * a21 +
                     (float) roi[(size_t) Y1 * A.pitch + X1] * a22;
        }
        __syncwarp();
        double sa = 0, sb = 0, sc = 0, sb1 = 0, sb2 = 0;
        for (int e = lane; e < 121; e += 32) {
            const int i = e / 11, j = e - 11 * i;
            const double m = s_mask[e];
            const float *sp = sub + (i + 1) * 13 + (j + 1);
            const double tgx = (double) (sp[1] - sp[-1]), tgy = (double) (sp[13] - sp[-13]);
            const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
            const double px = j - 5, py = i - 5;
            sa += gxx, sb += gxy, sc += gyy;
            sb1 += gxx * px + gxy * py;
            sb2 += gxy * px + gyy * py;
        }
        for (int o = 16; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o), sb += __shfl_xor_sync(0xffffffffu, sb, o), sc += __shfl_xor_sync(0xffffffffu, sc, o);
            sb1 += __shfl_xor_sync(0xffffffffu, sb1, o), sb2 += __shfl_xor_sync(0xffffffffu, sb2, o);
        }
        const double det = sa * sc - sb * sb;
        if (fabs(det) <= 2.220446049250313e-16 * 2.220446049250313e-16) break;
        const double scale = 1.0 / det;
        const float nx = (float) ((double) cIx + sc * scale * sb1 - sb * scale * sb2), ny = (float) ((double) cIy - sb * scale * sb1 + sa * scale * sb2);
        err = (nx - cIx) * (nx - cIx) + (ny - cIy) * (ny - cIy);
        cIx = nx, cIy = ny;
        if (cIx < 0 || cIx >= (float) R.w || cIy < 0 || cIy >= (float) R.h) break;
    } while (++iter < max_iter && (double) err > eps2);
    if (fabsf(cIx - cTx) > (float) half_win || fabsf(cIy - cTy) > (float) half_win) cIx = cTx, cIy = cTy;
    if (lane == 0) xy[0] = cIx, xy[1] = cIy;
}

}  // namespace icg

// ======================================================================================================= C ABI
using namespace icg;

struct icg_detect {
    int W, H, pitch, max_blocks, cap, roi_cap, device;
    cudaStream_t stream;
    bool own_stream;
    uint8_t *d_img, *d_mask;
    DetRect *d_rois;
    int *d_maxc, *d_ncand, *d_out_n, *d_overflow;
    unsigned int *d_maxkey;
    float *d_eig, *d_out_xy;
    unsigned long long *d_cand;
    // pinned staging
    uint8_t *h_stage;
    size_t h_stage_bytes;
};

extern "C" {

int icg_detect_create(icg_detect **out, int width, int height, int max_blocks, int max_corners_per_block, int max_roi_pixels, int device, void *stream) {
    if (!out || width < 16 || height < 16 || max_blocks < 1 || max_corners_per_block < 1 || max_roi_pixels < 9) {
        set_error("icg_detect_create: bad arguments");
        return ICG_EINVAL;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("icg_detect_create: no CUDA device (this library has no CPU fallback)");
        return ICG_ENODEVICE;
    }
    if (device < 0 || device >= ndev) {
        set_error("icg_detect_create: device %d out of range", device);
        return ICG_EINVAL;
    }
    ICG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ICG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("icg_detect_create: device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
        return ICG_ENODEVICE;
    }
    icg_detect *h = new icg_detect();
    h->W = width, h->H = height, h->pitch = (width + 15) & ~15, h->max_blocks = max_blocks, h->cap = max_corners_per_block, h->roi_cap = max_roi_pixels;
    h->device = device;
    h->own_stream = stream == nullptr;
    if (stream)
        h->stream = (cudaStream_t) stream;
    else
        ICG_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    ICG_CUDA(cudaMalloc(&h->d_img, (size_t) h->pitch * height));
    ICG_CUDA(cudaMalloc(&h->d_mask, (size_t) h->pitch * height));
    ICG_CUDA(cudaMalloc(&h->d_rois, sizeof(DetRect) * max_blocks));
    ICG_CUDA(cudaMalloc(&h->d_maxc, sizeof(int) * max_blocks));
    ICG_CUDA(cudaMalloc(&h->d_ncand, sizeof(int) * (2 * max_blocks + 2)));
    h->d_out_n = h->d_ncand + max_blocks;
    h->d_overflow = h->d_ncand + 2 * max_blocks;
    ICG_CUDA(cudaMalloc(&h->d_maxkey, sizeof(unsigned int) * max_blocks));
    ICG_CUDA(cudaMalloc(&h->d_eig, sizeof(float) * (size_t) max_blocks * max_roi_pixels));
    ICG_CUDA(cudaMalloc(&h->d_out_xy, sizeof(float) * 2 * (size_t) max_blocks * max_corners_per_block));
    ICG_CUDA(cudaMalloc(&h->d_cand, sizeof(unsigned long long) * (size_t) max_blocks * DET_MAX_CAND));
    h->h_stage_bytes = sizeof(DetRect) * max_blocks + sizeof(int) * (3 * max_blocks + 4) + sizeof(float) * 2 * (size_t) max_blocks * max_corners_per_block;
    ICG_CUDA(cudaMallocHost(&h->h_stage, h->h_stage_bytes));
    ICG_CUDA(raise_dynamic_smem((const void *) detect_select, (size_t) ((sizeof(unsigned long long) * DET_MAX_CAND))));
    *out = h;
    return ICG_OK;
}

void icg_detect_destroy(icg_detect *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->d_img), cudaFree(h->d_mask), cudaFree(h->d_rois), cudaFree(h->d_maxc), cudaFree(h->d_ncand), cudaFree(h->d_maxkey), cudaFree(h->d_eig);
    cudaFree(h->d_out_xy), cudaFree(h->d_cand);
    cudaFreeHost(h->h_stage);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

static int run_detect(icg_detect *h, const uint8_t *d_img, const uint8_t *d_mask, int n_blocks, const icg_rect *rois, const int32_t *max_corners, double quality,
                      double min_distance, int do_select, int do_subpix, int n_given, float *out_xy, int32_t *out_n, int pitch = 0, size_t frame_stride = 0,
                      int n_frames = 1) {
    // n_frames > 1: `rois` (n_blocks of them) apply to every frame; device block index = frame * n_blocks + roi
    const int rois_per_frame = n_blocks;
    n_blocks *= n_frames;
    DetRect *hr = (DetRect *) h->h_stage;
    int *hm = (int *) (hr + h->max_blocks);
    int *hn = hm + h->max_blocks;          // out_n (also input counts when !do_select)
    int *hov = hn + 2 * h->max_blocks;      // overflow flag
    float *hxy = (float *) (hov + 4);
    int maxw = 0, maxh = 0;
    for (int b = 0; b < n_blocks; b++) {
        const icg_rect &r = rois[b % rois_per_frame];
        if (r.w < 3 || r.h < 3 || r.x < 0 || r.y < 0 || r.x + r.w > h->W || r.y + r.h > h->H || (size_t) r.w * r.h > (size_t) h->roi_cap) {
            set_error("icg_detect: block %d ROI (%d,%d,%d,%d) outside the %dx%d frame or larger than max_roi_pixels", b, r.x, r.y, r.w, r.h, h->W, h->H);
            return ICG_EINVAL;
        }
        hr[b] = DetRect{r.x, r.y, r.w, r.h};
        hm[b] = max_corners ? max_corners[b] : h->cap;
        maxw = std::max(maxw, r.w), maxh = std::max(maxh, r.h);
    }
    cudaStream_t s = h->stream;
    ICG_CUDA(cudaMemcpyAsync(h->d_rois, hr, sizeof(DetRect) * n_blocks, cudaMemcpyHostToDevice, s));
    ICG_CUDA(cudaMemcpyAsync(h->d_maxc, hm, sizeof(int) * n_blocks, cudaMemcpyHostToDevice, s));
    ICG_CUDA(cudaMemsetAsync(h->d_ncand, 0, sizeof(int) * (2 * h->max_blocks + 2), s));
    ICG_CUDA(cudaMemsetAsync(h->d_maxkey, 0, sizeof(unsigned int) * h->max_blocks, s));
    DetArgs A;
    A.img = d_img, A.mask = d_mask, A.W = h->W, A.H = h->H, A.pitch = pitch ? pitch : h->pitch, A.n_blocks = n_blocks, A.cap = h->cap, A.roi_cap = h->roi_cap;
    A.img_stride = frame_stride, A.bpf = rois_per_frame;
    A.rois = h->d_rois, A.max_corners = h->d_maxc, A.eig = h->d_eig, A.maxkey = h->d_maxkey, A.cand = h->d_cand, A.ncand = h->d_ncand;
    A.out_xy = h->d_out_xy, A.out_n = h->d_out_n, A.overflow = h->d_overflow, A.quality = quality, A.min_distance = min_distance;
    if (do_select) {
        detect_eig<<<dim3((maxw + ET_W - 1) / ET_W, (maxh + ET_H - 1) / ET_H, n_blocks), 256, 0, s>>>(A);
        detect_nms<<<dim3((maxw + 31) / 32, (maxh + 7) / 8, n_blocks), 256, 0, s>>>(A);
        detect_select<<<n_blocks, 1024, sizeof(unsigned long long) * DET_MAX_CAND, s>>>(A);
        count_launch(3);
    } else {
        // corners supplied by the caller (cv::cornerSubPix drop-in): out_xy holds n_given points of block 0
        memcpy(hxy, out_xy, sizeof(float) * 2 * n_given);
        hn[0] = n_given;
        ICG_CUDA(cudaMemcpyAsync(h->d_out_xy, hxy, sizeof(float) * 2 * n_given, cudaMemcpyHostToDevice, s));
        ICG_CUDA(cudaMemcpyAsync(h->d_out_n, hn, sizeof(int), cudaMemcpyHostToDevice, s));
    }
    if (do_subpix) {
        // cornerSubPix(win (5,5), zeroZone (-1,-1), COUNT+EPS 20 / 0.01): eps squared (IG/tracking/tracking.cc:623-625,651)
        detect_subpix<<<dim3((h->cap + 3) / 4, n_blocks), 128, 0, s>>>(A, 5, 20, 0.01 * 0.01);
        count_launch();
    }
    ICG_CHECK_LAUNCH();
    ICG_CUDA(cudaMemcpyAsync(hn, h->d_out_n, sizeof(int) * n_blocks, cudaMemcpyDeviceToHost, s));
    ICG_CUDA(cudaMemcpyAsync(hov, h->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, s));
    ICG_CUDA(cudaMemcpyAsync(hxy, h->d_out_xy, sizeof(float) * 2 * (size_t) n_blocks * h->cap, cudaMemcpyDeviceToHost, s));
    ICG_CUDA(cudaStreamSynchronize(s));
    if (*hov) {
        set_error("icg_detect: internal capacity exceeded (code %d: 1 = more than %d NMS candidates in a block, 2 = min-distance grid too fine, 3 = cell overflow)", *hov,
                  DET_MAX_CAND);
        return ICG_EUNSUPPORTED;
    }
    for (int b = 0; b < n_blocks; b++) {
        if (out_n) out_n[b] = hn[b];
        memcpy(out_xy + (size_t) b * h->cap * 2, hxy + (size_t) b * h->cap * 2, sizeof(float) * 2 * hn[b]);
    }
    return ICG_OK;
}

int icg_detect_blocks(icg_detect *h, const uint8_t *img, const uint8_t *mask, int stride, int n_blocks, const icg_rect *rois, const int32_t *max_corners,
                      double quality, double min_distance, int do_subpix, float *out_xy, int32_t *out_n) {
    if (!h || !img || !rois || !out_xy || !out_n || n_blocks < 1 || n_blocks > h->max_blocks || stride < h->W) {
        set_error("icg_detect_blocks: bad arguments");
        return ICG_EINVAL;
    }
    ICG_CUDA(cudaSetDevice(h->device));
    ICG_CUDA(cudaMemcpy2DAsync(h->d_img, h->pitch, img, stride, h->W, h->H, cudaMemcpyHostToDevice, h->stream));
    if (mask) ICG_CUDA(cudaMemcpy2DAsync(h->d_mask, h->pitch, mask, stride, h->W, h->H, cudaMemcpyHostToDevice, h->stream));
    return run_detect(h, h->d_img, mask ? h->d_mask : nullptr, n_blocks, rois, max_corners, quality, min_distance, 1, do_subpix, 0, out_xy, out_n);
}

int icg_detect_blocks_dev(icg_detect *h, int n_frames, const uint8_t *dev_img, int pitch, size_t frame_stride, const uint8_t *dev_mask, int n_blocks,
                          const icg_rect *rois, const int32_t *max_corners, double quality, double min_distance, int do_subpix, float *out_xy, int32_t *out_n) {
    if (!h || !dev_img || !rois || !out_xy || !out_n || n_frames < 1 || n_blocks < 1 || (long long) n_frames * n_blocks > h->max_blocks || pitch < h->W) {
        set_error("icg_detect_blocks_dev: bad arguments (n_frames * n_blocks must be <= max_blocks of the handle)");
        return ICG_EINVAL;
    }
    ICG_CUDA(cudaSetDevice(h->device));
    return run_detect(h, dev_img, dev_mask, n_blocks, rois, max_corners, quality, min_distance, 1, do_subpix, 0, out_xy, out_n, pitch, frame_stride, n_frames);
}

int icg_corner_subpix(icg_detect *h, const uint8_t *img, int stride, float *corners_xy, int n) {
    if (!h || !img || !corners_xy || n < 0 || n > h->cap || stride < h->W) {
        set_error("icg_corner_subpix: bad arguments (n must be <= max_corners_per_block of the handle)");
        return ICG_EINVAL;
    }
    if (n == 0) return ICG_OK;
    ICG_CUDA(cudaSetDevice(h->device));
    ICG_CUDA(cudaMemcpy2DAsync(h->d_img, h->pitch, img, stride, h->W, h->H, cudaMemcpyHostToDevice, h->stream));
    icg_rect full = {0, 0, h->W, h->H};
    int32_t cnt = 0;
    if ((size_t) h->W * h->H > (size_t) h->roi_cap) {
        // the eig plane is not needed for sub-pixel refinement; only the ROI geometry is
    }
    // run_detect validates roi_cap against w*h; sub-pixel refinement does not touch the eig plane, so bypass that check
    int saved = h->roi_cap;
    h->roi_cap = h->W * h->H;
    int rc = run_detect(h, h->d_img, nullptr, 1, &full, nullptr, 0.0, 0.0, 0, 1, n, corners_xy, &cnt);
    h->roi_cap = saved;
    return rc;
}

}  // extern "C"
