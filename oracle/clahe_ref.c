/* oracle/clahe_ref.c -- CPU restatement of cv::CLAHE::apply for 8-bit images (TEST INFRASTRUCTURE ONLY).
 *
 * The reference equalises every frame with `clahe_ = cv::createCLAHE(3.0, cv::Size(21, 21))`
 * (ic_gvins/ic_gvins/tracking/tracking.cc:62) and `clahe_->apply(frame_cur_->image(), frame_cur_->image())` (:141).
 * OpenCV is an un-vendored dependency of the reference (ic_gvins/CMakeLists.txt:24); the algorithm restated here is OpenCV's
 * modules/imgproc/src/clahe.cpp (CLAHE_CalcLut_Body + CLAHE_Interpolation_Body), pinned bit-exactly against cv2 4.13.0 by
 * tests/golden/clahe_golden.npz (tests/test_oracle_clahe.py).  Compile with -ffp-contract=off (the float sequence matters).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}
static uint8_t sat_u8_from_float(float v) {
    long r = lrintf(v); /* cvRound: round half to even (default rounding mode) */
    return (uint8_t) (r < 0 ? 0 : r > 255 ? 255 : r);
}

/* per-tile look-up tables: lut[(ty * tiles_x + tx) * 256 + v] */
void icgo_clahe_lut(const uint8_t *src, int W, int H, int stride, double clip_limit_, int tiles_x, int tiles_y, uint8_t *lut) {
    /* CLAHE_Impl::apply: pad to a multiple of the grid with BORDER_REFLECT_101 (right / bottom) */
    int Wp = W, Hp = H;
    if (W % tiles_x != 0 || H % tiles_y != 0) {
        Wp = W + (tiles_x - (W % tiles_x));
        Hp = H + (tiles_y - (H % tiles_y));
    }
    const int tw = Wp / tiles_x, th = Hp / tiles_y, area = tw * th;
    const float lut_scale = (float) 255 / area;
    int clip = 0;
    if (clip_limit_ > 0.0) {
        clip = (int) (clip_limit_ * area / 256);
        if (clip < 1) clip = 1;
    }
    for (int ty = 0; ty < tiles_y; ty++)
        for (int tx = 0; tx < tiles_x; tx++) {
            int hist[256];
            memset(hist, 0, sizeof(hist));
            for (int y = ty * th; y < (ty + 1) * th; y++) {
                const uint8_t *row = src + (size_t) reflect101(y, H) * stride;
                for (int x = tx * tw; x < (tx + 1) * tw; x++) hist[row[reflect101(x, W)]]++;
            }
            if (clip > 0) {
                int clipped = 0;
                for (int i = 0; i < 256; i++)
                    if (hist[i] > clip) {
                        clipped += hist[i] - clip;
                        hist[i] = clip;
                    }
                const int batch = clipped / 256;
                int residual = clipped - batch * 256;
                for (int i = 0; i < 256; i++) hist[i] += batch;
                if (residual != 0) {
                    int step = 256 / residual;
                    if (step < 1) step = 1;
                    for (int i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++;
                }
            }
            uint8_t *l = lut + (size_t) (ty * tiles_x + tx) * 256;
            int sum = 0;
            for (int i = 0; i < 256; i++) {
                sum += hist[i];
                l[i] = sat_u8_from_float(sum * lut_scale);
            }
        }
}

/* cv::CLAHE::apply(src, dst) for CV_8UC1; dst may alias src */
void icgo_clahe_apply(const uint8_t *src, int W, int H, int stride, double clip_limit, int tiles_x, int tiles_y, uint8_t *dst, int dst_stride) {
    uint8_t *lut = (uint8_t *) malloc((size_t) tiles_x * tiles_y * 256);
    icgo_clahe_lut(src, W, H, stride, clip_limit, tiles_x, tiles_y, lut);
    int Wp = W, Hp = H;
    if (W % tiles_x != 0 || H % tiles_y != 0) {
        Wp = W + (tiles_x - (W % tiles_x));
        Hp = H + (tiles_y - (H % tiles_y));
    }
    const int tw = Wp / tiles_x, th = Hp / tiles_y;
    const float inv_tw = 1.0f / tw, inv_th = 1.0f / th;
    for (int y = 0; y < H; y++) {
        const float tyf = y * inv_th - 0.5f;
        int ty1 = (int) floorf(tyf), ty2 = ty1 + 1;
        const float ya = tyf - ty1, ya1 = 1.0f - ya;
        if (ty1 < 0) ty1 = 0;
        if (ty2 > tiles_y - 1) ty2 = tiles_y - 1;
        const uint8_t *p1 = lut + (size_t) ty1 * tiles_x * 256, *p2 = lut + (size_t) ty2 * tiles_x * 256;
        for (int x = 0; x < W; x++) {
            const float txf = x * inv_tw - 0.5f;
            int tx1 = (int) floorf(txf), tx2 = tx1 + 1;
            const float xa = txf - tx1, xa1 = 1.0f - xa;
            if (tx1 < 0) tx1 = 0;
            if (tx2 > tiles_x - 1) tx2 = tiles_x - 1;
            const int v = src[(size_t) y * stride + x];
            const int i1 = tx1 * 256 + v, i2 = tx2 * 256 + v;
            const float res = (p1[i1] * xa1 + p1[i2] * xa) * ya1 + (p2[i1] * xa1 + p2[i2] * xa) * ya;
            dst[(size_t) y * dst_stride + x] = sat_u8_from_float(res);
        }
    }
    free(lut);
}
