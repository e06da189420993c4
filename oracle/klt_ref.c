/*
 * oracle/klt_ref.c -- CPU restatement (TEST INFRASTRUCTURE, never shipped, never the product path)
 * of the OpenCV routines that IC-GVINS' visual front end bottoms out in:
 *
 *   cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(21,21), 3,
 *                            TermCriteria(COUNT+EPS, 30, 0.01), OPTFLOW_USE_INITIAL_FLOW)
 *       called at ic_gvins/ic_gvins/tracking/tracking.cc:385,390,487,493
 *   forward/backward gate + border gate            tracking.cc:396-403, 499-506, 841-849
 *
 * The arithmetic itself is NOT under /root/reference: it lives in the un-vendored system dependency
 * OpenCV (module video, lkpyramid.cpp; module imgproc pyrDown), version unpinned by the reference
 * (README.md:56-58) and pinned in this container to opencv-python-headless 4.13.0.  This file restates the
 * published algorithm in plain C.  It is pinned by tests/test_oracle_klt.py
 * against golden vectors generated from cv2 4.13.0 with tests/golden/make_klt_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static inline int reflect101(int p, int len) {
    /* cv::borderInterpolate(p, len, BORDER_REFLECT_101) */
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* cv::pyrDown for CV_8UC1: separable [1 4 6 4 1], reflect-101, (sum+128)>>8. */
void icgo_pyr_down(const uint8_t *src, int W, int H, int sstride, uint8_t *dst, int dstride) {
    int Wd = (W + 1) / 2, Hd = (H + 1) / 2;
    static const int k[5] = {1, 4, 6, 4, 1};
    for (int y = 0; y < Hd; y++) {
        for (int x = 0; x < Wd; x++) {
            int sum = 0;
            for (int i = -2; i <= 2; i++) {
                int sy          = reflect101(2 * y + i, H);
                const uint8_t *r = src + (size_t) sy * sstride;
                int rs          = 0;
                for (int j = -2; j <= 2; j++) rs += k[j + 2] * r[reflect101(2 * x + j, W)];
                sum += k[i + 2] * rs;
            }
            dst[(size_t) y * dstride + x] = (uint8_t) ((sum + 128) >> 8);
        }
    }
}

/* Scharr derivative used inside LK: int16 interleaved (Ix,Iy), no scaling, reflect-101 source. */
void icgo_scharr(const uint8_t *src, int W, int H, int sstride, int16_t *d /* H*W*2 */) {
    for (int y = 0; y < H; y++) {
        const uint8_t *r0 = src + (size_t) reflect101(y - 1, H) * sstride;
        const uint8_t *r1 = src + (size_t) y * sstride;
        const uint8_t *r2 = src + (size_t) reflect101(y + 1, H) * sstride;
        for (int x = 0; x < W; x++) {
            int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
            /* t0 = vertical smooth (3,10,3); t1 = vertical diff */
            int t0m = 3 * (r0[xm] + r2[xm]) + 10 * r1[xm];
            int t0p = 3 * (r0[xp] + r2[xp]) + 10 * r1[xp];
            int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
            d[((size_t) y * W + x) * 2 + 0] = (int16_t) (t0p - t0m);
            d[((size_t) y * W + x) * 2 + 1] = (int16_t) (3 * (t1m + t1p) + 10 * t1c);
        }
    }
}

static inline int cv_round_f(float v) {
    /* cvRound(float) = cvtss2si: round-half-to-even */
    return (int) lrintf(v);
}

typedef struct {
    int W, H;
    uint8_t *img;   /* padded by win, reflect-101 */
    int16_t *deriv; /* padded by win, zeros */
    int pstride;    /* padded width */
} icgo_level;

static void make_level(icgo_level *L, const uint8_t *src, int W, int H, int sstride, int win, int with_deriv) {
    int PW = W + 2 * win, PH = H + 2 * win;
    L->W       = W;
    L->H       = H;
    L->pstride = PW;
    L->img     = (uint8_t *) malloc((size_t) PW * PH);
    for (int y = 0; y < PH; y++) {
        const uint8_t *r = src + (size_t) reflect101(y - win, H) * sstride;
        for (int x = 0; x < PW; x++) L->img[(size_t) y * PW + x] = r[reflect101(x - win, W)];
    }
    L->deriv = NULL;
    if (with_deriv) {
        int16_t *d = (int16_t *) malloc((size_t) W * H * 2 * sizeof(int16_t));
        icgo_scharr(src, W, H, sstride, d);
        L->deriv = (int16_t *) calloc((size_t) PW * PH * 2, sizeof(int16_t));
        for (int y = 0; y < H; y++)
            memcpy(L->deriv + ((size_t) (y + win) * PW + win) * 2, d + (size_t) y * W * 2, (size_t) W * 2 * sizeof(int16_t));
        free(d);
    }
}

#define ICGO_USE_INITIAL_FLOW 4

/*
 * calcOpticalFlowPyrLK restatement (SURVEY A.3 + the level-0 "err" epilogue of lkpyramid.cpp that
 * clears status when the final window origin leaves [-win, cols) x [-win, rows); the reference passes
 * an err vector at tracking.cc:385, so that epilogue is live).
 * next_xy: in = initial flow (if flags & USE_INITIAL_FLOW), out = tracked positions.
 * Returns the number of pyramid levels used - 1 (effective maxLevel).
 */
int icgo_calc_optical_flow_pyr_lk(const uint8_t *prev, const uint8_t *next, int W, int H, int stride,
                                  const float *prev_xy, float *next_xy, uint8_t *status, float *err, int n, int win,
                                  int max_level, int max_iter, double eps, int flags, double min_eig_thr) {
    icgo_level P[8], N[8];
    uint8_t *pbuf[8], *nbuf[8];
    int nl = 0;
    if (max_level > 7) max_level = 7;
    /* buildOpticalFlowPyramid: stop when a level is not larger than the window */
    {
        int w = W, h = H;
        const uint8_t *ps = prev, *ns = next;
        int pstr = stride, nstr = stride;
        for (int l = 0; l <= max_level; l++) {
            if (l > 0) {
                int w2 = (w + 1) / 2, h2 = (h + 1) / 2;
                if (w2 <= win || h2 <= win) break;
                pbuf[l] = (uint8_t *) malloc((size_t) w2 * h2);
                nbuf[l] = (uint8_t *) malloc((size_t) w2 * h2);
                icgo_pyr_down(ps, w, h, pstr, pbuf[l], w2);
                icgo_pyr_down(ns, w, h, nstr, nbuf[l], w2);
                ps = pbuf[l];
                ns = nbuf[l];
                pstr = nstr = w2;
                w = w2;
                h = h2;
            } else {
                pbuf[0] = nbuf[0] = NULL;
            }
            make_level(&P[l], ps, w, h, pstr, win, 1);
            make_level(&N[l], ns, w, h, nstr, win, 0);
            nl++;
        }
    }
    max_level = nl - 1;

    if (max_iter < 0) max_iter = 0;
    if (max_iter > 100) max_iter = 100;
    if (eps < 0) eps = 0;
    if (eps > 10) eps = 10;
    double eps2 = eps * eps;

    for (int i = 0; i < n; i++) {
        status[i] = 1;
        if (err) err[i] = 0;
    }

    const float half     = (float) ((win - 1) * 0.5f);
    const float FLT_SC   = 1.f / (1 << 20);
    int16_t *Iwin        = (int16_t *) malloc((size_t) win * win * sizeof(int16_t));
    int16_t *dIwin       = (int16_t *) malloc((size_t) win * win * 2 * sizeof(int16_t));
    float *cur           = (float *) malloc((size_t) n * 2 * sizeof(float));

    for (int level = max_level; level >= 0; level--) {
        const icgo_level *LI = &P[level], *LJ = &N[level];
        const int cols = LI->W, rows = LI->H, ps = LI->pstride;
        const float scale = (float) (1. / (1 << level));
        for (int p = 0; p < n; p++) {
            float px = prev_xy[2 * p] * scale, py = prev_xy[2 * p + 1] * scale;
            float nx, ny;
            if (level == max_level) {
                if (flags & ICGO_USE_INITIAL_FLOW) {
                    nx = next_xy[2 * p] * scale;
                    ny = next_xy[2 * p + 1] * scale;
                } else {
                    nx = px;
                    ny = py;
                }
            } else {
                nx = cur[2 * p] * 2.f;
                ny = cur[2 * p + 1] * 2.f;
            }
            cur[2 * p]     = nx;
            cur[2 * p + 1] = ny;

            px -= half;
            py -= half;
            int ipx = (int) floorf(px), ipy = (int) floorf(py);
            if (ipx < -win || ipx >= cols || ipy < -win || ipy >= rows) {
                if (level == 0) {
                    status[p] = 0;
                    if (err) err[p] = 0;
                }
                continue;
            }
            float a = px - ipx, b = py - ipy;
            int iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << 14));
            int iw01 = cv_round_f(a * (1.f - b) * (1 << 14));
            int iw10 = cv_round_f((1.f - a) * b * (1 << 14));
            int iw11 = (1 << 14) - iw00 - iw01 - iw10;

            float fA11 = 0, fA12 = 0, fA22 = 0;
            for (int y = 0; y < win; y++) {
                const uint8_t *src  = LI->img + (size_t) (y + ipy + win) * ps + ipx + win;
                const int16_t *dsrc = LI->deriv + ((size_t) (y + ipy + win) * ps + ipx + win) * 2;
                for (int x = 0; x < win; x++, dsrc += 2) {
                    int ival  = (src[x] * iw00 + src[x + 1] * iw01 + src[x + ps] * iw10 + src[x + ps + 1] * iw11 + (1 << 8)) >> 9;
                    int ixval = (dsrc[0] * iw00 + dsrc[2] * iw01 + dsrc[2 * ps] * iw10 + dsrc[2 * ps + 2] * iw11 + (1 << 13)) >> 14;
                    int iyval = (dsrc[1] * iw00 + dsrc[3] * iw01 + dsrc[2 * ps + 1] * iw10 + dsrc[2 * ps + 3] * iw11 + (1 << 13)) >> 14;
                    Iwin[y * win + x]            = (int16_t) ival;
                    dIwin[(y * win + x) * 2]     = (int16_t) ixval;
                    dIwin[(y * win + x) * 2 + 1] = (int16_t) iyval;
                    fA11 += (float) (ixval * ixval);
                    fA12 += (float) (ixval * iyval);
                    fA22 += (float) (iyval * iyval);
                }
            }
            float A11 = fA11 * FLT_SC, A12 = fA12 * FLT_SC, A22 = fA22 * FLT_SC;
            float D      = A11 * A22 - A12 * A12;
            float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
            if (minEig < min_eig_thr || D < FLT_EPSILON) {
                if (level == 0) status[p] = 0;
                continue;
            }
            D = 1.f / D;

            nx -= half;
            ny -= half;
            float pdx = 0, pdy = 0;
            for (int j = 0; j < max_iter; j++) {
                int inx = (int) floorf(nx), iny = (int) floorf(ny);
                if (inx < -win || inx >= cols || iny < -win || iny >= rows) {
                    if (level == 0) status[p] = 0;
                    break;
                }
                a    = nx - inx;
                b    = ny - iny;
                iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << 14));
                iw01 = cv_round_f(a * (1.f - b) * (1 << 14));
                iw10 = cv_round_f((1.f - a) * b * (1 << 14));
                iw11 = (1 << 14) - iw00 - iw01 - iw10;
                float fb1 = 0, fb2 = 0;
                for (int y = 0; y < win; y++) {
                    const uint8_t *J = LJ->img + (size_t) (y + iny + win) * ps + inx + win;
                    for (int x = 0; x < win; x++) {
                        int diff = ((J[x] * iw00 + J[x + 1] * iw01 + J[x + ps] * iw10 + J[x + ps + 1] * iw11 + (1 << 8)) >> 9) -
                                   Iwin[y * win + x];
                        fb1 += (float) (diff * dIwin[(y * win + x) * 2]);
                        fb2 += (float) (diff * dIwin[(y * win + x) * 2 + 1]);
                    }
                }
                float b1 = fb1 * FLT_SC, b2 = fb2 * FLT_SC;
                float dx = (float) ((A12 * b2 - A22 * b1) * D);
                float dy = (float) ((A12 * b1 - A11 * b2) * D);
                nx += dx;
                ny += dy;
                cur[2 * p]     = nx + half;
                cur[2 * p + 1] = ny + half;
                if ((double) dx * dx + (double) dy * dy <= eps2) break;
                if (j > 0 && fabs(dx + pdx) < 0.01 && fabs(dy + pdy) < 0.01) {
                    cur[2 * p] -= dx * 0.5f;
                    cur[2 * p + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx;
                pdy = dy;
            }

            if (status[p] && err && level == 0) {
                float fx = cur[2 * p] - half, fy = cur[2 * p + 1] - half;
                int inx = (int) floorf(fx), iny = (int) floorf(fy);
                if (inx < -win || inx >= cols || iny < -win || iny >= rows) {
                    status[p] = 0;
                    continue;
                }
                float aa = fx - inx, bb = fy - iny;
                iw00 = cv_round_f((1.f - aa) * (1.f - bb) * (1 << 14));
                iw01 = cv_round_f(aa * (1.f - bb) * (1 << 14));
                iw10 = cv_round_f((1.f - aa) * bb * (1 << 14));
                iw11 = (1 << 14) - iw00 - iw01 - iw10;
                float errval = 0.f;
                for (int y = 0; y < win; y++) {
                    const uint8_t *J = LJ->img + (size_t) (y + iny + win) * ps + inx + win;
                    for (int x = 0; x < win; x++) {
                        int diff = ((J[x] * iw00 + J[x + 1] * iw01 + J[x + ps] * iw10 + J[x + ps + 1] * iw11 + (1 << 8)) >> 9) -
                                   Iwin[y * win + x];
                        errval += (float) abs(diff);
                    }
                }
                err[p] = errval * 1.f / (32 * win * win);
            }
        }
    }
    for (int p = 0; p < n; p++) {
        next_xy[2 * p]     = cur[2 * p];
        next_xy[2 * p + 1] = cur[2 * p + 1];
    }

    for (int l = 0; l < nl; l++) {
        free(P[l].img);
        free(P[l].deriv);
        free(N[l].img);
        if (l > 0) {
            free(pbuf[l]);
            free(nbuf[l]);
        }
    }
    free(Iwin);
    free(dIwin);
    free(cur);
    return max_level;
}

/*
 * Forward + backward LK with the reference's gate (tracking.cc:385-403):
 *   status = st_fwd && st_bwd && !isOnBorder(fwd) && dist(bwd, prev) < 0.5
 * isOnBorder: x < 5 || y < 5 || x > W-5 || y > H-5 (tracking.cc:847-849), distance in double (tracking.cc:841-845).
 */
void icgo_track_fb(const uint8_t *prev, const uint8_t *next, int W, int H, int stride, const float *prev_xy,
                   float *next_xy /* in: init, out: fwd */, float *back_xy /* out */, uint8_t *status, int n, int win,
                   int max_level, int max_iter, double eps, double fb_thr, double border) {
    uint8_t *st2 = (uint8_t *) malloc((size_t) n);
    float *err   = (float *) malloc((size_t) n * sizeof(float));
    icgo_calc_optical_flow_pyr_lk(prev, next, W, H, stride, prev_xy, next_xy, status, err, n, win, max_level, max_iter,
                                  eps, ICGO_USE_INITIAL_FLOW, 1e-4);
    memcpy(back_xy, prev_xy, (size_t) n * 2 * sizeof(float));
    icgo_calc_optical_flow_pyr_lk(next, prev, W, H, stride, next_xy, back_xy, st2, err, n, win, max_level, max_iter,
                                  eps, ICGO_USE_INITIAL_FLOW, 1e-4);
    for (int k = 0; k < n; k++) {
        float x = next_xy[2 * k], y = next_xy[2 * k + 1];
        int onb   = x < 5.0 || y < 5.0 || x > (W - border) || y > (H - border);
        double dx = (double) (back_xy[2 * k] - prev_xy[2 * k]); /* float subtraction, then widened (tracking.cc:842) */
        double dy = (double) (back_xy[2 * k + 1] - prev_xy[2 * k + 1]);
        double d  = sqrt(dx * dx + dy * dy);
        status[k] = (status[k] && st2[k] && !onb && d < fb_thr) ? 1 : 0;
    }
    free(st2);
    free(err);
}
