// fundamental.cu -- SURVEY.md 8f rank 3 (first half): the fundamental-matrix outlier gate of the front end.  HOST code in this round:
// <= 300 point pairs per frame and a sequentially adaptive hypothesis loop (the iteration bound shrinks as better models are found, the
// subsets come from one cv::RNG stream), so the reference-exact result is defined by a serial order.  (The B200 plan -- all subsets of the
// stream drawn up front, every hypothesis scored in parallel, then the serial acceptance replayed -- is in DESIGN.md section 7.)
//
// Replaces cv::findFundamentalMat(pts_new_undis, pts_cur_undis, cv::FM_RANSAC, reprojection_error_std_, 0.99, status)
// (IG/tracking/tracking.cc:547; only `status` is consumed, :549-553).  OpenCV is an un-vendored dependency of the reference; its
// algorithm (calib3d fundam.cpp / ptsetreg.cpp) is restated here and pinned against cv2 4.13.0 by tests/golden/fundamental_golden.npz:
// identical inlier masks.
#include <float.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace icg;

namespace {

struct CvRng {  // cv::RNG: multiply-with-carry, CV_RNG_COEFF = 4164903690
    uint64_t state;
    explicit CvRng(uint64_t s = 0xffffffffffffffffull) : state(s ? s : 0xffffffffffffffffull) {}
    unsigned next() {
        state = (uint64_t) (unsigned) state * 4164903690u + (unsigned) (state >> 32);
        return (unsigned) state;
    }
    int uniform(int a, int b) { return a == b ? a : (int) (next() % (unsigned) (b - a) + a); }
};

// haveCollinearPoints (fundam.cpp): the LAST of `count` points against every pair of earlier ones
bool collinear_last(const float *p /* count x 2 */, int count) {
    const int i = count - 1;
    for (int j = 0; j < i; j++) {
        const double dx1 = p[2 * j] - p[2 * i], dy1 = p[2 * j + 1] - p[2 * i + 1];
        for (int k = 0; k < j; k++) {
            const double dx2 = p[2 * k] - p[2 * i], dy2 = p[2 * k + 1] - p[2 * i + 1];
            if (fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
        }
    }
    return false;
}

// cv::solveCubic: c[0] x^3 + c[1] x^2 + c[2] x + c[3] = 0, real roots in OpenCV's order
int solve_cubic(const double c[4], double r[3]) {
    double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0) return a3 == 0 ? -1 : 0;
            r[0] = -a3 / a2;
            return 1;
        }
        double d = a2 * a2 - 4 * a1 * a3;
        if (d >= 0) {
            d = sqrt(d);
            const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
            if (fabs(q1) > fabs(q2)) {
                r[0] = q1 / a1;
                r[1] = a3 / q1;
            } else {
                r[0] = q2 / a1;
                r[1] = a3 / q2;
            }
            return d > 0 ? 2 : 1;
        }
        return 0;
    }
    a0 = 1. / a0;
    a1 *= a0, a2 *= a0, a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9), R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54), Qcubed = Q * Q * Q;
    double d = Qcubed - R * R;
    if (d > 0) {
        const double theta = acos(R / sqrt(Qcubed)), sqrtQ = sqrt(Q), t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
        r[0] = t0 * cos(t1) - t2;
        r[1] = t0 * cos(t1 + (2. * M_PI / 3)) - t2;
        r[2] = t0 * cos(t1 + (4. * M_PI / 3)) - t2;
        return 3;
    }
    if (d == 0) {
        if (R >= 0) {
            r[0] = -2 * pow(R, 1. / 3) - a1 / 3;
            r[1] = pow(R, 1. / 3) - a1 / 3;
        } else {
            r[0] = 2 * pow(-R, 1. / 3) - a1 / 3;
            r[1] = -pow(-R, 1. / 3) - a1 / 3;
        }
        return 2;
    }
    d = sqrt(-d);
    double e = pow(d + fabs(R), 1. / 3);
    if (R > 0) e = -e;
    r[0] = (e + Q / e) - a1 * (1. / 3);
    return 1;
}

// right null space of the 7 x 9 system: one-sided (Hestenes) Jacobi on the columns of A (rows padded to 9 with zeros); the two columns of
// V whose A-images have the smallest norms span it.  Returned as f1 = v7, f2 = v8 in the descending-singular-value order of cv::SVDecomp.
void null_space_7x9(const double a[7 * 9], double f1[9], double f2[9]) {
    double G[9][7], V[9][9];  // column-major: G[c] = column c of A
    for (int c = 0; c < 9; c++) {
        for (int r = 0; r < 7; r++) G[c][r] = a[r * 9 + c];
        for (int r = 0; r < 9; r++) V[c][r] = r == c ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 60; sweep++) {
        bool rotated = false;
        for (int p = 0; p < 8; p++)
            for (int q = p + 1; q < 9; q++) {
                double al = 0, be = 0, ga = 0;
                for (int r = 0; r < 7; r++) al += G[p][r] * G[p][r], be += G[q][r] * G[q][r], ga += G[p][r] * G[q][r];
                if (ga == 0.0 || fabs(ga) <= 1e-16 * sqrt(al * be)) continue;
                rotated = true;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int r = 0; r < 7; r++) {
                    const double x = G[p][r], y = G[q][r];
                    G[p][r] = cs * x - sn * y, G[q][r] = sn * x + cs * y;
                }
                for (int r = 0; r < 9; r++) {
                    const double x = V[p][r], y = V[q][r];
                    V[p][r] = cs * x - sn * y, V[q][r] = sn * x + cs * y;
                }
            }
        if (!rotated) break;
    }
    int order[9];
    double nrm[9];
    for (int c = 0; c < 9; c++) {
        order[c] = c, nrm[c] = 0;
        for (int r = 0; r < 7; r++) nrm[c] += G[c][r] * G[c][r];
    }
    std::stable_sort(order, order + 9, [&](int x, int y) { return nrm[x] > nrm[y]; });
    memcpy(f1, V[order[7]], sizeof(double) * 9);
    memcpy(f2, V[order[8]], sizeof(double) * 9);
}

// run7Point (fundam.cpp): up to three fundamental matrices through seven correspondences
int run_7point(const float *m1, const float *m2, double F[3][9]) {
    double a[7 * 9], f1[9], f2[9], c[4], r[3] = {0, 0, 0};
    for (int i = 0; i < 7; i++) {
        const double x0 = m1[2 * i], y0 = m1[2 * i + 1], x1 = m2[2 * i], y1 = m2[2 * i + 1];
        double *row = a + i * 9;
        row[0] = x1 * x0, row[1] = x1 * y0, row[2] = x1, row[3] = y1 * x0, row[4] = y1 * y0, row[5] = y1, row[6] = x0, row[7] = y0, row[8] = 1;
    }
    null_space_7x9(a, f1, f2);
    for (int i = 0; i < 9; i++) f1[i] -= f2[i];
    double t0 = f2[4] * f2[8] - f2[5] * f2[7], t1 = f2[3] * f2[8] - f2[5] * f2[6], t2 = f2[3] * f2[7] - f2[4] * f2[6];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
           f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
           f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7], t1 = f1[3] * f1[8] - f1[5] * f1[6], t2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
           f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
           f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    const int n = solve_cubic(c, r);
    if (n < 1 || n > 3) return n;
    for (int k = 0; k < n; k++) {
        double lambda = r[k], mu = 1.;
        const double s = f1[8] * r[k] + f2[8];
        if (fabs(s) > DBL_EPSILON) {  // normalise so that F(3,3) == 1
            mu = 1. / s;
            lambda *= mu;
            F[k][8] = 1.;
        } else {
            F[k][8] = 0.;
        }
        for (int i = 0; i < 8; i++) F[k][i] = f1[i] * lambda + f2[i] * mu;
    }
    return n;
}

// FMEstimatorCallback::computeError + findInliers (ptsetreg.cpp): err = max(d1^2 s1, d2^2 s2) rounded to float, inlier <=> err <= thresh^2
int find_inliers(const float *m1, const float *m2, int count, const double *F, double thresh, uint8_t *mask) {
    const float t = (float) (thresh * thresh);
    int good = 0;
    for (int i = 0; i < count; i++) {
        const double x1 = m1[2 * i], y1 = m1[2 * i + 1], x2 = m2[2 * i], y2 = m2[2 * i + 1];
        double a = F[0] * x1 + F[1] * y1 + F[2], b = F[3] * x1 + F[4] * y1 + F[5], c = F[6] * x1 + F[7] * y1 + F[8];
        const double s2 = 1. / (a * a + b * b), d2 = x2 * a + y2 * b + c;
        a = F[0] * x2 + F[3] * y2 + F[6], b = F[1] * x2 + F[4] * y2 + F[7], c = F[2] * x2 + F[5] * y2 + F[8];
        const double s1 = 1. / (a * a + b * b), d1 = x1 * a + y1 * b + c;
        const float err = (float) std::max(d1 * d1 * s1, d2 * d2 * s2);
        good += mask[i] = err <= t;
    }
    return good;
}

int update_num_iters(double p, double ep, int model_points, int max_iters) {  // RANSACUpdateNumIters
    p = std::min(std::max(p, 0.), 1.), ep = std::min(std::max(ep, 0.), 1.);
    double num = std::max(1. - p, DBL_MIN), denom = 1. - pow(1. - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num), denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int) lrint(num / denom);
}

}  // namespace

extern "C" {

int icg_find_fundamental_mat_ransac(const float *pts1_xy, const float *pts2_xy, int n, double threshold, double confidence, int max_iters,
                                    uint8_t *status, double *F9) {
    if (!pts1_xy || !pts2_xy || !status || n < 0 || max_iters < 1) {
        set_error("icg_find_fundamental_mat_ransac: bad arguments");
        return ICG_EINVAL;
    }
    if (n < 15) {  // cv::findFundamentalMat switches to other estimators below 15 points; the reference only calls it with >= 15 (tracking.cc:546)
        set_error("icg_find_fundamental_mat_ransac: needs at least 15 point pairs (got %d)", n);
        return ICG_EUNSUPPORTED;
    }
    if (threshold <= 0) threshold = 3;
    if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
    CvRng rng;
    std::vector<uint8_t> mask(n), best(n, 0);
    double bestF[9] = {0};
    int niters = max_iters, max_good = 0;
    float ms1[14], ms2[14];
    for (int iter = 0; iter < niters; iter++) {
        // RANSACPointSetRegistrator::getSubset: 7 distinct indices, the subset redrawn while either image has collinear points
        bool found = false;
        for (int attempts = 0; attempts < 10000; attempts++) {
            int idx[7];
            for (int i = 0; i < 7; i++) {
                int v;
                bool dup;
                do {
                    v = rng.uniform(0, n);
                    dup = false;
                    for (int j = 0; j < i; j++) dup = dup || idx[j] == v;
                } while (dup);
                idx[i] = v;
                ms1[2 * i] = pts1_xy[2 * v], ms1[2 * i + 1] = pts1_xy[2 * v + 1];
                ms2[2 * i] = pts2_xy[2 * v], ms2[2 * i + 1] = pts2_xy[2 * v + 1];
            }
            if (collinear_last(ms1, 7) || collinear_last(ms2, 7)) continue;
            found = true;
            break;
        }
        if (!found) {
            if (iter == 0) {
                memset(status, 0, n);
                if (F9) memset(F9, 0, sizeof(double) * 9);
                return ICG_OK;
            }
            break;
        }
        double F[3][9];
        const int nmodels = run_7point(ms1, ms2, F);
        if (nmodels <= 0) continue;
        for (int k = 0; k < nmodels; k++) {
            const int good = find_inliers(pts1_xy, pts2_xy, n, F[k], threshold, mask.data());
            if (good > std::max(max_good, 6)) {
                std::swap(mask, best);
                memcpy(bestF, F[k], sizeof(bestF));
                max_good = good;
                niters = update_num_iters(confidence, (double) (n - good) / n, 7, niters);
            }
        }
    }
    memcpy(status, best.data(), n);
    if (F9) memcpy(F9, bestF, sizeof(bestF));
    return ICG_OK;
}

// Tracking::triangulatePoint (IG/tracking/tracking.cc:796-808) for n point pairs: rows of the 4 x 4 design matrix
//   pc0.x * P0.row(2) - P0.row(0),  pc0.y * P0.row(2) - P0.row(1),  pc1.x * P1.row(2) - P1.row(0),  pc1.y * P1.row(2) - P1.row(1)
// (P = T_c_w, 3 x 4 row-major), pw = right singular vector of the smallest singular value, dehomogenised (scale / sign invariant, so any
// accurate SVD gives the reference's Eigen::JacobiSVD result up to rounding).  Host function: <= 300 pairs, the caller interleaves it
// with map bookkeeping (:715-789).
int icg_triangulate_points(const double *Tcw0 /* n x 12 */, const double *Tcw1 /* 12 */, const double *pc0_xy, const double *pc1_xy, int n,
                           double *pw_xyz) {
    if ((!Tcw0 || !Tcw1 || !pc0_xy || !pc1_xy || !pw_xyz) && n > 0) {
        set_error("icg_triangulate_points: bad arguments");
        return ICG_EINVAL;
    }
    for (int i = 0; i < n; i++) {
        const double *P0 = Tcw0 + 12 * (size_t) i, *P1 = Tcw1;
        double G[4][4], V[4][4];  // column-major columns of the design matrix / of V
        for (int c = 0; c < 4; c++) {
            G[c][0] = pc0_xy[2 * i] * P0[8 + c] - P0[c];
            G[c][1] = pc0_xy[2 * i + 1] * P0[8 + c] - P0[4 + c];
            G[c][2] = pc1_xy[2 * i] * P1[8 + c] - P1[c];
            G[c][3] = pc1_xy[2 * i + 1] * P1[8 + c] - P1[4 + c];
            for (int r = 0; r < 4; r++) V[c][r] = r == c ? 1.0 : 0.0;
        }
        for (int sweep = 0; sweep < 60; sweep++) {  // one-sided Jacobi: orthogonalise the columns
            bool rotated = false;
            for (int p = 0; p < 3; p++)
                for (int q = p + 1; q < 4; q++) {
                    double al = 0, be = 0, ga = 0;
                    for (int r = 0; r < 4; r++) al += G[p][r] * G[p][r], be += G[q][r] * G[q][r], ga += G[p][r] * G[q][r];
                    if (ga == 0.0 || fabs(ga) <= 1e-16 * sqrt(al * be)) continue;
                    rotated = true;
                    const double zeta = (be - al) / (2.0 * ga);
                    const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                    for (int r = 0; r < 4; r++) {
                        const double x = G[p][r], y = G[q][r];
                        G[p][r] = cs * x - sn * y, G[q][r] = sn * x + cs * y;
                        const double vx = V[p][r], vy = V[q][r];
                        V[p][r] = cs * vx - sn * vy, V[q][r] = sn * vx + cs * vy;
                    }
                }
            if (!rotated) break;
        }
        int best = 0;
        double bn = 1e300;
        for (int c = 0; c < 4; c++) {
            double nn = 0;
            for (int r = 0; r < 4; r++) nn += G[c][r] * G[c][r];
            if (nn < bn) bn = nn, best = c;
        }
        for (int k = 0; k < 3; k++) pw_xyz[3 * (size_t) i + k] = V[best][k] / V[best][3];
    }
    return ICG_OK;
}

}  // extern "C"
