// oracle/ba_ref.hpp -- CPU FP64 restatement of IC-GVINS' sliding-window factor graph (TEST INFRASTRUCTURE ONLY).
//
// PARITY UNPINNED at the Ceres boundary: the reference solves with Ceres 2.0/2.1 (README.md:46,
// ic_gvins/CMakeLists.txt:41), which is neither vendored nor installed here, and the reference ships no tests or
// golden vectors for this path.  The factor arithmetic below follows the reference's own
// headers line by line (cited at each function); the trust-region loop restates Ceres' published
// LEVENBERG_MARQUARDT / DENSE_SCHUR algorithm with the defaults the reference leaves in place
// (ic_gvins/ic_gvins/ic_gvins.cc:1143-1146).  Self-consistency is pinned by tests/test_oracle_ba.py
// (finite-difference Jacobians through Plus, Schur-vs-full normal equations, convergence to ground truth).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace icgo {

// ------------------------------------------------------------------------------------------- small fixed-size algebra
struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

struct M3 {
    double m[9];  // row-major
    double &operator()(int r, int c) { return m[3 * r + c]; }
    double operator()(int r, int c) const { return m[3 * r + c]; }
};
inline M3 m3_identity() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline M3 operator*(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
    return r;
}
inline V3 operator*(const M3 &a, V3 v) {
    return {a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
inline M3 operator*(double s, const M3 &a) {
    M3 r;
    for (int i = 0; i < 9; i++) r.m[i] = s * a.m[i];
    return r;
}
inline M3 operator*(const M3 &a, double s) { return s * a; }
inline M3 operator+(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 9; i++) r.m[i] = a.m[i] + b.m[i];
    return r;
}
inline M3 operator-(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 9; i++) r.m[i] = a.m[i] - b.m[i];
    return r;
}
inline M3 operator-(const M3 &a) { return -1.0 * a; }
inline M3 transpose(const M3 &a) { return {{a.m[0], a.m[3], a.m[6], a.m[1], a.m[4], a.m[7], a.m[2], a.m[5], a.m[8]}}; }
// Rotation::skewSymmetric (ic_gvins/ic_gvins/common/rotation.h:97-101)
inline M3 skew(V3 v) { return {{0, -v.z, v.y, v.z, 0, -v.x, -v.y, v.x, 0}}; }

struct Q {  // Eigen::Quaterniond semantics; constructor order (w, x, y, z)
    double w, x, y, z;
};
inline V3 vec(Q q) { return {q.x, q.y, q.z}; }
inline Q operator*(Q a, Q b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
inline Q q_inverse(Q q) {  // Eigen: conjugate / squaredNorm
    double n2 = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z;
    return {q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
}
inline Q q_normalized(Q q) {
    double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return {q.w / n, q.x / n, q.y / n, q.z / n};
}
inline M3 q_matrix(Q q) {  // Eigen::QuaternionBase::toRotationMatrix
    double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    return {{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)}};
}
inline V3 q_rotate(Q q, V3 v) {  // Eigen: q * v
    V3 uv = cross(vec(q), v);
    uv    = uv + uv;
    return v + q.w * uv + cross(vec(q), uv);
}
// Rotation::rotvec2quaternion (rotation.h:72-76): AngleAxis(|v|, v/|v|); zero vector -> identity
inline Q rotvec2quaternion(V3 rv) {
    double angle = norm(rv);
    V3 axis      = angle > 0 ? rv / angle : rv;
    double s = std::sin(0.5 * angle), c = std::cos(0.5 * angle);
    return {c, s * axis.x, s * axis.y, s * axis.z};
}
// Rotation::quaternionleft / quaternionright bottom-right 3x3 (rotation.h:103-119)
inline M3 qleft_br(Q q) { return q.w * m3_identity() + skew(vec(q)); }
inline M3 qright_br(Q p) { return p.w * m3_identity() - skew(vec(p)); }

inline Q pose_q(const double *pose) { return {pose[6], pose[3], pose[4], pose[5]}; }
inline V3 pose_p(const double *pose) { return {pose[0], pose[1], pose[2]}; }

// PoseParameterization::Plus (ic_gvins/ic_gvins/factors/pose_parameterization.h:34-49)
void pose_plus(const double *x, const double *delta, double *x_plus_delta);

// ------------------------------------------------------------------------------------------- cost functions
// ceres::CostFunction contract: Evaluate(parameters, residuals, jacobians); jacobians[i] row-major nres x global size.
struct CostFunction {
    virtual ~CostFunction() {}
    virtual int num_residuals() const = 0;
    virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
};

// ReprojectionFactor : SizedCostFunction<2,7,7,7,1,1> (factors/reprojection_factor.h:36-147)
struct ReprojectionFactor : CostFunction {
    V3 pts0, pts1, vel0, vel1;
    double td0, td1, std_;
    ReprojectionFactor(V3 p0, V3 p1, V3 v0, V3 v1, double t0, double t1, double s) : pts0(p0), pts1(p1), vel0(v0), vel1(v1), td0(t0), td1(t1), std_(s) {}
    int num_residuals() const override { return 2; }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
};

// GnssFactor : SizedCostFunction<3,7> (factors/gnss_factor.h:31-71)
struct GnssFactor : CostFunction {
    V3 blh, std_, lever;
    GnssFactor(V3 b, V3 s, V3 l) : blh(b), std_(s), lever(l) {}
    int num_residuals() const override { return 3; }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
};

// Preintegration result consumed by PreintegrationFactor (preintegration/preintegration_earth.cc:37-164).
struct Preintegration {
    bool normal = false;  // PreintegrationNormal (iswithearth: false, preintegration/preintegration_normal.cc) instead of PreintegrationEarth
    double delta_time = 0;
    V3 dp{0, 0, 0}, dv{0, 0, 0};
    Q dq{1, 0, 0, 0};
    V3 bg{0, 0, 0}, ba{0, 0, 0};  // delta_state_.bg / ba: linearisation biases
    double jacobian[225];         // 15x15 row-major
    double covariance[225];
    V3 gravity{0, 0, 9.8}, iewn{0, 0, 0};
    Q q0{1, 0, 0, 0};
    std::vector<double> pn;  // (dt, px, py, pz) per IMU epoch (preintegration_earth.cc:234)
    // working state for integrationProcess
    V3 cur_p{0, 0, 0}, cur_v{0, 0, 0};
    Q cur_q{1, 0, 0, 0};
    double noise[144];
    double corr_time = 3600;
};

// PreintegrationFactor : CostFunction 15 x (7,9,7,9) (preintegration/preintegration_factor.h:45-69)
struct PreintegrationFactor : CostFunction {
    const Preintegration *pre;
    explicit PreintegrationFactor(const Preintegration *p) : pre(p) {}
    int num_residuals() const override { return 15; }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
};

// ImuErrorFactor 6x9 (preintegration/imu_error_factor.h:45-91)
struct ImuErrorFactor : CostFunction {
    int num_residuals() const override { return 6; }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
};
// ImuPosePriorFactor 6x7 (preintegration/imu_pose_prior_factor.h:42-68)
struct ImuPosePriorFactor : CostFunction {
    double pose[7], sqrt_info[6];
    int num_residuals() const override { return 6; }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
};
// ImuMixPriorFactor 9x9 (preintegration/imu_mix_prior_factor.h:40-75)
struct ImuMixPriorFactor : CostFunction {
    double mix[9], mix_std[9];
    int num_residuals() const override { return 9; }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
};
// MarginalizationFactor (factors/marginalization_factor.h:47-101): e = e0 + J0 dx
struct MarginalizationFactor : CostFunction {
    int r = 0;                       // remained (local) size == number of residuals
    std::vector<int> block_size;     // global sizes (7, 9, 7, 1)
    std::vector<int> block_index;    // column offset of each block in J0 (local coordinates)
    std::vector<double> x0;          // concatenated linearisation points (global sizes)
    std::vector<double> J0, e0;      // J0 row-major r x r
    int num_residuals() const override { return r; }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
};

// ------------------------------------------------------------------------------------------- IMU propagation (B3, host side)
// PreintegrationEarth::resetState / integrationProcess / updateJacobianAndCovariance / setNoiseMatrix
// (preintegration/preintegration_earth.cc:205-338)
void preint_reset(Preintegration &P, V3 p, Q q, V3 v, V3 bg, V3 ba, V3 iewn, V3 gravity, double gyr_arw, double acc_vrw, double gyr_bias_std,
                  double acc_bias_std, double corr_time);
void preint_add_imu(Preintegration &P, const double *imu_pre /* dt, dtheta[3], dvel[3] */, const double *imu_cur);
void imu_sqrt_information(const double *cov, double *U);

// ------------------------------------------------------------------------------------------- window problem + LM
struct WindowProblem {
    int K = 0, L = 0;
    std::vector<double> pose, mix;  // K*7, K*9
    double ext[8];                  // t(3) q_xyzw(4) td
    std::vector<double> invdepth;   // L
    bool ext_const = false, td_const = false;
    // reprojection factors
    std::vector<int> f_lm, f_ref, f_obs;
    std::vector<double> f_const;  // F*14: pts0 pts1 vel0 vel1 td0 td1
    std::vector<uint8_t> f_active;
    double reproj_std = 1.5 / 787.0;
    bool reproj_huber = true;
    // imu
    std::vector<Preintegration> preint;  // between node k and k+1
    bool has_imu_error = true;
    bool has_pose_prior = false, has_mix_prior = false;
    ImuPosePriorFactor pose_prior;
    ImuMixPriorFactor mix_prior;
    // gnss
    std::vector<int> gnss_node;
    std::vector<double> gnss_blh, gnss_std;  // n*3
    V3 lever{0, 0, 0};
    bool gnss_huber = true;
    // marginalization prior
    bool has_marg = false;
    MarginalizationFactor marg;
    std::vector<int> marg_block_type, marg_block_node;  // type 0 pose, 1 mix, 2 ext, 3 td
};

struct SolveSummary {
    int iterations = 0;            // LM iterations executed (successful + unsuccessful)
    int num_successful_steps = 0;  // summary.num_successful_steps (ic_gvins.cc:1186)
    int termination = 0;           // 0 NO_CONVERGENCE (max iterations), 1 CONVERGENCE, 2 FAILURE
    double initial_cost = 0, final_cost = 0;
    double final_radius = 0;
};

// ceres::Solver::Solve(LEVENBERG_MARQUARDT, DENSE_SCHUR) restatement; updates the problem's parameters in place.
SolveSummary solve(WindowProblem &P, int max_num_iterations, int num_threads);
// Problem::EvaluateResidualBlock(id, false, &cost, ...) for the two chi2 passes (ic_gvins.cc:1251,1278)
void reproj_costs(const WindowProblem &P, std::vector<double> &cost);
void gnss_costs(const WindowProblem &P, std::vector<double> &cost);

// MarginalizationInfo::marginalization driven as GVINS::gvinsMarginalization does (ic_gvins.cc:1412-1640,
// factors/marginalization_info.h:73-253): removes the num_marg oldest nodes and the landmarks anchored in them.
struct MargOut {
    int m = 0, r = 0;                         // marginalizedSize(), remainedSize()
    std::vector<int> block_type, block_node;  // remained blocks (type 0 pose 1 mix 2 ext 3 td; node index after the removal)
    std::vector<double> x0;                   // remainedBlockData(), concatenated global sizes
    std::vector<double> J0, e0;               // linearizedJacobians() (r x r row-major), linearizedResiduals()
    std::vector<double> Hp, bp;               // the Schur complement itself (for invariant comparisons)
};
void marginalize(WindowProblem &W, int num_marg, MargOut &out);
void sym_eig_jacobi(std::vector<double> &A, int n, std::vector<double> &evals, std::vector<double> &V);

}  // namespace icgo
