// icg_factors.hpp -- header-only C++ cost functions with the reference's class names and the exact Ceres signature
//     bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const
// (IG/factors/reprojection_factor.h:55, IG/preintegration/preintegration_factor.h:45, IG/factors/gnss_factor.h:43,
//  IG/preintegration/imu_error_factor.h:45, imu_pose_prior_factor.h:42, imu_mix_prior_factor.h:40, IG/factors/marginalization_factor.h:47),
// forwarding to the C ABI (include/icgvins_b200.h): the arithmetic runs on the device.
//
// With Ceres headers present (`#include <ceres/ceres.h>` before this file, or -DICG_WITH_CERES) the classes derive from
// ceres::SizedCostFunction / ceres::CostFunction with the reference's block sizes, so `problem.AddResidualBlock(new icg_b200::ReprojectionFactor(...),
// loss, pose0, pose1, extrinsic, invdepth, td)` compiles unchanged at IG/ic_gvins.cc:1826-1831, :1870-1872, :1877-1887, :1896-1903, :1158-1161.
// Without Ceres (this image) they are plain classes with the same constructor / Evaluate surface, which is what tests/test_shims_gpu.py runs.
//
// These per-factor entry points are the SEAM (drop-in, testable one factor at a time); the fast path is the batched window solve
// (icg_b200::WindowSolver in icg_shims.hpp), which evaluates all factors of all windows in a handful of launches.
#pragma once
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/icgvins_b200.h"

#if defined(ICG_WITH_CERES) || defined(CERES_PUBLIC_CERES_H_)
#include <ceres/ceres.h>
#define ICG_SIZED_BASE(...) : public ceres::SizedCostFunction<__VA_ARGS__>
#define ICG_DYN_BASE : public ceres::CostFunction
#define ICG_OVERRIDE override
#else
#define ICG_SIZED_BASE(...)
#define ICG_DYN_BASE
#define ICG_OVERRIDE
#endif

namespace icg_b200 {

// One evaluation handle per process (the reference evaluates factors from Ceres' worker threads: calls are serialised here; the handle
// itself is not re-entrant).  Created on first use on device 0 (or ICG_DEVICE).
class FactorContext {
public:
    static FactorContext &instance() {
        static FactorContext c;
        return c;
    }
    icg_ba *handle() { return h_; }
    std::mutex &mutex() { return m_; }

private:
    FactorContext() {
        const char *dev = std::getenv("ICG_DEVICE");
        if (icg_ba_create(&h_, 1, 2, 1, 1, 1, 1, dev ? std::atoi(dev) : 0, nullptr) != ICG_OK)
            throw std::runtime_error(std::string("icg_b200::FactorContext: ") + icg_last_error());
    }
    ~FactorContext() { icg_ba_destroy(h_); }
    icg_ba *h_ = nullptr;
    std::mutex m_;
};

inline bool factor_ok(int rc) { return rc == ICG_OK; }

// ReprojectionFactor(pts0, pts1, vel0, vel1, td0, td1, std)  -- SizedCostFunction<2, 7, 7, 7, 1, 1> (reprojection_factor.h:36-53)
// Vector3d arguments are passed as pointers to their 3 doubles (Eigen: v.data()).
class ReprojectionFactor ICG_SIZED_BASE(2, 7, 7, 7, 1, 1) {
public:
    ReprojectionFactor(const double *pts0, const double *pts1, const double *vel0, const double *vel1, double td0, double td1, double std) : std_(std) {
        std::memcpy(c_, pts0, 24), std::memcpy(c_ + 3, pts1, 24), std::memcpy(c_ + 6, vel0, 24), std::memcpy(c_ + 9, vel1, 24);
        c_[12] = td0, c_[13] = td1;
    }
    // parameters: pose0[7], pose1[7], extrinsic[7], invdepth[1], td[1]
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const ICG_OVERRIDE {
        FactorContext &ctx = FactorContext::instance();
        std::lock_guard<std::mutex> lk(ctx.mutex());
        return factor_ok(icg_ba_reproj_evaluate(ctx.handle(), parameters[0], parameters[1], parameters[2], parameters[3], parameters[4], c_, std_, residuals, jacobians));
    }

private:
    double c_[14], std_;
};

// PreintegrationFactor(preintegration) -- CostFunction 15 x (7, 9, 7, 9) (preintegration_factor.h:31-69).  The preintegration object of the
// reference is represented by its blob (ICG_IMU_BLOB_DOUBLES doubles, icg_imu_preintegrate).
class PreintegrationFactor ICG_DYN_BASE {
public:
    explicit PreintegrationFactor(const double *imu_blob) : blob_(imu_blob, imu_blob + ICG_IMU_BLOB_DOUBLES) {
#if defined(ICG_WITH_CERES) || defined(CERES_PUBLIC_CERES_H_)
        *mutable_parameter_block_sizes() = std::vector<int>{7, 9, 7, 9};
        set_num_residuals(15);
#endif
    }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const ICG_OVERRIDE {
        FactorContext &ctx = FactorContext::instance();
        std::lock_guard<std::mutex> lk(ctx.mutex());
        return factor_ok(icg_ba_imu_evaluate(ctx.handle(), blob_.data(), parameters[0], parameters[1], parameters[2], parameters[3], residuals, jacobians));
    }

private:
    std::vector<double> blob_;
};

// GnssFactor(gnss {blh, std}, lever) -- SizedCostFunction<3, 7> (gnss_factor.h:31-71)
class GnssFactor ICG_SIZED_BASE(3, 7) {
public:
    GnssFactor(const double *blh, const double *std3, const double *lever) {
        std::memcpy(blh_, blh, 24), std::memcpy(std_, std3, 24), std::memcpy(lever_, lever, 24);
    }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const ICG_OVERRIDE {
        FactorContext &ctx = FactorContext::instance();
        std::lock_guard<std::mutex> lk(ctx.mutex());
        return factor_ok(icg_ba_gnss_evaluate(ctx.handle(), parameters[0], blh_, std_, lever_, residuals, jacobians));
    }

private:
    double blh_[3], std_[3], lever_[3];
};

// ImuErrorFactor -- SizedCostFunction<6, 9> (imu_error_factor.h:31-91)
class ImuErrorFactor ICG_SIZED_BASE(6, 9) {
public:
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const ICG_OVERRIDE {
        FactorContext &ctx = FactorContext::instance();
        std::lock_guard<std::mutex> lk(ctx.mutex());
        return factor_ok(icg_ba_imu_error_evaluate(ctx.handle(), parameters[0], residuals, jacobians));
    }
};

// ImuPosePriorFactor(pose, std) -- SizedCostFunction<6, 7> (imu_pose_prior_factor.h:31-68)
class ImuPosePriorFactor ICG_SIZED_BASE(6, 7) {
public:
    ImuPosePriorFactor(const double *pose7, const double *std6) { std::memcpy(pose_, pose7, 56), std::memcpy(std_, std6, 48); }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const ICG_OVERRIDE {
        FactorContext &ctx = FactorContext::instance();
        std::lock_guard<std::mutex> lk(ctx.mutex());
        return factor_ok(icg_ba_pose_prior_evaluate(ctx.handle(), parameters[0], pose_, std_, residuals, jacobians));
    }

private:
    double pose_[7], std_[6];
};

// ImuMixPriorFactor(mix, std) -- CostFunction 9 x 9 (imu_mix_prior_factor.h:31-75)
class ImuMixPriorFactor ICG_SIZED_BASE(9, 9) {
public:
    ImuMixPriorFactor(const double *mix9, const double *std9) { std::memcpy(mix_, mix9, 72), std::memcpy(std_, std9, 72); }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const ICG_OVERRIDE {
        FactorContext &ctx = FactorContext::instance();
        std::lock_guard<std::mutex> lk(ctx.mutex());
        return factor_ok(icg_ba_mix_prior_evaluate(ctx.handle(), parameters[0], mix_, std_, residuals, jacobians));
    }

private:
    double mix_[9], std_[9];
};

// MarginalizationFactor(marginalization_info) -- CostFunction r x (remained block sizes) (marginalization_factor.h:31-101).  The
// MarginalizationInfo of the reference is represented by the prior icg_ba_marginalize returns (block types, x0, J0, e0).
class MarginalizationFactor ICG_DYN_BASE {
public:
    MarginalizationFactor(int r, const std::vector<int32_t> &block_type, const std::vector<double> &x0, const std::vector<double> &J0, const std::vector<double> &e0)
        : r_(r), type_(block_type), x0_(x0), J0_(J0), e0_(e0) {
#if defined(ICG_WITH_CERES) || defined(CERES_PUBLIC_CERES_H_)
        for (int32_t t : type_) mutable_parameter_block_sizes()->push_back(t == 1 ? 9 : t == 3 ? 1 : 7);
        set_num_residuals(r_);
#endif
    }
    bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const ICG_OVERRIDE {
        FactorContext &ctx = FactorContext::instance();
        std::lock_guard<std::mutex> lk(ctx.mutex());
        return factor_ok(icg_ba_marg_factor_evaluate(ctx.handle(), r_, (int) type_.size(), type_.data(), parameters, x0_.data(), J0_.data(), e0_.data(), residuals,
                                                     jacobians));
    }

private:
    int r_;
    std::vector<int32_t> type_;
    std::vector<double> x0_, J0_, e0_;
};

}  // namespace icg_b200
