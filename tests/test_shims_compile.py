"""The header-only C++ shims (ic_gvins_b200/host/icg_shims.hpp: the OpenCV / MarginalizationInfo call signatures over the C ABI) must
compile as plain C++17 against include/icgvins_b200.h and link against the library's exports.  CPU only (no compute calls)."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_shim_header_compiles_and_links():
    lib = os.path.join(ROOT, "ic_gvins_b200", "libicgvins_b200.so")
    if not os.path.exists(lib):
        from ic_gvins_b200 import build
        build.build()
    src = r'''
#include "ic_gvins_b200/host/icg_shims.hpp"
#include "ic_gvins_b200/host/icg_factors.hpp"
// instantiate every shim member so that all C-ABI symbols are referenced (never executed: no GPU here)
int main(int argc, char **) {
    if (argc > 1000) {
        icg_b200::KltContext k(1280, 560);
        icg_b200::Mat a{nullptr, 560, 1280, 1280};
        std::vector<icg_b200::Point2f> p, q;
        std::vector<uint8_t> st;
        std::vector<float> err;
        k.calcOpticalFlowPyrLK(a, a, p, q, st, err, icg_b200::Size(21, 21), 3, icg_b200::TermCriteria(3, 30, 0.01), 4);
        k.trackForwardBackward(a, a, p, q, st);
        icg_b200::CameraModel cm(icg_camera{787, 787, 640, 280, 0, 0, 0, 0, 0, 0});
        cm.undistortPoints(p);
        cm.distortPoints(p);
        const double pc[3] = {0.1, 0.2, 1.0};
        cm.distortCameraPoint(pc);
        icg_b200::Clahe c(1280, 560);
        c.apply(nullptr, 1280, nullptr, 1280);
        icg_b200::BlockDetector d(1280, 560, 18, 32, 213 * 186);
        std::vector<std::vector<icg_b200::Point2f>> f;
        d.detect(a, nullptr, {}, {}, 0.01, 40.0, f);
        icg_b200::WindowSolver s(10, 300, 2700);
        icg_ba_problem P{};
        icg_ba_summary o[2];
        int32_t culled[2];
        s.Solve(P, 5);
        s.gvinsOptimization(P, 20, o, culled);
        s.marginalization(P, 1);
        const double v[16] = {0};
        const double *params[5] = {v, v, v, v, v};
        double r[16], *J[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
        icg_b200::ReprojectionFactor(v, v, v, v, 0, 0, 1).Evaluate(params, r, J);
        icg_b200::GnssFactor(v, v, v).Evaluate(params, r, J);
        icg_b200::ImuErrorFactor().Evaluate(params, r, J);
        icg_b200::ImuPosePriorFactor(v, v).Evaluate(params, r, J);
        icg_b200::ImuMixPriorFactor(v, v).Evaluate(params, r, J);
        std::vector<double> blob(ICG_IMU_BLOB_DOUBLES);
        icg_b200::PreintegrationFactor(blob.data()).Evaluate(params, r, J);
        icg_b200::MarginalizationFactor(1, {3}, {0.0}, {1.0}, {0.0}).Evaluate(params, r, J);
    }
    return 0;
}
'''
    with tempfile.TemporaryDirectory() as td:
        cpp = os.path.join(td, "shim_test.cpp")
        exe = os.path.join(td, "shim_test")
        open(cpp, "w").write(src)
        r = subprocess.run(["g++", "-std=c++17", "-Wall", "-I", ROOT, cpp, "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([exe], capture_output=True, text=True)  # argc == 1: nothing is called, the binary only has to load
        assert r.returncode == 0, r.stderr
