/*
 * oracle/cxx_driver_ba.cpp -- TEST INFRASTRUCTURE: one C++ translation unit that only speaks the reference's BundlerLib interface
 * (ref Dependencies/BundlerLib/Include/BundlerLib.h:20-66), written the way Core/MAGESLAM/Source/BundleAdjustment/BundleAdjust.cpp
 * drives it (ref BundleAdjust.cpp:60-180: allocate, set cameras / points / observations, StepBundleAdjustment with a Huber schedule,
 * read poses and points back). The SAME file is compiled twice by oracle/Makefile:
 *   _ref/ba_driver_ref   -I <reference>/Dependencies/BundlerLib/Include, linked with the reference's BundlerLib.cpp + g2o objects
 *   _ref/ba_driver_b200  -I include/mageslam_b200 (the header-compatible shim),  linked with libmage_b200.so
 * and tests/test_cxx_boundary.py runs both and compares what they print. The problem is generated in here (fixed LCG), so both
 * builds see identical inputs.
 */
#include "BundlerLib.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace {
struct Lcg {
    uint64_t s;
    double uni() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)(s >> 11) / 9007199254740992.0; }
    double sym(double a) { return (2.0 * uni() - 1.0) * a; }
};
Eigen::Matrix3f rotZYX(double rz, double ry, double rx)
{
    return (Eigen::AngleAxisd(rz, Eigen::Vector3d::UnitZ()) * Eigen::AngleAxisd(ry, Eigen::Vector3d::UnitY()) * Eigen::AngleAxisd(rx, Eigen::Vector3d::UnitX()))
        .toRotationMatrix().cast<float>();
}
}

int main(int argc, char** argv)
{
    const int K = argc > 1 ? atoi(argv[1]) : 10, P = argc > 2 ? atoi(argv[2]) : 2000, D = argc > 3 ? atoi(argv[3]) : 4;
    const int steps = argc > 4 ? atoi(argv[4]) : 2;
    const bool tethers = argc > 5 && atoi(argv[5]) != 0;
    Lcg rng{ 0x9E3779B97F4A7C15ull };
    const Eigen::Vector4f intr(320.f, 240.f, 500.f, 500.f);                 // cx, cy, fx, fy (ref BundlerLib.cpp:193-200)

    // cameras on an arc looking at a box of points (world -> camera: x_c = R x_w + t, as g2o's SBACam)
    std::vector<Eigen::Matrix3f> R(K), Rn(K);
    std::vector<Eigen::Vector3f> t(K), tn(K);
    for (int k = 0; k < K; k++) {
        const double a = 0.03 * (k - 0.5 * (K - 1)) / std::max(K, 1);
        R[k] = rotZYX(0.01 * std::sin(3.0 * k), a, 0.01 * std::cos(2.0 * k));
        const Eigen::Vector3f centre((float)(0.25 * (k - 0.5 * (K - 1))), (float)(0.02 * std::sin(1.7 * k)), (float)(0.05 * std::cos(0.7 * k)));
        t[k] = -R[k] * centre;
        Rn[k] = k < 2 ? R[k] : Eigen::Matrix3f(rotZYX(rng.sym(0.004), rng.sym(0.004), rng.sym(0.004)) * R[k]);
        tn[k] = k < 2 ? t[k] : Eigen::Vector3f(t[k] + Eigen::Vector3f((float)rng.sym(0.02), (float)rng.sym(0.02), (float)rng.sym(0.02)));
    }
    std::vector<Eigen::Vector3f> X(P), Xn(P);
    for (int p = 0; p < P; p++) {
        X[p] = Eigen::Vector3f((float)rng.sym(3.0), (float)rng.sym(2.0), (float)(5.5 + rng.sym(2.5)));
        Xn[p] = X[p] + Eigen::Vector3f((float)rng.sym(0.03), (float)rng.sym(0.03), (float)rng.sym(0.03));
    }
    struct Obs { int cam, pt; Eigen::Vector2f uv; float info; };
    std::vector<Obs> obs;
    for (int p = 0; p < P; p++)
        for (int d = 0; d < D; d++) {
            const int k = (int)((p + (size_t)d * (K / D > 0 ? K / D : 1) + (d ? (int)(rng.uni() * 2) : 0)) % K);
            const Eigen::Vector3f xc = R[k] * X[p] + t[k];
            Eigen::Vector2f uv(intr[2] * xc[0] / xc[2] + intr[0] + (float)rng.sym(0.7), intr[3] * xc[1] / xc[2] + intr[1] + (float)rng.sym(0.7));
            if (rng.uni() < 0.01) uv += Eigen::Vector2f((float)rng.sym(40.0), (float)rng.sym(40.0));          // a few gross outliers
            obs.push_back({ k, p, uv, 1.0f / (1.0f + 0.44f * (float)(p % 4)) });                              // per-octave information scale
        }

    mage::BundlerLib ba{ mage::BundlerParameters{ false } };
    ba.AllocateCameras(K);
    for (int k = 0; k < K; k++)
        ba.SetCameraPose(k, Eigen::Map<const Eigen::Vector3f>(tn[k].data()), Eigen::Map<const Eigen::Matrix3f>(Rn[k].data()),
                         Eigen::Map<const Eigen::Vector4f>(intr.data()), k < 2);
    ba.AllocateMapPoints(P);
    for (int p = 0; p < P; p++) ba.SetMapPoint(p, Eigen::Map<const Eigen::Vector3f>(Xn[p].data()));
    ba.AllocateObservations(obs.size());
    for (size_t i = 0; i < obs.size(); i++) ba.SetObservation(i, Eigen::Map<const Eigen::Vector2f>(obs[i].uv.data()), obs[i].cam, obs[i].pt, obs[i].info);
    if (tethers) {
        ba.AllocateFixedDistanceConstraints(2);
        ba.SetFixedDistanceConstraint(0, 2, 3, (t[2] - t[3]).norm(), 500.0f);
        ba.SetFixedDistanceConstraint(1, 4, 6, (t[4] - t[6]).norm(), 900.0f);
        ba.AllocateRelativeRotationConstraints(1);
        ba.SetRelativeRotationConstraint(0, 3, 5, Eigen::Quaternionf(Eigen::Matrix3f(R[3].transpose() * R[5])), 1000.0f);
        ba.AllocateRelativeTransformConstraints(1);
        const Eigen::Matrix3f dR = R[7 % K] * R[2].transpose();
        const Eigen::Vector3f dt = t[7 % K] - dR * t[2];
        ba.SetRelativeTransformConstraint(0, 2, 7 % K, Eigen::Map<const Eigen::Vector3f>(dt.data()), Eigen::Quaternionf(dR), 2e5f);
    }

    const float huber[10] = { 2.5f, 2.5f, 2.0f, 2.0f, 1.8f, 1.8f, 1.8f, 1.8f, 1.8f, 1.8f };
    for (int s = 0; s < steps; s++) {
        std::vector<unsigned int> outliers;
        const float mean = ba.StepBundleAdjustment(gsl::span<const float>(huber, 10), s + 1 < steps ? 1e9f : 36.0f, outliers);
        printf("step %d mean %.9g lambda %.9g outliers %zu", s, mean, ba.GetCurrentLambda(), outliers.size());
        unsigned long long h = 1469598103934665603ull;
        for (unsigned int o : outliers) h = (h ^ o) * 1099511628211ull;
        printf(" outlier_hash %llx\n", h);
    }
    for (int k = 0; k < K; k++) {
        Eigen::Vector3f p; Eigen::Matrix3f r;
        ba.GetPose(k, Eigen::Map<Eigen::Vector3f>(p.data()), Eigen::Map<Eigen::Matrix3f>(r.data()));
        printf("pose %d", k);
        for (int i = 0; i < 3; i++) printf(" %.9g", p[i]);
        for (int i = 0; i < 9; i++) printf(" %.9g", r.data()[i]);
        printf("\n");
    }
    for (int p = 0; p < P; p++) {
        Eigen::Vector3f x;
        ba.GetPoint(p, Eigen::Map<Eigen::Vector3f>(x.data()));
        printf("point %d %.9g %.9g %.9g\n", p, x[0], x[1], x[2]);
    }
    return 0;
}
