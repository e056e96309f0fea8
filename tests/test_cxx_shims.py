"""CPU test: the C++ host-side shims compile (syntax + types) -- BundlerLib.h against the reference's Eigen/GSL when the
reference tree is present (build container only), OrbDetector.hpp standalone."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDEP = "/root/reference/Dependencies"


def _compile(src, incs):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.cpp")
        open(p, "w").write(src)
        cmd = ["g++", "-std=c++14", "-fsyntax-only", "-w", p] + ["-I" + i for i in incs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]


def test_orb_detector_shim_compiles():
    _compile('#include "mageslam_b200/OrbDetector.hpp"\n'
             'int f(){ mage_b200::OrbDetector d(7,2000,1.2f,8,31,10,true,1.5f,0.9f,20,1.1f,2.0f,32,32); std::vector<mage_dmatch> m; '
             'mage_camera_calibration c{}; mage_keypoint k{}; mage_b200::UndistortKeypoints(&k, 1, c, c); '
             'return (int)sizeof(d) + (int)mage_b200::Match(nullptr,nullptr,0,nullptr,nullptr,0,nullptr,30,1,m); }\n',
             [os.path.join(ROOT, "include")])


@pytest.mark.skipif(not os.path.isdir(REFDEP), reason="reference tree (Eigen/GSL headers) not present on this box")
def test_bundlerlib_shim_is_source_compatible_with_reference_call_sites():
    # the call pattern of TrackLocalMap::OptimizeCameraPose (reference TrackLocalMap.cpp:439-494) against the shim
    _compile('#include "mageslam_b200/BundlerLib.h"\n'
             'float f(){ mage::BundlerLib b{ mage::BundlerParameters{ true } }; float p[3]={0,0,0}, R[9]={1,0,0,0,1,0,0,0,1}, k[4]={320,240,500,500}, uv[2]={1,2};\n'
             ' b.AllocateCameras(1); b.AllocateMapPoints(1); b.AllocateObservations(1);\n'
             ' b.SetCameraPose(0, Eigen::Map<const Eigen::Vector3f>(p), Eigen::Map<const Eigen::Matrix3f>(R), Eigen::Map<const Eigen::Vector4f>(k), false);\n'
             ' b.SetMapPoint(0, Eigen::Map<const Eigen::Vector3f>(p)); b.SetObservation(0, Eigen::Map<const Eigen::Vector2f>(uv), 0, 0, 1.0f);\n'
             ' b.AllocateFixedDistanceConstraints(1); b.SetFixedDistanceConstraint(0, 0, 1); b.AllocateRelativeRotationConstraints(1);\n'
             ' b.SetRelativeRotationConstraint(0, 0, 1, Eigen::Quaternionf::Identity(), 2.f); b.AllocateRelativeTransformConstraints(1);\n'
             ' b.SetRelativeTransformConstraint(0, 0, 1, Eigen::Map<const Eigen::Vector3f>(p), Eigen::Quaternionf::Identity(), 3.f);\n'
             ' std::vector<unsigned int> out; float hub[3]={2,2,2}; float e = b.StepBundleAdjustment(hub, 25.f, out);\n'
             ' b.GetPose(0, Eigen::Map<Eigen::Vector3f>(p), Eigen::Map<Eigen::Matrix3f>(R)); b.GetPoint(0, Eigen::Map<Eigen::Vector3f>(p)); b.SetCurrentLambda(1.f);\n'
             ' return e + b.GetCurrentLambda(); }\n',
             [os.path.join(ROOT, "include"), REFDEP + "/eigen-git-mirror", REFDEP + "/GSL/include"])
