"""A few mage_optimize_camera_pose calls (300 points, 3 and 4 iterations) for ncu captures of k_ba_step_t<true>."""
import sys
sys.path.insert(0, ".")
from mageslam_b200 import synth
from mageslam_b200.tracking import OptimizeCameraPose
p = synth.ba_problem(K=1, P=300, obs_per_point=1, n_fixed=0, pose_sigma=0.03, outlier_frac=0.05, seed=5)
for it in (3, 4, 3, 4, 3, 4):
    r = OptimizeCameraPose(p["cam_pos"][0], p["cam_rot"][0], p["intrinsics"][0], p["points"], p["obs_uv"], p["obs_info"], it, 25.0, 2.0)
print(len(r[2]), r[3])
