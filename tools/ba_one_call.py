import sys
sys.path.insert(0, ".")
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib
prob = synth.ba_problem()
b = BundlerLib().load(prob)
print(b.StepBundleAdjustment([1.8] * 3, 1e9), b.stats())
