"""diagnostic: the camera / surfel normal-equation kernels timed ALONE (serialised by ncu-like env: LVI_LIN_SERIAL=1 puts every type on one stream)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lvi_exc_b200 import synth, workload
from lvi_exc_b200.backend import CudaBackend, CudaProblem
seq = synth.make_sequence(synth.default_config(duration=60.0))
b = CudaBackend(0)
pd, info = workload.lvi_stage_problem(seq, b)
prob = CudaProblem(b, pd)
prob.bench_iterations(3)
b.kernel_timing(True); b.kernel_times()
n = 10
ms = prob.bench_iterations(n)
kt = b.kernel_times()
print("debug", os.environ.get("LVI_LIN_DEBUG"), "phases", [round(float(x), 3) for x in ms])
for k, (c, t) in sorted(kt.items(), key=lambda kv: -kv[1][1])[:8]:
    if "linearize" in k: print("  %-42s %4d launches %8.3f ms/iter" % (k, c, t / n))
