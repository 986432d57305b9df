"""diagnostic: per-kernel CUDA-event times of the LM iteration's kernels on the C2 stage-S4 problem"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lvi_exc_b200 import synth, workload
from lvi_exc_b200.backend import CudaBackend, CudaProblem
dur = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seq = synth.make_sequence(synth.default_config(duration=dur))
b = CudaBackend(0)
pd, info = workload.lvi_stage_problem(seq, b)
prob = CudaProblem(b, pd)
prob.bench_iterations(3)
b.kernel_timing(True); b.kernel_times()
n = 10
ms = prob.bench_iterations(n)
kt = b.kernel_times()
print("phases", [round(float(x), 3) for x in ms])
for k, (c, t) in sorted(kt.items(), key=lambda kv: -kv[1][1])[:16]:
    print("  %-42s %4d launches %8.3f ms/iter" % (k, c, t / n))
