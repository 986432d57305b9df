"""diagnostics: where does the end-to-end solve spend its time (not a test)"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from lvi_exc_b200 import synth, workload
from lvi_exc_b200.backend import CudaBackend, CudaProblem
seq = synth.make_sequence(synth.default_config(duration=60.0))
cb = CudaBackend(0)
pd, info = workload.lvi_stage_problem(seq, cb)
saved = pd.clone_params()
tol0 = dict(function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
for steps in (10, 20, 20):
    pd.restore_params(saved)
    t0 = time.perf_counter(); p = CudaProblem(cb, pd); t1 = time.perf_counter()
    s = p.solve(steps, **tol0); t2 = time.perf_counter(); p.close(); t3 = time.perf_counter()
    print(f"steps {steps}: create {t1-t0:.3f} solve {t2-t1:.3f} destroy {t3-t2:.3f} | iters {s.num_iterations} ok {s.num_successful_steps} bad {s.num_unsuccessful_steps} "
          f"jac_ms {s.time_jacobian_ms:.1f} lin_ms {s.time_linear_solve_ms:.1f} total_ms {s.time_total_ms:.1f}")
    print("   costs", ["%.6e" % s.log_cost[k] for k in range(0, s.n_log, 3)])
