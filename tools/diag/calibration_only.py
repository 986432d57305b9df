import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lvi_exc_b200 import synth, pipeline
from lvi_exc_b200.backend import CudaBackend
seq = synth.make_sequence(synth.default_config(duration=60.0))
b = CudaBackend(0)
for rep in range(2):
    t=time.perf_counter(); res = pipeline.run_calibration(seq, b); b.synchronize()
    print('calibration wall', time.perf_counter()-t, [(s['name'], round(s['time_ms'],1)) for s in res['stages']], file=sys.stderr)
