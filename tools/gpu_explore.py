"""exploration: GPU vs oracle pipeline, stage by stage (not a test)"""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from lvi_exc_b200 import pipeline, synth
from lvi_exc_b200.backend import CudaBackend
from tests.oracle_backend import OracleBackend
dur = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
seq = synth.make_sequence(synth.default_config(duration=dur, n_landmarks=int(sys.argv[2]) if len(sys.argv) > 2 else 800))
cb = CudaBackend(0)
t = time.time(); og = pipeline.run_calibration(seq, cb, verbose=True); tg = time.time() - t
t = time.time(); oo = pipeline.run_calibration(seq, OracleBackend(), verbose=True); to = time.time() - t
print("gpu wall", tg, "oracle wall", to)
cg, co = og["calib"], oo["calib"]
print("assoc", og["assoc_counts"], oo["assoc_counts"], og.get("n_lm_plane"), oo.get("n_lm_plane"))
print("dq_L", pipeline.quat_angle(cg.q_LtoI, co.q_LtoI), "dp_L", np.linalg.norm(cg.p_LinI - co.p_LinI))
print("dq_C", pipeline.quat_angle(cg.q_CtoI, co.q_CtoI), "dp_C", np.linalg.norm(cg.p_CinI - co.p_CinI))
for a, b in zip(og["stages"], oo["stages"]):
    print(a["name"], a["iterations"], b["iterations"], "%.9e %.9e" % (a["final_cost"], b["final_cost"]), "ms %.1f %.1f" % (a["time_ms"], b["time_ms"]))
