"""one C3-shaped map pass (de-skew -> voxel build -> surfels -> association) at a reduced duration: for ncu captures and kernel-variant
sweeps.  usage: c3_pass.py [duration_s] [leaves] [passes]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from lvi_exc_b200 import pipeline, synth, workload
from lvi_exc_b200.backend import CudaBackend

dur = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
leaves = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 3
backend = CudaBackend(0)
cfg = synth.default_config(duration=dur)
times = synth.scan_times(cfg)
S, H, W = len(times), cfg.rings, cfg.az_steps
raw_d = torch.empty((S, H, W, 8), dtype=torch.float32, device="cuda:0")
for c0 in range(0, S, 200):
    n = min(200, S - c0)
    raw, _ = synth.make_lattice_scans(cfg, leaves, c0, n)
    raw_d[c0:c0 + n] = torch.from_numpy(raw.view(np.float32).reshape(n, H, W, 8)).to(raw_d.device)


class Seq:
    pass


seq = Seq()
seq.cfg, seq.scan_times, seq.gt = cfg, times, synth.gt_extrinsics()
seq.map_time, seq.end_time = float(times[0]), float(times[-1] + 1.0 / cfg.scan_rate)
seq.imu_t, seq.gyro, seq.accel = synth.make_imu(cfg)
pc = pipeline.PipelineConfig()
mgr = workload.make_manager(seq, pc)
mgr.calib.q_LtoI, mgr.calib.p_LinI = seq.gt["q_LtoI"], seq.gt["p_LinI"]
backend.kernel_timing(True)
for it in range(passes):
    if it == passes - 1:
        backend.kernel_times()
        backend.synchronize(); t0 = time.perf_counter()
    ta = time.perf_counter()
    base = mgr._base()
    tb = time.perf_counter()
    batch = backend.undistort(base, raw_d, seq.map_time, True)
    backend.synchronize(); tc = time.perf_counter()
    smap = backend.build_surfel_map(backend.map_cloud(batch), pc.ndt_resolution, pc.plane_lambda_refine)
    backend.synchronize(); td = time.perf_counter()
    sp = backend.associate(smap, batch, raw_d, pc.associated_radius, pc.k_per_ring, pc.time_downsample)
    backend.synchronize(); te = time.perf_counter()
    stage_ms = [1e3 * (tb - ta), 1e3 * (tc - tb), 1e3 * (td - tc), 1e3 * (te - td)]
    nl, npl = smap.num_leaves, smap.num_planes
    smap.close(); batch.close()
backend.synchronize()
wall = time.perf_counter() - t0
kt = backend.kernel_times()
print("host wall per stage [problem data, undistort, map + surfels, associate] ms:", [round(x, 3) for x in stage_ms])
print(f"points {S*H*W} leaves {nl} planes {npl} assoc {len(sp)} wall_ms {1e3*wall:.3f} kernels_ms {sum(v[1] for v in kt.values()):.3f}")
for k, v in sorted(kt.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  {k:44s} {v[0]:4d} launches {v[1]:8.3f} ms")
