"""The map path sharded over the GPUs of a node (lvi_map_build_sharded / lvi_associate_sharded, SURVEY §8e collectives 1-2)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from lvi_exc_b200 import pipeline, synth, workload

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_sharded_entry_points_at_world_1_equal_the_single_gpu_calls(cuda_backend):
    b = cuda_backend
    seq = synth.make_sequence(synth.default_config(duration=2.0, n_landmarks=50), with_camera=False)
    mgr = workload.make_manager(seq, pipeline.PipelineConfig())
    mgr.calib.q_LtoI, mgr.calib.p_LinI = seq.gt["q_LtoI"], seq.gt["p_LinI"]
    rot = b.undistort(mgr._base(), seq.scans_raw, None, False)
    in_map = b.transform(rot, seq.loam_poses)
    keys = pipeline.check_key_scan(seq.loam_poses)
    a = b.build_surfel_map(b.map_cloud(in_map, keys), 0.5, 0.6)
    s = b.build_surfel_map_sharded(b.map_cloud(in_map, keys), 0.5, 0.6)
    assert a.num_planes == s.num_planes > 10 and all(np.array_equal(a.planes[k], s.planes[k]) for k in a.planes)
    pa = b.associate(a, in_map, seq.scans_raw, 0.05, 2, 10)
    ps = b.associate_sharded(s, in_map, seq.scans_raw, 0.05, 2, 10)
    assert len(pa) > 100 and pa.tobytes() == ps.tobytes()
    a.close(); s.close(); in_map.close(); rot.close()


@pytest.mark.gpu
def test_map_path_sharded_over_two_gpus_equals_the_single_gpu_build():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
                        str(ROOT / "tests" / "mgpu" / "map_shard_check.py"), "4.0"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"] and out["same_planes"] and out["same_points"] and out["planes"] > 10 and out["selected"] > 100
