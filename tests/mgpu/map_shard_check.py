"""torchrun script: the map path sharded over N GPUs against the single-GPU build of the concatenated scans.

Every rank generates the same synthetic sequence, keeps the scans of its time chunk, de-skews them, and calls lvi_map_build_sharded /
lvi_associate_sharded.  Rank 0 also runs the whole sequence through a single-GPU context and compares: identical plane sets (p4, Pi, boxes,
voxel index, inlier counts), identical associated points (time stamps, raw / map coordinates, plane ids), identical counts.  Prints one
JSON line on rank 0."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lvi_exc_b200 import pipeline, synth, workload  # noqa: E402
from lvi_exc_b200.backend import CudaBackend  # noqa: E402
from lvi_exc_b200.dist import make_backend, shard_range  # noqa: E402


def local_map_path(b, mgr, seq, lo, hi, keys, sharded):
    raw = seq.scans_raw[lo:hi]
    rot = b.undistort(mgr._base(), raw, None, False)
    in_map = b.transform(rot, seq.loam_poses[lo:hi])
    rot.close()
    cloud = b.map_cloud(in_map, keys[lo:hi])
    pc = pipeline.PipelineConfig()
    if sharded:
        smap = b.build_surfel_map_sharded(cloud, pc.ndt_resolution, pc.plane_lambda_first)
        sp = b.associate_sharded(smap, in_map, raw, pc.associated_radius, pc.k_per_ring, pc.time_downsample)
    else:
        smap = b.build_surfel_map(cloud, pc.ndt_resolution, pc.plane_lambda_first)
        sp = b.associate(smap, in_map, raw, pc.associated_radius, pc.k_per_ring, pc.time_downsample)
    return smap, sp, b.last_n_all


def main():
    duration = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    b, dist, rank, world = make_backend()
    seq = synth.make_sequence(synth.default_config(duration=duration, n_landmarks=50), with_camera=False)
    mgr = workload.make_manager(seq, pipeline.PipelineConfig())
    mgr.calib.q_LtoI, mgr.calib.p_LinI = seq.gt["q_LtoI"], seq.gt["p_LinI"]
    S = len(seq.scan_times)
    keys = pipeline.check_key_scan(seq.loam_poses)
    lo, hi = shard_range(S, rank, world)
    smap, sp, n_all = local_map_path(b, mgr, seq, lo, hi, keys, sharded=True)
    out = dict(world=world, scans=S, planes=int(smap.num_planes), selected=int(len(sp)), associated=int(n_all), shard_stats=smap.shard_stats)
    ok = True
    if rank == 0:
        b1 = CudaBackend(int(os.environ.get("LOCAL_RANK", "0")))
        smap1, sp1, n_all1 = local_map_path(b1, mgr, seq, 0, S, keys, sharded=False)
        same_planes = smap1.num_planes == smap.num_planes and all(np.array_equal(smap1.planes[k], smap.planes[k]) for k in smap1.planes)
        same_points = len(sp1) == len(sp) and sp1.tobytes() == sp.tobytes()
        ok = bool(same_planes and same_points and n_all1 == n_all)
        out.update(single_gpu=dict(planes=int(smap1.num_planes), selected=int(len(sp1)), associated=int(n_all1)), same_planes=bool(same_planes),
                   same_points=bool(same_points), ok=ok)
        smap1.close(); b1.close()
        print(json.dumps(out), flush=True)
    smap.close()
    if dist is not None:
        dist.barrier()
    b.close()
    if dist is not None:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
