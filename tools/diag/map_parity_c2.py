"""diagnostic: the map path (voxel build, surfels, association) of the CUDA library against the CPU oracle at C2 size, step by step"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from lvi_exc_b200 import synth, workload, pipeline
from lvi_exc_b200.backend import CudaBackend
from tests import oracle_binding as ob

dur = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
degenerate = int(sys.argv[2]) if len(sys.argv) > 2 else 0
seq = synth.make_sequence(synth.default_config(duration=dur, degenerate=degenerate), with_camera=False)
b = CudaBackend(0)
mgr = workload.make_manager(seq)
mgr.calib.q_LtoI, mgr.calib.p_LinI = seq.gt["q_LtoI"], seq.gt["p_LinI"]
batch = b.undistort(mgr._base(), seq.scans_raw, seq.map_time, True)
scans_map = batch.numpy()
cloud = scans_map.reshape(-1, 8)
for lam in (0.6, 0.7):
    gmap = b.build_surfel_map(b.map_cloud(batch), 0.5, lam)
    gl = gmap.export_leaves()
    ov = ob.OracleVoxelMap(cloud, 0.5); osf = ob.OracleSurfels(ov, lam)
    ol = ov.export()
    print("lam", lam, "leaves", gmap.num_leaves, ov.num_leaves, "planes", gmap.num_planes, osf.count)
    print(" keys", np.array_equal(gl["keys"], ol["keys"]), "npts", np.array_equal(gl["nr_points"], ol["nr_points"]),
          "start", np.array_equal(gl["leaf_start"], ol["leaf_start"]), "pidx", np.array_equal(gl["point_index"], ol["point_index"]))
    if not np.array_equal(gl["point_index"], ol["point_index"]):
        bad = np.nonzero(gl["point_index"] != ol["point_index"])[0]
        print("  first point_index mismatches", bad[:10], gl["point_index"][bad[:10]], ol["point_index"][bad[:10]])
    gp, op = gmap.planes, osf.export()
    same_set = np.array_equal(gp["leaf_key"], op["leaf_key"])
    print(" plane set", same_set)
    if not same_set:
        sg, so = set(gp["leaf_key"].tolist()), set(op["leaf_key"].tolist())
        print("  only gpu", sorted(sg - so)[:10], "only oracle", sorted(so - sg)[:10])
        for k in sorted(sg ^ so)[:5]:
            i = int(np.nonzero(ol["keys"] == k)[0][0])
            ev_o = np.sort(ol["evals"][i])[::-1]; ev_g = np.sort(gl["evals"][i])[::-1]
            print("   leaf", k, "n", ol["nr_points"][i], "p oracle", 2 * (ev_o[1] - ev_o[2]) / ev_o.sum(), "p gpu", 2 * (ev_g[1] - ev_g[2]) / ev_g.sum())
    else:
        print(" ninl", np.array_equal(gp["n_inliers"], op["n_inliers"]), "box", np.array_equal(gp["box_min"], op["box_min"]) and np.array_equal(gp["box_max"], op["box_max"]),
              "p4 maxdiff", np.abs(gp["p4"] - op["p4"]).max())
    for unf in ("", "1"):
        if unf: os.environ["LVI_ASSOC_UNFUSED"] = "1"
        else: os.environ.pop("LVI_ASSOC_UNFUSED", None)
        sp_g = b.associate(gmap, batch, seq.scans_raw, 0.05, 2, 1)
        sp_o, n_all = osf.associate(scans_map, seq.scans_raw, 0.05, 2, 1, mode=1)
        ok = len(sp_g) == len(sp_o) and all(np.array_equal(sp_g[f], sp_o[f]) for f in ("timestamp", "point", "point_in_map", "plane_id"))
        print(" assoc unfused=%r" % unf, len(sp_g), len(sp_o), ok)
        if not ok:
            tg = set(zip(sp_g["timestamp"].tolist(), sp_g["plane_id"].tolist())); to = set(zip(sp_o["timestamp"].tolist(), sp_o["plane_id"].tolist()))
            print("  only gpu", sorted(tg - to)[:6], "only oracle", sorted(to - tg)[:6])
    gmap.close()
