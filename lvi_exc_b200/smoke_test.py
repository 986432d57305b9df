"""One small invocation of the whole hot path on cuda:0, checked against the CPU oracle (test infrastructure).
Called by __graft_entry__.smoke()."""
from __future__ import annotations

import numpy as np

from . import pipeline, synth, workload
from .backend import CudaBackend, CudaProblem


def run(verbose: bool = True) -> None:
    from tests.oracle_backend import OracleBackend   # the checker
    from tests import oracle_binding as ob

    cfg = synth.default_config(duration=2.0, n_landmarks=200)
    seq = synth.make_sequence(cfg)
    cb, orc = CudaBackend(0), OracleBackend()
    # ---- map path: undistort -> transform -> voxel covariance -> surfels -> association
    mgr = pipeline.TrajectoryManager(pipeline.CameraIntrinsics(), seq.map_time, seq.end_time, 0.02, 0.2)
    init = pipeline.perturbed_initial_extrinsics(seq.gt)
    mgr.calib.q_LtoI, mgr.calib.p_LinI = init["q_LtoI"], init["p_LinI"]
    scans_map_g = cb.transform(cb.undistort(mgr._base(), seq.scans_raw, None, False), seq.loam_poses)
    scans_map_o = orc.transform(orc.undistort(mgr._base(), seq.scans_raw, None, False), seq.loam_poses)
    sm = scans_map_g.cpu().numpy()
    fin = np.isfinite(scans_map_o[..., 0])
    assert np.array_equal(fin, np.isfinite(sm[..., 0]))
    assert np.abs(sm[fin][:, :3] - scans_map_o[fin][:, :3]).max() < 2e-5, "undistort/transform mismatch"
    # feed both sides the SAME cloud so that indices must agree bit for bit
    cloud = scans_map_o.reshape(-1, 8)
    gmap = cb.build_surfel_map(cloud, 0.5, 0.6)
    omap = orc.build_surfel_map(cloud, 0.5, 0.6)
    gl, ol = gmap.export_leaves(), omap.vmap.export()
    assert np.array_equal(gl["keys"], ol["keys"]) and np.array_equal(gl["nr_points"], ol["nr_points"]), "voxel keys/counts differ"
    assert np.array_equal(gl["point_index"], ol["point_index"]), "leaf point lists differ"
    assert np.allclose(gl["cov"], ol["cov"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(gmap.planes["leaf_key"], omap.planes["leaf_key"]), "surfel plane sets differ"
    assert np.allclose(gmap.planes["p4"], omap.planes["p4"], rtol=0, atol=1e-6)
    sp_g = cb.associate(gmap, scans_map_o, seq.scans_raw, 0.05, 2, 10)
    sp_o = orc.associate(omap, scans_map_o, seq.scans_raw, 0.05, 2, 10)
    assert len(sp_g) == len(sp_o) and np.array_equal(sp_g["plane_id"], sp_o["plane_id"]) and np.array_equal(sp_g["timestamp"], sp_o["timestamp"]), \
        "association differs from the oracle"
    # ---- solve path: S0 (gyro-only SO3 fit) then S1 (IMU + surfel) against the oracle from identical inputs
    mgr.feed_imu(seq.imu_t, seq.gyro, seq.accel)
    pd_g, pd_o = mgr.problem_so3(), mgr.problem_so3()
    s_g, s_o = cb.solve(pd_g, 30), orc.solve(pd_o, 30)
    s0_cost = s_g.final_cost
    assert abs(s_g.final_cost - s_o.final_cost) <= 1e-6 * max(1.0, s_o.final_cost), (s_g.final_cost, s_o.final_cost)
    assert np.abs(pd_g.so3_knots - pd_o.so3_knots).max() < 1e-6
    # S1 from a state near the optimum (control points sampled from the ground truth)
    mgr1 = workload.make_manager(seq)
    pd_g = mgr1.problem_surfel(omap.planes_Pi, sp_o, seq.map_time)
    pd_o = mgr1.problem_surfel(omap.planes_Pi, sp_o, seq.map_time)
    ev_g, ev_o = CudaProblem(cb, pd_g).evaluate(gradient=False), ob.OracleProblem(pd_o).evaluate(gradient=False)
    assert abs(ev_g["cost"] - ev_o["cost"]) <= 1e-9 * ev_o["cost"], (ev_g["cost"], ev_o["cost"])
    assert np.abs(ev_g["residuals"] - ev_o["residuals"]).max() <= 1e-7 * max(1.0, np.abs(ev_o["residuals"]).max())
    s_g, s_o = cb.solve(pd_g, 5), orc.solve(pd_o, 5)
    rel = abs(s_g.final_cost - s_o.final_cost) / s_o.final_cost
    assert s_g.num_iterations == s_o.num_iterations and rel < 1e-6, (s_g.final_cost, s_o.final_cost)
    assert pipeline.quat_angle(pd_g.lidar_q, pd_o.lidar_q) < 1e-4 and np.abs(pd_g.lidar_p - pd_o.lidar_p).max() < 1e-3
    if verbose:
        print(f"smoke ok: leaves {gmap.num_leaves}, planes {gmap.num_planes}, surfel points {len(sp_g)}, "
              f"S0 cost {s0_cost:.6e}, S1(5 it) cost gpu {s_g.final_cost:.6e} oracle {s_o.final_cost:.6e}, "
              f"kernel launches {cb.launches}")
    cb.close()
