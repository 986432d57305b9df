"""Generates tests/golden/velodyne_251370668_20k.npy from the reference's own real-data fixture
/root/reference/src/ndt_omp/data/251370668.pcd (binary PCD v0.7, fields x y z intensity float32): every 3rd point of the
first 60,000 -> 20,000 x 3 float32.  Run in the build container (the reference tree is not present on the GPU box)."""
import sys
from pathlib import Path

import numpy as np

src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/ndt_omp/data/251370668.pcd")
raw = src.read_bytes()
hdr_end = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
header = raw[:hdr_end].decode()
npts = int([l for l in header.splitlines() if l.startswith("POINTS")][0].split()[1])
pts = np.frombuffer(raw[hdr_end:hdr_end + npts * 16], dtype=np.float32).reshape(npts, 4)
out = np.ascontiguousarray(pts[:60000:3, :3])
np.save(Path(__file__).resolve().parent / "velodyne_251370668_20k.npy", out)
print(out.shape, out.min(axis=0), out.max(axis=0))
