"""Where the near-diagonal worker tasks stand relative to the chain (LVI_TRACE_FACTOR dump): for tile (j + s, j), times relative to the moment the
chain issues the flagged copy of W_j (chain stamp 4 of column j)."""
import sys
import numpy as np
raw = open(sys.argv[1], 'rb').read()
NT, TPC, T, RB = np.frombuffer(raw[:16], np.int32)
tr = np.frombuffer(raw[16:], np.uint64).reshape(NT, TPC, 8).astype(np.float64)
t0 = tr[tr > 1e12].min(); tr = np.where(tr > 1e12, (tr - t0) / 1e3, np.nan)
d = tr[:, 0, :]
split = NT // 2
j = np.arange(30, split - 30)
pc = lambda x: np.round(np.nanpercentile(x, [5, 25, 50, 75, 90, 95]), 2)
wpub = d[j, 4]
print("percentiles 5 25 50 75 90 95; all relative to 'W_j flagged copy issued' of the task's own column j")
for s in (2, 3, 4, 8, 16, 22):
    p = tr[:, s, :]
    print(f"s={s:2d}: fetched {pc(p[j,0]-wpub)} old done {pc(p[j,2]-wpub)} fresh in {pc(p[j,3]-wpub)} last upd done {pc(p[j,4]-wpub)} W in {pc(p[j,5]-wpub)} stored {pc(p[j,7]-wpub)}")
print("chain: D_j final rel. W_{j-1} issued", pc(d[j, 2] - d[j - 1, 4]), " W_j issued rel. D_j final", pc(d[j, 4] - d[j, 2]))
hl = tr[:, T + 1, :]
print("helpers: T tile in rel. D_j final", pc(hl[j, 2] - d[j, 2]), " rel. stored time of tile (j+1, j-1)", pc(hl[j, 2] - tr[j - 1, 2, 7]))
