"""Per-task latency statistics from an LVI_TRACE_FACTOR dump (diagnostics; complements analyze_factor_trace.py)."""
import sys
import numpy as np
raw = open(sys.argv[1], 'rb').read()
NT, TPC, T, RB = np.frombuffer(raw[:16], np.int32)
tr = np.frombuffer(raw[16:], np.uint64).reshape(NT, TPC, 8).astype(np.float64)
t0 = tr[tr > 1e12].min(); tr = np.where(tr > 1e12, (tr - t0) / 1e3, np.nan)
d = tr[:, 0, :]; p0 = tr[:, TPC - 1, :]; p1 = tr[:, 1, :]
split = NT // 2
per = np.diff(d[:split, 7])
pc = lambda x: np.round(np.nanpercentile(x, [5, 25, 50, 75, 90, 95]), 2)
print("percentiles 5 25 50 75 90 95")
print("column period", pc(per[5:split - 30]), "mean %.2f" % np.nanmean(per[5:split - 30]))
j = np.arange(30, split - 30)
print("chain: wait for Dpre", pc(d[j, 6] - d[j, 5]), " potrf (+ wait for Ppre)", pc(d[j, 3] - d[j, 2]))
for name, p in (("Dpre task", p0), ("Ppre task", p1), ("tile s=5", tr[:, 5, :]), ("tile s=15", tr[:, 15, :])):
    print(name, "fetched -> old updates done", pc(p[j, 2] - p[j, 0]), "| wait for fresh inputs", pc(p[j, 3] - p[j, 2]), "| last update", pc(p[j, 4] - p[j, 3]))
print("Dpre_j stored relative to the chain needing it", pc(p0[j, 7] - d[j - 1, 5]))
print("Ppre stored relative to potrf start + 2.8 us ", pc(p1[j, 7] - d[j, 2] - 2.8))
fl = d[:split, 7]; fl = fl[~np.isnan(fl)]
print("look-ahead (columns) when the Dpre task is fetched", pc([jj - np.searchsorted(fl, p0[jj, 0]) for jj in j]))
