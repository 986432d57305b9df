"""Critical-path breakdown of band_factor_ll_kernel from an LVI_TRACE_FACTOR dump (diagnostics).

The chain CTAs stamp, per block column j (row j*TPC of the trace): 2 D_j final in shared memory (start of the diagonal-block phase),
3 Cholesky + inverse done (and Ppre_j complete in shared memory), 4 flagged copy of W_j issued, 5 X_j computed (in shared memory),
6 Dpre_{j+1} available to the warps that build the next diagonal tile, 7 D_{j+1} built and the stores of column j issued."""
import sys
import numpy as np
raw = open(sys.argv[1], "rb").read()
NT, TPC, T, RB = np.frombuffer(raw[:16], np.int32)
tr = np.frombuffer(raw[16:], np.uint64).reshape(NT, TPC, 8).astype(np.float64)
t0 = tr[tr > 1e12].min()
tr = np.where(tr > 1e12, (tr - t0) / 1e3, np.nan)   # us
print(f"NT {NT} TPC {TPC} T {T} RB {RB}; last stamp {np.nanmax(tr):.1f} us")
d = tr[:, 0, :]
pub = d[:, 7]
split = int(np.nanargmin(np.diff(pub))) + 1 if NT > 2 and np.nanmin(np.diff(pub)) < 0 else NT
names = [None, None, "Cholesky + inverse (+ Ppre complete)", "flagged W stores", "X = Ppre W^T", "wait for Dpre (next column)", "rank-32 update -> D_{j+1} || stores"]
for c, (lo, hi) in enumerate(((0, split), (split, NT))):
    if hi - lo < 8:
        continue
    sl = slice(lo + 3, hi - 3)
    print(f"chain {c}: columns {lo}..{hi - 1}, first flag {pub[lo]:.1f} us, last flag {pub[hi - 1]:.1f} us, "
          f"period {np.nanmean(np.diff(pub[sl])):.2f} us per column")
    for k in range(2, 7):
        print(f"   {names[k]:<40} {np.nanmean(d[sl, k + 1] - d[sl, k]):6.2f}")
    dd = np.diff(pub[lo:hi])
    print("   period along the chain (30-column means):", " ".join(f"{np.nanmean(dd[k:k + 30]):.1f}" for k in range(0, len(dd), 30)))

