"""Critical-path breakdown of band_factor_ll_kernel from an LVI_TRACE_FACTOR dump (diagnostics).

The chain CTAs stamp, per block column j (row j*TPC of the trace): 0 loop top, 1 Dpre_j in registers, 2 D_j final (after the rank-32
update with X_{j-1}), 3 potrf + inverse done (and Ppre_j in shared memory), 4 W_j stores issued, 5 X_j computed, 6 X_j stores issued,
7 flags released."""
import sys
import numpy as np
raw = open(sys.argv[1], "rb").read()
NT, TPC, T, RB = np.frombuffer(raw[:16], np.int32)
tr = np.frombuffer(raw[16:], np.uint64).reshape(NT, TPC, 8).astype(np.float64)
t0 = tr[tr > 0].min()
tr = np.where(tr > 0, (tr - t0) / 1e3, np.nan)   # us
print(f"NT {NT} TPC {TPC} T {T} RB {RB}; last stamp {np.nanmax(tr):.1f} us")
d = tr[:, 0, :]
pub = d[:, 7]
split = int(np.nanargmin(np.diff(pub))) + 1 if NT > 2 and np.nanmin(np.diff(pub)) < 0 else NT
names = ["wait for Dpre", "rank-32 update with X_{j-1}", "potrf + inverse (+ Ppre in)", "W stores", "X = Ppre W^T", "X stores", "barrier + flags"]
for c, (lo, hi) in enumerate(((0, split), (split, NT))):
    if hi - lo < 8:
        continue
    sl = slice(lo + 3, hi - 3)
    print(f"chain {c}: columns {lo}..{hi - 1}, first flag {pub[lo]:.1f} us, last flag {pub[hi - 1]:.1f} us, "
          f"period {np.nanmean(np.diff(pub[sl])):.2f} us per column")
    for k in range(7):
        print(f"   {names[k]:<32} {np.nanmean(d[sl, k + 1] - d[sl, k]):6.2f}")
    dd = np.diff(pub[lo:hi])
    print("   period along the chain (30-column means):", " ".join(f"{np.nanmean(dd[k:k + 30]):.1f}" for k in range(0, len(dd), 30)))

# worker tasks next to the chain: row 1 = Ppre task (j,1), row 2 = tile (j+2,j); stamps: 0 fetched, 1 step k=j-2 starts, 2 step k=j-1 starts,
# 3 its inputs are in shared memory, 4 accumulation done, 5 (row 2) W_j in, 7 flagged copy stored
if TPC > 2 and T >= 2:
    p1, p2 = tr[:, 1, :], tr[:, 2, :]
    lo, hi = 6, split - 6
    j = np.arange(lo, hi)
    m = np.nanmean
    print("worker path, chain 0 (us, means):")
    print("   tile (j+2,j):  W_j stores issued -> W_j in          %6.2f" % m(p2[j, 5] - d[j, 4]))
    print("                  W_j in -> flagged copy stored         %6.2f" % m(p2[j, 7] - p2[j, 5]))
    print("                  idle before W_j (accum done -> W in)  %6.2f" % m(p2[j, 5] - p2[j, 4]))
    print("   Ppre_j:        L(j+1,j-1) stored -> inputs of k=j-1 in %6.2f" % m(p1[j, 3] - p2[j - 1, 7]))
    print("                  X_{j-1} stores issued -> inputs in      %6.2f" % m(p1[j, 3] - d[j - 1, 6]))
    print("                  step k=j-1 start -> inputs in           %6.2f" % m(p1[j, 3] - p1[j, 2]))
    print("                  inputs in -> Ppre stored                %6.2f" % m(p1[j, 7] - p1[j, 3]))
    print("   chain:         Ppre_j stored -> chain has it (stamp 3) %6.2f" % m(d[j, 3] - p1[j, 7]))
    print("                  W_{j-1} stores issued -> Ppre_j stored  %6.2f" % m(p1[j, 7] - d[j - 1, 4]))
