"""Critical-path breakdown of band_factor_ll_kernel from an LVI_TRACE_FACTOR dump (diagnostics)."""
import sys
import numpy as np
raw = open(sys.argv[1], "rb").read()
NT, TPC, T, RB = np.frombuffer(raw[:16], np.int32)
tr = np.frombuffer(raw[16:], np.uint64).reshape(NT, TPC, 8).astype(np.float64)
t0 = tr[tr > 0].min()
tr = np.where(tr > 0, (tr - t0) / 1e3, np.nan)   # us
print(f"NT {NT} TPC {TPC} T {T} RB {RB}; total {np.nanmax(tr):.1f} us")
d = tr[:, 0, :]          # diagonal tasks
p = tr[:, 1, :] if T >= 1 else None  # first sub-diagonal panel tile
names_d = ["fetch", "last dep seen", "last tiles loaded", "accum done", "potrf+inv done", "-", "-", "published"]
sl = slice(20, NT - 20)
print("diagonal task, mean us between marks:")
for a, b in ((0, 1), (1, 2), (2, 3), (3, 4), (4, 7)):
    print(f"  {names_d[a]:>18} -> {names_d[b]:<18} {np.nanmean(d[sl, b] - d[sl, a]):8.2f}")
if p is not None:
    names_p = ["fetch", "last dep seen", "last tiles loaded", "accum done", "W flag seen", "W loaded", "X computed", "published"]
    print("panel task (j+1,j), mean us between marks:")
    for a in range(1, 8):
        print(f"  {names_p[a-1]:>18} -> {names_p[a]:<18} {np.nanmean(p[sl, a] - p[sl, a-1]):8.2f}")
    sl1 = slice(21, NT // 2 - 39)
    sl0 = slice(20, NT // 2 - 40)
    m = np.nanmean
    print("pivot chain of the first chain, mean us per column:")
    print("  potrf + inverse (accum done -> W ready)        %.2f" % m(d[sl0, 4] - d[sl0, 3]))
    print("  W hand-off (W ready -> panel has W)            %.2f" % m(p[sl0, 5] - d[sl0, 4]))
    print("  panel solve X = P W^T                          %.2f" % m(p[sl0, 6] - p[sl0, 5]))
    print("  L(j+1,j) hand-off (X done -> next diag has it) %.2f" % m(d[sl1, 2] - p[sl0, 6]))
    print("  last rank-32 update of the next diagonal       %.2f" % m(d[sl1, 3] - d[sl1, 2]))
    print("  column period                                  %.2f" % m(d[sl1, 3] - d[sl0, 3]))
