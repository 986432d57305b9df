"""Per-kernel totals of an ncu `--metrics gpu__time_duration.sum --csv` launch list (diagnostics)."""
import collections
import csv
import sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(r for r in rows if "Kernel Name" in r)
start = rows.index(hdr) + 1
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # launches to skip at the front (map path + warm-up)
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows[start + skip:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    n = r[ki][:64]
    tot[n] = tot.get(n, 0) + v
    cnt[n] += 1
all_ = sum(tot.values())
for n, v in sorted(tot.items(), key=lambda x: -x[1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
    print(f"{v / 1e3:10.1f} us {100 * v / all_:5.1f}%  {cnt[n]:4d} launches  {v / cnt[n] / 1e3:9.1f} us avg  {n}")
