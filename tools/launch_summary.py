"""Aggregates an ncu launch list (csv from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv --log-file X`) per kernel name into a markdown table:  python tools/launch_summary.py X.csv "title" [first_kernel_regex] > out.md
With a first-kernel pattern only the LAST complete period starting at that kernel is kept (one window / one step)."""
import collections, csv, re, sys
path, title = sys.argv[1], sys.argv[2]
first = sys.argv[3] if len(sys.argv) > 3 else None
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
per = collections.OrderedDict()
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}
for r in rows:
    d = per.setdefault(int(r[0]), dict(name=r[4]))
    d[r[12]] = float(r[14].replace(",", "")) * U.get(r[13], 1)
L = list(per.values())
if first:
    idx = [i for i, l in enumerate(L) if re.search(first, l["name"])]
    if len(idx) >= 2:
        L = L[idx[-2]:idx[-1]]
    elif idx:
        L = L[idx[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for l in L:
    n = re.sub(r"\(.*", "", l["name"])
    n = re.sub(r"^void\s+", "", n)
    n = n.replace("tcv::", "")
    a = agg[n]
    a[0] += 1
    a[1] += l.get("gpu__time_duration.sum", 0.0)
    a[2] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"# {title}\n")
print("Per-launch times under ncu are cold-cache and serialised: compare SHARES with the CUDA-event timings, not absolutes.\n")
print("| kernel | launches | total us | share | DRAM GB (read+write) |\n|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e9:.3f} |")
print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot:.1f} | 100% | {sum(a[2] for a in agg.values()) / 1e9:.3f} |")
