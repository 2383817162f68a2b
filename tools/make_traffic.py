"""Joins an ncu launch list of ONE recorded window (116 launches, plan order) with bench.py's
--dump-calls table and writes profiles/traffic.json: measured DRAM bytes per launch per kernel kind.

python tools/make_traffic.py gpurun_out/launches_traffic.csv gpurun_out/calls.jsonl profiles/traffic.json"""
import collections, csv, json, sys
launch_csv, calls_jsonl, out = sys.argv[1:4]
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 14 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    d = per.setdefault(int(r[0]), dict(name=r[4]))
    d[r[12]] = float(r[14]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6,
                               "usecond": 1e3, "msecond": 1e6, "nsecond": 1}.get(r[13], 1)
launches = list(per.values())
# the capture window may start mid-plan; the plan is periodic, so rotate it to start at the first kernel
start = next(i for i, L in enumerate(launches) if "trimask_raw" in L["name"])
nxt = next((i for i, L in enumerate(launches) if i > start and "trimask_raw" in L["name"]), None)
launches = launches[start:nxt] if nxt is not None else launches[start:] + launches[:start]
calls = sorted((json.loads(l) for l in open(calls_jsonl)), key=lambda c: c["idx"])
# one C-ABI call may launch several kernels (preprocess: 2-4, gca_prep: 2); align by kernel-name hints
HINT = {"tcv_conv2d": ("conv_", "igemm_tc"), "tcv_gemm_tn_tc": ("igemm_tc",), "tcv_gemm_tn_f32": ("gemm_tn_f32",),
        "tcv_preprocess_eval": ("trimask_raw", "dilate_", "preprocess_"), "tcv_gca_prep": ("gca_scales", "gca_prep"),
        "tcv_gca_values": ("gca_values",), "tcv_gca_softmax": ("gca_softmax",), "tcv_gca_fold": ("gca_fold",),
        "tcv_tam_attend": ("tam_attend",), "tcv_avgpool2": ("avgpool2",), "tcv_unknown_os8": ("unknown_os8",),
        "tcv_postprocess_eval": ("postprocess",), "tcv_pad_reflect1": ("pad_reflect1",)}
agg = collections.defaultdict(lambda: dict(launches=0, dram_bytes=0.0, ns=0.0))
li = 0
for c in calls:
    kind = c["kind"]
    fn = "tcv_conv2d" if kind.startswith("conv_") else ("tcv_gemm_tn_tc" if kind.endswith("_tc") and "gemm" in kind else
                                                        ("tcv_gemm_tn_f32" if "gemm" in kind else kind))
    hints = HINT.get(fn, (fn,))
    n = 0
    while li < len(launches) and any(h in launches[li]["name"] for h in hints):
        L = launches[li]
        a = agg[kind]
        a["launches"] += 1
        a["dram_bytes"] += L.get("dram__bytes_read.sum", 0) + L.get("dram__bytes_write.sum", 0)
        a["ns"] += L.get("gpu__time_duration.sum", 0)
        li += 1
        n += 1
        if fn in ("tcv_conv2d", "tcv_gemm_tn_tc", "tcv_gemm_tn_f32", "tcv_gca_values", "tcv_gca_softmax", "tcv_gca_fold",
                  "tcv_tam_attend", "tcv_avgpool2", "tcv_unknown_os8", "tcv_postprocess_eval", "tcv_pad_reflect1"):
            break
    assert n > 0, (c, launches[li]["name"] if li < len(launches) else None)
assert li == len(launches), (li, len(launches))
res = {k: dict(launches=v["launches"], dram_bytes_per_launch=v["dram_bytes"] / v["launches"],
               dram_bytes_total=v["dram_bytes"], ncu_ms_total=v["ns"] / 1e6) for k, v in agg.items()}
json.dump(res, open(out, "w"), indent=1)
for k, v in sorted(res.items(), key=lambda kv: -kv[1]["dram_bytes_total"]):
    print(f"{k:22s} launches={v['launches']:3d} dram/launch={v['dram_bytes_per_launch']/1e6:9.1f} MB total={v['dram_bytes_total']/1e9:6.2f} GB")
