"""profiles/traffic.json from an ncu launch list of the headline window (tools/ncu_round2.sh): measured DRAM bytes per launch
for every kernel kind of bench.py's breakdown.   python tools/make_traffic2.py profiles/r02_launches_1080p_window.csv profiles/traffic.json
Only the LAST window of the list is used (from its preprocess kernel on: the replayed plan, not the recording pass)."""
import collections, csv, json, re, sys
src, out = sys.argv[1:3]
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "msecond": 1e6, "nsecond": 1}
per = collections.OrderedDict()
for r in csv.reader(open(src, errors="replace")):
    if len(r) > 14 and r[0].isdigit():
        d = per.setdefault(int(r[0]), dict(name=r[4]))
        d[r[12]] = float(r[14].replace(",", "")) * U.get(r[13], 1)
L = list(per.values())
starts = [i for i, l in enumerate(L) if "trimask_raw" in l["name"]]
L = L[starts[-1]:]
KIND = [("conv_tc2p_kernel", "conv_tc2p"), ("conv_tc2_kernel", "conv_tc2"), ("conv_tc3_kernel", "conv_tc3"),
        ("igemm_tc_kernel", "conv_tc"), ("conv_direct_kernel", "conv_direct"), ("gca_rowstats_kernel", "tcv_gca_softmax"),
        ("gca_shift_add_kernel", "tcv_gca_shift_add"), ("gca_prep_grid_kernel", "tcv_gca_prep_grid"),
        ("gca_values_parity_kernel", "tcv_gca_values_parity"), ("gca_unfold_parity_kernel", "tcv_gca_unfold_parity"),
        ("tam_attend_kernel", "tcv_tam_attend"), ("head_conv_tanh01_kernel", "tcv_head_conv_tanh01"),
        ("avgpool2_kernel", "tcv_avgpool2"), ("space_to_depth2", "tcv_space_to_depth2"), ("pad_reflect1", "tcv_pad_reflect1")]
agg = collections.defaultdict(lambda: dict(launches=0, dram_bytes=0.0, ns=0.0))
ng = 0
for l in L:
    n = l["name"]
    kind = next((k for pat, k in KIND if pat in n), None)
    if "gemm_tc2_kernel" in n:                    # scores, aggregation, scores, aggregation
        kind = "gca_scores_gemm_tc" if ng % 2 == 0 else "gca_pv_gemm_tc"
        ng += 1
    if kind is None:
        continue
    a = agg[kind]
    a["launches"] += 1
    a["dram_bytes"] += l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0)
    a["ns"] += l.get("gpu__time_duration.sum", 0)
res = {k: dict(launches=v["launches"], dram_bytes_per_launch=v["dram_bytes"] / v["launches"], dram_bytes_total=v["dram_bytes"],
               ncu_ms_total=v["ns"] / 1e6) for k, v in agg.items()}
json.dump(res, open(out, "w"), indent=1)
for k, v in sorted(res.items(), key=lambda kv: -kv[1]["dram_bytes_total"]):
    print(f"{k:24s} launches={v['launches']:3d} dram/launch={v['dram_bytes_per_launch']/1e6:9.1f} MB total={v['dram_bytes_total']/1e9:6.2f} GB")
print("window total %.2f GB" % (sum(v["dram_bytes_total"] for v in res.values()) / 1e9))
