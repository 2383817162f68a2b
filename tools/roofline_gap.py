"""Per-launch roofline gap of one window from bench.py --dump-calls:  python tools/roofline_gap.py calls.jsonl > out.md
ideal(launch) = max(3 x algorithmic FLOP / sustained bf16 peak  [bf16x3 split: every MMA is issued three times],
                    algorithmic bytes / measured HBM copy bandwidth);  lost = measured - ideal."""
import collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
    dict(bf16_tflops_sustained=1373.9, hbm_gbs=6545.9)
TF, HBM = pk["bf16_tflops_sustained"], pk["hbm_gbs"]
TENSOR = ("conv_tc", "conv_tc2", "conv_tc3", "gca_scores_gemm_tc", "gca_pv_gemm_tc")
rows = [json.loads(l) for l in open(sys.argv[1])]
out = []
for r in rows:
    fl, by = r.get("flops", 0), r.get("bytes", 0)
    t_mma = 3 * fl / (TF * 1e12) * 1e3 if r["kind"] in TENSOR else 0.0
    t_hbm = by / (HBM * 1e9) * 1e3
    ideal = max(t_mma, t_hbm)
    out.append(dict(lost=r["ms"] - ideal, ms=r["ms"], ideal=ideal, mma=t_mma, hbm=t_hbm, kind=r["kind"],
                    shape=r.get("shape", ""), layer=r.get("layer", "")))
tot, ide = sum(o["ms"] for o in out), sum(o["ideal"] for o in out)
print(f"# Roofline gap per launch ({os.path.basename(sys.argv[1])})\n")
print(f"Sum of per-launch times (eager replay, CUDA events) {tot:.2f} ms; sum of per-launch ideals {ide:.2f} ms "
      f"(tensor peak {TF} TFLOP/s sustained with the 3x split, HBM {HBM} GB/s).\n")
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
for o in out:
    a = agg[o["kind"]]
    a[0] += o["ms"]; a[1] += o["ideal"]; a[2] += 1
print("| kernel kind | launches | ms | ideal ms | lost ms |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1][0] - kv[1][1])):
    print(f"| `{k}` | {a[2]} | {a[0]:.3f} | {a[1]:.3f} | {a[0] - a[1]:.3f} |")
print("\n| lost ms | ms | ideal (mma / hbm) | kind | shape | layer |\n|---|---|---|---|---|---|")
for o in sorted(out, key=lambda o: -o["lost"])[:30]:
    print(f"| {o['lost']:.3f} | {o['ms']:.3f} | {o['ideal']:.3f} ({o['mma']:.3f} / {o['hbm']:.3f}) | `{o['kind']}` | {o['shape']} | {o['layer']} |")
