"""SASS digest of the shipped library: per kernel, how many tcgen05 MMA (UTC*MMA), TMEM load (LDTM), TMA load / store
(UTMALDG / UTMASTG), legacy HMMA and cp.async (LDGSTS) instructions it contains -- the evidence B200_PROFILING.md asks for.
    python tools/sass_digest.py [lib.so] > profiles/r02_sass_digest.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tcvom_b200", "lib", "libtcvom_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = collections.OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("UTC*MMA.2CTA", r"\bUTC[A-Z]*MMA\.2CTA"), ("LDTM", r"\bLDTM"),
                               ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.2CTA", r"\bUTMALDG[.0-9A-Z]*\.2CTA"),
                               ("UTMASTG", r"\bUTMASTG"), ("UTCBAR", r"\bUTCBAR"), ("HMMA", r"\bHMMA"), ("LDGSTS", r"\bLDGSTS"),
                               ("LDG.256", r"\bLDG\.[.A-Z0-9]*256"), ("STG.256", r"\bSTG\.[.A-Z0-9]*256")])
rows, cur, cnt = [], None, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, cnt))
        cur, cnt = m.group(1), collections.Counter()
        continue
    if cur:
        for k, p in PAT.items():
            if re.search(p, line):
                cnt[k] += 1
if cur:
    rows.append((cur, cnt))
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "").replace("tcv::", "")
    except Exception:
        return n
tot = collections.Counter()
print("# SASS digest of tcvom_b200/lib/libtcvom_b200.so (sm_100a; `cuobjdump -sass`, round 2)\n")
print("Kernels that contain tensor-core / TMA instructions (all other kernels: plain CUDA-core code, no HMMA anywhere).\n")
print("| kernel | " + " | ".join(PAT) + " |\n|---|" + "---|" * len(PAT))
for n, c in sorted(rows, key=lambda r: -(r[1]["UTC*MMA"] * 1000 + r[1]["UTMALDG"] + r[1]["UTMASTG"])):
    tot.update(c)
    if c["UTC*MMA"] or c["UTMALDG"] or c["UTMASTG"] or c["HMMA"]:
        print(f"| `{demangle(n)}` | " + " | ".join(str(c[k]) for k in PAT) + " |")
print("| **whole library** (%d kernels) | " % len(rows) + " | ".join(str(tot[k]) for k in PAT) + " |")
