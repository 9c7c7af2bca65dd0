"""Summarise a tools/gemm_tune.py table: TFLOP/s per tile config and the per-shape winner."""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1])]
sel = sys.argv[2] if len(sys.argv) > 2 else "big"
tot = dict(L=0.0, M=0.0, S=0.0, auto=0.0, best=0.0)
for r in rows:
    fl = 2.0 * r["m"] * r["n"] * r["k"] * (0.5 if r["lower"] else 1)
    best = min("LMS", key=lambda c: r[c])
    for c in "LMS":
        tot[c] += r[c]
    tot["auto"] += r["auto"]
    tot["best"] += r[best]
    big = r["m"] >= 4096 and r["n"] >= 2048
    if (sel == "big" and big) or (sel == "small" and not big and r["k"] in (128, 512, 2048)) or sel == "all":
        print(r["lower"], r["m"], r["n"], r["k"], "us L %.1f M %.1f S %.1f auto %.1f" % (r["L"], r["M"], r["S"], r["auto"]),
              " TF L %.1f M %.1f S %.1f" % tuple(fl / (r[c] * 1e-6) / 1e12 for c in "LMS"), best)
print({k: round(v / 1e3, 1) for k, v in tot.items()})
