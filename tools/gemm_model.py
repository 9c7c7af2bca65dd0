"""Check the analytical tile-choice model of gemm_f64.cu::dispatch against a tools/gemm_tune.py table:
prints prediction error per config and the regret of the model's choice versus the measured best."""
import json
import statistics as st
import sys

SMS = 148
# b: tile edge; r: co-resident CTAs per SM; s1: us per k for one CTA alone on an SM;
# rate: flop/us per SM at full residency; t0: un-overlapped prologue + epilogue (us)
CFG = {"L": dict(b=128, r=1, s1=0.139, rate=234e3, t0=9.0),
       "M": dict(b=64, r=2, s1=0.044, rate=231e3, t0=4.5),
       "S": dict(b=32, r=5, s1=0.016, rate=220e3, t0=3.0)}


def ntiles(r, b):
    tm = -(-r["m"] // b)
    tn = -(-r["n"] // b)
    return tm * (tm + 1) // 2 if r["lower"] else tm * tn


def pred(r, c):
    p = CFG[c]
    T = ntiles(r, p["b"])
    k = r["k"]
    w = 2.0 * p["b"] * p["b"]
    t = 4.0 + p["t0"]
    if T <= p["r"] * SMS:                      # one wave, spread breadth-first over the SMs
        q = -(-T // SMS)
        t += k * max(p["s1"], q * w / p["rate"])
    elif c == "S":                             # many small CTAs de-synchronise: near-continuous refill
        t += k * -(-T // SMS) * w / p["rate"]
    else:                                      # slots free up r at a time: waves of r * SMS tiles
        waves = -(-T // (p["r"] * SMS))
        t += waves * k * p["r"] * w / p["rate"] + (p["t0"] * (waves - 1) if p["r"] == 1 else 0)
    return t


rows = [json.loads(l) for l in open(sys.argv[1])]
err = {c: [] for c in "LMS"}
tot = dict(model=0.0, best=0.0, auto=0.0)
worst = []
for r in rows:
    pr = {c: pred(r, c) for c in "LMS"}
    for c in "LMS":
        err[c].append(pr[c] / r[c])
    ch = min("LMS", key=lambda c: pr[c])
    best = min("LMS", key=lambda c: r[c])
    tot["model"] += r[ch]
    tot["best"] += r[best]
    tot["auto"] += r["auto"]
    worst.append((r[ch] / r[best], r, ch, best, pr))
for c in "LMS":
    print(c, "pred/meas median %.2f min %.2f max %.2f" % (st.median(err[c]), min(err[c]), max(err[c])))
print({k: round(v / 1e3, 1) for k, v in tot.items()})
worst.sort(key=lambda x: -x[0])
for w in worst[:12]:
    print("%.2f" % w[0], w[1], w[2], w[3], {k: round(v, 1) for k, v in w[4].items()})
