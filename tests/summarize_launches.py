"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
    python tests/summarize_launches.py gpurun_out/dec_launches_b64.csv [title]"""
import collections
import csv
import re
import sys


def main(path, title=""):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        us = val / 1000 if r["Metric Unit"] in ("ns", "nsecond") else val
        key = re.sub(r"\(.*", "", r["Kernel Name"]).replace("aries::<unnamed>::", "").replace("void ", "")[:56]
        key += " grid=" + r.get("Grid Size", "")
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# {title or path}\n\n{n} launches, {tot:.0f} us in total (cold-cache, serialised: compare shares, not absolutes)\n")
    print("| kernel | launches | total us | us each | share |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {t / c:.2f} | {100 * t / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
