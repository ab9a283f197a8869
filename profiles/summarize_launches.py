"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel device time of one bench step.
usage: python profiles/summarize_launches.py launches.csv [step_index]"""
import collections
import csv
import re
import sys


def main(path, step=3):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
    idx = [i for i, n in enumerate(names) if "preprocess_kernel" in n] + [len(names)]
    s, e = idx[step], idx[step + 1]
    agg = collections.OrderedDict()
    for n, v in zip(names[s:e], vals[s:e]):
        k = re.sub(r"\(.*", "", n).replace("void ", "").replace("unnamed>::", "").replace("vtd::<", "")[:64]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f%% |" % (k, c, v / 1000, 100 * v / tot))
    print("| **total** | %d | %.1f | |" % (e - s, tot / 1000))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
