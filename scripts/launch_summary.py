"""Summarise an ncu launch list (gpu__time_duration.sum per launch, --csv) into a per-kernel table
and the launch sequence of the last complete step.
usage: python scripts/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import io
import sys

rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
recs = list(csv.DictReader(io.StringIO("".join(rows))))


def short(n):
    n = n.replace("void ", "").replace("fbkst::", "")
    return n.split("(")[0][:48]


starts = [i for i, x in enumerate(recs) if "cmvn_stats" in x["Kernel Name"]]
n_steps = max(1, len(starts) - 1)
agg = collections.OrderedDict()
for x in recs[starts[0]:starts[-1]] if len(starts) > 1 else recs:
    try:
        v = float(x["Metric Value"].replace(",", "")) / 1e3
    except ValueError:
        continue
    agg.setdefault(short(x["Kernel Name"]), []).append(v)
tot = sum(sum(v) for v in agg.values())
print("## per-kernel totals over %d complete steps (ncu gpu__time_duration, cold cache, serialised)\n" % n_steps)
print("| kernel | launches/step | us/step | share | avg us | min | max |")
print("|---|---|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("| %s | %.1f | %.1f | %.1f%% | %.2f | %.2f | %.2f |" % (
        k, len(v) / n_steps, sum(v) / n_steps, 100 * sum(v) / tot, sum(v) / len(v), min(v), max(v)))
print("\nsum of kernel time per step: %.1f us\n" % (tot / n_steps))
if len(starts) > 1:
    print("## launch sequence of the last complete step\n")
    print("| # | kernel | grid | block | us |")
    print("|---|---|---|---|---|")
    for i, x in enumerate(recs[starts[-2]:starts[-1]]):
        print("| %d | %s | %s | %s | %.2f |" % (i, short(x["Kernel Name"]), x["Grid Size"], x["Block Size"],
                                              float(x["Metric Value"]) / 1e3))
