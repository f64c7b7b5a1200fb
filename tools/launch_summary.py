"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/launch_summary.py launches.csv [--step-only]
--step-only keeps libghr's step kernels (no FP32 probes, no torch fill / copy kernels of the harness): the shares
are then shares of the fwd+bwd step, comparable with bench.py's stage_ms."""
import collections
import csv
import sys


def main(path, step_only=False):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                name = d["Kernel Name"].split("(")[0][:70]
                if step_only and ("ghr::" not in name or "probe" in name):
                    continue
                v = float(d["Metric Value"].replace(",", ""))
                unit = d["Metric Unit"]
                v = v / 1000.0 if unit in ("ns", "nsecond") else (v * 1000.0 if unit in ("ms", "msecond") else v)
                agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':72s} {'n':>4s} {'sum_us':>10s} {'mean_us':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:72s} {len(v):4d} {sum(v):10.1f} {sum(v)/len(v):9.1f} {100*sum(v)/tot:5.1f}%")
    print(f"{'TOTAL':72s} {sum(len(v) for v in agg.values()):4d} {tot:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1], "--step-only" in sys.argv)
