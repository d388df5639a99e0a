"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.

usage: python tools/summarize_launches.py gpurun_out/r01b_launches.csv [launches_per_step] > profiles/<name>.txt

ncu's per-launch times are cold-cache and serialised: compare SHARES, not absolute times, with bench.py's numbers."""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)                      # drop the argument list
    name = name.replace("void ", "").replace("__nv_bfloat16", "bf16")
    name = re.sub(r"at::native::(\(anonymous namespace\)::)?", "aten::", name)
    return name[:110]


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((short(r["Kernel Name"]), v * scale))
    total = sum(t for _, t in rows)
    agg = OrderedDict()
    for k, t in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    print("# %s: %d launches, %.2f ms total (ncu-serialised, cold cache)" % (path, len(rows), total / 1e3))
    ours = sum(t for k, t in rows if k.startswith("rb::"))
    print("# kernels of libradtts_b200.so (rb::*): %.1f %% of the profiled time, %d launches"
          % (100.0 * ours / max(total, 1e-9), sum(1 for k, _ in rows if k.startswith("rb::"))))
    print("%-112s %8s %11s %9s %7s" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print("%-112s %8d %11.1f %9.2f %6.2f%%" % (k, n, t, t / n, 100.0 * t / total))


if __name__ == "__main__":
    main()
