"""Reads an `ncu --page raw --csv` export and writes the small JSON bench.py picks `roofline.traffic` from.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > profiles/X_raw.csv
    python tools/ncu_traffic.py profiles/X_raw.csv <kernel-name substring> <min duration us> profiles/rNN_in_layer_traffic.json
"""
import csv
import json
import sys


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def main():
    path, pat, min_us, out = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    picked = []
    for r in rows[2:]:
        if pat not in r[col["Kernel Name"]]:
            continue
        dur, du = num(r[col["gpu__time_duration.sum"]]), units[col["gpu__time_duration.sum"]]
        dur_us = dur / 1e3 if du.startswith("n") else (dur * 1e3 if du.startswith("m") else dur)
        if dur_us < min_us:
            continue

        def bytes_of(name):
            v, u = num(r[col[name]]), units[col[name]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        picked.append({"duration_us": dur_us, "read": bytes_of("dram__bytes_read.sum"), "write": bytes_of("dram__bytes_write.sum"),
                       "tensor_active_pct": num(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])
                       if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in col else None})
    assert picked, "no launch of %r longer than %g us in %s" % (pat, min_us, path)
    n = len(picked)
    d = {"kernel": pat, "launches": n, "dram_bytes_per_launch": sum(p["read"] + p["write"] for p in picked) / n,
         "dram_read_bytes": sum(p["read"] for p in picked) / n, "dram_write_bytes": sum(p["write"] for p in picked) / n,
         "duration_us_under_ncu": sum(p["duration_us"] for p in picked) / n,
         "tensor_pipe_active_pct": None if picked[0]["tensor_active_pct"] is None else sum(p["tensor_active_pct"] for p in picked) / n,
         "source": "ncu --set full --clock-control none, %s" % path}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d))


if __name__ == "__main__":
    main()
