"""Summarises a tools/graph_timeline.py JSON: per kernel name the launches, the summed duration, the time it ran ALONE
(no other kernel executing: that is what it costs the step) and its share of time it ran next to other kernels.

    python tools/timeline_summary.py gpurun_out/rNN_timeline.json > profiles/rNN_step_timeline.md
"""
import collections
import json
import re
import sys


def short(n):
    n = re.sub(r"^void ", "", n).replace("rb::", "").replace("at::native::", "aten:")
    return n[:64]


rows = json.load(open(sys.argv[1]))
# the replay's body starts after the launch gap that follows the static-batch copies
t_start = rows[0][0]
for i in range(1, len(rows)):
    if rows[i][0] - (rows[i - 1][0] + rows[i - 1][1]) > 500:
        t_start = rows[i][0]
        break
rows = [r for r in rows if r[0] >= t_start]
ev = []
for i, (ts, d, s, n) in enumerate(rows):
    ev.append((ts, 1, i))
    ev.append((ts + d, -1, i))
ev.sort()
active, last = set(), None
excl, shared = collections.defaultdict(float), collections.defaultdict(float)
idle = 0.0
for t, k, i in ev:
    if last is not None and t > last:
        if not active:
            idle += t - last
        elif len(active) == 1:
            excl[next(iter(active))] += t - last
        else:
            for j in active:
                shared[j] += (t - last) / len(active)
    if k == 1:
        active.add(i)
    else:
        active.discard(i)
    last = t
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, (ts, d, s, n) in enumerate(rows):
    a = agg[short(n)]
    a[0] += 1
    a[1] += d
    a[2] += excl[i]
    a[3] += shared[i]
span = max(r[0] + r[1] for r in rows) - t_start
print("# One replay of the captured train step (cfg2, 32 x <=800 frames, bf16): kernel timeline from CUPTI")
print()
print("Source: `%s` (`tools/graph_timeline.py`; taken under the torch profiler, so ~2 %% slower than the bench)." % sys.argv[1])
print("%d kernels, span %.2f ms, sum of durations %.2f ms, GPU idle inside the span %.2f ms." % (
    len(rows), span / 1e3, sum(r[1] for r in rows) / 1e3, idle / 1e3))
print("`alone` = time during which this kernel was the only one executing (what it costs the step); `shared` = its")
print("share of the time it ran concurrently with other kernels (time divided by the number of kernels running).")
print()
print("| kernel | launches | sum us | alone us | shared us |")
print("|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1][2] + kv[1][3]))[:48]:
    print("| `%s` | %d | %.0f | %.0f | %.0f |" % (k.replace("|", "/"), a[0], a[1], a[2], a[3]))
print()
print("total alone %.2f ms, shared %.2f ms" % (sum(a[2] for a in agg.values()) / 1e3, sum(a[3] for a in agg.values()) / 1e3))
