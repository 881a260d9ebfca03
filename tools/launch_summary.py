"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file in.csv ...`):

    python tools/launch_summary.py in.csv out.csv [launches_per_step]

The list of `bench.py --profile-mode` holds a warm-up step followed by the profiled step; with launches_per_step given,
only the LAST that many launches (the second step) are summarised."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    src, dst = sys.argv[1], sys.argv[2]
    last = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, data = rows[0], rows[1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    if last:
        data = data[-last:]
    agg = OrderedDict()
    for r in data:
        name = re.sub(r"\(.*$", "", r[ik]).strip()
        t = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[iu], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "ms", "share"])
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, round(t, 4), round(t / tot, 4)])
    print(f"{dst}: {len(data)} launches, {tot:.3f} ms")


if __name__ == "__main__":
    main()
