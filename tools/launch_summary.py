"""Per-kernel summary of an ncu launch list (CSV written by
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file X ...):
launches, total time, share of the listed time, DRAM bytes read / written.

    python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rN_launches_summary.txt
"""
import collections
import csv
import re
import sys


def main(path, what=""):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[1:]:
        if not r[0].isdigit():
            continue
        per.setdefault(r[0], {"name": r[col["Kernel Name"]]})[r[col["Metric Name"]]] = (r[col["Metric Unit"]], float(r[col["Metric Value"]].replace(",", "")))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for k in per.values():
        nm = re.sub(r"\(.*", "", k["name"])
        nm = nm.replace("svx::<unnamed>::", "").replace("void ", "")
        a = agg[nm[-64:]]
        a[0] += 1
        for key, idx in (("gpu__time_duration.sum", 1), ("dram__bytes_read.sum", 2), ("dram__bytes_write.sum", 3)):
            if key in k:
                u, v = k[key]
                a[idx] += v * scale.get(u, 1.0)
    tot = sum(a[1] for a in agg.values())
    print(f"# {what} ({len(per)} launches listed)")
    print("# kernel, launches, total us, share, dram read MB, dram write MB")
    for nm, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{nm:64s} {a[0]:6d} {a[1]:12.1f} {100 * a[1] / tot:6.1f}% {a[2] / 1e6:10.1f} {a[3] / 1e6:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
