"""profiles/traffic.json from an ncu launch list (the CSV of tools/gpu_round2_lean.sh): DRAM bytes per launch and the share
of the listed time of each of our kernel families -- bench.py reads the dominant kernel's figure as `roofline.traffic`.

    python tools/traffic_json.py profiles/launches_r2_v40.csv.gz "<profiled command, code version>" > profiles/traffic.json
"""
import collections
import csv
import gzip
import json
import re
import sys

FAMILIES = ["gemm_tf32_kernel", "gemm_bf16_kernel", "mlp_fused_kernel", "conv3_slab_kernel", "winattn_umma_kernel",
            "lnrows_reg_kernel", "lnsample", "conv3to1_kernel", "mergefuse_kernel", "metrics_kernel"]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, what):
    fh = gzip.open(path, "rt", errors="replace") if path.endswith(".gz") else open(path, errors="replace")
    rows = [r for r in csv.reader(fh) if len(r) > 14]
    col = {h: i for i, h in enumerate(rows[0])}
    per = collections.OrderedDict()
    for r in rows[1:]:
        if r[0].isdigit():
            per.setdefault(r[0], {"name": r[col["Kernel Name"]]})[r[col["Metric Name"]]] = \
                float(r[col["Metric Value"]].replace(",", "")) * SCALE.get(r[col["Metric Unit"]], 1.0)
    total = sum(k.get("gpu__time_duration.sum", 0.0) for k in per.values())
    out = {}
    for fam in FAMILIES:
        ks = [k for k in per.values() if re.search(fam, k["name"])]
        if not ks:
            continue
        out[fam] = {
            "dram_bytes_per_launch": sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in ks) / len(ks),
            "launches": len(ks),
            "us_per_launch": sum(k.get("gpu__time_duration.sum", 0.0) for k in ks) / len(ks),
            "share_of_listed_time": sum(k.get("gpu__time_duration.sum", 0.0) for k in ks) / total,
            "source": f"{path} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                      f"--clock-control none -c 4400, {what})",
        }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
