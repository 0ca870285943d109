"""Summarises an .ncu-rep (ncu --set full) into a few lines per captured launch: duration, DRAM bytes, tensor-pipe
and memory utilisation, occupancy.  Runs in the build container (no GPU needed):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rN_<kernel>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col.get("Kernel Name", 4)
    print(f"# {path}: {len(data)} captured launch(es)")
    for r in data:
        print(f"\n== {r[name_i][:110]}")
        for key, label in KEYS:
            if key in col and r[col[key]] != "":
                print(f"  {label:28s} {r[col[key]]:>16s} {units[col[key]]}   [{key}]")
        # any tensor-pipe metrics present in this ncu version
        for h, i in col.items():
            if "pipe_tensor" in h and "pct" in h and r[i] not in ("", "0"):
                print(f"  {h:60s} {r[i]:>10s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
