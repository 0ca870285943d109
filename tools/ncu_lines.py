"""Warp-stall samples of one captured launch aggregated per CUDA source line (needs -lineinfo and --import-source on):

    python tools/ncu_lines.py report.ncu-rep [launch_index] [top]
"""
import collections
import csv
import io
import subprocess
import sys


def main(path, launch=0, top=30):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip",
                          str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    agg = collections.Counter()
    text = {}
    fname, S, cur = "", None, None
    total = 0
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if "# Samples" in r:
            S = r.index("# Samples")
            continue
        if S is None or len(r) <= S:
            continue
        if r[0].isdigit():           # a CUDA source line; the SASS rows that follow belong to it
            cur = (fname, int(r[0]))
            text[cur] = r[1].strip()
            continue
        if r[S].isdigit() and cur and r[2] not in ("-", ""):
            agg[cur] += int(r[S])
            total += int(r[S])
    print(f"# {path} launch {launch}: {total} samples")
    for (f, ln), s in agg.most_common(top):
        print(f"{100.0 * s / max(total, 1):5.1f}%  {f}:{ln:<5d} {text[(f, ln)][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 30)
