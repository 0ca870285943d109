"""sums of gpurun_out/op_breakdown.json by op-name pattern:  python tools/op_sum.py <label> [pattern ...]"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.load(open(os.path.join(ROOT, "gpurun_out", "op_breakdown.json")))
pats = sys.argv[2:] or [r"\.attn$", r"^merger", r"^refiner", r"^decoder", r"^resnet", r"\.mlp$"]
print(sys.argv[1], "total %.3f ms |" % sum(r[1] for r in d),
      " ".join("%s %.3f" % (p, sum(r[1] for r in d if re.search(p, r[0]))) for p in pats))
