"""Groups a per-op breakdown (profiles/r1_op_breakdown_v*.json) by layer kind: python tools/op_groups.py file.json [top]"""
import collections
import json
import sys


def key(n):
    p = n.split('.')
    if p[0] == 'swin' and p[1].isdigit():
        return 'swin.s' + p[1] + '.' + p[-1]
    if p[0] == 'resnet':
        return 'resnet.' + p[1] + ('.' + p[-1] if len(p) > 2 else '')
    if p[0] in ('decoder', 'refiner'):
        return p[0] + '.' + p[1].split('+')[0] + ('' if len(p) < 3 or p[2].startswith('p') else '.' + p[2])
    return '.'.join(p[:2])


ops = json.load(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
g = collections.OrderedDict()
PEAK_TF, PEAK_GBS = 689.55, 6550.7
ops = [tuple(o) + (0.0,) * (4 - len(o)) for o in ops]
for n, ms, fl, nb in ops:
    a = g.setdefault(key(n), [0, 0, 0, 0.0])
    a[0] += ms
    a[1] += fl
    a[2] += 1
    a[3] += max(fl / (PEAK_TF * 1e9), nb / (PEAK_GBS * 1e6))
tot = sum(v[0] for v in g.values())
print(f"total {tot:.3f} ms over {len(ops)} ops")
for k, v in sorted(g.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k:32s} {v[0]:7.3f} ms  n={v[2]:3d}  {v[1] / 1e9:9.1f} GF  {v[1] / v[0] / 1e9 if v[0] else 0:7.1f} TF/s"
          f"  floor {v[3]:6.3f} ms  ({100 * v[3] / v[0] if v[0] else 0:4.0f}% of roofline)  gap {v[0] - v[3]:6.3f}")
mods = collections.Counter()
for n, ms, fl, nb in ops:
    mods[n.split('.')[0]] += ms
print({k: round(v, 2) for k, v in mods.most_common()})
