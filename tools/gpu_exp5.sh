#!/bin/bash
# four gather warps (and + early stage release) in the attention kernel: timing, then the GPU tier, stress and bench with the
# faster variant library
# (variant libraries first:  tools/build_variant.sh g4 svx_winattn -DSVX_WU_LOAD_WARPS=4 -DSVX_WU_EARLY_RELEASE=0 ; g4e = -DSVX_WU_LOAD_WARPS=4
#  -DSVX_WU_EARLY_RELEASE=1 ; base was 2 warps / no early release then.  Result: profiles/r2_winattn_gather_warps.txt; g4e is the default now)
O=gpurun_out; mkdir -p $O
L=swinvox_b200/libswinvox_b200
{
timeout 60 python tools/winattn_time.py 192 2>&1 | tail -1 | sed 's/^/base: /'
SVX_LIB_PATH=${L}_g4.so timeout 60 python tools/winattn_time.py 192 2>&1 | tail -1 | sed 's/^/g4:   /'
SVX_LIB_PATH=${L}_g4e.so timeout 60 python tools/winattn_time.py 192 2>&1 | tail -1 | sed 's/^/g4e:  /'
} > $O/exp5.txt 2>&1
best=$(python - <<'PY'
import re
t={}
for l in open('gpurun_out/exp5.txt'):
    m=re.match(r'(\w+):.*H56 (\d+) us.*H56s (\d+) us.*H28 (\d+) us.*H28s (\d+) us', l)
    if m: t[m.group(1)]=sum(int(x) for x in m.groups()[1:])
print(min(t, key=t.get) if t else 'base')
PY
)
echo "best variant: $best" >> $O/exp5.txt
if [ "$best" = "base" ]; then lib=$L.so; else lib=${L}_$best.so; fi
(SVX_LIB_PATH=$lib timeout 100 python tools/stress_winattn.py 300 64 3 tf32 2>&1 | grep -E "^stress|Error|FAILED|stall" >> $O/exp5.txt) &
SVX_LIB_PATH=$lib timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 >> $O/exp5.txt
wait
SVX_LIB_PATH=$lib timeout 200 python bench.py --no-eager --cpu-seconds 3 > $O/bench_exp5_$best.json 2> $O/bench_exp5.err
cp $O/op_breakdown.json $O/op_breakdown_exp5_$best.json
python tools/op_sum.py "$best" >> $O/exp5.txt
cut -c1-330 $O/bench_exp5_$best.json >> $O/exp5.txt
cat $O/exp5.txt
