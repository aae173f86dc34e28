#!/bin/bash
# Run on the GPU box (via gpurun): launch list + one --set full capture of the step's top kernels, exported to CSV
# there so that only small files travel back.  Usage: tools/ncu_export.sh <tag> [batch]
set -u
TAG=${1:-r1}; B=${2:-16}
OUT=gpurun_out
python tools/ncu_step.py $B > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/ncu_step.py $B > $OUT/${TAG}_list.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_tc_kernel|attn_tc_kernel|gn_apply|gn_stats|layernorm" -c 40 -f -o /tmp/${TAG}_prof \
    python tools/ncu_step.py $B > $OUT/${TAG}_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > $OUT/${TAG}_prof_raw.csv 2>/dev/null
ls -la /tmp/${TAG}_prof.ncu-rep
SZ=$(stat -c %s /tmp/${TAG}_prof.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -lt 40000000 ]; then cp /tmp/${TAG}_prof.ncu-rep $OUT/; fi
tail -n 2 $OUT/${TAG}_list.log; tail -n 2 $OUT/${TAG}_full.log
