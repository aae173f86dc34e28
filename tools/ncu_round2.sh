#!/bin/bash
# Round-2 Nsight Compute session (run on the GPU box via gpurun): launch list of one B=16 step, --set full of the first
# kernels of every family, the attention kernel's source page, and the fused MSDeformAttn kernel of the TESTR head.
set -u
OUT=gpurun_out; TAG=${1:-r2}
python tools/ncu_step.py 16 > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
python tools/one_testr.py 16 > $OUT/${TAG}_plain_testr.log 2>&1 || { echo "plain testr run failed"; tail -5 $OUT/${TAG}_plain_testr.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/${TAG}_launches.csv python tools/ncu_step.py 16 > $OUT/${TAG}_list.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"gemm_tc|attn_tc_kernel|attn_kvs_kernel|gn_apply|gn_stats|row_stats" -c 70 -f -o /tmp/${TAG}_prof \
    python tools/ncu_step.py 16 > $OUT/${TAG}_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > $OUT/${TAG}_prof_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv -k regex:"attn_tc_kernel" > $OUT/${TAG}_attn_source.csv 2>/dev/null
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"msda_fused" -c 2 -f -o /tmp/${TAG}_msda \
    python tools/one_testr.py 16 > $OUT/${TAG}_msda.log 2>&1
ncu -i /tmp/${TAG}_msda.ncu-rep --page raw --csv > $OUT/${TAG}_msda_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_msda.ncu-rep --page source --csv > $OUT/${TAG}_msda_source.csv 2>/dev/null
ls -la /tmp/${TAG}_prof.ncu-rep /tmp/${TAG}_msda.ncu-rep
tail -n 2 $OUT/${TAG}_list.log $OUT/${TAG}_full.log $OUT/${TAG}_msda.log
