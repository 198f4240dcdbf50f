#!/bin/bash
# Round-2 evidence run (one B200, under gpurun):  tests, bench line, ncu launch list of the bench command, ncu --set full of
# the headline kernel (Miller loop alone and Miller + final exponentiation) and of the MSM bucket kernel.
# Outputs land in gpurun_out/ with the prefix given as $1; tools/profile_summaries.py turns them into profiles/*.md.
set -x
P=${1:-r2}
O=gpurun_out
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -m gpu -x -q > $O/${P}_gputests.log 2>&1; tail -3 $O/${P}_gputests.log; fi
python bench.py > $O/${P}_bench_n1.json 2> $O/${P}_bench_n1.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file $O/${P}_bench_launches.csv python bench.py --steps 3 --warmup 3 > $O/${P}_bench_ncu.json 2> $O/${P}_bench_ncu.err
CID=5 N=11840 ncu --set full --clock-control none --import-source on -k regex:vm_pairing_kernel -s 2 -c 2 -f \
    -o $O/${P}_pair python tools/pair_probe.py > $O/${P}_pair_ncu.log 2>&1
LG=20 REPS=1 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel -s 1 -c 1 -f \
    -o $O/${P}_msm_acc python tools/msm_probe.py > $O/${P}_msm_acc_ncu.log 2>&1
# gpurun brings back at most 64 MiB: export the pages that the summaries need here and drop the bulky report
ncu -i $O/${P}_msm_acc.ncu-rep --page raw --csv > $O/${P}_msm_acc_raw.csv 2>/dev/null
ncu -i $O/${P}_msm_acc.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $O/${P}_msm_acc_sass.csv.gz
rm -f $O/${P}_msm_acc.ncu-rep
ncu -i $O/${P}_pair.ncu-rep --page raw --csv > $O/${P}_pair_raw.csv 2>/dev/null
LG=20 REPS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/${P}_msm_launches.csv python tools/msm_probe.py > $O/${P}_msm_ncu.log 2>&1
LG=20 python tools/msm_probe.py > $O/${P}_msm_plain.log 2>&1; grep ms $O/${P}_msm_plain.log
ls -la $O | grep ${P}_
