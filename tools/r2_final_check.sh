set -x
O=gpurun_out; P=r3e
python -m pytest tests -m gpu -x -q > $O/${P}_gputests.log 2>&1; tail -3 $O/${P}_gputests.log
python bench.py > $O/${P}_bench_n1.json 2> $O/${P}_bench_n1.err
LG=20 REPS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/${P}_msm_launches.csv python tools/msm_probe.py > $O/${P}_msm_ncu.log 2>&1
python tools/msm_sizes_probe.py > $O/${P}_msm_sizes.log 2>&1; cat $O/${P}_msm_sizes.log | tr '\n' ' '
