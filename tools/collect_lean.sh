# Lean round-end evidence run on one B200 (gpurun -- bash tools/collect_lean.sh): bench, reference arm, bench launch list,
# config 4 / 5 and post-processing lines; every program runs plain before any profiler pass.  Outputs: gpurun_out/r95_*.
TAG=r95
P=gpurun_out
BQ="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-modes"
timeout 400 python bench.py > $P/${TAG}_bench.log 2> $P/${TAG}_bench.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $P/${TAG}_ref.log 2>&1; echo ref rc=$?
timeout 200 $BQ > $P/${TAG}_plain_bench.log 2>&1 && \
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv_tc_kernel|features_kernel|conv1_direct|pool_planar|mask_head|regions_kernel|average_kernel|scan_counts|hot_bits" -c 1200 --csv --log-file $P/${TAG}_launches_bench.csv $BQ > $P/${TAG}_ncu_bench.log 2>&1; echo ncu rc=$?
timeout 300 python tools/bench_aux.py silence > $P/${TAG}_silence.log 2>&1; echo silence rc=$?
timeout 300 python tools/bench_aux.py long > $P/${TAG}_long.log 2>&1; echo long rc=$?
timeout 200 python tools/bench_aux.py postproc > $P/${TAG}_postproc.log 2>&1; echo postproc rc=$?
tail -1 $P/${TAG}_bench.log | cut -c1-200
tail -1 $P/${TAG}_silence.log | cut -c1-200; tail -1 $P/${TAG}_long.log | cut -c1-200
