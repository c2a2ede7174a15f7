#!/bin/bash
# Round-end evidence run on one B200 (gpurun -- bash tools/collect_profiles.sh TAG): every program first runs plain
# (exit 0 without ncu), then under ncu; outputs land in gpurun_out/TAG_* and are summarised into profiles/ by hand.
TAG=${1:-r60}
P=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
PR="python tools/prof_run.py --windows 1010 --max-batch 1005"
BQ="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-modes"
python bench.py > $P/${TAG}_bench.log 2> $P/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $P/${TAG}_ref.log 2>&1
$BQ > $P/${TAG}_plain_bench.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv_tc_kernel|features_kernel|conv1_direct|pool_planar|mask_head|regions_kernel|average_kernel|scan_counts|mel_to_planar" -c 1200 --csv --log-file $P/${TAG}_launches_bench.csv $BQ > $P/${TAG}_ncu_bench.log 2>&1
$PR > $P/${TAG}_plain.log 2>&1 && \
  ncu --metrics $M --clock-control none --csv --log-file $P/${TAG}_launches_f16x3.csv $PR > $P/${TAG}_ncu1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 17 -c 2 -f -o $P/${TAG}_prof_conv $PR > $P/${TAG}_ncu2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:features_kernel -c 1 -f -o $P/${TAG}_prof_feat $PR > $P/${TAG}_ncu3.log 2>&1
python tools/bench_aux.py silence > $P/${TAG}_silence.log 2>&1
python tools/bench_aux.py long > $P/${TAG}_long.log 2>&1
python tools/bench_aux.py long --pcm16 > $P/${TAG}_long_pcm16.log 2>&1
python tools/tc_role_profile.py 1005 f16x3 > $P/${TAG}_roles_f16x3.log 2>&1
tail -2 $P/${TAG}_bench.log | cut -c1-600
tail -1 $P/${TAG}_long.log | cut -c1-300; tail -1 $P/${TAG}_long_pcm16.log | cut -c1-300; tail -1 $P/${TAG}_silence.log | cut -c1-300
ls -la $P/${TAG}_*
