#!/bin/bash
# Round-end evidence run on one B200 (gpurun -- bash tools/collect_profiles.sh TAG): every program first runs plain
# (exit 0 without ncu), then under ncu; outputs land in gpurun_out/TAG_* and are copied / summarised into profiles/.
TAG=${1:-r2}
P=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
PR="python tools/prof_run.py --windows 1010 --max-batch 1005"
BQ="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-other-modes --corpus-files 0"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $P/${TAG}_clocks.csv &
SMI=$!
python bench.py > $P/${TAG}_bench.json 2> $P/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $P/${TAG}_ref.json 2> $P/${TAG}_ref.err
python tools/scale_parity.py > $P/${TAG}_scale_parity.json 2> $P/${TAG}_scale_parity.err
python tools/precision_study.py > $P/${TAG}_precision_study.txt 2>&1
python tools/bench_aux.py long > $P/${TAG}_long.json 2> $P/${TAG}_long.err
python tools/bench_aux.py long --pcm16 > $P/${TAG}_long_pcm16.json 2> $P/${TAG}_long_pcm16.err
python tools/bench_aux.py silence > $P/${TAG}_silence.json 2> $P/${TAG}_silence.err
python tools/bench_aux.py config1 > $P/${TAG}_config1.json 2> $P/${TAG}_config1.err
python tools/tc_role_profile.py 1005 f16x3 > $P/${TAG}_roles_f16x3.txt 2>&1
kill $SMI
$BQ > $P/${TAG}_plain_bench.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv_tc_kernel|features_kernel|conv1_direct|pool_planar|mask_head|regions_kernel|average_kernel|scan_counts|compact_flags" -c 1200 --csv --log-file $P/${TAG}_launches_bench.csv $BQ > $P/${TAG}_ncu_bench.log 2>&1
$PR > $P/${TAG}_plain.log 2>&1 && \
  ncu --metrics $M --clock-control none --csv --log-file $P/${TAG}_launches_f16x3.csv $PR > $P/${TAG}_ncu1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 17 -c 2 -f -o $P/${TAG}_prof_conv $PR > $P/${TAG}_ncu2.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:features_kernel -c 1 -f -o $P/${TAG}_prof_feat $PR > $P/${TAG}_ncu3.log 2>&1
tail -c 400 $P/${TAG}_bench.json; echo
tail -c 300 $P/${TAG}_long.json; echo
ls -la $P/${TAG}_*
