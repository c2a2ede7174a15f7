#!/bin/bash
# After the last kernel change: the ncu traffic record of THIS build first, then the bench lines that quote it,
# the scale-parity record (K1 / classifier bits against the frozen results of the real reference), the role profile.
P=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
PR="python tools/prof_run.py --windows 1010 --max-batch 1005"
$PR > $P/r2_plain.log 2>&1 && \
  ncu --metrics $M --clock-control none --csv --log-file $P/r2_launches_f16x3.csv $PR > $P/r2_ncu1.log 2>&1
python tools/traffic_record.py $P/r2_launches_f16x3.csv 1010 f16x3 > profiles/r2_traffic.json 2> $P/r2_traffic.err
cp profiles/r2_traffic.json $P/r2_traffic.json
python bench.py > $P/r2_bench.json 2> $P/r2_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $P/r2_ref.json 2> $P/r2_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > $P/r2_smoke.log 2>&1
python -m pytest tests -m gpu -q > $P/r2_tests.log 2>&1
python tools/scale_parity.py > $P/r2_scale_parity.json 2> $P/r2_scale_parity.err
python tools/tc_role_profile.py 1005 f16x3 > $P/r2_roles_f16x3.txt 2>&1
python tools/k1_time.py SS_K1_PACKED > $P/r2_k1_packed.txt 2>&1
tail -3 $P/r2_tests.log; tail -2 $P/r2_smoke.log; head -c 300 $P/r2_bench.json; echo; tail -c 400 $P/r2_scale_parity.json; tail -1 $P/r2_k1_packed.txt
