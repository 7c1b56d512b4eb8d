#!/bin/bash
# The round-end evidence run on one B200 (gpurun): GPU tests, parity report, the bench line, then -- each only after its own command
# exited 0 without ncu -- the launch list and the full captures of the two march kernels (C3, and C5 with 2 views).
# Outputs under gpurun_out/final5/.
O=gpurun_out/final5; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
DIFFRENDER_LIB=differender_b200/libdiffrender_dbg.so timeout 900 python tools/bounds_suite.py > $O/bounds_suite.log 2>&1; tail -1 $O/bounds_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python tools/parity_report.py > $O/parity_report.txt 2> $O/parity_report.err; echo "parity report exit $?"
python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench exit $?"
python tools/bench_line.py final $O/bench_n1.json
B="python bench.py --warmup 3 --no-cpu-baseline --no-e2e --no-others --cuda-profiler-range"
$B --steps 2 > $O/cap_line2.json 2> $O/cap_line2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv $B --steps 2 > $O/ncu_launch.log 2>&1
$B --steps 1 > $O/cap_line.json 2> $O/cap_line.err && \
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:'fwd_kernel|bwd_kernel' -c 2 -o $O/c3_full -f $B --steps 1 > $O/ncu_full.log 2>&1
C5="python bench.py --config c5 --views 2 --warmup 1 --steps 1 --no-cpu-baseline --no-e2e --no-others --cuda-profiler-range"
$C5 > $O/cap_line_c5.json 2> $O/cap_line_c5.err && \
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:'fwd_kernel|bwd_kernel' -c 2 -o $O/c5_full -f $C5 > $O/ncu_full_c5.log 2>&1
ls -la $O
