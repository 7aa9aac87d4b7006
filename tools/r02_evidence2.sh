#!/bin/bash
# Second evidence session of round 2 (after the guarded CholeskyQR2 / Jacobi changes).  Run under gpurun, 1 GPU:
#   gpurun --timeout 1800 -- 'bash tools/r02_evidence2.sh'
# Every ncu pass runs only after the same command exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
O=gpurun_out/r02b; mkdir -p $O
python -m pytest tests -m gpu -q -rA 2>&1 | grep -v "^PASSED" | tail -30 > $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; cat $O/smoke.log
python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; rc=$?; echo "bench rc=$rc"; cut -c1-400 $O/bench_n1.json
python tools/config_bench.py > $O/config_bench.jsonl 2> $O/config_bench.err; cut -c1-260 $O/config_bench.jsonl
python tools/orth_check.py > $O/orth_check.log 2>&1; tail -1 $O/orth_check.log
python tools/qr_time.py > $O/qr_time.log 2>&1; cat $O/qr_time.log
python tools/jacobi_sweeps.py > $O/jacobi_sweeps.log 2>&1; cut -c1-200 $O/jacobi_sweeps.log
python tools/apps_bench.py --fast > $O/apps_bench.log 2>&1; tail -6 $O/apps_bench.log | cut -c1-300
if [ $rc -eq 0 ]; then
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/bench_for_ncu_plain.json 2> /dev/null && \
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 3000 --csv --log-file $O/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_launches.out 2>&1
  python tools/orth_one.py 200000 100 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_chol_inv -s 2 -c 1 -o $O/chol_inv_full python tools/orth_one.py 200000 100 > $O/ncu_chol.out 2>&1
  python tools/jacobi_one.py > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:k_jacobi_cl2 -c 1 -o $O/jacobi_cl2_full python tools/jacobi_one.py > $O/ncu_jacobi.out 2>&1
fi
ls -la $O
