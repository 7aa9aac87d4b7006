#!/bin/bash
# One GPU-box session that produces the round-2 evidence files (copied to profiles/ afterwards).  Run under gpurun, 1 GPU:
#   gpurun --timeout 2400 -- 'bash tools/r02_evidence.sh'
# Rules of /opt/skills/guides/B200_PROFILING.md: every ncu pass runs only after the same command exited 0 without ncu;
# numbers printed under ncu are never bench values.
set -u
O=gpurun_out/r02; mkdir -p $O
python -m pytest tests -m gpu -q -rA 2>&1 | grep -v "^PASSED" | tail -30 > $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; cat $O/smoke.log
python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; rc=$?; echo "bench rc=$rc"; cut -c1-400 $O/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
python tools/config_bench.py > $O/config_bench.jsonl 2> $O/config_bench.err; cat $O/config_bench.jsonl | cut -c1-260
python tools/qr_time.py > $O/qr_time.log 2>&1; cat $O/qr_time.log
python tools/spmm_bench.py > $O/spmm_bench.jsonl 2> $O/spmm_bench.err; tail -3 $O/spmm_bench.jsonl | cut -c1-300
python tools/apps_bench.py > $O/apps_bench.log 2>&1; tail -6 $O/apps_bench.log | cut -c1-300
if [ $rc -eq 0 ]; then
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/bench_for_ncu_plain.json 2> /dev/null && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/ncu_launches.out 2>&1
  python tools/gemm_one.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_gemm_an -c 1 -o $O/gemm_an_full python tools/gemm_one.py > $O/ncu_gemm.out 2>&1
  python tools/qr_one.py 25000 100 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:k_node_factor_cl -c 1 -o $O/node_factor_full python tools/qr_one.py 25000 100 > $O/ncu_node.out 2>&1
  python tools/jacobi_one.py > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:k_jacobi_cl -c 1 -o $O/jacobi_cl_full python tools/jacobi_one.py > $O/ncu_jacobi.out 2>&1
  python tools/spmm_one.py > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:k_csr_spmm_rm -c 1 -o $O/spmm_full python tools/spmm_one.py > $O/ncu_spmm.out 2>&1
fi
ls -la $O
