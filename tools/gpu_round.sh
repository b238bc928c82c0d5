#!/bin/bash
# One GPU session: parity tests, smoke, bench, launch list, full ncu capture of the gcn0 kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/round_status.txt
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/round_status.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/round_status.txt
python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_nograph.json 2>> gpurun_out/bench.err; echo "bench nograph rc=$?" | tee -a gpurun_out/round_status.txt
python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain_step.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_step.log 2>&1
echo "ncu launches rc=$?" | tee -a gpurun_out/round_status.txt
python tools/prof_gcn0.py > gpurun_out/plain_gcn0.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:gcn0 -s 6 -c 6 -o gpurun_out/prof_gcn0 -f python tools/prof_gcn0.py > gpurun_out/ncu_gcn0.log 2>&1
echo "ncu gcn0 rc=$?" | tee -a gpurun_out/round_status.txt
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench.json gpurun_out/bench_nograph.json; tail -3 gpurun_out/bench.err
