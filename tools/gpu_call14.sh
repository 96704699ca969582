#!/bin/bash
# round 2: table path after the first-host word / galloping jump / pipelined walk / warp-aggregated interning / block-sum offsets
set -x
mkdir -p gpurun_out/r02m/bench
O=gpurun_out/r02m
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/build_smoke.txt 2>&1; echo "smoke rc=$?"
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.txt 2>&1; echo "rc=$?" >> $O/pytest_gpu.txt; tail -5 $O/pytest_gpu.txt | cut -c1-200
timeout 200 python bench.py --steps 20 --warmup 5 > $O/bench/bench_tightly-100k.json 2> $O/bench/bench_tightly-100k.err; echo "default rc=$?"
for w in evenly-100k tightly-100k-deep tightly-50k-1m; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench/bench_$w.json 2> $O/bench/bench_$w.err; echo "$w rc=$?"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
cut -c1-330 $O/bench/*.json
