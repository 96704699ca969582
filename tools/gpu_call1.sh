#!/bin/bash
# round 2, GPU call 1: evidence fixes + fused table path + wire formats
set -x
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
python -c "import __graft_entry__ as g; g.build()" > $O/build.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -15 $O/pytest_gpu.txt
timeout 300 python bench.py > $O/bench_tightly-100k.json 2> $O/bench_default.err; echo "bench rc=$?"
tail -c 1500 $O/bench_default.err
GANGPACK_TABLES=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_tightly-100k_scan.json 2> $O/bench_scan.err; echo "scan rc=$?"
for w in tightly-100k-deep evenly-100k-deep evenly-100k tightly-50k-1m; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"
  GANGPACK_TABLES=0 timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${w}_scan.json 2> $O/bench_${w}_scan.err; echo "$w scan rc=$?"
done
timeout 300 python bench.py --wire int64 --no-cpu-baseline > $O/bench_tightly-100k_int64wire.json 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>&1
BENCH_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches.log 2>&1
BENCH_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_pack_tables -s 3 -c 1 -o $O/r02_pack_tables \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_tables.log 2>&1
GANGPACK_TRACE=1 python tools/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1
ls -la $O
