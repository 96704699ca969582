#!/bin/bash
# round 2, GPU call 2: thread-per-app table kernel, FIFO warp-first, sort, zones, reservations, multi
set -x
mkdir -p gpurun_out/r02b
O=gpurun_out/r02b
python -c "import __graft_entry__ as g; g.build()" > $O/build.txt 2>&1
# new / changed areas first, then everything
timeout 900 python -m pytest tests/test_gpu_wire_and_tables.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/pytest_tables.txt 2>&1; echo "rc=$?" >> $O/pytest_tables.txt; tail -5 $O/pytest_tables.txt
timeout 900 python -m pytest tests/test_gpu_zones.py tests/test_gpu_multi.py tests/test_gpu_reference_scenarios.py -m gpu -q > $O/pytest_new.txt 2>&1; echo "rc=$?" >> $O/pytest_new.txt; tail -15 $O/pytest_new.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_host_cpp.py -m gpu -q > $O/pytest_parity.txt 2>&1; echo "rc=$?" >> $O/pytest_parity.txt; tail -15 $O/pytest_parity.txt
timeout 300 python bench.py > $O/bench_tightly-100k.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 800 $O/bench_default.err
GANGPACK_CHUNK_APPS=50000 timeout 300 python bench.py --no-cpu-baseline > $O/bench_tightly-100k_chunk50k.json 2>/dev/null
GANGPACK_CHUNK_APPS=1000000 timeout 300 python bench.py --no-cpu-baseline > $O/bench_tightly-100k_nochunk.json 2>/dev/null
GANGPACK_TABLES=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_tightly-100k_scan.json 2> $O/bench_scan.err
for w in evenly-100k tightly-100k-deep fifo-10k fifo-da-50k tightly-50k-1m; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"
done
GANGPACK_TABLES=0 timeout 300 python bench.py --workload tightly-100k-deep --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_tightly-100k-deep_scan.json 2>/dev/null
timeout 300 python tools/multi_bench.py --config 3 --devices 1 > $O/multi_c3_1gpu.json 2> $O/multi_c3.err
GANGPACK_TRACE=1 python tools/e2e_breakdown.py > $O/e2e_breakdown.txt 2>&1
BENCH_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches.log 2>&1
ls -la $O
