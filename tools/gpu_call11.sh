#!/bin/bash
# round 2: re-check after the FIFO handshake barrier + the zone-FIFO test fix: FIFO / zone tests, racecheck, fifo benches
set -x
mkdir -p gpurun_out/r02j
O=gpurun_out/r02j
python -c "import __graft_entry__ as g; g.build()" > $O/build.txt 2>&1
timeout 420 python -m pytest tests/test_gpu_zones.py tests/test_gpu_parity.py tests/test_gpu_reference_scenarios.py tests/test_gpu_multi.py tests/test_host_cpp.py -m gpu -q > $O/pytest_fifo_zones.txt 2>&1; echo "rc=$?" >> $O/pytest_fifo_zones.txt; tail -6 $O/pytest_fifo_zones.txt | cut -c1-300
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_zones.py tests/test_gpu_parity.py tests/test_gpu_wire_and_tables.py -m gpu -q \
   -k "(fifo_with_single_az and (1 or 2)) or fifo_zones_blocks or golden_fifo or (random_fifo and (0- or 1-)) or multi_group or degenerate or tables_multi_group" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.log; tail -5 $O/sanitizer_racecheck.log | cut -c1-300
for w in fifo-10k fifo-da-50k; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"; cut -c1-200 $O/bench_$w.json
done
