#!/bin/bash
# round 2, final 1-GPU call: smoke, the whole -m gpu suite, zone-FIFO timing, every bench workload, launch list, sanitizers
set -x
mkdir -p gpurun_out/r02i/bench
O=gpurun_out/r02i
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/build_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/build_smoke.txt | cut -c1-300
timeout 400 python -m pytest tests/test_gpu_zones.py tests/test_host_cpp.py -m gpu -q > $O/pytest_zones_host.txt 2>&1; echo "rc=$?" >> $O/pytest_zones_host.txt; tail -25 $O/pytest_zones_host.txt | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "rc=$?" >> $O/pytest_gpu.txt; tail -8 $O/pytest_gpu.txt | cut -c1-300
timeout 200 python tools/zone_fifo_bench.py > $O/zone_fifo_bench.json 2> $O/zone_fifo_bench.err; cat $O/zone_fifo_bench.json; tail -3 $O/zone_fifo_bench.err
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench/bench_tightly-100k.json 2> $O/bench/bench_tightly-100k.err; echo "default rc=$?"
for w in evenly-100k tightly-100k-deep fifo-10k; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 > $O/bench/bench_$w.json 2> $O/bench/bench_$w.err; echo "$w rc=$?"
done
for w in evenly-100k-deep tightly-50k-1m fifo-da-50k minfrag-100k; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/bench/bench_$w.json 2> $O/bench/bench_$w.err; echo "$w rc=$?"
done
GANGPACK_TABLES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench/bench_tightly-100k_scan-path.json 2> $O/bench/bench_scan.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench/bench_reference.json 2> $O/bench/bench_reference.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
# sanitizers (time-boxed): the zone-FIFO kernel, the warp-first FIFO kernel after the fixes, the small parity tests of the other families
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_zones.py tests/test_gpu_parity.py -m gpu -q \
   -k "fifo_with_single_az or fifo_zones_blocks or degenerate or golden or (random_fifo and 0-) or multi_group or reschedule or single_az_packers" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.log; tail -4 $O/sanitizer_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_zones.py tests/test_gpu_parity.py tests/test_gpu_wire_and_tables.py -m gpu -q \
   -k "(fifo_with_single_az and 1) or golden_fifo or (random_fifo and 0-) or multi_group or degenerate or tables_multi_group" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.log; tail -4 $O/sanitizer_racecheck.log
cut -c1-400 $O/bench/*.json
