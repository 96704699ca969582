#!/bin/bash
# round 2, last call: the suite exactly as the driver runs it, captures the verdict asked for (distribute-evenly decision
# kernel, minimal-fragmentation, the zone-FIFO kernel), kernel-only times of the node-order sort
set -x
mkdir -p gpurun_out/r02l
O=gpurun_out/r02l
python -c "import __graft_entry__ as g; g.build()" > $O/build.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu --durations=10 > $O/pytest_gpu.txt 2>&1; echo "rc=$?" >> $O/pytest_gpu.txt; tail -16 $O/pytest_gpu.txt | cut -c1-200
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/sort_launches.csv python tools/sort_bench.py > $O/sort_bench.txt 2>&1; cat $O/sort_bench.txt
BENCH_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gp_decide_tables -s 3 -c 1 -o $O/r02_decide_evenly python bench.py --workload evenly-100k --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_decide_evenly.log 2>&1
BENCH_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gp_pack_independent -s 2 -c 1 -o $O/r02_minfrag python bench.py --workload minfrag-100k --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_minfrag.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gp_pack_fifo_zones_cta -s 1 -c 1 -o $O/r02_zone_fifo python tools/zone_fifo_bench.py > $O/ncu_zone_fifo.log 2>&1
ls -la $O
