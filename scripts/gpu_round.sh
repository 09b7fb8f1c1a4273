#!/bin/bash
# One GPU session: tests, bench lines for the BASELINE configs, ncu launch list + DRAM traffic of one pipeline step.
set -x
O=gpurun_out
T=${1:-r}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/${T}_pytest.log
python scripts/ccl_bench.py 16 > $O/${T}_ccl.log 2>&1
python scripts/profile_sam2.py tiny 64 64 > $O/${T}_prof_tiny.log 2>&1
python bench.py --steps 10 --warmup 3 > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err
python bench.py --workload sam2 --variant base_plus --batch 256 --chunk 64 --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err
python bench.py --workload nodes4096 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > $O/${T}_bench_cfg4.json 2> $O/${T}_bench_cfg4.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file $O/${T}_ncu_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > $O/${T}_ncu_dram.log 2>&1
tail -3 $O/${T}_pytest.log; cat $O/${T}_ccl.log | head -3
