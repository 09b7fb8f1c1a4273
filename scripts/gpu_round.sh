#!/bin/bash
# One GPU session: tests, bench lines for the BASELINE configs, ncu launch list + DRAM traffic of one pipeline step.
set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/s5_pytest.log
python scripts/ccl_bench.py 16 > $O/s5_ccl.log 2>&1
python bench.py --steps 10 --warmup 3 > $O/s5_bench_cfg2.json 2> $O/s5_bench_cfg2.err
python bench.py --workload sam2 --variant base_plus --batch 256 --chunk 64 --steps 3 --warmup 3 --no-cpu-baseline > $O/s5_bench_cfg3.json 2> $O/s5_bench_cfg3.err
python bench.py --workload nodes4096 --batch 64 --steps 5 --warmup 3 --no-cpu-baseline > $O/s5_bench_cfg4.json 2> $O/s5_bench_cfg4.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file $O/s5_ncu_dram.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > $O/s5_ncu_dram.log 2>&1
tail -3 $O/s5_pytest.log; cat $O/s5_ccl.log | head -3
