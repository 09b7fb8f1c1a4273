#!/bin/bash
# Final GPU session of a round: full test suite, the default bench line, the ncu launch list (time + DRAM bytes) of one step of the
# same command, and ncu --set full of the labelling passes.  Run through gpurun from the repo root; everything lands in gpurun_out/.
O=gpurun_out
T=${1:-final}
python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || tail -5 $O/${T}_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
    --log-file $O/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extras > $O/${T}_ncu_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_ccl -c 10 -o $O/${T}_ccl python scripts/ccl_bench.py 64 > $O/${T}_ccl_ncu.log 2>&1
tail -2 $O/${T}_ccl_ncu.log
python -c "
import json; d=json.load(open('$O/${T}_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['extra']['cfg4_ccl']['roofline']['frac'], d['extra']['cfg3']['value'])"
