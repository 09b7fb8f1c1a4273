#!/bin/bash
# Final GPU session (short form): full test suite, smoke(), the default bench line, the ncu launch list (time + DRAM bytes) of one
# step of the same command.  Run through gpurun from the repo root; everything lands in gpurun_out/.
O=gpurun_out
T=${1:-final}
python -m pytest tests -m gpu -x -q --durations=8 > $O/${T}_pytest.log 2>&1; tail -2 $O/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -3 $O/${T}_smoke.log
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err || tail -5 $O/${T}_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
    --log-file $O/${T}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extras > $O/${T}_ncu_bench.log 2>&1
python -c "
import json; d=json.load(open('$O/${T}_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['extra']['cfg4_ccl']['roofline']['frac'], d['extra']['cfg3']['value'])"
