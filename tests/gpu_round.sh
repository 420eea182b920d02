#!/bin/bash
# One GPU-box session: full -m gpu suite, default benches, ncu launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -150) > gpurun_out/t.log 2>&1
for c in default cfg3 cfg4; do
  (timeout 400 python bench.py --config $c --no-cpu 2>&1 | tail -3) > gpurun_out/b_$c.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1
tail -5 gpurun_out/t.log
