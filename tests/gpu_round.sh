#!/bin/bash
# One GPU-box session: full -m gpu suite, bench lines of every config, ncu launch list + full captures.  Outputs under gpurun_out/.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/t.log 2>&1
(timeout 400 python bench.py 2>&1 | tail -1) > gpurun_out/b_cfg2.log 2>&1
for c in default cfg3 cfg4 cfg5; do
  (timeout 400 python bench.py --config $c 2>&1 | tail -1) > gpurun_out/b_$c.log 2>&1
done
(timeout 300 python bench.py --impl reference --steps 5 2>&1 | tail -1) > gpurun_out/b_ref.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1
timeout 700 ncu --set full --clock-control none --import-source on -k regex:"ss_(smooth|narrow|solve)" -s 1404 -c 9 -f -o gpurun_out/phys_r2b python tests/profile_physics.py > gpurun_out/prof.log 2>&1
tail -3 gpurun_out/t.log
