#!/bin/bash
# Final GPU-box session of a round: full -m gpu suite, bench lines of every config, reference arm, cfg1.  Outputs under gpurun_out/.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/t.log 2>&1
(timeout 400 python bench.py 2>&1 | tail -1) > gpurun_out/b_cfg2.log 2>&1
for c in default cfg3 cfg4 cfg5; do
  (timeout 400 python bench.py --config $c 2>&1 | tail -1) > gpurun_out/b_$c.log 2>&1
done
(timeout 300 python bench.py --impl reference --steps 5 2>&1 | tail -1) > gpurun_out/b_ref.log 2>&1
(timeout 300 python tests/bench_cfg1.py 2>&1 | tail -3) > gpurun_out/cfg1.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s.log 2>&1
tail -3 gpurun_out/t.log; cat gpurun_out/s.log gpurun_out/cfg1.log
