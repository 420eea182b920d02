#!/bin/bash
# 8-GPU session: weak and strong scaling of cfg2, cfg5 sharded over the box.  Outputs under gpurun_out/.
mkdir -p gpurun_out
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" 2>&1 | grep "^{" | tail -1; }
run --steps 10 > gpurun_out/scale_n${N}_cfg2_weak.json
run --steps 10 --scaling strong > gpurun_out/scale_n${N}_cfg2_strong.json
run --config cfg5 > gpurun_out/scale_n${N}_cfg5.json
cat gpurun_out/scale_n${N}_*.json | cut -c1-400
