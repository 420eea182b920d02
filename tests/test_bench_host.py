"""Host logic of bench.py that needs no GPU: the counter-based ctrl stream must be the same on the device side
(torch) and on the CPU-baseline side (numpy), for any env offset (multi-GPU ranks) and period."""
import json
import os

import numpy as np
import torch

import bench


def test_ctrl_stream_identical_in_numpy_and_torch(arrays_E):
    A, _ = arrays_E
    lo, hi = A["actuator_ctrlrange"][:, 0].copy(), A["actuator_ctrlrange"][:, 1].copy()
    tlo, thi = torch.tensor(lo, dtype=torch.float64), torch.tensor(hi, dtype=torch.float64)
    for env0, nenv, period in ((0, 64, 0), (4096, 33, 7), (8 * 4096 - 5, 5, 123)):
        a = bench.ctrl_np(0, env0, nenv, period, lo, hi)
        b = bench.ctrl_torch(0, env0, nenv, period, tlo, thi, "cpu").numpy()
        assert a.shape == (nenv, len(lo))
        assert np.array_equal(a.astype(np.float32), b)
        assert np.all(a >= lo) and np.all(a <= hi)
    # different envs / periods / actuators draw different values; the same key draws the same value
    x = bench.ctrl_np(0, 0, 8, 0, lo, hi)
    assert len(np.unique(x[:, 2])) == 8 and not np.array_equal(x, bench.ctrl_np(0, 0, 8, 1, lo, hi))
    assert np.array_equal(x[3], bench.ctrl_np(0, 3, 1, 0, lo, hi)[0])          # rank offset = env offset


def test_committed_bench_line_has_the_contract_keys():
    line = json.load(open(os.path.join(bench.ROOT, "profiles", "bench_r1_n1.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["metric"] == bench.METRIC and "workload" in line["config"] and line["gpu_launches"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-12
