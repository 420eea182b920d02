python -m pytest tests/test_parity_r2_gpu.py -m gpu -q -s -k "frames_actuators" 2>&1 | grep -v "^$" | grep -n "^E  \|Error\|passed\|failed\|>  " | cut -c1-300 | head -40
