SS_VERBOSE=1 python tests/tune_physics.py 2>&1 | tail -2
python tests/determinism_stress.py 2>&1 | tail -2
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
