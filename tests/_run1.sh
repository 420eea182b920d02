echo scrubnan; SS_SYNC=72 SS_SETS=1 NST=1 python tests/determinism_stress2.py 2>&1 | tail -8
echo plain2sets; NST=2 python tests/determinism_stress2.py 2>&1 | tail -8
python tests/determinism_stress.py 2>&1 | tail -4
