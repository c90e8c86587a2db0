export STEP_NMS=1
STEP_LANES=8 STEP_PROFILE=0 python scripts/one_step.py C2 20 2>&1 | head -1 | cut -c1-110
STEP_LANES=8 STEP_PROFILE=0 RSGPU_DENSE_SERIAL=0 python scripts/one_step.py C2 20 2>&1 | head -1 | cut -c1-110
python scripts/ab_steps.py C3 1 2>&1 | tail -2
