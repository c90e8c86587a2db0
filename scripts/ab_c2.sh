export STEP_LANES=8 STEP_NMS=1 STEP_PROFILE=0 RSGPU_ICP_IMPL=split
for cfg in "A=1" "RSGPU_DENSE_BPS=2" "RSGPU_DENSE_BPS=1" "RSGPU_DENSE_BPS=2 RSGPU_ICP_IMPL=persistent" "RSGPU_DENSE_BPS=2 RSGPU_ICP_IMPL=persistent RSGPU_ICP_CTAS=32" "RSGPU_DENSE_BPS=2 RSGPU_DENSE_SERIAL=0"; do
  env $cfg python scripts/one_step.py C2 20 2>&1 | head -1 | cut -c1-150
done
STEP_LANES=1 STEP_PROFILE=1 RSGPU_DENSE_BPS=2 python scripts/one_step.py C2 5 2>&1 | head -1 | cut -c1-300
STEP_TRACE=1 python scripts/one_step.py C2 10 2>&1 | cut -c1-100
