# ncu evidence of round 2 (run under gpurun, one GPU): metric passes over every dense launch of one C2 step and over the C5
# search launch, full captures of the two dominant dense kernels and of the C5 kernel, the launch list of one C2 step
set -x
M=smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
ncu --metrics $M --clock-control none -k regex:'db_|DeviceScan' --csv --log-file gpurun_out/r02_dense_c2_metrics.csv python scripts/dense_one.py C2 all 1 > gpurun_out/r02_dense_c2_metrics.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:db_search_kernel -s 2 -c 1 -f -o gpurun_out/r02_db_search python scripts/dense_one.py C2 all 1 > gpurun_out/r02_db_search.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:db_prefilter_kernel -s 2 -c 1 -f -o gpurun_out/r02_db_prefilter python scripts/dense_one.py C2 all 1 > gpurun_out/r02_db_prefilter.log 2>&1
ncu --metrics $M --clock-control none -k regex:radius_search -s 1 -c 1 --csv --log-file gpurun_out/r02_nn_c5_metrics.csv python scripts/nn_one.py 10000000 0.10 64 4000000 > gpurun_out/r02_nn_c5_metrics.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:radius_search -s 1 -c 1 -f -o gpurun_out/r02_nn_c5 python scripts/nn_one.py 10000000 0.10 64 4000000 > gpurun_out/r02_nn_c5.log 2>&1
STEP_LANES=1 STEP_NMS=1 STEP_PROFILE=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r02_c2.csv python scripts/one_step.py C2 1 > gpurun_out/launches_r02_c2.log 2>&1
