mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_dgrad_tc.py tests/test_gpu_tc_nets.py tests/test_gpu_quad.py tests/test_gpu_bench_path.py -m gpu -x -q > gpurun_out/g26_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g26_gputest.log
tail -3 gpurun_out/g26_gputest.log
ADC_V2_NO_RING=1 python tools/prof_step.py ged 128 > gpurun_out/g26_prof_ged_noring.txt 2>&1; head -4 gpurun_out/g26_prof_ged_noring.txt; grep "tc_conv_rowseg" gpurun_out/g26_prof_ged_noring.txt | head -12
python tools/prof_step.py ged 128 > gpurun_out/g26_prof_ged_ring.txt 2>&1; head -4 gpurun_out/g26_prof_ged_ring.txt;  grep "tc_conv_rowseg" gpurun_out/g26_prof_ged_ring.txt | head -12
