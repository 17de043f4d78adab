mkdir -p gpurun_out
python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_dgrad_tc.py tests/test_gpu_tc_nets.py tests/test_gpu_quad.py tests/test_gpu_bench_path.py tests/test_gpu_s2_bwd_tc.py -m gpu -x -q > gpurun_out/g27_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g27_gputest.log
tail -3 gpurun_out/g27_gputest.log
ADC_V2_NO_EPI12=1 python tools/prof_step.py ged 128 > gpurun_out/g27_prof_ged_epi8.txt 2>&1; head -4 gpurun_out/g27_prof_ged_epi8.txt | tail -2
python tools/prof_step.py ged 128 > gpurun_out/g27_prof_ged_epi12.txt 2>&1; head -4 gpurun_out/g27_prof_ged_epi12.txt | tail -2;  grep "tc_conv_rowseg" gpurun_out/g27_prof_ged_epi12.txt | head -14
ADC_V2_NO_EPI12=1 python tools/prof_step.py gud > gpurun_out/g27_prof_gud_epi8.txt 2>&1; head -6 gpurun_out/g27_prof_gud_epi8.txt
python tools/prof_step.py gud > gpurun_out/g27_prof_gud_epi12.txt 2>&1; head -6 gpurun_out/g27_prof_gud_epi12.txt
