mkdir -p gpurun_out
python -m pytest tests/test_gpu_nconv.py tests/test_gpu_bench_path.py -m gpu -x -q > gpurun_out/g23_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g23_gputest.log
tail -3 gpurun_out/g23_gputest.log
python tools/prof_step.py ged 128 > gpurun_out/g23_prof_ged.txt 2>&1; head -12 gpurun_out/g23_prof_ged.txt; grep "nconv_bwd_data" gpurun_out/g23_prof_ged.txt | head
