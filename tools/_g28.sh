mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/g28_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g28_gputest.log
tail -3 gpurun_out/g28_gputest.log
python bench.py > gpurun_out/g28_bench.json 2> gpurun_out/g28_bench.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('gpurun_out/g28_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
for c in ged gud gms s2d ud full; do python tools/prof_step.py $c > gpurun_out/g28_prof_$c.txt 2>&1; head -1 gpurun_out/g28_prof_$c.txt; done
B=128 STEPS=3 MARK=2 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches_bench_step_b128.csv python tools/prof_bench_step.py > gpurun_out/g28_ncu.log 2>&1; tail -2 gpurun_out/g28_ncu.log
