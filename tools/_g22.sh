mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/g22_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g22_gputest.log
tail -3 gpurun_out/g22_gputest.log
python tools/prof_step.py ged 128 > gpurun_out/g22_prof_ged.txt 2>&1; head -24 gpurun_out/g22_prof_ged.txt
python bench.py --no-extras --no-cpu-baseline > gpurun_out/g22_bench.json 2> gpurun_out/g22_bench.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('gpurun_out/g22_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
