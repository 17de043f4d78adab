mkdir -p gpurun_out
python -m pytest tests/test_gpu_bench_path.py tests/test_gpu_optim.py tests/test_gpu_dropin.py tests/test_gpu_nets.py tests/test_gpu_tc_nets.py tests/test_gpu_feeder.py -m gpu -x -q > gpurun_out/g29_gputest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g29_gputest.log
tail -3 gpurun_out/g29_gputest.log
run() { name=$1; shift; env "$@" python bench.py --no-extras --no-cpu-baseline > gpurun_out/g29_$name.json 2> gpurun_out/g29_$name.err; python -c "
import json,sys; d=json.load(open('gpurun_out/g29_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
run base ADC_ADOPT_GRADS=0
run adopt ADC_ADOPT_GRADS=1
run base2 ADC_ADOPT_GRADS=0
run adopt2 ADC_ADOPT_GRADS=1
