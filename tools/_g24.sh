mkdir -p gpurun_out
run() { name=$1; shift; env "$@" python bench.py --no-extras --no-cpu-baseline > gpurun_out/g24_$name.json 2> gpurun_out/g24_$name.err; python -c "
import json,sys; d=json.load(open('gpurun_out/g24_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'])"; }
run base X=1
run after ADC_WGRAD_AFTER=1
run after_res40 ADC_WGRAD_AFTER=1 ADC_WGRAD_SMEM_RESERVE=40
run res40 ADC_WGRAD_SMEM_RESERVE=40
run after_res80 ADC_WGRAD_AFTER=1 ADC_WGRAD_SMEM_RESERVE=80
