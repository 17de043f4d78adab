mkdir -p gpurun_out
run() { name=$1; shift; env "$@" python bench.py --no-extras --no-cpu-baseline > gpurun_out/g30_$name.json 2> gpurun_out/g30_$name.err; python -c "
import json,sys; d=json.load(open('gpurun_out/g30_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
run off1 ADC_SUM_OFFPATH=1
run off0 ADC_SUM_OFFPATH=0
run off1b ADC_SUM_OFFPATH=1
run off0b ADC_SUM_OFFPATH=0
