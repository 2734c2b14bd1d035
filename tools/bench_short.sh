python bench.py --steps 10 --warmup 3 > gpurun_out/rx_bench.json 2> gpurun_out/rx_bench.err; python -c "
import json; d=json.load(open('gpurun_out/rx_bench.json')); print(d['ms_per_step'], d['value'], {k:round(v['ms_per_launch'],4) for k,v in d['kernels'].items()})"
