for dbg in 0 1; do
SPARSEPOLY_B200_DEBUG=$dbg python bench.py --workload psgd --rows-per-gpu 1000000 --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']
print('dbg=$dbg samples/s',round(l['value']/1e6,2),'M  ms/step',round(l['ms_per_step'],3),'kernel_ms',{k:round(v/l['steps'],3) for k,v in l['kernel_ms'].items()}, 'e2e', round(l['e2e']['value']/1e6,2),'M', 'whole frac', round(r['whole_step_frac'],3))"
done
