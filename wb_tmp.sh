for cfg in "$@"; do
  set -- $cfg
  SPARSEPOLY_B200_SWEEP=$1 SPARSEPOLY_B200_HORIZON=$2 SPARSEPOLY_B200_WINDOW=$4 SPARSEPOLY_B200_NEAR=$5 python bench.py --workload pcd --scale $3 --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']
g=l['geometry']
print('$cfg', {k:(round(v,3) if isinstance(v,float) else v) for k,v in g.items() if k in ('sweep','window','horizon','near','hot_frac','max_slots')},'s/epoch',round(l['value'],4),'us/step',round(r['us_per_sequential_step'],4),'e2e',round(l['e2e']['value'],3), 'nzfrac', round(l['p_nonzero_frac'],4))"
done
