for g in "1 256" "1 128" "2 256" "4 128" "8 64" "16 32" "8 32" "16 64" "4 256"; do
  set -- $g
  SPARSEPOLY_B200_NCTA=$1 SPARSEPOLY_B200_THREADS=$2 python bench.py --workload pcd --scale 0.1 --steps 1 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); r=l['roofline']
print('C,T=',l['geometry'],'s/epoch',round(l['value'],3),'us/step',round(r['us_per_sequential_step'],3))"
done
