timeout 900 python -m pytest tests/test_parity_gpu.py -x -q --timeout 120 -k "pcd or linear or ragged or cluster or C1 or C2s or all_subsets" 2>&1 | tail -5
for g in "16 32" "16 64" "8 64" "8 128" "4 128" "4 256" "16 128"; do
  set -- $g
  SPARSEPOLY_B200_NCTA=$1 SPARSEPOLY_B200_THREADS=$2 timeout 300 python bench.py --workload pcd --scale 0.1 --steps 1 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read()); r=l['roofline']
    print('C,T=',l['geometry'],'s/epoch',round(l['value'],3),'us/step',round(r['us_per_sequential_step'],3))
except Exception as e: print('fail',e)"
done
