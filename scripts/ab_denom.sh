# usage: bash scripts/ab_denom.sh "<gamma> ..." "<denom> ..."   (C2 pcd bench in other sparsity regimes)
for g in $1; do for dn in $2; do
SPARSEPOLY_B200_SPEC_DENOM=$dn timeout 150 python bench.py --workload pcd --gamma $g --steps 2 --warmup 3 --no-cpu --no-also 2> gpurun_out/abd.err | tail -1 > gpurun_out/abd.json
python -c "
import json; l=json.loads(open('gpurun_out/abd.json').read()); print('gamma=$g denom=$dn', round(l['value'],4), 's/epoch', l['p_nonzero_frac_by_order'], l['zero_update_speculation']['positions'], l['zero_update_speculation']['rejected'])"
done; done
