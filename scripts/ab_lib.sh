# A/B of two builds of the library on one box: scripts/ab_lib.sh "<gamma> ..."
for g in $1; do for lib in "" "$PWD/sparsepoly_b200/libsp_alt.so"; do
SPARSEPOLY_B200_LIB=$lib timeout 150 python bench.py --workload pcd --gamma $g --steps 2 --warmup 3 --no-cpu --no-also 2> gpurun_out/abl.err | tail -1 > gpurun_out/abl.json
python -c "
import json; l=json.loads(open('gpurun_out/abl.json').read()); print('gamma=$g lib=${lib:-default}', round(l['value'],4), 's/epoch', l['p_nonzero_frac_by_order'], l['zero_update_speculation']['positions'], l['zero_update_speculation']['rejected'])"
done; done
