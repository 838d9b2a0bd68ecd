for spec in 1 0; do
SPARSEPOLY_B200_SPEC=$spec timeout 150 python bench.py --workload pcd --steps 2 --warmup 3 --no-cpu --no-also 2> gpurun_out/ab_$spec.err | tail -1 > gpurun_out/ab_$spec.json
python -c "
import json; l=json.loads(open('gpurun_out/ab_$spec.json').read()); print('spec=$spec', l['value'], l['roofline']['us_per_sequential_step'], l['p_nonzero_frac_by_order'], l['zero_update_speculation'], l['kernel_ms'])"
done
