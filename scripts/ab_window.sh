# usage: bash scripts/ab_window.sh 96 160 192 256   (window sizes to try on the C2 pcd bench)
for win in "$@"; do
SPARSEPOLY_B200_WINDOW=$win SPARSEPOLY_B200_SWEEP=window timeout 150 python bench.py --workload pcd --steps 2 --warmup 3 --no-cpu --no-also 2> gpurun_out/abw_$win.err | tail -1 > gpurun_out/abw_$win.json
python -c "
import json; l=json.loads(open('gpurun_out/abw_$win.json').read()); print('window=$win', l['value'], l['roofline']['us_per_sequential_step'], l['geometry'], l['zero_update_speculation']['positions'], l['zero_update_speculation']['rejected'])"
done
