cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_normals_clusters.py tests/test_gpu_c1_full_res.py -m gpu -x -q 2>&1 | tail -2
python tools/knn_once.py 6
timeout 300 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'knn_finish_kernel' -c 1 python tools/knn_once.py 1 2>&1 | grep -E "duration|hit_rate|dram__|issue_active"
for i in 1 2; do
python bench.py --no-primitives --no-ransac --no-faithful --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['parity_checked']['ok'])"
done
