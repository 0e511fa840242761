set -x
cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests/test_gpu_normals_clusters.py tests/test_gpu_c1_full_res.py tests/test_gpu_services.py tests/test_golden.py -m gpu -x -q ) > gpurun_out/r02e_gputest.log 2>&1
tail -4 gpurun_out/r02e_gputest.log
python tools/knn_once.py 5
python tools/knn_once.py 5 voxel
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio --clock-control none -k regex:'knn_collect_kernel|knn_finish_kernel' -c 2 python tools/knn_once.py 1 2>&1 | grep -E "knn_|duration|inst_executed|issue_active|warps_active|scoreboard" > gpurun_out/r02e_knn_quick.log
cat gpurun_out/r02e_knn_quick.log
( time timeout 400 python bench.py --no-primitives --no-ransac --no-faithful --no-cpu-baseline ) > gpurun_out/r02e_bench_frames.json 2> gpurun_out/r02e_bench_frames.err
python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_frames.json')); print(d['value'], d['e2e']['value'], d['parity_checked']['ok'])"
