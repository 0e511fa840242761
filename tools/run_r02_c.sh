set -x
cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02c_gputest.log 2>&1
tail -4 gpurun_out/r02c_gputest.log
python tools/knn_once.py 5
python tools/knn_once.py 5 voxel
python tools/tc_check.py numerics > gpurun_out/r02_plane_tc_numerics.log 2>&1
tail -5 gpurun_out/r02_plane_tc_numerics.log
( time timeout 400 python bench.py --no-primitives --no-ransac ) > gpurun_out/r02c_bench_frames.json 2> gpurun_out/r02c_bench_frames.err
tail -c 400 gpurun_out/r02c_bench_frames.err
python -c "
import json; d=json.load(open('gpurun_out/r02c_bench_frames.json')); print(d['value'], d['e2e']['value'], d['frames']['faithful']['frames_per_s'], d['parity_checked']['ok'], d['frames']['faithful'].get('parity_ok'))"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'knn_collect_kernel|knn_finish_kernel' -c 2 -f -o gpurun_out/r02c_knn python tools/knn_once.py 1 > gpurun_out/r02c_knn_ncu.log 2>&1
tail -2 gpurun_out/r02c_knn_ncu.log
