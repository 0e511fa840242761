set -x
cd $GRAFT_REPO_ROOT
python tools/faithful_once.py 3 16 > gpurun_out/r02_faithful_once.log 2>&1
cat gpurun_out/r02_faithful_once.log
PITT_TRACE=1 python tools/faithful_once.py 2 > gpurun_out/r02_faithful_trace.log 2>&1
tail -40 gpurun_out/r02_faithful_trace.log
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__cycles_active.avg,sm__cycles_elapsed.avg --clock-control none --csv --log-file gpurun_out/r02_faithful_launches.csv python tools/faithful_once.py 2 > gpurun_out/r02_faithful_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'knn_collect_kernel|knn_finish_kernel' -c 2 -f -o gpurun_out/r02_knn python tools/knn_once.py 1 > gpurun_out/r02_knn_ncu.log 2>&1
tail -3 gpurun_out/r02_knn_ncu.log
python tools/knn_once.py 5
python tools/knn_once.py 5 voxel
