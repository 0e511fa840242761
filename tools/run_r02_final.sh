set -x
cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_gputest.log 2>&1
tail -4 gpurun_out/r02_gputest.log
( time timeout 500 python bench.py ) > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 300 gpurun_out/r02_bench_n1.err
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__cycles_active.avg,sm__cycles_elapsed.avg --clock-control none --csv --log-file gpurun_out/r02_frame_launches.csv python tools/frame_once.py 2 > gpurun_out/r02_frame_once_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'knn_collect_kernel|knn_finish_kernel' -c 2 -f -o gpurun_out/r02_knn python tools/knn_once.py 1 > gpurun_out/r02_knn_ncu.log 2>&1
tail -2 gpurun_out/r02_knn_ncu.log
