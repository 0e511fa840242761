set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_gputest.log 2>&1
tail -3 gpurun_out/r02_gputest.log
( time timeout 400 python bench.py ) > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 600 gpurun_out/r02_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,sm__cycles_active.avg,sm__cycles_elapsed.avg --clock-control none --csv --log-file gpurun_out/r02_frame_launches.csv python tools/frame_once.py 2 > gpurun_out/r02_frame_once_ncu.log 2>&1
python tools/frame_once.py 5 > gpurun_out/r02_frame_once.log 2>&1
cat gpurun_out/r02_frame_once.log
