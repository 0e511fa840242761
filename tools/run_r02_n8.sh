cd $GRAFT_REPO_ROOT
nproc; nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 --no-primitives --no-ransac --no-cpu-baseline > gpurun_out/r02_bench_n8_frames.json 2> gpurun_out/r02_bench_n8_frames.err
tail -c 300 gpurun_out/r02_bench_n8_frames.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n8_frames.json').read().strip().splitlines()[-1]); print(d['n_gpus'], 'fps', d['value'], 'e2e', d['e2e']['value'], d['parity_checked']['ok'], 'ctx', d['frames']['contexts_per_gpu'], 'blocking', d['frames']['blocking_sync'], 'cores', d['frames']['host_cores'], 'faithful', d['frames']['faithful']['frames_per_s'], 'host ms', d['frames']['host_cpu_ms_per_frame'])"
