cd $GRAFT_REPO_ROOT
nproc
for cores in 0-7 0-5; do
echo "== 2 GPUs on host cores $cores"
taskset -c $cores python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-primitives --no-ransac --no-faithful --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ctx', d['frames']['contexts_per_gpu'], 'blocking', d['frames']['blocking_sync'], 'cores', d['frames']['host_cores'], 'host ms/frame', round(d['frames']['host_cpu_ms_per_frame'],2))"
done
echo "== 8 contexts per GPU on 8 cores"
taskset -c 0-7 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 --no-primitives --no-ransac --no-faithful --no-cpu-baseline --frame-contexts 8 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ctx', d['frames']['contexts_per_gpu'], 'blocking', d['frames']['blocking_sync'], 'host ms/frame', round(d['frames']['host_cpu_ms_per_frame'],2))"
echo "== 24 contexts per GPU on 8 cores"
taskset -c 0-7 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-primitives --no-ransac --no-faithful --no-cpu-baseline --frame-contexts 24 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ctx', d['frames']['contexts_per_gpu'], 'blocking', d['frames']['blocking_sync'], 'host ms/frame', round(d['frames']['host_cpu_ms_per_frame'],2))"
