cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3; nproc
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -5 gpurun_out/r02_bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['e2e']['value'], d['parity_checked'], d['frames'].get('normals',{}).get('ms'), d['frames']['faithful']['frames_per_s'], d['ransac']['value'] if 'ransac' in d else None)"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err
tail -3 gpurun_out/r02_bench_ref_n2.err; cat gpurun_out/r02_bench_ref_n2.json | cut -c1-600
