cd $GRAFT_REPO_ROOT
python tools/determinism2.py 256 12 2>&1 | tail -4
( time timeout 500 python bench.py ) > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 200 gpurun_out/r02_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json')); print(d['value'], d['e2e']['value'], d['parity_checked']['ok'], d['frames']['frame_latency_ms'], d['frames']['frame_latency_device_ms'], d['frames']['normals']['ms'], d['frames']['faithful']['frames_per_s'])"
