cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_c1_full_res.py tests/test_gpu_services.py tests/test_golden.py tests/test_gpu_prefilter.py -m gpu -x -q 2>&1 | tail -2
python tools/frame_once.py 6
for c in 8 16; do
python bench.py --no-primitives --no-ransac --no-faithful --no-cpu-baseline --frame-contexts $c --steps 12 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('contexts', d['frames']['contexts_per_gpu'], 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['parity_checked']['ok'], 'latency', round(d['frames']['frame_latency_ms'],2), round(d['frames']['frame_latency_device_ms'],2))"
done
python tools/determinism2.py 128 6 2>&1 | tail -2
