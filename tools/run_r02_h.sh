cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests/test_gpu_c1_full_res.py tests/test_gpu_services.py tests/test_gpu_prefilter.py tests/test_gpu_split.py -m gpu -x -q ) 2>&1 | tail -4
for i in 1 2; do
python bench.py --no-primitives --no-ransac --no-faithful --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['parity_checked']['ok'], d['parity_checked']['resident_equals_e2e_on_all_frames'])"
done
compute-sanitizer --tool memcheck python tools/frames_probe.py 4 2>&1 | tail -5
