cd $GRAFT_REPO_ROOT
python bench.py --no-primitives --no-faithful --no-cpu-baseline --steps 10 2>gpurun_out/sanity.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['parity_checked']['ok'], 'roofline', round(d['roofline']['frac'],3), 'roofline_frame', d['roofline_frame']['frac'], d['roofline_frame']['kernel_ms'], sorted(d.keys()))"
tail -3 gpurun_out/sanity.err
