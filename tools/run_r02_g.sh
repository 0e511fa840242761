cd $GRAFT_REPO_ROOT
for c in 8 12 24 32; do
python bench.py --no-primitives --no-ransac --no-faithful --no-cpu-baseline --frame-contexts $c --steps 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('contexts', d['frames']['contexts_per_gpu'], 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'blocking', d['frames']['blocking_sync'], 'host ms/frame', round(d['frames']['host_cpu_ms_per_frame'],2))"
done
