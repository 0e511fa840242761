cd $GRAFT_REPO_ROOT
run() {
  python bench.py --no-primitives --no-ransac --no-faithful --no-cpu-baseline --steps 8 2>/dev/null | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print(os.environ.get('TAG'), 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['parity_checked']['ok'], 'normals ms', round(d['frames']['normals']['ms'],3))"
}
TAG=default run
TAG=need1.0 PITT_KNN_NEED=1.0 run
TAG=need1.4 PITT_KNN_NEED=1.4 run
TAG=mcap640 PITT_KNN_MCAP=640 run
TAG=mcap1000 PITT_KNN_MCAP=1000 run
TAG=cavg2.0 PITT_KNN_CAVG=2.0 run
TAG=cavg3.2 PITT_KNN_CAVG=3.2 run
