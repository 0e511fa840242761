cd $GRAFT_REPO_ROOT
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_gputest.log 2>&1
tail -4 gpurun_out/r02_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
