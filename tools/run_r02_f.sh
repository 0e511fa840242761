cd $GRAFT_REPO_ROOT
python tools/knn_probe.py 2>&1 | grep -E "full|voxel"
for w in 1 2 4 8; do echo "wide mult $w"; PITT_KNN_WIDE=$w python tools/knn_once.py 6; done
for w in 2 4 8; do PITT_KNN_WIDE=$w timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'knn_collect_kernel|knn_finish_kernel' -c 2 python tools/knn_once.py 1 2>&1 | grep -E "duration" ; done
