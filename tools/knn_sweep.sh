for mc in 1536 3072 6144; do
  export PITT_KNN_MCAP=$mc
  python tools/knn_probe.py 2>&1 | grep -E "full" | sed "s/^/mcap=$mc /"
done
