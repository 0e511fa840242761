for cfg in "1.2 1536" "1.2 768" "1.4 768" "1.2 512"; do
  set -- $cfg
  export PITT_KNN_NEED=$1 PITT_KNN_MCAP=$2
  python tools/knn_once.py 2 full > gpurun_out/knn_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg,smsp__inst_executed.sum --clock-control none -s 14 -c 14 --csv --log-file "gpurun_out/knn_launches_$1_$2.csv" python tools/knn_once.py 2 full > gpurun_out/ncu_knn.log 2>&1
done
