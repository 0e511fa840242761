set -x
timeout 120 python tools/tc_check.py numerics > gpurun_out/tc_numerics.log 2>&1; echo "rc=$?" >> gpurun_out/tc_numerics.log; tail -8 gpurun_out/tc_numerics.log
for cfg in "20000 512" "100003 700" "4099 300" "1000000 5000"; do
timeout 120 python tools/tc_check.py parity $cfg >> gpurun_out/tc_parity.log 2>&1; echo "rc=$?" >> gpurun_out/tc_parity.log
done
tail -12 gpurun_out/tc_parity.log
timeout 120 python tools/tc_check.py time 1000000 5000 > gpurun_out/tc_time.log 2>&1; echo "rc=$?" >> gpurun_out/tc_time.log; tail -6 gpurun_out/tc_time.log
nvidia-smi --query-gpu=name,memory.used --format=csv
