timeout 120 python tools/tc_check.py numerics 2>&1 | tail -4
for cfg in "20000 512" "100003 700" "4099 300" "777 256" "1000000 5000"; do
timeout 120 python tools/tc_check.py parity $cfg 2>&1 | tail -2
done
TC_VARIANTS=8,9,12 timeout 100 python tools/tc_check.py cta 1000000 5000 2>&1 | grep -v "est CTAs\|blocks 0" | tail -9
TC_VARIANTS=0,1,4 timeout 120 python tools/tc_check.py time 1000000 5000 2>&1 | tail -5
