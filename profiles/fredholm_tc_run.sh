for mode in h2 b3; do
  B2_FREDHOLM_MODE=$mode timeout 200 python profiles/fredholm_tc_check.py --time > gpurun_out/fr_${mode}_32.log 2>&1; echo "rc=$?" >> gpurun_out/fr_${mode}_32.log
done
B2_FREDHOLM_MODE=h2 B2_FREDHOLM_BK=64 timeout 200 python profiles/fredholm_tc_check.py --time --time-only > gpurun_out/fr_h2_64.log 2>&1
B2_FREDHOLM_MODE=h2 B2_FREDHOLM_CONCAT=0 timeout 200 python profiles/fredholm_tc_check.py --time --time-only > gpurun_out/fr_h2_noconcat.log 2>&1
B2_FREDHOLM_MODE=h2 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fr_launches_h2.csv python profiles/fredholm_tc_check.py --time --time-only > gpurun_out/fr_ncu.log 2>&1
B2_FREDHOLM_MODE=h2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fredholm_tc_kernel -c 2 -o gpurun_out/r02_fredholm_tc_h2 -f python profiles/fredholm_tc_check.py --time --time-only > gpurun_out/fr_ncu_full.log 2>&1
for f in gpurun_out/fr_h2_32.log gpurun_out/fr_b3_32.log gpurun_out/fr_h2_64.log gpurun_out/fr_h2_noconcat.log; do echo == $f; grep -c OK $f; grep "FAIL\|us \|rc=\|Error\|error" $f | head -12; done
