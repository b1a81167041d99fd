for mode in h2 b3; do for bk in 32 64; do
  B2_FREDHOLM_MODE=$mode B2_FREDHOLM_BK=$bk timeout 200 python profiles/fredholm_tc_check.py --time > gpurun_out/fr_${mode}_${bk}.log 2>&1; echo "rc=$?" >> gpurun_out/fr_${mode}_${bk}.log
done; done
B2_FREDHOLM_MODE=h2 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fr_launches_h2.csv python profiles/fredholm_tc_check.py --time --time-only > gpurun_out/fr_ncu.log 2>&1
for f in gpurun_out/fr_h2_32.log gpurun_out/fr_h2_64.log gpurun_out/fr_b3_32.log gpurun_out/fr_b3_64.log; do echo == $f; grep -c OK $f; grep "FAIL\|us \|rc=\|Error\|error" $f | head -20; done
