rm -f gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
NCU="ncu --profile-from-start off --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_r01c.csv python profiles/ncu_driver.py > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
timeout 900 $NCU --set full --import-source on -k regex:attention_kernel -s 3 -c 1 -o gpurun_out/ncu_attn64 -f python profiles/ncu_driver.py > gpurun_out/ncu_attn.log 2>&1; echo "attn rc=$?"
timeout 900 $NCU --set full --import-source on -k regex:decode_attn_kernel -s 3 -c 1 -o gpurun_out/ncu_dattn -f python profiles/ncu_driver.py > gpurun_out/ncu_dattn.log 2>&1; echo "dattn rc=$?"
