ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --rollout-steps 8 --no-train --no-cpu-baseline --no-extra > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_ -s 40 -c 10 -o gpurun_out/r2_tc python bench.py --steps 1 --warmup 3 --rollout-steps 4 --no-cpu-baseline --no-train --no-extra --dtype bf16 --no-graph > gpurun_out/r2_ncu_b.log 2>&1
STEPS=6 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 220 --csv --log-file gpurun_out/r2_train_launches.csv python tools/train_eager.py > gpurun_out/r2_ncu_c.log 2>&1
DLWPCS_TC_TRACE=1 timeout 100 python tools/trace_step.py > gpurun_out/r2_trace_final.txt 2>&1
