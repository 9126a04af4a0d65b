mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_train_r2b.csv python bench.py --workload train --steps 1 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/ncu_t2b.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fwd_r2b.csv python bench.py --workload forward --steps 1 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/ncu_f2b.log 2>&1; echo "ncu fwd rc=$?"
timeout 300 python tools/kbench.py --json gpurun_out/kbench_r2.json 2>&1 | tail -30
