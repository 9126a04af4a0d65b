set -x
python -m pytest tests/test_gpu_train.py -m gpu -q -k "two_rank" > gpurun_out/c7_dp.log 2>&1; echo "dp rc=$?" > gpurun_out/c7_status.txt
run() { name=$1; shift; env "$@" python bench.py --gpus 2 --workload train --steps 10 --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/c7_$name.json 2> gpurun_out/c7_$name.err; echo "$name rc=$?" >> gpurun_out/c7_status.txt; }
python bench.py --gpus 1 --workload train --steps 10 --warmup 5 --no-roofline --no-cpu-baseline > gpurun_out/c7_n1.json 2> gpurun_out/c7_n1.err
run bf16 MAPDIT_GRAD_REDUCE=bf16
run fp32 MAPDIT_GRAD_REDUCE=fp32
run bf16_cta8 MAPDIT_GRAD_REDUCE=bf16 NCCL_MAX_CTAS=8
run bf16_cta16 MAPDIT_GRAD_REDUCE=bf16 NCCL_MAX_CTAS=16
run bf16_cta4 MAPDIT_GRAD_REDUCE=bf16 NCCL_MAX_CTAS=4
run fp32_cta8 MAPDIT_GRAD_REDUCE=fp32 NCCL_MAX_CTAS=8
cat gpurun_out/c7_status.txt
for f in n1 bf16 fp32 bf16_cta8 bf16_cta16 bf16_cta4 fp32_cta8; do python -c "
import json,sys
d=json.load(open('gpurun_out/c7_$f.json')); print('$f', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], d['validation'].get('loss_last'))"; done
tail -3 gpurun_out/c7_dp.log
