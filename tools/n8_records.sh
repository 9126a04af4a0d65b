#!/bin/bash
# Round-2 records at 8 GPUs (one node, torchrun via bench.py --gpus 8): BASELINE configs 3 (the bench line), 5 and 4.
set -x
python bench.py --gpus 8 --steps 10 --warmup 5 > gpurun_out/n8_c3_default.json 2> gpurun_out/n8_c3_default.err; echo "c3 rc=$?" > gpurun_out/n8_status.txt
python bench.py --gpus 8 --model DiT-XL/2 --input-size 64 --batch 32 --workload train --steps 5 --warmup 3 --no-roofline > gpurun_out/n8_c5_train.json 2> gpurun_out/n8_c5.err; echo "c5 rc=$?" >> gpurun_out/n8_status.txt
python bench.py --gpus 8 --model DiT-L/2 --batch 64 --workload sample --sampling-steps 250 --modulation adaln --steps 2 --warmup 3 --no-roofline > gpurun_out/n8_c4_map.json 2> gpurun_out/n8_c4_map.err; echo "c4 map rc=$?" >> gpurun_out/n8_status.txt
python bench.py --gpus 8 --model DiT-L/2 --batch 64 --workload sample --sampling-steps 250 --flags-off --steps 2 --warmup 3 --no-roofline > gpurun_out/n8_c4_off.json 2> gpurun_out/n8_c4_off.err; echo "c4 off rc=$?" >> gpurun_out/n8_status.txt
cat gpurun_out/n8_status.txt
