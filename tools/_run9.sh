mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n8_r2.json 2> gpurun_out/bench_n8_r2.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_n8_r2.json').read());print('N8 train', d['value'], d['ms_per_step'], 'sample', d['sample50']['value'], 'adaln', d['adaln']['train']['value'], d['adaln']['sample50']['value'], d['clocks'])"
