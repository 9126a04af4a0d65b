mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/dp_check.py > gpurun_out/dp_check_n2_r2.log 2>&1; echo "dp_check rc=$?"; tail -3 gpurun_out/dp_check_n2_r2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_r2.json 2> gpurun_out/bench_n2_r2.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_n2_r2.json'));print('N2 train', d['value'], d['ms_per_step'], 'sample', d['sample50']['value'], 'adaln', d['adaln']['train']['value'], d['adaln']['sample50']['value'])"
