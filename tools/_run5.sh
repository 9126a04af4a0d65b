mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_r2g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r2g.log
tail -3 gpurun_out/t_r2g.log
timeout 300 python bench.py --workload train --no-cpu-baseline --no-roofline --steps 5 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('rot train', d['value'], d['ms_per_step'], d['clocks'])"
timeout 300 python bench.py --workload sample --no-cpu-baseline --no-roofline 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('rot sample', d['value'], d['ms_per_step'], d['clocks'])"
