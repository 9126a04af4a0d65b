timeout 600 python -m pytest tests/test_gpu_train.py -x -q -k "attention_backward" 2>&1 | tail -2
timeout 120 python tools/kbench.py attn_bwd 2>&1 | tail -1
