"""GEMM micro-benchmark sweep (CUDA events, isolated kernels) used to locate what bounds gemm_tc / gemm_tc2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mapdit_b200 import _lib, ops


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    e[0].record()
    for i in range(iters):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(e[i].elapsed_time(e[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]


M = 65536
mk = lambda *s: (torch.randn(*s, device="cuda") * 0.05).bfloat16()
for two in (0, 1):
    _lib.set_option("gemm_2cta", two)
    for (N, K) in [(3072, 768), (768, 3072), (2304, 768), (3072, 3072), (768, 768), (1024, 1024), (4096, 1024)]:
        a, b = mk(M, K), mk(N, K)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        out2 = torch.empty_like(out)
        res = {}
        res["store"] = timeit(lambda: ops.gemm_bf16(a, b, out))
        res["mpsilu"] = timeit(lambda: ops.gemm_bf16(a, b, out, epilogue=_lib.EPI_MPSILU))
        res["mpsilu+z"] = timeit(lambda: ops.gemm_bf16(a, b, out, epilogue=_lib.EPI_MPSILU, out2=out2))
        if N % 192 == 0:
            res["qknorm"] = timeit(lambda: ops.gemm_bf16(a, b, out, epilogue=_lib.EPI_QKNORM, tokens=256, head_dim=64, qk_cols=2 * N // 3))
        fl = 2.0 * M * N * K
        print(f"2cta={two} N={N} K={K}: " + "  ".join(f"{k} {v*1e3:.0f}us {fl/v/1e9:.0f}TF" for k, v in res.items()), flush=True)
    # cuBLAS reference point for the same shapes (library GEMM, plain store)
    if two == 1:
        for (N, K) in [(3072, 768), (768, 3072), (2304, 768), (3072, 3072)]:
            a, b = mk(M, K), mk(N, K)
            t = timeit(lambda: torch.matmul(a, b.t()))
            print(f"cublas N={N} K={K}: {t*1e3:.0f}us {2.0*M*N*K/t/1e9:.0f}TF", flush=True)
