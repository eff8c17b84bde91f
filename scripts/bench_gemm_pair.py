"""Single-CTA vs CTA-pair GEMM (register/LSU epilogue vs TMA-out epilogue) at the step's shapes (CUDA events, operands > L2)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_kernels import gemm

SHAPES = [("proj_resid", 98304, 768, 768, 2), ("fc1_gelu", 98304, 3072, 768, 1), ("fc2_resid", 98304, 768, 3072, 2),
          ("f32", 98304, 768, 768, 0), ("act", 98304, 768, 768, 4)]
res = {}
for mode, tma in (("0", "0"), ("1", "0"), ("1", "1")):
    os.environ["BD_GEMM_PAIR"] = mode
    os.environ["BD_GEMM_TMA_EPI"] = tma
    for name, M, N, K, epi in SHAPES:
        try:
            res[f"{name}.pair{mode}.tma{tma}"] = gemm(M, N, K, epi)
        except Exception as e:
            res[f"{name}.pair{mode}.tma{tma}"] = {"error": str(e)[:200]}
            break
print(json.dumps(res, indent=1))
