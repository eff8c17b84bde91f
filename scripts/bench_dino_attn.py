import os, sys, subprocess, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.bench_kernels import attention
print(json.dumps({"peel" if not os.environ.get("BD_ATT2_NOPEEL") else "nopeel": attention(384, 12, 64, 261, variants=(2,))}))
