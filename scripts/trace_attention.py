"""Debug aid: role timeline of the persistent attention kernel (CTA 0, first item)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxdreamer_b200 import _lib
lib = _lib.load()
L, heads, hd, seq = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (64, 8, 96, 1536)))
seq_pad = (seq + 127) // 128 * 128
Q = torch.randn(L * heads, seq_pad, hd, device="cuda").to(torch.bfloat16)
K = torch.randn_like(Q)
Vt = torch.randn(L * heads, hd, seq_pad, device="cuda").to(torch.bfloat16)
O = torch.empty(L * seq, heads * hd, device="cuda", dtype=torch.bfloat16)
def run():
    _lib.check(lib.bd_attention(_lib.ptr(Q), _lib.ptr(K), _lib.ptr(Vt), _lib.ptr(O), L, heads, hd, seq, seq_pad, hd ** -0.5, 1, 2, None))
for _ in range(3):
    run()
torch.cuda.synchronize()
tr = torch.zeros(4 * 512, dtype=torch.int64, device="cuda")
lib.bd_debug_attention_trace(_lib.ptr(tr))
run()
torch.cuda.synchronize()
lib.bd_debug_attention_trace(None)
ev = tr.cpu()[3 * 512:].view(64, 8)
t = tr.cpu()[:512].view(128, 4)
sm = tr.cpu()[512:3 * 512].view(2, 64, 8)
allv = tr.cpu()[:3 * 512]
t0 = int(allv[allv > 0].min())
bkv = 96 if hd == 96 else 128
n_kv = (seq + bkv - 1) // bkv
print("item j | MMA: waitP0 start/end, waitP1 start/end | SM0: wait start, S seen, ld done, arrive, turn granted, exp done | SM1: ...   (cycles since first stamp)")
rel = lambda x: int(x) - t0 if int(x) > 0 else -1
for idx in range(min(2 * n_kv + 4, 64)):
    print(f"{idx // n_kv:2d} {idx % n_kv:2d} | " + " ".join(f"{rel(v):7d}" for v in t[idx]) + " | " + " ".join(f"{rel(v):7d}" for v in sm[0, idx, :6]) + " | " +
          " ".join(f"{rel(v):7d}" for v in sm[1, idx, :6]))
print("item | producer: q_empty wait start / done | MMA: item start, q_full seen, k_full seen | softmax g0: last pv_done seen")
for i in range(4):
    print(f"{i:2d} | " + " ".join(f"{int(x) - t0 if int(x) > 0 else -1:7d}" for x in ev[i, :6]))
