"""Launches the attention kernel a few times at one shape (for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxdreamer_b200 import _lib
lib = _lib.load()
L, heads, hd, seq, variant = (int(x) for x in sys.argv[1:6])
seq_pad = (seq + 127) // 128 * 128
Q = torch.randn(L * heads, seq_pad, hd, device="cuda").to(torch.bfloat16)
K = torch.randn_like(Q)
Vt = torch.randn(L * heads, hd, seq_pad, device="cuda").to(torch.bfloat16)
O = torch.empty(L * seq, heads * hd, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    _lib.check(lib.bd_attention(_lib.ptr(Q), _lib.ptr(K), _lib.ptr(Vt), _lib.ptr(O), L, heads, hd, seq, seq_pad, hd ** -0.5, 1, variant, None))
torch.cuda.synchronize()
