"""e2e (host buffers -> poses) of bd_forward_host at BASELINE config 2 for different numbers of image chunks
(BOXDREAMER_B200_HOST_CHUNKS is read per call)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxdreamer_b200 import BoxDreamer, synth
from boxdreamer_b200.config import make_config
B, T, S = 64, 6, 224
m = BoxDreamer(make_config(S), precision="bf16")
m.load_state_dict(synth.synth_decoder_state_dict(0), strict=True)
m.rgb_encoder.model.load_state_dict(synth.synth_dino_state_dict(0), strict=True)
m = m.cuda().eval()
d = synth.synth_inputs(B, T, S, seed=1235, dtype=torch.bfloat16)
mask = torch.zeros(B, T, dtype=torch.bool); mask[torch.arange(B), d["query_idx"]] = True
hi, hb, hq = d["images"].contiguous().pin_memory(), d["bbox_feat"].contiguous().pin_memory(), d["query_idx"].pin_memory()
hK, hX = d["non_ndc_intrinsics"][mask].float().contiguous().pin_memory(), d["bbox_3d"][mask].float().contiguous().pin_memory()
eng = m._engine_for(hi.cuda(), B, T)
for ch in (1, 2, 3, 4, 6):
    os.environ["BOXDREAMER_B200_HOST_CHUNKS"] = str(ch)
    for _ in range(3):
        eng.forward_host(hi, hb, hq, hX, hK)
    t0 = time.perf_counter()
    for _ in range(10):
        eng.forward_host(hi, hb, hq, hX, hK)
    dt = (time.perf_counter() - t0) / 10
    print(f"chunks {ch}: {dt * 1e3:7.2f} ms  {B / dt:7.1f} queries/s")
