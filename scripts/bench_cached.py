"""Queries sharing one reference set (video / demo setting): the references' encoder tokens are
computed once (`encode_references`), each step encodes only the B query crops.  CUDA events, inputs resident, bf16."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxdreamer_b200 import BoxDreamer, synth
from boxdreamer_b200.config import make_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
R = 5
m = BoxDreamer(make_config(224), precision="bf16")
m.load_state_dict(synth.synth_decoder_state_dict(0), strict=True)
m.rgb_encoder.model.load_state_dict(synth.synth_dino_state_dict(0), strict=True)
m = m.cuda().eval()
d = synth.synth_inputs(B, R + 1, 224, seed=56, dtype=torch.bfloat16)
cache = m.encode_references(d["images"][0, :R].cuda(), d["bbox_feat"][0, :R].cuda())
q, X, K = d["images"][:, R].cuda().contiguous(), d["bbox_3d"][:, R].cuda(), d["non_ndc_intrinsics"][:, R].cuda()
for _ in range(3):
    m.forward_with_references(q, cache, X, K)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    out = m.forward_with_references(q, cache, X, K)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(json.dumps({"metric": "queries_per_sec (shared reference set, cached encoder tokens)", "value": B / ms * 1e3, "ms_per_step": ms,
                  "queries": B, "reference_views": R, "dtype": "bf16"}))
