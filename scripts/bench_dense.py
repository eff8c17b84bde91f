"""Dense-reference path timed end to end through the drop-in module on one GPU:
B queries x R reference views, DINO pre-selection -> multi-round over sub-batches of 5 -> pooled robust PnP -> fine pass.
CUDA events on the current stream, inputs resident in HBM, random-init weights (synth seed 0), bf16 tensor path."""
import json, os, sys, copy
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from boxdreamer_b200 import BoxDreamer, synth
from boxdreamer_b200.config import make_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
R = int(sys.argv[2]) if len(sys.argv) > 2 else 16
steps, warm = 5, 2
cfg = make_config(224)
cfg["modules"]["dense_cfg"].update(dict(enable=True, filter_enable=True, filter="dino", filter_topk=10, multi_round=True,
                                        sub_batch_size=5, fine_level=True, fine_topk=5, dense_mem_friendly=False))
m = BoxDreamer(cfg, precision="bf16")
m.load_state_dict(synth.synth_decoder_state_dict(0), strict=True)
m.rgb_encoder.model.load_state_dict(synth.synth_dino_state_dict(0), strict=True)
m = m.cuda().eval()
m.write_pred_bbox = True
data = synth.synth_inputs(B, R + 1, 224, seed=55)
dev = {k: (v.to(torch.bfloat16).cuda() if torch.is_tensor(v) and v.is_floating_point() else (v.cuda() if torch.is_tensor(v) else v))
       for k, v in data.items()}
def run():
    return m({k: v for k, v in dev.items()})
for _ in range(warm):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    out = run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print(json.dumps({"metric": "queries_per_sec (dense multi-round)", "value": B / ms * 1e3, "ms_per_step": ms, "queries": B,
                  "reference_views": R, "dense_cfg": dict(cfg["modules"]["dense_cfg"]), "dtype": "bf16",
                  "views_after_fine_pass": int(out["images"].shape[1])}))
