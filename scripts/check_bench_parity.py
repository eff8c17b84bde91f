"""Stand-alone check of bench.py's pose_err_vs_reference leg (exact GPU path vs the oracle on a 1-query sample)."""
import sys, importlib.util, torch
sys.path.insert(0, '/root/repo')
spec = importlib.util.spec_from_file_location("bench", "/root/repo/bench.py"); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
from boxdreamer_b200 import synth
from oracle import boxdreamer_oracle as O
dec, dino = synth.synth_decoder_state_dict(0), synth.synth_dino_state_dict(0)
data = synth.synth_inputs(1, 2, 224, seed=1235)
with torch.no_grad():
    ref = O.forward(data, dec, dino)
print(b.parity_vs_oracle(data, ref, dec, dino))
