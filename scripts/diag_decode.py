"""Dump the scores + kernel decode of the exact-decode test case for offline analysis."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bonito_b200 import synth
from bonito_b200.crf.model import Model
from bonito_b200.decode import beam_search
spec = synth.model_spec("hac", n_lstm=5)
weights = synth.make_weights(spec, seed=25)
model = Model(synth.model_config(spec))
model.load_state_dict(synth.state_dict_from_weights(spec, weights))
model.use_koi(batchsize=32, chunksize=1998, quantize=False)
model = model.half().eval().to("cuda")
x = synth.squiggle(6, 3996, seed=8).half().cuda()
with torch.inference_mode():
    scores = model(x)
    seq, qstring, moves = beam_search(scores, scale=1.05, offset=0.2)
np.savez_compressed("gpurun_out/diag_decode.npz", scores=scores.cpu().numpy(), seq=seq.numpy(), q=qstring.numpy(), moves=moves.numpy())
print("saved")
