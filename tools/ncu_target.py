"""Small fixed workload for ncu: full-size turbo engine, B=24: log-mel, one encoder pass, a few eager
decode steps.  Usage: python tools/ncu_target.py [B] [n_decode_steps] [encode 0/1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import helpers
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.engine import WhisperEngine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n_dec = int(sys.argv[2]) if len(sys.argv) > 2 else 4
do_enc = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dims = WhisperDims.large_v3_turbo()
eng = WhisperEngine(dims, helpers.random_state_dict(dims, 0, "hf"), device="cuda:0", max_batch=B)
eng.load_pcm([helpers.synth_clip(i) for i in range(B)])
eng.use_graphs = False
eng.finish_check_every = 0
eng.features(B)
if do_enc:
    eng.encode(B)
prompts = torch.tensor([[50258, -1, 50360]] * B, dtype=torch.int32)
eng.decode(B, prompts, n_steps=n_dec)
torch.cuda.synchronize()
print("done", eng.stats)
