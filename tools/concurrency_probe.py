"""How well do two engine contexts overlap on one GPU?  decode||decode, decode||encode, vs alone."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.engine import WhisperEngine
B = 24
dims = WhisperDims.large_v3_turbo()
e0 = WhisperEngine(dims, helpers.random_state_dict(dims, 0, "hf"), device="cuda:0", max_batch=B, own_stream=True)
e1 = WhisperEngine(dims, None, device="cuda:0", max_batch=B, shared_weights=e0.w, own_stream=True)
clips = [helpers.synth_clip(i) for i in range(B)]
prompts = torch.tensor([[50258, -1, 50360]] * B, dtype=torch.int32)
for e in (e0, e1):
    with torch.cuda.stream(e.stream):
        e.load_pcm(clips); e.features(B); e.encode(B); e.finish_check_every = 0; e.decode(B, prompts, n_steps=8)
torch.cuda.synchronize()

def dec(e, n=447):
    with torch.cuda.stream(e.stream):
        e.decode(B, prompts, n_steps=n); e.stream.synchronize()
def enc(e, n=3):
    with torch.cuda.stream(e.stream):
        for _ in range(n): e.encode(B)
        e.stream.synchronize()
def timed(fns):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ths = [threading.Thread(target=f) for f in fns]
    [t.start() for t in ths]; [t.join() for t in ths]
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
print("decode alone        ms", round(timed([lambda: dec(e0)]), 1))
print("decode || decode    ms", round(timed([lambda: dec(e0), lambda: dec(e1)]), 1))
print("3x encode alone     ms", round(timed([lambda: enc(e0)]), 1))
print("decode || 3x encode ms", round(timed([lambda: dec(e0), lambda: enc(e1)]), 1))
# same-thread issue of both decodes (no GIL contention): alternate replays on the two streams
g0, g1 = e0._graph_for(B), e1._graph_for(B)
torch.cuda.synchronize(); t0 = time.perf_counter()
for s in range(447):
    with torch.cuda.stream(e0.stream): g0.replay()
    with torch.cuda.stream(e1.stream): g1.replay()
torch.cuda.synchronize(); print("decode || decode, single host thread ms", round((time.perf_counter() - t0) * 1e3, 1))
