"""Aggregate cost of a decode step when N engine contexts decode concurrently on one GPU (447 fixed steps each,
no early stop, token values irrelevant), and of 2 encoder passes next to them.  Usage: decode_concurrency.py [N]"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, helpers
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.engine import WhisperEngine
B = 24
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dims = WhisperDims.large_v3_turbo()
engs = [WhisperEngine(dims, helpers.random_state_dict(dims, 0, "hf"), device="cuda:0", max_batch=B, own_stream=True)]
for _ in range(N - 1):
    engs.append(WhisperEngine(dims, None, device="cuda:0", max_batch=B, shared_weights=engs[0].w, own_stream=True))
clips = [helpers.synth_clip(i) for i in range(B)]
prompts = torch.tensor([[50258, 50259, 50360]] * B, dtype=torch.int32)
for e in engs:
    with torch.cuda.stream(e.stream):
        e.load_pcm(clips); e.features(B); e.encode(B); e.finish_check_every = 0; e.decode(B, prompts, n_steps=8)
torch.cuda.synchronize()

def dec(e, n=447):
    with torch.cuda.stream(e.stream):
        e.decode(B, prompts, n_steps=n); e.stream.synchronize()
def enc(e, n=2):
    with torch.cuda.stream(e.stream):
        for _ in range(n): e.encode(B)
        e.stream.synchronize()
def timed(fns):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ths = [threading.Thread(target=f) for f in fns]
    [t.start() for t in ths]; [t.join() for t in ths]
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for k in range(1, N + 1):
    ms = min(timed([(lambda e=e: dec(e)) for e in engs[:k]]) for _ in range(2))
    print(f"{k} contexts decoding 447 steps each: {ms:7.1f} ms -> {ms / (447 * k) * 1e3:6.1f} us per context-step", flush=True)
ms = min(timed([(lambda e=e: enc(e)) for e in engs[:1]]) for _ in range(2))
print(f"2 encoder passes alone: {ms:.1f} ms")
if N >= 2:
    ms = min(timed([(lambda e=e: dec(e)) for e in engs[:N - 1]] + [lambda: enc(engs[N - 1], 2 * (N - 1))]) for _ in range(2))
    print(f"{N - 1} contexts decoding || 1 context doing {2 * (N - 1)} encoder passes: {ms:.1f} ms")
