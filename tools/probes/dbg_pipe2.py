import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = os.environ.get("BLOCKING", "0")
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, helpers
from oracle import whisper_ref as R
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.pipeline import B200WhisperPipeline
tok = helpers.build_tokenizer()
pipes = {}
for v in ("decisive", "varied"):
    sd = helpers.variant_state_dict(R.WhisperDims(**helpers.TINY), v)
    pipes[v] = B200WhisperPipeline(sd, WhisperDims(**helpers.TINY), tok, devices=["cuda:0"], max_batch=4)
for v, pp in pipes.items():
    for ci, e in enumerate(pp.scheduler.flat_engines):
        t = e.w["tok_emb_frag"]
        print(v, ci, tuple(t.shape), t.dtype, t.is_contiguous(), hex(t.data_ptr()), t.data_ptr() % 512, "emb", tuple(e.w["tok_emb"].shape),
              "n_parts", e.n_parts, "part_val", tuple(e.part_val.shape), "dxn", tuple(e.dxn.shape), "sup", tuple(e.sup_bits.shape), flush=True)
pcm = np.concatenate([helpers.synth_clip(0), helpers.synth_clip(1, kind="mod"), helpers.synth_clip(2, seconds=11.3, kind="mod")])
steps = sys.argv[1:] or ["d0", "d5", "d60", "v5"]
for s in steps:
    v = "decisive" if s[0] == "d" else "varied"
    st = int(s[1:])
    cl = 60 if st == 60 else 30
    st = 5 if st == 60 else st
    try:
        r = pipes[v](pcm, chunk_length_s=cl, stride_length_s=st, batch_size=24, return_timestamps=True)
        torch.cuda.synchronize()
        print(s, "ok", len(r["chunks"]), flush=True)
    except Exception as ex:
        print(s, "FAILED:", type(ex).__name__, str(ex)[:200], flush=True)
        break
