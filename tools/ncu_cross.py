"""Decode cross-attention kernel alone (for ncu --set full): turbo shapes, B = 24."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
sys.argv = [sys.argv[0]]
import bench, helpers
from turbo_whisper_workspace_b200.config import WhisperDims
from turbo_whisper_workspace_b200.engine import WhisperEngine
dims = WhisperDims.large_v3_turbo()
eng = WhisperEngine(dims, helpers.random_state_dict(dims, 0, "hf"), device="cuda:0", max_batch=24)
eng.ckv.normal_()
ms, nbytes = bench.cross_attn_roofline_probe(eng, 24, iters=8)
print("cross-attn ms", ms, "GB/s", nbytes / ms / 1e6)
