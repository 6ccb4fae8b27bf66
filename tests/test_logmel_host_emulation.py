"""CPU check of the log-mel KERNEL's own arithmetic (csrc/logmel.cu): tests/csrc/logmel_host_test.cu includes the kernel
source with TW_HOST_TEST and runs its `__host__ __device__` phases — sample staging with reflect padding, the
single-frame radix 8 / 5 / 5 FFT in a per-warp workspace, real post-process + power, sparse mel projection, log10, the
store-then-raise-to-the-floor order of the one-launch kernel — serially on the host in the kernel's own warp / lane
structure.  Compared with the numpy oracle (pinned to transformers by tests/test_oracle_golden.py) at the stated
log-mel tolerance.  The binary is built by __graft_entry__.build() into oracle/_ref/."""
import os
import subprocess

import numpy as np
import pytest

import helpers
from oracle import logmel_ref as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "logmel_host_test")


@pytest.fixture(scope="module")
def binary():
    if not os.path.exists(BIN):
        import __graft_entry__ as g
        g.build_host_tests()
    if not os.path.exists(BIN):
        pytest.skip("host emulation binary not built (no nvcc)")
    return BIN


@pytest.mark.parametrize("name,clip", [
    ("noise30", lambda: helpers.synth_clip(0)),
    ("mod11", lambda: helpers.synth_clip(2, seconds=11.3, kind="mod")),      # zero tail: the per-clip floor is active
    ("tiny", lambda: helpers.synth_clip(5, seconds=0.05)),
    ("silence", lambda: np.zeros(1000, np.float32)),
])
def test_kernel_phases_on_host_match_oracle(binary, tmp_path, name, clip):
    from turbo_whisper_workspace_b200.ops import slaney_mel_filters
    x = clip()
    pcm, fb, out = tmp_path / "pcm.f32", tmp_path / "fb.f32", tmp_path / "out.f32"
    x.astype(np.float32).tofile(pcm)
    slaney_mel_filters().astype(np.float32).tofile(fb)
    subprocess.run([binary, str(pcm), str(len(x)), str(fb), str(out)], check=True)
    got = np.fromfile(out, dtype=np.float32).reshape(128, 3000)
    want = L.log_mel(x)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
