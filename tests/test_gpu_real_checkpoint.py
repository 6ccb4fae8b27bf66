"""Opt-in test against the reference's ONLY golden vector: ref:examples/Test1/ChrisAndAlexDiTest.flac and its
insanely-fast-whisper transcript ref:examples/Test1/output.json (3 chunks [0.0, 6.24] [6.24, 16.38] [16.38, 19.74] + text).
It needs the real openai checkpoint, which does not exist offline, so the test is skipped unless

    TWB200_WHISPER_CHECKPOINT=/path/to/openai-whisper-large-v3(-turbo)     # a local save_pretrained directory

is set (and the two reference files are readable: TWB200_EXAMPLE_FLAC / TWB200_EXAMPLE_JSON override their paths).
The FLAC (192 kHz) is decoded in-process and resampled on the GPU; chunk boundaries must agree with the golden to
0.02 s (one timestamp step) and the text exactly up to surrounding whitespace."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu

CKPT = os.environ.get("TWB200_WHISPER_CHECKPOINT", "")
FLAC = os.environ.get("TWB200_EXAMPLE_FLAC", "/root/reference/examples/Test1/ChrisAndAlexDiTest.flac")
GOLD = os.environ.get("TWB200_EXAMPLE_JSON", "/root/reference/examples/Test1/output.json")


@pytest.mark.skipif(not (CKPT and os.path.isdir(CKPT) and os.path.exists(FLAC) and os.path.exists(GOLD)),
                    reason="needs a real Whisper checkpoint (TWB200_WHISPER_CHECKPOINT) and the reference's example files")
def test_reference_example_transcript(cuda_device):
    from transformers import WhisperForConditionalGeneration, WhisperTokenizer
    from turbo_whisper_workspace_b200.pipeline import B200WhisperPipeline
    gold = json.load(open(GOLD))
    model = WhisperForConditionalGeneration.from_pretrained(CKPT)
    pipe = B200WhisperPipeline.from_hf_model(model, WhisperTokenizer.from_pretrained(CKPT), devices=[cuda_device],
                                             max_batch=4, contexts_per_device=1)
    # the call that produced the golden (insanely-fast-whisper: chunk_length_s=30, batch_size=24, return_timestamps=True)
    out = pipe(FLAC, chunk_length_s=30, batch_size=24, generate_kwargs={"task": "transcribe"}, return_timestamps=True)
    assert out["text"].strip() == gold["text"].strip()
    assert len(out["chunks"]) == len(gold["chunks"])
    for got, want in zip(out["chunks"], gold["chunks"]):
        assert got["text"].strip() == want["text"].strip()
        for a, b in zip(got["timestamp"], want["timestamp"]):
            assert abs(a - b) <= 0.02 + 1e-9
