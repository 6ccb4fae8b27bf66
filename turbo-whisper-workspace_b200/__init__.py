"""turbo-whisper-workspace_b200 — B200-native engine for the Whisper transcription hot path of
crmorton/Turbo-Whisper-Workspace (the HF ASR pipeline behind
``AudioProcessingPipeline.process_audio(task='transcribe')``, ref:vocalis/core/audio_pipeline.py:323-369).

Host side is Python/PyTorch (device memory, streams); all arithmetic runs in hand-written sm_100a
CUDA behind the C ABI declared in ``include/twb200.h`` (``libtwb200.so``).  There is no CPU path.
"""
__version__ = "0.1.0"
