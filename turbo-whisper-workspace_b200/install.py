"""Zero-touch installation into the reference: make ``transformers.pipeline("automatic-speech-recognition", ...)``
return a :class:`B200WhisperPipeline` for Whisper checkpoints.

The reference builds its transcription object inside ``AudioProcessingPipeline.load_transcription_model`` with

    from transformers import pipeline
    self.transcription_model = pipeline("automatic-speech-recognition", model=model_name, device=device,
                                        torch_dtype=torch.float16 if self.gpu_available else torch.float32)

(ref:vocalis/core/audio_pipeline.py:187-200; the import happens at call time, so an attribute patched on the
``transformers`` module before the first transcription is what it picks up).  ``install()`` patches exactly that
attribute: ASR requests for a Whisper model are answered with the B200 engine, every other task / model is forwarded to
the original factory untouched.  ``uninstall()`` restores it.  Nothing in the reference's source changes.

    import turbo_whisper_workspace_b200.install as twb
    twb.install()                       # once, at application start-up (before the first transcription)
    ...                                 # vocalis.* runs unmodified: load_transcription_model() now yields the engine
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Sequence, Tuple

_ORIGINAL: Optional[Callable[..., Any]] = None
_OPTIONS: Dict[str, Any] = {}

ASR_TASK = "automatic-speech-recognition"


def _default_loader(model: Any, tokenizer: Any = None) -> Tuple[Any, Any]:
    """``model``: a hub id / local directory (str) or a loaded WhisperForConditionalGeneration -> (model, tokenizer)."""
    from transformers import WhisperForConditionalGeneration, WhisperTokenizer
    name = model if isinstance(model, str) else getattr(getattr(model, "config", None), "_name_or_path", None)
    if isinstance(model, str):
        model = WhisperForConditionalGeneration.from_pretrained(model)
    if tokenizer is None or isinstance(tokenizer, str):
        tokenizer = WhisperTokenizer.from_pretrained(tokenizer or name)
    return model, tokenizer


def _is_whisper(model: Any) -> bool:
    if isinstance(model, str):
        return "whisper" in model.lower()
    return getattr(getattr(model, "config", None), "model_type", None) == "whisper"


def pipeline(task: Optional[str] = None, model: Any = None, *args, **kwargs):
    """Replacement for ``transformers.pipeline`` while installed.  The arguments the reference passes that only make
    sense for the library implementation (``device``, ``torch_dtype`` / ``dtype``) are accepted and ignored: the engine
    runs bf16 on the CUDA devices given to :func:`install`."""
    if _ORIGINAL is None:
        raise RuntimeError("turbo_whisper_workspace_b200.install.pipeline called while not installed")
    if task != ASR_TASK or model is None or not _is_whisper(model):
        return _ORIGINAL(task, model, *args, **kwargs)
    from .pipeline import B200WhisperPipeline
    loader = _OPTIONS.get("loader") or _default_loader
    hf_model, tokenizer = loader(model, kwargs.get("tokenizer"))
    build = _OPTIONS.get("builder")
    if build is not None:                                   # tests / custom schedulers
        pipe = build(hf_model, tokenizer)
    else:
        pipe = B200WhisperPipeline.from_hf_model(hf_model, tokenizer, devices=_OPTIONS["devices"],
                                                 max_batch=_OPTIONS["max_batch"],
                                                 contexts_per_device=_OPTIONS["contexts_per_device"])
    if hasattr(pipe, "num_beams"):
        pipe.num_beams = _OPTIONS.get("num_beams", 1)
    return pipe


def install(devices: Optional[Sequence[Any]] = None, max_batch: int = 24, contexts_per_device: int = 4,
            loader: Optional[Callable[..., Tuple[Any, Any]]] = None,
            builder: Optional[Callable[[Any, Any], Any]] = None, num_beams: int = 1) -> None:
    """Patch ``transformers.pipeline``.  ``devices``: CUDA devices of the engine (default: all visible ones);
    ``loader(model, tokenizer) -> (WhisperForConditionalGeneration, tokenizer)`` and ``builder(model, tokenizer) ->
    callable`` override checkpoint loading and engine construction.  ``num_beams``: default beam count of the returned
    pipeline's calls — 1 = greedy (the north star's mode); 5 reproduces what the reference's literal call decodes with
    under transformers >= 4.53, whose ASR pipeline defaults to ``num_beams=5`` and the reference passes only
    ``generate_kwargs={"task": task}`` (ref:vocalis/core/audio_pipeline.py:348)."""
    global _ORIGINAL
    import transformers
    if devices is None:
        import torch
        devices = [f"cuda:{i}" for i in range(max(1, torch.cuda.device_count()))]
    _OPTIONS.update(devices=list(devices), max_batch=int(max_batch), contexts_per_device=int(contexts_per_device),
                    loader=loader, builder=builder, num_beams=max(1, int(num_beams)))
    if _ORIGINAL is None:
        _ORIGINAL = transformers.pipeline
        transformers.pipeline = pipeline


def uninstall() -> None:
    global _ORIGINAL
    if _ORIGINAL is not None:
        import transformers
        transformers.pipeline = _ORIGINAL
        _ORIGINAL = None
    _OPTIONS.clear()
