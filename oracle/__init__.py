"""CPU oracle for the Whisper transcription hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and only as the checker / the timed CPU baseline.  The product path
(``turbo-whisper-workspace_b200``) never imports this package and fails loudly without its CUDA
library.

The reference (crmorton/Turbo-Whisper-Workspace) delegates all arithmetic of the path to the
third-party ``transformers`` package (pinned ``==4.54.1`` in ref:.devcontainer/requirements.txt:6;
``>=4.30.0`` in ref:requirements.txt:4), which is not vendored under /root/reference.  The modules
here restate the published algorithm of that dependency (file:line citations are for the installed
transformers 5.5.0, written ``$TF/``) and are pinned against outputs of the installed library:
``tests/golden/make_golden.py`` generates the committed fixtures, ``tests/test_oracle_*.py`` check
the restatement against them (and live against transformers when importable).

Parity status: the reference's own repository holds no test, golden vector or fixture for this
path (ref:tests/__init__.py is empty; ref:examples/Test1/output.json needs the real checkpoint and a
FLAC decoder, neither available offline).  The oracle is therefore pinned against the *installed
third-party implementation* (transformers 5.5.0) rather than against reference-owned vectors.
"""
