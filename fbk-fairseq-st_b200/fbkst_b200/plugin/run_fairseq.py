#!/usr/bin/env python
"""Launcher: fairseq's own CLI with the B200 plugin, under a current numpy / Python.

    python run_fairseq.py train    DATA --user-dir <this directory> --arch conv_transformer_big2_b200 \
        --task speech_translation_with_transcription --criterion ctc_multi_loss_b200 ...
    python run_fairseq.py generate DATA --user-dir <this directory> --path checkpoint.pt ...

It restores the numpy aliases the 2020 code base imports (compat.apply_numpy), makes sure the reference tree
(FBKST_REFERENCE_ROOT, default: the current directory) is importable, and calls ``fairseq_cli.<cmd>.cli_main``
(fairseq_cli/train.py:327, fairseq_cli/generate.py) unchanged."""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    if len(sys.argv) < 2 or sys.argv[1] not in ("train", "generate", "validate", "interactive"):
        sys.exit(__doc__)
    cmd = sys.argv.pop(1)
    sys.path.insert(0, HERE)
    import compat  # noqa: E402
    compat.apply_numpy()
    root = os.environ.get("FBKST_REFERENCE_ROOT", os.getcwd())
    if root not in sys.path:
        sys.path.insert(0, root)
    importlib.import_module("fairseq_cli.%s" % cmd).cli_main()


if __name__ == "__main__":
    main()
