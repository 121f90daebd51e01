"""Print the per-parameter gradient errors of the training chain vs the oracle (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fbk-fairseq-st_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import test_gpu_train as T

T.TOL = 1e9
T.COS_MIN = -2.0
cases = {
    "tiny": (dict(embed_dim=128, ffn_dim=256, heads=2, layers=3, conv_channels=64, feat_dim=40, vocab=64,
                  distance_penalty="log", ctc_layer=2, ctc_strategy="weighted"), [97, 64, 30], 3, 40),
    "mha": (dict(embed_dim=256, ffn_dim=512, heads=4, layers=2, conv_channels=64, feat_dim=40, vocab=50,
                 distance_penalty=None, ctc_layer=0, ctc_strategy="avg"), [201, 160, 77, 40], 5, 40),
    "big2": (dict(embed_dim=512, ffn_dim=2048, heads=8, layers=3, conv_channels=64, feat_dim=40, vocab=1005,
                  distance_penalty="log", ctc_layer=2, ctc_strategy="avg"), [601, 598, 411, 203], 7, 40),
}
for name in sys.argv[1:] or list(cases):
    cfg, lens, seed, feat = cases[name]
    import pytest
    w, c = T.run_case(cfg, lens, seed, feat)
    print("==", name)
    for k, v in w.items():
        print("  %-50s matched-pattern err %.4f   cosine vs fp32 oracle %.5f" % (k, v, c[k]))
