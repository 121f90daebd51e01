"""Host-side restatement of the penalty look-up table of ``attention_fwd_wide_kernel`` (csrc/attention_wide.cu).

The log-distance penalty (conv_transformer_layer.py:22-27) of query row i and key k is ``lut[lut_off + k - i]``.
A softmax thread owns one row and reads the 128 consecutive entries of a key tile with LDS.128, which needs
16-byte alignment although consecutive rows start one float apart: the kernel keeps FOUR copies, copy r holding
``lut[o + r]`` at index o, and a row whose first index is = r (mod 4) reads copy r at the aligned index below it.
The copies start ``r * n + pad[r]`` floats into the buffer, pad = 0 / 12 / 20 / 28: with n a multiple of 32 and
no pads, rows q+1 .. q+4 of a quarter-warp read the SAME four banks in different copies -- 16 wavefronts per
LDS.128 instead of 4, which cost the first version of the kernel 16 us of 80 (profiles/r02z_attention_wide.txt).
Checked here: alignment, bounds, values, and that the eight rows of every quarter-warp hit eight different
4-bank groups for every tile and query tile."""
import math

BM = BN = 128


def lut_floats(L):
    return ((L + BM - 1) // BM) * BM + ((L + BN - 1) // BN) * BN


def lut_copy(r, n):  # aw_lut_copy
    return r * n + (4 + 8 * r if r else 0)


def build(L):
    nq = (L + BM - 1) // BM
    n, off = lut_floats(L), nq * BM
    buf = [None] * (4 * n + 32)
    for r in range(4):
        for oo in range(n):
            d = abs(oo + r - off)
            buf[lut_copy(r, n) + oo] = -math.log2(d) if d > 1 else 0.0
    return buf, n, off


def row_base(q, q0, n, off):
    """float index of the row's first LDS.128 for key tile 0 (the kernel adds k0 = 128 * tile)"""
    i = q0 + q
    r = (4 - (q & 3)) & 3
    return lut_copy(r, n) + (off - i - r), r


def test_alignment_bounds_and_values():
    for L in (100, 375, 700, 1450):
        buf, n, off = build(L)
        nq, nkv = (L + BM - 1) // BM, (L + BN - 1) // BN
        for qt in range(nq):
            for q in range(BM):
                base, r = row_base(q, qt * BM, n, off)
                assert base % 4 == 0 and base >= 0
                for tile in range(nkv):
                    for t in (0, 1, 63, 64, 127):
                        idx = base + tile * BN + t
                        assert idx < len(buf) and buf[idx] is not None
                        d = abs(tile * BN + t - (qt * BM + q))
                        assert buf[idx] == (-math.log2(d) if d > 1 else 0.0)


def test_quarter_warps_are_bank_conflict_free():
    """LDS.128 is served a quarter-warp (8 lanes x 16 B) per wavefront when the 8 addresses fall into 8
    different groups of four banks (32 banks of 4 B)."""
    for L in (375, 1450):
        _, n, off = build(L)
        assert n % 32 == 0  # the condition under which unpadded copies collide
        for q8 in range(0, BM, 8):
            for chunk in range(0, BN, 4):  # every LDS.128 of the tile
                groups = {((row_base(q, 0, n, off)[0] + chunk) // 4) % 8 for q in range(q8, q8 + 8)}
                assert len(groups) == 8, (L, q8, chunk, groups)
        # and the unpadded layout does collide (what the first version measured)
        groups = {((r * n + (off - q - r)) // 4) % 8 for q in range(8) for r in [(4 - (q & 3)) & 3]}
        assert len(groups) < 8
