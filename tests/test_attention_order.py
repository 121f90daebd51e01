"""Host-side restatement of the tile order used by ``attention_fwd_dec_kernel`` (csrc/attention_tcgen05.cu).

The two softmax groups of a CTA own different work items; the TMA warp (K loads), the QK issuer and the
event-driven PV warp each walk the two groups' tile streams in ONE fixed interleaved order: take the group
whose turn it is, or the other one when that stream is exhausted, and flip the turn after every tile
(``ATD_PICK`` + ``turn ^= 1``).  The kernel relies on three properties of that order, checked here for all
stream lengths up to 12 x 12 tiles:

  1. every walker produces the same sequence (the 3-slot K ring is indexed by position in it);
  2. within a group the tiles appear in order (single-buffered S / P / V / O per group);
  3. the QK issued right after the first tile of each group is always the SAME group's next tile when that
     tile exists (v1 of the kernel issued "the next element of the order" in the PV iteration of (g, c) and
     needed it to be (g, c + 1)).
"""
import itertools


def merged(n_a, n_b):
    c, n, turn, out = [0, 0], [n_a, n_b], 0, []
    while c[0] < n[0] or c[1] < n[1]:
        g = turn if c[turn] < n[turn] else turn ^ 1
        out.append((g, c[g]))
        c[g] += 1
        turn ^= 1
    return out


def test_order_is_a_function_of_the_stream_lengths_only():
    for n_a, n_b in itertools.product(range(13), repeat=2):
        assert merged(n_a, n_b) == merged(n_a, n_b)
        assert len(merged(n_a, n_b)) == n_a + n_b


def test_tiles_of_a_group_appear_in_order():
    for n_a, n_b in itertools.product(range(13), repeat=2):
        seq = merged(n_a, n_b)
        for g, n in ((0, n_a), (1, n_b)):
            assert [c for gg, c in seq if gg == g] == list(range(n))


def test_balanced_streams_alternate_and_leftovers_run_alone():
    for n_a, n_b in itertools.product(range(13), repeat=2):
        seq = merged(n_a, n_b)
        k = min(n_a, n_b)
        assert [g for g, _ in seq[: 2 * k]] == [0, 1] * k
        rest = {g for g, _ in seq[2 * k:]}
        assert rest <= ({0} if n_a > n_b else {1})


def test_successor_property_of_the_qk_lookahead():
    """Pop the order once per group up front (the prologue QKs), then, walking the SAME order for PV, pop
    one more element whenever the PV's group still has a tile left: the popped element is always that
    group's next tile."""
    for n_a, n_b in itertools.product(range(13), repeat=2):
        seq, n = merged(n_a, n_b), [n_a, n_b]
        pos, issued = 0, [0, 0]

        def pop():
            nonlocal pos
            g, c = seq[pos]
            pos += 1
            issued[g] += 1
            return g, c
        if seq:
            pop()
        if any(issued[g] == 0 and n[g] > 0 for g in (0, 1)):
            pop()
        for g, c in seq:
            assert issued[g] > c  # the tile's QK was issued before its PV
            if issued[g] < n[g] and issued[g] == c + 1:
                assert pop() == (g, c + 1)
        assert pos == len(seq)
