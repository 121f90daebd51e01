"""Host-side restatement of the guard-row rule of fbkst_ctc_compress (csrc/ctc.cu).

After CTC compression the remaining layers run with a device-side row limit of max_new_len * B rows (row
m = t * B + b) which the row-limited GEMMs round up to their 256-row (CTA pair) or 128-row (single CTA) tiles:
rows up to the end of the last tile are read-modify-written by every residual epilogue.  fbkst_ctc_compress zeroes
G = ceil(512 / B) + 1 time steps past max_new_len, which must cover every such row for every batch size, or the
rows would carry values from one forward to the next (they grew to inf within ~70 forwards)."""


def guard(B):
    return (512 + B - 1) // B + 1


def last_time_step_touched(max_new, B, tile):
    rows = max_new * B
    rows_rounded = (rows + tile - 1) // tile * tile
    return (rows_rounded - 1) // B  # time step of the last row of the last tile


def test_guard_covers_the_tile_rounding_for_every_batch_size():
    for B in list(range(1, 130)) + [192, 256, 384, 512, 1000]:
        for max_new in (1, 2, 7, 86, 126, 130, 173, 375, 1499):
            for tile in (128, 256):
                assert last_time_step_touched(max_new, B, tile) < max_new + guard(B), (B, max_new, tile)


def test_guard_is_not_wasteful_at_the_bench_shapes():
    assert guard(64) == 9 and guard(8) == 65 and guard(48) == 12
