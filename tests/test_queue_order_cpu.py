"""The order in which the trace kernel hands out (pixel, sample) work items (regenerate_slot, csrc/rtb200_trace.cuh): image rows
from the bottom up, all samples of a row before the next row, x innermost. This is a Python restatement of that index
arithmetic (host logic only, no GPU): it must be a bijection onto the [sample][pixel] sample buffer whatever the batch size,
visit the bottom row first and the top row last, and hand a warp 32 consecutive x of one (row, sample)."""
import numpy as np
import pytest


def decode(my, width, rows_local, s_count):
    x = my % width
    t = my // width
    s_local = t % s_count
    rr = t // s_count
    y_local = rows_local - 1 - rr
    lp = y_local * width + x
    return s_local, y_local, x, s_local * (rows_local * width) + lp     # last: samplebuf index


@pytest.mark.parametrize("width,rows,s_count", [(40, 30, 4), (7, 3, 1), (64, 5, 16), (2, 2, 5)])
def test_queue_order_is_a_bijection_bottom_rows_first(width, rows, s_count):
    total = width * rows * s_count
    my = np.arange(total)
    s, y, x, idx = decode(my, width, rows, s_count)
    assert sorted(idx.tolist()) == list(range(total))                    # every (sample, pixel) exactly once
    assert y[0] == rows - 1 and y[-1] == 0                               # bottom row first, top row last
    assert np.all(np.diff(y) <= 0)                                       # rows never go back down
    last_row_items = my[y == 0]
    assert last_row_items.min() == total - width * s_count              # the whole top row (all its samples) closes the queue
    w = my[: min(32, width)]
    s0, y0, x0, i0 = decode(w, width, rows, s_count)
    assert len(set(s0.tolist())) == 1 and len(set(y0.tolist())) == 1 and np.all(np.diff(i0) == 1)   # coalesced sample-buffer writes
