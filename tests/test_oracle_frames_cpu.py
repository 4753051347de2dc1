"""CPU check of the frozen oracle frames the GPU image gate compares against (tests/golden/oracle_frames_4096spp_v1.npz): the file holds
all three views with the parameters the GPU test uses, and a window of every view, re-rendered live by the oracle, reproduces it bit for bit
(DE_ORACLE_FRAMES_ALL=1 re-renders the whole frames, ~12 minutes on 8 cores)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import gen_oracle_frames as gof  # noqa: E402


def test_fixture_holds_every_view_with_the_gate_parameters():
    z = np.load(gof.OUT)
    assert tuple(z["params"]) == (128, 64, 256, 128, 4096, 4242)
    for view in gof.VIEWS:
        acc, acc2 = gof.load(view)
        assert acc.shape == acc2.shape == (64, 128, 3) and acc.dtype == acc2.dtype == np.float32
        assert np.isfinite(acc).all() and np.isfinite(acc2).all() and acc.mean() > 0   # single channels may be negative: out-of-gamut wavelengths
        # Cauchy-Schwarz per pixel: (sum x)^2 <= n sum x^2
        assert (acc.astype(np.float64) ** 2 <= 4096.0 * acc2.astype(np.float64) * (1 + 1e-4) + 1e-12).all()


@pytest.mark.parametrize("view", gof.VIEWS)
def test_live_oracle_render_reproduces_the_frozen_frame(view):
    """A pixel's sum depends on nothing but its own (pixel, sample) keys, so a window of the frame re-rendered live must equal the same
    window of the frozen frame bit for bit.  Default: a 16x8 window in the middle of the frame (1/64 of it, a few seconds per view);
    DE_ORACLE_FRAMES_ALL=1: the whole frame."""
    orc, s = gof.scene_for(view)
    win = (0, 0, gof.W, gof.H) if os.environ.get("DE_ORACLE_FRAMES_ALL") else (56, 28, 16, 8)
    acc, acc2, _ = orc.render(s, gof.SPP, seed=gof.SEED, second_moment=True, window=win)
    want, want2 = gof.load(view)
    x0, y0, w, h = win
    sl = (slice(y0, y0 + h), slice(x0, x0 + w))
    assert np.abs(want[sl]).sum() > 0
    assert np.array_equal(acc[sl], want[sl]) and np.array_equal(acc2[sl], want2[sl])
