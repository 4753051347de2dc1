"""Cross-check of the oracle against the reference running on REAL Taichi (SURVEY.md 8c deliverable 3).  Skips -- loudly, with the
reason -- wherever `taichi` or the reference checkout is missing, which includes this project's build container and GPU box;
the three assumptions it pins are listed in oracle/taichi_harness.py."""
import os

import numpy as np
import pytest

from oracle import taichi_harness as th

_ok, _why = th.available()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_harness_reports_why_it_cannot_run():
    ok, why = th.available()
    assert isinstance(ok, bool) and isinstance(why, str) and why
    if not ok:
        with pytest.raises(RuntimeError):
            th.TaichiReference({}, (64, 32))


@pytest.mark.skipif(not _ok, reason=_why)
def test_runtime_semantics_the_oracle_assumes():
    p = th.probe_semantics("cpu")
    assert p["bisection_iterations"] == 8, p                       # lib/colour.py:26 -> 256 wavelength bins
    assert max(p["libm_rel_err"]) < 4e-7, p                        # fast_math stays within ~3 ulp of libm
    if "unsupported_on_this_arch" not in p["sample_lod"]:
        assert p["sample_lod"]["clamp_to_edge"] and p["sample_lod"]["weight_error"] < 1e-6, p


@pytest.mark.skipif(not _ok, reason=_why)
def test_reference_frame_matches_the_oracle_within_monte_carlo_error():
    """64x32, 256 spp, florida: the reference on Taichi (its own ti.random stream) against the oracle (Philox): independent samples of
    the same estimator, compared box-wise with the oracle's second moments."""
    import digital_earth_b200 as de
    from oracle import oracle as orc
    W, H, spp = 64, 32, 256
    tex = de.textures.synthetic(256, 128, cloud_cover=0.6, seed=3)
    cfg = de.load_config(os.path.join(ROOT, "digital-earth_b200", "assets", "configs", "config - florida.txt"))
    ref = th.TaichiReference(tex, (W, H))
    try:
        acc_t, secs = ref.render(cfg, spp)
    finally:
        ref.close()
    s = orc.Scene(tex, W, H, cam_pos=cfg["cam_pos"], look_at=cfg["look_at"], up=cfg["up"], fov=cfg["fov"], aspect_scale=cfg["aspect_scale"],
                  sun_angle=cfg["sun_angle"], sun_path_rot=cfg["sun_path_rot"])
    acc_o, acc2_o, _ = orc.render(s, spp, seed=99, second_moment=True)
    mu_t, mu_o = acc_t / spp, acc_o / spp
    var = np.maximum(acc2_o / spp - mu_o ** 2, 0) / spp
    box = lambda a: a.reshape(H // 8, 8, W // 8, 8, 3).mean((1, 3))  # noqa: E731
    z = (box(mu_t) - box(mu_o)) / np.sqrt(2.0 * box(var) / 64.0 + 1e-16)
    print("taichi (%s) %.1f s for %d paths; mean %.6g vs oracle %.6g; box z: mean %.2f, |z|>4: %.1f%%"
          % (ref.arch, secs, W * H * spp, mu_t.mean(), mu_o.mean(), z.mean(), 100 * np.mean(np.abs(z) > 4)))
    assert abs(mu_t.mean() - mu_o.mean()) < 0.03 * mu_o.mean()
    assert abs(z.mean()) < 0.5 and np.mean(np.abs(z) > 4.0) < 0.03
