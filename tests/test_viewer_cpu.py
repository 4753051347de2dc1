"""Viewer shim: camera controller and reset rules of earth_viewer.py:23-318,
exercised on a scripted window and a recording stand-in for the renderer (no GPU)."""
import math
import os

import numpy as np
import pytest

from digital_earth_b200 import viewer as V
from digital_earth_b200 import load_config


class _Field:
    def __init__(self, v): self.v = v
    def __getitem__(self, k): return self.v
    def __setitem__(self, k, v): self.v = v


class FakeRenderer:
    def __init__(self):
        self.calls = []
        self.fov, self.aspect_scale, self.exposure = _Field(0.2356), _Field(1.0), _Field(2.5)
        self.gamma, self.selected_crf = _Field(1.0), _Field(0)
        self.sun_angle, self.sun_path_rot = _Field(math.radians(60)), _Field(math.radians(-45))
        self.spp = 0
    def set_camera_pos(self, *p): self.calls.append(("cam", p))
    def set_look_at(self, *p): self.calls.append(("look", p))
    def set_up(self, *p): self.calls.append(("up", p))
    def copy_textures(self): self.calls.append(("tex",))
    def reset_framebuffer(self): self.calls.append(("reset",)); self.spp = 0
    def accumulate(self): self.spp += 1
    def fetch_image(self): return np.full((16, 8, 3), min(1.0, self.spp / 10.0), np.float32)


def make(events, tmp_path, **kw):
    win = V.ScriptedWindow(events, **kw)
    return V.EarthViewer(win, renderer=FakeRenderer(), config_path=str(tmp_path / "config.txt"), screenshot_dir=str(tmp_path / "shots")), win


def test_rotation_is_counterclockwise_known_answer():
    # known answer of the Euler-Rodrigues formula the reference uses (lib/math_utils.py:88-102)
    got = V.rotate_about([4, 4, 1], 1.2, [3, 5, 0])
    assert np.allclose(got, [2.74911638, 4.77180932, 1.91629719], atol=1e-7)
    assert np.allclose(V.rotate_about([0, 0, 1], math.pi / 2, [1, 0, 0]), [0, 1, 0], atol=1e-12)


def test_defaults_and_idle_frames_do_not_reset(tmp_path):
    v, win = make([{}] * 5, tmp_path)
    assert np.allclose(v.camera.position, (-1.5e7, 0, 1.5e7)) and np.allclose(v.camera.look_at, 0)  # earth_viewer.py:26-27
    assert v.start() == 5 and v.resets == 0 and v.renderer.spp == 5
    assert win.last_image.shape == (16, 8, 3)


def test_forward_motion_speed_follows_altitude(tmp_path):
    v, _ = make([{"keys": ["w"]}], tmp_path)
    p0 = v.camera.position.copy()
    alt = np.linalg.norm(p0) - V.PLANET_R
    v.step(0.5)
    moved = np.linalg.norm(v.camera.position - p0)
    assert math.isclose(moved, 0.05 * 30.0 * min(alt, V.PLANET_R * 0.5) * 0.5, rel_tol=1e-12)  # earth_viewer.py:135-140
    assert np.allclose((v.camera.position - p0) / moved, -p0 / np.linalg.norm(p0))                # towards the look-at point
    assert v.resets == 1 and ("reset",) in v.renderer.calls
    v2, _ = make([{"keys": ["w", V.SHIFT]}], tmp_path)
    v2.step(0.5)
    assert math.isclose(np.linalg.norm(v2.camera.position - p0), 3 * moved, rel_tol=1e-12)


def test_strafe_uses_up_cross_dir_and_pole_fallback(tmp_path):
    v, _ = make([{"keys": ["a"]}, {"keys": ["a"]}], tmp_path)
    p0 = v.camera.position.copy()
    v.step(1.0)
    left = np.cross([0, 1, 0], V._unit(-p0))
    assert np.allclose(V._unit(v.camera.position - p0), V._unit(left))
    v.camera._camera_pos[:] = (0.0, 2.0e7, 0.0)   # looking straight down the up axis: |cos| > 0.999
    v.camera._lookat_pos[:] = 0.0
    p1 = v.camera.position.copy()
    v.step(1.0)
    assert np.allclose(V._unit(v.camera.position - p1), (-1.0, 0.0, 0.0))


def test_never_ends_below_the_surface(tmp_path):
    v, _ = make([{"keys": ["w"]}], tmp_path)
    v.camera._camera_pos[:] = (0.0, 0.0, V.PLANET_R + 1000.0)
    v.camera._lookat_pos[:] = 0.0
    v.step(10.0)   # 0.05 * 30 * 1000 m * 10 s = 15 km forward would end underground: stepped back twice
    assert np.linalg.norm(v.camera.position) >= V.PLANET_R
    assert math.isclose(np.linalg.norm(v.camera.position), V.PLANET_R + 1000.0 + 15000.0, rel_tol=1e-9)


def test_q_and_e_switch_the_up_vector(tmp_path):
    v, _ = make([{"keys": ["q"]}, {"keys": ["e"]}], tmp_path)
    v.step(1.0)
    assert np.allclose(v.camera.up, V._unit((-1.5e7, 0, 1.5e7)))
    assert ("up", tuple(v.camera.up)) in v.renderer.calls
    v.step(1.0)
    assert np.allclose(v.camera.up, (0, 1, 0)) and v.resets == 2


def test_mouse_drag_turns_the_view_not_the_position(tmp_path):
    ev = [{"keys": [V.RMB], "cursor": (0.5, 0.5)}, {"keys": [V.RMB], "cursor": (0.6, 0.5)}, {"cursor": (0.9, 0.9)}, {"keys": [V.RMB], "cursor": (0.1, 0.1)}]
    v, _ = make(ev, tmp_path)
    p0, out0 = v.camera.position.copy(), v.camera.look_at - v.camera.position
    v.step(1.0)
    assert v.resets == 0                     # first frame of a drag only latches the cursor
    v.step(1.0)
    out1 = v.camera.look_at - v.camera.position
    assert np.allclose(v.camera.position, p0) and math.isclose(np.linalg.norm(out1), np.linalg.norm(out0), rel_tol=1e-12)
    assert np.allclose(out1, V.rotate_about([0, 1, 0], -0.1 * 3.0, out0))   # dx = last - now = -0.1, 3 rad per unit
    assert v.resets == 1
    v.step(1.0); v.step(1.0)
    assert v.resets == 1                     # releasing the button forgets the drag origin


def test_reset_rules_of_the_sliders(tmp_path):
    ev = [{"controls": {"exposure": 3.0}}, {"controls": {"gamma": 2.2, "selected_crf": 3}}, {"controls": {"sun_angle": 1.0}},
          {"controls": {"fov": 0.3}}, {"controls": {"aspect_scale": 1.1}}, {"controls": {"sun_path_rot": 0.2}}, {"controls": {"sun_path_rot": 0.2}}]
    v, _ = make(ev, tmp_path)
    seen = []
    for _ in ev:
        v.step(0.03)
        seen.append(v.resets)
    assert seen == [0, 0, 1, 2, 3, 4, 4]     # earth_viewer.py:262-303: sun and projection reset, look controls do not
    r = v.renderer
    assert (r.exposure[None], r.gamma[None], r.selected_crf[None], r.sun_angle[None], r.fov[None], r.aspect_scale[None], r.sun_path_rot[None]) == (3.0, 2.2, 3, 1.0, 0.3, 1.1, 0.2)
    with pytest.raises(KeyError):
        make([{"controls": {"bogus": 1}}], tmp_path)[0].step(0.1)


def test_i_writes_and_o_reads_the_ten_line_scene_file(tmp_path):
    v, _ = make([{"keys": ["i"], "controls": {}}, {"keys": ["w"]}, {"keys": ["o"]}], tmp_path)
    v.state["exposure"] = 1.25
    v.step(1.0)
    cfg = load_config(str(tmp_path / "config.txt"))   # the same parser the CLI uses
    assert cfg["cam_pos"] == (-1.5e7, 0.0, 1.5e7) and cfg["exposure"] == 1.25 and cfg["selected_crf"] == 0
    assert open(tmp_path / "config.txt").read().count("\n") == 9          # no trailing newline (earth_viewer.py:221)
    v.step(1.0)
    assert not np.allclose(v.camera.position, (-1.5e7, 0.0, 1.5e7))
    v.state["exposure"] = 9.0
    v.step(1.0)
    assert np.allclose(v.camera.position, (-1.5e7, 0.0, 1.5e7)) and v.state["exposure"] == 1.25 and v.renderer.exposure[None] == 1.25


def test_load_scene_applies_a_shipped_config(tmp_path):
    v, _ = make([{}], tmp_path)
    path = os.path.join(os.path.dirname(V.__file__), "assets", "configs", "config - florida.txt")
    v.load_scene(path)
    cfg = load_config(path)
    assert np.allclose(v.camera.position, cfg["cam_pos"]) and v.renderer.fov[None] == cfg["fov"] and v.resets == 1


def test_sinks_png_sequence_mjpeg_and_screenshot(tmp_path):
    v, win = make([{}, {"keys": ["p"]}], tmp_path, sink=str(tmp_path / "frames"))
    v.start()
    assert sorted(os.listdir(tmp_path / "frames")) == ["frame_0000.png", "frame_0001.png"]
    shots = os.listdir(tmp_path / "shots")
    assert len(shots) == 1 and shots[0].startswith("earth_viewer-") and shots[0].endswith(".jpg")
    v2, _ = make([{}, {}, {}], tmp_path, sink=str(tmp_path / "out.mjpeg"))
    v2.start()
    data = open(tmp_path / "out.mjpeg", "rb").read()
    assert data.count(b"\xff\xd8\xff") == 3 and data.endswith(b"\xff\xd9")
