"""Headless stand-in for the reference's interactive viewer (earth_viewer.py:23-318).

The reference couples its render loop to a Taichi GGUI window.  There is no display on a GPU box, so
the loop here talks to a small *window protocol* instead:

    window.running                  bool
    window.is_pressed(key)          'w' 'a' 's' 'd' 'q' 'e' 'i' 'o' 'p' 'g', CTRL, SPACE, SHIFT, RMB
    window.get_cursor_pos()         (x, y) in [0, 1]^2
    window.controls()               {} or {"sun_angle": .., "sun_path_rot": .., "fov": .., "aspect_scale": ..,
                                     "exposure": .., "selected_crf": .., "gamma": ..}  (what the sliders did)
    window.show(image)              present a (W, H, 3) image

`ScriptedWindow` implements it from a list of per-frame events and writes the frames to a PNG
sequence or an MJPEG file; any windowing toolkit can implement the same five members.

What is mirrored, with the reference line each rule comes from:
  * Camera: WASD + CTRL/SPACE motion, speed 30 * clamp(altitude, 0, R/2) * 0.05 * dt, SHIFT x3, bounce
    back when the step ends below the surface, `q` = up along the local zenith, `e` = +Y up, RMB-drag
    rotation by 3 rad per unit of cursor travel, left = up x dir with the (-1, 0, 0) fallback
    (earth_viewer.py:23-163);
  * `i` writes and `o` reads the 10-line config.txt (earth_viewer.py:100-126, 213-236);
  * the frame buffer is reset when the camera, the sun or the projection (fov, aspect) changes and
    NOT when exposure / gamma / camera response change (earth_viewer.py:203-211, 262-314);
  * `p` saves screenshot/<main>-<timestamp>.jpg (earth_viewer.py:243-250).
"""
import math
import os
import time

import numpy as np

from .config import load_config
from .screenshot import save_screenshot, to_uint8_image

PLANET_R = 6371000.0  # lib/volume_rendering_models.py:34
CTRL, SPACE, SHIFT, RMB = "Control", " ", "Shift", "RMB"  # same spellings as ti.ui.CTRL / SPACE / SHIFT / RMB
SCREEN_RES = (1920, 1080)
UP_DIR = (0.0, 1.0, 0.0)
HELP_MSG = """
====================================================
Camera:
* Drag with your right mouse button to rotate
* Press W/A/S/D (+ Ctrl/Space, Shift = fast) to move, Q/E to change the up vector
* I / O save / load config.txt, P saves a screenshot
====================================================
"""


def _unit(v):
    v = np.asarray(v, dtype=np.float64)
    return v / math.sqrt(float(np.sum(v * v)))


def rotate_about(axis, theta, v):
    """v rotated counterclockwise by theta about axis (Rodrigues; the same rotation as the
    Euler-Rodrigues matrix of lib/math_utils.py:88-102 applied to v)."""
    k = _unit(axis)
    v = np.asarray(v, dtype=np.float64)
    c, s = math.cos(theta), math.sin(theta)
    return v * c + np.cross(k, v) * s + k * float(np.dot(k, v)) * (1.0 - c)


class Camera:
    """earth_viewer.py:23-163 on the window protocol."""

    def __init__(self, window, up, config_path="config.txt"):
        self._window = window
        self._lookat_pos = np.array((0.0, 0.0, 0.0))
        self._camera_pos = np.array((-15000000.0, 0.0, 15000000.0))
        self._up = _unit(up)
        self._last_mouse_pos = None
        self.config_path = config_path

    position = property(lambda self: self._camera_pos)
    look_at = property(lambda self: self._lookat_pos)
    up = property(lambda self: self._up)
    target_dir = property(lambda self: _unit(self._lookat_pos - self._camera_pos))

    def set_up(self, new_up):
        self._up = np.asarray(new_up, dtype=np.float64)

    def altitude(self):
        return float(np.linalg.norm(self._camera_pos)) - PLANET_R

    def _left(self, tgtdir):
        if abs(float(np.dot(self._up, tgtdir))) > 0.999:
            return np.array([-1.0, 0.0, 0.0])
        return np.cross(self._up, tgtdir)

    def update_camera(self, elapsed_time):
        moved = self._update_by_keys(elapsed_time)
        return self._update_by_mouse() or moved

    def _update_by_mouse(self):
        win = self._window
        if not win.is_pressed(RMB):
            self._last_mouse_pos = None
            return False
        pos = np.array(win.get_cursor_pos(), dtype=np.float64)
        if self._last_mouse_pos is None:
            self._last_mouse_pos = pos
            return False
        dx, dy = self._last_mouse_pos - pos
        self._last_mouse_pos = pos
        out_dir = self._lookat_pos - self._camera_pos
        left = self._left(_unit(out_dir))
        scale = 3.0
        new_out = rotate_about(left, dy * scale, rotate_about(self._up, dx * scale, out_dir))
        self._lookat_pos = self._camera_pos + new_out
        return True

    def _update_by_keys(self, elapsed_time):
        win = self._window
        tgt = self.target_dir
        left = self._left(tgt)
        step_dir = np.zeros(3)
        pressed = False
        for key, d in (("w", tgt), ("a", left), ("s", -tgt), ("d", -left), (CTRL, -self._up), (SPACE, self._up)):
            if win.is_pressed(key):
                pressed = True
                step_dir = step_dir + d
        if win.is_pressed("q"):
            pressed = True
            self.set_up(_unit(self._camera_pos))
        if win.is_pressed("e"):
            pressed = True
            self.set_up(np.array((0.0, 1.0, 0.0)))
        if win.is_pressed("i"):  # camera part of the scene file; the viewer appends the rest
            with open(self.config_path, "w") as f:
                for v in (self._camera_pos, self._lookat_pos, self._up):
                    f.write("%s %s %s\n" % (str(v[0]), str(v[1]), str(v[2])))
        if win.is_pressed("o"):
            with open(self.config_path) as f:
                rows = [f.readline().split() for _ in range(3)]
            for dst, row in zip((self._camera_pos, self._lookat_pos, self._up), rows):
                dst[:] = [float(x) for x in row[:3]]
            pressed = True
        if not pressed:
            return False
        step_dir = step_dir * 0.05
        speed = 30.0 * max(min(self.altitude(), PLANET_R * 0.5), 0.0)
        if win.is_pressed(SHIFT):
            speed *= 3.0
        cam_step = step_dir * speed * elapsed_time
        self._lookat_pos = self._lookat_pos + cam_step
        self._camera_pos = self._camera_pos + cam_step
        if float(np.linalg.norm(self._camera_pos)) < PLANET_R:
            self._lookat_pos = self._lookat_pos - cam_step * 2
            self._camera_pos = self._camera_pos - cam_step * 2
        return True


class ScriptedWindow:
    """Window protocol driven by a list of per-frame events; frames go to a sink.

    events[k] is a dict: {"keys": iterable of pressed keys, "cursor": (x, y), "controls": {...}}.
    sink: None (keep the last frame only), a directory (frame_%04d.png) or a path ending in .mjpeg
    (concatenated JPEGs, playable with ffplay / any MJPEG reader).
    """

    def __init__(self, events, sink=None, quality=90):
        self.events = list(events)
        self.frame = 0
        self.sink = sink
        self.quality = quality
        self.last_image = None
        self.shown = 0
        if sink and not sink.endswith(".mjpeg"):
            os.makedirs(sink, exist_ok=True)
        elif sink:
            open(sink, "wb").close()

    @property
    def running(self):
        return self.frame < len(self.events)

    def _ev(self):
        return self.events[self.frame] if self.frame < len(self.events) else {}

    def is_pressed(self, key):
        return key in self._ev().get("keys", ())

    def get_cursor_pos(self):
        return tuple(self._ev().get("cursor", (0.5, 0.5)))

    def controls(self):
        return dict(self._ev().get("controls", {}))

    def show(self, image):
        self.last_image = image
        if self.sink:
            if self.sink.endswith(".mjpeg"):
                import io
                from PIL import Image
                buf = io.BytesIO()
                Image.fromarray(to_uint8_image(image)).save(buf, format="JPEG", quality=self.quality)
                with open(self.sink, "ab") as f:
                    f.write(buf.getvalue())
            else:
                save_screenshot(image, os.path.join(self.sink, "frame_%04d.png" % self.shown))
        self.shown += 1
        self.frame += 1


class EarthViewer:
    """earth_viewer.py:166-318: progressive render loop with the reference's reset rules."""

    RESETTING = ("sun_angle", "sun_path_rot", "fov", "aspect_scale")
    NON_RESETTING = ("exposure", "selected_crf", "gamma")

    def __init__(self, window, renderer=None, screen_res=SCREEN_RES, up=UP_DIR, textures=None, config_path="config.txt",
                 screenshot_dir="screenshot", spp_per_frame=1, clock=time.time):
        self.window = window
        self.camera = Camera(window, up=up, config_path=config_path)
        if renderer is None:
            from .renderer import Renderer
            renderer = Renderer(image_res=screen_res, up=up, textures=textures)
        self.renderer = renderer
        self.renderer.set_camera_pos(*self.camera.position)
        self.config_path = config_path
        self.screenshot_dir = screenshot_dir
        self.spp_per_frame = spp_per_frame
        self.clock = clock
        self.resets = 0
        self.frames = 0
        os.makedirs(screenshot_dir, exist_ok=True)
        self.renderer.copy_textures()
        self.state = {k: self.renderer.__getattribute__(k)[None] for k in self.RESETTING + self.NON_RESETTING}

    def load_scene(self, path):
        """What pressing `o` does with a full 10-line file, callable without a key press."""
        cfg = load_config(path)
        cam = self.camera
        cam._camera_pos[:] = cfg["cam_pos"]
        cam._lookat_pos[:] = cfg["look_at"]
        cam._up = np.array(cfg["up"], dtype=np.float64)
        for k in self.RESETTING + self.NON_RESETTING:
            self.state[k] = cfg[k]
        self._push_camera()
        self._push_state()
        self.renderer.reset_framebuffer()
        self.resets += 1

    def _push_camera(self):
        r, cam = self.renderer, self.camera
        r.set_camera_pos(*cam.position)
        r.set_look_at(*cam.look_at)
        r.set_up(*cam.up)

    def _push_state(self):
        for k, v in self.state.items():
            self.renderer.__getattribute__(k)[None] = v

    def step(self, elapsed_time):
        """One pass of the reference's `while window.running` body; returns the presented image."""
        win, r = self.window, self.renderer
        reset = False
        if self.camera.update_camera(elapsed_time):
            self._push_camera()
            reset = True
        if win.is_pressed("i"):  # the camera wrote lines 1-3 (above); append lines 4-10
            s = self.state
            with open(self.config_path, "a") as f:
                f.write("\n".join(str(x) for x in (s["fov"], s["aspect_scale"], s["exposure"], int(s["selected_crf"]), s["gamma"], s["sun_angle"])) + "\n")
                f.write(str(s["sun_path_rot"]))
        if win.is_pressed("o"):
            cfg = load_config(self.config_path)
            for k in self.RESETTING + self.NON_RESETTING:
                self.state[k] = cfg[k]
        for _ in range(self.spp_per_frame):
            r.accumulate()
        img = r.fetch_image()
        if win.is_pressed("p"):
            save_screenshot(img, os.path.join(self.screenshot_dir, "earth_viewer-%s.jpg" % time.strftime("%Y-%m-%d-%H%M%S")))
        for k, v in win.controls().items():
            if k not in self.state:
                raise KeyError("unknown control %r" % k)
            if v != self.state[k]:
                self.state[k] = v
                reset = reset or k in self.RESETTING
        self._push_state()
        if reset:
            r.reset_framebuffer()
            self.resets += 1
        win.show(img)
        self.frames += 1
        return img

    def start(self):
        print(HELP_MSG)
        elapsed = 1.0  # the reference's first frame moves with dt = 1 s
        while self.window.running:
            t = self.clock()
            self.step(elapsed)
            elapsed = self.clock() - t
        return self.frames


def main(argv=None):
    """python -m digital_earth_b200.viewer --script events.json --sink frames/ [--config "config - florida.txt"]"""
    import argparse
    import json
    from .render import parse_textures
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("--script", required=True, help="JSON list of per-frame events (see ScriptedWindow)")
    ap.add_argument("--sink", default="frames", help="directory for a PNG sequence, or a .mjpeg file")
    ap.add_argument("--config", default=None)
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--textures", default="textures")
    ap.add_argument("--spp-per-frame", type=int, default=1)
    a = ap.parse_args(argv)
    with open(a.script) as f:
        events = json.load(f)
    W, H = (int(x) for x in a.res.split("x"))
    v = EarthViewer(ScriptedWindow(events, a.sink), screen_res=(W, H), textures=parse_textures(a.textures, a.config or ""), spp_per_frame=a.spp_per_frame)
    if a.config:
        v.load_scene(a.config)
    n = v.start()
    print("%d frames, %d frame-buffer resets" % (n, v.resets))
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main())
