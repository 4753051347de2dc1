"""The reference's scene file: 10 text lines written on key `i` and read on key `o`
(earth_viewer.py:100-105 camera part, :213-222 viewer part; read back at :107-126,:224-236).

    line 1-3  camera position / look-at / up      three floats each
    line 4    fov (tangent half-height)            line 5  aspect_scale
    line 6    exposure                             line 7  selected_crf (int)
    line 8    gamma                                line 9  sun_angle [rad]
    line 10   sun_path_rot [rad]                   (no trailing newline)
"""

KEYS = ("cam_pos", "look_at", "up", "fov", "aspect_scale", "exposure", "selected_crf", "gamma", "sun_angle", "sun_path_rot")


def load_config(path):
    with open(path) as f:
        cam = f.readline().split()
        look = f.readline().split()
        up = f.readline().split()
        cfg = {
            "cam_pos": tuple(float(x) for x in cam[:3]),
            "look_at": tuple(float(x) for x in look[:3]),
            "up": tuple(float(x) for x in up[:3]),
            "fov": float(f.readline()),
            "aspect_scale": float(f.readline()),
            "exposure": float(f.readline()),
            "selected_crf": int(float(f.readline())),  # the viewer writes an int; tolerate "12.0"
            "gamma": float(f.readline()),
            "sun_angle": float(f.readline()),
            "sun_path_rot": float(f.readline()),
        }
    return cfg


def save_config(path, cfg):
    with open(path, "w") as f:
        for k in ("cam_pos", "look_at", "up"):
            v = cfg[k]
            f.write(str(float(v[0])) + " " + str(float(v[1])) + " " + str(float(v[2])) + "\n")
        f.write(str(float(cfg["fov"])) + "\n")
        f.write(str(float(cfg["aspect_scale"])) + "\n")
        f.write(str(float(cfg["exposure"])) + "\n")
        f.write(str(int(cfg["selected_crf"])) + "\n")
        f.write(str(float(cfg["gamma"])) + "\n")
        f.write(str(float(cfg["sun_angle"])) + "\n")
        f.write(str(float(cfg["sun_path_rot"])))
