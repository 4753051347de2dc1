"""Pack the reference's DATA inputs (not code) into digital-earth_b200/assets/.

Run once in the build container (needs /root/reference); the outputs are committed so
nothing reads /root/reference at run time.  Sources:
  LUT/CIE.dat                  2x441x3 f32  (renderer.py:97-107)
  LUT/srgb2spec.dat            300x3 f16    (renderer.py:109-117)
  LUT/ozone_cross_section.dat  441 f32      (renderer.py:119-125)
  LUT/camera_response_functions/*.rf  1024 rows x (irradiance, R, G, B)  (renderer.py:147-167)
  config - *.txt               the three shipped scene files
CRF order: the reference uses os.listdir order (filesystem dependent, renderer.py:154) with
Neutral.rf moved to index 0; we fix it to sorted(names) with Neutral first.
"""
import os
import shutil
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "digital-earth_b200", "assets")


def load_rf(path):
    with open(path) as f:
        rows = [list(map(float, ln.split()))[1:] for ln in f.readlines()]
    return np.asarray(rows, dtype=np.float32)


def main():
    cie = np.fromfile(os.path.join(REF, "LUT/CIE.dat"), dtype=np.float32, count=441 * 2 * 3).reshape(2, 441, 3)
    s2s = np.fromfile(os.path.join(REF, "LUT/srgb2spec.dat"), dtype=np.float16, count=900).reshape(300, 3)
    o3 = np.fromfile(os.path.join(REF, "LUT/ozone_cross_section.dat"), dtype=np.float32, count=441)
    d = os.path.join(REF, "LUT/camera_response_functions")
    names = sorted(n for n in os.listdir(d) if (n.endswith(".rf") or n.endswith(".txt")) and "README" not in n)
    names.insert(0, names.pop(names.index("Neutral.rf")))
    crf = np.stack([load_rf(os.path.join(d, n)) for n in names])  # (n, 1024, 3)
    assert crf.shape[1:] == (1024, 3)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "luts.npz"), cie=cie, srgb2spec=s2s, o3=o3, crf=crf, crf_names=np.array(names))
    for n in os.listdir(REF):
        if n.startswith("config - ") and n.endswith(".txt"):
            shutil.copyfile(os.path.join(REF, n), os.path.join(OUT, "configs", n))
    print("packed", cie.shape, s2s.shape, o3.shape, crf.shape, names)


if __name__ == "__main__":
    main()
