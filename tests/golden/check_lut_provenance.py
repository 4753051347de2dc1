"""Known-answer pin of the shipped LUT data: run the reference's own generator
scripts (LUT/ozone_cross_section_generator.py, LUT/srgb2spec_generator.py) in a scratch directory and
compare their output, byte for byte, with the LUT/*.dat files packed into assets/luts.npz.

Build-container only (needs /root/reference); the verdict is committed as tests/golden/lut_provenance.json
and asserted by tests/test_host_cpu.py together with the sha256 of the packed arrays.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("DE_REFERENCE", "/root/reference")


def main():
    z = np.load(os.path.join(ROOT, "digital-earth_b200", "assets", "luts.npz"))
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copy(os.path.join(REF, "LUT", "O3_cross_section_Serdyuchenko_2014.txt"), tmp)
        for script in ("ozone_cross_section_generator.py", "srgb2spec_generator.py"):
            src = open(os.path.join(REF, "LUT", script)).read().replace("import cv2\n", "")  # cv2 is imported but unused upstream
            p = os.path.join(tmp, script)
            open(p, "w").write(src)
            subprocess.check_call([sys.executable, p], cwd=tmp, stdout=subprocess.DEVNULL)
        o3 = np.fromfile(os.path.join(tmp, "ozone_cross_section.dat"), dtype=np.float32)
        s2s = np.fromfile(os.path.join(tmp, "srgb2spec.dat"), dtype=np.float16).reshape(-1, 3)
    out["ozone_regenerated_bit_exact"] = bool(o3.shape == z["o3"].shape and np.array_equal(o3.view(np.uint32), z["o3"].view(np.uint32)))
    out["srgb2spec_regenerated_bit_exact"] = bool(s2s.shape == z["srgb2spec"].shape and np.array_equal(s2s.view(np.uint16), z["srgb2spec"].view(np.uint16)))
    for k in ("cie", "srgb2spec", "o3", "crf"):
        out["sha256_" + k] = hashlib.sha256(np.ascontiguousarray(z[k]).tobytes()).hexdigest()
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "lut_provenance.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
