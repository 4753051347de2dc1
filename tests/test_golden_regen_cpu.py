"""The committed golden fixtures are reproducible: when the reference checkout is present (the build container; never the
GPU box), run the generators -- which import the UNMODIFIED /root/reference/{pathtracer,renderer}.py and lib/*.py on the
Taichi stand-in of oracle/ti_shim -- into a scratch file and require every array to equal the committed .npz bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DE_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "pathtracer.py")), reason="reference checkout not present (GPU box)")


@pytest.mark.parametrize("script,env_key,fixture", [("gen_golden.py", "DE_GOLDEN_OUT", "golden_v1.npz"),
                                                     ("gen_golden_preview.py", "DE_GOLDEN_PREVIEW_OUT", "golden_preview_v1.npz")])
def test_fixture_regenerates_bit_for_bit(tmp_path, script, env_key, fixture):
    out = str(tmp_path / fixture)
    env = dict(os.environ, **{env_key: out})
    env.pop("DE_GOLDEN_PATHS", None)                      # the default (64 paths per view) is what the fixture was made with
    subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", script)], check=True, env=env, capture_output=True, timeout=900)
    new, old = np.load(out), np.load(os.path.join(ROOT, "tests", "golden", fixture))
    assert sorted(new.files) == sorted(old.files)
    bad = [k for k in old.files if new[k].shape != old[k].shape or new[k].dtype != old[k].dtype or new[k].tobytes() != old[k].tobytes()]
    assert not bad, "arrays differ from the committed fixture: %s" % bad
