"""Texture manifest and ingest (lib/textures.py:1-79, renderer.py:60-94).

File names / resolutions per TEXTURE_QUALITY are the reference's.  `load_directory` reads the
NASA maps when a user has them; `synthetic` builds the procedural stand-ins used by every test
and bench here (the real maps are not redistributable / not available offline).
Arrays handed to the C-ABI are uint8 [h][w][c] with row 0 = south (v = 0).
"""
import os

import numpy as np

from . import synth

TEXTURE_QUALITY = 2  # lib/textures.py:1
TEX_RES_4K, TEX_RES_8K, TEX_RES_10K = (3840, 1920), (8100, 4050), (10800, 5400)
TEX_RES_16K, TEX_RES_21K = (16200, 8100), (21600, 10800)
CIE_LUT_RES = (441, 2)
O3_CROSSEC_LUT_RES = 441
SLOTS = synth.SLOTS

# slot -> (file, resolution) for quality 0 / 1 / 2   (lib/textures.py:10-79)
MANIFEST = {
    0: {"albedo": ("earth_color_4K.png", TEX_RES_4K), "topography": ("topography_4K.png", TEX_RES_4K),
        "ocean": ("earth_landocean_4K.png", TEX_RES_4K), "clouds": ("earth_clouds_4K.png", TEX_RES_4K),  # CLOUDS_TEX_RES is undefined
        "bathymetry": ("earth_bathymetry_4k.png", TEX_RES_4K), "emissive": ("earth_nightlights_4K.png", TEX_RES_4K),  # upstream at quality 0
        "stars": ("stars_8K.jpg", TEX_RES_8K)},
    1: {"albedo": ("earth_color_10K.png", TEX_RES_10K), "topography": ("topography_10K.png", TEX_RES_10K),
        "ocean": ("earth_landocean_8K.png", TEX_RES_8K), "clouds": ("earth_clouds_8K.png", TEX_RES_8K),
        "bathymetry": ("earth_bathymetry_10k.png", TEX_RES_10K), "emissive": ("earth_nightlights_10K.png", TEX_RES_10K),
        "stars": ("stars_16K.png", TEX_RES_16K)},
    2: {"albedo": ("earth_color_21K.png", TEX_RES_21K), "topography": ("topography_21K.png", TEX_RES_21K),
        "ocean": ("earth_landocean_16K.png", TEX_RES_16K), "clouds": ("earth_clouds_21K.png", TEX_RES_21K),
        "bathymetry": ("earth_bathymetry_21k.png", TEX_RES_21K), "emissive": ("earth_nightlights_21K.png", TEX_RES_21K),
        "stars": ("stars_16K.png", TEX_RES_16K)},
}
RGB_SLOTS = ("albedo", "stars")


def from_image_array(img, rgb):
    """Decoded image (H, W[, C]) with row 0 = TOP  ->  uint8 [h][w][c], row 0 = south."""
    a = np.asarray(img)
    if a.ndim == 2:
        a = a[:, :, None]
    a = a[::-1, :, :3] if rgb else a[::-1, :, :1]
    if rgb and a.shape[2] == 1:
        a = np.repeat(a, 3, axis=2)
    return np.ascontiguousarray(a, dtype=np.uint8)


def load_directory(directory="textures", quality=TEXTURE_QUALITY):
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    out = {}
    for slot, (fname, res) in MANIFEST[quality].items():
        path = os.path.join(directory, fname)
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: the reference expects the NASA maps in textures/ (README.md:31-32). "
                "Pass textures=digital_earth_b200.textures.synthetic(...) to render procedural stand-ins.")
        out[slot] = from_image_array(np.array(Image.open(path)), slot in RGB_SLOTS)
    return out


def synthetic(width=2048, height=1024, cloud_cover=0.5, hurricane=False, seed=0):
    return synth.make_textures(width, height, cloud_cover=cloud_cover, hurricane=hurricane, seed=seed)


def load_luts(assets_dir=None):
    assets_dir = assets_dir or os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
    z = np.load(os.path.join(assets_dir, "luts.npz"))
    return {"cie": z["cie"], "srgb2spec": z["srgb2spec"], "o3": z["o3"], "crf": z["crf"], "crf_names": [str(s) for s in z["crf_names"]]}


def load_crf_directory(directory):
    """renderer.py:147-167 on a user directory of .rf/.txt curves; Neutral first, the rest SORTED
    (the reference uses os.listdir order, which is filesystem dependent)."""
    names = sorted(n for n in os.listdir(directory) if (n.endswith(".txt") or n.endswith(".rf")) and "README" not in n)
    names.insert(0, names.pop(names.index("Neutral.rf")))
    data = []
    for n in names:
        with open(os.path.join(directory, n)) as f:
            data.append([list(map(float, ln.split()))[1:] for ln in f.readlines()])
    return np.asarray(data, dtype=np.float32), names  # (n, 1024, 3)
