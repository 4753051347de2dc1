"""Seeded procedural equirectangular textures.

The NASA maps the reference expects (lib/textures.py:10-27, README.md:31-32) are not
available offline, so every test / bench / smoke run uses these stand-ins at the
resolutions BASELINE.json names (2048x1024, 8192x4096).  Layout of every map:
uint8 [h][w][c], row 0 = v=0 = south pole (the reference's ti.tools.imread arrays are
[x][y] with y up; the C-ABI takes the transposed row-major form, see include/de_api.h).

Definitions follow SURVEY.md section 8(d): fBm fields by spectral synthesis (periodic in u),
  albedo rgb8 (seed 1) . topography r8 = clip((fBm-.5)*3,0,1) (seed 11) . landocean r8 = smoothed
  (topography==0) . clouds r8 = clip((fBm-q)/(max-q),0,1)^.5 for a coverage quantile (seed 22,
  optional spiral "hurricane") . bathymetry r8 (seed 33) . nightlights r8 sparse (seed 44) .
  stars rgb8 sparse points (seed 55).
"""
import numpy as np

SLOTS = ("albedo", "topography", "ocean", "clouds", "bathymetry", "emissive", "stars")


def _fbm(w, h, seed, beta=2.2):
    """Power-law noise in [0,1], periodic in both axes, float32 [h][w]."""
    rng = np.random.default_rng(seed)
    fy = np.fft.fftfreq(h)[:, None] * h
    fx = np.fft.rfftfreq(w)[None, :] * w
    # equirect: u spans 360 deg over w texels, v spans 180 deg over h -> isotropic in angle
    k = np.sqrt((fx * 0.5) ** 2 + fy ** 2).astype(np.float32)
    k[0, 0] = 1.0
    amp = k ** (-beta * 0.5)
    amp[0, 0] = 0.0
    ph = rng.random(amp.shape, dtype=np.float32) * np.float32(2 * np.pi)
    spec = (amp * np.cos(ph) + 1j * amp * np.sin(ph)).astype(np.complex64)
    f = np.fft.irfft2(spec, s=(h, w)).astype(np.float32)
    f -= f.min()
    f /= max(float(f.max()), 1e-20)
    return f


def _u8(x):
    return np.clip(np.rint(x * 255.0), 0, 255).astype(np.uint8)


def _smooth(x, n=2):
    for _ in range(n):
        x = (x + np.roll(x, 1, 1) + np.roll(x, -1, 1)) / 3.0
        x = (x + np.vstack([x[:1], x[:-1]]) + np.vstack([x[1:], x[-1:]])) / 3.0
    return x


def make_textures(w, h, cloud_cover=0.5, hurricane=False, seed=0):
    """Return {slot: uint8 array [h][w] or [h][w][3]}; deterministic in (w, h, args, seed)."""
    topo_f = _fbm(w, h, seed + 11)
    sea_level = float(np.quantile(topo_f, 0.62))
    topo = np.clip((topo_f - sea_level) / max(1.0 - sea_level, 1e-6) * 1.6, 0.0, 1.0)
    land = (topo > 0).astype(np.float32)
    ocean = _smooth(1.0 - land, 1)

    c = _fbm(w, h, seed + 22, beta=2.6)
    q = float(np.quantile(c, 1.0 - cloud_cover))
    clouds = np.clip((c - q) / max(float(c.max()) - q, 1e-6), 0.0, 1.0) ** 0.5
    if hurricane:
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        cx, cy = 0.62 * w, 0.64 * h
        dx, dy = (xx - cx) / w * 2.0, (yy - cy) / h
        r = np.sqrt(dx * dx + dy * dy) / 0.09
        th = np.arctan2(dy, dx)
        arms = 0.5 + 0.5 * np.cos(2.0 * th - 5.0 * np.log(r + 0.05))
        blob = np.exp(-r * r * 0.6) * (0.55 + 0.45 * arms) * (1.0 - np.exp(-(r / 0.08) ** 2))
        clouds = np.clip(np.maximum(clouds, blob * 1.2), 0.0, 1.0)

    veg = _fbm(w, h, seed + 1, beta=2.0)
    lat = np.abs(np.linspace(-1.0, 1.0, h, dtype=np.float32))[:, None]
    desert = np.clip(1.2 - np.abs(lat - 0.3) * 4.0, 0.0, 1.0) * (veg < 0.5)
    ice = np.clip((lat - 0.8) * 8.0, 0.0, 1.0)
    land_rgb = np.stack([0.16 + 0.30 * desert + 0.10 * veg, 0.22 + 0.16 * desert + 0.18 * veg, 0.08 + 0.10 * desert + 0.04 * veg], -1)
    sea_rgb = np.stack([0.02 + 0.03 * veg, 0.06 + 0.06 * veg, 0.16 + 0.12 * veg], -1)
    rgb = land_rgb * land[..., None] + sea_rgb * (1.0 - land[..., None])
    rgb = rgb * (1.0 - ice[..., None]) + 0.85 * ice[..., None]

    bathy = _fbm(w, h, seed + 33)
    nl = _fbm(w, h, seed + 44, beta=1.2)
    lights = np.clip((nl - float(np.quantile(nl, 0.97))) * 12.0, 0.0, 1.0) * land

    rng = np.random.default_rng(seed + 55)
    stars = np.zeros((h, w, 3), np.float32)
    n_stars = max(16, (w * h) // 2048)
    sy, sx = rng.integers(0, h, n_stars), rng.integers(0, w, n_stars)
    mag = rng.random(n_stars, dtype=np.float32) ** 6
    tint = 0.7 + 0.3 * rng.random((n_stars, 3), dtype=np.float32)
    stars[sy, sx] = mag[:, None] * tint

    return {
        "albedo": _u8(rgb), "topography": _u8(topo), "ocean": _u8(ocean), "clouds": _u8(clouds),
        "bathymetry": _u8(bathy), "emissive": _u8(lights), "stars": _u8(stars),
    }
