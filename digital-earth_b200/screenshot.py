"""Screenshot output as earth_viewer.py:244-250: `screenshot/<main>-%Y-%m-%d-%H%M%S.jpg`, written
from the (W, H, 3) float image `fetch_image()` returns ([x][y], y up) exactly like
ti.tools.imwrite does: clip to [0,1], scale to uint8, swap axes and flip so row 0 is the top."""
import os
from datetime import datetime

import numpy as np


def to_uint8_image(img):
    """(W, H, 3) float in [0,1], y up  ->  (H, W, 3) uint8, top row first."""
    if hasattr(img, "detach"):
        img = img.detach().cpu().numpy()
    a = np.asarray(img, dtype=np.float32)
    a = (np.clip(a, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)
    return np.ascontiguousarray(a.swapaxes(0, 1)[::-1])


def save_screenshot(img, path=None, main_filename=None):
    from PIL import Image
    if path is None:
        import __main__
        timestamp = datetime.today().strftime("%Y-%m-%d-%H%M%S")
        main_filename = main_filename or os.path.split(getattr(__main__, "__file__", "main.py"))[1]
        os.makedirs(os.path.join(os.getcwd(), "screenshot"), exist_ok=True)
        path = os.path.join(os.getcwd(), "screenshot", f"{main_filename}-{timestamp}.jpg")
    Image.fromarray(to_uint8_image(img)).save(path)
    print(f"Screenshot has been saved to {path}")
    return path
