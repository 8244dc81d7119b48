import numpy as np
import cv2


def rescale(img, scale, anti_aliasing=False, multichannel=False, **_):
    """Bilinear rescale (skimage's order=1 default), enough for the data loaders' down-scaling in tests."""
    h, w = img.shape[:2]
    nh, nw = int(round(h * scale)), int(round(w * scale))
    return cv2.resize(np.asarray(img, dtype=np.float32), (nw, nh), interpolation=cv2.INTER_LINEAR)
