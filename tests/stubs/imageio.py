"""Stand-in for the few `imageio` calls of the reference (test infrastructure): imread / imwrite / imsave through OpenCV, mimwrite as
a frame dump (`<name>.frames/%04d.png`) since no video encoder is available offline."""
import os
import numpy as np
import cv2


def imread(path, as_gray=False, **_):
    img = cv2.imread(str(path), cv2.IMREAD_GRAYSCALE if as_gray else cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3:
        img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
    return img


def imwrite(path, image, **_):
    img = np.asarray(image)
    if img.dtype != np.uint8:
        img = (np.clip(img, 0, 1) * 255).astype(np.uint8)
    if img.ndim == 3 and img.shape[2] >= 3:
        img = img[..., [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]
    os.makedirs(os.path.dirname(os.path.abspath(str(path))), exist_ok=True)
    if not cv2.imwrite(str(path), img):
        raise IOError(path)


imsave = imwrite


def mimwrite(path, images, **_):
    d = str(path) + '.frames'
    os.makedirs(d, exist_ok=True)
    for i, im in enumerate(images):
        imwrite(os.path.join(d, '%04d.png' % i), im)
