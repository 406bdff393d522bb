"""Panel assembly and PNG output (reference utils/image.py:10-64), PIL only."""
import os

import numpy as np
import PIL.Image


def imgrid(imarray, cols=5, pad=1):
    if imarray.dtype != np.uint8:
        raise ValueError("imgrid input imarray must be uint8")
    n, h, w, c = imarray.shape
    rows = int(np.ceil(n / float(cols)))
    grid = np.full((rows * (h + pad) + pad, cols * (w + pad) + pad, c), 255, dtype=np.uint8)
    for i in range(n):
        r, q = divmod(i, cols)
        y, x = pad + r * (h + pad), pad + q * (w + pad)
        grid[y:y + h, x:x + w] = imarray[i]
    return grid


def save_im(im, filename, fmt="png"):
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    PIL.Image.fromarray(np.asarray(im)).save(filename + "." + fmt)
