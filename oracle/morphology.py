"""Oracle (test infrastructure): the scipy.ndimage calls the reference wraps
in ``filters/extension_filters.py``, restated in NumPy.

Third-party dependency: scipy.ndimage (un-vendored, unpinned by the
reference: ``cguerrero/requirements.txt`` is empty).  Operative pin: the
scipy in this image.  These restatements follow the published algorithms
(binary erosion/dilation by a structuring element with ``border_value=0``;
grey dilation with a flat square = maximum filter with ``mode='reflect'``)
and are cross-checked against scipy itself in tests/test_oracle_golden.py.
"""
import numpy as np


def _as_bool(a):
    # scipy: input array -> "non-zero elements are True" (NaN is non-zero)
    return np.asarray(a) != 0


def _structure(structure):
    if structure is None:                          # generate_binary_structure(2, 1)
        return np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=bool)
    return np.asarray(structure) != 0


def _shifted(p, dy, dx, fill):
    """p shifted so that out[y, x] = p[y + dy, x + dx], outside = fill."""
    ny, nx = p.shape
    out = np.full_like(p, fill)
    ys = slice(max(0, -dy), min(ny, ny - dy))
    xs = slice(max(0, -dx), min(nx, nx - dx))
    yd = slice(max(0, dy), min(ny, ny + dy))
    xd = slice(max(0, dx), min(nx, nx + dx))
    out[ys, xs] = p[yd, xd]
    return out


def binary_erosion(image, structure=None, iterations=1):
    """BinaryErosion.apply (extension_filters.py:218-235):
    scipy.ndimage.binary_erosion(input, iterations=n) -- cross structuring
    element, border_value=0 (outside counts as False).  Returns bool."""
    p = _as_bool(image)
    s = _structure(structure)
    cy, cx = s.shape[0] // 2, s.shape[1] // 2
    for _ in range(iterations):
        out = np.ones_like(p)
        for y in range(s.shape[0]):
            for x in range(s.shape[1]):
                if s[y, x]:
                    out &= _shifted(p, y - cy, x - cx, False)
        p = out
    return p


def binary_dilation(image, structure=None, iterations=1):
    """scipy.ndimage.binary_dilation with border_value=0 (used inside closing)."""
    p = _as_bool(image)
    s = _structure(structure)
    cy, cx = s.shape[0] // 2, s.shape[1] // 2
    for _ in range(iterations):
        out = np.zeros_like(p)
        for y in range(s.shape[0]):
            for x in range(s.shape[1]):
                if s[y, x]:
                    # dilation reflects the structure: out[c] |= p[c - offset]
                    out |= _shifted(p, -(y - cy), -(x - cx), False)
        p = out
    return p


def binary_closing(image, structure=None):
    """BinaryClosing.apply (extension_filters.py:276-293):
    scipy.ndimage.binary_closing = dilation then erosion, both with
    border_value=0 -- so the one-cell frame is always False afterwards."""
    return binary_erosion(binary_dilation(image, structure), structure)


def grey_dilation_square(image, size=7):
    """GreyDilation.apply (extension_filters.py:328-345) with size=(s, s), s
    odd: flat structuring element == maximum filter, mode='reflect'
    (d c b a | a b c d | d c b a).  dtype preserved.  Inputs are NaN-free in
    the pipeline (products of 0/1 masks and majority values)."""
    a = np.asarray(image)
    h = size // 2
    p = np.pad(a, h, mode='symmetric')
    out = None
    for dy in range(size):
        for dx in range(size):
            blk = p[dy:dy + a.shape[0], dx:dx + a.shape[1]]
            out = blk.copy() if out is None else np.maximum(out, blk)
    return out
