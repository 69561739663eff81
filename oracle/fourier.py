"""Oracle (test infrastructure): the 2-D Fourier stripe-removal stage.

Follows ``filters/custom_filters.py:834-1101`` (FourierInitial,
FourierProcessQuarters, DetectApplyFourier) and the wrappers in
``filters/extension_filters.py:348-480``.

Third-party dependency: ``scipy.fftpack.fft2/ifft2/fftshift/ifftshift``
(un-vendored; pocketfft/DUCC backend; float32 in -> complex64 out).  It is the
operative oracle for the transforms; ``dft2_direct`` below is the textbook
definition used to cross-check it on small sizes.
"""
import numpy as np
from scipy import fftpack

from . import stencils

MARGIN = 10                                        # custom_filters.py:911


def dft2_direct(a, inverse=False):
    """Definition of the 2-D DFT in complex128 (O(n^3); small inputs only)."""
    a = np.asarray(a, dtype=np.complex128)
    ny, nx = a.shape
    sign = 2j if inverse else -2j
    wy = np.exp(sign * np.pi * np.outer(np.arange(ny), np.arange(ny)) / ny)
    wx = np.exp(sign * np.pi * np.outer(np.arange(nx), np.arange(nx)) / nx)
    out = wy @ a @ wx
    return out / (ny * nx) if inverse else out


def fourier_initial(image):
    """FourierInitial.apply (custom_filters.py:859-877): fft2 -> fftshift ->
    abs.  Returns (|F| shifted, F shifted).  float32 input gives complex64 /
    float32 (scipy.fftpack keeps single precision)."""
    f = fftpack.fft2(image)
    fs = fftpack.fftshift(f)
    return np.abs(fs), fs


def quarter_geometry(ny, nx):
    """divmod bookkeeping of FourierProcessQuarters.__init__ (:907-911)."""
    my, y_odd = divmod(ny, 2)
    mx, x_odd = divmod(nx, 2)
    return my, y_odd, mx, x_odd


def first_quarters(fabs):
    """_get_firsts_quarters (:936-948)."""
    ny, nx = fabs.shape
    my, _, mx, x_odd = quarter_geometry(ny, nx)
    q1 = fabs[:my - MARGIN, :mx - MARGIN]
    q2 = fabs[:my - MARGIN, mx + MARGIN + x_odd:nx]
    return q1, q2


def assemble_mask(m1, m2, ny, nx):
    """_fill_complete_quarters + _getting_reversed_masks + _fill_complete_mask
    (:968-1050): the two quarter masks are placed in (my, mx) zero blocks,
    point-mirrored into the lower half, and tiled; with odd sizes the middle
    row / column stays zero."""
    my, y_odd, mx, x_odd = quarter_geometry(ny, nx)
    c1 = np.zeros((my, mx))
    c2 = np.zeros((my, mx))
    c1[:my - MARGIN, :mx - MARGIN] = m1            # :988-989
    c2[:my - MARGIN, MARGIN:mx] = m2               # :990-991
    c3 = c2[::-1, ::-1]                            # :1023
    c4 = c1[::-1, ::-1]                            # :1024
    full = np.zeros((ny, nx))
    full[:my, :mx] = c1
    full[:my, mx + x_odd:nx] = c2
    full[my + y_odd:ny, :mx] = c3
    full[my + y_odd:ny, mx + x_odd:nx] = c4
    return full


def process_quarters(fabs, mask_fn=None):
    """FourierProcessQuarters.apply (:913-934)."""
    mask_fn = mask_fn or stencils.mask_fourier
    q1, q2 = first_quarters(fabs)
    return assemble_mask(mask_fn(q1), mask_fn(q2), *fabs.shape)


def apply_mask(mask, fshift):
    """Tail of DetectApplyFourier (:1097-1100): (1 - mask) * F_shift ->
    ifftshift -> ifft2 -> abs.  float64 mask promotes to complex128."""
    g = fshift * (1 - mask)
    return np.abs(fftpack.ifft2(fftpack.ifftshift(g)))


def detect_apply_fourier(image, mask_fn=None):
    """DetectApplyFourier.apply (custom_filters.py:1077-1101).
    Returns (corrected float64, mask float64, |F| shifted)."""
    fabs, fshift = fourier_initial(image)
    mask = process_quarters(fabs, mask_fn)
    return apply_mask(mask, fshift), mask, fabs
