"""Oracle (test infrastructure): window geometry of the reference iterators.

Restates ``/root/reference/cguerrero/hydrodem/sliding_window.py`` without the
per-cell generator: all windows of a grid are produced at once as a
``(rows, cols, ws, ws)`` float32 array together with the ``(j, i)`` centres
in the reference's raster order.
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view


class OracleWindowError(ValueError):
    """Raised where the reference raises one of its window exceptions."""


def check_window_size(shape, ws):
    """Size guards of ``SlidingWindow.window_size`` (sliding_window.py:150-156).

    The too-large test comes first, then the even test, exactly as there.
    """
    if any(ws > n for n in shape):
        raise OracleWindowError("high")
    if ws % 2 != 1:
        raise OracleWindowError("even")


def nan_offsets(ws, circular=False, inner_size=None, no_center=False):
    """Window-local (row, col) indices that the variant sets to NaN.

    circular  -> 4 corner cells only (sliding_window.py:485-499)
    inner     -> central inner_size^2 block except the centre (:624-653)
    no_center -> the centre cell (:720-736)
    """
    out = []
    if circular:
        out += [(0, 0), (0, ws - 1), (ws - 1, 0), (ws - 1, ws - 1)]
    c = ws // 2
    if inner_size is not None:
        if inner_size > ws:
            raise OracleWindowError("inner")
        r = inner_size // 2
        out += [(y, x) for y in range(c - r, c + r + 1)
                for x in range(c - r, c + r + 1) if not (y == x == c)]
    if no_center:
        out.append((c, c))
    return out


def all_windows(grid, ws, *, circular=False, inner_size=None, no_center=False,
                ignore_border=False):
    """All windows of ``grid`` in raster order.

    Follows ``SlidingWindow.__iter__`` (sliding_window.py:158-196): the grid
    is cast to float32 (:132), only centres with a complete window are
    visited, each window is a copy with the variant's cells set to NaN
    (:280-300).  ``ignore_border`` first pads the grid with ws//2 NaN cells
    on every side (:400-418); the returned centres are then in *padded*
    coordinates, as in the reference.

    Returns (windows[rows, cols, ws, ws] float32, jj[rows], ii[cols]).
    """
    g = np.asarray(grid).astype('float32')
    check_window_size(g.shape, ws)
    h = ws // 2
    if ignore_border:
        g = np.pad(g, h, mode='constant', constant_values=np.nan)
    view = sliding_window_view(g, (ws, ws)).copy()
    for (y, x) in nan_offsets(ws, circular, inner_size, no_center):
        view[:, :, y, x] = np.nan
    jj = np.arange(h, g.shape[0] - h)
    ii = np.arange(h, g.shape[1] - h)
    return view, jj, ii


def window_at(grid, ws, j, i, **variant):
    """One window by centre, ``SlidingWindow.__getitem__`` (:198-262)."""
    g = np.asarray(grid).astype('float32')
    check_window_size(g.shape, ws)
    h = ws // 2
    if variant.pop('ignore_border', False):
        g = np.pad(g, h, mode='constant', constant_values=np.nan)
    ny, nx = g.shape
    # _check_border (:264-271) -- note its upper bound is one cell too
    # generous (ny - right_down + 1); numpy slicing then returns a short
    # window.  The oracle keeps the reference's test.
    if not (j >= h and i >= h and j <= ny - (h + 1) + 1 and i <= nx - (h + 1) + 1):
        raise OracleWindowError("border")
    win = g[j - h:j + h + 1, i - h:i + h + 1].copy()
    for (y, x) in nan_offsets(ws, **variant):
        win[y, x] = np.nan
    return win
