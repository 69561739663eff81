"""Oracle (test infrastructure): the reference's windowed filters, vectorised.

Each function restates one class of
``/root/reference/cguerrero/hydrodem/filters/custom_filters.py`` (file:line
cited per function) with the same input casts, output dtype, border
behaviour and aliasing, but without the per-cell Python loop.
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

from .windows import check_window_size


def _f32(grid):
    # SlidingWindow.grid setter: value.astype('float32') (sliding_window.py:132)
    return np.asarray(grid).astype('float32')


def _footprint(ws, circular):
    fp = np.ones((ws, ws), dtype=bool)
    if circular:
        fp[0, 0] = fp[0, -1] = fp[-1, 0] = fp[-1, -1] = False  # sliding_window.py:499
    return fp


def majority_min_count(ws):
    """Smallest count that passes ``count > (ws**2 - 1) * 0.7``
    (custom_filters.py:71), evaluated with the same Python float expression."""
    thr = (ws ** 2 - 1) * 0.7
    return int(np.floor(thr)) + 1


def majority(image, ws, rows_per_chunk=64):
    """MajorityFilter.apply (custom_filters.py:48-73).

    Mode of the ws*ws window with the four corners NaN'ed (CircularWindow,
    centre included).  Counter keys compare with ``==`` so every NaN is its
    own key of count 1 and -0.0/0.0 share one; the mode is written only if
    its count exceeds (ws^2-1)*0.7.  Output: float64 zeros, untouched border
    of ws//2.

    Independent algorithm (sort + run test instead of a hash count): a value
    reaches count >= m iff sorted[k+m-1] == sorted[k] for some k.
    """
    g = _f32(image)
    check_window_size(g.shape, ws)
    out = np.zeros(g.shape)                       # :66  float64
    h = ws // 2
    m = majority_min_count(ws)
    keep = _footprint(ws, True).ravel()
    n = int(keep.sum())
    if m > n:
        return out                                # can never fire (e.g. ws=3)
    view = sliding_window_view(g, (ws, ws))
    rows = view.shape[0]
    for r0 in range(0, rows, rows_per_chunk):
        v = view[r0:r0 + rows_per_chunk]
        flat = v.reshape(v.shape[0], v.shape[1], ws * ws)[:, :, keep]
        s = np.sort(flat, axis=-1)                # NaN sort last, never equal
        hit = s[:, :, m - 1:] == s[:, :, :n - m + 1]
        anyhit = hit.any(axis=-1)
        first = hit.argmax(axis=-1)
        val = np.take_along_axis(s, first[..., None], axis=-1)[..., 0]
        blk = out[h + r0:h + r0 + v.shape[0], h:g.shape[1] - h]
        blk[anyhit] = val[anyhit]
    return out


def expand(image, ws):
    """ExpandFilter.apply (custom_filters.py:102-125): 1.0 where any non-NaN
    cell of the corner-less window is > 0; float64 zeros elsewhere and on the
    ws//2 border."""
    g = _f32(image)
    check_window_size(g.shape, ws)
    out = np.zeros(g.shape)                       # :119
    h = ws // 2
    pos = g > 0                                   # NaN > 0 is False (:122-123)
    view = sliding_window_view(pos, (ws, ws))
    fp = _footprint(ws, True)
    hit = (view & fp).any(axis=(-1, -2))
    out[h:g.shape[0] - h, h:g.shape[1] - h][hit] = 1
    return out


def np_pairwise_sum_f32(values):
    """numpy's float32 add.reduce for n <= 8 contiguous values, spelled out
    (the order the CUDA kernel has to reproduce for CorrectNANValues):
    n < 8 -> sequential left to right; n == 8 -> eight accumulators combined
    ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)).  Checked against ``np.sum`` in
    tests/test_oracle_golden.py."""
    v = np.asarray(values, dtype=np.float32)
    n = v.size
    assert n <= 8
    if n < 8:
        acc = np.float32(-0.0) if n else np.float32(0.0)
        for x in v:
            acc = np.float32(acc + x)
        return acc
    r = v
    return np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3])) +
                      np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))


def correct_nan(dem):
    """CorrectNANValues.apply (custom_filters.py:286-317), window_size=3.

    Interior cells whose float32 value is < 0 (MaskNegatives :465-486 gated by
    ``int(v) == 1`` sliding_window.py:192) are replaced by the float32 mean of
    their 8 neighbours that are non-NaN and >= 0, read from a float32
    *snapshot* taken before any write (sliding_window.py:132).  Writes into
    ``dem`` and returns the same object (:316-317).  No valid neighbour gives
    NaN.
    """
    check_window_size(dem.shape, 3)
    snap = _f32(dem)
    ny, nx = snap.shape
    cand = np.argwhere(snap[1:ny - 1, 1:nx - 1] < 0) + 1
    for j, i in cand:
        win = snap[j - 1:j + 2, i - 1:i + 2].copy()
        win[1, 1] = np.nan                        # NoCenterWindow
        neigh = win[~np.isnan(win)]               # :314
        pos = neigh[neigh >= 0]                   # :315
        with np.errstate(invalid='ignore', divide='ignore'):
            dem[j, i] = pos.mean() if pos.size else np.float32(np.nan)  # :316
    return dem


def isolated_points(mask):
    """IsolatedPoints.apply (custom_filters.py:345-366), window_size=3.

    Interior cells with ``int(float32(v)) == 1`` are rewritten to 1.0 if any
    of the 8 neighbours (snapshot values, NaN skipped) is > 0, else 0.0.
    In place; returns the same object.
    """
    check_window_size(mask.shape, 3)
    snap = _f32(mask)
    ny, nx = snap.shape
    core = snap[1:ny - 1, 1:nx - 1]
    with np.errstate(invalid='ignore'):
        gate = np.trunc(core) == 1                # int(v) == 1  (NaN would raise in the reference)
    pos = snap > 0
    view = sliding_window_view(pos, (3, 3))
    fp = np.ones((3, 3), dtype=bool)
    fp[1, 1] = False
    neigh_any = (view & fp).any(axis=(-1, -2))
    tgt = mask[1:ny - 1, 1:nx - 1]
    tgt[gate] = np.where(neigh_any[gate], 1.0, 0.0)
    return mask


def quadratic_terms(ws):
    """r0..r3 and the coordinate grids of QuadraticFilter (custom_filters.py:240-246).
    Note the half-pixel asymmetric offsets -ws/2+1 .. ws/2."""
    values = np.linspace(-ws / 2 + 1, ws / 2, ws)
    xx, yy = np.meshgrid(values, values)
    r0 = ws ** 2
    r1 = (xx * xx).sum()
    r2 = (xx * xx * xx * xx).sum()
    r3 = (xx * xx * yy * yy).sum()
    return xx, yy, r0, r1, r2, r3


def quadratic_kernel(ws):
    """The fixed ws*ws correlation kernel equivalent to custom_filters.py:252-256:
    K = ((x^2 + y^2) r1 - (r2 + r3)) / (2 r1^2 - r0 (r2 + r3)); sums to 1."""
    xx, yy, r0, r1, r2, r3 = quadratic_terms(ws)
    den = 2 * r1 ** 2 - r0 * (r2 + r3)
    return ((xx * xx + yy * yy) * r1 - (r2 + r3)) / den


def quadratic(dem, ws, rows_per_chunk=32):
    """QuadraticFilter.apply (custom_filters.py:226-257).

    smoothed = dem.copy() (keeps dtype and the ws//2 border, :249); interior
    cells get ((s2+s3) r1 - s1 (r2+r3)) / (2 r1^2 - r0 (r2+r3)) with the
    window cast to float32, s1 a float32 sum and s2, s3 float64 sums
    (:252-254).  Tolerance class: the oracle reproduces the float32 s1 with
    numpy's own reduction but in a different blocking than the per-window
    call, so agreement with the reference is ~1e-7 relative, not bitwise.
    """
    g = _f32(dem)
    check_window_size(g.shape, ws)
    xx, yy, r0, r1, r2, r3 = quadratic_terms(ws)
    den = 2 * r1 ** 2 - r0 * (r2 + r3)
    out = dem.copy()
    h = ws // 2
    view = sliding_window_view(g, (ws, ws))
    wxx = (xx * xx)
    wyy = (yy * yy)
    for r0_ in range(0, view.shape[0], rows_per_chunk):
        v = view[r0_:r0_ + rows_per_chunk]
        flat = np.ascontiguousarray(v).reshape(v.shape[0], v.shape[1], ws * ws)
        s1 = flat.sum(axis=-1)                                   # float32 pairwise
        s2 = (flat * wxx.ravel()).sum(axis=-1)                   # float64
        s3 = (flat * wyy.ravel()).sum(axis=-1)
        res = ((s2 + s3) * r1 - s1 * (r2 + r3)) / den
        out[h + r0_:h + r0_ + v.shape[0], h:g.shape[1] - h] = res
    return out


def groves_correction(dem, groves_class, ws=15, trace=None):
    """GrovesCorrection.apply (custom_filters.py:704-732), one iteration:
    smooth = Quadratic(dem); hi = dem - smooth; tall = (hi > 1.5)*1;
    keep = 1 - groves_class*tall; result = keep*hi + smooth.
    ``trace``: a list that receives ``hi`` of this iteration (the parity tests use it to find the cells that sit ON
    the 1.5 m threshold, where a tolerance-class difference upstream legitimately flips the decision)."""
    smooth = quadratic(dem, ws)
    hi = dem - smooth
    if trace is not None:
        trace.append(hi)
    tall = (hi > 1.5) * 1
    keep = 1 - groves_class * tall
    return hi * keep + smooth


def groves_corrections_iter(dem, groves_class, iterations=3, ws=15, trace=None):
    """GrovesCorrectionsIter (custom_filters.py:755-767)."""
    for _ in range(iterations):
        dem = groves_correction(dem, groves_class, ws, trace)
    return dem


def _box_sum_f64(a, ws):
    """Sum over the ws*ws window clipped to the image, via a float64 integral
    image (exact enough: 53-bit accumulation of float32 data)."""
    ny, nx = a.shape
    h = ws // 2
    ii = np.zeros((ny + 1, nx + 1), dtype=np.float64)
    ii[1:, 1:] = a.astype(np.float64).cumsum(0).cumsum(1)
    y0 = np.clip(np.arange(ny) - h, 0, ny)
    y1 = np.clip(np.arange(ny) + h + 1, 0, ny)
    x0 = np.clip(np.arange(nx) - h, 0, nx)
    x1 = np.clip(np.arange(nx) + h + 1, 0, nx)
    s = ii[y1][:, x1] - ii[y0][:, x1] - ii[y1][:, x0] + ii[y0][:, x0]
    cnt = (y1 - y0)[:, None] * (x1 - x0)[None, :]
    return s, cnt


def hollow_mean(image, ws=55, inner=5):
    """Mean of the IgnoreBorderInnerSliding window used by BlanksFourier
    (custom_filters.py:416-421): ws*ws window clipped to the image (NaN
    padding, sliding_window.py:400-418) minus the central inner*inner block,
    centre included (InnerWindow + NoCenterWindow).  NaNs inside the image are
    skipped as np.nanmean does.  float64 integral-image restatement of a
    float32 pairwise nanmean => tolerance class."""
    g = _f32(image)
    nan = np.isnan(g)
    z = np.where(nan, np.float32(0), g)
    valid = (~nan).astype(np.float64)
    s_big, _ = _box_sum_f64(z, ws)
    c_big, _ = _box_sum_f64(valid, ws)
    s_in, _ = _box_sum_f64(z, inner)
    c_in, _ = _box_sum_f64(valid, inner)
    with np.errstate(invalid='ignore', divide='ignore'):
        return ((s_big - s_in) / (c_big - c_in)).astype(np.float32)


def blanks_fourier(image, ws=55, inner=5):
    """BlanksFourier.apply (custom_filters.py:395-427): mask = centre >
    4*hollow_mean for every cell; returns (mask float64, image*(1-mask))."""
    check_window_size(np.asarray(image).shape, ws)
    mean = hollow_mean(image, ws, inner)
    with np.errstate(invalid='ignore'):
        mask = (image > (4 * mean)).astype(np.float64)
    return mask, image * (1 - mask)


def detect_blanks_fourier(quarter):
    """DetectBlanksFourier.apply (custom_filters.py:441-462): two passes of
    BlanksFourier(55), masks added."""
    final = np.zeros(quarter.shape)
    for _ in (0, 1):
        m, quarter = blanks_fourier(quarter, 55, 5)
        final += m
    return final


def mask_fourier(quarter):
    """MaskFourier (custom_filters.py:559-561): DetectBlanksFourier ->
    IsolatedPoints(3) -> ExpandFilter(13)."""
    return expand(isolated_points(detect_blanks_fourier(quarter)), 13)


def mean3_round(image, weights=None):
    """PostProcessingFinal (custom_filters.py:1124-1125) = Convolve() then
    Around(): scipy.ndimage.convolve(ones(3,3), mode='reflect') / 9
    (extension_filters.py:183-184) then np.around (:130).

    NumPy restatement of the published ndimage algorithm: double accumulator
    starting at 0, footprint visited in row-major order, 'reflect' =
    (d c b a | a b c d | d c b a)."""
    w = np.ones((3, 3)) if weights is None else np.asarray(weights, dtype=np.float64)
    return np.around(convolve_reflect(image, w) / w.size)


def convolve_reflect(image, weights):
    """scipy.ndimage.convolve(image, weights) with the default mode='reflect',
    origin 0, restated: out = sum_k w_flipped[k] * in[shifted], double
    accumulation in row-major footprint order (NI_Correlate)."""
    a = np.asarray(image)
    w = np.asarray(weights, dtype=np.float64)
    ky, kx = w.shape
    hy, hx = ky // 2, kx // 2
    # convolve == correlate with the reversed kernel (and, for even sizes, a
    # shifted origin; only odd sizes are supported here as in the pipeline)
    assert ky % 2 == 1 and kx % 2 == 1
    wr = w[::-1, ::-1]
    p = np.pad(a.astype(np.float64), ((hy, hy), (hx, hx)), mode='symmetric')
    out = np.zeros(a.shape, dtype=np.float64)
    for dy in range(ky):
        for dx in range(kx):
            if wr[dy, dx] != 0:                    # ndimage skips zero weights
                out = out + p[dy:dy + a.shape[0], dx:dx + a.shape[1]] * wr[dy, dx]
    if a.dtype == np.float32:
        return out.astype(np.float32)
    return out


def median(dem, ws, circular=False):
    """NEW stage N1 (no reference code -- parity unpinned).  Definition
    (SURVEY.md section 8a): smoothed = dem.copy(); every interior cell gets
    np.nanmedian of its float32 window (square, or corner-less if
    ``circular``); the ws//2 border is unchanged (QuadraticFilter convention,
    custom_filters.py:249)."""
    g = _f32(dem)
    check_window_size(g.shape, ws)
    out = dem.copy()
    h = ws // 2
    keep = _footprint(ws, circular).ravel()
    view = sliding_window_view(g, (ws, ws))
    flat = view.reshape(view.shape[0], view.shape[1], ws * ws)[:, :, keep]
    s = np.sort(flat, axis=-1)                    # NaN last
    k = (~np.isnan(s)).sum(axis=-1)
    lo = np.clip((k - 1) // 2, 0, None)
    hi = np.clip(k // 2, 0, None)
    a = np.take_along_axis(s, lo[..., None], axis=-1)[..., 0]
    b = np.take_along_axis(s, hi[..., None], axis=-1)[..., 0]
    with np.errstate(invalid='ignore', over='ignore'):
        # np.nanmedian -> np.median -> mean of the two middle values:
        # float32 add then true-divide by 2 (numpy _median / mean).
        med = np.where(lo == hi, a, (a + b) / np.float32(2))
    med = np.where(k == 0, np.float32(np.nan), med).astype(np.float32)
    out[h:g.shape[0] - h, h:g.shape[1] - h] = med
    return out
