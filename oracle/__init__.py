"""CPU oracle for the HydroDEM raster-conditioning hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker (or as the timed CPU baseline), never as a fallback for the CUDA path.

What it is: a vectorised NumPy (plus a small C library, ``oracle/c``)
restatement of the reference algorithms in
``/root/reference/cguerrero/hydrodem/filters/*.py`` and
``sliding_window.py``.  Every function cites the reference file:line it
follows.  The reference's per-cell Python loops are replaced by
``sliding_window_view`` reductions so that 1024^2..3601^2 inputs finish in
seconds; the arithmetic (dtype, order of casts, thresholds) is kept.

Pinning:
  * stages that exist in the reference are pinned two ways --
    (1) against the reference's own stored goldens G1-G4
        (``tests/golden/ref_*.npz``, converted from
        ``cguerrero/tests/resources/tests_expected.zip``), and
    (2) against outputs of the reference classes themselves, imported from
        ``/root/reference`` in the build container on seeded synthetic inputs
        (``tests/golden/gen_golden.py`` -> ``tests/golden/run_*.npz``).
  * third-party arithmetic the reference delegates to -- ``scipy.ndimage``
    (binary_erosion, binary_closing, grey_dilation, convolve) and
    ``scipy.fftpack`` (fft2/ifft2/fftshift/ifftshift) -- is un-vendored and
    unpinned in the reference (empty requirements.txt).  The operative pin is
    this image's numpy 2.3 / scipy 1.x; the oracle restates the published
    algorithms in NumPy (``morphology.py``) and the tests cross-check the
    restatement against scipy on the same inputs.
  * median, sink-fill and D8 DO NOT EXIST in the reference: **parity
    unpinned**.  Their oracle is this repo's own definition (SURVEY.md
    section 8a rows N1-N3), cross-checked by two independent algorithms
    (iterative Planchon-Darboux vs priority-flood; sorting vs
    ``scipy.ndimage.median_filter``).
"""
from . import windows, stencils, morphology, fourier, hydrology, chain  # noqa: F401
