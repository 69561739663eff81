"""Oracle (test infrastructure): the composed filters and the full
conditioning chain, restated from ``filters/custom_filters.py`` and
``hydro_dem_process.py``.
"""
import numpy as np

from . import stencils, morphology, fourier, hydrology


def tidying_lagoons(majority_img):
    """TidyingLagoons.apply (custom_filters.py:587-610): BinaryErosion(2) ->
    ExpandFilter(7) -> * majority image -> GreyDilation((7,7))."""
    er = morphology.binary_erosion(majority_img, iterations=2)
    ex = stencils.expand(er, 7)
    prod = majority_img * ex                       # ProductFilter(factor=majority) :607
    return morphology.grey_dilation_square(prod, 7)


def lagoons_detection(hsheds):
    """LagoonsDetection.apply (custom_filters.py:633-661).  Mutates and
    aliases ``hsheds`` exactly like the reference (CorrectNANValues is in
    place).  Returns the results dict of ComposedFilterResults."""
    res = {}
    res["CorrectNANValues"] = stencils.correct_nan(hsheds)
    res["MajorityFilter"] = stencils.majority(res["CorrectNANValues"], 11)
    res["TidyingLagoons"] = tidying_lagoons(res["MajorityFilter"])
    res["MaskPositives"] = (res["TidyingLagoons"] > 0.0) * 1      # :509-510, int64
    return res


def srtm_branch(srtm_raw, groves_class_raw):
    """SRTM.process (image_srtm.py:65-79): Fourier correction (:125-126),
    BinaryClosing(ones(3,3)) of the groves class (:177-178),
    GrovesCorrectionsIter(iterations=3) (:199)."""
    corrected, mask, fabs = fourier.detect_apply_fourier(srtm_raw)
    groves = morphology.binary_closing(groves_class_raw, np.ones((3, 3)))
    trace = []
    out = stencils.groves_corrections_iter(corrected, groves, 3, trace=trace)
    return out, dict(fourier=corrected, mask=mask, fabs=fabs, groves=groves, groves_hi=trace)


def final_terms(srtm, lag, rivers):
    """HydroDEMProcess._prepare_final_terms (hydro_dem_process.py:60-91)."""
    mask_rl = lag["MaskPositives"] + rivers
    not_rl = 1 - mask_rl
    first = srtm * not_rl
    third = lag["CorrectNANValues"] * rivers
    return first, lag["TidyingLagoons"], third


def conditioning_chain(srtm_raw, groves_class_raw, hsheds, rivers=None, with_hydrology=True):
    """Full chain of HydroDEMProcess.start (hydro_dem_process.py:122-153)
    minus GDAL I/O and the sequential river routing (out of scope, SURVEY.md
    section 2 row 3b: ``rivers`` is a given 0/1 raster, default zeros), plus
    the two NEW stages sink-fill and D8 on the final DEM."""
    srtm, aux = srtm_branch(srtm_raw, groves_class_raw)
    lag = lagoons_detection(hsheds)
    if rivers is None:
        rivers = np.zeros(srtm.shape, dtype=np.int64)
    t1, t2, t3 = final_terms(srtm, lag, rivers)
    dem_complete = t1 + t2 + t3
    final = stencils.mean3_round(dem_complete)     # PostProcessingFinal :1124-1125
    out = dict(srtm=srtm, lagoons=lag, dem_complete=dem_complete, final=final, **aux)
    if with_hydrology:
        out["filled"] = hydrology.sinkfill(final)
        out["d8"] = hydrology.d8(out["filled"])
    return out
