"""Oracle (test infrastructure): the flood-extent scores of the reference's ``Stats`` class (stats.py:6-93), restated on
arrays -- the same NumPy expressions, minus GDAL and Config.  Pinned by tests/golden/run_stats.npz, which holds outputs of
the reference class itself (driven through stub ``gdal`` / ``config_loader`` modules by tests/golden/gen_golden.py)."""
import math

import numpy as np


def totals(ndwi):
    """Stats._totals (stats.py:21-25)."""
    total_positives = np.count_nonzero(ndwi)
    total_negatives = np.count_nonzero(1 - ndwi)
    return total_positives, total_negatives


def values(ndwi, file, threshold=0.0001):
    """Stats._set_values (stats.py:63-86): TP, FN, P, FP, TN, N."""
    out = {}
    ndwi_complement = 1 - ndwi
    mask = file > threshold
    out["TP"] = np.count_nonzero(mask * ndwi)
    out["FN"] = np.count_nonzero(((ndwi - mask) > 0) * 1)
    out["P"] = out["TP"] + out["FN"]
    out["FP"] = np.count_nonzero(((mask - ndwi) > 0) * 1)
    out["TN"] = np.count_nonzero((1 - mask) * ndwi_complement)
    out["N"] = out["FP"] + out["TN"]
    return out


def scores(v, total_values):
    """The eight score functions (stats.py:27-61), in the order of ``stats_functions`` (:16-18).  f1_score adds FP twice
    (:50-52) -- reproduced, not corrected."""
    sensitivity = v["TP"] / v["P"]
    specificity = v["TN"] / v["N"]
    num = v["TP"] * v["TN"] - v["FP"] * v["FN"]
    den = (v["TP"] + v["FP"]) * (v["TP"] + v["FN"]) * (v["TN"] + v["FP"]) * (v["TN"] + v["FN"])
    return {"accuracy": (v["TN"] + v["TP"]) / total_values, "sensitivity": sensitivity,
            "BACC": (sensitivity + specificity) / 2., "f1_score": (2. * v["TP"]) / (2. * v["TP"] + v["FP"] + v["FP"]),
            "MCC": num / math.sqrt(den), "precision": v["TP"] / (v["TP"] + v["FP"]), "specificity": specificity,
            "fall_out": v["FP"] / (v["FP"] + v["TN"])}
