"""CPU: the oracle of the flood-extent scores (oracle/stats.py) against outputs of the reference's own Stats class
(tests/golden/run_stats.npz, generated through stub gdal / config_loader modules)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import stats as ostats

MASKS = ("u8", "i16", "f32", "f32odd", "u8odd")


@pytest.mark.parametrize("name", MASKS)
def test_oracle_counts_and_scores_match_the_reference_class(name):
    g = load_golden("run_stats")
    ndwi = g[f"ndwi_{name}"]
    with np.errstate(all="ignore"):
        assert list(ostats.totals(ndwi)) == g[f"totals_{name}"].tolist()
        for d in range(3):
            v = ostats.values(ndwi, g[f"sim_{d + 1}"])
            assert [v[k] for k in ("TP", "FN", "P", "FP", "TN", "N")] == g[f"counts_{name}"][d].tolist()
            sc = ostats.scores(v, int(g[f"totals_{name}"].sum()))
            np.testing.assert_array_equal([sc[k[:-2]] for k in g["score_keys"]], g[f"scores_{name}"][d])


def test_reference_quirks_are_kept():
    """uint8 differences wrap (a mask value of 2 under a wet cell is a false negative AND a false positive), NaN times
    zero counts as non-zero, f1_score counts FP twice."""
    g = load_golden("run_stats")
    assert not np.array_equal(g["counts_u8odd"], g["counts_u8"])
    assert not np.array_equal(g["counts_f32odd"], g["counts_f32"])
    v = {"TP": 10, "FN": 5, "P": 15, "FP": 2, "TN": 20, "N": 22}
    assert ostats.scores(v, 37)["f1_score"] == 20. / 24.
