"""GPU: hd_confusion_counts through hydrodem_b200.stats.Stats against the reference class's own outputs and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from hydrodem_b200 import geotiff
    from hydrodem_b200.stats import Stats
    from oracle import stats as ostats


@pytest.mark.parametrize("name", ["u8", "i16", "f32", "f32odd", "u8odd"])
def test_stats_match_the_reference_class(name):
    """Counts, totals and all eight scores for three simulated rasters (float32, float32, float64): identical to what the
    reference's Stats class returned (same Python arithmetic on the same integers -> the same floats)."""
    g = load_golden("run_stats")
    st = Stats(g[f"ndwi_{name}"], [g["sim_1"], g["sim_2"], g["sim_3"]], "x")
    got = st.get_stats()
    assert [r["day"] for r in got] == [1, 2, 3]
    keys = [str(k) for k in g["score_keys"]]
    assert [k for k in got[0] if k != "day"] == keys
    np.testing.assert_array_equal([[r[k] for k in keys] for r in got], g[f"scores_{name}"])
    assert [st.total_positives, st.total_negatives] == g[f"totals_{name}"].tolist()
    for d in range(3):
        st._set_values(g[f"sim_{d + 1}"])
        assert [st.values_file[k] for k in ("TP", "FN", "P", "FP", "TN", "N")] == g[f"counts_{name}"][d].tolist()
        assert all(isinstance(v, int) for v in st.values_file.values())


@pytest.mark.parametrize("shape,mdtype", [((1201, 1333), np.uint8), ((257, 4099), np.float32), ((3000, 17), np.int16)])
def test_counts_on_larger_rasters_and_from_files(tmp_path, shape, mdtype):
    rng = np.random.default_rng(shape[0])
    ndwi = (rng.random(shape) < 0.4).astype(mdtype)
    sim = np.where(rng.random(shape) < 0.35, rng.random(shape), 0.0).astype(np.float32)
    want = ostats.values(ndwi, sim)
    st = Stats(ndwi)
    st._set_values(sim)
    assert st.values_file == {k: int(v) for k, v in want.items()}
    assert (st.total_positives, st.total_negatives) == tuple(int(x) for x in ostats.totals(ndwi))
    if mdtype != np.uint8:                                 # a partition for signed / float 0-1 masks; uint8 differences wrap:
        assert sum(st.values_file[k] for k in ("TP", "FN", "FP", "TN")) == ndwi.size      # a mismatch is FN and FP there
    else:
        assert st.values_file["FN"] == st.values_file["FP"] == int(np.count_nonzero((sim > np.float32(0.0001)) != ndwi))
    geotiff.write_geotiff(tmp_path / "ndwi.tif", ndwi)
    geotiff.write_geotiff(tmp_path / "day1.tif", sim)
    from_files = Stats(str(tmp_path / "ndwi.tif"), [str(tmp_path / "day1.tif")], "f").get_stats()
    # NumPy >= 2 returns numpy.int64 from count_nonzero, so the reference's four-factor MCC product wraps once a raster has
    # more than ~55 000 cells per class (math.sqrt then even raises on a negative product); the counts here are Python
    # ints (what count_nonzero returned when the reference was written) and the product is exact
    sc = ostats.scores({k: int(v) for k, v in want.items()}, ndwi.size)
    assert -1.0 <= sc["MCC"] <= 1.0
    assert from_files == [{"day": 1, **{k + "_f": sc[k] for k in ("accuracy", "sensitivity", "BACC", "f1_score", "MCC",
                                                                 "precision", "specificity", "fall_out")}}]
