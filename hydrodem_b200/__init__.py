"""hydrodem_b200 -- B200-native HydroDEM raster-conditioning hot path.

``hydrodem_b200.filters`` mirrors the reference's ``filters`` package (same class names and
behaviour, CUDA underneath); ``hydrodem_b200.pipeline.ConditioningChain`` runs the whole chain
device-resident; ``include/hydrodem_b200.h`` is the C ABI.  See DESIGN.md / INTEGRATION.md.
"""
import importlib
import sys

__all__ = ["install_as_reference_filters"]


def install_as_reference_filters():
    """Make the reference's flat imports (``from filters.custom_filters import MajorityFilter``,
    ``from exceptions import WindowSizeEvenError``; image_srtm.py:7-8, sliding_window.py:7) resolve to this
    package, so image_srtm.py / image_hsheds.py / hydro_dem_process.py run on the GPU unchanged."""
    pkg = importlib.import_module("hydrodem_b200.filters")
    sys.modules["filters"] = pkg
    for sub in ("custom_filters", "extension_filters", "simple_filters", "new_filters"):
        sys.modules[f"filters.{sub}"] = importlib.import_module(f"hydrodem_b200.filters.{sub}")
    sys.modules["exceptions"] = importlib.import_module("hydrodem_b200.exceptions")
    return pkg
