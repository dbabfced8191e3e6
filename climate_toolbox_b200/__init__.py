"""climate_toolbox_b200 -- B200-native grid->region aggregation hot path.

Drop-in for the aggregation path of ClimateImpactLab/climate_toolbox
(``climate_toolbox.aggregations.aggregations.weighted_aggregate_grid_to_regions``
plus the gridcell transforms, lon standardisation and leap-day removal that
feed it).  The sub-package layout mirrors the reference's so that
``from climate_toolbox_b200.aggregations.aggregations import ...`` reads like
the reference import; :func:`install_as_climate_toolbox` registers the same
modules under the reference's package name.

Arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI in
``include/ctb.h`` (``libctb.so``); there is no CPU fallback.
"""
import sys as _sys

__version__ = "0.1.0"

from ._xr import DataArray, Dataset  # noqa: E402,F401


def install_as_climate_toolbox():
    """Alias this package's modules as ``climate_toolbox.*`` (the reference's
    import paths) in ``sys.modules``.  Refuses to shadow a real install."""
    import importlib

    if "climate_toolbox" in _sys.modules and \
            getattr(_sys.modules["climate_toolbox"], "__b200__", False) is False:
        raise RuntimeError("a different 'climate_toolbox' is already imported")
    me = _sys.modules[__name__]
    me.__b200__ = True
    _sys.modules["climate_toolbox"] = me
    for sub in ("aggregations", "aggregations.aggregations", "transformations",
                "transformations.transformations", "utils", "utils.utils", "io", "io.io"):
        _sys.modules["climate_toolbox." + sub] = importlib.import_module(__name__ + "." + sub)
    return me
