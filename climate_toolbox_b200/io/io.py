"""Loaders that put climate data into the form the aggregation path requires
(mirror of ``/root/reference/climate_toolbox/io/io.py:6-58``)."""
from __future__ import annotations

from .._xr import Dataset, from_any, to_like
from ..utils.utils import rename_coords_to_lon_and_lat, convert_lons_split

__all__ = ["standardize_climate_data", "load_bcsd", "load_gmfd", "load_best"]


def standardize_climate_data(ds):
    """
    Read climate data and standardize units to lon and lat, lon to -180 to 180
    (reference ``io.py:6-24``).  No data is moved: coordinates are relabelled and
    the lon sort is recorded as a lazy permutation.
    """
    ds = rename_coords_to_lon_and_lat(ds)
    ds = convert_lons_split(ds, lon_name="lon")
    return ds


def load_bcsd(fp, varname, lon_name="lon", broadcast_dims=("time",)):
    """
    Read and prepare climate data (reference ``io.py:27-58``).

    ``fp`` may be a Dataset (this package's, holding numpy arrays or CUDA tensors, or
    ``xarray``'s) -- passed through like the reference's ``hasattr(fp, "sel_points")``
    branch -- or a netCDF path, which needs ``xarray`` (not part of this image; netCDF
    decode is outside the hot path).  ``varname``, ``lon_name`` and ``broadcast_dims``
    are accepted and ignored, as in the reference.
    """
    if isinstance(fp, Dataset) or hasattr(fp, "data_vars"):
        ds = fp
    else:
        try:
            import xarray as xr
        except ImportError as e:  # pragma: no cover
            raise ImportError("reading netCDF paths needs xarray; pass an in-memory Dataset") from e
        with xr.open_dataset(fp) as ds:  # pragma: no cover
            ds.load()
    return standardize_climate_data(ds)


def load_gmfd(fp, varname, lon_name="lon", broadcast_dims=("time",)):
    pass  # stub in the reference too (io.py:61-62)


def load_best(fp, varname, lon_name="lon", broadcast_dims=("time",)):
    pass  # stub in the reference too (io.py:65-66)
