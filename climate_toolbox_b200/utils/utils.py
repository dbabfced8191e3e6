"""
Handy functions for standardizing the format of climate data
(mirror of ``/root/reference/climate_toolbox/utils/utils.py:10-80``).

The two functions on the aggregation path -- :func:`convert_lons_split`
(reference ``:33-40``) and :func:`remove_leap_days` (``:77-80``) -- copy the whole
dataset in the reference.  Here they only relabel / record positions (lazy takes
of ``_xr.Variable``); the plan builder folds the lon permutation into CSR column
indices and the kernel reads the time positions as an index list.
"""
from __future__ import annotations

import numpy as np

from .._xr import Dataset, DataArray, Deferred, Variable, from_any, to_like

__all__ = ["convert_kelvin_to_celsius", "convert_lons_mono", "convert_lons_split",
           "rename_coords_to_lon_and_lat", "rename_coords_to_longitude_and_latitude",
           "remove_leap_days"]


def convert_kelvin_to_celsius(df, temp_name):
    """Convert Kelvin to Celsius (reference ``:10-20``).  The subtraction is
    deferred and fuses into the aggregation kernel as ``(x - 273.15)^1``."""
    like = df
    ds = from_any(df)
    var = ds._vars[temp_name]
    attrs = dict(var.attrs)
    attrs.update(units="C", valid_min=-108.78788, valid_max=62.02828)
    src = var if var.deferred is None else Variable(var.dims, var.values, var.attrs)
    out = ds.copy()
    out._vars[temp_name] = Variable(var.dims, None, attrs, None,
                                    Deferred("poly", (273.15, 1.0), (src,)))
    return to_like(out, like)


def _relabel_and_sort(ds, lon_name, fn):
    like = ds
    ds = from_any(ds)
    if lon_name not in ds._coords:
        raise KeyError(lon_name)
    new = fn(np.asarray(ds._coords[lon_name].values))
    out = ds.copy()
    c = out._coords[lon_name]
    out._coords[lon_name] = Variable(c.dims, new, c.attrs)
    # ds.sel(lon=np.sort(lon)): exact-label orthogonal selection -> lazy take
    return to_like(out.sel(**{lon_name: np.sort(new)}), like)


def convert_lons_mono(ds, lon_name="longitude"):
    """Convert longitude from -180-180 to 0-360 (reference ``:23-30``)"""
    return _relabel_and_sort(ds, lon_name, lambda lon: lon % 360)


def convert_lons_split(ds, lon_name="longitude"):
    """Convert longitude from 0-360 to -180-180 (reference ``:33-40``)"""
    return _relabel_and_sort(ds, lon_name, lambda lon: (lon + 180) % 360 - 180)


def _rename(ds, pairs):
    like = ds
    ds = from_any(ds)
    for old_names, new in pairs:
        for old in old_names:
            if old in ds.coords:
                ds = ds.rename({old: new})
                break
    if "z" in ds.coords:
        ds = ds.drop("z").squeeze()
    return to_like(ds, like)


def rename_coords_to_lon_and_lat(ds):
    """Rename Dataset spatial coord names to: lat, lon (reference ``:43-57``)"""
    return _rename(ds, ((("latitude",), "lat"), (("longitude", "long"), "lon")))


def rename_coords_to_longitude_and_latitude(ds):
    """Rename Dataset spatial coord names to: latitude, longitude (reference ``:60-74``)"""
    return _rename(ds, ((("lat",), "latitude"), (("lon", "long"), "longitude")))


def remove_leap_days(ds):
    """Drop Feb 29 steps (reference ``:77-80``) as a lazy time take."""
    like = ds
    ds = from_any(ds)
    keep = ~((ds["time.month"].values == 2) & (ds["time.day"].values == 29))
    return to_like(ds.loc[{"time": keep}], like)
