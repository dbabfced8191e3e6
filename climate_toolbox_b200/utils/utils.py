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
           "remove_leap_days", "season_boundaries", "get_daily_growing_season_mask", "GrowingSeasonMask"]


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


# ---------------------------------------------------------------------------
# growing-season mask (reference ``:83-153``), fused into the aggregation
# ---------------------------------------------------------------------------
def _growing_days_arrays(growing_days):
    """(planting [lat][lon], harvest [lat][lon], latitude, longitude) of a crop-calendar Dataset: data
    variable ``variable`` over (z, latitude, longitude) with z = 1 (planting day) and 2 (harvest day)."""
    gd = from_any(growing_days)
    v = gd["variable"]
    arr = np.asarray(v.transpose("z", "latitude", "longitude").values, dtype=np.float64)
    z = list(np.asarray(gd["z"].values).tolist())
    return (arr[z.index(1)], arr[z.index(2)], np.asarray(gd["latitude"].values, dtype=np.float64),
            np.asarray(gd["longitude"].values, dtype=np.float64))


def season_boundaries(growing_days):
    """Returns the sorted start and end date of growing season (reference ``:83-116``).

    The crop calendar's longitudes are on 0..360: they are shifted by -180 and the grid is sorted by
    longitude (the reference does the shift in place on the caller's Dataset; this does not).
    Returns ``(min_day, max_day)`` DataArrays over (latitude, longitude): the two ``z`` planes sorted
    per gridcell (NaN last, like ``np.sort``)."""
    plant, harv, lat, lon = _growing_days_arrays(growing_days)
    lon = lon - 180
    order = np.argsort(lon, kind="stable")
    lon = lon[order]
    both = np.sort(np.stack([plant[:, order], harv[:, order]], axis=2), axis=2)
    coords = {"latitude": lat, "longitude": lon}
    return (DataArray(both[:, :, 0], dims=("latitude", "longitude"), coords=coords),
            DataArray(both[:, :, 1], dims=("latitude", "longitude"), coords=coords))


class GrowingSeasonMask:
    """What :func:`get_daily_growing_season_mask` returns: the lat x lon x time mask of the reference
    in factored form -- a (first day, last day, wrap) triple per gridcell and the day of year of every
    time step.  Pass it as ``season_mask=`` to ``weighted_aggregate_grid_to_regions``: the gate is
    applied inside the kernel (a gridcell-day outside its season adds nothing to the weighted sum,
    exactly like multiplying the data by the mask, whose 0 and NaN both vanish in the skip-NaN sum).
    ``.values`` materialises the dense (lat, lon, time) array of 1 / 0 / NaN the reference returns."""

    dims = ("lat", "lon", "time")

    def __init__(self, lat, lon, time, first, last, wrap, missing):
        self.lat, self.lon, self.time = lat, lon, time
        self.first, self.last, self.wrap, self.missing = first, last, wrap, missing

    @property
    def day_of_year(self):
        return _day_of_year(self.time)

    @property
    def shape(self):
        return (len(self.lat), len(self.lon), len(self.time))

    def gate_words(self):
        """uint32 [lat][lon]: first | last << 9 | wrap << 18 (include/ctb.h)."""
        return (self.first.astype(np.uint32) | (self.last.astype(np.uint32) << np.uint32(9)) |
                (self.wrap.astype(np.uint32) << np.uint32(18)))

    @property
    def values(self):
        import torch
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        doy = torch.as_tensor(self.day_of_year.astype(np.int64), device=dev)[None, None, :]
        first = torch.as_tensor(self.first.astype(np.int64), device=dev)[:, :, None]
        last = torch.as_tensor(self.last.astype(np.int64), device=dev)[:, :, None]
        wrap = torch.as_tensor(self.wrap, device=dev)[:, :, None]
        on = ((doy >= first) & (doy <= last)) != wrap
        out = on.to(torch.float64)
        out[torch.as_tensor(self.missing, device=dev)] = float("nan")
        return out.cpu().numpy()


def _day_of_year(time):
    t = np.asarray(time)
    if np.issubdtype(t.dtype, np.datetime64):
        d = t.astype("datetime64[D]")
        return ((d - d.astype("datetime64[Y]").astype("datetime64[D]")).astype(np.int64) + 1).astype(np.int32)
    if np.issubdtype(t.dtype, np.integer):       # the YYYYDDD integers tas_poly writes
        return (t % 1000).astype(np.int32)
    import pandas as pd
    return pd.DatetimeIndex(t).dayofyear.values.astype(np.int32)


def get_daily_growing_season_mask(lat, lon, time, growing_days_path):
    """
    Constructs a mask for days in the within calendar growing season (reference ``:119-153``).

    Parameters
    ----------
    lat, lon, time : coordinate arrays (or DataArray coords) of the climate data
    growing_days_path : str or Dataset
        the crop calendar: variable ``variable`` over (z, latitude, longitude), z = 1 planting day,
        z = 2 harvest day, longitudes on 0..360 (a path needs a netCDF reader, which this image does
        not have; pass the Dataset)

    Returns
    -------
    GrowingSeasonMask
        over lat x lon x time: 1 inside the season, 0 outside; seasons that wrap around the new year
        (harvest < planting) are the complement of [min, max]; a missing harvest day gives 1 everywhere,
        a missing planting day NaN -- the reference's ``where / fillna(1 - mask) / where`` chain.
    """
    if isinstance(growing_days_path, str):
        raise NotImplementedError("reading netCDF needs xarray's backends; pass the crop-calendar Dataset")
    plant, harv, _, _ = _growing_days_arrays(growing_days_path)
    min_day, max_day = season_boundaries(growing_days_path)
    glat = np.asarray(min_day.coords["latitude"].values, dtype=np.float64)
    glon = np.asarray(min_day.coords["longitude"].values, dtype=np.float64)
    lon_src = np.asarray(from_any(growing_days_path)["longitude"].values, dtype=np.float64) - 180
    order = np.argsort(lon_src, kind="stable")
    plant, harv = plant[:, order], harv[:, order]
    mn, mx = np.asarray(min_day.values), np.asarray(max_day.values)
    with np.errstate(invalid="ignore"):
        # day-of-year comparisons against (possibly fractional) day numbers, as integers:
        # doy >= mn <=> doy >= ceil(mn);  doy <= mx <=> doy <= floor(mx);  a NaN bound is never met
        first = np.where(np.isnan(mn), 511, np.clip(np.ceil(mn), 0, 511)).astype(np.int32)
        last = np.where(np.isnan(mx), 0, np.clip(np.floor(mx), 0, 511)).astype(np.int32)
        empty = np.isnan(mn) | np.isnan(mx)
        first = np.where(empty, 511, first)
        last = np.where(empty, 0, last)
        wrap = ~(harv >= plant)                  # harvest < planting, or either missing: fillna(1 - mask)
    missing = np.isnan(plant)
    wrap = wrap & ~missing                       # planting missing: NaN, never counted
    # the data's grid: every (lat, lon) label must exist in the calendar (exact match, like xarray alignment)
    lat = np.asarray(getattr(lat, "values", lat), dtype=np.float64)
    lon = np.asarray(getattr(lon, "values", lon), dtype=np.float64)
    time = np.asarray(getattr(time, "values", time))
    ii = _match(glat, lat, "lat")
    jj = _match(glon, lon, "lon")
    sel = np.ix_(ii, jj)
    return GrowingSeasonMask(lat, lon, time, first[sel], last[sel], wrap[sel], missing[sel])


def _match(have, want, name):
    order = np.argsort(have, kind="stable")
    pos = np.searchsorted(have[order], want)
    pos = np.clip(pos, 0, len(have) - 1)
    idx = order[pos]
    if not np.array_equal(have[idx], want):
        raise KeyError("not all {} values of the data are in the crop calendar".format(name))
    return idx
