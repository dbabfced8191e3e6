"""numpy/pandas restatement of climate_toolbox's grid->region aggregation path.

TEST INFRASTRUCTURE.  This file is the *checker* the CUDA path is compared
with.  It is never imported by the product package ``climate_toolbox_b200``.

What it restates (all paths relative to /root/reference):

* ``climate_toolbox/aggregations/aggregations.py:8-32``   -> :func:`reindex_spatial_data_to_regions`
* ``climate_toolbox/aggregations/aggregations.py:35-84``  -> :func:`aggregate_reindexed_data_to_regions`
* ``climate_toolbox/aggregations/aggregations.py:87-124`` -> :func:`weighted_aggregate_grid_to_regions`
* ``climate_toolbox/aggregations/aggregations.py:127-152``-> :func:`prepare_spatial_weights_data`
* ``climate_toolbox/transformations/transformations.py:7-93``    -> :func:`snyder_edd`
* ``climate_toolbox/transformations/transformations.py:96-147``  -> :func:`snyder_gdd`
* ``climate_toolbox/transformations/transformations.py:160-208`` -> :func:`tas_poly`
* ``climate_toolbox/utils/utils.py:23-40``  -> :func:`convert_lons_mono`, :func:`convert_lons_split`
* ``climate_toolbox/utils/utils.py:77-80``  -> :func:`leap_day_keep_mask`

The arithmetic of the reference lives in third-party, un-vendored, unpinned
dependencies (``xarray`` / ``pandas`` / ``numpy`` -- ``requirements_dev.txt:11-15``
lists them without versions; ``xarray`` and ``toolz`` are NOT installed in this
image and there is no network), so the reference itself cannot be executed here.
The semantics restated are those of the xarray calls at the cited lines:

* ``Dataset.sel(lon=<DataArray>, lat=<DataArray>)`` = exact float64 label match
  (``pandas.Index.get_indexer``; a miss raises ``KeyError``), pointwise gather,
  the pair (lat, lon) replaced by ONE dim ``reshape_index`` at the position of
  the first of the two;
* ``.where(w > 0).fillna(backup)`` = per-ROW fallback weight;
* ``groupby(agglev).sum(dim="reshape_index")`` = skip-NaN sum per sorted unique
  label (``pd.factorize(sort=True)``; NaN labels dropped); all-NaN group -> 0.

PARITY PINNING.  The reference's own tests pin only: two Snyder known answers
(``tests/test_climate_toolbox.py:231-278``), the two lon-conversion arrays
(``:175-190``), the leap-day removal (``:193-213``) and shapes / no-NaN
properties of the aggregation (``:109-135``).  ``tests/test_oracle.py`` checks
the oracle against every one of those.  For the aggregated VALUES themselves
the reference holds no golden vector and cannot be run here: **parity
unpinned** for the aggregation arithmetic -- this restatement is the
definition of parity, cross-checked only against an independent scipy-CSR
formulation (``tests/test_oracle.py::test_oracle_vs_scipy_csr``).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

__all__ = [
    "time_group_sum",
    "season_boundaries",
    "growing_season_mask",
    "convert_lons_mono",
    "convert_lons_split",
    "leap_day_keep_mask",
    "exact_label_index",
    "reindex_spatial_data_to_regions",
    "effective_weights",
    "aggregate_reindexed_data_to_regions",
    "weighted_aggregate_grid_to_regions",
    "prepare_spatial_weights_data",
    "snyder_edd",
    "snyder_edd_units",
    "snyder_gdd_units",
    "snyder_gdd",
    "tas_poly",
    "tas_poly_time_labels",
]


# ---------------------------------------------------------------------------
# utils.py
# ---------------------------------------------------------------------------
def convert_lons_mono(lon):
    """utils.py:23-30.  -180..180 -> 0..360, then sort the axis.

    Returns ``(new_sorted_labels, perm)`` where ``perm[j_new] = j_old`` so that
    ``data_new = np.take(data_old, perm, axis=lon_axis)``.
    """
    lon = np.asarray(lon, dtype=np.float64) % 360
    perm = _sel_positions(lon, np.sort(lon))
    return lon[perm], perm


def convert_lons_split(lon):
    """utils.py:33-40.  0..360 -> -180..180, then sort the axis."""
    lon = (np.asarray(lon, dtype=np.float64) + 180) % 360 - 180
    perm = _sel_positions(lon, np.sort(lon))
    return lon[perm], perm


def _sel_positions(labels, wanted):
    """``ds.sel(lon=wanted)`` on an axis labelled ``labels`` (exact match)."""
    return exact_label_index(labels, wanted, "lon")


def leap_day_keep_mask(times):
    """utils.py:77-80.  True for every step that is NOT Feb 29."""
    idx = pd.DatetimeIndex(np.asarray(times))
    return ~((idx.month == 2) & (idx.day == 29))


# ---------------------------------------------------------------------------
# aggregations.py
# ---------------------------------------------------------------------------
def exact_label_index(grid_labels, wanted, name="label"):
    """Exact-equality label -> position lookup (aggregations.py:27 via xarray
    ``.sel`` without ``method=``).  Raises ``KeyError`` on any miss."""
    index = pd.Index(np.asarray(grid_labels))
    pos = index.get_indexer(np.asarray(wanted))
    if (pos < 0).any():
        raise KeyError("not all values found in index {!r}".format(name))
    return pos.astype(np.int64)


def reindex_spatial_data_to_regions(x, dims, lat, lon, df):
    """aggregations.py:8-32.

    ``x`` has named axes ``dims`` containing ``"lat"`` and ``"lon"``.  Returns
    ``(x_reindexed, new_dims, lat_pos, lon_pos)``: the pair (lat, lon) is
    replaced by ``"reshape_index"`` (length ``len(df)``) at the position of the
    first of the two dims.
    """
    dims = tuple(dims)
    a_lat, a_lon = dims.index("lat"), dims.index("lon")
    i = exact_label_index(lat, df["lat"].values, "lat")
    j = exact_label_index(lon, df["lon"].values, "lon")
    first = min(a_lat, a_lon)
    # move (lat, lon) to the front, gather pointwise, move the new axis back
    rest = [d for d in range(len(dims)) if d not in (a_lat, a_lon)]
    xt = np.transpose(np.asarray(x), [a_lat, a_lon] + rest)
    g = xt[i, j]  # (nnz, *rest)
    new_dims = [d for k, d in enumerate(dims) if k not in (a_lat, a_lon)]
    new_dims.insert(first, "reshape_index")
    # axis 0 of g is reshape_index; rest follow in original relative order
    g = np.moveaxis(g, 0, first)
    return g, tuple(new_dims), i, j


def effective_weights(df, aggwt, backup_aggwt="areawt"):
    """aggregations.py:69-73: ``w.where(w > 0).fillna(backup)`` per row."""
    w = np.asarray(df[aggwt].values, dtype=np.float64)
    b = np.asarray(df[backup_aggwt].values, dtype=np.float64)
    w = np.where(w > 0, w, np.nan)          # NaN, 0 and negatives -> NaN
    return np.where(np.isnan(w), b, w)      # fillna(backup): NaN backup stays NaN


def _group_codes(labels):
    """Sorted unique group labels + integer code per row (NaN label -> -1)."""
    codes, uniques = pd.factorize(np.asarray(labels), sort=True)
    return codes.astype(np.int64), np.asarray(uniques)


def aggregate_reindexed_data_to_regions(
    x, dims, df, aggwt, agglev, backup_aggwt="areawt", fast=True
):
    """aggregations.py:35-84.

    ``x`` has a ``"reshape_index"`` axis (one entry per row of ``df``).  Returns
    ``(out, out_dims, region_labels)`` with ``agglev`` taking the place of
    ``reshape_index``; float64.

    ``num = sum_k nan->0(x_k * w_k)``, ``den = sum_k nan->0(w_k)`` -- NaN data
    drops out of the numerator only, its weight still counts in ``den``.
    """
    dims = tuple(dims)
    ax = dims.index("reshape_index")
    w = effective_weights(df, aggwt, backup_aggwt)
    codes, labels = _group_codes(df[agglev].values)
    R = len(labels)

    xm = np.moveaxis(np.asarray(x), ax, -1)  # (..., nnz)
    prod = xm * w                            # f32*f64 -> f64 (aggregations.py:78)
    prod = np.where(np.isnan(prod), 0.0, prod)
    wz = np.where(np.isnan(w), 0.0, w)

    keep = codes >= 0
    num = np.zeros(prod.shape[:-1] + (R,), dtype=np.float64)
    den = np.zeros(R, dtype=np.float64)
    if fast and keep.any():
        order = np.argsort(codes[keep], kind="stable")
        kidx = np.flatnonzero(keep)[order]
        sc = codes[kidx]
        starts = np.flatnonzero(np.r_[True, sc[1:] != sc[:-1]])
        present = sc[starts]
        num[..., present] = np.add.reduceat(prod[..., kidx], starts, axis=-1)
        den[present] = np.add.reduceat(wz[kidx], starts)
    else:
        for r in range(R):
            m = codes == r
            num[..., r] = prod[..., m].sum(axis=-1)
            den[r] = wz[m].sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        out = num / den
    out = np.moveaxis(out, -1, ax)
    out_dims = tuple(agglev if d == "reshape_index" else d for d in dims)
    return out, out_dims, labels


def weighted_aggregate_grid_to_regions(x, dims, lat, lon, df, aggwt, agglev,
                                       backup_aggwt="areawt"):
    """aggregations.py:87-124 (with ``weights`` given -- the ``weights=None``
    default is a TypeError in the reference, :118-119 vs :128)."""
    g, gd, _, _ = reindex_spatial_data_to_regions(x, dims, lat, lon, df)
    return aggregate_reindexed_data_to_regions(g, gd, df, aggwt, agglev, backup_aggwt)


def prepare_spatial_weights_data(weights_file):
    """aggregations.py:127-152, *intent* of :144 (``df.set_value`` was removed in
    pandas 1.0): relabel ``pix_cent_x == 180.125 -> -179.875``.  ``:147``
    ``drop_duplicates()`` discards its result -- duplicates are kept."""
    df = pd.read_csv(weights_file)
    df.loc[df["pix_cent_x"] == 180.125, "pix_cent_x"] = -179.875
    df.index.names = ["reshape_index"]
    return df.rename(columns={"pix_cent_x": "lon", "pix_cent_y": "lat"})


# ---------------------------------------------------------------------------
# transformations.py
# ---------------------------------------------------------------------------
def snyder_edd(tasmin, tasmax, threshold):
    """transformations.py:62-89 in float64."""
    tasmin = np.asarray(tasmin, dtype=np.float64)
    tasmax = np.asarray(tasmax, dtype=np.float64)
    assert not (tasmax < tasmin).any(), "values encountered where tasmin > tasmax"
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = (tasmax + tasmin) / 2
        width = (tasmax - tasmin) / 2
        theta = np.arcsin((threshold - mean) / width)
        res = np.where(
            tasmin < threshold,
            np.where(
                tasmax > threshold,
                ((mean - threshold) * (np.pi / 2 - theta) + width * np.cos(theta)) / np.pi,
                0,
            ),
            mean - threshold,
        )
    return res


def snyder_edd_units(threshold, units):
    """transformations.py:91"""
    return "degreedays_{}{}".format(threshold, units)


def snyder_gdd(tasmin, tasmax, threshold_low, threshold_high):
    """transformations.py:139-141"""
    return snyder_edd(tasmin, tasmax, threshold_low) - snyder_edd(tasmin, tasmax, threshold_high)


def snyder_gdd_units(threshold_low, threshold_high, units):
    """transformations.py:143-145"""
    return "degreedays_{}-{}{}".format(threshold_low, threshold_high, units)


def tas_poly_time_labels(times_kept):
    """transformations.py:195-199: ``year * 1000 + (1..n)`` (ordinal position
    after leap-day removal, not calendar day-of-year)."""
    idx = pd.DatetimeIndex(np.asarray(times_kept))
    return np.asarray(idx.year, dtype=np.int64) * 1000 + np.arange(1, len(idx) + 1)


def tas_poly(tas, times, power, time_axis=0):
    """transformations.py:183-199.  Arithmetic in float64 (inputs up-cast before
    the first operation -- DESIGN.md "dtype policy").  Returns
    ``(values, new_time_labels)``."""
    keep = leap_day_keep_mask(times)
    tas = np.compress(keep, np.asarray(tas), axis=time_axis).astype(np.float64)
    if tas.shape[time_axis] > 365:
        raise ValueError
    out = (tas - 273.15) ** power
    return out, tas_poly_time_labels(np.asarray(times)[keep])


# ---------------------------------------------------------------------------
# temporal sums after the aggregation
# ---------------------------------------------------------------------------
def time_group_sum(values, dims, labels, time_dim="time"):
    """Sum of the daily region values over runs of equal ``labels`` along ``time_dim`` -- what
    ``out.groupby(label).sum()`` gives for period labels that are runs of consecutive steps
    (``EDD_P = sum_d EDD_d``, /root/reference/climate_toolbox/transformations/transformations.py:17-21).
    A plain sum: NaN and infinities propagate.  Returns (array, unique labels in order)."""
    values = np.asarray(values, dtype=np.float64)
    labels = np.asarray(labels)
    ax = list(dims).index(time_dim)
    n = values.shape[ax]
    assert len(labels) == n
    change = np.ones(n, dtype=bool)
    change[1:] = labels[1:] != labels[:-1]
    starts = np.flatnonzero(change)
    ends = np.r_[starts[1:], n]
    parts = [np.take(values, np.arange(a, b), axis=ax).sum(axis=ax, keepdims=True) for a, b in zip(starts, ends)]
    out = np.concatenate(parts, axis=ax) if parts else np.take(values, [], axis=ax)
    return out, labels[change]


# ---------------------------------------------------------------------------
# growing-season mask
# ---------------------------------------------------------------------------
def season_boundaries(planting, harvest):
    """/root/reference/climate_toolbox/utils/utils.py:83-116 on plain arrays [lat][lon]: the two z planes
    sorted per gridcell (np.sort: NaN last) -> (min_day, max_day)."""
    both = np.sort(np.stack([np.asarray(planting, float), np.asarray(harvest, float)], axis=2), axis=2)
    return both[:, :, 0], both[:, :, 1]


def growing_season_mask(planting, harvest, day_of_year):
    """utils.py:119-153 on plain arrays: the dense [lat][lon][time] mask of 1 / 0 / NaN.
    ``mask = (doy >= min) & (doy <= max)``; ``.where(harvest >= planting)`` blanks wrap-around seasons
    (and gridcells with a missing date), ``.fillna(1 - mask)`` fills those with the complement,
    ``.where(~isnan(planting))`` blanks gridcells without a planting date."""
    planting, harvest = np.asarray(planting, float), np.asarray(harvest, float)
    mn, mx = season_boundaries(planting, harvest)
    doy = np.asarray(day_of_year)[None, None, :]
    with np.errstate(invalid="ignore"):
        mask = ((doy >= mn[:, :, None]) & (doy <= mx[:, :, None])).astype(np.float64)
        keep = (harvest >= planting)[:, :, None]
    out = np.where(keep, mask, np.nan)
    out = np.where(np.isnan(out), 1.0 - mask, out)
    out = np.where(np.isnan(planting)[:, :, None], np.nan, out)
    return out
