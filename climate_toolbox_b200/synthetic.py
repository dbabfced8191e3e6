"""Deterministic synthetic grids, weight tables and fields.

Used by the parity tests and by ``bench.py``; the real CIL segment-weights file and
BCSD netCDFs are not part of the reference repository.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

__all__ = ["grid_labels", "weights_table", "tas_field", "CONFIGS"]

CONFIGS = {
    # name: (resolution deg, n_regions, T)
    "config1": (1.0, 3000, 365),
    "config2": (0.25, 24378, 1460),
}


def grid_labels(d, lon_0_360=False):
    """Cell centres ``lat_i = -90 + d/2 + d*i``; ``lon_j = -180 + d/2 + d*j`` (or the
    0..360 variant ``d/2 + d*j`` that exercises the lon-roll path)."""
    nlat, nlon = int(round(180 / d)), int(round(360 / d))
    lat = -90 + d / 2 + d * np.arange(nlat)
    lon = (0.0 if lon_0_360 else -180.0) + d / 2 + d * np.arange(nlon)
    return lat, lon


def weights_table(d, n_regions, seed=1234, land_frac=0.30, dup_frac=0.35):
    """Synthetic segment-weights table with the reference's columns
    (``lat, lon, hierid, ISO, areawt, popwt, cropwt``; aggregations.py:49-53,100-106).

    Land mask = Gaussian-smoothed noise thresholded at its (1-land_frac) quantile;
    ``n_regions`` Voronoi regions over land cells; ``dup_frac`` of land cells also
    belong to their 2nd-nearest region (border pixels -> duplicated gridcells);
    rows randomly permuted; ``popwt`` has 20 % rows zero/NaN, ``cropwt`` 60 %.
    Labels are always the standardised -180..180 ones.
    """
    from scipy.ndimage import gaussian_filter
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    lat, lon = grid_labels(d)
    nlat, nlon = len(lat), len(lon)
    field = gaussian_filter(rng.standard_normal((nlat, nlon)), sigma=nlat / 30.0, mode="wrap")
    land = field > np.quantile(field, 1.0 - land_frac)
    li, lj = np.nonzero(land)
    n_land = len(li)
    n_regions = min(n_regions, n_land)
    seeds = rng.choice(n_land, size=n_regions, replace=False)
    tree = cKDTree(np.c_[li[seeds], lj[seeds]].astype(np.float64))
    k = 2 if n_regions > 1 else 1
    _, nn = tree.query(np.c_[li, lj].astype(np.float64), k=k)
    nn = nn.reshape(n_land, k)
    dup = rng.random(n_land) < dup_frac if k == 2 else np.zeros(n_land, bool)
    ri = np.r_[li, li[dup]]
    rj = np.r_[lj, lj[dup]]
    reg = np.r_[nn[:, 0], nn[dup, 1]] if k == 2 else nn[:, 0]
    perm = rng.permutation(len(ri))
    ri, rj, reg = ri[perm], rj[perm], reg[perm]
    n = len(ri)

    def holes(w, frac):
        w = w.copy()
        bad = rng.random(n) < frac
        nan = bad & (rng.random(n) < 0.5)
        w[bad] = 0.0
        w[nan] = np.nan
        return w

    df = pd.DataFrame({
        "lat": lat[ri], "lon": lon[rj],
        "hierid": np.char.add("R", np.char.zfill(reg.astype(str), 5)),
        "ISO": np.char.add("C", np.char.zfill((reg % 180).astype(str), 3)),
        "areawt": rng.random(n) * np.cos(np.deg2rad(lat[ri])),
        "popwt": holes(rng.lognormal(0.0, 2.0, n), 0.20),
        "cropwt": holes(rng.lognormal(0.0, 2.0, n), 0.60),
    })
    df.index.names = ["reshape_index"]
    return df


def tas_field(T, nlat, nlon, seed=7, nan_frac=0.0, dtype=np.float32):
    """``tas = 288 + 10 N(0,1)`` K plus a ``(tasmin, tasmax)`` pair with
    ``tasmax >= tasmin`` everywhere (transformations.py:66).  Layout [T][lat][lon]."""
    rng = np.random.default_rng(seed)
    tas = (288.0 + 10.0 * rng.standard_normal((T, nlat, nlon))).astype(dtype)
    spread_lo = np.abs(3.0 * rng.standard_normal((T, nlat, nlon))).astype(dtype)
    spread_hi = np.abs(3.0 * rng.standard_normal((T, nlat, nlon))).astype(dtype)
    tasmin, tasmax = tas - spread_lo, tas + spread_hi
    if nan_frac > 0:
        m = rng.random((T, nlat, nlon)) < nan_frac
        tas[m] = np.nan
        m2 = rng.random((T, nlat, nlon)) < nan_frac
        tasmin[m2] = np.nan
        m3 = rng.random((T, nlat, nlon)) < nan_frac
        tasmax[m3] = np.nan
        # a NaN tasmax with a finite tasmin is allowed by the reference's assert
    return tas, tasmin, tasmax
