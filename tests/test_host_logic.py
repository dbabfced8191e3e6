"""Host-side mirror of the reference interface (named-array shim, lon standardisation,
leap-day removal, renames): transcriptions of the reference's own utils/io tests
(/root/reference/tests/test_climate_toolbox.py:138-228).  CPU only -- these paths
move no data and launch no kernels."""
import numpy as np
import pandas as pd
import pytest

import oracle
from climate_toolbox_b200 import DataArray, Dataset
from climate_toolbox_b200.io import standardize_climate_data, load_bcsd
from climate_toolbox_b200.transformations.transformations import ordinal, tas_poly, snyder_edd
from climate_toolbox_b200.utils.utils import (
    convert_lons_mono, convert_lons_split, remove_leap_days,
    rename_coords_to_lon_and_lat, rename_coords_to_longitude_and_latitude,
    convert_kelvin_to_celsius)


@pytest.fixture
def clim_data(ref_fix):
    lat, lon, time, temp, _ = ref_fix
    return Dataset({"temperature": (["lat", "lon", "time"], temp)},
                   coords={"lon": lon, "lat": lat, "time": time})


def test_clim_data(clim_data):
    assert not clim_data.temperature.isnull().any()
    assert clim_data.temperature.shape == (90, 180, 10)
    assert clim_data.dims["time"] == 10


def test_rename_coords_to_lon_and_lat():
    ds = Dataset(coords={"z": [1.20, 2.58], "long": [156.6, 38.48]})
    ds = rename_coords_to_lon_and_lat(ds)
    coords = ds.coords
    assert "z" not in coords
    assert "lon" in coords and "long" not in coords


def test_rename_coords_to_lon_and_lat2():
    ds = Dataset(coords={"latitude": [71.32, 72.58], "longitude": [156.6, 38.48]})
    ds = rename_coords_to_lon_and_lat(ds)
    coords = ds.coords
    assert "lat" in coords and "latitude" not in coords
    assert "lon" in coords and "longitude" not in coords


def test_rename_coords_to_longitude_and_latitude():
    ds = Dataset(coords={"lat": [71.32, 72.58], "lon": [156.6, 38.48]})
    ds = rename_coords_to_longitude_and_latitude(ds)
    coords = ds.coords
    assert "latitude" in coords and "lat" not in coords
    assert "longitude" in coords and "lon" not in coords


def test_rename_coords_to_longitude_and_latitude_with_clim_data(clim_data):
    ds = rename_coords_to_longitude_and_latitude(clim_data)
    coords = ds.coords
    assert "latitude" in coords and "lat" not in coords
    assert "longitude" in coords and "lon" not in coords
    assert ds.temperature.dims == ("latitude", "longitude", "time")


def test_convert_lons_mono():
    ds = Dataset(coords={"lon": [-156.6, -38.48]})
    expected = np.array([203.4, 321.52])
    ds = convert_lons_mono(ds, lon_name="lon")
    np.testing.assert_array_equal(ds.lon.values, expected)


def test_convert_lons_split():
    ds = Dataset(coords={"longitude": [300, 320]})
    expected = np.array([-60, -40])
    ds = convert_lons_split(ds)
    np.testing.assert_array_equal(ds.longitude.values, expected)


def test_convert_lons_split_is_lazy_and_matches_oracle(clim_data):
    ds = convert_lons_split(clim_data, lon_name="lon")
    new, perm = oracle.convert_lons_split(clim_data.lon.values)
    np.testing.assert_array_equal(ds.lon.values, new)
    var = ds._vars["temperature"]
    assert var.physical is clim_data._vars["temperature"].physical      # no data moved
    np.testing.assert_array_equal(var.takes["lon"], perm)
    np.testing.assert_array_equal(ds.temperature.values,
                                  np.take(clim_data.temperature.values, perm, axis=1))


def test_remove_leap_days():
    da = DataArray(np.random.rand(4, 3),
                   [("time", pd.date_range("2000-02-27", periods=4)), ("space", ["IA", "IL", "IN"])])
    leap_day = np.datetime64("2000-02-29")
    full = da.values
    da = remove_leap_days(da)
    assert leap_day not in da.coords["time"].values
    assert da.shape == (3, 3)
    np.testing.assert_array_equal(da.values, full[[0, 1, 3]])


def test_remove_leap_days_with_clim_data(clim_data):
    leap_day = np.datetime64("2000-02-29")
    da = remove_leap_days(clim_data)
    assert leap_day not in da.coords["time"].values


def test_convert_kelvin_to_celsius_units(clim_data):
    ds = convert_kelvin_to_celsius(clim_data, "temperature")
    assert "C" in ds.data_vars["temperature"].units
    assert ds._vars["temperature"].deferred.kind == "poly"


def test_standardize_climate_data(clim_data):
    ds = standardize_climate_data(clim_data)
    coordinates = ds.coords
    assert "lat" in coordinates and "latitude" not in coordinates
    assert "lon" in coordinates and "longitude" not in coordinates
    assert ds.lon.values.min() < 0 and np.all(np.diff(ds.lon.values) > 0)
    assert load_bcsd(clim_data, "temperature").lon.values.min() < 0


def test_sel_exact_label_keyerror(clim_data):
    with pytest.raises(KeyError):
        clim_data.sel(lon=np.array([0.1250001]))


def test_ordinal():
    assert [ordinal(n) for n in (1, 2, 3, 4, 11, 12, 13, 21, 22, 101)] == \
        ["1st", "2nd", "3rd", "4th", "11th", "12th", "13th", "21st", "22nd", "101st"]


def test_tas_poly_metadata():
    t = pd.date_range("2000-02-27", periods=5)
    ds = Dataset({"tas": (("time", "lat", "lon"), np.zeros((5, 2, 2), dtype=np.float32))},
                 coords={"time": t, "lat": [0.0, 1.0], "lon": [0.0, 1.0]})
    ds1 = tas_poly(ds, 3, "tas-poly-3")
    v = ds1["tas-poly-3"]
    assert v.shape == (4, 2, 2) and v.dims == ("time", "lat", "lon")
    np.testing.assert_array_equal(ds1.time.values, oracle.tas_poly_time_labels(t[[0, 1, 3, 4]]))
    assert v.attrs["units"] == "C^3" and v.attrs["variable"] == "tas-poly-3"
    assert v.attrs["long_title"] == "Daily average temperature (degrees C) raised to the 3rd power"
    assert tas_poly(ds, 1, "tas")["tas"].attrs["units"] == "C"
    big = Dataset({"tas": (("time", "lat"), np.zeros((400, 1)))},
                  coords={"time": pd.date_range("2001-01-01", periods=400), "lat": [0.0]})
    with pytest.raises(ValueError):
        tas_poly(big, 1, "tas")


def test_snyder_preconditions():
    a = DataArray(np.array([1.0, 2.0]), dims=("x",), attrs={"units": "K"})
    b = DataArray(np.array([2.0, 1.0]), dims=("x",), attrs={"units": "K"})
    with pytest.raises(AssertionError):
        snyder_edd(a, b, 1.5)          # tasmax < tasmin somewhere
    c = DataArray(np.array([2.0, 3.0]), dims=("x",), attrs={"units": "C"})
    with pytest.raises(AssertionError):
        snyder_edd(a, c, 1.5)          # unit mismatch
    d = DataArray(np.array([2.0, 3.0]), dims=("x",))
    with pytest.raises(AttributeError):
        snyder_edd(a, d, 1.5)          # no units attr
    r = snyder_edd(a, DataArray(np.array([2.0, 3.0]), dims=("x",), attrs={"units": "K"}), 281.15)
    assert r.units == "degreedays_281.15K"


def test_weights_none_is_typeerror_like_the_reference(clim_data):
    from climate_toolbox_b200.aggregations.aggregations import weighted_aggregate_grid_to_regions
    with pytest.raises(TypeError):
        weighted_aggregate_grid_to_regions(clim_data, "temperature", "popwt", "ISO")


def test_stacked_weight_columns_for_one_pass_multi_weight():
    """SURVEY 8-f2 host logic: K copies of the rows, virtual region k*R + code, NaN labels dropped;
    aggregating the stacked frame with the oracle equals one oracle pass per column."""
    import oracle
    from climate_toolbox_b200 import synthetic
    from climate_toolbox_b200.aggregations.aggregations import _stack_weight_columns
    lat, lon = synthetic.grid_labels(5.0)
    df = synthetic.weights_table(5.0, 40, seed=3).copy()
    df.loc[df.index[::11], "hierid"] = np.nan
    cols = ["popwt", "cropwt"]
    st, labels, combos, present = _stack_weight_columns(df, cols, ["hierid"], "areawt")
    labels = labels["hierid"]
    R = len(labels)
    assert [(c[0], c[1], c[2]) for c in combos] == [("hierid", "popwt", 0), ("hierid", "cropwt", R)]
    assert len(st) == 2 * len(df) and list(labels) == sorted(set(df.hierid.dropna()))
    assert np.isnan(st["_lev"].values).sum() == 2 * df.hierid.isna().sum()
    np.testing.assert_array_equal(present, np.arange(2 * R))
    x = np.random.default_rng(0).normal(280, 10, (6, len(lat), len(lon)))
    got = oracle.weighted_aggregate_grid_to_regions(x, ("time", "lat", "lon"), lat, lon, st, "_w", "_lev", "_bk")[0]
    for k, c in enumerate(cols):
        ref = oracle.weighted_aggregate_grid_to_regions(x, ("time", "lat", "lon"), lat, lon, df, c, "hierid")[0]
        np.testing.assert_allclose(got[:, k * R:(k + 1) * R], ref, rtol=1e-13, equal_nan=True)
    # several region levels: every (level, weight) pair gets its own block of virtual regions
    st2, labels2, combos2, present2 = _stack_weight_columns(df, ["popwt"], ["hierid", "ISO"], "areawt")
    R2 = len(labels2["ISO"])
    assert [(c[0], c[2], c[3]) for c in combos2] == [("hierid", 0, R), ("ISO", R, R2)] and len(st2) == 2 * len(df)
    got2 = oracle.weighted_aggregate_grid_to_regions(x, ("time", "lat", "lon"), lat, lon, st2, "_w", "_lev", "_bk")[0]
    ref_iso = oracle.weighted_aggregate_grid_to_regions(x, ("time", "lat", "lon"), lat, lon, df, "popwt", "ISO")[0]
    sel = np.flatnonzero((present2 >= R) & (present2 < R + R2))
    np.testing.assert_allclose(got2[:, sel], ref_iso, rtol=1e-13, equal_nan=True)


# ----------------------------------------------------------------------------
# growing-season mask (reference utils.py:83-153): factored form vs the dense oracle
# ----------------------------------------------------------------------------
def _crop_calendar(nlat, nlon, rng, lon0=0.125, d=2.0):
    """Synthetic crop calendar on the reference's 0..360 longitude convention: planting / harvest day per
    gridcell, some seasons wrapping around the new year, fractional days, missing dates."""
    from climate_toolbox_b200 import Dataset
    lat = -90 + d / 2 + d * np.arange(nlat)
    lon360 = lon0 + d * np.arange(nlon)
    plant = rng.integers(1, 366, (nlat, nlon)).astype(np.float64)
    harv = rng.integers(1, 366, (nlat, nlon)).astype(np.float64)        # about half of them < planting
    plant[rng.random((nlat, nlon)) < 0.1] += 0.5
    harv[rng.random((nlat, nlon)) < 0.05] = np.nan
    plant[rng.random((nlat, nlon)) < 0.05] = np.nan
    gd = Dataset({"variable": (("z", "latitude", "longitude"), np.stack([plant, harv]))},
                 coords={"z": np.array([1, 2]), "latitude": lat, "longitude": lon360})
    return gd, lat, lon360, plant, harv


def test_growing_season_mask_matches_oracle():
    import oracle
    from climate_toolbox_b200.utils.utils import get_daily_growing_season_mask, season_boundaries
    rng = np.random.default_rng(5)
    gd, lat, lon360, plant, harv = _crop_calendar(12, 20, rng)
    time = pd.date_range("2003-11-20", periods=120).values          # crosses the new year
    lon = lon360 - 180                                               # the calendar's shift (utils.py:87)
    order = np.argsort(lon)
    mn, mx = season_boundaries(gd)
    omn, omx = oracle.season_boundaries(plant[:, order], harv[:, order])
    np.testing.assert_array_equal(mn.values, omn)
    np.testing.assert_array_equal(mx.values, omx)
    np.testing.assert_array_equal(mn.coords["longitude"].values, lon[order])
    m = get_daily_growing_season_mask(lat, lon[order], time, gd)
    assert m.shape == (12, 20, 120) and m.dims == ("lat", "lon", "time")
    doy = pd.DatetimeIndex(time).dayofyear.values
    ref = oracle.growing_season_mask(plant[:, order], harv[:, order], doy)
    got = m.values
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_array_equal(np.nan_to_num(got, nan=-1), np.nan_to_num(ref, nan=-1))
    assert np.isnan(ref).any() and (ref == 0).any() and (ref == 1).any()
    # a subset / permutation of the grid's labels selects the matching gridcells; an unknown label raises
    m2 = get_daily_growing_season_mask(lat[3:7], lon[order][::-1][:5], time, gd)
    np.testing.assert_array_equal(np.nan_to_num(m2.values, nan=-1),
                                  np.nan_to_num(ref[3:7][:, ::-1][:, :5], nan=-1))
    with pytest.raises(KeyError):
        get_daily_growing_season_mask(lat + 0.01, lon[order], time, gd)


def test_weights_csv_disk_cache_round_trip(tmp_path):
    """the parsed weights frame persists across processes (Arrow IPC next to nothing else);
    the cached frame equals the parsed one, and a rewritten CSV is parsed again."""
    import time

    from climate_toolbox_b200.aggregations import aggregations as A
    rng = np.random.default_rng(0)
    n = 500
    csv = pd.DataFrame({"pix_cent_x": rng.choice([-179.875, 0.125, 180.125], n), "pix_cent_y": rng.normal(size=n),
                        "hierid": ["R%03d" % i for i in rng.integers(0, 40, n)], "popwt": rng.random(n),
                        "areawt": rng.random(n)})
    p = tmp_path / "w.csv"
    csv.to_csv(p, index=False)
    cache = tmp_path / "cache"
    A.prepare_spatial_weights_data.cache_clear()
    a = A.prepare_spatial_weights_data(str(p), cache_dir=str(cache))
    files = list(cache.iterdir())
    assert len(files) == 1 and files[0].suffix == ".feather"
    A.prepare_spatial_weights_data.cache_clear()          # "next process"
    b = A.prepare_spatial_weights_data(str(p), cache_dir=str(cache))
    pd.testing.assert_frame_equal(a, b)
    assert b.index.name == "reshape_index" and not (b["lon"] == 180.125).any()
    plain = oracle.prepare_spatial_weights_data(str(p))
    pd.testing.assert_frame_equal(b.reset_index(drop=True), plain.reset_index(drop=True), check_dtype=False)
    # a rewritten file has another key
    time.sleep(0.01)
    csv.iloc[:10].to_csv(p, index=False)
    A.prepare_spatial_weights_data.cache_clear()
    c = A.prepare_spatial_weights_data(str(p), cache_dir=str(cache))
    assert len(c) == 10 and len(list(cache.iterdir())) == 2
