"""The oracle against every golden vector / known answer the reference's own tests hold
for this path (/root/reference/tests/test_climate_toolbox.py), plus an independent
formulation.  CPU only."""
import os

import numpy as np
import pandas as pd
import pytest

import oracle
from conftest import reference_fixtures

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_fixture.npz")


def test_reference_test_reindex_spatial_weights(ref_fix):
    # reference tests :109-116
    lat, lon, time, temp, df = ref_fix
    assert not np.isnan(temp).any()
    g, dims, i, j = oracle.reindex_spatial_data_to_regions(temp, ("lat", "lon", "time"), lat, lon, df)
    assert g.shape == (len(df["lon"]), len(time)) == (100, 10)
    assert "reshape_index" in dims
    assert np.array_equal(lat[i], df["lat"].values) and np.array_equal(lon[j], df["lon"].values)
    assert np.array_equal(g, temp[i, j, :])


def test_reference_test_weighting(ref_fix):
    # reference tests :119-135
    lat, lon, time, temp, df = ref_fix
    assert np.isnan(df["popwt"].values).any()
    g, dims, _, _ = oracle.reindex_spatial_data_to_regions(temp, ("lat", "lon", "time"), lat, lon, df)
    assert not np.isnan(g).any()
    for aggwt in ("popwt", "areawt"):
        out, od, labels = oracle.aggregate_reindexed_data_to_regions(g, dims, df, aggwt, "ISO")
        assert od == ("ISO", "time") and out.shape == (len(labels), 10)
        assert not np.isnan(out).any()
        assert list(labels) == sorted(set(df["ISO"]))


def test_time_major_dim_order(ref_fix):
    lat, lon, time, temp, df = ref_fix
    x = np.ascontiguousarray(np.transpose(temp, (2, 0, 1)))
    a, ad, _ = oracle.weighted_aggregate_grid_to_regions(x, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
    b, bd, _ = oracle.weighted_aggregate_grid_to_regions(temp, ("lat", "lon", "time"), lat, lon, df, "popwt", "hierid")
    assert ad == ("time", "hierid") and bd == ("hierid", "time")
    np.testing.assert_array_equal(a, b.T)


def test_golden_fixture_unchanged(ref_fix):
    lat, lon, time, temp, df = ref_fix
    G = np.load(GOLD, allow_pickle=True)
    np.testing.assert_array_equal(G["temp_checksum"], [temp.sum(), temp[3, 5, 7]])
    g, gd, i, j = oracle.reindex_spatial_data_to_regions(temp, ("lat", "lon", "time"), lat, lon, df)
    np.testing.assert_array_equal(G["lat_pos"], i)
    np.testing.assert_array_equal(G["lon_pos"], j)
    np.testing.assert_array_equal(G["reindexed"], g)
    for aggwt in ("popwt", "areawt"):
        for agglev in ("ISO", "hierid"):
            v, _, labels = oracle.aggregate_reindexed_data_to_regions(g, gd, df, aggwt, agglev)
            np.testing.assert_allclose(v, G["agg_{}_{}".format(aggwt, agglev)], rtol=1e-14)
            np.testing.assert_array_equal(labels, G["labels_" + agglev])
    # SURVEY.md 8c check values from an independent restatement
    np.testing.assert_allclose(G["agg_popwt_ISO"][0, :3], [58.62485545, 53.81246705, 46.68234678], atol=5e-9)


def test_fast_and_loop_paths_agree(ref_fix):
    lat, lon, time, temp, df = ref_fix
    g, gd, _, _ = oracle.reindex_spatial_data_to_regions(temp, ("lat", "lon", "time"), lat, lon, df)
    a = oracle.aggregate_reindexed_data_to_regions(g, gd, df, "popwt", "hierid", fast=True)[0]
    b = oracle.aggregate_reindexed_data_to_regions(g, gd, df, "popwt", "hierid", fast=False)[0]
    np.testing.assert_allclose(a, b, rtol=1e-14)


def test_oracle_vs_scipy_csr():
    """Independent formulation: Y = D^-1 A nan0(X) with A a scipy CSR matrix."""
    import scipy.sparse as sp
    from climate_toolbox_b200 import synthetic

    lat, lon = synthetic.grid_labels(2.0)
    df = synthetic.weights_table(2.0, 200, seed=11)
    tas, _, _ = synthetic.tas_field(12, len(lat), len(lon), seed=2, nan_frac=0.01, dtype=np.float64)
    out, dims, labels = oracle.weighted_aggregate_grid_to_regions(
        tas, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
    w = oracle.effective_weights(df, "popwt")
    codes, _ = pd.factorize(df["hierid"].values, sort=True)
    i = np.searchsorted(lat, df["lat"].values)
    j = np.searchsorted(lon, df["lon"].values)
    cell = i * len(lon) + j
    ok = ~np.isnan(w)
    A = sp.csr_matrix((w[ok], (codes[ok], cell[ok])), shape=(len(labels), len(lat) * len(lon)))
    X = np.nan_to_num(tas.reshape(12, -1), nan=0.0).T
    den = np.asarray(A.sum(axis=1)).ravel()
    ref = (A @ X / den[:, None]).T
    np.testing.assert_allclose(out, ref, rtol=1e-12)


def test_weight_fallback_semantics():
    # aggregations.py:73: NaN, 0 and negative primary weights fall back per ROW; NaN backup stays NaN
    df = pd.DataFrame({"popwt": [1.0, 0.0, -2.0, np.nan, np.nan, 3.0],
                       "areawt": [9.0, 8.0, 7.0, 6.0, np.nan, -1.0]})
    w = oracle.effective_weights(df, "popwt")
    np.testing.assert_array_equal(w, [1.0, 8.0, 7.0, 6.0, np.nan, 3.0])
    # backup == primary: zeros and negatives survive
    np.testing.assert_array_equal(oracle.effective_weights(df, "areawt"), df["areawt"].values)


def test_nan_semantics():
    """NaN data drops from the numerator only; all-NaN region-day -> 0; den == 0 -> NaN;
    NaN region labels are dropped."""
    lat, lon = np.array([0.0, 1.0]), np.array([0.0, 1.0, 2.0])
    x = np.array([[[1.0, np.nan, 3.0], [np.nan, np.nan, 6.0]]])  # (time=1, lat, lon)
    df = pd.DataFrame({"lat": [0, 0, 0, 1, 1, 1, 1.0], "lon": [0, 1, 2, 0, 1, 2, 2.0],
                       "r": ["a", "a", "b", "c", "c", "d", np.nan],
                       "w": [1.0, 3.0, 2.0, 1.0, 1.0, 0.0, 5.0], "areawt": [0.0] * 7})
    out, dims, labels = oracle.weighted_aggregate_grid_to_regions(x, ("time", "lat", "lon"), lat, lon, df, "w", "r")
    assert list(labels) == ["a", "b", "c", "d"]
    np.testing.assert_array_equal(out[0, :3], [1.0 / 4.0, 3.0, 0.0])
    assert np.isnan(out[0, 3])  # den = 0 (w=0 falls back to areawt=0)


def test_label_miss_raises_keyerror(ref_fix):
    lat, lon, time, temp, df = ref_fix
    bad = df.copy()
    bad.loc[3, "lon"] = bad.loc[3, "lon"] + 1e-9
    with pytest.raises(KeyError):
        oracle.reindex_spatial_data_to_regions(temp, ("lat", "lon", "time"), lat, lon, bad)


def test_reference_test_convert_lons():
    # reference tests :175-190
    new, perm = oracle.convert_lons_mono([-156.6, -38.48])
    np.testing.assert_array_equal(new, np.array([203.4, 321.52]))
    new, perm = oracle.convert_lons_split([300, 320])
    np.testing.assert_array_equal(new, np.array([-60, -40]))


def test_lon_split_is_exact_roll():
    for d in (1.0, 0.25):
        lon360 = d / 2 + d * np.arange(int(360 / d))
        new, perm = oracle.convert_lons_split(lon360)
        n = len(lon360)
        np.testing.assert_array_equal(perm, np.roll(np.arange(n), n // 2))
        np.testing.assert_array_equal(new, -180 + d / 2 + d * np.arange(n))


def test_reference_test_remove_leap_days():
    # reference tests :193-213
    t = pd.date_range("2000-02-27", periods=4)
    keep = oracle.leap_day_keep_mask(t)
    assert np.datetime64("2000-02-29") not in t.values[keep] and keep.sum() == 3
    t = pd.date_range(start=pd.Timestamp(2000, 1, 1), periods=10, freq="D")
    assert oracle.leap_day_keep_mask(t).all()


def test_reference_test_snyder_edd():
    # reference tests :231-251
    tmax = np.array([280.4963, 280.7887])
    tmin = np.array([278.902, 278.23163])
    res = oracle.snyder_edd(tmin, tmax, 273.15 + 8)
    assert oracle.snyder_edd_units(273.15 + 8, "K") == "degreedays_281.15K"
    assert res.sum() == 0.0


def test_reference_test_snyder_gdd():
    # reference tests :254-278
    tmax = np.array([280.4963, 280.7887])
    tmin = np.array([278.902, 278.23163])
    res = oracle.snyder_gdd(tmin, tmax, 273.15 + 1, 273.15 + 8)
    assert oracle.snyder_gdd_units(273.15 + 1, 273.15 + 8, "K") == "degreedays_274.15-281.15K"
    assert res.sum() == pytest.approx(11, 0.1)
    assert res.sum() == pytest.approx(10.909315, abs=1e-6)  # SURVEY.md section 4 probe value


def test_snyder_branches_and_nan():
    e = 10.0
    tmin = np.array([12.0, 2.0, 5.0, np.nan, 5.0, 8.0])
    tmax = np.array([20.0, 8.0, 15.0, 15.0, np.nan, 8.0])
    r = oracle.snyder_edd(tmin, tmax, e)
    assert r[0] == 16.0 - 10.0 and r[1] == 0.0
    M, W = 10.0, 5.0
    th = np.arcsin((e - M) / W)
    assert r[2] == pytest.approx(((M - e) * (np.pi / 2 - th) + W * np.cos(th)) / np.pi)
    assert np.isnan(r[3]) and r[4] == 0.0 and r[5] == 0.0
    with pytest.raises(AssertionError):
        oracle.snyder_edd(np.array([5.0]), np.array([4.0]), e)


def test_tas_poly():
    t = pd.date_range("2000-02-27", periods=5)
    x = np.arange(5 * 2, dtype=np.float32).reshape(5, 2) + 270
    out, labels = oracle.tas_poly(x, t, 3)
    assert out.shape == (4, 2) and out.dtype == np.float64
    np.testing.assert_array_equal(labels, 2000 * 1000 + np.arange(1, 5))
    np.testing.assert_allclose(out, (x[[0, 1, 3, 4]].astype(np.float64) - 273.15) ** 3)
    with pytest.raises(ValueError):
        oracle.tas_poly(np.zeros((400, 1)), pd.date_range("2001-01-01", periods=400), 1)


def test_prepare_spatial_weights_data(tmp_path):
    p = tmp_path / "w.csv"
    pd.DataFrame({"pix_cent_x": [180.125, 10.125, 10.125], "pix_cent_y": [0.125, 1.125, 1.125],
                  "hierid": ["a", "b", "b"], "areawt": [1.0, 2.0, 2.0]}).to_csv(p, index=False)
    df = oracle.prepare_spatial_weights_data(str(p))
    assert list(df["lon"]) == [-179.875, 10.125, 10.125] and len(df) == 3  # duplicates kept
    assert df.index.names == ["reshape_index"] and "lat" in df
