"""Parity of the CUDA path (through the public API and the C-ABI) against the oracle.

Tolerances: index mapping bit-exact; aggregates |got-ref| <= 1e-9 * max(|ref|, S) where
S = sum(w|f(x)|)/sum(w) is the cancellation-aware scale of SURVEY.md 7.3-6 (odd powers of
degC and EDD near 0 cancel to ~0, where a pure relative error is undefined)."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

import oracle
from conftest import rel_err
from climate_toolbox_b200 import DataArray, Dataset, synthetic
from climate_toolbox_b200 import _engine as E
from climate_toolbox_b200 import _native as N
from climate_toolbox_b200.aggregations.aggregations import (
    _aggregate_reindexed_data_to_regions, _reindex_spatial_data_to_regions,
    weighted_aggregate_grid_to_regions, weighted_aggregate_grid_to_regions_multi)
from climate_toolbox_b200.io import load_bcsd
from climate_toolbox_b200.transformations.transformations import snyder_edd, snyder_gdd, tas_poly

pytestmark = pytest.mark.gpu
TOL = 1e-9
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_fixture.npz")


def check(got, ref, scale=None, tol=TOL):
    err = rel_err(got, ref, scale)
    assert err <= tol, err
    return err


def oracle_agg(x, dims, lat, lon, df, aggwt, agglev, backup="areawt"):
    ref, rd, labels = oracle.weighted_aggregate_grid_to_regions(x, dims, lat, lon, df, aggwt, agglev, backup)
    scale = oracle.weighted_aggregate_grid_to_regions(np.abs(x), dims, lat, lon, df, aggwt, agglev, backup)[0]
    return ref, rd, labels, scale


@pytest.fixture
def clim_data(ref_fix):
    lat, lon, time, temp, _ = ref_fix
    return Dataset({"temperature": (["lat", "lon", "time"], temp)},
                   coords={"lon": lon, "lat": lat, "time": time})


# ----------------------------------------------------------------------------
# transcriptions of the reference's hot-path tests (tests/test_climate_toolbox.py:109-135)
# ----------------------------------------------------------------------------
def test_reindex_spatial_weights(clim_data, ref_fix):
    weights = ref_fix[4]
    assert not clim_data.temperature.isnull().any()
    ds = _reindex_spatial_data_to_regions(clim_data, weights)
    assert ds.temperature.shape == (len(ds["lon"]), len(ds["time"]))
    assert "reshape_index" in ds.dims
    G = np.load(GOLD, allow_pickle=True)
    np.testing.assert_array_equal(ds.temperature.values, G["reindexed"])  # gather is bit-exact


def test_weighting(clim_data, ref_fix):
    weights = ref_fix[4]
    assert np.isnan(weights["popwt"].values).any()
    ds = _reindex_spatial_data_to_regions(clim_data, weights)
    assert not ds.temperature.isnull().any()
    G = np.load(GOLD, allow_pickle=True)

    wtd = _aggregate_reindexed_data_to_regions(ds, "temperature", "popwt", "ISO", weights)
    assert not wtd.temperature.isnull().any()
    assert wtd.temperature.dims == ("ISO", "time")
    check(wtd.temperature.values, G["agg_popwt_ISO"])

    # the reference re-uses the same reindexed ds with another weight
    wtd = _aggregate_reindexed_data_to_regions(ds, "temperature", "areawt", "ISO", weights)
    assert not wtd.temperature.isnull().any()
    check(wtd.temperature.values, G["agg_areawt_ISO"])


def test_weighting_on_materialised_reindexed_data(clim_data, ref_fix):
    """A caller that built the (reshape_index, time) array itself."""
    weights = ref_fix[4]
    G = np.load(GOLD, allow_pickle=True)
    for dims, arr in ((("reshape_index", "time"), G["reindexed"]),
                      (("time", "reshape_index"), np.ascontiguousarray(G["reindexed"].T))):
        ds = Dataset({"temperature": (dims, arr)}, coords={"time": ref_fix[2]})
        wtd = _aggregate_reindexed_data_to_regions(ds, "temperature", "popwt", "hierid", weights)
        exp = G["agg_popwt_hierid"]
        assert wtd.temperature.dims == tuple("hierid" if d == "reshape_index" else d for d in dims)
        check(wtd.temperature.values, exp if dims[0] == "reshape_index" else exp.T)


@pytest.mark.parametrize("aggwt", ["popwt", "areawt"])
@pytest.mark.parametrize("agglev", ["ISO", "hierid"])
@pytest.mark.parametrize("layout", ["lat_lon_time", "time_lat_lon"])
def test_weighted_aggregate_on_reference_fixture(ref_fix, aggwt, agglev, layout):
    lat, lon, time, temp, weights = ref_fix
    G = np.load(GOLD, allow_pickle=True)
    if layout == "lat_lon_time":
        ds = Dataset({"temperature": (["lat", "lon", "time"], temp)},
                     coords={"lon": lon, "lat": lat, "time": time})
        exp, dims = G["agg_{}_{}".format(aggwt, agglev)], (agglev, "time")
    else:
        ds = Dataset({"temperature": (["time", "lat", "lon"], np.ascontiguousarray(temp.transpose(2, 0, 1)))},
                     coords={"lon": lon, "lat": lat, "time": time})
        exp, dims = G["agg_{}_{}".format(aggwt, agglev)].T, ("time", agglev)
    out = weighted_aggregate_grid_to_regions(ds, "temperature", aggwt, agglev, weights=weights)
    assert out.temperature.dims == dims
    np.testing.assert_array_equal(out.coords[agglev].values, G["labels_" + agglev])
    np.testing.assert_array_equal(out.coords["time"].values, time.values)
    assert out.temperature.values.dtype == np.float64
    check(out.temperature.values, exp)
    # the caller's dataset is not mutated (the reference adds agglev/aggwt to it)
    assert agglev not in ds.coords and aggwt not in ds.data_vars


def test_index_map_bit_exact_with_lon_roll_and_takes(ref_fix):
    """K0: row -> physical gridcell, exact float64 label match, lon roll folded in."""
    d = 1.0
    lat, lon360 = synthetic.grid_labels(d, lon_0_360=True)
    df = synthetic.weights_table(d, 500, seed=3)
    lon_std, perm = oracle.convert_lons_split(lon360)
    i = oracle.exact_label_index(lat, df["lat"].values)
    j = oracle.exact_label_index(lon_std, df["lon"].values)
    expect = (i * len(lon360) + perm[j]).astype(np.int32)
    grid = E.GridSpec(lat, lon_std, None, perm, len(lat), len(lon360))
    plan = E.get_plan(grid, df, "popwt", "hierid", cache=False)
    np.testing.assert_array_equal(plan.row_cells(), expect)
    w = plan.row_weights()
    np.testing.assert_array_equal(w, oracle.effective_weights(df, "popwt"))  # bit-exact incl. NaN
    codes, labels = pd.factorize(df["hierid"].values, sort=True)
    wz = np.nan_to_num(w, nan=0.0)
    den = np.array([wz[codes == r].sum() for r in range(len(labels))])
    np.testing.assert_allclose(plan.den(), den, rtol=1e-14)
    info = plan.info
    assert info["n_regions"] == len(labels) and info["n_rows"] == len(df)
    kept = ~np.isnan(w) & (w != 0)
    assert info["nnz"] == kept.sum()
    assert info["n_cells_distinct"] == len(np.unique(expect[kept]))
    plan.close()


def test_label_miss_raises_keyerror(clim_data, ref_fix):
    weights = ref_fix[4].copy()
    weights.loc[7, "lat"] = weights.loc[7, "lat"] + 1e-10
    with pytest.raises(KeyError, match="lat"):
        weighted_aggregate_grid_to_regions(clim_data, "temperature", "popwt", "ISO", weights=weights)
    with pytest.raises(KeyError):
        weighted_aggregate_grid_to_regions(clim_data, "nope", "popwt", "ISO", weights=ref_fix[4])
    with pytest.raises(KeyError):
        weighted_aggregate_grid_to_regions(clim_data, "temperature", "nowt", "ISO", weights=ref_fix[4])


# ----------------------------------------------------------------------------
# synthetic configs at oracle-friendly sizes
# ----------------------------------------------------------------------------
_TABLES = {}


def _table(d, R, seed=1234):
    """Synthetic weights tables are deterministic and slow to build at 0.25 degree: one per session."""
    key = (d, R, seed)
    if key not in _TABLES:
        _TABLES[key] = synthetic.weights_table(d, R, seed=seed)
    return _TABLES[key]


def _config(d, R, T, seed=1234, nan_frac=0.001, lon_0_360=False, dtype=np.float32):
    lat, lon = synthetic.grid_labels(d, lon_0_360=lon_0_360)
    df = _table(d, R, seed)
    tas, tmin, tmax = synthetic.tas_field(T, len(lat), len(lon), seed=7, nan_frac=nan_frac, dtype=dtype)
    return lat, lon, df, tas, tmin, tmax


@pytest.mark.parametrize("where", ["host", "device"])
@pytest.mark.parametrize("variant", [N.VARIANT_STAGED, N.VARIANT_DIRECT])
def test_config1_shape_areawt(where, variant):
    lat, lon, df, tas, _, _ = _config(1.0, 3000, 70)
    data = tas if where == "host" else torch.from_numpy(tas).cuda()
    ds = Dataset({"tas": (("time", "lat", "lon"), data)},
                 coords={"time": pd.date_range("2001-01-01", periods=70), "lat": lat, "lon": lon})
    out = weighted_aggregate_grid_to_regions(ds, "tas", "areawt", "hierid", weights=df, variant=variant)
    ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, df, "areawt", "hierid")
    assert out.tas.dims == rd and list(out.hierid.values) == list(labels)
    check(out.tas.values, ref, scale)


@pytest.mark.parametrize("mode", ["packed_pageable", "packed_pinned_leap", "pulled_pinned_leap", "pulled_pinned",
                                  "zero_copy_pinned", "chunked_pinned", "chunked_pageable"])
def test_host_input_paths(mode):
    """Host arrays.  Default: a compact plan -- the referenced gridcells are packed on the host
    (ctb_host_pack) and only they cross PCIe.  Alternatives: whole time chunks copied
    double-buffered against the kernel, or pinned memory read in place by the kernel.
    70 days with a 1 MB chunk budget forces several chunks and a ragged last one."""
    lat, lon, df, tas, _, _ = _config(1.0, 3000, 70)
    if "pinned" in mode:
        host = torch.empty(tas.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(torch.from_numpy(tas))
        arr = host.numpy()
    else:
        arr = tas
    tix = None
    exp_in = tas
    if mode.endswith("leap"):
        tix = np.delete(np.arange(70), [3, 40])        # a time take folded into the packing
        exp_in = tas[tix]
    T = exp_in.shape[0]
    grid = E.GridSpec(lat, lon)
    compact = mode.startswith("packed") or mode.startswith("pulled")
    plan = E.get_plan(grid, df, "areawt", "hierid", compact=compact)
    if compact:
        # referenced pieces + padding of every run to a 64-byte boundary of the packed plane
        assert 4 * plan.info["n_pieces_distinct"] <= plan.info["n_packed_cells"] < len(lat) * len(lon)
        assert plan.info["n_packed_cells"] % 16 == 0
    n0 = E.launch_count()
    out = E.aggregate_host(plan, [arr.reshape(70, -1)], N.LAYOUT_TIME_MAJOR, arr.shape[1] * arr.shape[2],
                           tix, T, chunk_bytes=1 << 20, zero_copy=(mode == "zero_copy_pinned"),
                           ingest="pull" if mode.startswith("pulled") else "pack")
    assert (E.launch_count() - n0 == 1) == (mode == "zero_copy_pinned")
    ref, rd, labels, scale = oracle_agg(exp_in, ("time", "lat", "lon"), lat, lon, df, "areawt", "hierid")
    check(out[0].cpu().numpy().T, ref, scale)


@pytest.mark.parametrize("aggwt,agglev", [("popwt", "hierid"), ("cropwt", "hierid"), ("popwt", "ISO")])
def test_quarter_degree_sample_vs_oracle(aggwt, agglev):
    """Full 0.25-degree grid and 24,378 regions, a few days (oracle finishes in seconds).
    agglev=ISO makes 180 huge regions -> exercises region splitting + the fix-up kernel."""
    lat, lon, df, tas, _, _ = _config(0.25, 24378, 5)
    ds = Dataset({"tas": (("time", "lat", "lon"), torch.from_numpy(tas).cuda())},
                 coords={"time": np.arange(5), "lat": lat, "lon": lon})
    ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, df, aggwt, agglev)
    for variant in (N.VARIANT_STAGED, N.VARIANT_DIRECT):
        out = weighted_aggregate_grid_to_regions(ds, "tas", aggwt, agglev, weights=df, variant=variant)
        check(out.tas.values, ref, scale)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_infinite_and_nan_values_take_the_checked_reduction(dtype):
    """The fused kernel reduces tiles of finite values without per-value checks (and with the
    entry ranges padded by zero-weight entries); a tile that staged a NaN or an infinity must
    take the checked path: +inf / -inf propagate, inf - inf = NaN, NaN products are skipped."""
    lat, lon, df, tas, _, _ = _config(1.0, 800, 40, nan_frac=0.0, dtype=dtype)
    ii, jj = np.searchsorted(lat, df.lat.values), np.searchsorted(lon, df.lon.values)
    big = df.groupby("hierid").size().sort_values().index[-3:]        # three multi-cell regions
    rows = [np.flatnonzero(df.hierid.values == r) for r in big]
    tas[3, ii[rows[0][0]], jj[rows[0][0]]] = np.inf                     # +inf alone
    tas[7, ii[rows[1][0]], jj[rows[1][0]]] = np.inf                     # +inf and -inf in one region-day
    k = next(q for q in rows[1] if (ii[q], jj[q]) != (ii[rows[1][0]], jj[rows[1][0]]))
    tas[7, ii[k], jj[k]] = -np.inf
    tas[33, ii[rows[2][0]], jj[rows[2][0]]] = np.nan                    # NaN next to -inf
    tas[33, ii[rows[2][-1]], jj[rows[2][-1]]] = -np.inf
    ds = Dataset({"tas": (("time", "lat", "lon"), torch.from_numpy(tas).cuda())},
                 coords={"time": np.arange(40), "lat": lat, "lon": lon})
    ref = oracle.weighted_aggregate_grid_to_regions(tas, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")[0]
    assert np.isinf(ref).any() and np.isnan(ref).any()
    for variant in (N.VARIANT_STAGED, N.VARIANT_DIRECT):
        out = weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df, variant=variant)
        check(out.tas.values, ref, tol=1e-9)


@pytest.mark.parametrize("where", ["host", "device"])
def test_several_weight_columns_in_one_pass(where):
    """popwt + areawt + cropwt from one pass over the data == three reference passes."""
    lat, lon, df, tas, _, _ = _config(1.0, 1500, 45)
    df = df.copy()
    df.loc[df.index[::97], "hierid"] = np.nan            # rows without a region are dropped
    data = tas if where == "host" else torch.from_numpy(tas).cuda()
    ds = Dataset({"tas": (("time", "lat", "lon"), data)},
                 coords={"time": np.arange(45), "lat": lat, "lon": lon})
    cols = ["popwt", "areawt", "cropwt"]
    n0 = E.launch_count()
    E.get_plan  # noqa: B018
    out = weighted_aggregate_grid_to_regions_multi(ds, "tas", cols, "hierid", df)
    for c in cols:
        ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, df, c, "hierid")
        got = out["tas_" + c]
        assert got.dims == rd and list(out.hierid.values) == list(labels)
        check(got.values if where == "host" else got.values, ref, scale)
    # and it is the same as three single calls
    one = weighted_aggregate_grid_to_regions(ds, "tas", "cropwt", "hierid", weights=df)
    np.testing.assert_allclose(out["tas_cropwt"].values, one.tas.values, rtol=1e-12, equal_nan=True)


def test_one_plan_on_two_streams_at_once():
    """Launches of one plan on different streams take different scheduler counters."""
    lat, lon, df, tas, _, _ = _config(1.0, 3000, 365, nan_frac=0.0)
    x = torch.from_numpy(tas).cuda().view(365, -1)
    plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid")
    ref = E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], None, 365, "identity", (), 1,
                             N.VARIANT_STAGED)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[torch.empty_like(ref) for _ in range(20)] for _ in streams]
    wss = [torch.empty(max(1, N.lib().ctb_aggregate_workspace_bytes(plan._h, 365, 1) // 8), dtype=torch.float64,
                       device="cuda") for _ in streams]
    for i in range(20):
        for s, o, ws in zip(streams, outs, wss):
            with torch.cuda.stream(s):
                E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], None, 365, "identity", (), 1,
                                   N.VARIANT_STAGED, out=o[i], workspace=ws)
    torch.cuda.synchronize()
    for o in outs:
        for t in o:
            assert torch.equal(t, ref)


def test_results_own_their_pinned_blocks():
    """Host results come back in pooled pinned blocks: a block may be reused only after every view
    of the result it carried is gone."""
    import gc
    lat, lon, df, tas, _, _ = _config(1.0, 500, 40, nan_frac=0.0, dtype=np.float64)   # x + c is exact enough in f64
    mk = lambda a: Dataset({"tas": (("time", "lat", "lon"), a)},
                           coords={"time": np.arange(40), "lat": lat, "lon": lon})
    r1 = weighted_aggregate_grid_to_regions(mk(tas), "tas", "areawt", "hierid", weights=df)
    keep = r1.tas.values[3:, ::2]                    # a view of a view
    snap = keep.copy()
    del r1
    gc.collect()
    r2 = weighted_aggregate_grid_to_regions(mk(tas + 5.0), "tas", "areawt", "hierid", weights=df)
    np.testing.assert_array_equal(keep, snap)        # not overwritten by the second call
    np.testing.assert_allclose(r2.tas.values[3:, ::2], snap + 5.0, rtol=1e-10)
    n_free = sum(len(v) for v in E._RESULT_POOL.values())
    del keep
    gc.collect()
    assert sum(len(v) for v in E._RESULT_POOL.values()) == n_free + 1   # the block went back to the pool
    r3 = weighted_aggregate_grid_to_regions(mk(tas - 1.0), "tas", "areawt", "hierid", weights=df)
    np.testing.assert_allclose(r2.tas.values[3:, ::2], snap + 5.0, rtol=1e-10)
    np.testing.assert_allclose(r3.tas.values[3:, ::2], snap - 1.0, rtol=1e-10)


def test_lon_0_360_and_leap_day_folded_into_the_kernel():
    lat, lon360, df, tas, _, _ = _config(1.0, 800, 45, lon_0_360=True)
    time = pd.date_range("2000-02-01", periods=45)
    ds = Dataset({"tas": (("time", "lat", "lon"), torch.from_numpy(tas).cuda())},
                 coords={"time": time, "lat": lat, "lon": lon360})
    ds = load_bcsd(ds, "tas")                 # rename + lazy lon roll
    out = weighted_aggregate_grid_to_regions(tas_poly(ds, 2, "tas2"), "tas2", "popwt", "hierid", weights=df)
    lon_std, perm = oracle.convert_lons_split(lon360)
    xt, tl = oracle.tas_poly(np.take(tas, perm, axis=2), time, 2)
    ref, rd, labels, scale = oracle_agg(xt, ("time", "lat", "lon"), lat, lon_std, df, "popwt", "hierid")
    assert out.tas2.shape == (44, len(labels)) and out.tas2.dims == rd
    np.testing.assert_array_equal(out.time.values, tl)
    check(out.tas2.values, ref, scale)


def test_fused_tas_poly_orders_1_to_4():
    """Config 3: four polynomial orders from one read of tas."""
    lat, lon, df, tas, _, _ = _config(1.0, 3000, 40)
    time = pd.date_range("2001-03-01", periods=40)
    ds = Dataset({"tas": (("time", "lat", "lon"), torch.from_numpy(tas).cuda())},
                 coords={"time": time, "lat": lat, "lon": lon})
    names = ["tas", "tas-poly-2", "tas-poly-3", "tas-poly-4"]
    ds1 = tas_poly(ds, [1, 2, 3, 4], names)
    weighted_aggregate_grid_to_regions(ds1, names, "popwt", "hierid", weights=df)   # builds the plan
    n0 = E.launch_count()
    out = weighted_aggregate_grid_to_regions(ds1, names, "popwt", "hierid", weights=df)
    assert E.launch_count() - n0 == 1          # one fused launch for all four orders
    for p, name in zip((1, 2, 3, 4), names):
        xt, _ = oracle.tas_poly(tas, time, p)
        ref, rd, labels, scale = oracle_agg(xt, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
        check(out[name].values, ref, scale)
    # a single order through the one-output instantiation agrees too
    one = weighted_aggregate_grid_to_regions(tas_poly(ds, 3, "p3"), "p3", "popwt", "hierid", weights=df)
    check(one.p3.values, out["tas-poly-3"].values, tol=1e-14)


@pytest.mark.parametrize("layout", ["time_lat_lon", "lat_lon_time"])
def test_fused_snyder_edd_gdd(layout):
    """Config 4: EDD at two thresholds and GDD from (tasmin, tasmax), cropwt."""
    lat, lon, df, _, tmin, tmax = _config(1.0, 3000, 33)
    dims = ("time", "lat", "lon")
    a, b = tmin, tmax
    if layout == "lat_lon_time":
        dims = ("lat", "lon", "time")
        a, b = (np.ascontiguousarray(v.transpose(1, 2, 0)) for v in (tmin, tmax))
    coords = {"time": np.arange(33), "lat": lat, "lon": lon}
    tn = DataArray(torch.from_numpy(a).cuda(), dims=dims, coords=coords, attrs={"units": "K"})
    tx = DataArray(torch.from_numpy(b).cuda(), dims=dims, coords=coords, attrs={"units": "K"})
    ds = Dataset(coords=coords)
    ds["edd10"] = snyder_edd(tn, tx, 283.15)
    ds["edd30"] = snyder_edd(tn, tx, 303.15)
    ds["gdd"] = snyder_gdd(tn, tx, 283.15, 303.15)
    assert ds["gdd"].units == "degreedays_283.15-303.15K"
    out = weighted_aggregate_grid_to_regions(ds, ["edd10", "edd30", "gdd"], "cropwt", "hierid", weights=df)
    for name, f in (("edd10", oracle.snyder_edd(a, b, 283.15)), ("edd30", oracle.snyder_edd(a, b, 303.15)),
                    ("gdd", oracle.snyder_gdd(a, b, 283.15, 303.15))):
        ref, rd, labels, scale = oracle_agg(f, dims, lat, lon, df, "cropwt", "hierid")
        assert out[name].dims == rd
        check(out[name].values, ref, scale)


def test_reference_snyder_known_answers():
    # reference tests :231-278 through the pointwise kernel (.values)
    tmax = DataArray(np.array([280.4963, 280.7887]), dims=("(latitude, longitude)",), attrs={"units": "K"})
    tmin = DataArray(np.array([278.902, 278.23163]), dims=("(latitude, longitude)",), attrs={"units": "K"})
    res = snyder_edd(tmin, tmax, threshold=273.15 + 8)
    assert res.units == "degreedays_281.15K"
    assert res.sum().item(0) == 0.0
    res = snyder_gdd(tmin, tmax, threshold_low=273.15 + 1, threshold_high=273.15 + 8)
    assert not res.units == "degreedays_281.15K"
    assert res.units == "degreedays_274.15-281.15K"
    assert res.sum().item(0) == pytest.approx(11, 0.1)
    G = np.load(GOLD, allow_pickle=True)
    check(res.values, G["gdd_274.15_281.15"], tol=1e-12)


def test_pointwise_transform_matches_oracle():
    _, _, _, tas, tmin, tmax = _config(2.0, 50, 6, nan_frac=0.01, dtype=np.float64)
    dims = ("time", "lat", "lon")
    tn = DataArray(tmin, dims=dims, attrs={"units": "K"})
    tx = DataArray(tmax, dims=dims, attrs={"units": "K"})
    # near tmax ~ e the closed form cancels to ~0 (in the oracle too): errors are measured
    # against the half-range W, the magnitude of the terms that cancel
    W = np.nan_to_num(np.abs(tmax - tmin) / 2) + 1.0
    check(snyder_edd(tn, tx, 290.0).values, oracle.snyder_edd(tmin, tmax, 290.0), scale=W, tol=1e-12)
    check(snyder_gdd(tn, tx, 285.0, 295.0).values, oracle.snyder_gdd(tmin, tmax, 285.0, 295.0),
          scale=W + np.nan_to_num(np.abs(oracle.snyder_edd(tmin, tmax, 285.0))), tol=1e-12)
    t = pd.date_range("2001-01-01", periods=6)
    ds = Dataset({"tas": (dims, tas)}, coords={"time": t})
    check(tas_poly(ds, 4, "p4").p4.values, oracle.tas_poly(tas, t, 4)[0], tol=1e-12)


# ----------------------------------------------------------------------------
# edge cases
# ----------------------------------------------------------------------------
def _tiny(nlat=7, nlon=9, T=5, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    lat = np.arange(nlat) * 0.5 - 1.0
    lon = np.arange(nlon) * 0.5 + 10.0
    x = rng.standard_normal((T, nlat, nlon)).astype(dtype) * 10
    return lat, lon, x, rng


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("variant", [N.VARIANT_STAGED, N.VARIANT_DIRECT])
def test_odd_grid_scalar_path_and_degenerate_regions(dtype, variant):
    """63 cells (not a multiple of 4 -> unaligned planes, scalar staging loads); single-row
    regions, a region whose weights are all NaN (den = 0 -> NaN), one whose data are all NaN
    (-> 0), NaN region labels (dropped), duplicate rows, negative backup weights."""
    lat, lon, x, rng = _tiny(dtype=dtype)
    x[:, 2, 3] = np.nan
    x[1, 4, 4] = np.inf
    rows = []
    for k in range(60):
        rows.append((lat[rng.integers(7)], lon[rng.integers(9)], "r%02d" % rng.integers(12),
                     rng.random(), rng.random()))
    rows += [(lat[0], lon[0], "single", 2.0, 1.0), (lat[1], lon[1], "allnanw", np.nan, np.nan),
             (lat[1], lon[2], "allnanw", 0.0, np.nan), (lat[2], lon[3], "allnandata", 1.0, 1.0),
             (lat[3], lon[3], np.nan, 1.0, 1.0), (lat[5], lon[5], "dup", 1.5, 1.0),
             (lat[5], lon[5], "dup", 1.5, 1.0), (lat[6], lon[8], "neg", -1.0, -2.0),
             (lat[6], lon[7], "neg", 3.0, 1.0), (lat[4], lon[4], "inf", 1.0, 1.0),
             (lat[0], lon[1], "zerow", 0.0, 0.0)]
    df = pd.DataFrame(rows, columns=["lat", "lon", "hierid", "popwt", "areawt"])
    ds = Dataset({"v": (("time", "lat", "lon"), x)}, coords={"time": np.arange(5), "lat": lat, "lon": lon})
    out = weighted_aggregate_grid_to_regions(ds, "v", "popwt", "hierid", weights=df, variant=variant)
    ref, rd, labels, scale = oracle_agg(x, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
    assert list(out.hierid.values) == list(labels)
    got = out.v.values
    li = list(labels)
    assert np.isnan(got[:, li.index("allnanw")]).all() and np.isnan(got[:, li.index("zerow")]).all()
    assert (got[:, li.index("allnandata")] == 0).all()
    assert got[1, li.index("inf")] == np.inf
    check(got, ref, scale)


def test_empty_inputs():
    lat, lon, x, rng = _tiny()
    df = pd.DataFrame({"lat": [lat[0]], "lon": [lon[0]], "hierid": ["a"], "popwt": [1.0], "areawt": [1.0]})
    ds = Dataset({"v": (("time", "lat", "lon"), x[:0])}, coords={"time": np.arange(0), "lat": lat, "lon": lon})
    out = weighted_aggregate_grid_to_regions(ds, "v", "popwt", "hierid", weights=df)
    assert out.v.shape == (0, 1)
    ds = Dataset({"v": (("time", "lat", "lon"), x)}, coords={"time": np.arange(5), "lat": lat, "lon": lon})
    out = weighted_aggregate_grid_to_regions(ds, "v", "popwt", "hierid", weights=df.iloc[:0])
    assert out.v.shape == (5, 0)


def test_ragged_time_and_2d_field():
    """T not a multiple of the 32-day tile; and a (lat, lon) field without a time axis."""
    lat, lon, df, tas, _, _ = _config(2.0, 300, 37)
    ds = Dataset({"tas": (("time", "lat", "lon"), tas)}, coords={"time": np.arange(37), "lat": lat, "lon": lon})
    out = weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df)
    ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
    check(out.tas.values, ref, scale)
    ds2 = Dataset({"tas": (("lat", "lon"), tas[0])}, coords={"lat": lat, "lon": lon})
    out2 = weighted_aggregate_grid_to_regions(ds2, "tas", "popwt", "hierid", weights=df)
    assert out2.tas.dims == ("hierid",)
    check(out2.tas.values, ref[0], scale[0])


def test_split_regions_small_tile():
    """Force regions to be larger than one staging tile (tiny smem budget): the two-phase
    partial-sum path must agree with the direct kernel and the oracle."""
    lat, lon, df, tas, _, _ = _config(1.0, 40, 35)
    ds = Dataset({"tas": (("time", "lat", "lon"), torch.from_numpy(tas).cuda())},
                 coords={"time": np.arange(35), "lat": lat, "lon": lon})
    ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
    out = weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df,
                                             variant=N.VARIANT_STAGED, smem_budget=8 * 1024)
    check(out.tas.values, ref, scale)
    lat_, lon_ = synthetic.grid_labels(1.0)
    plan = E.get_plan(E.GridSpec(lat_, lon_), df, "popwt", "hierid", smem_budget=8 * 1024)
    assert plan.info["n_split_regions"] > 0 and plan.info["n_scratch_slots"] > plan.info["n_split_regions"]


def test_extra_leading_dims_and_permuted_layout():
    """(model, time, lat, lon) flattens onto the time axis; (lon, time, lat) is permuted."""
    lat, lon, df, tas, _, _ = _config(2.0, 200, 12)
    x4 = tas.reshape(3, 4, len(lat), len(lon))
    ds = Dataset({"tas": (("model", "time", "lat", "lon"), x4)},
                 coords={"model": ["a", "b", "c"], "time": np.arange(4), "lat": lat, "lon": lon})
    out = weighted_aggregate_grid_to_regions(ds, "tas", "areawt", "hierid", weights=df)
    ref, rd, labels, scale = oracle_agg(x4, ("model", "time", "lat", "lon"), lat, lon, df, "areawt", "hierid")
    assert out.tas.dims == rd == ("model", "time", "hierid")
    check(out.tas.values, ref, scale)
    xp = np.ascontiguousarray(tas.transpose(2, 0, 1))
    ds = Dataset({"tas": (("lon", "time", "lat"), xp)}, coords={"time": np.arange(12), "lat": lat, "lon": lon})
    out = weighted_aggregate_grid_to_regions(ds, "tas", "areawt", "hierid", weights=df)
    ref, rd, labels, scale = oracle_agg(xp, ("lon", "time", "lat"), lat, lon, df, "areawt", "hierid")
    assert out.tas.dims == rd
    check(out.tas.values, ref, scale)


def test_prepare_spatial_weights_data_csv(tmp_path):
    lat, lon, df, tas, _, _ = _config(2.0, 100, 3)
    csv = df.rename(columns={"lon": "pix_cent_x", "lat": "pix_cent_y"}).reset_index(drop=True)
    # the reference's out-of-bounds relabel: 180.125 -> -179.875 (aggregations.py:144)
    p = tmp_path / "weights.csv"
    csv.to_csv(p, index=False)
    ds = Dataset({"tas": (("time", "lat", "lon"), tas)}, coords={"time": np.arange(3), "lat": lat, "lon": lon})
    out = weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=str(p))
    ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon,
                                        oracle.prepare_spatial_weights_data(str(p)), "popwt", "hierid")
    check(out.tas.values, ref, scale)


def test_prepare_spatial_weights_data_csv_relabels_180_125(tmp_path):
    """aggregations.py:144: pix_cent_x 180.125 is the gridcell at -179.875 -- through the CSV path of the
    drop-in (the grid must hold -179.875, so a small 0.25-degree strip)."""
    rng = np.random.default_rng(11)
    lat = np.array([10.125, 10.375, 10.625, 10.875])
    lon = -179.875 + 0.25 * np.arange(16)
    tas = (280 + 10 * rng.standard_normal((5, len(lat), len(lon)))).astype(np.float32)
    la, lo = np.meshgrid(lat, lon, indexing="ij")
    csv = pd.DataFrame({"pix_cent_x": lo.ravel(), "pix_cent_y": la.ravel(),
                        "hierid": ["R%d" % (i % 5) for i in range(la.size)],
                        "popwt": rng.lognormal(0, 1, la.size), "areawt": rng.random(la.size) + 0.1})
    west = csv["pix_cent_x"] == -179.875
    csv.loc[west, "pix_cent_x"] = 180.125          # as the segment-weights files write the wrapped column
    assert west.sum() == len(lat)
    p = tmp_path / "weights.csv"
    csv.to_csv(p, index=False)
    ds = Dataset({"tas": (("time", "lat", "lon"), tas)}, coords={"time": np.arange(5), "lat": lat, "lon": lon})
    out = weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=str(p))
    w = oracle.prepare_spatial_weights_data(str(p))
    assert (w["lon"] == -179.875).sum() == len(lat) and not (w["lon"] == 180.125).any()
    ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, w, "popwt", "hierid")
    check(out.tas.values, ref, scale)
    # ... and the relabelled column takes part: dropping it changes the answer
    ref2, _, _, _ = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, w[w["lon"] != -179.875], "popwt", "hierid")
    assert np.nanmax(np.abs(ref - ref2)) > 1e-3


# ----------------------------------------------------------------------------
# BASELINE.json configs 3, 4, 5 and the multi-weight pass AT THEIR OWN SHAPE
# (0.25 degree, 24,378 regions, 866+ bundles), a few days against the oracle
# ----------------------------------------------------------------------------
def _put(a, where):
    return a if where == "host" else torch.from_numpy(a).cuda()


@pytest.mark.parametrize("where", ["host", "device"])
def test_config3_shape_poly_orders_quarter_degree(where):
    """Config 3: tas_poly orders 1-4 fused, 0.25 degree -> 24,378 regions, popwt."""
    lat, lon, df, tas, _, _ = _config(0.25, 24378, 4)
    time = pd.date_range("2001-03-01", periods=4)
    ds = Dataset({"tas": (("time", "lat", "lon"), _put(tas, where))},
                 coords={"time": time, "lat": lat, "lon": lon})
    names = ["tas", "tas-poly-2", "tas-poly-3", "tas-poly-4"]
    out = weighted_aggregate_grid_to_regions(tas_poly(ds, [1, 2, 3, 4], names), names, "popwt", "hierid", weights=df)
    for p, name in zip((1, 2, 3, 4), names):
        xt, _ = oracle.tas_poly(tas, time, p)
        ref, rd, labels, scale = oracle_agg(xt, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
        assert out[name].dims == rd and out[name].shape == (4, 24378)
        check(out[name].values, ref, scale)


@pytest.mark.parametrize("where", ["host", "device"])
def test_config4_shape_snyder_cropwt_quarter_degree(where):
    """Config 4: Snyder EDD at two thresholds + GDD from (tasmin, tasmax), cropwt (60 % of the
    rows fall back to areawt), 0.25 degree -> 24,378 regions."""
    lat, lon, df, _, tmin, tmax = _config(0.25, 24378, 3)
    dims = ("time", "lat", "lon")
    coords = {"time": np.arange(3), "lat": lat, "lon": lon}
    tn = DataArray(_put(tmin, where), dims=dims, coords=coords, attrs={"units": "K"})
    tx = DataArray(_put(tmax, where), dims=dims, coords=coords, attrs={"units": "K"})
    ds = Dataset(coords=coords)
    ds["edd10"] = snyder_edd(tn, tx, 283.15)
    ds["edd30"] = snyder_edd(tn, tx, 303.15)
    ds["gdd"] = snyder_gdd(tn, tx, 283.15, 303.15)
    out = weighted_aggregate_grid_to_regions(ds, ["edd10", "edd30", "gdd"], "cropwt", "hierid", weights=df)
    W = np.nan_to_num((tmax.astype(np.float64) - tmin) / 2)      # magnitude of the terms that cancel
    for name, f in (("edd10", oracle.snyder_edd(tmin, tmax, 283.15)), ("edd30", oracle.snyder_edd(tmin, tmax, 303.15)),
                    ("gdd", oracle.snyder_gdd(tmin, tmax, 283.15, 303.15))):
        ref, rd, labels, scale = oracle_agg(f, dims, lat, lon, df, "cropwt", "hierid")
        scale_w = oracle.weighted_aggregate_grid_to_regions(W, dims, lat, lon, df, "cropwt", "hierid")[0]
        assert out[name].shape == (3, 24378)
        check(out[name].values, ref, scale + scale_w)


def test_several_weight_columns_quarter_degree():
    """popwt + areawt + cropwt from ONE pass at 0.25 degree == three oracle passes."""
    lat, lon, df, tas, _, _ = _config(0.25, 24378, 3)
    ds = Dataset({"tas": (("time", "lat", "lon"), torch.from_numpy(tas).cuda())},
                 coords={"time": np.arange(3), "lat": lat, "lon": lon})
    cols = ["popwt", "areawt", "cropwt"]
    out = weighted_aggregate_grid_to_regions_multi(ds, "tas", cols, "hierid", df)
    for c in cols:
        ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, df, c, "hierid")
        assert out["tas_" + c].dims == rd and list(out.hierid.values) == list(labels)
        check(out["tas_" + c].values, ref, scale)


def test_several_region_levels_and_weights_in_one_pass():
    """hierid + ISO x popwt + areawt from ONE pass over the data == four oracle passes;
    the 180 ISO regions are far larger than a tile (split + fix-up), at 0.25 degree."""
    lat, lon, df, tas, _, _ = _config(0.25, 24378, 3)
    ds = Dataset({"tas": (("time", "lat", "lon"), torch.from_numpy(tas).cuda())},
                 coords={"time": np.arange(3), "lat": lat, "lon": lon})
    n0 = E.launch_count()
    out = weighted_aggregate_grid_to_regions_multi(ds, "tas", ["popwt", "areawt"], ["hierid", "ISO"], df)
    for lev in ("hierid", "ISO"):
        for c in ("popwt", "areawt"):
            ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, df, c, lev)
            got = out["tas_{}_{}".format(c, lev)]
            assert got.dims == rd and list(out[lev].values) == list(labels)
            check(got.values, ref, scale)
    # one streaming launch (+ the fix-up of the split regions), not four
    out = weighted_aggregate_grid_to_regions_multi(ds, "tas", ["popwt", "areawt"], ["hierid", "ISO"], df)
    n1 = E.launch_count()
    out = weighted_aggregate_grid_to_regions_multi(ds, "tas", ["popwt", "areawt"], ["hierid", "ISO"], df)
    assert E.launch_count() - n1 == 2


def test_config5_slice_model_years_through_a_buffer_pool():
    """Config 5 (ensemble streaming): model-years from a pool of resident buffers through ONE plan and
    one reused output buffer, as bench.py --workload config5 does -- every model-year's output is
    compared (3 days of each against the oracle, all of it against the direct kernel)."""
    T = 64
    lat, lon, df, _, _, _ = _config(0.25, 24378, 1)
    plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid")
    g = torch.Generator(device="cuda").manual_seed(11)
    pool = [288.0 + 10.0 * torch.randn((T, len(lat) * len(lon)), generator=g, device="cuda", dtype=torch.float32)
            for _ in range(2)]
    out = torch.empty((1, plan.R, T), dtype=torch.float64, device="cuda")
    ok = torch.from_numpy(plan.den() > 0).cuda()
    for y in range(4):                                   # 4 model-years through 2 buffers
        x = pool[y % 2]
        if y >= 2:
            x.add_(1.0)                                  # the "next model-year" lands in the same buffer
        E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], None, T, out=out)
        direct = E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, x.shape[1], None, T,
                                    variant=N.VARIANT_DIRECT)
        assert torch.allclose(out[0][ok], direct[0][ok], rtol=1e-12, atol=0)
        days = [0, 31, 63]
        sl = x[days].cpu().numpy().reshape(3, len(lat), len(lon))
        ref, rd, labels, scale = oracle_agg(sl, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
        check(out[0][:, days].cpu().numpy().T, ref, scale)


# ----------------------------------------------------------------------------
# plan cache: in-place edits of the weights frame must rebuild the plan
# ----------------------------------------------------------------------------
def test_plan_cache_sees_in_place_edits_of_the_weights_frame():
    lat, lon, df, tas, _, _ = _config(1.0, 800, 6)
    df = df.copy()
    ds = Dataset({"tas": (("time", "lat", "lon"), torch.from_numpy(tas).cuda())},
                 coords={"time": np.arange(6), "lat": lat, "lon": lon})

    def both():
        out = weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df)
        ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
        assert list(out.hierid.values) == list(labels)
        check(out.tas.values, ref, scale)
        return out.tas.values

    a = both()
    # (1) region labels edited in the MIDDLE of the frame, in place (first and last row untouched)
    mid = df.index[len(df) // 3: len(df) // 3 + 200]
    df.loc[mid, "hierid"] = df["hierid"].values[0]
    b = both()
    assert not np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
    # (2) one weight column reversed in place: same multiset of values, other positions
    df["popwt"] = df["popwt"].values[::-1].copy()
    both()
    # (3) two rows' weights swapped
    i, j = df.index[10], df.index[500]
    wi, wj = df.loc[i, "areawt"], df.loc[j, "areawt"]
    df.loc[i, "areawt"], df.loc[j, "areawt"] = wj, wi
    both()
    # the multi-weight entry point keeps its own stacked frame: same hazard
    m1 = weighted_aggregate_grid_to_regions_multi(ds, "tas", ["popwt", "areawt"], "hierid", df)
    df.loc[mid, "hierid"] = df["hierid"].values[-1]
    m2 = weighted_aggregate_grid_to_regions_multi(ds, "tas", ["popwt", "areawt"], "hierid", df)
    ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, df, "areawt", "hierid")
    check(m2["tas_areawt"].values, ref, scale)
    assert m1["tas_areawt"].shape[1] != m2["tas_areawt"].shape[1] or not np.array_equal(
        np.nan_to_num(m1["tas_areawt"].values), np.nan_to_num(m2["tas_areawt"].values))


def test_public_api_from_pinned_memory_pulls_and_returns_in_chunks():
    """Pinned host source through the public API: the GPU pulls the referenced pieces (ctb_pull_pack) and
    the result comes back in time chunks (ctb_copy_rows_to_host) while later chunks are pulled."""
    lat, lon, df, tas, _, _ = _config(1.0, 3000, 330)
    host = torch.empty(tas.shape, dtype=torch.float32, pin_memory=True)
    host.copy_(torch.from_numpy(tas))
    time = pd.date_range("2004-01-01", periods=330)             # crosses Feb 29: a lazy time take on top
    ds = Dataset({"tas": (("time", "lat", "lon"), host.numpy())}, coords={"time": time, "lat": lat, "lon": lon})
    ds1 = tas_poly(ds, 1, "t1")
    n0 = E.TRANSFER_BYTES["d2h"]
    out = weighted_aggregate_grid_to_regions(ds1, "t1", "popwt", "hierid", weights=df)
    xt, tl = oracle.tas_poly(tas, time, 1)
    ref, rd, labels, scale = oracle_agg(xt, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
    assert out.t1.shape == ref.shape == (329, len(labels))
    check(out.t1.values, ref, scale)
    assert E.TRANSFER_BYTES["d2h"] - n0 == 8 * ref.size         # every result byte copied exactly once
    forced = weighted_aggregate_grid_to_regions(ds1, "t1", "popwt", "hierid", weights=df, ingest="pack")
    np.testing.assert_array_equal(forced.t1.values, out.t1.values)


def test_back_to_back_host_calls_do_not_share_staging():
    """Two host-input calls with different data and keep_on_device=True: the second call packs into
    the pinned staging buffers the first call's last H2D copies may still be reading."""
    lat, lon, df, tas, _, _ = _config(1.0, 3000, 96, nan_frac=0.0)
    mk = lambda a: Dataset({"tas": (("time", "lat", "lon"), a)},
                           coords={"time": np.arange(a.shape[0]), "lat": lat, "lon": lon})
    tas2 = (tas + 7.0).astype(np.float32)
    refs = [oracle_agg(a, ("time", "lat", "lon"), lat, lon, df, "areawt", "hierid") for a in (tas, tas2)]
    for _ in range(3):
        r1 = weighted_aggregate_grid_to_regions(mk(tas), "tas", "areawt", "hierid", weights=df, keep_on_device=True)
        r2 = weighted_aggregate_grid_to_regions(mk(tas2), "tas", "areawt", "hierid", weights=df, keep_on_device=True)
        torch.cuda.synchronize()
        for r, (ref, rd, labels, scale) in zip((r1, r2), refs):
            check(r["tas"].data.cpu().numpy(), ref, scale)


# ----------------------------------------------------------------------------
# fused time reduction: annual sums inside the kernel
# ----------------------------------------------------------------------------
@pytest.mark.parametrize("where", ["host", "device"])
@pytest.mark.parametrize("kind", ["identity", "edd", "cell_major"])
def test_time_groups_annual_sums(where, kind):
    """time_groups='year' == groupby(year).sum() of the daily result; years of 365 days do not align
    with the 32-day tiles, 800 days of host input arrive in several chunks."""
    lat, lon, df, tas, tmin, tmax = _config(1.0, 3000, 800)
    time = pd.date_range("2001-01-01", periods=800)
    dims = ("time", "lat", "lon")
    coords = {"time": time, "lat": lat, "lon": lon}
    if kind == "edd":
        tn = DataArray(_put(tmin, where), dims=dims, coords=coords, attrs={"units": "K"})
        tx = DataArray(_put(tmax, where), dims=dims, coords=coords, attrs={"units": "K"})
        ds = Dataset(coords=coords)
        ds["v"] = snyder_edd(tn, tx, 288.15)
        f = oracle.snyder_edd(tmin, tmax, 288.15)
        fdims = dims
    elif kind == "cell_major":
        fdims = ("lat", "lon", "time")
        f = np.ascontiguousarray(tas.transpose(1, 2, 0))
        ds = Dataset({"v": (fdims, _put(f, where))}, coords=coords)
    else:
        ds = Dataset({"v": (dims, _put(tas, where))}, coords=coords)
        f, fdims = tas, dims
    out = weighted_aggregate_grid_to_regions(ds, "v", "popwt", "hierid", weights=df, time_groups="year")
    ref, rd, labels, scale = oracle_agg(f, fdims, lat, lon, df, "popwt", "hierid")
    ref_y, years = oracle.time_group_sum(ref, rd, time.year.values)
    scale_y, _ = oracle.time_group_sum(np.nan_to_num(scale), rd, time.year.values)
    assert out.v.dims == rd and list(out.time.values) == list(years) == [2001, 2002, 2003]
    check(out.v.values, ref_y, scale_y)


def test_time_groups_blocks_and_split_regions():
    """Block-length groups shorter than a tile (several groups per 32-day tile) and regions split
    over bundles (partial rows reduced by the fix-up kernel)."""
    lat, lon, df, tas, _, _ = _config(1.0, 40, 75)
    ds = Dataset({"tas": (("time", "lat", "lon"), torch.from_numpy(tas).cuda())},
                 coords={"time": np.arange(75), "lat": lat, "lon": lon})
    ref, rd, labels, scale = oracle_agg(tas, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
    for p, budget in ((7, 0), (10, 8 * 1024), (75, 0), (1, 0)):
        out = weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df, time_groups=p,
                                                 smem_budget=budget)
        ref_g, gl = oracle.time_group_sum(ref, rd, np.arange(75) // p)
        scale_g, _ = oracle.time_group_sum(np.nan_to_num(scale), rd, np.arange(75) // p)
        assert out.tas.shape == ref_g.shape
        check(out.tas.values, ref_g, scale_g)


# ----------------------------------------------------------------------------
# growing-season mask fused into the kernel (reference utils.py:83-153)
# ----------------------------------------------------------------------------
def _calendar(lat, lon_std, rng):
    """Crop calendar on the grid of the data, stored the reference's way: longitudes on 0..360 (= +180)."""
    nlat, nlon = len(lat), len(lon_std)
    plant = rng.integers(1, 366, (nlat, nlon)).astype(np.float64)
    harv = rng.integers(1, 366, (nlat, nlon)).astype(np.float64)
    harv[rng.random((nlat, nlon)) < 0.05] = np.nan
    plant[rng.random((nlat, nlon)) < 0.05] = np.nan
    gd = Dataset({"variable": (("z", "latitude", "longitude"), np.stack([plant, harv]))},
                 coords={"z": np.array([1, 2]), "latitude": lat, "longitude": lon_std + 180})
    return gd, plant, harv


@pytest.mark.parametrize("where", ["host", "device"])
@pytest.mark.parametrize("kind", ["identity", "edd_annual", "cell_major", "lon360_leap"])
def test_growing_season_mask_fused(where, kind):
    """aggregate(ds, season_mask=m) == oracle aggregate of x * dense mask (1 / 0 / NaN)."""
    from climate_toolbox_b200.utils.utils import get_daily_growing_season_mask, remove_leap_days
    rng = np.random.default_rng(21)
    T = 420 if kind == "edd_annual" else 75
    lat, lon, df, tas, tmin, tmax = _config(1.0, 3000, T, lon_0_360=(kind == "lon360_leap"))
    time = pd.date_range("2003-12-01" if kind != "lon360_leap" else "2004-02-01", periods=T)
    lon_std = lon if kind != "lon360_leap" else oracle.convert_lons_split(lon)[0]
    gd, plant, harv = _calendar(lat, lon_std, rng)
    dims = ("time", "lat", "lon")
    coords = {"time": time, "lat": lat, "lon": lon}
    doy = time.dayofyear.values
    dense = oracle.growing_season_mask(plant, harv, doy).transpose(2, 0, 1)        # (time, lat, lon) on lon_std
    kw = {}
    if kind == "identity":
        ds = Dataset({"v": (dims, _put(tas, where))}, coords=coords)
        f, fdims = tas.astype(np.float64), dims
    elif kind == "cell_major":
        fdims = ("lat", "lon", "time")
        x = np.ascontiguousarray(tas.transpose(1, 2, 0))
        ds = Dataset({"v": (fdims, _put(x, where))}, coords=coords)
        f, dense = x.astype(np.float64), dense.transpose(1, 2, 0)
    elif kind == "edd_annual":
        tn = DataArray(_put(tmin, where), dims=dims, coords=coords, attrs={"units": "K"})
        tx = DataArray(_put(tmax, where), dims=dims, coords=coords, attrs={"units": "K"})
        ds = Dataset(coords=coords)
        ds["v"] = snyder_edd(tn, tx, 288.15)
        f, fdims = oracle.snyder_edd(tmin, tmax, 288.15), dims
        kw = {"time_groups": "year"}
    else:   # 0..360 longitudes (lazy roll) and a leap day removed (lazy time take): the gate follows both
        ds = remove_leap_days(load_bcsd(Dataset({"v": (dims, _put(tas, where))}, coords=coords), "v"))
        keep = ~((time.month == 2) & (time.day == 29))
        _, perm = oracle.convert_lons_split(lon)
        f, fdims = np.take(tas, perm, axis=2)[keep].astype(np.float64), dims
        dense, time = dense[keep], time[keep]
    m = get_daily_growing_season_mask(ds["lat"], ds["lon"], ds["time"], gd)
    out = weighted_aggregate_grid_to_regions(ds, "v", "cropwt", "hierid", weights=df, season_mask=m, **kw)
    ref, rd, labels, scale = oracle_agg(f * dense, fdims, lat, lon_std, df, "cropwt", "hierid")
    scale = np.nan_to_num(scale)
    if kw:
        ref, years = oracle.time_group_sum(ref, rd, time.year.values)
        scale, _ = oracle.time_group_sum(scale, rd, time.year.values)
        assert list(out.time.values) == list(years)
    assert out.v.dims == rd
    check(out.v.values, ref, scale + 1e-300)
    plain = weighted_aggregate_grid_to_regions(ds, "v", "cropwt", "hierid", weights=df, **kw)
    assert not np.allclose(np.nan_to_num(plain.v.values), np.nan_to_num(out.v.values))      # the gate did something


# ----------------------------------------------------------------------------
# full size (BASELINE.json configs[1]): size-independent properties
# ----------------------------------------------------------------------------
def test_full_size_properties():
    """0.25 deg x 24,378 regions x 365 days on the device: (a) a constant field aggregates to
    the constant wherever den > 0; (b) linearity agg(a*x + b) = a*agg(x) + b; (c) staged and
    direct kernels agree; (d) a 3-day slice matches the oracle."""
    d, R, T = 0.25, 24378, 365
    lat, lon = synthetic.grid_labels(d)
    df = _table(d, R)
    g = torch.Generator(device="cuda").manual_seed(7)
    x = 288.0 + 10.0 * torch.randn((T, len(lat), len(lon)), generator=g, device="cuda", dtype=torch.float32)
    coords = {"time": np.arange(T), "lat": lat, "lon": lon}

    def agg(t, **kw):
        ds = Dataset({"tas": (("time", "lat", "lon"), t)}, coords=coords)
        return weighted_aggregate_grid_to_regions(ds, "tas", "popwt", "hierid", weights=df,
                                                  keep_on_device=True, **kw)["tas"].data

    y = agg(x)
    assert y.shape == (T, R)
    plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid")
    ok = torch.from_numpy(plan.den() > 0).cuda()
    c = agg(torch.full_like(x, 3.25))
    assert torch.allclose(c[:, ok], torch.full_like(c[:, ok], 3.25), rtol=1e-12, atol=0)
    a, b = 0.5, 16.0    # exact in fp32: a*x + b introduces fp32 rounding, so compare loosely
    y2 = agg(x * a + b)
    assert torch.allclose(y2[:, ok], a * y[:, ok] + b, rtol=1e-5)
    yd = agg(x, variant=N.VARIANT_DIRECT)
    assert torch.allclose(y[:, ok], yd[:, ok], rtol=1e-12, atol=0)
    sl = x[100:103].cpu().numpy()
    ref, rd, labels, scale = oracle_agg(sl, ("time", "lat", "lon"), lat, lon, df, "popwt", "hierid")
    check(y[100:103].cpu().numpy(), ref, scale)


# ----------------------------------------------------------------------------
# randomized parity sweep (small sizes, many shapes)
# ----------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(12))
def test_randomized_parity(seed):
    """Random grid sizes (odd and even, so both the 16-byte and the scalar staging paths),
    dtypes, layouts, time lengths (ragged 32-day blocks), NaN densities, weight holes, region
    counts from 1 to more than one tile's worth, host and device inputs, random transforms."""
    rng = np.random.default_rng(1000 + seed)
    nlat, nlon = int(rng.integers(3, 40)), int(rng.integers(3, 60))
    if seed % 3 == 0:
        nlon = (nlon // 4 + 1) * 4                       # aligned planes -> vector loads
    T = int(rng.choice([1, 2, 31, 32, 33, 64, 77]))
    dtype = np.float32 if rng.random() < 0.6 else np.float64
    lat = np.sort(rng.choice(np.arange(-90, 90, 0.25), nlat, replace=False))
    lon = np.sort(rng.choice(np.arange(-180, 180, 0.25), nlon, replace=False))
    n_rows = int(rng.integers(1, 4 * nlat * nlon))
    n_reg = int(rng.integers(1, max(2, n_rows // 3)))
    wts = rng.lognormal(0, 1, n_rows)
    wts[rng.random(n_rows) < 0.2] = np.nan
    wts[rng.random(n_rows) < 0.1] = 0.0
    area = rng.random(n_rows)
    area[rng.random(n_rows) < 0.05] = np.nan
    labels = rng.integers(0, n_reg, n_rows).astype(object)
    labels[rng.random(n_rows) < 0.03] = np.nan
    df = pd.DataFrame({"lat": rng.choice(lat, n_rows), "lon": rng.choice(lon, n_rows),
                       "hierid": labels, "popwt": wts, "areawt": area})
    x = (280 + 15 * rng.standard_normal((T, nlat, nlon))).astype(dtype)
    x[rng.random(x.shape) < rng.choice([0.0, 0.01, 0.3])] = np.nan
    spread = np.abs(4 * rng.standard_normal(x.shape)).astype(dtype)
    dims = ("time", "lat", "lon")
    layout_llt = rng.random() < 0.3
    on_device = rng.random() < 0.5

    def put(a):
        if layout_llt:
            a = np.ascontiguousarray(a.transpose(1, 2, 0))
        return torch.from_numpy(a).cuda() if on_device else a

    d = ("lat", "lon", "time") if layout_llt else dims
    coords = {"time": np.arange(T), "lat": lat, "lon": lon}
    kind = ["identity", "poly", "edd", "gdd"][seed % 4]
    ds = Dataset({"tas": (d, put(x))}, coords=coords)
    if kind == "identity":
        f, name = x.astype(np.float64), "tas"
    elif kind == "poly":
        p = int(rng.integers(1, 5))
        from climate_toolbox_b200._xr import Deferred, Variable
        ds._vars["v"] = Variable(d, None, None, None, Deferred("poly", (273.15, float(p)), (ds._vars["tas"],)))
        f, name = (x.astype(np.float64) - 273.15) ** p, "v"
    else:
        tn = DataArray(put(x - spread), dims=d, coords=coords, attrs={"units": "K"})
        tx = DataArray(put(x + spread), dims=d, coords=coords, attrs={"units": "K"})
        lo_, hi_ = 283.15, 291.0
        if kind == "edd":
            ds["v"] = snyder_edd(tn, tx, lo_)
            f = oracle.snyder_edd(x - spread, x + spread, lo_)
        else:
            ds["v"] = snyder_gdd(tn, tx, lo_, hi_)
            f = oracle.snyder_gdd(x - spread, x + spread, lo_, hi_)
        name = "v"
    out = weighted_aggregate_grid_to_regions(ds, name, "popwt", "hierid", weights=df)
    ref, rd, labels_o, scale = oracle_agg(f, dims, lat, lon, df, "popwt", "hierid")
    got = out[name]
    assert list(out.hierid.values) == list(labels_o)
    got_v = got.values if got.dims == rd else got.transpose(*rd).values
    # EDD/GDD: the kernel's polynomial form of the straddling branch is exact to 3e-14 * max(|EDD|, W)
    # (W = half the daily range, here <= max|spread|), so W joins the scale of the 1e-9 test
    scale = scale + (np.abs(spread).max() if kind in ("edd", "gdd") else 0.0)
    check(got_v, ref, scale)


@pytest.mark.parametrize("engine", [0, 1])
def test_push_rows_copies_the_column_block(engine):
    """ctb_push_rows (the multi-GPU gather as a push) on one GPU: the "peers" are two more buffers of this
    device.  The column block arrives bit-exact, everything else stays untouched, the source is skipped."""
    import ctypes as C

    from climate_toolbox_b200 import _engine as E
    from climate_toolbox_b200 import _native as N
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    n_rows, ld, t0, n = 997, 1460, 183, 181
    src = torch.randn((n_rows, ld), generator=g, device=dev, dtype=torch.float64)
    peers = [torch.full((n_rows, ld), -7.0, device=dev, dtype=torch.float64) for _ in range(2)]
    keep = src.clone()
    arr = (C.c_void_p * 3)(peers[0].data_ptr(), src.data_ptr(), peers[1].data_ptr())
    N.check(N.lib().ctb_push_rows(C.c_void_p(src.data_ptr()), ld, t0, n, n_rows, 3, arr, engine, E._stream_ptr(dev)))
    torch.cuda.synchronize()
    assert torch.equal(src, keep)
    for p in peers:
        assert torch.equal(p[:, t0:t0 + n], src[:, t0:t0 + n])
        assert bool((p[:, :t0] == -7.0).all()) and bool((p[:, t0 + n:] == -7.0).all())
    # a block wider than one pass of the copy kernel (256 columns per warp and pass), odd sizes
    wide = torch.zeros((5, ld), device=dev, dtype=torch.float64)
    arr1 = (C.c_void_p * 1)(wide.data_ptr())
    N.check(N.lib().ctb_push_rows(C.c_void_p(src.data_ptr()), ld, 3, 777, 5, 1, arr1, engine, E._stream_ptr(dev)))
    torch.cuda.synchronize()
    assert torch.equal(wide[:, 3:780], src[:5, 3:780]) and float(wide[:, 780:].abs().sum()) == 0.0


@pytest.mark.parametrize("kind,params,n_out", [("identity", (), 1), ("poly", (273.15, 1, 2), 2)])
def test_fused_peer_epilogue_on_one_gpu(kind, params, n_out):
    """The gather fused into the kernel's epilogue (ctb_agg_opts.peer_out / peer_row), driven on ONE GPU: the
    "peers" are two buffers of this device.  Both receive what the plain launch writes -- rows permuted by
    peer_row, columns offset like a time shard -- including the split regions the fix-up kernel finishes."""
    from climate_toolbox_b200 import _engine as E
    from climate_toolbox_b200 import _native as N
    lat, lon, df, tas, _, _ = _config(1.0, 300, 70, seed=9)
    df = df.copy()
    df.loc[df.index[:4000], "hierid"] = df["hierid"].iloc[0]        # one region larger than a tile: split rows
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(tas).to(dev).view(tas.shape[0], -1)
    T, ncell = x.shape
    plan = E.get_plan(E.GridSpec(lat, lon), df, "popwt", "hierid", device=dev)
    ref = E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, ncell, None, T, kind, params, n_out)
    t_total, t0 = T + 37, 21                                           # this "rank" owns days [21, 21 + T)
    bufs = [torch.full((n_out, plan.R, t_total), -5.0, device=dev, dtype=torch.float64) for _ in range(2)]
    row = torch.as_tensor(plan.region_order(), device=dev)
    E.aggregate_device(plan, x, None, N.LAYOUT_TIME_MAJOR, ncell, None, T, kind, params, n_out,
                       out=E._OffsetOut(bufs[0], t0), out_ld=t_total,
                       peer_ptrs=[b.data_ptr() + 8 * t0 for b in bufs], peer_row=row)
    torch.cuda.synchronize()
    for b in bufs:
        got = b.index_select(1, row.to(torch.int64))[:, :, t0:t0 + T]
        assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(ref))
        assert bool((b[:, :, :t0] == -5.0).all()) and bool((b[:, :, t0 + T:] == -5.0).all())


def test_ipc_buffer_alloc_and_free():
    """ctb_ipc_alloc hands out device memory + a 64-byte handle for the other ranks; the owner frees it.
    (Opening a handle needs a second process: __graft_entry__.smoke() and bench.py --gpus N do that.)"""
    import ctypes as C

    from climate_toolbox_b200 import _native as N
    from climate_toolbox_b200.parallel import _tensor_from_ptr
    ptr = C.c_void_p()
    handle = (C.c_ubyte * N.IPC_HANDLE_BYTES)()
    N.check(N.lib().ctb_ipc_alloc(8 * 1000, 0, C.byref(ptr), handle))
    assert ptr.value and any(handle)
    t = _tensor_from_ptr(ptr.value, (10, 100), torch.device("cuda", 0), None)
    t.copy_(torch.arange(1000, dtype=torch.float64, device="cuda").view(10, 100))
    assert float(t.sum()) == 999 * 1000 / 2
    del t
    torch.cuda.synchronize()
    N.check(N.lib().ctb_ipc_free(ptr, 0))
