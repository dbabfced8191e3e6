import os
import sys

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def reference_fixtures():
    """Transcription of /root/reference/tests/test_climate_toolbox.py:33-106
    (``pd.Timestamp`` instead of the removed ``pd.datetime``).  The draw order --
    ``seed(42)``, clim_data, then weights -- is the one pytest produces there."""
    lat = np.arange(-89.875, 90, 2)
    lon = np.arange(0.125, 360.0, 2)
    time = pd.date_range(start=pd.Timestamp(2000, 1, 1), periods=10, freq="D")
    np.random.seed(42)
    temp = np.random.rand(len(lat), len(lon), len(time)) * 100

    df = pd.DataFrame()
    lats = np.random.choice(lat, 100)
    lons = np.random.choice(lon, 100)
    df["lat"] = lats
    df["lon"] = lons
    df["areawt"] = np.random.random(len(df["lon"]))
    tmp = np.random.random(len(df["lon"]))
    tmp[::5] = np.nan
    df["popwt"] = tmp
    df["hierid"] = np.random.choice(np.arange(1, 25), len(lats))
    mapping = {h: np.random.choice(np.arange(1, 5)) for h in df["hierid"].values}
    df["ISO"] = [mapping[i] for i in df["hierid"]]
    df.index.names = ["reshape_index"]
    return lat, lon, time, temp, df


@pytest.fixture
def ref_fix():
    return reference_fixtures()


def rel_err(got, ref, scale=None):
    """|got-ref| / max(|ref|, scale): the cancellation-aware relative error of
    SURVEY.md 7.3-6 (scale = sum|w f(x)| / sum w when given)."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.array_equal(np.isnan(got), np.isnan(ref)), "NaN pattern differs"
    m = ~np.isnan(ref) & np.isfinite(ref)
    assert np.array_equal(got[~m & ~np.isnan(ref)], ref[~m & ~np.isnan(ref)]), "inf pattern differs"
    if not m.any():
        return 0.0
    den = np.abs(ref[m])
    if scale is not None:
        den = np.maximum(den, np.broadcast_to(scale, ref.shape)[m])
    den = np.maximum(den, np.finfo(np.float64).tiny)
    return float(np.max(np.abs(got[m] - ref[m]) / den))
