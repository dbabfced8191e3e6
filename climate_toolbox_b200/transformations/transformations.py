"""Gridcell-level transforms (mirror of
``/root/reference/climate_toolbox/transformations/transformations.py:7-214``).

Each function returns variables holding a *deferred* transform: aggregating them
fuses the arithmetic into the gather kernel (no full-grid temporaries);
``.values`` evaluates them with the pointwise CUDA kernel.  Arithmetic is float64
from the first operation whatever the storage dtype (dtype policy).
"""
from __future__ import annotations

import numpy as np
import torch

from .._xr import DataArray, Dataset, Deferred, Variable, from_any, to_like
from ..utils.utils import remove_leap_days, convert_kelvin_to_celsius  # noqa: F401

__all__ = ["snyder_edd", "snyder_gdd", "validate_edd_snyder_agriculture", "tas_poly", "ordinal"]


def _source(da):
    v = da.variable
    return v if v.deferred is None else Variable(v.dims, v.values, v.attrs)


def _any_less(a, b):
    """``(a < b).any()`` on the data's own device (precondition check only)."""
    pa, pb = a.physical, b.physical
    if isinstance(pa, torch.Tensor) and isinstance(pb, torch.Tensor) and not a.takes and not b.takes:
        return bool((pa < pb).any().item())
    return bool((a.values < b.values).any())


def snyder_edd(tasmin, tasmax, threshold, check=True):
    r"""
    Snyder exceedance degree days/cooling degree days (reference ``:7-93``).

    .. math::

        EDD_d = \begin{cases}
            ((M - e)(\pi/2 - \theta) + w\cos\theta)/\pi & tmin_d < e < tmax_d \\
            0 & tmax_d < e \\
            M - e & \text{otherwise} \end{cases}
        \quad M = (tmax_d + tmin_d)/2,\ w = (tmax_d - tmin_d)/2,\ \theta = \arcsin((e - M)/w)

    ``tasmin`` / ``tasmax`` : DataArray with a ``units`` attribute; ``threshold`` in
    the same units.  ``check=False`` skips the ``tasmax >= tasmin`` scan.
    """
    tasmin, tasmax = from_any(tasmin), from_any(tasmax)
    # Check for unit agreement (AttributeError if a `.units` attr is missing)
    assert tasmin.units == tasmax.units
    smin, smax = _source(tasmin), _source(tasmax)
    if check:
        assert not _any_less(smax, smin), "values encountered where tasmin > tasmax"
    attrs = {"units": "degreedays_{}{}".format(threshold, tasmax.attrs["units"])}
    var = Variable(smin.dims, None, attrs, None,
                   Deferred("edd", (float(threshold),), (smin, smax)))
    return DataArray._wrap(var, tasmax._coords, tasmax.name)


def snyder_gdd(tasmin, tasmax, threshold_low, threshold_high, check=True):
    r"""
    Snyder growing degree days (reference ``:96-147``):
    ``GDD = EDD(threshold_low) - EDD(threshold_high)`` per gridcell-day.
    """
    tasmin, tasmax = from_any(tasmin), from_any(tasmax)
    assert tasmin.units == tasmax.units
    smin, smax = _source(tasmin), _source(tasmax)
    if check:
        assert not _any_less(smax, smin), "values encountered where tasmin > tasmax"
    attrs = {"units": "degreedays_{}-{}{}".format(threshold_low, threshold_high,
                                                  tasmax.attrs["units"])}
    var = Variable(smin.dims, None, attrs, None,
                   Deferred("gdd", (float(threshold_low), float(threshold_high)), (smin, smax)))
    return DataArray._wrap(var, tasmax._coords, tasmax.name)


def validate_edd_snyder_agriculture(ds, thresholds):
    """reference ``:150-157``"""
    msg_null = "hierid dims do not match 24378"

    assert ds.hierid.shape == (24378,), msg_null

    for threshold in thresholds:
        assert threshold in list(ds.refTemp)
    return


def tas_poly(ds, power, varname):
    """
    Daily average temperature (degrees C), raised to a power (reference ``:160-208``).

    Leap years are removed before counting days (uses a 365 day calendar).
    ``power`` / ``varname`` may also be equal-length lists: all orders are then
    produced from ONE read of ``tas`` when aggregated.
    """
    like = ds
    ds = from_any(ds)
    powers = [power] if np.isscalar(power) else list(power)
    names = [varname] if isinstance(varname, str) else list(varname)
    if len(powers) != len(names):
        raise ValueError("power and varname must have the same length")

    # remove leap years
    ds = remove_leap_days(ds)
    tas = _source(ds["tas"])

    # Replace datetime64[ns] 'time' with YYYYDDD int 'day'
    if ds.dims["time"] > 365:
        raise ValueError
    ds1 = Dataset()
    for k, c in ds._coords.items():
        if k != "time":
            ds1._coords[k] = c
    year = np.asarray(ds["time.year"].values, dtype=np.int64)
    ds1._coords["time"] = Variable(("time",), year * 1000 + np.arange(1, len(year) + 1))

    for p, name in zip(powers, names):
        powername = ordinal(p)
        description = (
            """
            Daily average temperature (degrees C){raised}

            Leap years are removed before counting days (uses a 365 day
            calendar).
            """.format(
                raised="" if p == 1 else (" raised to the {powername} power".format(powername=powername))
            )
        ).strip()
        attrs = {"units": "C^{}".format(p) if p > 1 else "C",
                 "long_title": description.splitlines()[0], "description": description,
                 "variable": name}
        # do transformation: (tas - 273.15) ** power, deferred
        ds1._vars[name] = Variable(tas.dims, None, attrs, None,
                                   Deferred("poly", (273.15, float(p)), (tas,)))
    return to_like(ds1, like)


def ordinal(n):
    """Converts numbers into ordinal strings"""
    return "%d%s" % (n, "tsnrhtdd"[(n // 10 % 10 != 1) * (n % 10 < 4) * n % 10:: 4])
