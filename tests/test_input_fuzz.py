"""Property tests of the input.data parser (InputDataPoroel.h:77-222 grammar): layout noise never changes the parsed
values; out-of-range values are always rejected."""
import math

import pytest
from hypothesis import given, settings, strategies as st

import helpers as H

capi = H.capi

KEYS = [  # (section, key, attribute, lo, hi) for Patterns::Double entries (ID:93-141)
    ("Properties", "Young modulus", "youngs_modulus", 1.0, 1e12),
    ("Properties", "Poisson ratio", "poisson_ratio", 0.0, 0.5),
    ("Properties", "Biot coefficient", "biot_coef", 0.1, 1.0),
    ("Properties", "Porosity", "poro", 1e-5, 0.99999),
    ("Properties", "Viscosity", "visc", 1e-6, 1.0),
    ("Properties", "Well radius", "r_well", 1e-2, 1e3),
    ("Solver", "Time step", "time_step", 1e-8, 1e6),
    ("Solver", "FSS tolerance", "fss_tol", 1e-20, 1e-1),
]
ws = st.text(alphabet=" \t", min_size=0, max_size=4)


@settings(max_examples=60, deadline=None)
@given(data=st.data())
def test_layout_noise_is_ignored(data):
    lines, expect = [], {}
    by_sec = {}
    for sec, key, attr, lo, hi in KEYS:
        v = data.draw(st.floats(min_value=lo, max_value=hi, allow_nan=False, allow_infinity=False))
        by_sec.setdefault(sec, []).append((key, attr, v))
    for sec, entries in by_sec.items():
        lines.append(data.draw(ws) + "subsection" + " " + data.draw(ws) + sec + data.draw(ws))
        for key, attr, v in entries:
            spaced_key = key.replace(" ", " " + data.draw(ws))  # ParameterHandler collapses runs of blanks inside names
            comment = data.draw(st.sampled_from(["", "  # a comment", "#x"]))
            lines.append(f"{data.draw(ws)}set {data.draw(ws)}{spaced_key}{data.draw(ws)}={data.draw(ws)}{v!r}{comment}")
            expect[attr] = v
            if data.draw(st.booleans()):
                lines.append(data.draw(ws) + "# " + data.draw(st.text(alphabet="abc =#", max_size=10)))
        lines.append(data.draw(ws) + "end" + data.draw(ws))
        lines.append("")
    d = capi.InputData(text="\n".join(lines))
    for attr, v in expect.items():
        assert getattr(d, attr) == v
    E, nu = expect["youngs_modulus"], expect["poisson_ratio"]
    if nu < 0.499:
        assert d.lame_constant == pytest.approx(E * nu / ((1 + nu) * (1 - 2 * nu)), rel=1e-12)
    assert d.shear_modulus == pytest.approx(0.5 * E / (1 + nu), rel=1e-14)


@settings(max_examples=40, deadline=None)
@given(i=st.integers(min_value=0, max_value=len(KEYS) - 1), up=st.booleans(), f=st.floats(min_value=1.001, max_value=1e3))
def test_out_of_range_values_are_rejected(i, up, f):
    sec, key, attr, lo, hi = KEYS[i]
    if key in ("Young modulus", "Well radius", "Time step") and up:
        return  # no upper bound declared (Patterns::Double(lo))
    bad = hi * f if up else (lo / f if lo > 0 else -f)
    if bad == lo or bad == hi or math.isinf(bad):
        return
    with pytest.raises(capi.HostError):
        capi.InputData(text=f"subsection {sec}\n  set {key} = {bad!r}\nend\n")


@settings(max_examples=30, deadline=None)
@given(vals=st.lists(st.tuples(st.integers(0, 5), st.integers(0, 1), st.floats(-1e-3, 1e-3, allow_nan=False)), min_size=0, max_size=6))
def test_boundary_lists_round_trip(vals):
    labels = ", ".join(str(v[0]) for v in vals)
    comps = ", ".join(str(v[1]) for v in vals)
    values = ", ".join(repr(v[2]) for v in vals)
    text = (f"subsection In situ\n  set Displacement boundary labels = {labels}\n  set Displacement boundary components = {comps}\n"
            f"  set Displacement boundary values = {values}\nend\n")
    d = capi.InputData(text=text)
    assert list(d.displacement_boundary_labels) == [v[0] for v in vals]
    assert list(d.displacement_boundary_components) == [v[1] for v in vals]
    assert list(d.displacement_boundary_values) == [v[2] for v in vals]
