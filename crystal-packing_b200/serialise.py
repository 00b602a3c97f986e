"""JSON state output in the layout of the reference's `serde_json::to_string(&final_state)`
(crystal-packing-cli/src/main.rs:263-265) — SURVEY.md §8(f) row 1, the first "next" row after the
hot path.

The field order follows the `#[derive(Serialize)]` declarations:
  PackedState / PotentialState  state/packed.rs:24-33, state/potential.rs:23-32
  Wallpaper                     wallpaper.rs:24-28       WyckoffSite   wallpaper.rs:48-55
  Cell2                         cell.rs:48-54            OccupiedSite  site.rs:13-19
  SharedValue -> bare f64       basis/shared_value.rs:63-70
  LineShape / Line2             shape/line_shape.rs:26-30, components/line2.rs:16-20
  MolecularShape2 / Atom2       shape/molecular_shape2.rs:20-24, components/atom2.rs:14-18
  LJShape2 / LJ2                shape/lj_shape.rs:20-24, components/lj2.rs:18-29
  Transform2 (newtype over nalgebra::Transform2 -> Matrix3 -> 9 numbers, column-major)
                                transform.rs:33-34
nalgebra `Point2` serialises as `[x, y]`, `Option<f64>` as a number or `null`, unit enum variants
as strings.  Floats are printed the way serde_json does (ryu: shortest round-trip digits, `1.0`
for integral values, exponent form outside 1e-5 <= |v| < 1e16).

UNVERIFIED against the Rust binary (it cannot be built in this environment): the layout is derived
from the type definitions and the pinned crate versions.  Host-side formatting only — nothing here
touches the Monte-Carlo path.
"""
import math

from . import capi

_FAMILY = {capi.CP_MONOCLINIC: "Monoclinic", capi.CP_ORTHORHOMBIC: "Orthorhombic",
           capi.CP_HEXAGONAL: "Hexagonal", capi.CP_TETRAGONAL: "Tetragonal"}


def format_f64(v):
    """serde_json / ryu formatting of an f64 (non-finite values become null, as serde_json does)."""
    if math.isnan(v) or math.isinf(v):
        return "null"
    if v == 0.0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    r = repr(abs(v))
    if "e" in r:
        mant, exp = r.split("e")
        exp = int(exp)
    else:
        mant, exp = r, 0
    ip, _, fp = mant.partition(".")
    digits = (ip + fp).lstrip("0")
    kk = len(ip) + exp if ip != "0" else exp - (len(fp) - len(fp.lstrip("0")))
    digits = digits.rstrip("0") or "0"
    n = len(digits)
    sign = "-" if v < 0 else ""
    if 0 < kk <= 16:
        if n <= kk:
            out = digits + "0" * (kk - n) + ".0"
        else:
            out = digits[:kk] + "." + digits[kk:]
    elif -5 < kk <= 0:
        out = "0." + "0" * (-kk) + digits
    else:
        e = kk - 1
        out = digits[0] + ("." + digits[1:] if n > 1 else "") + "e" + str(e)
    return sign + out


def _obj(pairs):
    return "{" + ",".join(f'"{k}":{v}' for k, v in pairs) + "}"


def _arr(items):
    return "[" + ",".join(items) + "]"


def _point(x, y):
    return _arr([format_f64(x), format_f64(y)])


def _shape_json(shape):
    p = shape._problem
    items = []
    for k in range(p.n_items):
        it = p.items[k]
        if p.kind == capi.CP_POLYGON:
            items.append(_obj([("start", _point(it[0], it[1])), ("end", _point(it[2], it[3]))]))
        elif p.kind == capi.CP_DISCS:
            items.append(_obj([("position", _point(it[0], it[1])), ("radius", format_f64(it[2]))]))
        else:
            items.append(_obj([("position", _point(it[0], it[1])), ("sigma", format_f64(it[2])),
                               ("epsilon", format_f64(it[3])), ("cutoff", format_f64(it[4]))]))
    return _obj([("name", f'"{shape.name}"'), ("items", _arr(items))])


def _transform_json(rows6):
    # rows 0-1 of the operator; row 2 is all zero (transform.rs:139).  Column-major 3x3.
    m = [[rows6[0], rows6[1], rows6[2]], [rows6[3], rows6[4], rows6[5]], [0.0, 0.0, 0.0]]
    return _arr([format_f64(m[r][c]) for c in range(3) for r in range(3)])


def state_to_json(state):
    """serde_json::to_string(&state) for a PackedState / PotentialState of this package."""
    p = state._problem
    family = f'"{_FAMILY[p.family]}"'
    wallpaper = _obj([("name", f'"{state.wallpaper.name}"'), ("family", family)])
    cell = _obj([("length", format_f64(p.dof[0])), ("ratio", format_f64(p.dof[1])),
                 ("angle", format_f64(p.dof[2])), ("family", family)])
    wyckoff = _obj([("letter", '"a"'),
                    ("symmetries", _arr([_transform_json(list(p.syms[s])) for s in range(p.n_syms)])),
                    ("num_rotations", "1"), ("mirror_primary", "false"), ("mirror_secondary", "false")])
    site = _obj([("wyckoff", wyckoff), ("x", format_f64(p.dof[3])), ("y", format_f64(p.dof[4])),
                 ("angle", format_f64(p.dof[5]))])
    return _obj([("wallpaper", wallpaper), ("shape", _shape_json(state.shape)), ("cell", cell),
                 ("occupied_sites", _arr([site]))])


def state_from_json(text, state_cls, shape_types):
    """The reverse direction for `--start-config` (declared but never read by the reference,
    main.rs:179-181): rebuilds a state of this package from JSON in the layout above."""
    import json

    from . import api

    d = json.loads(text)
    items = d["shape"]["items"]
    if "start" in items[0]:
        shape = api.LineShape.from_radial(d["shape"]["name"], [1.0] * len(items))
        for k, it in enumerate(items):
            for c, v in enumerate(list(it["start"]) + list(it["end"])):
                shape._problem.items[k][c] = v
    elif "radius" in items[0]:
        shape = api.MolecularShape2.circle()
        shape._problem.n_items = len(items)
        for k, it in enumerate(items):
            shape._problem.items[k][0], shape._problem.items[k][1] = it["position"]
            shape._problem.items[k][2] = it["radius"]
    else:
        shape = api.LJShape2.circle()
        shape._problem.n_items = len(items)
        for k, it in enumerate(items):
            shape._problem.items[k][0], shape._problem.items[k][1] = it["position"]
            shape._problem.items[k][2], shape._problem.items[k][3] = it["sigma"], it["epsilon"]
            shape._problem.items[k][4] = float("nan") if it["cutoff"] is None else it["cutoff"]
    if not isinstance(shape, tuple(shape_types)):
        raise TypeError(f"{state_cls.__name__} cannot hold a {type(shape).__name__}")
    shape.name = d["shape"]["name"]
    api.refresh_shape(shape)  # area and enclosing radius are functions of the items
    site = d["occupied_sites"][0]
    group = api.group_from_symmetries(d["wallpaper"]["name"], d["wallpaper"]["family"],
                                      site["wyckoff"]["symmetries"])
    state = state_cls.from_group(shape, group)
    cell = d["cell"]
    state.dof = [cell["length"], cell["ratio"], cell["angle"], site["x"], site["y"], site["angle"]]
    return state
