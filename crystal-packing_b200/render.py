"""Text and SVG output of a state — SURVEY.md §8(f) rows 1 (second half) and 3:

  State::as_positions          crystal-packing/src/state/packed.rs:97-106, potential.rs:123-132
  ToSVG for PackedState / PotentialState / Cell2 / the shapes / Transform2
                               crystal-packing/src/to_svg.rs:15-275 (written by the CLI next to the
                               JSON state, crystal-packing-cli/src/main.rs:266)

Host-side formatting only: nothing here touches the Monte-Carlo path, and nothing here is timed.
The 3x3 matrices are evaluated with the reference's expressions (site.rs:33-45, transform.rs:81-112,
cell.rs:110-112,216-263 and nalgebra's product order) using libm sin/cos, which is
what the Rust crate executes; `Engine.images` returns the same numbers as the kernels evaluate them.

UNVERIFIED against the Rust binary (it cannot be built in this environment): number formatting
follows Rust's `{}` / `{:?}` for f64 and f32, the `svg 0.10.0` crate's writer (attributes sorted by
name, one node per line, path parameters narrowed to f32) and nalgebra 0.27.1's derived `Debug`.
"""
import math
import struct

from . import capi
from .serialise import format_f64


# ---------------------------------------------------------------------------------------- matrices
def _matmul(a, b):
    # nalgebra Matrix3 * Matrix3: ((a_i0 b_0j) + a_i1 b_1j) + a_i2 b_2j
    return [[((a[i][0] * b[0][j]) + a[i][1] * b[1][j]) + a[i][2] * b[2][j] for j in range(3)] for i in range(3)]


def transform_new(angle, translation):
    """Transform2::new (transform.rs:81-87)."""
    s, c = math.sin(angle), math.cos(angle)
    return [[c, -s, translation[0]], [s, c, translation[1]], [0.0, 0.0, 1.0]]


def identity():
    return [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]


def position(m):
    """Transform2::position = transform * origin (transform.rs:93-95): the translation column,
    divided by m22 when that is not zero (nalgebra's general transform)."""
    x, y, n = m[0][2], m[1][2], m[2][2]
    return (x / n, y / n) if n != 0.0 else (x, y)


def set_position(m, p):
    out = [row[:] for row in m]
    out[0][2], out[1][2] = p
    return out


def periodic(m, period, offset):
    """Transform2::periodic (transform.rs:107-112)."""
    x, y = position(m)
    x = math.fmod(math.fmod(x - offset, period) + period, period) + offset
    y = math.fmod(math.fmod(y - offset, period) + period, period) + offset
    return set_position(m, (x, y))


class _Cell:
    """Cell2 accessors (cell.rs:91-101,126-129,216-263) over the DOFs of a state."""

    def __init__(self, dof):
        self.length, self.ratio, self.angle_value = dof[0], dof[1], dof[2]

    def a(self):
        return self.length

    def b(self):
        return self.length * self.ratio

    def angle(self):
        return self.angle_value

    def to_cartesian(self, x, y):
        yb = y * self.b()
        return x * self.a() + yb * math.cos(self.angle_value), yb * math.sin(self.angle_value)

    def to_cartesian_isometry(self, m):
        return set_position(m, self.to_cartesian(*position(m)))

    def to_cartesian_translate(self, m, ix, iy):
        px, py = position(m)
        return set_position(m, self.to_cartesian(px + float(ix), py + float(iy)))

    def periodic_images(self, m, shells, zero):
        for ix in range(-shells, shells + 1):  # iproduct!: x outer, y inner
            for iy in range(-shells, shells + 1):
                if not zero and ix == 0 and iy == 0:
                    continue
                yield self.to_cartesian_translate(m, ix, iy)

    def get_corners(self):
        return [self.to_cartesian(x, y) for x, y in ((-0.5, -0.5), (-0.5, 0.5), (0.5, 0.5), (0.5, -0.5))]


def relative_positions(state):
    """OccupiedSite::positions over the occupied sites (site.rs:33-45, packed.rs:117-119)."""
    p = state._problem
    site = transform_new(p.dof[5], (p.dof[3], p.dof[4]))
    out = []
    for s in range(p.n_syms):
        r = p.syms[s]
        sym = [[r[0], r[1], r[2]], [r[3], r[4], r[5]], [0.0, 0.0, 0.0]]  # row 2 stays zero (transform.rs:139)
        out.append(periodic(_matmul(sym, site), 1.0, -0.5))
    return out


def cartesian_positions(state):
    """PackedState::cartesian_positions (packed.rs:112-115)."""
    cell = _Cell(state.dof)
    return [cell.to_cartesian_isometry(m) for m in relative_positions(state)]


# ---------------------------------------------------------------------------------------- numbers
def display_f64(v):
    """Rust `{}` of an f64: shortest round-trip digits, never an exponent, no trailing `.0`."""
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    r = repr(float(v))
    if "e" in r or "E" in r:
        mant, exp = r.lower().split("e")
        exp = int(exp)
        sign = "-" if mant.startswith("-") else ""
        mant = mant.lstrip("-")
        ip, _, fp = mant.partition(".")
        digits = ip + fp
        point = len(ip) + exp
        if point <= 0:
            r = sign + "0." + "0" * (-point) + digits.rstrip("0")
        elif point >= len(digits):
            r = sign + digits + "0" * (point - len(digits))
        else:
            r = sign + digits[:point] + "." + digits[point:]
    if r.endswith(".0"):
        r = r[:-2]
    return r


def debug_f64(v):
    """Rust `{:?}` of an f64: as `{}` but integral values keep `.0` and magnitudes outside
    [1e-5, 1e16) use the exponent form (the same rule serde_json's ryu output follows)."""
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    return format_f64(v)


def display_f32(v):
    """Rust `{}` of `v as f32`: shortest digits that round-trip through f32."""
    f = struct.unpack("f", struct.pack("f", v))[0]
    if f == 0.0:
        return "-0" if math.copysign(1.0, f) < 0 else "0"
    for digits in range(1, 10):
        s = f"{f:.{digits}g}"
        if struct.unpack("f", struct.pack("f", float(s)))[0] == f:
            break
    return display_f64(float(s))


# ---------------------------------------------------------------------------------------- as_positions
def debug_transform(m):
    """`{:?}` of the reference's Transform2 newtype: derived Debug down to nalgebra's column-major
    ArrayStorage."""
    cols = ", ".join("[" + ", ".join(debug_f64(m[r][c]) for r in range(3)) + "]" for c in range(3))
    return f"Transform2(Transform {{ matrix: Matrix {{ data: ArrayStorage([{cols}]) }}, _phantom: PhantomData }})"


def as_positions(state):
    """State::as_positions (packed.rs:97-106): the cell's Display line (cell.rs:67-77), the word
    `Positions`, then the Debug form of every Cartesian transform."""
    cell = _Cell(state.dof)
    lines = [f"Cell2 {{ a: {display_f64(cell.a())}, b: {display_f64(cell.b())}, angle: {display_f64(cell.angle())} }}",
             "Positions"]
    lines += [debug_transform(m) for m in cartesian_positions(state)]
    return "\n".join(lines) + "\n"


# ---------------------------------------------------------------------------------------- SVG
class _Node:
    """A node of the `svg` crate: attributes are written sorted by name, children one per line."""

    def __init__(self, name, **attrs):
        self.name, self.attrs, self.children = name, {}, []
        for k, v in attrs.items():
            self.set(k, v)

    def set(self, key, value):
        self.attrs[key.replace("_", "-")] = value if isinstance(value, str) else display_f64(value)
        return self

    def add(self, child):
        self.children.append(child)
        return self

    def __str__(self):
        attrs = "".join(f' {k}="{v}"' for k, v in sorted(self.attrs.items()))
        if not self.children:
            return f"<{self.name}{attrs}/>"
        return f"<{self.name}{attrs}>\n" + "\n".join(str(c) for c in self.children) + f"\n</{self.name}>"


def _path_data(points, close=True):
    # element::path::Data: `M` for the first point, `L` for the rest, `z`; parameters are f32
    cmds = [("M" if k == 0 else "L") + display_f32(x) + "," + display_f32(y) for k, (x, y) in enumerate(points)]
    return " ".join(cmds + (["z"] if close else []))


def cell_svg(cell):
    """ToSVG for Cell2 (to_svg.rs:58-78)."""
    return _Node("g").add(_Node("path", fill="None", stroke="grey").set("stroke-width", 0.1)
                          .set("d", _path_data(cell.get_corners())))


def shape_svg(shape):
    """ToSVG for LineShape (to_svg.rs:45-56: one closed path from the first start through every end),
    MolecularShape2 (:92-102: a circle of `radius` per atom) and LJShape2 (:33-43: radius sigma / 2)."""
    g = _Node("g")
    items = shape.items
    if shape.kind == capi.CP_POLYGON:
        points = [(items[0][0], items[0][1])] + [(it[2], it[3]) for it in items]
        return g.add(_Node("path").set("d", _path_data(points)))
    for it in items:
        r = it[2] if shape.kind == capi.CP_DISCS else it[2] / 2.0
        g.add(_Node("circle", r=r, cx=it[0], cy=it[1]))
    return g


def use_svg(m):
    """ToSVG for Transform2 (to_svg.rs:104-125): `matrix(m00 m10 m01 m11 m02 m12)`."""
    vals = (m[0][0], m[1][0], m[0][1], m[1][1], m[0][2], m[1][2])
    return _Node("use", transform="matrix(" + " ".join(display_f64(v) for v in vals) + ")")


def as_svg(state):
    """ToSVG for PackedState / PotentialState (to_svg.rs:127-275): the view box from three times the
    cell corners padded by the enclosing radius, the cell and the shape as definitions, the 3 x 3
    block of cells, every shape of the cell in blue and its first shell of periodic images in green."""
    cell = _Cell(state.dof)
    padding = state.shape.enclosing_radius()
    box = (0.0, 0.0, 0.0, 0.0)
    for x, y in cell.get_corners():
        x, y = x * 3.0, y * 3.0
        box = (min(x - padding, box[0]), min(y - padding, box[1]),
               max(2.0 * (x + padding), box[2]), max(2.0 * (y + padding), box[3]))
    doc = _Node("svg", xmlns="http://www.w3.org/2000/svg", viewBox=" ".join(display_f64(v) for v in box))
    doc.add(_Node("defs").add(cell_svg(cell).set("id", "cell")).add(shape_svg(state.shape).set("id", "mol")))
    for m in cell.periodic_images(identity(), 1, True):
        doc.add(use_svg(m).set("href", "#cell"))
    for rel in relative_positions(state):
        doc.add(use_svg(cell.to_cartesian_isometry(rel)).set("href", "#mol").set("fill", "blue"))
        for m in cell.periodic_images(rel, 1, False):
            doc.add(use_svg(m).set("href", "#mol").set("fill", "green"))
    return str(doc)


def save_svg(path, state):
    """svg::save(path, &state.as_svg()) (main.rs:266)."""
    with open(path, "w") as f:
        f.write(as_svg(state))
