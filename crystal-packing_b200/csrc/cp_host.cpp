// cp_host.cpp — host-side constructors of libcpgpu.so: shapes, plane groups, the symmetry-string
// parser, initial states and annealing schedules.  These run once per problem on the CPU in the
// reference too; nothing here is on the Monte-Carlo loop.  Compiled with -ffp-contract=off so the
// constants handed to the kernels carry the roundings the reference's own constructors produce.
//
// Reference citations are relative to the reference tree (crystal-packing/src/... unless noted).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "cp_internal.h"

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;  // std::f64::consts::PI

double origin_distance(double x, double y) { return std::sqrt(x * x + y * y); }

// f64::max as used by the enclosing_radius folds: NaN operands are ignored
double fold_max(double acc, double v) { return std::fmax(acc, v); }

// MolecularShape2::overlap_area (shape/molecular_shape2.rs:103-105)
double lens_half(double r, double d) { return (r * r) * std::acos(d / r) - d * std::sqrt((r * r) - (d * d)); }

// MolecularShape2::circle_overlap (shape/molecular_shape2.rs:107-117)
double circle_overlap(const double *a1, const double *a2) {
    double ddx = a1[0] - a2[0], ddy = a1[1] - a2[1];
    double dist = std::sqrt(ddx * ddx + ddy * ddy);
    if (dist < a1[2] + a2[2]) {
        double d1 = (dist * dist + a1[2] * a1[2] - a2[2] * a2[2]) / (2. * dist);
        double d2 = (dist * dist + a2[2] * a2[2] - a1[2] * a1[2]) / (2. * dist);
        return lens_half(a1[2], d1) + lens_half(a2[2], d2);
    }
    return 0.;
}

void finish_shape(cp_problem *p) {
    const int n = p->n_items;
    double radius = -1.7976931348623157e308;  // f64::MIN seed of the fold
    for (int i = 0; i < n; ++i) {
        double v = origin_distance(p->items[i][0], p->items[i][1]);
        if (p->kind == CP_DISCS) v = v + p->items[i][2];          // molecular_shape2.rs:66-74
        if (p->kind == CP_LJ) v = v + p->items[i][2] / 2.;        // lj_shape.rs:52-60
        radius = fold_max(radius, v);                             // line_shape.rs:86-93
    }
    p->enclosing_radius = radius;
    p->area = 0.;
    if (p->kind == CP_POLYGON) {
        // line_shape.rs:60-72
        double angle_term = std::sin(2. * kPi / static_cast<double>(n));
        double sum = 0.;
        for (int i = 0; i < n; ++i)
            sum += 0.5 * angle_term * origin_distance(p->items[i][0], p->items[i][1]) *
                   origin_distance(p->items[i][2], p->items[i][3]);
        p->area = sum;
    } else if (p->kind == CP_DISCS) {
        // molecular_shape2.rs:39-52 (pairwise lens subtraction only)
        double total = 0., overlap = 0.;
        for (int i = 0; i < n; ++i) total += kPi * (p->items[i][2] * p->items[i][2]);
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) overlap += circle_overlap(p->items[i], p->items[j]);
        p->area = total - overlap;
    }
}

struct GroupRow {
    const char *name;
    int family;
    std::vector<const char *> ops;
};

const std::vector<GroupRow> &group_table() {
    // wallpaper.rs:134-176; "pg" is the parser's alias of p1g1 (wallpaper.rs:100-110)
    static const std::vector<GroupRow> rows = {
        {"p1", CP_MONOCLINIC, {"x,y"}},
        {"p2", CP_MONOCLINIC, {"x,y", "-x,-y"}},
        {"p1m1", CP_ORTHORHOMBIC, {"x,y", "-x,y"}},
        {"p1g1", CP_ORTHORHOMBIC, {"x,y", "-x,y+1/2"}},
        {"pg", CP_ORTHORHOMBIC, {"x,y", "-x,y+1/2"}},
        {"p2mm", CP_ORTHORHOMBIC, {"x,y", "-x,-y", "-x,y", "x,-y"}},
        {"p2mg", CP_ORTHORHOMBIC, {"x,y", "-x, -y", "-x+1/2, y", "x+1/2, -y"}},
        {"p2gg", CP_ORTHORHOMBIC, {"x,y", "-x, -y", "-x+1/2, y+1/2", "x+1/2, -y+1/2"}},
    };
    return rows;
}

}  // namespace

extern "C" {

cp_status cp_shape_polygon(const double *radii, int n, cp_problem *out) {
    // LineShape::from_radial (shape/line_shape.rs:128-146); radii == NULL -> polygon(n) (:148-150)
    if (!out) return cp_fail(CP_ERR_INVALID_ARG, "cp_shape_polygon: null output");
    if (n < 3)
        return cp_fail(CP_ERR_INVALID_ARG,
                       "The number of points provided is too few to create a 2D shape.");
    if (n > CP_MAX_ITEMS) return cp_fail(CP_ERR_INVALID_ARG, "cp_shape_polygon: more than CP_MAX_ITEMS sides");
    std::memset(out, 0, sizeof *out);
    out->kind = CP_POLYGON;
    out->n_items = n;
    const double dtheta = 2. * kPi / static_cast<double>(n);
    for (int index = 0; index < n; ++index) {
        double r1 = radii ? radii[index] : 1., r2 = radii ? radii[(index + 1) % n] : 1.;
        double angle = static_cast<double>(index) * dtheta;
        out->items[index][0] = r1 * std::sin(angle);
        out->items[index][1] = r1 * std::cos(angle);
        out->items[index][2] = r2 * std::sin(angle + dtheta);
        out->items[index][3] = r2 * std::cos(angle + dtheta);
    }
    finish_shape(out);
    return CP_OK;
}

cp_status cp_shape_trimer(int kind, double radius, double angle, double distance, cp_problem *out) {
    if (!out) return cp_fail(CP_ERR_INVALID_ARG, "cp_shape_trimer: null output");
    if (kind != CP_DISCS && kind != CP_LJ)
        return cp_fail(CP_ERR_INVALID_ARG, "cp_shape_trimer: kind must be CP_DISCS or CP_LJ");
    std::memset(out, 0, sizeof *out);
    out->kind = kind;
    out->n_items = 3;
    const double half = angle * (kPi / 180.) / 2.;  // angle.to_radians() / 2.
    if (kind == CP_LJ) {
        // LJShape2::from_trimer (shape/lj_shape.rs:111-131)
        double x_base = distance * std::sin(half);
        double y_base = 1. / 3. * distance * std::cos(half);
        const double px[3] = {0., -x_base, x_base};
        const double py[3] = {-2. * y_base, y_base, y_base};
        const double rr[3] = {1., radius, radius};
        for (int i = 0; i < 3; ++i) {
            out->items[i][0] = px[i];
            out->items[i][1] = py[i];
            out->items[i][2] = 2. * rr[i];  // sigma
            out->items[i][3] = 1.;          // epsilon (LJ2::default, lj2.rs:31-40)
            out->items[i][4] = 3.5;         // cutoff Some(3.5)
        }
    } else {
        // MolecularShape2::from_trimer (shape/molecular_shape2.rs:141-162)
        out->items[0][0] = 0.;
        out->items[0][1] = -2. / 3. * distance * std::cos(half);
        out->items[0][2] = 1.;
        out->items[1][0] = -distance * std::sin(half);
        out->items[1][1] = 1. / 3. * distance * std::cos(half);
        out->items[1][2] = radius;
        out->items[2][0] = distance * std::sin(half);
        out->items[2][1] = 1. / 3. * distance * std::cos(half);
        out->items[2][2] = radius;
    }
    finish_shape(out);
    return CP_OK;
}

cp_status cp_shape_circle(int kind, cp_problem *out) {
    if (!out) return cp_fail(CP_ERR_INVALID_ARG, "cp_shape_circle: null output");
    if (kind != CP_DISCS && kind != CP_LJ)
        return cp_fail(CP_ERR_INVALID_ARG, "cp_shape_circle: kind must be CP_DISCS or CP_LJ");
    std::memset(out, 0, sizeof *out);
    out->kind = kind;
    out->n_items = 1;
    if (kind == CP_LJ) {
        // LJShape2::circle -> LJ2::new(0,0,1) (lj_shape.rs:146-151, lj2.rs:74-80)
        out->items[0][2] = 1.;
        out->items[0][3] = 1.;
        out->items[0][4] = std::nan("");
    } else {
        out->items[0][2] = 1.;  // Atom2::new(0,0,1) (molecular_shape2.rs:167-172)
    }
    finish_shape(out);
    return CP_OK;
}

cp_status cp_shape_finish(cp_problem *shape) {
    if (!shape) return cp_fail(CP_ERR_INVALID_ARG, "cp_shape_finish: null shape");
    if (shape->kind < CP_POLYGON || shape->kind > CP_LJ || shape->n_items < 1 || shape->n_items > CP_MAX_ITEMS)
        return cp_fail(CP_ERR_INVALID_ARG, "cp_shape_finish: kind / n_items out of range");
    finish_shape(shape);
    return CP_OK;
}

cp_status cp_parse_operations(const char *sym_ops, double out[6]) {
    // Transform2::from_operations (transform.rs:124-183).  Quirks kept: single digits only, '*'
    // behaves like '/', a constant's sign comes from the last '-' seen.
    if (!sym_ops || !out) return cp_fail(CP_ERR_INVALID_ARG, "cp_parse_operations: null argument");
    std::string s(sym_ops);
    size_t b = s.find_first_not_of("()"), e = s.find_last_not_of("()");
    s = (b == std::string::npos) ? std::string() : s.substr(b, e - b + 1);
    std::vector<std::string> pieces;
    size_t start = 0;
    while (start <= s.size()) {
        size_t comma = s.find(',', start);
        if (comma == std::string::npos) {
            if (start < s.size()) pieces.push_back(s.substr(start));  // split_terminator drops a trailing ""
            break;
        }
        pieces.push_back(s.substr(start, comma - start));
        start = comma + 1;
    }
    if (pieces.size() < 2) return cp_fail(CP_ERR_PARSE, "Not enough dimensions in input");
    if (pieces.size() > 2) return cp_fail(CP_ERR_PARSE, "Too many dimensions in input");
    for (int i = 0; i < 6; ++i) out[i] = 0.;
    for (int row = 0; row < 2; ++row) {
        double sign = 1., constant = 0.;
        char pending = 0;
        for (char c : pieces[row]) {
            switch (c) {
            case 'x': out[row * 3 + 0] = sign; sign = 1.; break;
            case 'y': out[row * 3 + 1] = sign; sign = 1.; break;
            case '*': case '/': pending = c; break;
            case '-': sign = -1.; break;
            case ' ': case '+': break;
            default:
                if (c >= '0' && c <= '9') {
                    double val = static_cast<double>(c - '0');
                    constant = pending ? sign * constant / val : sign * val;
                    pending = 0;
                    sign = 1.;
                } else {
                    return cp_fail(CP_ERR_PARSE, (std::string("Found invalid value: '") + c + "'").c_str());
                }
            }
        }
        out[row * 3 + 2] = constant;
    }
    return CP_OK;
}

cp_status cp_group_lookup(const char *name, int32_t *family, int32_t *n_syms,
                          double syms[CP_MAX_SYMS][6]) {
    if (!name || !family || !n_syms || !syms) return cp_fail(CP_ERR_INVALID_ARG, "cp_group_lookup: null argument");
    for (const GroupRow &row : group_table()) {
        if (std::strcmp(row.name, name) != 0) continue;
        *family = row.family;
        *n_syms = static_cast<int32_t>(row.ops.size());
        for (size_t i = 0; i < row.ops.size(); ++i) {
            cp_status st = cp_parse_operations(row.ops[i], syms[i]);
            if (st != CP_OK) return st;
        }
        return CP_OK;
    }
    return cp_fail(CP_ERR_PARSE, "Invalid Value");  // WallpaperGroups::from_str (wallpaper.rs:109)
}

cp_status cp_problem_from_group(cp_problem *p, const char *group) {
    if (!p || !group) return cp_fail(CP_ERR_INVALID_ARG, "cp_problem_from_group: null argument");
    if (p->n_items < 1 || p->n_items > CP_MAX_ITEMS)
        return cp_fail(CP_ERR_INVALID_ARG, "cp_problem_from_group: shape fields are not set");
    cp_status st = cp_group_lookup(group, &p->family, &p->n_syms, p->syms);
    if (st != CP_OK) return st;
    // initialise (packed.rs:169-189 / potential.rs:154-174): cell side 4*R*N (hard) or 2*R*N (LJ)
    const double num_shapes = static_cast<double>(p->n_syms);
    const double max_cell_size = (p->kind == CP_LJ ? 2. : 4.) * p->enclosing_radius * num_shapes;
    p->dof[0] = max_cell_size;                                        // Cell2::from_family (cell.rs:200-214)
    p->dof[1] = 1.;
    p->dof[2] = p->family == CP_HEXAGONAL ? kPi / 3. : kPi / 2.;
    const double position = -0.5 + 0.5 / static_cast<double>(p->n_syms);  // site.rs:51-63
    p->dof[3] = position;
    p->dof[4] = position;
    p->dof[5] = 0.;
    return CP_OK;
}

cp_status cp_build_schedule(const cp_build_optimiser *o, cp_schedule *out) {
    // BuildOptimiser::build (crystal-packing-cli/src/main.rs:122-143)
    if (!o || !out) return cp_fail(CP_ERR_INVALID_ARG, "cp_build_schedule: null argument");
    double kt_ratio;
    if (!std::isnan(o->kt_ratio))
        kt_ratio = 1. - o->kt_ratio;
    else if (!std::isnan(o->kt_finish))
        kt_ratio = std::pow(o->kt_finish / o->kt_start, 1. / static_cast<double>(o->steps));
    else
        kt_ratio = 0.1;
    out->kt_start = o->kt_start;
    out->kt_ratio = kt_ratio;
    out->max_step_size = o->max_step_size;
    out->steps = o->steps;
    out->inner_steps = o->inner_steps < o->steps ? o->inner_steps : o->steps;
    out->convergence = o->convergence;
    return CP_OK;
}

cp_status cp_pipeline_schedules(const cp_build_optimiser *o, cp_schedule out[3]) {
    // analyse_state (crystal-packing-cli/src/main.rs:224-252)
    if (!o || !out) return cp_fail(CP_ERR_INVALID_ARG, "cp_pipeline_schedules: null argument");
    cp_build_optimiser quick = *o;  // .steps(100).kt_start(0.).convergence(None)
    quick.steps = 100;
    quick.kt_start = 0.;
    quick.convergence = std::nan("");
    cp_build_optimiser last = *o;  // .kt_start(0.)
    last.kt_start = 0.;
    cp_status st = cp_build_schedule(&quick, &out[0]);
    if (st == CP_OK) st = cp_build_schedule(o, &out[1]);
    if (st == CP_OK) st = cp_build_schedule(&last, &out[2]);
    return st;
}

}  // extern "C"
