"""`packing` — the reference CLI's option surface (crystal-packing-cli/src/main.rs:27-64,154-213) on
top of the GPU engine; SURVEY.md §8(f) row 2.

    python -m crystal_packing_b200.cli p2 --outfile out --replications 1000 -s 10000 polygon --sides 5

Same positional group, shape sub-commands (`polygon --sides`, `trimer -d -a -r`, `circle`),
`-p/--potential {LJ,Hard}`, `--replications`, `-s/--steps`, `--kt-start`, `--kt-finish`,
`--kt-ratio`, `--max-step-size`, `--inner-steps`, `--convergence`, `--outfile`, `--start-config`
(read here; the reference declares it and ignores it).  Writes `<outfile>.json` in the reference's
serde layout and `<outfile>.svg` (main.rs:263-266; render.py).  Extra:
`--rng {pcg,philox}` (pcg reproduces the reference's random stream) and `--device`.
"""
import argparse
import sys


def build_parser():
    ap = argparse.ArgumentParser(prog="packing")
    ap.add_argument("-v", "--verbosity", action="count", default=0)
    # WallpaperGroups::from_str accepts exactly these (wallpaper.rs:100-110); `variants()` advertises
    # "p2m1", which the reference's own parser rejects
    ap.add_argument("wallpaper", help="p1 p2 p1m1 p1g1 pg p2mm p2mg p2gg")
    ap.add_argument("-p", "--potential", choices=["LJ", "Hard"], default="Hard")
    ap.add_argument("--outfile", required=True)
    ap.add_argument("--start-config")
    ap.add_argument("--replications", type=int, default=100)
    ap.add_argument("-s", "--steps", type=int, default=100)
    ap.add_argument("--kt-start", type=float, default=0.1)
    ap.add_argument("--kt-finish", type=float)
    ap.add_argument("--kt-ratio", type=float)
    ap.add_argument("--max-step-size", type=float, default=0.01)
    ap.add_argument("--inner-steps", type=int, default=1000)
    ap.add_argument("--convergence", type=float)
    ap.add_argument("--rng", choices=["pcg", "philox"], default="pcg")
    ap.add_argument("--device", type=int, default=0)
    sub = ap.add_subparsers(dest="shape", required=True)
    poly = sub.add_parser("polygon")
    poly.add_argument("--sides", type=int, default=4)
    tri = sub.add_parser("trimer")
    tri.add_argument("-d", "--distance", type=float, default=1.0)
    tri.add_argument("-a", "--angle", type=float, default=120.0)
    tri.add_argument("-r", "--radius", type=float, default=0.637556)
    sub.add_parser("circle")
    return ap


def make_state(args):
    from . import api

    group = api.WallpaperGroups.from_str(args.wallpaper)
    lj = args.potential == "LJ"
    if args.shape == "polygon":
        if lj:  # main.rs:330-332
            raise SystemExit("Error: Polygon with a LJ potential is not yet implemented")
        return api.PackedState.from_group(api.LineShape.polygon(args.sides), group)
    if args.shape == "trimer":
        if lj:
            return api.PotentialState.from_group(api.LJShape2.from_trimer(args.radius, args.angle, args.distance), group)
        return api.PackedState.from_group(api.MolecularShape2.from_trimer(args.radius, args.angle, args.distance), group)
    if lj:
        return api.PotentialState.from_group(api.LJShape2.circle(), group)
    return api.PackedState.from_group(api.MolecularShape2.circle(), group)


def make_optimiser(args):
    from . import api

    return api.BuildOptimiser(steps=args.steps, kt_start=args.kt_start, kt_finish=args.kt_finish,
                              kt_ratio=args.kt_ratio, max_step_size=args.max_step_size,
                              inner_steps=args.inner_steps, convergence=args.convergence)


def main(argv=None):
    from . import api

    args = build_parser().parse_args(argv)
    state = make_state(args)
    if args.start_config:
        loaded = type(state).from_json(open(args.start_config).read())
        state.dof = loaded.dof
    engine = api.Engine(args.device)
    rng = api.RNG_PCG64MCG if args.rng == "pcg" else api.RNG_PHILOX
    best = api.monte_carlo_best_packing(state, make_optimiser(args), args.replications, engine=engine, rng=rng)
    print(f"Final score: {best.best_score}", file=sys.stderr)  # info!("Final score: {}") (main.rs:256-261)
    out = args.outfile[:-5] if args.outfile.endswith(".json") else args.outfile
    with open(out + ".json", "w") as f:
        f.write(best.to_json())
    with open(out + ".svg", "w") as f:  # svg::save(outfile.with_extension("svg"), ..) (main.rs:266)
        f.write(best.as_svg())
    return 0


if __name__ == "__main__":
    sys.exit(main())
