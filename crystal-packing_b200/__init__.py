"""crystal-packing_b200 — B200-native engine for crystal-packing's Monte-Carlo packing optimisation.

The product is `lib/libcpgpu.so` (hand-written sm_100a kernels behind the C ABI declared in
`include/cp_gpu.h`).  This Python layer is the host-side mirror of the reference's interface for
that path — same names, argument meaning and error behaviour as the Rust crate
(`PackedState::from_group`, `State::score`, `MCOptimiser::new(..).optimise_state`, `BuildOptimiser`,
and the GPU drop-in for the CLI's `analyse_state`, `monte_carlo_best_packing`).  Every compute call
goes through the C ABI to the GPU; there is no CPU fallback and importing this package never
touches `oracle/`.

The directory name carries a hyphen, so import it as `crystal_packing_b200` through the loader
module of that name at the repository root.
"""
from .api import (  # noqa: F401
    CP_DISCS,
    CP_LJ,
    CP_POLYGON,
    RNG_PCG64MCG,
    RNG_PHILOX,
    BuildOptimiser,
    CrystalPackingError,
    Engine,
    LJShape2,
    LineShape,
    MCOptimiser,
    MultiEngine,
    MolecularShape2,
    PackedState,
    PackedState2,
    PotentialState,
    PotentialState2,
    Transform2,
    WallpaperGroup,
    WallpaperGroups,
    default_engine,
    library_path,
    load_library,
    monte_carlo_best_packing,
)
from . import api, capi  # noqa: F401
