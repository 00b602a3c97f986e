// Compiled by tests/test_cpp_host.py against include/crystal_packing.hpp + libcpgpu.so.
// argv[1] = "host": only host-side constructors (runs without a GPU).
// argv[1] = "gpu":  tests/packing.rs:14-52 through the C++ mirror (square, p2, MCOptimiser::new(0,0,0.001,1000,100,0,None)).
#include <cstdio>
#include <cstring>

#include "crystal_packing.hpp"

int main(int argc, char **argv) {
    try {
        cp::LineShape square = cp::LineShape::from_radial("Square", {1., 1., 1., 1.});
        cp::PackedState state = cp::PackedState::from_group(square, "p2");
        std::printf("area %a radius %a shapes %d dof", square.area(), square.enclosing_radius(), state.total_shapes());
        for (double v : state.problem.dof) std::printf(" %a", v);
        std::printf(" basis %zu\n", state.generate_basis().size());
        cp::BuildOptimiser opts;
        opts.steps = 100;
        cp::MCOptimiser m = opts.build(7);
        std::printf("kt_ratio %a inner %llu\n", m.schedule.kt_ratio, (unsigned long long)m.schedule.inner_steps);
        try {
            cp::LineShape::from_radial("x", {1., 1.});
            return 2;
        } catch (const cp::Error &e) {
            std::printf("error %d %s\n", (int)e.status, e.what());
        }
        if (argc > 1 && std::strcmp(argv[1], "gpu") == 0) {
            cp::Engine eng(0);
            double init = *state.score(eng);
            cp::PackedState fin = cp::MCOptimiser(0., 0., 0.001, 1000, 100, 0, std::nullopt).optimise_state(state, eng);
            double final_score = *fin.score(eng);
            std::printf("init %.17g final %.17g\n", init, final_score);
            cp::BuildOptimiser o2;
            o2.steps = 300;
            o2.inner_steps = 100;
            cp::BestPacking d;
            cp::PackedState best = cp::monte_carlo_best_packing(state, o2, 32, eng, CP_RNG_PCG64MCG, &d);
            std::printf("best %.17g replica %llu\n", d.best.result.score, (unsigned long long)d.best.replica);
            return init < final_score ? 0 : 3;
        }
        return 0;
    } catch (const cp::Error &e) {
        std::printf("cp::Error %d: %s\n", (int)e.status, e.what());
        return 1;
    }
}
