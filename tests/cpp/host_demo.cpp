// Compiled by tests/test_cpp_host.py against include/crystal_packing.hpp + libcpgpu.so.
// argv[1] = "host": only host-side constructors (runs without a GPU).
// argv[1] = "multi" [ids..]: cp_multi over a device list against one context (needs a GPU).
// argv[1] = "gpu":  tests/packing.rs:14-52 through the C++ mirror (square, p2, MCOptimiser::new(0,0,0.001,1000,100,0,None)).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

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
        if (argc > 1 && std::strcmp(argv[1], "multi") == 0) {
            // argv[2..]: device ids (may repeat).  The same 96 replicas on one context and sharded over
            // the list must give the same best state and the same per-replica results, for all three
            // state types (packed.rs / potential.rs behind one entry point).
            std::vector<int> ids;
            for (int k = 2; k < argc; ++k) ids.push_back(std::atoi(argv[k]));
            if (ids.empty()) ids = {0, 0};
            cp::Engine one(ids[0]);
            cp::MultiEngine many(ids);
            cp::BuildOptimiser o;
            o.steps = 300;
            o.inner_steps = 100;
            auto same = [&](auto st, const char *what) {
                cp::BestPacking a, b;
                cp::monte_carlo_best_packing(st, o, 97, one, CP_RNG_PCG64MCG, &a, 5);
                cp::monte_carlo_best_packing(st, o, 97, many, CP_RNG_PCG64MCG, &b, 5);
                bool ok = std::memcmp(&a.best, &b.best, sizeof a.best) == 0 &&
                          std::memcmp(a.replicas.data(), b.replicas.data(), sizeof(cp_replica_result) * 97) == 0;
                std::printf("multi %s devices %d best %.17g replica %llu %s\n", what, many.devices(), b.best.result.score,
                            (unsigned long long)b.best.replica, ok ? "same" : "DIFFERENT");
                return ok;
            };
            bool ok = same(cp::PackedState::from_group(cp::LineShape::polygon(5), "p2"), "polygon");
            ok &= same(cp::PackedState::from_group(cp::MolecularShape2::from_trimer(0.637556, 120., 1.), "p2gg"), "trimer");
            ok &= same(cp::PotentialState::from_group(cp::LJShape2::circle(), "p1"), "lj");
            std::uint64_t begin = 0, count = 0;
            cp::check(cp_multi_shard(many.raw(), many.devices() - 1, &begin, &count));
            std::printf("last shard %llu +%llu\n", (unsigned long long)begin, (unsigned long long)count);
            return ok ? 0 : 4;
        }
        return 0;
    } catch (const cp::Error &e) {
        std::printf("cp::Error %d: %s\n", (int)e.status, e.what());
        return 1;
    }
}
