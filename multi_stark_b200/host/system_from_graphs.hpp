// A system assembled from compiled circuits handed over as msgpu_graph_desc (include/msgpu.h): the generic entry point for
// a caller holding its own `CircuitSpec` / `ConstraintGraph` (reference src/system.rs:115-203, src/graph.rs:62-76) instead
// of one of the named benchmark systems.
#pragma once
#include "program.hpp"
#include "system.hpp"

namespace msh {

// preprocessed[i]: HOST pointer to pre_heights[i] x descs[i].pre_width canonical values, or NULL when circuit i has none.
inline SystemShape system_from_descs(const CommitmentParameters& cp, const FriParameters& fp, const msgpu_graph_desc* descs, uint32_t n,
                                     const uint64_t* const* preprocessed, const uint64_t* pre_heights) {
    if (n == 0 || !descs) throw std::runtime_error("a system needs at least one circuit");
    std::vector<SystemShape::CompiledInput> inputs;
    for (uint32_t i = 0; i < n; i++) {
        const msgpu_graph_desc& d = descs[i];
        SystemShape::CompiledInput in;
        in.graph = graph_from_desc(d);
        in.main_width = d.main_width;
        const size_t want_s2 = std::max<size_t>(in.graph.lookups.size(), 1) * ExtensionParams().degree;
        if (d.stage2_width != want_s2) throw std::runtime_error("graph descriptor: stage2_width must be max(lookups, 1) * 2");
        if (d.pre_width) {
            if (!preprocessed || !preprocessed[i] || !pre_heights || pre_heights[i] == 0)
                throw std::runtime_error("graph descriptor: circuit with preprocessed columns needs its preprocessed trace");
            in.has_preprocessed = true;
            in.preprocessed = Matrix((size_t)pre_heights[i], d.pre_width);
            for (size_t k = 0; k < in.preprocessed.values.size(); k++) {
                if (preprocessed[i][k] >= GL_P) throw std::runtime_error("preprocessed value is not canonical");
                in.preprocessed.values[k] = Fp(preprocessed[i][k]);
            }
        }
        inputs.push_back(std::move(in));
    }
    return SystemShape::build_compiled(cp, fp, std::move(inputs));
}

}  // namespace msh
