// ORACLE (test infrastructure, not product code): the system handle shared by the oracle's C entry points.
#pragma once
#include "cpu_fri.hpp"
#include "cpu_verify.hpp"

struct OrcSystem {
    msh::SystemShape shape;
    // System::new tail (src/system.rs:180-196): the preprocessed commitment, made on first use
    std::unique_ptr<orc::CpuBackend> backend;
    std::unique_ptr<msh::Prover> prover;
    msh::Prover& get_prover() {
        if (!prover) {
            backend = std::make_unique<orc::CpuBackend>(shape);
            prover = std::make_unique<msh::Prover>(shape, *backend);
        }
        return *prover;
    }
};
