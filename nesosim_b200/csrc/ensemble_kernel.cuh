// ensemble_kernel.cuh -- season-resident path for small grids (placeholder until the kernel lands).
#pragma once
namespace nesosim {
struct EnsembleState {
    bool derived_valid = false;
};
inline void ensemble_release(EnsembleState &) {}
}  // namespace nesosim
