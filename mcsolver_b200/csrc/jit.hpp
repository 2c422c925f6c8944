// Runtime specialisation of the colour-pass kernel (NVRTC): see struct_pass.cuh.
#pragma once
#include "structured.hpp"

namespace mcg {
struct StructArgs;
// Launch the JIT-specialised pass kernel of `colour` if JIT is enabled for this system and the
// specialisation compiled; returns false when the caller should launch the offline kernel instead.
bool jit_launch_pass(mcg_system *s, int colour, int mode, const StructArgs &a, int q0, int rowsPerBlock, int nrb, uint64_t sweep,
                     double pAtt, dim3 grid, dim3 block);
}  // namespace mcg
