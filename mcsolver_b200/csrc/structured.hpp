// Structured (descriptor-driven) path: entry points used by engine.cu; implemented in structured.cu.
#pragma once
#include "system.hpp"

namespace mcg {
void structured_create(mcg_system *s, const mcg_lattice_desc *d);
void structured_destroy(StructuredSystem *st);
void structured_init_spins(mcg_system *s, double flunc);
void structured_set_spins(mcg_system *s, int r, const double *spins);
void structured_get_spins(mcg_system *s, int r, double *spins);
void structured_measure_sums(mcg_system *s);
// n colour sweeps; fusedMeasure: the last sweep also produces the raw measurement sums in d_sums
void structured_sweeps(mcg_system *s, int64_t n, double pAtt, bool fusedMeasure);
void structured_colour_order(const mcg_system *s, int32_t *order);
void structured_rng_layout(const mcg_system *s, int32_t *stride, int32_t *group);
struct WolffArgs;
int structured_wolff_step(mcg_system *s, const WolffArgs &w, bool primed, bool needResidual, int force);   // returns kernels launched
uint64_t structured_jit_key(const mcg_system *s, int colour);   // cache key (= cubin file name) of a colour's specialised module
int structured_jit_check(const mcg_lattice_desc *d, int precision, std::string &report);
// slab decomposition of one lattice along its first axis (structured.cu, last section)
void structured_create_slab(mcg_system *s, const mcg_lattice_desc *global, int rank, int world, const char *commId);
void structured_slab_plan(const mcg_lattice_desc *global, int precision, int rank, int world, int32_t *info);
void structured_slab_exchange_all(mcg_system *s);
bool structured_is_slab(const mcg_system *s);
void structured_slab_info(const mcg_system *s, int32_t *info);
}  // namespace mcg
