#include "structured.hpp"
namespace mcg {
struct StructuredSystem { int dummy; };
void structured_create(mcg_system *, const mcg_lattice_desc *) { throw Error(MCG_ERR_ARG, "structured path not built yet"); }
void structured_destroy(StructuredSystem *st) { delete st; }
void structured_init_spins(mcg_system *, double) {}
void structured_set_spins(mcg_system *, int, const double *) {}
void structured_get_spins(mcg_system *, int, double *) {}
void structured_measure_sums(mcg_system *) {}
void structured_sweeps(mcg_system *, int64_t, double, bool) {}
void structured_colour_order(const mcg_system *, int32_t *) {}
}  // namespace mcg
