// Parallel tempering over the replica ladder (new capability, no reference counterpart).
#include "system.hpp"
using namespace mcg;
extern "C" {
MCG_API int mcg_pt_swap_local(mcg_system *, int, uint64_t) { set_last_error("parallel tempering: not built yet"); return MCG_ERR_STATE; }
MCG_API int mcg_pt_energies(mcg_system *, double *) { set_last_error("parallel tempering: not built yet"); return MCG_ERR_STATE; }
MCG_API int mcg_pt_apply(mcg_system *, const double *, const double *) { set_last_error("parallel tempering: not built yet"); return MCG_ERR_STATE; }
}
