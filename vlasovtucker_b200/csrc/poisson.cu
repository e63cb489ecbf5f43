// Poisson path (src/poisson.cpp) — device implementation.
#include "vt_internal.h"

namespace vt {
struct PoissonData {
    int dummy;
};
void poisson_destroy(PoissonData* p) { delete p; }
}  // namespace vt

extern "C" {
int vt_poisson_setup(vt_ctx*, const double*, const double*, const uint8_t*, const double*, const double*) { return 1; }
int vt_poisson_update_bc_values(vt_ctx*, const double*, const double*) { return 1; }
int vt_poisson_solve(vt_ctx*, const double*, double*, double*) { return 1; }
int vt_poisson_stats(vt_ctx*, int*, double*) { return 1; }
}
