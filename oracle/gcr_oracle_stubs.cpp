// temporary: solvers not restated yet (replaced as the oracle grows)
#include "gcr_oracle.h"
extern "C" {
int gcro_solve_e5_nister(const double *, const uint32_t *, double *) { return 0; }
int gcro_solve_e5_stewenius(const double *, const uint32_t *, int, const double *, double *) { return 0; }
int gcro_estimate_e5(const double *, const uint32_t *, double *) { return 0; }
void gcro_normalize_correspondences(const double *, int, const double *, const double *, double *) {}
int gcro_solve_p3p(const double *, const uint32_t *, double *) { return 0; }
}
