/* fplll/fplll.h -- compile-only stand-in for fpLLL.
 *
 * INTEGRATION-TEST INFRASTRUCTURE. The reference pulls fpLLL into every generator through
 * distribution.h -> lattice_sample.h (two matrix members of Lattice_Alpha), although generation
 * never reduces a lattice. The generators are built here without fpLLL:
 * this header gives the two matrix types a trivial definition and integration/stubs/lattice_stub.cpp
 * provides lattice_alpha_init/clear/map that do nothing (init/clear) or abort (map, sampling only).
 */
#ifndef QB200_FPLLL_STUB_H
#define QB200_FPLLL_STUB_H

namespace fplll {
template <class T>
class ZZ_mat {};
template <class T>
class FP_mat {};
}  // namespace fplll

#endif
