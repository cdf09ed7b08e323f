// lattice_stub.cpp -- see fplll/fplll.h in this directory.
#include "lattice_sample.h"

#include "errors.h"

void lattice_alpha_init(Lattice_Alpha* const, const Parameters* const) {}
void lattice_alpha_clear(Lattice_Alpha* const) {}
void lattice_alpha_map(mpz_t, mpz_t, const Lattice_Alpha* const, const Parameters* const) {
  critical("lattice_alpha_map(): sampling is not available in the fpLLL-free generator build.");
}
