// TEST INFRASTRUCTURE ONLY: nothing from lat/lattice-functions.h is reachable from the search code.
#include "lat/kaldi-lattice.h"
