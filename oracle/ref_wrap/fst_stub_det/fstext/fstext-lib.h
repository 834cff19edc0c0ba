// TEST INFRASTRUCTURE ONLY (oracle/_ref determinizer build): lat/push-lattice.h and lat/minimize-lattice.h include the
// whole fstext library; their .cc files need OpenFst containers, TopSort and Connect only (the stand-in's fst/fstlib.h).
#ifndef B2K_ORACLE_FST_STUB_DET_FSTEXT_LIB_H_
#define B2K_ORACLE_FST_STUB_DET_FSTEXT_LIB_H_
#include "fst/fstlib.h"
#include "fstext/fstext-utils.h"
#include "fstext/lattice-weight.h"
#endif
