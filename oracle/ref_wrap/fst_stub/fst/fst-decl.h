// TEST INFRASTRUCTURE ONLY: the forward declarations nnet3 / hmm headers ask <fst/fst-decl.h> for, served by the same
// container-only stand-in as fst/fstlib.h so that a translation unit can include both the decoder headers and the nnet3
// headers (oracle/check_shims.py).
#ifndef B2K_FST_STUB_FST_DECL_H_
#define B2K_FST_STUB_FST_DECL_H_
#include "fst/fstlib.h"
#endif
