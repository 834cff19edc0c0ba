// TEST INFRASTRUCTURE ONLY: see fst/fstlib.h in this directory (MemoryPool lives there).
#include "fst/fstlib.h"
