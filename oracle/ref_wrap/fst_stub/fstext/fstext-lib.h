// TEST INFRASTRUCTURE ONLY: the reference decoder needs nothing from fstext beyond OpenFst itself.
#include "fst/fstlib.h"
