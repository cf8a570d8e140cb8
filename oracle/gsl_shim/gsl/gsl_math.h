/* oracle GSL shim (test infrastructure): see gsl_shim_all.h */
#include "gsl_shim_all.h"
