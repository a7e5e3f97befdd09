// TEST INFRASTRUCTURE.  Stand-in for <gsl/gsl_math.h>: the real header pulls in <math.h>, through which the
// reference (galileon.cc:366) reaches the global-namespace isnan/fabs/pow of its era's compilers.
#pragma once
#include <math.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
using std::isnan;
