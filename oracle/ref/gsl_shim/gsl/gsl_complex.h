// TEST INFRASTRUCTURE.  Empty stand-in: camb/galileon.cc includes <gsl/gsl_complex.h> but uses nothing from it.
#pragma once
#include <cmath>
