// TEST INFRASTRUCTURE.  Minimal stand-in for <gsl/gsl_errno.h> (status codes only).
#pragma once
enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_EINVAL = 4, GSL_EBADFUNC = 9 };
