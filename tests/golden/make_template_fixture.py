"""Generates tests/golden/highL_template.f64 from the reference's fiducial template.

Source: /root/reference/camb/HighLExtrapTemplate_lenspotentialCls.dat (read at camb/modules.f90:1222-1245;
columns L, TT, EE, BB, TE, PP, TP, EP).  Only the four columns the reference keeps (TT, EE, TE, PP) for
L = 2..8000 are stored, as little-endian float64[7999][4].  The file is DATA the reference ships as the required
input of its l-interpolation (camb/modules.f90:1046-1089); it is not source code.
Run here (the container with /root/reference); the output is committed.
"""
import os
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/camb/HighLExtrapTemplate_lenspotentialCls.dat"
a = np.loadtxt(src)
assert a[0, 0] == 2 and a[-1, 0] >= 8000
a = a[a[:, 0] <= 8000]
out = np.ascontiguousarray(a[:, [1, 2, 4, 5]], dtype="<f8")
assert out.shape == (7999, 4)
out.tofile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "highL_template.f64"))
print("wrote", out.shape)
