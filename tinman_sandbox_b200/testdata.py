"""Host-side mirror of the reference's TestData (pointers_only/data_structures.hpp:78-89): same member
names, same array layouts, same closed-form synthetic initialisation (data_structures.cpp:38-92,117-163),
same update_time_levels (data_structures.cpp:174-180). Pure host code: it prepares inputs for the CUDA
path and never computes the RHS itself.
"""
from __future__ import annotations

import math

import numpy as np

from .capi import FIELD_NAMES, field_shape

# GLL derivative matrix literals for np=4 (data_structures.cpp:150-163); Dvv[i][j] = values[j*np + i]
_DVV_VALUES = (
    -3.0000000000000000, -0.80901699437494745, 0.30901699437494745, -0.50000000000000000,
    4.0450849718747373, 0.00000000000000000, -1.11803398874989490, 1.54508497187473700,
    -1.5450849718747370, 1.11803398874989490, 0.00000000000000000, -4.04508497187473730,
    0.5000000000000000, -0.30901699437494745, 0.80901699437494745, 3.000000000000000000,
)


class TestData:
    __test__ = False  # not a pytest class

    def __init__(self, nelem, nlev=72, qsize_d=1, ntl=3, alloc=None):
        """alloc(shape) -> float64 C-contiguous array (e.g. a view of pinned memory); default np.zeros."""
        self.nelem, self.nlev, self.qsize_d, self.ntl = nelem, nlev, qsize_d, ntl
        alloc = alloc or (lambda shape: np.zeros(shape, dtype=np.float64))
        self.arrays = {n: alloc(field_shape(n, nelem, nlev, qsize_d, ntl)) for n in FIELD_NAMES}
        self.ctl = np.array([0, nelem, 0, 1, 2, 0], dtype=np.int32)  # nets nete n0 np1 nm1 qn0
        self.dt2 = 1.0
        self.consts = np.zeros(6)  # rrearth eta_ave_w cp Rwater_vapor Rgas kappa
        self.dvv = np.zeros((4, 4))
        self.ps0 = 0.0
        self.hyai = np.zeros(nlev + 1)

    # ---- data_structures.cpp:165-172 -----------------------------------------------------------
    def init_data(self, elem_offset=0):
        """Closed-form fields of the reference; elem_offset = global index of this slice's first element
        (a rank's nets in the global numbering), so that slices of a partitioned run carry the same values
        as the corresponding elements of a single-process run."""
        E, L, A = self.nelem, self.nlev, self.arrays
        ip = np.arange(1, 5, dtype=np.float64).reshape(4, 1)   # iip
        jp = np.arange(1, 5, dtype=np.float64).reshape(1, 4)   # jjp
        ie = (np.arange(E, dtype=np.float64) + 1 + elem_offset)
        il = np.arange(1, L + 1, dtype=np.float64)
        sin_ij = np.array([[math.sin(i + j) for j in range(1, 5)] for i in range(1, 5)])
        cos_i3j = np.array([[math.cos(i + 3 * j) for j in range(1, 5)] for i in range(1, 5)])
        A["elem_fcor"][...] = sin_ij
        A["elem_metdet"][...] = ip * jp
        A["elem_rmetdet"][...] = 1.0 / (ip * jp)
        A["elem_spheremp"][...] = 2 * ip + 0 * jp
        A["elem_state_phis"][...] = ip + jp
        A["elem_D"][...] = 0.0
        A["elem_D"][..., 0, 0] = 1.0
        A["elem_D"][..., 1, 1] = 2.0
        A["elem_Dinv"][...] = 0.0
        A["elem_Dinv"][..., 0, 0] = 1.0
        A["elem_Dinv"][..., 1, 1] = 0.5
        A["elem_derived_phi"][...] = cos_i3j + il.reshape(L, 1, 1)
        A["elem_derived_vn0"][...] = 1.0
        A["elem_derived_pecnd"][...] = 1.0
        A["elem_derived_omega_p"][...] = jp * jp + 0 * ip
        A["elem_derived_eta_dot_dpdn"][...] = 0.0
        qd = np.array([[[1.0 + math.sin(i * j * k) for j in range(1, 5)] for i in range(1, 5)]
                       for k in range(1, L + 1)])
        A["elem_state_Qdp"][...] = 0.0
        A["elem_state_Qdp"][:, 0, 0] = qd
        e5 = ie.reshape(E, 1, 1, 1, 1)
        t5 = np.arange(1, self.ntl + 1, dtype=np.float64).reshape(1, self.ntl, 1, 1, 1)
        l5 = il.reshape(1, 1, L, 1, 1)
        i5, j5 = ip.reshape(1, 1, 1, 4, 1), jp.reshape(1, 1, 1, 1, 4)
        # evaluation order of the reference expressions is kept (left to right) so values are bit-identical
        A["elem_state_dp3d"][...] = (((10.0 * l5 + e5) + i5) + j5) + t5
        base = (((1.0 + 0.5 * l5) + i5) + j5) + 0.2 * e5
        A["elem_state_v"][..., 0] = base + 2.0 * t5
        A["elem_state_v"][..., 1] = base + 3.0 * t5
        A["elem_state_T"][...] = ((((1000.0 - l5) - i5) - j5) + 0.1 * e5) + t5
        # Constants, Control, HVCoord, Derivative (data_structures.cpp:117-163)
        Rwv, Rgas, cp = 461.5, 287.04, 1005.0
        self.consts[:] = (1.0 / 6.376e6, 1.0, cp, Rwv, Rgas, Rgas / cp)
        self.ctl[:] = (0, E, 0, 1, 2, 0)
        self.dt2 = 1.0
        self.ps0 = 10.0
        self.hyai[:] = L + 1 - np.arange(L + 1, dtype=np.float64)
        v = np.array(_DVV_VALUES).reshape(4, 4)
        self.dvv[...] = v.T  # Dvv[i][j] = values[j*np + i]
        return self

    # ---- data_structures.cpp:174-180 -----------------------------------------------------------
    def update_time_levels(self):
        nets, nete, n0, np1, nm1, qn0 = [int(x) for x in self.ctl]
        self.ctl[:] = (nets, nete, np1, nm1, n0, qn0)
