"""oracle/harness.py — TEST INFRASTRUCTURE ONLY.

ctypes loaders for the two CPU oracles of the compute_and_apply_rhs path:

  * ``RefOracle(nlev)``  — the REAL reference (compute_and_apply_rhs_test/cxx/pointers_only/*.cpp)
    compiled by oracle/Makefile into oracle/_ref/libcaar_ref_L<nlev>.so (nlev 72 or 128 only: PLEV is
    compile-time in the reference, config.h.in:3).
  * ``PortOracle()``     — our plain-C restatement oracle/caar_oracle.c (any nlev >= 2).

Both expose ``init(nelem)`` (the reference's closed-form synthetic data, PO/data_structures.cpp:38-92),
``run(state, ncalls, nthreads)`` (PO/compute_and_apply_rhs.cpp:15-278) and ``norms(state)``
(PO/compute_and_apply_rhs.cpp:372-399) on a ``State`` of numpy arrays in the reference's host layout.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

FIELD_NAMES = (
    "elem_D", "elem_Dinv", "elem_fcor", "elem_spheremp", "elem_metdet", "elem_rmetdet",
    "elem_state_dp3d", "elem_state_v", "elem_state_T", "elem_state_phis", "elem_state_Qdp",
    "elem_derived_eta_dot_dpdn", "elem_derived_omega_p", "elem_derived_phi", "elem_derived_pecnd",
    "elem_derived_vn0",
)
MUTATED = ("elem_state_dp3d", "elem_state_v", "elem_state_T", "elem_derived_eta_dot_dpdn",
           "elem_derived_omega_p", "elem_derived_phi", "elem_derived_vn0")


def field_shape(name: str, E: int, L: int, Q: int = 1, ntl: int = 3):
    """Host shapes, PO/data_structures.cpp:14-31."""
    return {
        "elem_D": (E, 4, 4, 2, 2), "elem_Dinv": (E, 4, 4, 2, 2),
        "elem_fcor": (E, 4, 4), "elem_spheremp": (E, 4, 4), "elem_metdet": (E, 4, 4),
        "elem_rmetdet": (E, 4, 4), "elem_state_phis": (E, 4, 4),
        "elem_state_dp3d": (E, ntl, L, 4, 4), "elem_state_T": (E, ntl, L, 4, 4),
        "elem_state_v": (E, ntl, L, 4, 4, 2),
        "elem_state_Qdp": (E, Q, 2, L, 4, 4),
        "elem_derived_eta_dot_dpdn": (E, L + 1, 4, 4),
        "elem_derived_omega_p": (E, L, 4, 4), "elem_derived_phi": (E, L, 4, 4),
        "elem_derived_pecnd": (E, L, 4, 4), "elem_derived_vn0": (E, L, 4, 4, 2),
    }[name]


@dataclass
class State:
    """A TestData (PO/data_structures.hpp:78-89) as numpy arrays + scalars."""
    nelem: int
    nlev: int
    qsize_d: int = 1
    ntl: int = 3
    arrays: dict = field(default_factory=dict)
    # Control (nets, nete, n0, np1, nm1, qn0) + dt2
    ctl: np.ndarray = None
    dt2: float = 1.0
    # Constants (rrearth, eta_ave_w, cp, Rwater_vapor, Rgas, kappa)
    consts: np.ndarray = None
    dvv: np.ndarray = None      # (4,4) row-major, Dvv[i][j]
    ps0: float = 10.0
    hyai: np.ndarray = None     # (nlev+1,)

    @classmethod
    def empty(cls, nelem, nlev, qsize_d=1, ntl=3):
        s = cls(nelem, nlev, qsize_d, ntl)
        for n in FIELD_NAMES:
            s.arrays[n] = np.zeros(field_shape(n, nelem, nlev, qsize_d, ntl), dtype=np.float64)
        s.ctl = np.zeros(6, dtype=np.int32)
        s.consts = np.zeros(6, dtype=np.float64)
        s.dvv = np.zeros((4, 4), dtype=np.float64)
        s.hyai = np.zeros(nlev + 1, dtype=np.float64)
        return s

    def copy(self):
        s = State(self.nelem, self.nlev, self.qsize_d, self.ntl)
        s.arrays = {k: v.copy() for k, v in self.arrays.items()}
        s.ctl = self.ctl.copy()
        s.dt2 = self.dt2
        s.consts = self.consts.copy()
        s.dvv = self.dvv.copy()
        s.ps0 = self.ps0
        s.hyai = self.hyai.copy()
        return s

    def ptr_table(self):
        tab = (C.POINTER(C.c_double) * 16)()
        for i, n in enumerate(FIELD_NAMES):
            a = self.arrays[n]
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
            tab[i] = a.ctypes.data_as(C.POINTER(C.c_double))
        return tab


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def build_port(force=False):
    so = os.path.join(HERE, "libcaar_oracle.so")
    src = os.path.join(HERE, "caar_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return so


def build_ref(force=False):
    """Builds oracle/_ref from /root/reference when that exists (this container); on the GPU box the
    prebuilt files shipped in oracle/_ref are used as they are."""
    if os.path.isdir("/root/reference/compute_and_apply_rhs_test"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    return REF_DIR


def ref_available(nlev=72):
    return os.path.exists(os.path.join(REF_DIR, f"libcaar_ref_L{nlev}.so"))


class PortOracle:
    kind = "port"

    def __init__(self):
        self.lib = C.CDLL(build_port())
        L = self.lib
        L.caar_oracle_run.restype = C.c_double
        L.caar_oracle_run_eulerian.restype = C.c_double
        L.caar_oracle_saxpby.restype = C.c_double
        L.caar_oracle_field_count.restype = C.c_size_t

    def init(self, nelem, nlev=72, qsize_d=1, ntl=3) -> State:
        s = State.empty(nelem, nlev, qsize_d, ntl)
        dt2, ps0 = C.c_double(), C.c_double()
        self.lib.caar_oracle_init(nelem, nlev, qsize_d, ntl, s.ptr_table(), _ip(s.ctl), C.byref(dt2),
                                  _dp(s.consts), _dp(s.dvv), C.byref(ps0), _dp(s.hyai))
        s.dt2, s.ps0 = dt2.value, ps0.value
        return s

    def run(self, s: State, ncalls=1, nthreads=1) -> float:
        return self.lib.caar_oracle_run(s.nlev, s.qsize_d, s.ntl, s.ptr_table(), _ip(s.ctl),
                                        C.c_double(s.dt2), _dp(s.consts), _dp(s.dvv), C.c_double(s.ps0),
                                        _dp(s.hyai), ncalls, nthreads)

    def run_eulerian(self, s: State, hybi, ncalls=1, nthreads=1) -> float:
        """rsplit == 0 branch (fortran/routine_extracted.F90:227-262); PARITY UNPINNED restatement."""
        hybi = np.ascontiguousarray(hybi, dtype=np.float64)
        assert hybi.size == s.nlev + 1
        return self.lib.caar_oracle_run_eulerian(s.nlev, s.qsize_d, s.ntl, s.ptr_table(), _ip(s.ctl),
                                                 C.c_double(s.dt2), _dp(s.consts), _dp(s.dvv), C.c_double(s.ps0),
                                                 _dp(s.hyai), _dp(hybi), ncalls, nthreads)

    def norms(self, s: State, tl=None):
        out = np.zeros(3)
        tl = int(s.ctl[3]) if tl is None else tl
        self.lib.caar_oracle_norms(s.nlev, s.ntl, s.ptr_table(), int(s.ctl[0]), int(s.ctl[1]), tl, _dp(out))
        return out

    def divergence_sphere(self, v, s: State, ie):
        """One 4x4 level of PO divergence_sphere (sphere_operators.cpp:50-89): v (4,4,2) -> (4,4)."""
        A = s.arrays
        out = np.zeros((4, 4))
        v = np.ascontiguousarray(v, dtype=np.float64)
        self.lib.caar_oracle_divergence_sphere(_dp(v), _dp(s.dvv), _dp(A["elem_Dinv"][ie]), _dp(A["elem_metdet"][ie]),
                                               _dp(A["elem_rmetdet"][ie]), C.c_double(s.consts[0]), _dp(out))
        return out

    def euler_step(self, s: State, vstar, qtens, qn0, qsize, dt):
        """qtens = Qdp(qn0) - dt*div(vstar*Qdp(qn0)) (level_vectorized_ppscan/EulerStepFunctor.hpp:33-66)."""
        assert vstar.shape == (s.nelem, s.nlev, 4, 4, 2) and qtens.shape == (s.nelem, s.qsize_d, s.nlev, 4, 4)
        self.lib.caar_oracle_euler_step(s.nlev, s.qsize_d, s.ptr_table(), _dp(vstar), _dp(qtens), int(s.ctl[0]),
                                        int(s.ctl[1]), int(qn0), int(qsize), C.c_double(dt), _dp(s.dvv),
                                        C.c_double(s.consts[0]))

    def preq_vertadv(self, T, v, eta_dp_deta, rpdel):
        """T, rpdel (L,4,4); v (L,4,4,2); eta_dp_deta (L+1,4,4) -> (T_vadv, v_vadv)
        (level_vectorized_ppscan/CaarFunctor.hpp:504-547)."""
        L = T.shape[0]
        Tv, vv = np.zeros((L, 4, 4)), np.zeros((L, 4, 4, 2))
        c = np.ascontiguousarray
        self.lib.caar_oracle_preq_vertadv(L, _dp(c(T)), _dp(c(v)), _dp(c(eta_dp_deta)), _dp(c(rpdel)), _dp(Tv), _dp(vv))
        return Tv, vv

    def sphere_wk(self, name, s: State, field, tensorvisc=None, nets=None, nete=None):
        """divergence_sphere_wk (field (E,L,4,4,2)) / laplace_simple / laplace_tensor (field (E,L,4,4), tensorvisc
        (E,4,4,2,2)) -> (E,L,4,4); level_vectorized_ppscan/SphereOperators.hpp:493-636."""
        op = {"divergence_sphere_wk": 0, "laplace_simple": 1, "laplace_tensor": 2}[name]
        out = np.zeros((s.nelem, s.nlev, 4, 4))
        f = np.ascontiguousarray(field)
        tv = np.ascontiguousarray(tensorvisc) if tensorvisc is not None else None
        self.lib.caar_oracle_sphere_wk(op, s.nlev, s.ptr_table(), _dp(f) if op == 0 else None,
                                       _dp(f) if op != 0 else None, _dp(tv) if tv is not None else None, _dp(out),
                                       int(s.ctl[0]) if nets is None else nets, int(s.ctl[1]) if nete is None else nete,
                                       _dp(s.dvv), C.c_double(s.consts[0]))
        return out

    def gradient_sphere(self, sc, s: State, ie):
        out = np.zeros((4, 4, 2))
        self.lib.caar_oracle_gradient_sphere(_dp(np.ascontiguousarray(sc)), _dp(s.dvv), _dp(s.arrays["elem_Dinv"][ie]),
                                             C.c_double(s.consts[0]), _dp(out))
        return out

    def vorticity_sphere(self, v, s: State, ie):
        out = np.zeros((4, 4))
        self.lib.caar_oracle_vorticity_sphere(_dp(np.ascontiguousarray(v)), _dp(s.dvv), _dp(s.arrays["elem_D"][ie]),
                                              _dp(s.arrays["elem_rmetdet"][ie]), C.c_double(s.consts[0]), _dp(out))
        return out

    def saxpby(self, a, b, x, y, sweeps=1, nthreads=1) -> float:
        return self.lib.caar_oracle_saxpby(C.c_double(a), C.c_double(b), _dp(x), _dp(y), C.c_size_t(x.size),
                                           sweeps, nthreads)


class RefOracle:
    kind = "reference"

    def __init__(self, nlev=72):
        path = os.path.join(REF_DIR, f"libcaar_ref_L{nlev}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.caar_ref_run.restype = C.c_double
        self.nlev = self.lib.caar_ref_nlev()
        assert self.nlev == nlev

    def init(self, nelem, nlev=None, qsize_d=1, ntl=3) -> State:
        assert nlev in (None, self.nlev) and qsize_d == 1 and ntl == 3
        s = State.empty(nelem, self.nlev)
        dt2, ps0 = C.c_double(), C.c_double()
        self.lib.caar_ref_init(nelem, s.ptr_table(), _ip(s.ctl), C.byref(dt2), _dp(s.consts), _dp(s.dvv),
                               C.byref(ps0), _dp(s.hyai))
        s.dt2, s.ps0 = dt2.value, ps0.value
        return s

    def run(self, s: State, ncalls=1, nthreads=1) -> float:
        assert s.nlev == self.nlev and s.qsize_d == 1 and s.ntl == 3
        return self.lib.caar_ref_run(s.ptr_table(), _ip(s.ctl), C.c_double(s.dt2), _dp(s.consts), _dp(s.dvv),
                                     C.c_double(s.ps0), _dp(s.hyai), ncalls, nthreads)

    def divergence_sphere(self, v, s: State, ie):
        """The reference's own divergence_sphere on one level of element ie."""
        out = np.zeros((4, 4))
        v = np.ascontiguousarray(v, dtype=np.float64)
        self.lib.caar_ref_divergence_sphere(_dp(v), s.ptr_table(), int(ie), _dp(s.dvv), C.c_double(s.consts[0]),
                                            _dp(out))
        return out

    def norms(self, s: State, tl=None):
        out = np.zeros(3)
        tl = int(s.ctl[3]) if tl is None else tl
        self.lib.caar_ref_norms(s.ptr_table(), int(s.ctl[0]), int(s.ctl[1]), tl, _dp(out))
        return out


class RefSaxpby:
    def __init__(self):
        path = os.path.join(REF_DIR, "libsaxpby_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.saxpby_ref_run.restype = C.c_double

    def run(self, a, b, x, y, sweeps=1) -> float:
        i1 = x.size // (128 * 256)
        assert i1 * 128 * 256 == x.size
        return self.lib.saxpby_ref_run(C.c_double(a), C.c_double(b), _dp(x), _dp(y), i1, sweeps)


class HommexxOracle:
    """The reference's HOMMEXX prototypes — level_vectorized_ppscan ("lv") / tiled_vectorized_ppscan ("tv") —
    compiled unmodified against oracle/kokkos_stub (oracle/Makefile) into oracle/_ref/libhommexx_<variant>_L<nlev>.so.
    Every method takes and returns arrays in the pointers_only conventions ([.][igp][jgp]([c]) per level) and
    converts to the order the reference's own code reads: Fortran memory order, which is also the index order of the
    HOMMEXX views (level_vectorized_ppscan/Elements.cpp:48-99,154-292)."""
    kind = "reference (HOMMEXX prototype under a serial Kokkos stand-in)"
    OPS = {"gradient_sphere": 0, "divergence_sphere": 1, "vorticity_sphere": 2, "divergence_sphere_wk": 3,
           "laplace_simple": 4, "laplace_tensor": 5, "laplace_tensor_replace": 6, "divergence_sphere_update": 7}

    def __init__(self, variant="lv", nlev=72):
        path = os.path.join(REF_DIR, f"libhommexx_{variant}_L{nlev}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.variant, self.nlev = variant, self.lib.hx_nlev()
        self.qsize_d = self.lib.hx_qsize_d()
        assert self.nlev == nlev and self.lib.hx_variant() == (1 if variant == "tv" else 0)

    @staticmethod
    def available(variant="lv", nlev=72):
        return os.path.exists(os.path.join(REF_DIR, f"libhommexx_{variant}_L{nlev}.so"))

    def run(self, s: State, ncalls=1):
        """compute_and_apply_rhs through Elements::pull_from_f90_pointers -> CaarFunctor -> push_to_f90_pointers.
        Needs s.qsize_d == the library's QSIZE_D and the reference's physical constants (they are compile-time in
        HOMMEXX: level_vectorized_ppscan/PhysicalConstants.hpp:11-18). ps0 is an `int` member there (Control.hpp:50)."""
        assert s.nlev == self.nlev and s.qsize_d == self.qsize_d and s.ntl == 3
        f = to_f90(s.arrays)
        A = lambda n: _dp(f[n])
        dvv_f90 = np.ascontiguousarray(s.dvv.T)
        self.lib.hx_caar_f90(s.nelem, A("elem_D"), A("elem_Dinv"), A("elem_fcor"), A("elem_spheremp"), A("elem_metdet"),
                             A("elem_state_phis"), A("elem_state_v"), A("elem_state_T"), A("elem_state_dp3d"),
                             A("elem_derived_phi"), A("elem_derived_pecnd"), A("elem_derived_omega_p"),
                             A("elem_derived_vn0"), A("elem_derived_eta_dot_dpdn"), A("elem_state_Qdp"),
                             int(s.ctl[0]), int(s.ctl[1]), int(s.ctl[4]), int(s.ctl[2]), int(s.ctl[3]), int(s.ctl[5]),
                             C.c_double(s.dt2), C.c_double(s.ps0), C.c_double(s.consts[1]), _dp(s.hyai), _dp(dvv_f90),
                             int(ncalls))
        back = from_f90(f)
        for n in MUTATED:
            s.arrays[n][...] = back[n]

    def sphere_op(self, name, field, s: State, ie, tensorvisc=None, alpha=1.0, beta=0.0, out=None):
        """One of OPS on every level of element ie. Scalars are (L,4,4), vectors (L,4,4,2) in pointers_only order;
        tensorvisc (4,4,2,2) like elem_D. divergence_sphere_update: out = beta*out + alpha*div(field)."""
        op = self.OPS[name]
        L = self.nlev
        A = s.arrays
        t4 = lambda a: np.ascontiguousarray(a.transpose(3, 2, 1, 0))            # [i][j][a][b] -> [b][a][j][i]
        t2 = lambda a: np.ascontiguousarray(a.T)
        vec_in = op in (1, 2, 3, 7)
        fin = np.ascontiguousarray(field.transpose(0, 3, 2, 1) if vec_in else field.transpose(0, 2, 1))
        vec_out = op == 0
        fout = np.zeros((L, 2, 4, 4) if vec_out else (L, 4, 4))
        if op == 7:
            fout[...] = out.transpose(0, 2, 1)
        tv = t4(tensorvisc) if tensorvisc is not None else None
        rc = self.lib.hx_sphere_op(op, _dp(t4(A["elem_D"][ie])), _dp(t4(A["elem_Dinv"][ie])), _dp(t2(A["elem_metdet"][ie])),
                                   _dp(t2(A["elem_spheremp"][ie])), _dp(tv) if tv is not None else None,
                                   _dp(t2(s.dvv)), _dp(fin), _dp(fout), C.c_double(alpha), C.c_double(beta))
        assert rc == 0
        return np.ascontiguousarray(fout.transpose(0, 3, 2, 1) if vec_out else fout.transpose(0, 2, 1))

    def preq_vertadv(self, T, v, eta_dp_deta, rpdel):
        """CaarFunctor::preq_vertadv: T, rpdel (L,4,4), v (L,4,4,2), eta_dp_deta (L+1,4,4) -> (T_vadv, v_vadv)."""
        L = self.nlev
        sw = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1))
        Tv, vv = np.zeros((L, 4, 4)), np.zeros((L, 2, 4, 4))
        self.lib.hx_preq_vertadv(_dp(sw(T)), _dp(np.ascontiguousarray(v.transpose(0, 3, 2, 1))), _dp(sw(eta_dp_deta)),
                                 _dp(sw(rpdel)), _dp(Tv), _dp(vv))
        return sw(Tv), np.ascontiguousarray(vv.transpose(0, 3, 2, 1))

    def euler_step(self, s: State, vstar, qn0, qsize, dt):
        """EulerStepFunctor::operator() (tiled_vectorized_ppscan/EulerStepFunctor.hpp:33-70) on every element ->
        qtens (E, qsize_d, L, 4, 4). TV only (the LV copy of the functor does not compile, see ref_hommexx_capi.cpp)."""
        assert self.variant == "tv" and s.qsize_d == self.qsize_d and s.nlev == self.nlev
        E, L, Q = s.nelem, s.nlev, s.qsize_d
        A = s.arrays
        dinv = np.ascontiguousarray(A["elem_Dinv"].transpose(0, 4, 3, 2, 1))
        met = np.ascontiguousarray(A["elem_metdet"].transpose(0, 2, 1))
        vs = np.ascontiguousarray(vstar.transpose(0, 1, 4, 3, 2))
        qdp = np.ascontiguousarray(A["elem_state_Qdp"].transpose(0, 2, 1, 3, 5, 4))
        qt = np.zeros((E, Q, L, 4, 4))
        self.lib.hx_euler_step(E, int(qsize), int(qn0), C.c_double(dt), _dp(dinv), _dp(met),
                               _dp(np.ascontiguousarray(s.dvv.T)), _dp(vs), _dp(qdp), _dp(qt))
        return np.ascontiguousarray(qt.transpose(0, 1, 2, 4, 3))


def best_oracle(nlev=72):
    """The real reference when its prebuilt library is present, else the restatement."""
    if ref_available(nlev):
        return RefOracle(nlev)
    return PortOracle()


def randomize(s: State, seed=20261018):
    """Non-trivial geometry and fields for parity tests, in the spirit of the reference's own random
    bench init (level_vectorized_ppscan/Elements.cpp:101-152: U(1/64,1) fields, |det D| >= 1/64,
    Dinv = D^-1) but with a fixed seed and physically-scaled magnitudes."""
    rng = np.random.default_rng(seed)
    E, L = s.nelem, s.nlev
    A = s.arrays
    D = rng.uniform(-1.0, 1.0, size=(E, 4, 4, 2, 2))
    det = D[..., 0, 0] * D[..., 1, 1] - D[..., 0, 1] * D[..., 1, 0]
    bad = np.abs(det) < 1.0 / 64
    D[bad] = np.array([[1.0, 0.25], [-0.25, 0.75]])
    det = D[..., 0, 0] * D[..., 1, 1] - D[..., 0, 1] * D[..., 1, 0]
    Dinv = np.empty_like(D)
    Dinv[..., 0, 0] = D[..., 1, 1] / det
    Dinv[..., 0, 1] = -D[..., 0, 1] / det
    Dinv[..., 1, 0] = -D[..., 1, 0] / det
    Dinv[..., 1, 1] = D[..., 0, 0] / det
    A["elem_D"][...] = D
    A["elem_Dinv"][...] = Dinv
    A["elem_metdet"][...] = np.abs(det)
    A["elem_rmetdet"][...] = 1.0 / np.abs(det)
    A["elem_fcor"][...] = rng.uniform(-1.4e-4, 1.4e-4, size=(E, 4, 4))
    A["elem_spheremp"][...] = rng.uniform(1.0 / 64, 1.0, size=(E, 4, 4))
    A["elem_state_phis"][...] = rng.uniform(0.0, 3.0e3, size=(E, 4, 4))
    A["elem_state_dp3d"][...] = rng.uniform(5.0, 15.0, size=A["elem_state_dp3d"].shape)
    A["elem_state_v"][...] = rng.uniform(-40.0, 40.0, size=A["elem_state_v"].shape)
    A["elem_state_T"][...] = rng.uniform(200.0, 300.0, size=A["elem_state_T"].shape)
    A["elem_state_Qdp"][...] = rng.uniform(0.0, 0.02, size=A["elem_state_Qdp"].shape) * 10.0
    A["elem_derived_eta_dot_dpdn"][...] = rng.uniform(-1.0, 1.0, size=A["elem_derived_eta_dot_dpdn"].shape)
    A["elem_derived_omega_p"][...] = rng.uniform(-1.0, 1.0, size=A["elem_derived_omega_p"].shape)
    A["elem_derived_phi"][...] = rng.uniform(0.0, 1.0, size=A["elem_derived_phi"].shape)
    A["elem_derived_pecnd"][...] = rng.uniform(-10.0, 10.0, size=A["elem_derived_pecnd"].shape)
    A["elem_derived_vn0"][...] = rng.uniform(-1.0, 1.0, size=A["elem_derived_vn0"].shape)
    s.consts[1] = 0.25  # eta_ave_w != 1 exercises the accumulators
    s.dt2 = 37.5
    s.ps0 = 1.0e3
    s.hyai[...] = np.linspace(0.002, 0.0, L + 1)
    return s


# ---- Fortran (F90 flat pointer) memory order of every array — test-side restatement of the convention HOMMEXX
# reads its F90 pointers in (level_vectorized_ppscan/Elements.cpp:48-99,164-292; the Fortran declarations are
# fortran/element_state_mod.F90:17-23 and fortran/element_mod.F90:69-121). C++ [igp][jgp] == Fortran (i,j).
_F90_AXES = {
    "elem_D": (0, 4, 3, 2, 1), "elem_Dinv": (0, 4, 3, 2, 1),                       # [e][b][a][j][i]
    "elem_fcor": (0, 2, 1), "elem_spheremp": (0, 2, 1), "elem_metdet": (0, 2, 1),  # [e][j][i]
    "elem_rmetdet": (0, 2, 1), "elem_state_phis": (0, 2, 1),
    "elem_state_dp3d": (0, 1, 2, 4, 3), "elem_state_T": (0, 1, 2, 4, 3),           # [e][tl][lev][j][i]
    "elem_state_v": (0, 1, 2, 5, 4, 3),                                            # [e][tl][lev][c][j][i]
    "elem_state_Qdp": (0, 2, 1, 3, 5, 4),                                          # [e][qni][iq][lev][j][i]
    "elem_derived_eta_dot_dpdn": (0, 1, 3, 2), "elem_derived_omega_p": (0, 1, 3, 2),
    "elem_derived_phi": (0, 1, 3, 2), "elem_derived_pecnd": (0, 1, 3, 2),          # [e][lev][j][i]
    "elem_derived_vn0": (0, 1, 4, 3, 2),                                           # [e][lev][c][j][i]
}


def to_f90(arrays: dict) -> dict:
    """C++ (pointers_only) layout -> Fortran memory order, as new contiguous arrays."""
    return {n: np.ascontiguousarray(a.transpose(_F90_AXES[n])) for n, a in arrays.items()}


def from_f90(f90: dict) -> dict:
    """Fortran memory order -> C++ (pointers_only) layout."""
    out = {}
    for n, a in f90.items():
        ax = _F90_AXES[n]
        inv = np.argsort(ax)
        out[n] = np.ascontiguousarray(a.transpose(inv))
    return out
