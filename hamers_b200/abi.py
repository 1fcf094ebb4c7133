"""ctypes binding of include/hamers_b200.h.

PyTorch appears here only as plumbing: device memory (tensors own the HBM-resident patch data)
and the CUDA stream the plan launches on.  All arithmetic happens in libhamers_b200.so.  There
is no CPU fallback: if the library is missing or no CUDA device exists, plan creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhamers_b200.so")

MAX_SPECIES = 4
GHOSTS = 4
SINGLE_SPECIES = 0
FIVE_EQN_ALLAIRE = 1
FOUR_EQN_CONSERVATIVE = 2      # SURVEY row f3 (needs species_R; reference-order kernels only)
MATH_EXACT = 0
MATH_FAST = 1
DIFF_NODE_SIXTH_ORDER, DIFF_MIDPOINT_SIXTH_ORDER = 0, 1

# every symbol include/hamers_b200.h declares (tests check the library exports them all)
SYMBOLS = [
    "hb2_last_error", "hb2_version", "hb2_constants", "hb2_device_count", "hb2_num_eqn", "hb2_num_comp", "hb2_num_ghosts",
    "hb2_cell_ghost_size", "hb2_cell_size", "hb2_side_size", "hb2_plan_create", "hb2_plan_destroy",
    "hb2_plan_set_stream", "hb2_plan_use_own_stream", "hb2_plan_synchronize", "hb2_plan_launch_count", "hb2_plan_workspace_bytes",
    "hb2_compute_flux_and_source_dev", "hb2_advance_stage_dev", "hb2_fused_stage_dev",
    "hb2_fill_ghosts_periodic_dev", "hb2_pack_box_dev", "hb2_unpack_box_dev", "hb2_pack_boxes_dev", "hb2_unpack_boxes_dev",
    "hb2_push_boxes_dev",
    "hb2_max_wave_speed_dev", "hb2_fused_stage_push_dev", "hb2_device_malloc", "hb2_device_free", "hb2_ipc_export",
    "hb2_ipc_open", "hb2_ipc_close",
    "hb2_compute_flux_and_source_host", "hb2_fused_stage_host", "hb2_probe_fp64_peak", "hb2_probe_hbm_bandwidth",
    "hb2_plan_set_profiling", "hb2_plan_get_profile", "hb2_advance_level_dev", "hb2_advance_level_host",
    # diffusive flux of the single-species Navier-Stokes application
    "hb2_diffusive_plan_create", "hb2_diffusive_plan_destroy", "hb2_diffusive_plan_set_stream", "hb2_diffusive_plan_launches", "hb2_diffusive_plan_set_math", "hb2_diffusive_plan_set_reconstructor",
    "hb2_compute_diffusive_flux_dev", "hb2_compute_diffusive_flux_host", "hb2_advance_stage_ns_dev",
    "hb2_diffusive_fill_ghosts_periodic_dev", "hb2_diffusive_extract_view_dev", "hb2_diffusive_accumulate_dev",
    "hb2_diffusive_divergence_accumulate_dev", "hb2_diffusive_max_spectral_radius_dev",
    # two-level AMR operators
    "hb2_amr_refine_dev", "hb2_amr_coarsen_dev", "hb2_amr_fluxsum_update_dev", "hb2_amr_coarsen_fluxsum_dev",
    "hb2_fill_ghosts_extrapolate_dev",
    # device-resident patch level (seam 2)
    "hb2_level_create", "hb2_level_destroy", "hb2_level_num_patches", "hb2_level_launch_count", "hb2_level_synchronize",
    "hb2_level_upload_patch", "hb2_level_download_patch", "hb2_level_patch_state_dev", "hb2_level_fill_ghosts",
    "hb2_level_advance_stage", "hb2_level_advance", "hb2_level_max_wave_speed", "hb2_level_advance_stage_patch",
    "hb2_level_end_stage", "hb2_level_advance_host",
]

WCNS5_JS, WCNS5_Z, WCNS6_LD = 0, 1, 2

KERNEL_KINDS = ("sensor", "xsweep", "ysweep", "zsweep", "advance", "fill_periodic", "pack", "unpack")

# SSP-RK3(3,3), the reference's default table (RungeKuttaLevelIntegrator.cpp:3894-3929), row-major [stage][m]
SSPRK3_ALPHA = np.array([[1.0, 0.0, 0.0], [3.0 / 4.0, 1.0 / 4.0, 0.0], [1.0 / 3.0, 0.0, 2.0 / 3.0]])
SSPRK3_BETA = np.array([[1.0, 0.0, 0.0], [0.0, 1.0 / 4.0, 0.0], [0.0, 0.0, 2.0 / 3.0]])
# weights of the flux / source sums kept for the AMR flux synchronisation (hb2_advance_stage_dev's gamma)
SSPRK3_GAMMA = np.array([[1.0 / 6.0, 0.0, 0.0], [0.0, 1.0 / 6.0, 0.0], [0.0, 0.0, 2.0 / 3.0]])


class PatchDescC(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("n", C.c_int32 * 3),
        ("flow_model", C.c_int32),
        ("num_species", C.c_int32),
        ("species_gamma", C.c_double * MAX_SPECIES),
        ("dx", C.c_double * 3),
        ("weno_p", C.c_int32),
        ("math", C.c_int32),
        ("device", C.c_int32),
        ("scheme", C.c_int32),
        ("weno_q", C.c_int32),
        ("weno_C", C.c_double),
        ("weno_alpha_tau", C.c_double),
        ("num_ghosts", C.c_int32),
        ("species_R", C.c_double * MAX_SPECIES),
    ]


class HamersB200Error(RuntimeError):
    pass


_LIB = None


def load_library():
    """Load libhamers_b200.so; raises (never falls back) when it is missing."""
    global _LIB, SO
    if _LIB is None:
        SO = os.environ.get("HAMERS_B200_LIB", SO)   # tuning variants (tools/build_variant.py)
        if not os.path.exists(SO):
            raise HamersB200Error(
                f"{SO} not found: build it with `python -m hamers_b200.build` (there is no CPU fallback)")
        lib = C.CDLL(SO)
        lib.hb2_last_error.restype = C.c_char_p
        lib.hb2_version.restype = C.c_char_p
        lib.hb2_level_launch_count.restype = C.c_int64
        lib.hb2_level_launch_count.argtypes = [C.c_void_p]
        for f in ("hb2_cell_ghost_size", "hb2_cell_size", "hb2_side_size", "hb2_plan_launch_count",
                  "hb2_plan_workspace_bytes"):
            getattr(lib, f).restype = C.c_int64
        lib.hb2_plan_create.argtypes = [C.POINTER(PatchDescC), C.POINTER(C.c_void_p)]
        lib.hb2_plan_destroy.argtypes = [C.c_void_p]
        lib.hb2_plan_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.hb2_plan_synchronize.argtypes = [C.c_void_p]
        lib.hb2_plan_launch_count.argtypes = [C.c_void_p]
        lib.hb2_plan_workspace_bytes.argtypes = [C.c_void_p]
        lib.hb2_plan_set_profiling.argtypes = [C.c_void_p, C.c_int32]
        lib.hb2_plan_get_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]
        _LIB = lib
    return _LIB


def _check(rc: int, what: str):
    if rc != 0:
        raise HamersB200Error(f"{what} failed ({rc}): {load_library().hb2_last_error().decode()}")


def _ptr_table(ptrs: Sequence[Optional[int]]):
    T = (C.c_void_p * len(ptrs))()
    for i, p in enumerate(ptrs):
        T[i] = p if p else None
    return T


def _dev_ptrs(t, n):
    """device pointers of the n leading-axis slices of a contiguous float64 CUDA tensor (or list of tensors)."""
    import torch

    if isinstance(t, (list, tuple)):
        out = []
        for x in t:
            out += _dev_ptrs(x, x.shape[0]) if x is not None else []
        return out
    assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous(), "need a contiguous float64 CUDA tensor"
    assert t.shape[0] == n, (t.shape, n)
    stride = t[0].numel() * 8
    base = t.data_ptr()
    return [base + i * stride for i in range(n)]


def _host_ptrs(a: np.ndarray, n):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.shape[0] == n
    stride = a[0].size * 8
    return [a.ctypes.data + i * stride for i in range(n)]


class Plan:
    """One patch shape of the hot path (mirrors hb2_plan_t).

    Array conventions (numpy / torch views of the SAMRAI layouts, x fastest):
      state  : (num_comp, [nz+8,] ny+8, nx+8)
      flux d : (num_eqn, ...) with extent +1 along direction d, ghost 0
      source : (num_eqn, [nz,] ny, nx)
    """

    def __init__(self, dim: int, n: Sequence[int], flow_model: int = SINGLE_SPECIES,
                 species_gamma: Sequence[float] = (1.4,), dx: Sequence[float] = (1.0, 1.0, 1.0),
                 weno_p: int = 2, math: int = MATH_EXACT, device: int = -1, scheme: int = 0, weno_q: int = 4,
                 weno_C: float = 1.0e9, weno_alpha_tau: float = 35.0, num_ghosts: int = GHOSTS,
                 species_R: Sequence[float] = ()):
        self.lib = load_library()
        d = PatchDescC()
        d.num_ghosts = int(num_ghosts)
        self.num_ghosts = int(num_ghosts)
        d.dim = dim
        for a in range(3):
            d.n[a] = int(n[a]) if a < dim else 1
            d.dx[a] = float(dx[a]) if a < dim else 1.0
        d.flow_model = flow_model
        d.num_species = len(species_gamma) if flow_model != SINGLE_SPECIES else 1
        for i, g in enumerate(species_gamma):
            d.species_gamma[i] = float(g)
        for i, r in enumerate(species_R):
            d.species_R[i] = float(r)
        d.weno_p = weno_p
        d.math = math
        d.device = device
        d.scheme, d.weno_q, d.weno_C, d.weno_alpha_tau = int(scheme), int(weno_q), float(weno_C), float(weno_alpha_tau)
        self.desc = d
        self.dim = dim
        self.n = tuple(int(n[a]) for a in range(dim))
        self.flow_model = flow_model
        self.num_species = d.num_species
        if flow_model == FOUR_EQN_CONSERVATIVE:
            self.neq = dim + 1 + d.num_species
        else:
            self.neq = dim + 2 if flow_model == SINGLE_SPECIES else dim + 2 * d.num_species
        self.ncomp = self.neq + 1 if flow_model == FIVE_EQN_ALLAIRE else self.neq
        self._h = C.c_void_p()
        _check(self.lib.hb2_plan_create(C.byref(d), C.byref(self._h)), "hb2_plan_create")

    # -- shapes ---------------------------------------------------------------------------
    @property
    def ghost_shape(self):
        return tuple(self.n[a] + 2 * self.num_ghosts for a in reversed(range(self.dim)))

    @property
    def cell_shape(self):
        return tuple(self.n[a] for a in reversed(range(self.dim)))

    def side_shape(self, direction: int):
        e = list(self.n)
        e[direction] += 1
        return tuple(reversed(e))

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if self._h:
            self.lib.hb2_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def use_torch_stream(self):
        """Launch on torch's current stream so that tensor ops and plan calls are ordered."""
        import torch

        _check(self.lib.hb2_plan_set_stream(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "hb2_plan_set_stream")
        return self

    def synchronize(self):
        _check(self.lib.hb2_plan_synchronize(self._h), "hb2_plan_synchronize")

    @property
    def launch_count(self) -> int:
        return int(self.lib.hb2_plan_launch_count(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.hb2_plan_workspace_bytes(self._h))

    def set_profiling(self, on: bool = True):
        _check(self.lib.hb2_plan_set_profiling(self._h, 1 if on else 0), "hb2_plan_set_profiling")

    def get_profile(self, reset: bool = True):
        """{kind: (total ms, launches)} measured with CUDA events on the launching stream."""
        ms = (C.c_double * len(KERNEL_KINDS))()
        n = (C.c_int64 * len(KERNEL_KINDS))()
        _check(self.lib.hb2_plan_get_profile(self._h, ms, n, 1 if reset else 0), "hb2_plan_get_profile")
        return {k: (ms[i], int(n[i])) for i, k in enumerate(KERNEL_KINDS)}

    def advance_level(self, U, dt: float, alpha=None, beta=None, periodic_mask: int = 7):
        """RungeKuttaLevelIntegrator::advanceLevel stage loop for one patch covering a periodic level."""
        alpha = SSPRK3_ALPHA if alpha is None else np.ascontiguousarray(alpha, dtype=np.float64)
        beta = SSPRK3_BETA if beta is None else np.ascontiguousarray(beta, dtype=np.float64)
        nst = alpha.shape[0]
        _check(self.lib.hb2_advance_level_dev(self._h, nst, alpha.ctypes.data_as(C.POINTER(C.c_double)),
                                              beta.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(dt),
                                              int(periodic_mask), _ptr_table(_dev_ptrs(U, self.ncomp))),
               "hb2_advance_level_dev")

    def advance_level_host(self, U: np.ndarray, dt: float, alpha=None, beta=None, periodic_mask: int = 7):
        alpha = SSPRK3_ALPHA if alpha is None else np.ascontiguousarray(alpha, dtype=np.float64)
        beta = SSPRK3_BETA if beta is None else np.ascontiguousarray(beta, dtype=np.float64)
        nst = alpha.shape[0]
        _check(self.lib.hb2_advance_level_host(self._h, nst, alpha.ctypes.data_as(C.POINTER(C.c_double)),
                                               beta.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(dt),
                                               int(periodic_mask), _ptr_table(_host_ptrs(U, self.ncomp))),
               "hb2_advance_level_host")
        return U

    # -- device-resident hot path ----------------------------------------------------------
    def compute_flux_and_source(self, Q, dt: float, flux, source=None):
        """ConvectiveFluxReconstructor::computeConvectiveFluxAndSourceOnPatch on device tensors.
        flux: list (per direction) of (neq, *side_shape) tensors; source: (neq, *cell_shape) or None."""
        qp = _ptr_table(_dev_ptrs(Q, self.ncomp))
        fp = _ptr_table([p for d in range(self.dim) for p in _dev_ptrs(flux[d], self.neq)])
        sp = _ptr_table(_dev_ptrs(source, self.neq)) if source is not None else None
        _check(self.lib.hb2_compute_flux_and_source_dev(self._h, qp, C.c_double(dt), fp, sp),
               "hb2_compute_flux_and_source_dev")

    def fused_stage(self, alpha, beta, U_int, dt: float, U_out, push=None):
        """computeFluxesAndSourcesOnPatch + advanceSingleStepOnPatch without materialising fluxes.
        push: optional ctypes table from push_table() -- the ghost fill of the new state fused into the update."""
        ncoef = len(alpha)
        tab = _ptr_table([p for m in range(ncoef) for p in _dev_ptrs(U_int[m], self.ncomp)])
        a = (C.c_double * ncoef)(*[float(x) for x in alpha])
        b = (C.c_double * ncoef)(*[float(x) for x in beta])
        _check(self.lib.hb2_fused_stage_push_dev(self._h, ncoef, a, b, tab, C.c_double(dt),
                                                 _ptr_table(_dev_ptrs(U_out, self.ncomp)), push), "hb2_fused_stage_push_dev")

    def push_table(self, base_by_offset):
        """base_by_offset: {(ox, oy[, oz]): device address of component 0 of the neighbour's U_out} (components are
        one ghost box apart); missing offsets get NULL entries."""
        tab = (C.c_void_p * (27 * self.ncomp))()
        comp_bytes = 8 * int(np.prod(self.ghost_shape))
        for o, base in base_by_offset.items():
            o3 = tuple(o) + (0,) * (3 - len(o))
            code = (o3[0] + 1) + 3 * (o3[1] + 1) + 9 * (o3[2] + 1)
            for c in range(self.ncomp):
                tab[code * self.ncomp + c] = int(base) + c * comp_bytes
        return tab

    def advance_stage(self, alpha, beta, U_int, F_int, S_int, U_out, gamma=None, F_acc=None, S_acc=None):
        """Euler::advanceSingleStepOnPatch from materialised fluxes.  F_int[m]: list per direction or None."""
        ncoef = len(alpha)
        nf = self.dim * self.neq
        ut = _ptr_table([p for m in range(ncoef) for p in _dev_ptrs(U_int[m], self.ncomp)])
        fl, sl = [], []
        for m in range(ncoef):
            if F_int[m] is None:
                fl += [None] * nf
                sl += [None] * self.neq
            else:
                fl += [p for d in range(self.dim) for p in _dev_ptrs(F_int[m][d], self.neq)]
                sl += _dev_ptrs(S_int[m], self.neq) if S_int[m] is not None else [None] * self.neq
        a = (C.c_double * ncoef)(*[float(x) for x in alpha])
        b = (C.c_double * ncoef)(*[float(x) for x in beta])
        g = (C.c_double * ncoef)(*[float(x) for x in (gamma if gamma is not None else [0.0] * ncoef)])
        fa = _ptr_table([p for d in range(self.dim) for p in _dev_ptrs(F_acc[d], self.neq)]) if F_acc is not None else None
        sa = _ptr_table(_dev_ptrs(S_acc, self.neq)) if S_acc is not None else None
        _check(self.lib.hb2_advance_stage_dev(self._h, ncoef, a, b, g, ut, _ptr_table(fl), _ptr_table(sl),
                                              _ptr_table(_dev_ptrs(U_out, self.ncomp)), fa, sa),
               "hb2_advance_stage_dev")

    def fill_ghosts_periodic(self, U, mask: int = 7):
        _check(self.lib.hb2_fill_ghosts_periodic_dev(self._h, _ptr_table(_dev_ptrs(U, self.ncomp)), int(mask)),
               "hb2_fill_ghosts_periodic_dev")

    def fill_ghosts_extrapolate(self, U, direction: int, side: int):
        """BDRY_COND::BASIC::FLOW on one face of the patch: ghost cells copy the adjacent interior cell."""
        _check(self.lib.hb2_fill_ghosts_extrapolate_dev(self._h, _ptr_table(_dev_ptrs(U, self.ncomp)), int(direction), int(side)),
               "hb2_fill_ghosts_extrapolate_dev")

    def pack_box(self, U, lo, hi, buffer):
        l = (C.c_int32 * 3)(*[int(lo[a]) if a < self.dim else 0 for a in range(3)])
        h = (C.c_int32 * 3)(*[int(hi[a]) if a < self.dim else 1 for a in range(3)])
        _check(self.lib.hb2_pack_box_dev(self._h, _ptr_table(_dev_ptrs(U, self.ncomp)), l, h,
                                         C.c_void_p(buffer.data_ptr())), "hb2_pack_box_dev")

    def unpack_box(self, U, lo, hi, buffer):
        l = (C.c_int32 * 3)(*[int(lo[a]) if a < self.dim else 0 for a in range(3)])
        h = (C.c_int32 * 3)(*[int(hi[a]) if a < self.dim else 1 for a in range(3)])
        _check(self.lib.hb2_unpack_box_dev(self._h, _ptr_table(_dev_ptrs(U, self.ncomp)), l, h,
                                           C.c_void_p(buffer.data_ptr())), "hb2_unpack_box_dev")

    def box_table(self, boxes, offsets):
        """ctypes tables of a list of (lo, hi) boxes and their buffer offsets (in doubles) for pack_boxes / unpack_boxes."""
        n = len(boxes)
        lo = (C.c_int32 * (3 * n))()
        hi = (C.c_int32 * (3 * n))()
        for b, (l, h) in enumerate(boxes):
            for a in range(3):
                lo[3 * b + a] = int(l[a]) if a < self.dim else 0
                hi[3 * b + a] = int(h[a]) if a < self.dim else 1
        off = (C.c_int64 * n)(*[int(o) for o in offsets])
        return n, lo, hi, off

    def pack_boxes(self, U, table, buffer):
        n, lo, hi, off = table
        _check(self.lib.hb2_pack_boxes_dev(self._h, _ptr_table(_dev_ptrs(U, self.ncomp)), n, lo, hi, off,
                                           C.c_void_p(buffer.data_ptr())), "hb2_pack_boxes_dev")

    def unpack_boxes(self, U, table, buffer):
        n, lo, hi, off = table
        _check(self.lib.hb2_unpack_boxes_dev(self._h, _ptr_table(_dev_ptrs(U, self.ncomp)), n, lo, hi, off,
                                             C.c_void_p(buffer.data_ptr())), "hb2_unpack_boxes_dev")

    def peer_box_table(self, boxes, bases, shifts):
        """ctypes tables for push_boxes: boxes = [(lo, hi)], bases[b] = device address of component 0 of the destination
        state array (this GPU's or an IPC-opened peer's, same ghost-box geometry), shifts[b] = cell offset subtracted."""
        n = len(boxes)
        lo = (C.c_int32 * (3 * n))()
        hi = (C.c_int32 * (3 * n))()
        sh = (C.c_int32 * (3 * n))()
        for b, (l, h) in enumerate(boxes):
            for a in range(3):
                lo[3 * b + a] = int(l[a]) if a < self.dim else 0
                hi[3 * b + a] = int(h[a]) if a < self.dim else 1
                sh[3 * b + a] = int(shifts[b][a]) if a < self.dim else 0
        dst = (C.c_void_p * n)(*[int(x) for x in bases])
        return n, lo, hi, dst, sh

    def push_boxes(self, U, table):
        """Store the boxes of U straight into the neighbouring patches' state arrays (hb2_push_boxes_dev)."""
        n, lo, hi, dst, sh = table
        _check(self.lib.hb2_push_boxes_dev(self._h, _ptr_table(_dev_ptrs(U, self.ncomp)), n, lo, hi, dst, sh,
                                           C.c_int64(int(np.prod(self.ghost_shape)))), "hb2_push_boxes_dev")

    def max_wave_speed(self, Q, out):
        _check(self.lib.hb2_max_wave_speed_dev(self._h, _ptr_table(_dev_ptrs(Q, self.ncomp)),
                                               C.c_void_p(out.data_ptr())), "hb2_max_wave_speed_dev")

    # -- host-buffer hot path (numpy in, numpy out; H2D/D2H inside the call) -----------------
    def compute_flux_and_source_host(self, Q: np.ndarray, dt: float, flux=None, source=None):
        if flux is None:
            flux = [np.empty((self.neq,) + self.side_shape(d)) for d in range(self.dim)]
        if source is None:
            source = np.zeros((self.neq,) + self.cell_shape)
        qp = _ptr_table(_host_ptrs(Q, self.ncomp))
        fp = _ptr_table([p for d in range(self.dim) for p in _host_ptrs(flux[d], self.neq)])
        sp = _ptr_table(_host_ptrs(source, self.neq))
        _check(self.lib.hb2_compute_flux_and_source_host(self._h, qp, C.c_double(dt), fp, sp),
               "hb2_compute_flux_and_source_host")
        return flux, source

    def fused_stage_host(self, alpha, beta, U_int, dt: float, U_out: Optional[np.ndarray] = None):
        ncoef = len(alpha)
        if U_out is None:
            U_out = np.zeros((self.ncomp,) + self.ghost_shape)
        tab = _ptr_table([p for m in range(ncoef) for p in _host_ptrs(U_int[m], self.ncomp)])
        a = (C.c_double * ncoef)(*[float(x) for x in alpha])
        b = (C.c_double * ncoef)(*[float(x) for x in beta])
        _check(self.lib.hb2_fused_stage_host(self._h, ncoef, a, b, tab, C.c_double(dt),
                                             _ptr_table(_host_ptrs(U_out, self.ncomp))), "hb2_fused_stage_host")
        return U_out


DIFF_GHOSTS = 6   # DiffusiveFluxReconstructorNodeSixthOrder.cpp:24


class DiffusiveDescC(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n", C.c_int32 * 3), ("dx", C.c_double * 3), ("species_gamma", C.c_double),
                ("species_c_v", C.c_double), ("species_mu", C.c_double), ("species_mu_v", C.c_double),
                ("species_c_p", C.c_double), ("species_Pr", C.c_double), ("device", C.c_int32)]


class DiffusivePlan:
    """SURVEY row f4 (mirrors hb2_diff_plan_t): DiffusiveFluxReconstructorNodeSixthOrder ("SIXTH_ORDER") of the
    single-species Navier-Stokes application with CONSTANT viscosities and PRANDTL conductivity, and the stage update
    that consumes its flux.  State arrays here carry SIX ghost cells: (dim + 2, [nz+12,] ny+12, nx+12)."""

    def __init__(self, dim: int, n: Sequence[int], dx: Sequence[float], species_gamma: float, species_c_v: float,
                 species_mu: float, species_mu_v: float, species_c_p: float, species_Pr: float, device: int = -1):
        self.lib = load_library()
        d = DiffusiveDescC()
        d.dim = dim
        for a in range(3):
            d.n[a] = int(n[a]) if a < dim else 1
            d.dx[a] = float(dx[a]) if a < dim else 1.0
        d.species_gamma, d.species_c_v, d.species_mu, d.species_mu_v = species_gamma, species_c_v, species_mu, species_mu_v
        d.species_c_p, d.species_Pr, d.device = species_c_p, species_Pr, device
        self.dim, self.n, self.neq = dim, tuple(int(x) for x in n[:dim]), dim + 2
        self._h = C.c_void_p()
        _check(self.lib.hb2_diffusive_plan_create(C.byref(d), C.byref(self._h)), "hb2_diffusive_plan_create")

    def close(self):
        if self._h:
            self.lib.hb2_diffusive_plan_destroy(self._h)
            self._h = C.c_void_p()

    def use_torch_stream(self):
        import torch

        _check(self.lib.hb2_diffusive_plan_set_stream(self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "hb2_diffusive_plan_set_stream")
        return self

    def set_reconstructor(self, reconstructor: int):
        """DIFF_NODE_SIXTH_ORDER ("SIXTH_ORDER", default) or DIFF_MIDPOINT_SIXTH_ORDER ("MIDPOINT_SIXTH_ORDER")"""
        _check(self.lib.hb2_diffusive_plan_set_reconstructor(self._h, C.c_int32(int(reconstructor))),
               "hb2_diffusive_plan_set_reconstructor")
        return self

    def set_math(self, math: int):
        """arithmetic of divergence_accumulate: MATH_EXACT (default, bit-identical to the oracle) or MATH_FAST (re-associated,
        <= 1e-12 relative; 3-D)."""
        _check(self.lib.hb2_diffusive_plan_set_math(self._h, C.c_int32(int(math))), "hb2_diffusive_plan_set_math")
        return self

    @property
    def launch_count(self) -> int:
        v = C.c_int64()
        _check(self.lib.hb2_diffusive_plan_launches(self._h, C.byref(v)), "hb2_diffusive_plan_launches")
        return v.value

    @property
    def ghost_shape(self):
        return tuple(x + 2 * DIFF_GHOSTS for x in reversed(self.n))

    def side_shape(self, d: int):
        s = list(self.n)
        s[d] += 1
        return tuple(reversed(s))

    def compute_diffusive_flux(self, Q, dt: float, flux):
        """DiffusiveFluxReconstructor::computeDiffusiveFluxOnPatch on device tensors (Q: ghost 6, all ghosts filled)."""
        qp = _ptr_table(_dev_ptrs(Q, self.neq))
        fp = _ptr_table([p for d in range(self.dim) for p in _dev_ptrs(flux[d], self.neq)])
        _check(self.lib.hb2_compute_diffusive_flux_dev(self._h, qp, C.c_double(dt), fp), "hb2_compute_diffusive_flux_dev")

    def compute_diffusive_flux_host(self, Q: np.ndarray, dt: float, flux=None):
        if flux is None:
            flux = [np.empty((self.neq,) + self.side_shape(d)) for d in range(self.dim)]
        qp = _ptr_table(_host_ptrs(Q, self.neq))
        fp = _ptr_table([p for d in range(self.dim) for p in _host_ptrs(flux[d], self.neq)])
        _check(self.lib.hb2_compute_diffusive_flux_host(self._h, qp, C.c_double(dt), fp), "hb2_compute_diffusive_flux_host")
        return flux

    def fill_ghosts_periodic(self, U, mask: int = 7):
        """Periodic same-level ghost fill of a six-ghost state tensor (neq, *ghost_shape), in place."""
        _check(self.lib.hb2_diffusive_fill_ghosts_periodic_dev(self._h, _ptr_table(_dev_ptrs(U, self.neq)), int(mask)),
               "hb2_diffusive_fill_ghosts_periodic_dev")

    def extract_view(self, U, num_ghosts: int, U_view):
        """Copy the num_ghosts-wide ghost box of the six-ghost state U into U_view (neq, n + 2 num_ghosts ...)."""
        _check(self.lib.hb2_diffusive_extract_view_dev(self._h, _ptr_table(_dev_ptrs(U, self.neq)), int(num_ghosts),
                                                       _ptr_table(_dev_ptrs(U_view, self.neq))),
               "hb2_diffusive_extract_view_dev")

    def accumulate(self, num_ghosts: int, beta: float, Fd, U):
        """U += beta (-div F_d) on the interior of U (neq, ... with num_ghosts ghosts); Fd: list per direction."""
        fp = _ptr_table([p for d in range(self.dim) for p in _dev_ptrs(Fd[d], self.neq)])
        _check(self.lib.hb2_diffusive_accumulate_dev(self._h, int(num_ghosts), C.c_double(beta), fp,
                                                     _ptr_table(_dev_ptrs(U, self.neq))), "hb2_diffusive_accumulate_dev")

    def divergence_accumulate(self, Q, dt: float, num_ghosts: int, beta: float, U):
        """U += beta (-div F_d(Q)) without writing the diffusive side flux (bit-identical to compute_diffusive_flux +
        accumulate).  Q: six-ghost state; U: (neq, ... with num_ghosts ghosts), not Q."""
        _check(self.lib.hb2_diffusive_divergence_accumulate_dev(self._h, _ptr_table(_dev_ptrs(Q, self.neq)), C.c_double(dt),
                                                                int(num_ghosts), C.c_double(beta),
                                                                _ptr_table(_dev_ptrs(U, self.neq))),
               "hb2_diffusive_divergence_accumulate_dev")

    def max_spectral_radius(self, Q, species_c_p_eos: float, out):
        """out[0] (1-element float64 CUDA tensor) = max diffusive spectral radius of the six-ghost state Q."""
        _check(self.lib.hb2_diffusive_max_spectral_radius_dev(self._h, _ptr_table(_dev_ptrs(Q, self.neq)),
                                                              C.c_double(species_c_p_eos), C.c_void_p(out.data_ptr())),
               "hb2_diffusive_max_spectral_radius_dev")

    def advance_stage_ns(self, num_ghosts: int, alpha, beta, U_int, Fc_int, Fd_int, S_int, U_out):
        """NavierStokes::advanceSingleStepOnPatch (conservative diffusive flux) on device tensors; rows with a zero
        coefficient may be None."""
        ncoef, neq, dim = len(alpha), self.neq, self.dim

        def rows(src, per_row, width):
            out = []
            for m in range(ncoef):
                out += per_row(src[m]) if src[m] is not None else [None] * width
            return _ptr_table(out)
        a = (C.c_double * ncoef)(*[float(x) for x in alpha])
        b = (C.c_double * ncoef)(*[float(x) for x in beta])
        _check(self.lib.hb2_advance_stage_ns_dev(
            self._h, int(num_ghosts), ncoef, a, b, rows(U_int, lambda u: _dev_ptrs(u, neq), neq),
            rows(Fc_int, lambda f: [p for d in range(dim) for p in _dev_ptrs(f[d], neq)], dim * neq),
            rows(Fd_int, lambda f: [p for d in range(dim) for p in _dev_ptrs(f[d], neq)], dim * neq),
            rows(S_int, lambda s_: _dev_ptrs(s_, neq), neq), _ptr_table(_dev_ptrs(U_out, neq))), "hb2_advance_stage_ns_dev")


def device_malloc(nbytes: int) -> int:
    p = C.c_void_p()
    _check(load_library().hb2_device_malloc(C.c_int64(int(nbytes)), C.byref(p)), "hb2_device_malloc")
    return p.value


def device_free(ptr: int):
    _check(load_library().hb2_device_free(C.c_void_p(ptr)), "hb2_device_free")


def ipc_export(ptr: int) -> bytes:
    h = (C.c_uint8 * 64)()
    _check(load_library().hb2_ipc_export(C.c_void_p(ptr), h), "hb2_ipc_export")
    return bytes(h)


def ipc_open(handle: bytes) -> int:
    h = (C.c_uint8 * 64)(*handle)
    p = C.c_void_p()
    _check(load_library().hb2_ipc_open(h, C.byref(p)), "hb2_ipc_open")
    return p.value


def ipc_close(ptr: int):
    _check(load_library().hb2_ipc_close(C.c_void_p(ptr)), "hb2_ipc_close")


class AmrPairC(C.Structure):
    _fields_ = [("dim", C.c_int32), ("nc", C.c_int32 * 3), ("nf", C.c_int32 * 3), ("ratio", C.c_int32 * 3),
                ("origin", C.c_int32 * 3), ("ghosts_c", C.c_int32), ("ghosts_f", C.c_int32), ("ncomp", C.c_int32),
                ("neq", C.c_int32), ("dxc", C.c_double * 3), ("dxf", C.c_double * 3)]


class AmrPair:
    """One coarse patch and one fine patch of a two-level hierarchy (mirrors hb2_amr_pair)."""

    def __init__(self, dim, nc, nf, ratio, origin, dxc, dxf, ncomp, neq, ghosts_c=GHOSTS, ghosts_f=GHOSTS):
        self.lib = load_library()
        d = AmrPairC()
        d.dim = dim
        for a in range(3):
            on = a < dim
            d.nc[a] = int(nc[a]) if on else 1
            d.nf[a] = int(nf[a]) if on else 1
            d.ratio[a] = int(ratio[a]) if on else 1
            d.origin[a] = int(origin[a]) if on else 0
            d.dxc[a] = float(dxc[a]) if on else 1.0
            d.dxf[a] = float(dxf[a]) if on else 1.0
        d.ghosts_c, d.ghosts_f, d.ncomp, d.neq = int(ghosts_c), int(ghosts_f), int(ncomp), int(neq)
        self.desc, self.dim, self.ncomp, self.neq = d, dim, int(ncomp), int(neq)

    @staticmethod
    def _stream():
        import torch

        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _box(self, lo, hi):
        l = (C.c_int32 * 3)(*[int(lo[a]) if a < self.dim else 0 for a in range(3)])
        h = (C.c_int32 * 3)(*[int(hi[a]) if a < self.dim else 1 for a in range(3)])
        return l, h

    def refine(self, Uc_old, Uc_new, tfrac: float, lo, hi, Uf):
        l, h = self._box(lo, hi)
        new = _ptr_table(_dev_ptrs(Uc_new, self.ncomp)) if Uc_new is not None else None
        _check(self.lib.hb2_amr_refine_dev(C.byref(self.desc), _ptr_table(_dev_ptrs(Uc_old, self.ncomp)), new, C.c_double(tfrac),
                                           l, h, _ptr_table(_dev_ptrs(Uf, self.ncomp)), self._stream()), "hb2_amr_refine_dev")

    def coarsen(self, Uf, lo, hi, Uc):
        l, h = self._box(lo, hi)
        _check(self.lib.hb2_amr_coarsen_dev(C.byref(self.desc), _ptr_table(_dev_ptrs(Uf, self.ncomp)), l, h,
                                            _ptr_table(_dev_ptrs(Uc, self.ncomp)), self._stream()), "hb2_amr_coarsen_dev")

    def fluxsum_update(self, F_fine, fluxsum):
        """F_fine: list per direction of (neq, ...) side tensors; fluxsum: list [2 dir + side] of (neq, tangential cells)."""
        _check(self.lib.hb2_amr_fluxsum_update_dev(C.byref(self.desc), _ptr_table(_dev_ptrs(F_fine, self.neq)),
                                                   _ptr_table(_dev_ptrs(fluxsum, self.neq)), self._stream()),
               "hb2_amr_fluxsum_update_dev")

    def coarsen_fluxsum(self, fluxsum, F_coarse):
        _check(self.lib.hb2_amr_coarsen_fluxsum_dev(C.byref(self.desc), _ptr_table(_dev_ptrs(fluxsum, self.neq)),
                                                    _ptr_table(_dev_ptrs(F_coarse, self.neq)), self._stream()),
               "hb2_amr_coarsen_fluxsum_dev")


class DeviceLevel:
    """Device-resident multi-patch level (mirrors hb2_level_t): all patches of one rank registered once.  boxes: list of
    (lo, hi) in level index space, hi exclusive."""

    def __init__(self, dim, boxes, level_n, periodic_mask=None, flow_model=SINGLE_SPECIES, species_gamma=(1.4,), species_R=(),
                 dx=(1.0, 1.0, 1.0), math=MATH_EXACT, scheme=0):
        self.lib = load_library()
        d = PatchDescC()
        d.dim = dim
        for a in range(3):
            d.n[a] = 1
            d.dx[a] = float(dx[a]) if a < dim else 1.0
        d.flow_model = flow_model
        d.num_species = len(species_gamma) if flow_model != SINGLE_SPECIES else 1
        for i, g in enumerate(species_gamma):
            d.species_gamma[i] = float(g)
        for i, r in enumerate(species_R):
            d.species_R[i] = float(r)
        d.weno_p, d.math, d.device, d.scheme = 2, math, -1, int(scheme)
        self.dim, self.boxes = dim, [(tuple(lo), tuple(hi)) for lo, hi in boxes]
        np_ = len(self.boxes)
        lo = (C.c_int32 * (3 * np_))(*[int(b[0][a]) if a < dim else 0 for b in self.boxes for a in range(3)])
        hi = (C.c_int32 * (3 * np_))(*[int(b[1][a]) if a < dim else 1 for b in self.boxes for a in range(3)])
        ln = (C.c_int32 * 3)(*[int(level_n[a]) if a < dim else 1 for a in range(3)])
        mask = (1 << dim) - 1 if periodic_mask is None else int(periodic_mask)
        self._h = C.c_void_p()
        _check(self.lib.hb2_level_create(C.byref(d), np_, lo, hi, ln, mask, C.byref(self._h)), "hb2_level_create")
        neq, ncomp = C.c_int32(), C.c_int32()
        self.lib.hb2_num_comp(C.byref(d), C.byref(ncomp))
        self.lib.hb2_num_eqn(C.byref(d), C.byref(neq))
        self.ncomp, self.neq = ncomp.value, neq.value

    def ghost_shape(self, p):
        lo, hi = self.boxes[p]
        return tuple(hi[a] - lo[a] + 2 * GHOSTS for a in reversed(range(self.dim)))

    def close(self):
        if self._h:
            self.lib.hb2_level_destroy(self._h)
            self._h = None

    def _host_table(self, arrays):
        """arrays: list per patch of C-contiguous float64 numpy arrays (ncomp, *ghost_shape)"""
        tab = (C.c_void_p * (len(arrays) * self.ncomp))()
        for p, a in enumerate(arrays):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.shape == (self.ncomp,) + self.ghost_shape(p)
            stride = a[0].size * 8
            for c in range(self.ncomp):
                tab[p * self.ncomp + c] = a.ctypes.data + c * stride
        return tab

    def upload(self, arrays):
        tab = self._host_table(arrays)
        for p in range(len(arrays)):
            sub = (C.c_void_p * self.ncomp)(*[tab[p * self.ncomp + c] for c in range(self.ncomp)])
            _check(self.lib.hb2_level_upload_patch(self._h, p, sub), "hb2_level_upload_patch")

    def download(self, arrays):
        tab = self._host_table(arrays)
        for p in range(len(arrays)):
            sub = (C.c_void_p * self.ncomp)(*[tab[p * self.ncomp + c] for c in range(self.ncomp)])
            _check(self.lib.hb2_level_download_patch(self._h, p, sub), "hb2_level_download_patch")

    def advance(self, dt, alpha=None, beta=None):
        a = np.ascontiguousarray(SSPRK3_ALPHA if alpha is None else alpha, dtype=np.float64)
        b = np.ascontiguousarray(SSPRK3_BETA if beta is None else beta, dtype=np.float64)
        _check(self.lib.hb2_level_advance(self._h, a.shape[0], a.ctypes.data_as(C.POINTER(C.c_double)),
                                          b.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(dt)), "hb2_level_advance")

    def advance_host(self, arrays, dt, alpha=None, beta=None):
        """One RK step on host arrays (pinned memory for asynchronous copies), pipelined over the patches; in place."""
        a = np.ascontiguousarray(SSPRK3_ALPHA if alpha is None else alpha, dtype=np.float64)
        b = np.ascontiguousarray(SSPRK3_BETA if beta is None else beta, dtype=np.float64)
        _check(self.lib.hb2_level_advance_host(self._h, a.shape[0], a.ctypes.data_as(C.POINTER(C.c_double)),
                                               b.ctypes.data_as(C.POINTER(C.c_double)), C.c_double(dt), self._host_table(arrays)),
               "hb2_level_advance_host")

    def synchronize(self):
        _check(self.lib.hb2_level_synchronize(self._h), "hb2_level_synchronize")

    @property
    def launch_count(self):
        return int(self.lib.hb2_level_launch_count(self._h))


class DeviceArray:
    """A cudaMalloc'ed float64 array that torch can view (torch.as_tensor(DeviceArray)) and peers can open over IPC."""

    def __init__(self, shape):
        self.shape = tuple(int(x) for x in shape)
        self.nbytes = 8 * int(np.prod(self.shape))
        self.ptr = device_malloc(self.nbytes)
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f8", "data": (self.ptr, False), "version": 3,
                                         "strides": None}

    def free(self):
        if self.ptr:
            device_free(self.ptr)
            self.ptr = 0


def constants():
    """The hard-switch constants compiled into the kernels (hb2_constants); no device needed."""
    out = (C.c_double * 7)()
    _check(load_library().hb2_constants(out), "hb2_constants")
    return list(out)


def device_count() -> int:
    n = C.c_int32()
    load_library().hb2_device_count(C.byref(n))
    return n.value


def probe_fp64_peak(device: int = -1, seconds_hint: float = 1.0) -> float:
    v = C.c_double()
    _check(load_library().hb2_probe_fp64_peak(int(device), C.c_double(seconds_hint), C.byref(v)), "hb2_probe_fp64_peak")
    return v.value


def probe_hbm_bandwidth(device: int = -1, nbytes: int = 1 << 30) -> float:
    v = C.c_double()
    _check(load_library().hb2_probe_hbm_bandwidth(int(device), C.c_int64(nbytes), C.byref(v)), "hb2_probe_hbm_bandwidth")
    return v.value
