"""Python host-side mirror of the reference's RD classes over the C ABI.

``RDHandle`` is a thin, pointer-level wrapper of ``glia_rd_t``; the mirror classes in
``glia_b200.host`` (SpectralOperators, DiffCoef, DiffusionSolver, PdeOperatorsRD,
DerivativeOperatorsRD) are built on top of it and keep the reference's method names.

Field arguments are anything exposing a device pointer: ``torch`` CUDA tensors
(``data_ptr()``) in the product, raw ``int`` addresses, or -- only in the emulator tests
-- NumPy arrays.  All work is done by libglia_rd.so; there is no Python/CPU compute path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    raise TypeError(f"cannot take a device pointer from {type(x)}")


class RDHandle:
    """``glia_rd_t``.  With ``nranks > 1`` this is one rank of a slab-decomposed handle
    (``glia_rd_create_slab``): ``n`` is the global grid, every field is this rank's local block
    ``[n0/nranks][n1][n2]``, every method is collective, and ``all_gather`` -- a callable
    ``bytes -> list[bytes]`` in rank order, e.g. built on ``torch.distributed.all_gather_object``
    -- carries the 64-byte IPC handles between the ranks at set-up."""

    def __init__(self, n, precision="f32", device=0, dt_ctx=0.5, lib_path=None, rank=0, nranks=1, all_gather=None,
                 nbatch=1):
        self.lib = _capi.load_library(lib_path)
        self.n = tuple(int(v) for v in (n if hasattr(n, "__len__") else (n, n, n)))
        self.precision = {"f32": 4, "f64": 8, 4: 4, 8: 8}[precision]
        self.np_dtype = np.float32 if self.precision == 4 else np.float64
        self.rank, self.nranks, self._all_gather = int(rank), int(nranks), all_gather
        if self.nranks > 1 and all_gather is None:
            raise ValueError("a slab handle needs all_gather(bytes) -> list[bytes] to exchange its IPC handles")
        self._h = C.c_void_p()
        arr = (C.c_int * 3)(*self.n)
        self.nbatch = int(nbatch)
        if self.nbatch > 1:   # ensemble handle: fields are [nbatch][n0][n1][n2]
            rc = self.lib.glia_rd_create_batch(C.byref(self._h), arr, self.precision, int(device), float(dt_ctx),
                                               self.nbatch)
        else:
            rc = self.lib.glia_rd_create_slab(C.byref(self._h), arr, self.precision, int(device), float(dt_ctx),
                                              self.rank, self.nranks)
        if rc != 0:
            msg = self.lib.glia_rd_last_error(self._h).decode() if self._h else "create failed"
            if self._h:
                self.lib.glia_rd_destroy(self._h)
                self._h = C.c_void_p()
            raise _capi.GliaRdError(msg)
        if self.nranks > 1:
            self._connect(0)

    def _connect(self, which):
        """collective: export this rank's arena, gather everybody's, map the peers."""
        mine = C.create_string_buffer(64)
        self._ck(self.lib.glia_rd_ipc_export(self._h, int(which), mine))
        blobs = self._all_gather(mine.raw)
        assert len(blobs) == self.nranks and all(len(b) == 64 for b in blobs)
        allh = C.create_string_buffer(b"".join(blobs), 64 * self.nranks)
        self._ck(self.lib.glia_rd_ipc_connect(self._h, int(which), allh))
        self._all_gather(b"ok")  # nobody proceeds before every rank has mapped its peers

    @property
    def local_shape(self):
        return (self.n[0] // self.nranks, self.n[1], self.n[2])

    # -- plumbing ---------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise _capi.GliaRdError(self.lib.glia_rd_last_error(self._h).decode())

    def close(self, collective=True):
        """Slab handles: every rank closes its mappings of the peers' arenas, all ranks meet, then each frees
        its own (disconnect -> barrier -> free, include/glia_rd.h)."""
        if getattr(self, "_h", None):
            if collective and getattr(self, "nranks", 1) > 1 and self._all_gather is not None:
                try:
                    self.lib.glia_rd_ipc_disconnect(self._h, -1)
                    self._all_gather(b"bye")
                except Exception:
                    pass
            self.lib.glia_rd_destroy(self._h)
            self._h = C.c_void_p()

    def wait_stream(self, stream=None):
        """Order work already enqueued on `stream` (a cudaStream_t address, a torch.cuda.Stream, or None for
        the legacy default stream) in front of this handle's later work."""
        s = getattr(stream, "cuda_stream", stream)
        self._ck(self.lib.glia_rd_wait_stream(self._h, int(s) if s else None))

    def __del__(self):
        try:
            self.close(collective=False)
        except Exception:
            pass

    @property
    def nreal(self):
        return self.n[0] * self.n[1] * self.n[2] // self.nranks * getattr(self, "nbatch", 1)

    @property
    def launch_count(self):
        return int(self.lib.glia_rd_launch_count(self._h))

    @property
    def stream(self):
        return self.lib.glia_rd_stream(self._h)

    # -- L0 -------------------------------------------------------------------------
    def fft_r2c(self, f, fhat):
        self._ck(self.lib.glia_rd_fft_r2c(self._h, _ptr(f), _ptr(fhat)))

    def fft_c2r(self, fhat, f):
        self._ck(self.lib.glia_rd_fft_c2r(self._h, _ptr(fhat), _ptr(f)))

    def gradient(self, gx, gy, gz, x, mask=7):
        self._ck(self.lib.glia_rd_gradient(self._h, _ptr(gx), _ptr(gy), _ptr(gz), _ptr(x), int(mask)))

    def divergence(self, div, dx, dy, dz):
        self._ck(self.lib.glia_rd_divergence(self._h, _ptr(div), _ptr(dx), _ptr(dy), _ptr(dz)))

    # -- L1 -------------------------------------------------------------------------
    def set_diffusion(self, k, kavg, k_scale):
        ka = (C.c_double * 3)(*[float(v) for v in kavg])
        self._ck(self.lib.glia_rd_set_diffusion(self._h, _ptr(k), ka, float(k_scale)))

    def set_diffusion_tissue(self, wm, gm, csf, k_scale, k_gm_wm, k_glm_wm, filter_sum):
        self._ck(self.lib.glia_rd_set_diffusion_tissue(self._h, _ptr(wm), _ptr(gm), _ptr(csf), float(k_scale),
                                                       float(k_gm_wm), float(k_glm_wm), float(filter_sum)))

    def set_secondary_k(self, kt):
        self._ck(self.lib.glia_rd_set_secondary_k(self._h, _ptr(kt)))

    def set_reaction(self, rho):
        self._ck(self.lib.glia_rd_set_reaction(self._h, _ptr(rho)))

    def set_reaction_tissue(self, wm, gm, csf, rho_scale, r_gm_wm, r_glm_wm):
        self._ck(self.lib.glia_rd_set_reaction_tissue(self._h, _ptr(wm), _ptr(gm), _ptr(csf), float(rho_scale),
                                                      float(r_gm_wm), float(r_glm_wm)))

    def set_coefficients_batch(self, wm, gm, csf, k_scales, k_gm_wm, k_glm_wm, filter_sum, rho_scales, r_gm_wm, r_glm_wm):
        """Ensemble handle: one (kappa, rho) pair per member over ONE member's tissue maps."""
        ks = (C.c_double * self.nbatch)(*[float(v) for v in k_scales])
        rs = (C.c_double * self.nbatch)(*[float(v) for v in rho_scales])
        self._ck(self.lib.glia_rd_set_coefficients_batch(self._h, _ptr(wm), _ptr(gm), _ptr(csf), ks, float(k_gm_wm),
                                                         float(k_glm_wm), float(filter_sum), rs, float(r_gm_wm),
                                                         float(r_glm_wm)))

    def batch_iterations(self, accumulated=True):
        out = (C.c_int * self.nbatch)()
        self._ck(self.lib.glia_rd_batch_iterations(self._h, out, int(bool(accumulated))))
        return list(out)

    def update_reac_diff(self, bg, gm, vt, csf, rho_scale, k_scale, gm_r_scale, gm_k_scale):
        """PdeOperatorsMassEffect::updateReacAndDiffCoefficients (src/pde/PdeOperatorsMassEffect.cpp:98-138)."""
        self._ck(self.lib.glia_rd_update_reac_diff(self._h, _ptr(bg), _ptr(gm), _ptr(vt), _ptr(csf), float(rho_scale),
                                                   float(k_scale), float(gm_r_scale), float(gm_k_scale)))

    def apply_D(self, dc, c, secondary=False):
        self._ck(self.lib.glia_rd_apply_D(self._h, _ptr(dc), _ptr(c), int(bool(secondary))))

    # -- L2 -------------------------------------------------------------------------
    def prec_factor(self):
        self._ck(self.lib.glia_rd_prec_factor(self._h))

    def diffusion_solve(self, c, dt):
        its = C.c_int(0)
        self._ck(self.lib.glia_rd_diffusion_solve(self._h, _ptr(c), float(dt), C.byref(its)))
        return its.value

    def set_ksp_tolerances(self, rtol=1e-6, abstol=1e-50, dtol=1e4, maxit=5000):
        self._ck(self.lib.glia_rd_set_ksp_tolerances(self._h, rtol, abstol, dtol, int(maxit)))

    # -- L2a ------------------------------------------------------------------------
    def set_splitting_order(self, order):
        self._ck(self.lib.glia_rd_set_splitting_order(self._h, int(order)))

    def resize_history(self, nt, dt):
        if self.nranks > 1:   # nobody may still map the old history arena when it is freed
            self._ck(self.lib.glia_rd_ipc_disconnect(self._h, 1))
            self._all_gather(b"resize")
        self._ck(self.lib.glia_rd_resize_history(self._h, int(nt), float(dt)))
        self.nt, self.dt = int(nt), float(dt)
        if self.nranks > 1:
            self._connect(1)

    def history_ptr(self, which, i):
        p = C.c_void_p()
        self._ck(self.lib.glia_rd_history(self._h, int(which), int(i), C.byref(p)))
        return p.value

    def reaction(self, c_t, c_lin, dt):
        self._ck(self.lib.glia_rd_reaction(self._h, _ptr(c_t), _ptr(c_lin), float(dt)))

    def solve_state(self, c0, cT=None, linearized=0):
        its = C.c_int(0)
        self._ck(self.lib.glia_rd_solve_state(self._h, _ptr(c0), _ptr(cT), int(linearized), C.byref(its)))
        return its.value

    def solve_adjoint(self, pT, p0=None, linearized=1, adjoint_store=True):
        its = C.c_int(0)
        self._ck(self.lib.glia_rd_solve_adjoint(self._h, _ptr(pT), _ptr(p0), int(linearized),
                                                int(bool(adjoint_store)), C.byref(its)))
        return its.value

    # -- L2b ------------------------------------------------------------------------
    def grad_kappa_rho(self, wm, gm, csf):
        out = (C.c_double * 6)()
        self._ck(self.lib.glia_rd_grad_kappa_rho(self._h, _ptr(wm), _ptr(gm), _ptr(csf), out))
        return np.array(list(out))

    def set_secondary_tissue(self, wm, gm, csf, k1, k2=0.0, k3=0.0):
        self._ck(self.lib.glia_rd_set_secondary_tissue(self._h, _ptr(wm), _ptr(gm), _ptr(csf), float(k1), float(k2),
                                                       float(k3)))

    def set_two_snapshot(self, d0, obs0=None):
        """two_time_points_: data at t = 0 and its observation mask (copied); d0 = None switches it off."""
        self._ck(self.lib.glia_rd_set_two_snapshot(self._h, _ptr(d0), _ptr(obs0)))

    def objective_gradient(self, c0, d1, wm, gm, csf, obs=None, beta=0.0, g_c0=None):
        """evaluateObjectiveAndGradient in field space -> dict(J, mismatch, reg, mismatch0, g6, its)."""
        J, g, its = (C.c_double * 4)(), (C.c_double * 6)(), (C.c_int * 2)()
        self._ck(self.lib.glia_rd_objective_gradient(self._h, _ptr(c0), _ptr(d1), _ptr(obs), float(beta), _ptr(wm),
                                                     _ptr(gm), _ptr(csf), J, _ptr(g_c0), g, its))
        return dict(J=J[0], mismatch=J[1], reg=J[2], mismatch0=J[3], g6=np.array(list(g)), its=(its[0], its[1]))

    def hessian_matvec(self, c0_tilde, y_c0, wm, gm, csf, obs=None, beta=0.0, diffusivity_inversion=False):
        """evaluateHessian in field space -> (hk[6], ksp_its[4]); y_c0 is filled."""
        hk, its = (C.c_double * 6)(), (C.c_int * 4)()
        self._ck(self.lib.glia_rd_hessian_matvec(self._h, _ptr(c0_tilde), _ptr(obs), float(beta),
                                                 int(bool(diffusivity_inversion)), _ptr(wm), _ptr(gm), _ptr(csf),
                                                 _ptr(y_c0), hk, its))
        return np.array(list(hk)), list(its)

    def probe_xsweep(self, what, local_mask, reps=20):
        ms = C.c_double(0)
        self._ck(self.lib.glia_rd_probe_xsweep(self._h, int(what), int(local_mask), int(reps), C.byref(ms)))
        return ms.value

    # -- smoother / MatProp / Phi (callers either side of the path) ------------------------
    def smooth(self, out, inp, sigma):
        """SpectralOperators::weierstrassSmoother; ``out`` may alias ``inp``."""
        self._ck(self.lib.glia_rd_smooth(self._h, _ptr(out), _ptr(inp), float(sigma)))

    def mat_prop(self, gm, wm, vt, csf, bg=None, filt=None):
        """MatProp::setValuesCustom: clips the maps in place, fills bg / filter; -> sum(filter)."""
        s = C.c_double(0)
        self._ck(self.lib.glia_rd_mat_prop(self._h, _ptr(gm), _ptr(wm), _ptr(vt), _ptr(csf), _ptr(bg), _ptr(filt),
                                           C.byref(s)))
        return s.value

    def phi_set(self, centers, sigma_phi, filt=None, sigma_smooth=0.0):
        ctr = np.ascontiguousarray(np.asarray(centers, dtype=np.float64).reshape(-1, 3))
        self._phi_np = ctr.shape[0]
        self._ck(self.lib.glia_rd_phi_set(self._h, self._phi_np, ctr.ctypes.data_as(C.POINTER(C.c_double)),
                                          float(sigma_phi), _ptr(filt), float(sigma_smooth)))

    def phi_apply(self, out, p):
        pv = np.ascontiguousarray(np.asarray(p, dtype=np.float64).ravel())
        assert pv.size == self._phi_np
        self._ck(self.lib.glia_rd_phi_apply(self._h, _ptr(out), pv.ctypes.data_as(C.POINTER(C.c_double))))

    def phi_apply_transpose(self, inp):
        out = np.zeros(self._phi_np, dtype=np.float64)
        self._ck(self.lib.glia_rd_phi_apply_transpose(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), _ptr(inp)))
        return out

    # -- data in / out, segmentation labels ------------------------------------------------
    def data_in(self, path, field):
        """dataIn: NetCDF classic file (variable ``data``, dims x y z) -> device field (local rows)."""
        self._ck(self.lib.glia_rd_data_in(self._h, str(path).encode(), _ptr(field)))

    def data_out(self, path, field):
        """dataOut: device field -> CDF-2 file laid out like the reference's."""
        self._ck(self.lib.glia_rd_data_out(self._h, str(path).encode(), _ptr(field)))

    def split_segmentation(self, seg, labels, wm=None, gm=None, vt=None, csf=None):
        """splitSegmentation (atlas form): labels = (wm, gm, vt, csf)."""
        lab = (C.c_int * 4)(*[int(v) for v in labels])
        self._ck(self.lib.glia_rd_split_segmentation(self._h, _ptr(seg), lab, _ptr(wm), _ptr(gm), _ptr(vt), _ptr(csf)))

    # -- per-kernel profile -----------------------------------------------------------
    def profile_begin(self):
        self._ck(self.lib.glia_rd_profile_begin(self._h))

    def profile_end(self):
        """-> {tag: (launches, total_ms)}"""
        buf = C.create_string_buffer(16384)
        self._ck(self.lib.glia_rd_profile_end(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            tag, cnt, ms = line.split()
            out[tag] = (int(cnt), float(ms))
        return out

    # -- timing / host entry ---------------------------------------------------------
    def timer_start(self):
        self._ck(self.lib.glia_rd_timer_start(self._h))

    def timer_stop_ms(self):
        ms = C.c_double(0)
        self._ck(self.lib.glia_rd_timer_stop_ms(self._h, C.byref(ms)))
        return ms.value

    def forward_adjoint(self, c0, d1, cT=None, p0=None):
        """Device-buffer forward + adjoint solve; returns (ksp_its_state, ksp_its_adjoint)."""
        ks, ka = C.c_int(0), C.c_int(0)
        self._ck(self.lib.glia_rd_forward_adjoint(self._h, _ptr(c0), _ptr(d1), _ptr(cT), _ptr(p0),
                                                  C.byref(ks), C.byref(ka)))
        return ks.value, ka.value

    def forward_adjoint_host(self, c0, d1, cT, p0):
        """Host-buffer (NumPy) end-to-end call: H2D, forward, adjoint, D2H inside."""
        for a in (c0, d1, cT, p0):
            assert isinstance(a, np.ndarray) and a.dtype == self.np_dtype and a.flags["C_CONTIGUOUS"]
        ks, ka = C.c_int(0), C.c_int(0)
        self._ck(self.lib.glia_rd_forward_adjoint_host(self._h, c0.ctypes.data, d1.ctypes.data, cT.ctypes.data,
                                                       p0.ctypes.data, C.byref(ks), C.byref(ka)))
        return ks.value, ka.value
