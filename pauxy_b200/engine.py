"""Device-side walker batch: thin Python owner of one pxb handle.

PyTorch is used for device memory (one arena tensor), streams and copies
only; every computation is a call into libpauxy_b200.so.  Walker state is a
structure of arrays in the arena, exposed here as torch views.
"""
import ctypes

import numpy
import torch

from . import _lib as L


def _round_up(x, m):
    return (x + m - 1) // m * m


class Engine(object):
    """One device's share of the walker population.

    Parameters mirror system.nbasis / nup / ndown / nfields, qmc.dt and
    propagator.exp_nmax of the reference (pauxy/propagation/continuous.py:37-40).
    """

    def __init__(self, nbasis, nup, ndown, nchol, nwalkers, dt, exp_order=6, device=None,
                 total_walkers=None, exchange='auto', free_projection=False, force_bias=True,
                 nbp=0, ndets=1, local_energy_weight=False, complex_one_body=False,
                 complex_cholesky=False):
        if not torch.cuda.is_available():
            raise RuntimeError("pauxy_b200.Engine needs a CUDA device (no CPU fallback)")
        self.lib = L.load()
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(device)
        self.M, self.na, self.nb, self.N, self.W = nbasis, nup, ndown, nchol, nwalkers
        self.ne = nup + ndown
        self.Wtot = total_walkers if total_walkers else nwalkers
        self.dt = dt
        self.Wp = _round_up(nwalkers, 4)
        self.Np = _round_up(nchol, 8)
        cfg = L.PxbConfig(nbasis, nup, ndown, nchol, nwalkers, exp_order,
                          self.device.index or 0, self.Wtot, dt, L.EXCHANGE_MODES[exchange],
                          (L.FLAG_FREE_PROJECTION if free_projection else 0) |
                          (0 if force_bias else L.FLAG_NO_FORCE_BIAS) |
                          (L.FLAG_LOCAL_ENERGY_WEIGHT if local_energy_weight else 0) |
                          (L.FLAG_COMPLEX_ONE_BODY if complex_one_body else 0) |
                          (L.FLAG_COMPLEX_CHOLESKY if complex_cholesky else 0), int(nbp or 0),
                          int(ndets))
        if ndets > L.MAX_DETS:
            raise L.PxbError(-4, "at most %d determinants in the trial" % L.MAX_DETS)
        self.ndets = int(ndets)
        self.complex_cholesky = bool(complex_cholesky)
        self._eshift_im = 0.0
        self.nbp = int(nbp or 0)
        self._h = ctypes.c_void_p()
        rc = self.lib.pxb_create(ctypes.byref(self._h), ctypes.byref(cfg))
        if rc != 0:
            raise L.PxbError(rc, "pxb_create: bad configuration")
        nbytes = ctypes.c_size_t()
        self._check(self.lib.pxb_arena_bytes(self._h, ctypes.byref(nbytes)))
        self.arena_bytes = nbytes.value
        with torch.cuda.device(self.device):
            self.arena = torch.empty(self.arena_bytes + 256, dtype=torch.uint8, device=self.device)
            base = self.arena.data_ptr()
            self._shift = (-base) % 256
            self._check(self.lib.pxb_bind_arena(self._h, base + self._shift, self.arena_bytes,
                                                self._stream()))
        self._make_views()
        self._xi_pinned = None
        self._xi_dev = None
        self.peers_attached = False

    # ------------------------------------------------------------------ utils
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.pxb_last_error(self._h)
            raise L.PxbError(rc, msg.decode() if msg else '')

    def _view(self, fid, dtype, shape=None):
        off, size = ctypes.c_size_t(), ctypes.c_size_t()
        self._check(self.lib.pxb_field(self._h, fid, ctypes.byref(off), ctypes.byref(size)))
        a = self._shift + off.value
        t = self.arena[a:a + size.value].view(dtype)
        return t if shape is None else t.view(shape)

    def _make_views(self):
        W, Wp, Np = self.W, self.Wp, self.Np
        f64, c128 = torch.float64, torch.complex128
        self.weight = self._view(L.F_WEIGHT, f64)[:W]
        self.unscaled_weight = self._view(L.F_UNSCALED_WEIGHT, f64)[:W]
        self.ot = self._view(L.F_OT, c128)[:W]
        self.hybrid_energy = self._view(L.F_HYBRID_ENERGY, c128)[:W]
        self.eloc = self._view(L.F_ELOC, c128, (Wp, 3))[:W]
        self.detR = self._view(L.F_DETR, f64)[:W]
        self.log_detR = self._view(L.F_LOG_DETR, f64)[:W]
        self.estimates = self._view(L.F_ESTIMATES, c128)
        self.counters = self._view(L.F_COUNTERS, torch.int64)
        self.parent_ix = self._view(L.F_PARENT_IX, torch.int32)
        self.xbar = self._view(L.F_XBAR, c128, (Wp, Np))[:W, :self.N]
        self.xshifted = self._view(L.F_XSHIFTED, c128, (Wp, Np))[:W, :self.N]
        self.cmf_cfb = self._view(L.F_CMF_CFB, c128, (Wp, 2))[:W]
        self.ovlp_new = self._view(L.F_OVLP_NEW, c128)[:W]
        self.total_weight = self._view(L.F_TOTAL_WEIGHT, f64)
        self.pairs = self._view(L.F_PAIRS, torch.int32)
        self.phase = self._view(L.F_PHASE, c128)[:W]
        if self.nbp > 0:
            self.bp_rdm = self._view(L.F_BP_RDM, c128, (2, self.M, self.M))
        self.bp_denom = self._view(L.F_BP_DENOM, c128)
        self.theta_sum = self._view(L.F_THETA_SUM, c128, (self.ne, self.M))
        self.walker_eloc = self._view(L.F_WALKER_ELOC, c128)[:W]
        ls = self._view(L.F_LOG_SHIFTS, f64)
        self.log_shifts = ls[:3]          # log_shift, detR_shift, log_detR_shift
        self.shift_sums = ls[8:11]        # population sums of |ot|, |detR|, |log_detR| (this device)
        # overlaps with the single determinants of the trial, [ndets, W] (MultiDetWalker.ovlps)
        od = self._view(L.F_OVLP_DET, c128)
        self.ovlp_det = od.view(self.ndets, od.numel() // self.ndets)[:, :W]

    def _dev(self, a, dtype):
        t = torch.as_tensor(numpy.ascontiguousarray(a, dtype=dtype))
        return t.to(self.device)

    def close(self):
        if self._h:
            self.lib.pxb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ setup
    def set_hamiltonian(self, hs_pot, rchol, bh1, h1rot, psi, mf_shift, ecore):
        """Upload the reference's arrays (numpy, reference layouts/dtypes)."""
        M, ne, N = self.M, self.ne, self.N
        assert hs_pot.shape == (M * M, N) and rchol.shape == (ne * M, N)
        assert bh1.shape == (2, M, M) and h1rot.shape == (ne, M) and psi.shape == (M, ne)
        if self.complex_cholesky:
            hs_dtype = numpy.complex128          # PXB_FLAG_COMPLEX_CHOLESKY: interleaved (re, im)
        else:
            hs_dtype = numpy.float64
            if numpy.iscomplexobj(hs_pot):
                if numpy.abs(hs_pot.imag).max() != 0.0:
                    raise L.PxbError(-4, "complex Cholesky vectors: create the Engine with complex_cholesky=True")
                hs_pot = hs_pot.real
        with torch.cuda.device(self.device):
            t = [self._dev(hs_pot, hs_dtype), self._dev(rchol, numpy.complex128),
                 self._dev(bh1, numpy.complex128), self._dev(h1rot, numpy.complex128),
                 self._dev(psi, numpy.complex128), self._dev(mf_shift, numpy.complex128)]
            self._check(self.lib.pxb_set_hamiltonian(self._h, *[x.data_ptr() for x in t],
                                                     float(numpy.real(ecore)), self._stream()))
        del t

    def set_trial_det(self, det, coeff, rchol=None, h1rot=None, psi=None):
        """Determinant `det` of a multi-determinant trial: its CI coefficient and (det >= 1, or to
        replace determinant 0) its half-rotated Cholesky vectors, one-body integrals and orbitals."""
        M, ne, N = self.M, self.ne, self.N
        c = complex(coeff)
        with torch.cuda.device(self.device):
            if rchol is None:
                self._check(self.lib.pxb_set_trial_det(self._h, int(det), c.real, c.imag, None, None,
                                                       None, self._stream()))
                return
            assert rchol.shape == (ne * M, N) and h1rot.shape == (ne, M) and psi.shape == (M, ne)
            t = [self._dev(rchol, numpy.complex128), self._dev(h1rot, numpy.complex128),
                 self._dev(psi, numpy.complex128)]
            self._check(self.lib.pxb_set_trial_det(self._h, int(det), c.real, c.imag,
                                                   *[x.data_ptr() for x in t], self._stream()))
        del t

    def _set_eshift(self, eshift):
        """Real part of the (possibly complex) energy shift; its imaginary part goes through
        pxb_set_eshift_imag when it changes."""
        e = complex(eshift)
        if e.imag != self._eshift_im:
            self._check(self.lib.pxb_set_eshift_imag(self._h, e.imag))
            self._eshift_im = e.imag
        return e.real

    def init_walkers(self, init_phi, total_walkers=None):
        tw = float(self.Wtot if total_walkers is None else total_walkers)
        with torch.cuda.device(self.device):
            t = self._dev(init_phi, numpy.complex128)
            self._check(self.lib.pxb_init_walkers(self._h, t.data_ptr(), tw, self._stream()))
            torch.cuda.current_stream(self.device).synchronize()

    # ------------------------------------------------------------------ state
    def set_phi(self, phi):
        """phi: [W, M, ne] complex128 (numpy or cuda tensor)."""
        t = phi if torch.is_tensor(phi) else self._dev(phi, numpy.complex128)
        assert tuple(t.shape) == (self.W, self.M, self.ne) and t.is_contiguous()
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_set_phi(self._h, t.data_ptr(), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()

    def get_phi(self):
        with torch.cuda.device(self.device):
            out = torch.empty((self.W, self.M, self.ne), dtype=torch.complex128, device=self.device)
            self._check(self.lib.pxb_get_phi(self._h, out.data_ptr(), self._stream()))
        return out

    # --------------------------------------------------------------- hot path
    def stage_xi(self, xi_host):
        """Copy host fields [W, N] float64 through pinned memory; returns the device tensor."""
        if self._xi_pinned is None:
            self._xi_pinned = torch.empty((self.W, self.N), dtype=torch.float64).pin_memory()
            self._xi_dev = torch.empty((self.W, self.N), dtype=torch.float64, device=self.device)
            self._xi_event = torch.cuda.Event()
        else:
            self._xi_event.synchronize()   # the previous H2D copy has drained the pinned buffer
        self._xi_pinned.numpy()[...] = xi_host
        self._xi_dev.copy_(self._xi_pinned, non_blocking=True)
        self._xi_event.record(torch.cuda.current_stream(self.device))
        return self._xi_dev

    def prefetch_xi(self, xi_pinned):
        """Start the host->device copy of the NEXT step's fields (pinned float64 [W, N]) on a
        copy stream, double-buffered, so that it overlaps the step in flight.  Returns the device
        tensor to hand to propagate(), which waits for the copy on the launch stream."""
        if getattr(self, '_pf', None) is None:
            with torch.cuda.device(self.device):
                self._pf = {'buf': [torch.empty((self.W, self.N), dtype=torch.float64,
                                                device=self.device) for _ in range(2)],
                            'ready': [torch.cuda.Event(), torch.cuda.Event()],
                            'free': [None, None], 'next': 0,
                            'stream': torch.cuda.Stream(self.device)}
        pf = self._pf
        i = pf['next']
        pf['next'] = 1 - i
        with torch.cuda.device(self.device):
            if pf['free'][i] is not None:
                pf['stream'].wait_event(pf['free'][i])   # the step that read this buffer is done
            with torch.cuda.stream(pf['stream']):
                pf['buf'][i].copy_(xi_pinned, non_blocking=True)
                pf['ready'][i].record(pf['stream'])
        return pf['buf'][i]

    def _xi_pointer(self, xi):
        """Device pointer of the fields (None: device Philox) + the prefetch slot they sit in; the
        launch stream is made to wait for a prefetched copy."""
        if xi is None:
            return None, None
        if not torch.is_tensor(xi):
            xi = self.stage_xi(xi)
        assert xi.dtype == torch.float64 and tuple(xi.shape) == (self.W, self.N)
        assert xi.is_contiguous()
        ptr = xi.data_ptr()
        slot = None
        pf = getattr(self, '_pf', None)
        if pf is not None:
            for i in range(2):
                if pf['buf'][i].data_ptr() == ptr:
                    slot = i
                    torch.cuda.current_stream(self.device).wait_event(pf['ready'][i])
        return ptr, slot

    def _xi_release(self, slot):
        if slot is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._pf['free'][slot] = ev

    def propagate(self, xi=None, eshift=0.0, step=1, seed=0, walker_offset=0):
        """xi: None (device Philox), numpy [W,N] or cuda float64 tensor [W,N]."""
        with torch.cuda.device(self.device):
            ptr, slot = self._xi_pointer(xi)
            self._check(self.lib.pxb_propagate(self._h, ptr, int(seed), int(walker_offset),
                                               self._set_eshift(eshift), int(step), self._stream()))
            self._xi_release(slot)

    def step(self, xi=None, eshift=0.0, step=1, seed=0, walker_offset=0, comb_r=0.0,
             ortho=False, pop=True, energy=True):
        """One whole driver step on this device as ONE library call (pxb_step): [re-orthogonalise]
        -> propagate -> [comb, plan overlapped with the local energy] -> accumulate; replayed from a
        CUDA graph from the second occurrence of a variant on."""
        flags = (L.STEP_ORTHO if ortho else 0) | (L.STEP_POP if pop else 0) | \
            (L.STEP_ENERGY if energy else 0)
        with torch.cuda.device(self.device):
            ptr, slot = self._xi_pointer(xi)
            self._check(self.lib.pxb_step(self._h, ptr, int(seed), int(walker_offset), self._set_eshift(eshift),
                                          int(step), float(comb_r), flags, self._stream()))
            self._xi_release(slot)

    def step_graphs(self, enable=None):
        """Enable / disable CUDA-graph replay in step(); returns the number of graph launches so far."""
        n = ctypes.c_longlong()
        self._check(self.lib.pxb_step_graphs(self._h, -1 if enable is None else int(bool(enable)),
                                             ctypes.byref(n)))
        return n.value

    def orthogonalise(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_orthogonalise(self._h, self._stream()))

    def local_energy(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_local_energy(self._h, self._stream()))

    def accumulate(self, with_energy=True):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_accumulate(self._h, 1 if with_energy else 0, self._stream()))

    def accumulate_theta(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_accumulate_theta(self._h, self._stream()))

    def zero_estimates(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_zero_estimates(self._h, self._stream()))

    # ----------------------------------------------------- population control
    def log_shift_enable(self, on=True):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_log_shift_enable(self._h, 1 if on else 0, self._stream()))

    def update_log_shifts(self, comm=None):
        """Walkers.update_log_ovlp (walkers/handler.py:456-475): device sums, all-reduce, running average."""
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_log_shift_sums(self._h, self._stream()))
            if comm is not None and comm.size > 1:
                comm.allreduce_sum_(self.shift_sums)
            self._check(self.lib.pxb_log_shift_update(self._h, self._stream()))

    def pop_control_comb(self, r):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_pop_control_comb(self._h, float(r), self._stream()))

    def pop_rescale(self, global_abs_weights):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_pop_rescale(self._h, global_abs_weights.data_ptr(),
                                                 int(global_abs_weights.numel()), self._stream()))

    def comb_plan(self, global_abs_weights, r):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_comb_plan(self._h, global_abs_weights.data_ptr(),
                                               int(global_abs_weights.numel()), float(r),
                                               self._stream()))

    # peer-memory comb (several devices of one node)
    def attach_peers(self, comm):
        """Exchange CUDA IPC handles of the arenas over `comm` and map every peer's arena
        (pxb_peer_export / pxb_peer_attach).  Returns True when the peer path is usable."""
        self.peers_attached = False
        if comm is None or comm.size == 1:
            return False
        hbuf = (ctypes.c_ubyte * 64)()
        off = ctypes.c_uint64()
        with torch.cuda.device(self.device):
            rc = self.lib.pxb_peer_export(self._h, hbuf, ctypes.byref(off))
            mine = numpy.zeros(80, dtype=numpy.uint8)
            if rc == 0:
                mine[:64] = numpy.frombuffer(bytes(hbuf), dtype=numpy.uint8)
                mine[64:72] = numpy.frombuffer(int(off.value).to_bytes(8, 'little'), dtype=numpy.uint8)
                mine[72] = 1
            allh = comm.allgather_tensor(torch.from_numpy(mine).to(self.device)).cpu().numpy()
            allh = allh.reshape(comm.size, 80)
            if not bool(allh[:, 72].all()):
                return False
            handles = numpy.ascontiguousarray(allh[:, :64])
            offsets = numpy.ascontiguousarray(allh[:, 64:72]).view(numpy.uint64).reshape(-1).copy()
            rc = self.lib.pxb_peer_attach(self._h, comm.rank, comm.size, handles.ctypes.data,
                                          offsets.ctypes.data)
            ok = torch.tensor([1.0 if rc == 0 else 0.0], dtype=torch.float64, device=self.device)
            ok = comm.allgather_tensor(ok).cpu().numpy()
            self.peers_attached = bool((ok > 0).all())
        return self.peers_attached

    def pop_control_comb_peers(self, global_abs_weights, r):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_pop_control_comb_peers(
                self._h, global_abs_weights.data_ptr(), int(global_abs_weights.numel()), float(r),
                self._stream()))

    def pop_plan(self, global_abs_weights, r):
        """Total weight + comb plan on the CURRENT stream (global_abs_weights None: one device)."""
        with torch.cuda.device(self.device):
            if global_abs_weights is None:
                ptr, n = None, 0
            else:
                ptr, n = global_abs_weights.data_ptr(), int(global_abs_weights.numel())
            self._check(self.lib.pxb_pop_plan(self._h, ptr, n, float(r), self._stream()))

    def reserve_sms(self, n):
        self._check(self.lib.pxb_reserve_sms(self._h, int(n)))

    def pop_pull(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_pop_pull(self._h, self._stream()))

    def pop_control_finish(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_pop_control_finish(self._h, self._stream()))

    def payload_doubles(self):
        n = ctypes.c_size_t()
        self._check(self.lib.pxb_payload_doubles(self._h, ctypes.byref(n)))
        return n.value

    def copy_walkers(self, src, dst):
        """src, dst: int32 cuda tensors of local slots."""
        if src.numel() == 0:
            return
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_copy_walkers(self._h, src.data_ptr(), dst.data_ptr(),
                                                  int(src.numel()), self._stream()))

    def pack_walkers(self, slots, buf):
        if slots.numel() == 0:
            return
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_pack_walkers(self._h, slots.data_ptr(), int(slots.numel()),
                                                  buf.data_ptr(), self._stream()))

    def unpack_walkers(self, slots, buf):
        if slots.numel() == 0:
            return
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_unpack_walkers(self._h, slots.data_ptr(), int(slots.numel()),
                                                    buf.data_ptr(), self._stream()))

    def set_weights(self, value):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_set_weights(self._h, float(value), self._stream()))

    # ------------------------------------------------------- back propagation
    def bp_steps(self):
        return int(self.lib.pxb_bp_steps(self._h))

    def back_propagate(self, nsteps, nstblz, init_walker=False):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_back_propagate(self._h, int(nsteps), int(nstblz),
                                                    1 if init_walker else 0, self._stream()))

    def bp_restore_weights(self, mode):
        self._check(self.lib.pxb_bp_restore_weights(self._h, {None: 0, 'partial': 1, 'full': 2}[mode]))

    def bp_reset(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_bp_reset(self._h, self._stream()))

    def bp_zero(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_bp_zero(self._h, self._stream()))

    def get_phi_bp(self, historic=False):
        with torch.cuda.device(self.device):
            out = torch.empty((self.W, self.M, self.ne), dtype=torch.complex128, device=self.device)
            self._check(self.lib.pxb_get_phi_bp(self._h, 1 if historic else 0, out.data_ptr(),
                                                self._stream()))
        return out

    # ------------------------------------------------------------ stage access
    def stage_greens(self, with_e1b=False):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_stage_greens(self._h, 1 if with_e1b else 0, self._stream()))

    def stage_force_bias_gemm(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_stage_force_bias_gemm(self._h, self._stream()))

    def stage_exchange(self):
        with torch.cuda.device(self.device):
            self._check(self.lib.pxb_stage_exchange(self._h, self._stream()))

    def exchange_is_eri(self):
        return bool(self.lib.pxb_exchange_mode(self._h) == L.EXCHANGE_MODES['eri'])

    def vhs_is_symmetric(self):
        return bool(self.lib.pxb_vhs_symmetric(self._h) == 1)

    def profile(self, enable=True):
        """Bracket every stage with CUDA events on the launch stream (see stage_times)."""
        self._check(self.lib.pxb_profile(self._h, 1 if enable else 0))

    def stage_times(self, reset=True):
        """{stage: (milliseconds, calls)} accumulated since the last reset (waits for the GPU)."""
        n = len(L.STAGES)
        ms = (ctypes.c_double * n)()
        calls = (ctypes.c_longlong * n)()
        self._check(self.lib.pxb_stage_times(self._h, ms, calls, n, 1 if reset else 0))
        return dict((L.STAGES[i], (ms[i], calls[i])) for i in range(n))

    def launch_count(self):
        return int(self.lib.pxb_launch_count(self._h))

    def _get(self, fn, shape):
        with torch.cuda.device(self.device):
            out = torch.empty(shape, dtype=torch.complex128, device=self.device)
            self._check(fn(self._h, out.data_ptr(), self._stream()))
        return out

    def get_theta(self):
        return self._get(self.lib.pxb_get_theta, (self.W, self.ne, self.M))

    def get_x(self):
        return self._get(self.lib.pxb_get_x, (2, self.W, self.N))

    def get_vhs(self):
        return self._get(self.lib.pxb_get_vhs, (self.W, self.M, self.M))

    def get_exx(self):
        return self._get(self.lib.pxb_get_exx, (2, self.W))

    def synchronize(self):
        torch.cuda.current_stream(self.device).synchronize()


def comb_plan_host(weights, r):
    """Bit-exact host comb (pauxy/walkers/handler.py:271-286) via the C library."""
    lib = L.load()
    w = numpy.ascontiguousarray(weights, dtype=numpy.float64)
    out = numpy.zeros(len(w), dtype=numpy.int32)
    rc = lib.pxb_comb_plan_host(w.ctypes.data, len(w), float(r), out.ctypes.data)
    if rc != 0:
        raise L.PxbError(rc, "comb sweep ran past the last walker")
    return out
