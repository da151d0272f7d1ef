"""Walker population on the device with the reference's handler interface.

`Walkers` mirrors pauxy.walkers.handler.Walkers (handler.py:19-164:
construction, :166-181 orthogonalise, :225-412 pop_control / comb /
pair_branch); `SingleDetWalker` objects are thin views of one slot of the
structure-of-arrays state owned by the Engine (pauxy/walkers/walker.py:24-61,
single_det.py:31-94).  Device->host copies happen only on attribute access.
"""
import numpy
import scipy.linalg
import torch


def get_input_value(inputs, key, default=0, alias=None, verbose=False):
    """Option lookup with aliases (pauxy/utils/io.py:304-323)."""
    val = inputs.get(key, None)
    if val is None and alias is not None:
        for a in alias:
            val = inputs.get(a, None)
            if val is not None:
                break
    return default if val is None else val


class SingleDetWalker(object):
    """View of walker `index` of the device batch."""

    def __init__(self, handler, index):
        self._h = handler
        self._i = index
        self.nup = handler.system.nup
        self.ndown = handler.system.ndown
        self.le_oratio = 1.0
        self.alive = 1
        self.field_configs = None
        self.stack = None

    def _scalar(self, name):
        return getattr(self._h.engine, name)[self._i].item()

    weight = property(lambda s: s._scalar('weight'),
                      lambda s, v: s._h.engine.weight.__setitem__(s._i, float(v)))
    unscaled_weight = property(lambda s: s._scalar('unscaled_weight'),
                               lambda s, v: s._h.engine.unscaled_weight.__setitem__(s._i, float(v)))
    ot = property(lambda s: s._scalar('ot'),
                  lambda s, v: s._h.engine.ot.__setitem__(s._i, complex(v)))
    ovlp = ot
    hybrid_energy = property(lambda s: s._scalar('hybrid_energy'),
                             lambda s, v: s._h.engine.hybrid_energy.__setitem__(s._i, complex(v)))
    phase = property(lambda s: s._scalar('phase'),
                     lambda s, v: s._h.engine.phase.__setitem__(s._i, complex(v)))
    detR = property(lambda s: s._scalar('detR'))
    log_detR = property(lambda s: s._scalar('log_detR'))
    total_weight = property(lambda s: s._h.engine.total_weight[0].item())

    @property
    def phi(self):
        return self._h.phi_host()[self._i]

    @phi.setter
    def phi(self, value):
        phi = self._h.engine.get_phi()
        phi[self._i] = torch.as_tensor(numpy.asarray(value, dtype=numpy.complex128)).to(phi.device)
        self._h.engine.set_phi(phi)
        self._h._phi_cache = None

    @property
    def phi_old(self):
        """walker.phi_old (walkers/walker.py:43), kept on the device when back propagating."""
        return self._h.engine.get_phi_bp(historic=True)[self._i].cpu().numpy()

    @property
    def E_L(self):
        return self._h.engine.eloc[self._i, 0].item().real

    def local_energy(self, system, two_rdm=None, rchol=None, eri=None, UVT=None):
        """(E, T, V) of this walker from the batched device evaluation
        (walkers/single_det.py:340-364 -> estimators/generic.py:156-221)."""
        self._h.engine.local_energy()
        e = self._h.engine.eloc[self._i].cpu().numpy()
        return (complex(e[0]), complex(e[1]), complex(e[2]))

    def greens_function(self, trial):
        """Host-side G / Gmod of this walker for inspection (single_det.py:295-321);
        the device keeps its own Theta."""
        phi = self.phi
        nup = self.nup
        self.Gmod, self.G, dets = [], [], []
        for sl in (slice(0, nup), slice(nup, nup + self.ndown)):
            ovlp = numpy.dot(phi[:, sl].T, trial.psi[:, sl].conj())
            gmod = numpy.dot(scipy.linalg.inv(ovlp), phi[:, sl].T)
            self.Gmod.append(gmod)
            self.G.append(numpy.dot(trial.psi[:, sl].conj(), gmod))
            dets.append(numpy.linalg.slogdet(ovlp))
        self.G = numpy.array(self.G)
        return dets[0][0] * dets[1][0] * numpy.exp(dets[0][1] + dets[1][1])

    def get_buffer(self):
        """Minimal communication payload of this walker (walkers/walker.py:63-131
        packs every numeric attribute; only these matter downstream)."""
        eng = self._h.engine
        slot = torch.tensor([self._i], dtype=torch.int32, device=eng.device)
        buf = torch.empty(eng.payload_doubles(), dtype=torch.float64, device=eng.device)
        eng.pack_walkers(slot, buf)
        return buf.cpu().numpy()

    def set_buffer(self, buff):
        eng = self._h.engine
        slot = torch.tensor([self._i], dtype=torch.int32, device=eng.device)
        buf = torch.as_tensor(numpy.ascontiguousarray(buff, dtype=numpy.float64)).to(eng.device)
        eng.unpack_walkers(slot, buf)
        self._h._phi_cache = None


class MultiDetWalker(SingleDetWalker):
    """View of walker `index` for a multi-determinant trial (pauxy/walkers/multi_det.py:8-300): the
    overlaps with the single determinants and their weights conj(c_i) <psi_i|phi> are device
    fields; the walker's overlap `ot` is their sum."""

    def __init__(self, handler, index, trial):
        SingleDetWalker.__init__(self, handler, index)
        self.ndets = trial.ndets
        self._coeffs = numpy.array(trial.coeffs)

    @property
    def ovlps(self):
        return self._h.engine.ovlp_det[:, self._i].cpu().numpy()

    @property
    def weights(self):
        return self._coeffs.conj() * self.ovlps

    eloc = property(lambda s: s._scalar('walker_eloc'))

    def greens_function(self, trial):
        """Host-side per-determinant Green's functions Gi [ndets, 2, M, M] for inspection
        (multi_det.py:198-231); returns the total overlap."""
        phi = self.phi
        nup = self.nup
        M = phi.shape[0]
        self.Gi = numpy.zeros((self.ndets, 2, M, M), dtype=numpy.complex128)
        tot = 0.0
        for ix in range(self.ndets):
            det = trial.psi[ix]
            ovlp = 1.0
            for s, sl in enumerate((slice(0, nup), slice(nup, nup + self.ndown))):
                O = numpy.dot(phi[:, sl].T, det[:, sl].conj())
                ovlp = ovlp * scipy.linalg.det(O)
                self.Gi[ix, s] = numpy.dot(det[:, sl].conj(), numpy.dot(scipy.linalg.inv(O), phi[:, sl].T))
            tot += trial.coeffs[ix].conj() * ovlp
        return tot


class Walkers(object):
    """Container of the walkers owned by this rank (device batch)."""

    def __init__(self, system, trial, qmc, engine, walker_opts=None, verbose=False, comm=None,
                 nprop_tot=None, nbp=None):
        walker_opts = walker_opts or {}
        if nbp is not None and int(nbp) != engine.nbp:
            raise ValueError("Walkers: nbp differs from the engine's field-history capacity")
        self.system = system
        self.engine = engine
        self.nwalkers = qmc.nwalkers
        self.ntot_walkers = qmc.ntot_walkers
        self.rank = 0 if comm is None else comm.rank
        self.walker_offset = self.rank * self.nwalkers  # global index rank*nw + i (handler.py:303-321)
        # walker restart (handler.py:43-47,148-161,432-485): same options and the same per-walker
        # record [weight, phase, ot, phi.ravel()] as the reference's 'walker_%d' datasets; the
        # container is one .npy per rank instead of an MPI-IO HDF5 file (no h5py in this image)
        self.write_freq = walker_opts.get('write_freq', 0)
        self.write_file = walker_opts.get('write_file', 'restart.h5')
        self.read_file = walker_opts.get('read_file', None)
        self.write_restart = self.write_freq > 0
        self.use_log_shift = walker_opts.get('use_log_shift', False)
        self.walker_type = 'SD' if trial.ndets == 1 else 'MSD'
        self.pcont_method = get_input_value(walker_opts, 'population_control', default='comb')
        self.min_weight = walker_opts.get('min_weight', 0.1)
        self.max_weight = walker_opts.get('max_weight', 4.0)
        # several devices: pull clones straight from peer memory (False: NCCL send/recv)
        self.peer_copy = walker_opts.get('peer_copy', True)
        # evaluate the local energy concurrently with the comb plan when the driver asks for it
        self.overlap = walker_opts.get('overlap_energy', True)
        self.reserve_sm_walkers = walker_opts.get('reserve_sm_walkers', 24576)
        self._side = None
        self.target_weight = qmc.ntot_walkers
        self.nw = qmc.nwalkers
        engine.init_walkers(trial.init, qmc.ntot_walkers)
        if self.use_log_shift:
            engine.log_shift_enable(True)
        if self.peer_copy and comm is not None and comm.size > 1:
            engine.attach_peers(comm)
        if trial.ndets == 1:
            self.walkers = [SingleDetWalker(self, i) for i in range(self.nwalkers)]
        else:               # handler.py:64-70
            self.walkers = [MultiDetWalker(self, i, trial) for i in range(self.nwalkers)]
        if self.read_file is not None:
            self.read_walkers(comm)
        self.buff_size = engine.payload_doubles()
        self._phi_cache = None
        self.last_parent_ix = None

    def phi_host(self):
        return self.engine.get_phi().cpu().numpy()

    # ----------------------------------------------------------------- fields
    def draw_fields(self, nfields, comm=None):
        """Auxiliary fields for this step from the GLOBAL legacy numpy stream in
        global walker order, one normal(size=(n_active, N)) block -- identical
        to the reference's per-walker draws (continuous.py:133).
        With several ranks every rank draws the whole block and keeps its rows,
        so results do not depend on the number of devices."""
        eng = self.engine
        w = eng.weight
        if comm is not None and comm.size > 1:
            gw = comm.allgather_tensor(w).cpu().numpy()
        else:
            gw = w.cpu().numpy()
        active = numpy.abs(gw) > 1e-8
        xi_active = numpy.random.normal(0.0, 1.0, (int(active.sum()), nfields))
        lo, hi = self.walker_offset, self.walker_offset + self.nwalkers
        xi = numpy.zeros((self.nwalkers, nfields))
        rows = numpy.cumsum(active) - 1
        mine = active[lo:hi]
        xi[mine] = xi_active[rows[lo:hi][mine]]
        return xi

    # ------------------------------------------------------------ re-ortho
    def orthogonalise(self, trial, free_projection):
        """handler.py:166-181; the free-projection branch (weight *= |detR|) is selected by the
        engine's PXB_FLAG_FREE_PROJECTION, set from the same propagator option."""
        self.engine.orthogonalise()
        self._phi_cache = None

    # --------------------------------------------------- population control
    def pop_control(self, comm, overlap_energy=False):
        """handler.py:225-254.  overlap_energy: the driver is going to evaluate the local energy
        right after this call (estimators/mixed.py:211-221); the comb then evaluates it for the
        walkers BEFORE the copies, concurrently with the (serial) comb plan -- a cloned walker's
        energy is its source's, and ELOC travels with the payload, so the estimator finds the
        same numbers."""
        if self.ntot_walkers == 1:
            return
        if self.use_log_shift:
            self.engine.update_log_shifts(comm)      # handler.py:228-229
        if self.pcont_method == "comb":
            self.comb(comm, overlap_energy)
        elif self.pcont_method == "pair_branch":
            self.pair_branch(comm)
        else:
            raise ValueError("Unknown population control method.")
        self._phi_cache = None

    def check_total_weight(self):
        """Deferred form of the reference's exit on a vanishing population (handler.py:236-241):
        the device paths only raise a flag; the driver polls it once per block."""
        self._check_total_weight()

    def _check_total_weight(self):
        if int(self.engine.counters[4].item()) != 0:
            # the reference prints and sys.exit()s (handler.py:236-241)
            raise RuntimeError("# Warning: total walker weight < 1e-8. Something is seriously wrong.")

    def comb(self, comm, overlap_energy=False):
        """handler.py:225-338.  One uniform draw from the global stream per call."""
        eng = self.engine
        r = numpy.random.random()
        multi = comm is not None and comm.size > 1
        if overlap_energy and self.overlap and (not multi or (self.peer_copy and eng.peers_attached)):
            main = torch.cuda.current_stream(eng.device)
            if self._side is None:
                self._side = torch.cuda.Stream(eng.device)
            side = self._side
            side.wait_stream(main)
            with torch.cuda.stream(side):
                gw = comm.allgather_tensor(torch.abs(eng.weight)) if multi else None
                eng.pop_plan(gw, r)             # serial sums + plan, needs only the weights
            # the plan is one CTA of serial float64 sums (bit-exact with the reference): with tens of
            # thousands of walkers it runs ~1 ms, so the persistent energy kernels leave it an SM
            reserve = 1 if self.ntot_walkers >= self.reserve_sm_walkers else 0
            if reserve:
                eng.reserve_sms(reserve)
            eng.local_energy()                  # X, exchange, ELOC of the walkers before the comb
            if reserve:
                eng.reserve_sms(0)
            main.wait_stream(side)
            if multi:
                # the all-gather above only orders the peers' SIDE streams: their ELOC / X are
                # final once their launch streams pass this barrier
                comm.stream_barrier(eng.device)
            eng.pop_pull()
            if multi:
                comm.stream_barrier(eng.device)   # nobody overwrites walker state while a peer pulls
            eng.pop_control_finish()
            return
        if not multi:
            eng.pop_control_comb(r)
            return
        gw = comm.allgather_tensor(torch.abs(eng.weight))
        if self.peer_copy and eng.peers_attached:
            # clones are pulled out of the peers' arenas over NVLink: no host round trip
            # (total weight < 1e-8 raises the sticky flag counters[4], see check_total_weight)
            eng.pop_control_comb_peers(gw, r)
            comm.stream_barrier(eng.device)
            eng.pop_control_finish()
            return
        eng.pop_rescale(gw)
        eng.comb_plan(gw, r)
        pairs = eng.pairs.cpu().numpy()
        self._check_total_weight()
        npairs = int(pairs[0])
        pl = pairs[1:1 + 2 * npairs].reshape(npairs, 2)
        self._move(comm, pl)
        eng.set_weights(1.0)

    def pair_branch(self, comm):
        """handler.py:340-412; the selection (stable sort + uniform draws) runs on
        the host over the gathered weights, walker data moves on the device."""
        eng = self.engine
        if comm is not None and comm.size > 1:
            gw_t = comm.allgather_tensor(torch.abs(eng.weight))
        else:
            gw_t = torch.abs(eng.weight)
        eng.pop_rescale(gw_t)
        total = eng.total_weight[0].item()
        self._check_total_weight()
        scale = total / self.target_weight
        gw = gw_t.cpu().numpy() / scale
        new_w, pairs = pair_branch_plan(gw, numpy.random.rand, self.min_weight, self.max_weight)
        lo, hi = self.walker_offset, self.walker_offset + self.nwalkers
        # cloned walkers take the averaged weight before their buffer is sent
        changed = [c for c, _ in pairs if lo <= c < hi]
        if changed:
            idx = torch.tensor([c - lo for c in changed], dtype=torch.long, device=eng.device)
            eng.weight[idx] = torch.as_tensor(new_w[changed]).to(eng.device)
        self._move(comm, pairs)

    def _move(self, comm, pairs):
        """Copy walker `clone` over walker `kill` for every (clone, kill) pair of
        GLOBAL indices (handler.py:301-334): device copy when both live here,
        packed send/recv between devices otherwise."""
        eng = self.engine
        local, out, inc = plan_moves(pairs, self.nw, self.rank)
        if local:
            src = torch.tensor([c for c, _ in local], dtype=torch.int32, device=eng.device)
            dst = torch.tensor([k for _, k in local], dtype=torch.int32, device=eng.device)
            eng.copy_walkers(src, dst)
        if comm is None or comm.size == 1:
            return
        pd = eng.payload_doubles()
        sends, recvs, unpack = [], [], []
        for peer in sorted(out):
            slots = torch.tensor(out[peer], dtype=torch.int32, device=eng.device)
            buf = torch.empty(len(out[peer]) * pd, dtype=torch.float64, device=eng.device)
            eng.pack_walkers(slots, buf)
            sends.append((peer, buf))
        for peer in sorted(inc):
            slots = torch.tensor(inc[peer], dtype=torch.int32, device=eng.device)
            buf = torch.empty(len(inc[peer]) * pd, dtype=torch.float64, device=eng.device)
            recvs.append((peer, buf))
            unpack.append((slots, buf))
        comm.exchange(sends, recvs)
        for slots, buf in unpack:
            eng.unpack_walkers(slots, buf)

    # ------------------------------------------------------------- restart
    def _restart_name(self, base):
        return '%s.rank%d.npy' % (base, self.rank)

    def get_write_buffers(self):
        """[nw, 3 + M*ne] complex: the reference's get_write_buffer (handler.py:432-435) for
        every walker of this rank."""
        eng = self.engine
        phi = eng.get_phi().cpu().numpy().reshape(self.nwalkers, -1)
        head = numpy.stack([eng.weight.cpu().numpy().astype(numpy.complex128),
                            eng.phase.cpu().numpy(), eng.ot.cpu().numpy()], axis=1)
        return numpy.concatenate([head, phi], axis=1)

    def _use_h5(self, base):
        from . import io
        return str(base).endswith(('.h5', '.hdf5')) and io.have_h5py()

    def write_walkers(self, comm=None):
        """handler.py:443-454.  HDF5 ('walker_%d' datasets, all ranks into one file) where h5py is
        installed and the name ends in .h5; otherwise one .npy per rank with the same records."""
        buffers = self.get_write_buffers()
        if self._use_h5(self.write_file):
            from . import io
            for r in range(1 if comm is None else comm.size):    # ranks take turns appending
                if r == self.rank:
                    io.write_walkers_h5(self.write_file, buffers, self.walker_offset, create=(r == 0))
                if comm is not None:
                    comm.barrier()
            return
        numpy.save(self._restart_name(self.write_file), buffers)

    def read_walkers(self, comm=None):
        """set_walker_from_buffer for every walker (handler.py:437-442, :477-485)."""
        eng = self.engine
        if self._use_h5(self.read_file):
            from . import io
            buff = io.read_walkers_h5(self.read_file, self.walker_offset, self.nwalkers)
        else:
            buff = numpy.load(self._restart_name(self.read_file))
        if buff.shape != (self.nwalkers, 3 + eng.M * eng.ne):
            raise ValueError("restart file %s does not match this walker population" % self.read_file)
        eng.set_phi(numpy.ascontiguousarray(buff[:, 3:].reshape(self.nwalkers, eng.M, eng.ne)))
        eng.weight.copy_(torch.as_tensor(numpy.ascontiguousarray(buff[:, 0].real)))
        eng.phase.copy_(torch.as_tensor(numpy.ascontiguousarray(buff[:, 1])))
        eng.ot.copy_(torch.as_tensor(numpy.ascontiguousarray(buff[:, 2])))
        self._phi_cache = None

    def set_total_weight(self, total_weight):
        self.engine.total_weight[0] = float(total_weight)


def plan_moves(pairs, nw, rank):
    """Split (clone, kill) pairs of GLOBAL walker indices by where the two ends
    live (rank = index // nw, slot = index % nw, handler.py:303-321).
    Returns (local [(src_slot, dst_slot)], out {peer: [src_slot]}, inc {peer:
    [dst_slot]}); per peer the two slot lists are in the same pair order on
    both sides, so packed buffers line up.  Vectorised: the pair list has a few
    thousand entries per step at 8 x 8192 walkers."""
    p = numpy.asarray(pairs, dtype=numpy.int64).reshape(-1, 2)
    local, out, inc = [], {}, {}
    if p.shape[0] == 0:
        return local, out, inc
    c, k = p[:, 0], p[:, 1]
    rc, rk = c // nw, k // nw
    lo = rank * nw
    here = (rc == rank) & (rk == rank)
    local = list(zip((c[here] - lo).tolist(), (k[here] - lo).tolist()))
    so = (rc == rank) & (rk != rank)
    for peer in numpy.unique(rk[so]).tolist():
        out[int(peer)] = (c[so & (rk == peer)] - lo).tolist()
    si = (rk == rank) & (rc != rank)
    for peer in numpy.unique(rc[si]).tolist():
        inc[int(peer)] = (k[si & (rc == peer)] - lo).tolist()
    return local, out, inc


def pair_branch_plan(abs_weights, rand, min_weight, max_weight):
    """Pair-branch selection of handler.py:340-386 on the gathered (rescaled)
    |weights|.  Returns (new_weights, [(clone, kill)]) with the reference's
    effective pairing: all messages carry the same tag on one rank, so the k-th
    cloned walker (ascending index) overwrites the k-th killed walker."""
    w = numpy.array(abs_weights, dtype=numpy.float64)
    order = numpy.argsort(w, kind='mergesort')
    ws = w[order].copy()
    s, e = 0, len(ws) - 1
    clones, kills = [], []
    while s < e:
        if ws[s] < min_weight or ws[e] > max_weight:
            wab = ws[s] + ws[e]
            r = rand()
            if r < ws[e] / wab:
                ws[e], ws[s] = 0.5 * wab, 0.0
                clones.append(int(order[e]))
                kills.append(int(order[s]))
            else:
                ws[s], ws[e] = 0.5 * wab, 0.0
                clones.append(int(order[s]))
                kills.append(int(order[e]))
            s += 1
            e -= 1
        else:
            break
    new_w = numpy.empty_like(ws)
    new_w[order] = ws
    return new_w, list(zip(sorted(clones), sorted(kills)))
