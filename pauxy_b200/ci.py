"""Matrix elements between determinants of an orthogonal (particle-hole) expansion: what the
multi-determinant trial needs at set-up for its variational energy and for the mean-field shift
(pauxy/estimators/ci.py:187-300 get_hmatel / get_one_body_matel, used by
trial_wavefunction/multi_slater.py:153-176,235-259).  Host numpy, setup only.

A determinant is a sorted array of occupied SPIN orbitals: alpha orbital p is p, beta orbital p is
p + nbasis.
"""
import numpy


def excitation(bra, ket):
    """(holes, particles, sign): the spin orbitals of `ket` missing in `bra`, those of `bra` missing
    in `ket`, and the phase picked up when both lists are brought into maximum coincidence (each
    differing orbital is moved to the front of its list in turn)."""
    bra = numpy.asarray(bra)
    ket = numpy.asarray(ket)
    holes = sorted(set(ket.tolist()) - set(bra.tolist()))
    parts = sorted(set(bra.tolist()) - set(ket.tolist()))
    moves = 0
    for k, o in enumerate(holes):
        moves += int(numpy.nonzero(ket == o)[0][0]) - k
    for k, o in enumerate(parts):
        moves += int(numpy.nonzero(bra == o)[0][0]) - k
    return holes, parts, (-1.0 if moves % 2 else 1.0)


def _spatial(orb, nbasis):
    return (orb, 0) if orb < nbasis else (orb - nbasis, 1)


def one_body_element(ints, bra, ket):
    """<bra| sum_pq ints[p,q] a_p^+ a_q |ket> for a spin-independent one-body operator."""
    nb = ints.shape[-1]
    holes, parts, sign = excitation(bra, ket)
    if len(holes) == 0:
        return sum(ints[_spatial(o, nb)[0], _spatial(o, nb)[0]] for o in bra)
    if len(holes) == 1:
        (i, si), (a, sa) = _spatial(holes[0], nb), _spatial(parts[0], nb)
        return sign * ints[i, a] if si == sa else 0.0
    return 0.0


def hamiltonian_element(system, bra, ket):
    """(H, one-body part incl. the core energy, two-body part) between two determinants by the
    Slater-Condon rules; two-electron integrals <ij|kl> = system.hijkl(i, j, k, l)."""
    nb = system.nbasis
    holes, parts, sign = excitation(bra, ket)
    nex = len(holes)
    if nex == 0:
        occ = [_spatial(o, nb) for o in bra]
        e1 = system.ecore + sum(system.H1[0, p, p] for p, _ in occ)
        e2 = 0.0
        for x, (p, sp) in enumerate(occ):
            for q, sq in occ[x + 1:]:
                e2 += system.hijkl(p, q, p, q)
                if sp == sq:
                    e2 -= system.hijkl(p, q, q, p)
        return numpy.array([e1 + e2, e1, e2])
    if nex == 1:
        (i, si), (a, sa) = _spatial(holes[0], nb), _spatial(parts[0], nb)
        e1 = system.H1[0, i, a]
        e2 = 0.0
        for o in bra:
            q, sq = _spatial(o, nb)
            if (q, sq) != (i, si):
                e2 += system.hijkl(i, q, a, q)
                if sq == si:
                    e2 -= system.hijkl(i, q, q, a)
        return sign * numpy.array([e1 + e2, e1, e2])
    if nex == 2:
        (i, si), (j, sj) = _spatial(holes[0], nb), _spatial(holes[1], nb)
        (a, sa), (b, sb) = _spatial(parts[0], nb), _spatial(parts[1], nb)
        v = 0.0
        if si == sa:
            v += system.hijkl(i, j, a, b)
        if si == sb:
            v -= system.hijkl(i, j, b, a)
        return sign * numpy.array([v, 0.0, v])
    return numpy.zeros(3)
