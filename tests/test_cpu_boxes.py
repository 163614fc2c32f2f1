"""The box-owning CPU workers of the bench's reference arm (oracle/cpu_boxes.py): threads that own one box each, exchange
ghost J through a mailbox and migrate leavers -- particles are conserved, every worker advances its particles in every
evaluation, and the ghost-exchanged J of the decomposed domain equals the J of the same particles in ONE box."""
import numpy as np

from picnic_b200 import decks, halo
from oracle import cpu_boxes
from oracle import oracle as orc

EPS = (0.0, 1.0e-3)


def _deck_fn(ncell):
    d = decks.deck_c3(ncell=ncell[0], ppc=2, dt=0.1, iter_max=21)
    d.species = decks.electron_proton((2, 2))
    d.ncell = tuple(ncell)
    return d


def _amps(deck):
    return 3.0e7, 5.0e8


def test_workers_conserve_particles_and_count_units():
    r = cpu_boxes.run(_deck_fn, _amps, 4, steps=3, warmup=1, n_outer=2, eps_outer=EPS, bn=10)
    assert r["workers"] == 4 and r["boxes"] == "2x2"
    assert r["particles"] == 4 * 10 * 10 * 4 * 2
    assert r["units"] == r["particles"] * 2 * 3          # migration keeps the total; n_outer x steps evaluations
    assert 1.0 < r["mean_picard_passes"] < 6.0


def test_decomposed_current_equals_single_box_current():
    bn, px, py = 8, 2, 2
    deck = _deck_fn((bn * px, bn * py))
    lay = halo.BoxLayout(2, deck.ncell, (bn, bn), deck.nghost, (1, 1))
    hub = cpu_boxes.ThreadHub(4)
    ws = [cpu_boxes.BoxWorker(deck, lay, r, hub, 3.0e7, 5.0e8, 1, EPS) for r in range(4)]
    # the same particles in one box that spans the domain
    lay1 = halo.BoxLayout(2, deck.ncell, deck.ncell, deck.nghost, (1, 1))
    one = cpu_boxes.BoxWorker(deck, lay1, 0, cpu_boxes.ThreadHub(1), 3.0e7, 5.0e8, 1, EPS)
    for k in range(2):
        be1 = one.species[k][1]
        for name in ("x", "xold", "v", "vold"):
            be1.p[name] = np.ascontiguousarray(np.concatenate([w.species[k][1].p[name] for w in ws], axis=1))
        be1.p["w"] = np.concatenate([w.species[k][1].p["w"] for w in ws])
        be1.p["id"] = np.concatenate([w.species[k][1].p["id"] for w in ws])
    import threading
    ts = [threading.Thread(target=w.step, args=(1,)) for w in ws]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    one.step(1)
    for comp, stag in enumerate(halo.STAG_J[2]):
        J1 = one.J[comp]
        assert np.abs(J1.a).max() > 0
        for w in ws:
            lo = w.lo
            hi = tuple(h + s for h, s in zip(w.hi, stag))
            got = w.grid._view(comp, lo, hi)
            ref = J1.a[tuple(slice(l - al, h - al + 1) for l, h, al in zip(lo, hi, J1.lo))]
            assert np.allclose(got, ref, rtol=1e-12, atol=1e-14 * np.abs(J1.a).max()), (comp, w.rank)
