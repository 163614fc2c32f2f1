#!/usr/bin/env python
"""bench.py -- particle-advances/s of the implicit push + deposit hot path.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE
JSON line (rank 0).  Under torchrun (N > 1) every rank owns one box of a periodic domain
tiled N boxes wide (weak scaling: the per-GPU box is fixed), with the ghost-J add-exchange
between neighbouring boxes over NCCL.

Workload (config.workload): BASELINE.json configs[2] "C3" -- 2D implicit energy-conserving
PIC, 512x512 cells, 2 species x 100 particles per cell (5.24e7 particles), CC1 gather and
deposit, rtol_particles = 1e-12, iter_max_particles = 21 -- the configuration the
north-star target (push+deposit vs HBM roofline on one B200) is quoted on.

One STEP = the particle side of one implicit time step under a Picard outer loop with
`n_outer` nonlinear function evaluations (PICTimeIntegrator_EM_ThetaImplicit.cpp:193-364):
    updateOldParticlePositions/Velocities
    n_outer x preRHSOp:  [E,B of that iteration] -> per species advanceParticlesIteratively
                         + setCurrentDensity (fused), sum over species, ghost add-exchange
    advanceVelocities_2ndHalf, advancePositions_2ndHalf, applyBCs (periodic)
    binTheParticles (cell sort) every `sort_every` steps (default 4; PICNIC itself bins only for
    collisions -- the sort is this engine's own locality device, any particle order is correct)
Units per step = particles x n_outer particle-advances.  Fields are smooth
analytic modes; outer iteration j sees the field scaled by (1 + eps_j), eps = 0, 1e-3, 1e-6,
which mimics the shrinking field updates of a converging Picard loop so that the particle
Picard loop does real work in every evaluation (mean passes k is reported).

`value`  : device-resident: the three field sets are already in HBM (field slots).
`e2e`    : the same step driven through the reference-facing C ABI with HOST buffers: every
           preRHSOp uploads E,B (6 components) from pinned host memory and reads the summed
           J (3 components) back to pinned host memory, inside the timed region.
`--impl reference`: the CPU oracle (the restatement of the reference algorithm; the
           reference itself needs Chombo/gfortran/MPI and cannot be built) on all
           host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from picnic_b200 import decks  # noqa: E402

BYTES_PER_ADVANCE_2D = 88.0   # (2D+7)*8 B, compulsory SoA traffic of one advance
EPS_OUTER = (0.0, 1.0e-3, 1.0e-6, 1.0e-9)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ncell", type=int, default=512, help="cells per direction of the per-GPU box")
    ap.add_argument("--ppc", type=int, default=10, help="particles per cell per direction per species")
    ap.add_argument("--n-outer", type=int, default=3)
    ap.add_argument("--sort-every", type=int, default=8,
                    help="locality-sort period in steps (the reference never sorts a collisionless deck; the sort only "
                         "keeps the fused kernel on its fast path: measured 8.31 / 8.07 / 8.16 ms per step at 4 / 8 / 16, "
                         "round 2, C3)")
    ap.add_argument("--workload", default="auto", choices=["auto", "c3", "c5"],
                    help="c3: BASELINE configs[2], one 512^2 box x 2 species x 100 ppc per GPU (the deck the single-GPU "
                         "roofline target is quoted on); c5: the per-GPU shard of BASELINE configs[4] (2048^2 cells x 256 ppc "
                         "over 8 GPUs = sixteen 512^2 boxes, two per GPU, 2 species x 128 ppc); auto = c3 at 1 GPU, c5 above")
    ap.add_argument("--boxes-per-gpu", type=int, default=0, help="0 = from the workload (c3: 1, c5: 2)")
    ap.add_argument("--sort", default="dual", choices=["dual", "cell"],
                    help="locality sort of the collisionless step: by dual cell (pgpu_sort_for_locality, default) or by "
                         "primal cell + quadrant (pgpu_bin_particles, what the collision step needs)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="ghost-J add-exchange between boxes: the library's peer-memory kernels (default) or "
                         "pack + NCCL send/recv + unpack-add")
    ap.add_argument("--migration", default="peer", choices=["peer", "nccl"],
                    help="particle migration between boxes: peer-memory inboxes with device-side counts (default) or "
                         "count all-gather + NCCL send/recv")
    ap.add_argument("--dt", type=float, default=0.1)
    ap.add_argument("--iter-max", type=int, default=21)
    ap.add_argument("--field-scale", type=float, default=1.0,
                    help="multiplies the amplitudes of the synthetic E, B (1 = a thermal electron's velocity changes by "
                         "~5 %% per step from E and turns by ~0.05 rad from B: mean Picard passes k = 2.2; 8 gives k >= 4)")
    ap.add_argument("--field-kmul", type=int, default=1,
                    help="multiplies the wave numbers of the synthetic fields (1: wavelengths of 512, 256 and 171 cells; "
                         "32: 16, 8 and 5.3 cells)")
    ap.add_argument("--no-variants", action="store_true",
                    help="skip the secondary C3 variants (iter_max 0, harder fields, explicit leap-frog)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="particles per species in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-collisions", action="store_true", help="skip the secondary C2 collision-pairs/s leg")
    ap.add_argument("--no-c4", action="store_true", help="skip the secondary C4 leg (1D, 1e8 particles)")
    ap.add_argument("--no-mass-matrix", action="store_true", help="skip the secondary mass-matrix leg")
    ap.add_argument("--no-c5-shard", action="store_true", help="skip the one-GPU run of the C5 per-GPU shard")
    a = ap.parse_args()
    if a.workload == "auto":
        a.workload = "c3" if a.gpus == 1 else "c5"
    if a.boxes_per_gpu == 0:
        a.boxes_per_gpu = 2 if a.workload == "c5" else 1
    # particles per cell per species as a lattice (ppc0, ppc1): c3 10 x 10 (or --ppc squared), c5 16 x 8 = 128
    a.ppc2 = (16, 8) if a.workload == "c5" else (a.ppc, a.ppc)
    return a


# ------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------
def box_layout(world):
    """Tile `world` equal square boxes (System.cpp:181-185) as px x py."""
    px = 1
    while px * px < world:
        px *= 2
    py = world // px
    assert px * py == world, "world size must be a power of two"
    return px, py


def make_deck(args, world):
    """The periodic domain of world * boxes_per_gpu square boxes of ncell^2 cells."""
    px, py = box_layout(world * args.boxes_per_gpu)
    d = decks.deck_c3(ncell=args.ncell, ppc=args.ppc, dt=args.dt, iter_max=args.iter_max)
    d.species = decks.electron_proton(tuple(args.ppc2))
    d.ncell = (args.ncell * px, args.ncell * py)
    return d, (px, py)


def rank_box(args, rank, layout):
    px, _ = layout
    bi, bj = rank % px, rank // px
    lo = (bi * args.ncell, bj * args.ncell)
    hi = (lo[0] + args.ncell - 1, lo[1] + args.ncell - 1)
    return lo, hi


def field_amplitudes(deck, scale=1.0):
    """E0, B0 such that omega_pe*dt ~ omega_ce*dt ~ 0.1-like kicks: a thermal electron's
    velocity changes by ~10 % per step from E and rotates by ~0.1 rad from B."""
    sp = deck.species[0]
    fn = abs(sp.fnorm_const(deck.units))
    vth = np.sqrt(decks.QE / decks.ME * sp.temperature_eV[0] / sp.mass) / decks.CVAC
    alpha = fn * deck.cnorm_dt / 2.0
    return scale * 0.05 * vth / alpha, scale * 0.05 / alpha


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        """Only the samples that arrived inside [t0, t1] (host clock around the timed region) count."""
        self.t0, self.t1 = t0, t1

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t0", None), getattr(self, "t1", None)
        for (ts, r) in self.rows:
            if t0 is not None and not (t0 <= ts <= t1 + 0.02):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# the GPU arm
# ------------------------------------------------------------------------------------------
class Box:
    """One Chombo box of this process: its grid, its species and its exchanges."""
    pass


class Engine:
    """The particle side of the implicit step, driven through the C ABI (capi).  A process owns `boxes_per_rank`
    boxes of the square equal-box decomposition (System.cpp:169-245); box b = bi + bj * px lives in process
    b // boxes_per_rank."""

    def __init__(self, args, rank, world, device, stream=None):
        import torch
        from picnic_b200 import capi
        self.torch, self.capi, self.args, self.rank, self.world = torch, capi, args, rank, world
        self.deck, self.layout = make_deck(args, world)
        deck = self.deck
        B = args.boxes_per_gpu
        self.nboxes = world * B
        E0, B0 = field_amplitudes(deck, args.field_scale)
        self.n_outer = args.n_outer
        self.boxes = []
        self.n_particles = 0
        self.h2d = self.d2h = 0
        for k in range(B):
            bx = Box()
            bx.id = rank * B + k
            bx.lo, bx.hi = rank_box(args, bx.id, self.layout)
            bx.grid = capi.Grid(2, deck.ncell, deck.xmin, deck.dx, deck.nghost, (1, 1), box_lo=bx.lo,
                                box_hi=bx.hi, volume_scale=deck.volume_scale)
            E, Bf = decks.analytic_fields(deck, bx.lo, bx.hi, E0=E0, B0=B0, kmul=args.field_kmul)
            # pinned host copies of the field sets of each outer iteration + pinned J read-back: ONE contiguous
            # buffer per direction of transfer (components back to back in the C ABI's packed order)
            nfield = bx.grid.fields_packed_size()
            nJ = bx.grid.current_packed_size()
            bx.host_fields = []
            for j in range(self.n_outer):
                t = torch.empty(nfield, dtype=torch.float64).pin_memory()
                h = t.numpy()
                off = 0
                for (lo, hi, a) in list(E) + list(Bf):
                    h[off:off + a.size] = (a * (1.0 + EPS_OUTER[j % len(EPS_OUTER)])).ravel(order="F")
                    off += a.size
                assert off == nfield
                bx.host_fields.append((h, t))
            tJ = torch.empty(nJ, dtype=torch.float64).pin_memory()
            bx.host_J = (tJ.numpy(), tJ)
            for j in range(self.n_outer):           # resident field slots for the device-timed arm
                bx.grid.fields_select(j)
                self._upload_fields(bx, j)
            bx.grid.fields_select(0)
            capi.check(capi.load().pgpu_synchronize())
            rng = np.random.default_rng(deck.seed + 1000 * bx.id)
            bx.species = []
            for sdef in deck.species:
                p = decks.load_species(deck, sdef, bx.lo, bx.hi, rng)
                sp = capi.Species(bx.grid, sdef.mass, sdef.charge, sdef.fnorm_const(deck.units), deck.units.cvac_norm,
                                  interp_N=deck.interp_N, interp_J=deck.interp_J, interp_E=deck.interp_E,
                                  rtol=deck.rtol, iter_max=deck.iter_max)
                sp.upload(p["x"], p["v"], p["w"], ids=p["id"])
                self._sort(sp)
                bx.species.append(sp)
                self.n_particles += sp.n
                del p
            self.h2d += nfield * 8 * self.n_outer
            self.d2h += nJ * 8 * self.n_outer
            bx.halo, bx.migration = None, []
            self.boxes.append(bx)
        self.species = [sp for bx in self.boxes for sp in bx.species]
        self.grid = self.boxes[0].grid
        self.step_no = 0
        self.halo_bytes = 0
        if self.nboxes > 1:
            # ghost add-exchange of J after every deposit and particle migration once per step, between the boxes of
            # this process (plain device pointers) and of the other processes (CUDA IPC), picnic_b200/halo.py
            from picnic_b200 import halo
            dev = torch.device("cuda", device)
            lay = halo.BoxLayout(2, deck.ncell, (args.ncell, args.ncell), deck.nghost, (1, 1))
            assert lay.world == self.nboxes
            for bx in self.boxes:
                assert lay.box(bx.id) == (tuple(bx.lo), tuple(bx.hi))
            comm = halo.DistComm(rank, world, stream=stream) if world > 1 else None
            if args.halo == "peer":
                for bx in self.boxes:
                    bx.halo = halo.PeerHaloExchange(lay, bx.id, bx.grid)
                hs = [bx.halo for bx in self.boxes]
                if world == 1:
                    halo.PeerHaloExchange.connect_local(hs)
                else:
                    halo.PeerHaloExchange.connect_mixed(hs, comm, B)
            else:
                assert B == 1, "--halo nccl is the one-box-per-process route"
                self.boxes[0].halo = halo.HaloExchange(lay, rank, comm,
                                                       halo.CapiGridBackend(self.grid, dev, on_torch_stream=True))
            self.halo_bytes = sum(bx.halo.bytes_per_exchange for bx in self.boxes)
            if args.migration == "peer":
                # leavers go straight into the neighbours' inboxes (peer stores), counts stay on the device
                for bx in self.boxes:
                    bx.migration = [halo.PeerMigration(lay, bx.id, sp, capacity=max(16384, sp.n // 512))
                                    for sp in bx.species]
                for k in range(len(deck.species)):
                    ms = [bx.migration[k] for bx in self.boxes]
                    if world == 1:
                        halo.PeerMigration.connect_local(ms)
                    else:
                        halo.PeerMigration.connect_mixed(ms, comm, B)
            else:
                assert B == 1, "--migration nccl is the one-box-per-process route"
                self.boxes[0].migration = [halo.Migration(lay, rank, comm,
                                                          halo.CapiSpeciesBackend(sp, dev, on_torch_stream=True))
                                           for sp in self.boxes[0].species]
        self.migrated = 0
        self.sections, self.stream = None, stream

    def _sort(self, sp):
        if self.args.sort == "dual":
            sp.sort_for_locality()
        else:
            sp.bin_particles()

    def _upload_fields(self, bx, j):
        """One H2D copy of the six field components from pinned memory (pgpu_fields_set_packed)."""
        capi = self.capi
        capi.check(capi.load().pgpu_fields_set_packed(bx.grid.h, bx.host_fields[j][0].ctypes.data))

    # device-side section timers: pairs of events on the engine stream, resolved after the region
    def _mark(self, name):
        if self.sections is None:
            return
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.stream)
        self.sections.append((name, ev, time.perf_counter()))

    def sections_begin(self):
        self.sections = []
        self._mark("start")

    def sections_end(self):
        """{section: total ms} of everything recorded since sections_begin (call after a synchronize)."""
        out, host, marks = {}, {}, self.sections
        self.sections = None
        for (_, e0, t0), (name, e1, t1) in zip(marks[:-1], marks[1:]):
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
            host[name] = host.get(name, 0.0) + (t1 - t0) * 1e3     # time the host spent issuing the section
        return out, host

    def pre_rhs_op(self, j, host_io):
        """PicSpeciesInterface::preRHSOp (PicSpeciesInterface.cpp:899-994) for outer iteration j, all boxes."""
        capi, lib = self.capi, self.capi.load()
        for bx in self.boxes:
            if host_io:
                self._upload_fields(bx, j)          # async H2D from pinned memory on the engine stream
            else:
                bx.grid.fields_select(j)
            bx.grid.current_zero()
        self._mark("fields_in")
        for bx in self.boxes:
            for sp in bx.species:
                capi.check(lib.pgpu_advance_particles_iteratively(sp.h, self.deck.dt, 1, None))
                bx.grid.current_add(sp)
        self._mark("advance_deposit")
        if self.nboxes > 1:
            if self.args.halo == "peer":
                # every send of a phase is enqueued before any receive of that phase: a receive spins on its
                # neighbours' arrival flags, and the boxes of one process share one stream
                for bx in self.boxes:
                    bx.halo.begin()
                for ph in range(self.boxes[0].halo.nphase):
                    for bx in self.boxes:
                        bx.halo.send(ph)
                    for bx in self.boxes:
                        bx.halo.recv_add(ph)
            else:
                self.boxes[0].halo.add_exchange()
            self._mark("ghost_J_exchange")
        for bx in self.boxes:
            bx.grid.current_finalize()
            if host_io:                             # ONE D2H copy of the three J components into pinned memory
                capi.check(lib.pgpu_current_get_packed_async(bx.grid.h, bx.host_J[0].ctypes.data))
        if host_io:
            capi.check(lib.pgpu_synchronize())      # the host reads J here (the field solve's residual)
        self._mark("J_out")

    def step(self, host_io=False):
        for sp in self.species:
            sp.update_old_positions()
            sp.update_old_velocities()
        if host_io:
            for bx in self.boxes:
                bx.grid.fields_select(0)
        for j in range(self.n_outer):
            self.pre_rhs_op(j, host_io)
        for sp in self.species:
            sp.finish_implicit_step((1, 1), (1, 1))   # 2nd-half v, 2nd-half x, periodic applyBCs
        self._mark("finish_step")
        if self.nboxes > 1:                            # remapOutcast: leavers to the owning box
            from picnic_b200 import halo
            ms = [m for bx in self.boxes for m in bx.migration]
            if self.args.migration == "peer":
                self.migrated += halo.migrate_all_peer(ms)
            else:
                self.migrated += halo.migrate_all(ms)
            self._mark("migration")
        self.step_no += 1
        if self.args.sort_every > 0 and self.step_no % self.args.sort_every == 0:
            for sp in self.species:
                self._sort(sp)
            self._mark("cell_sort")

    def sync(self):
        self.capi.check(self.capi.load().pgpu_synchronize())

    def destroy(self):
        for bx in self.boxes:
            for m in bx.migration:
                if hasattr(m, "destroy"):
                    m.destroy()
            if bx.halo is not None and hasattr(bx.halo, "destroy"):
                bx.halo.destroy()
            for sp in bx.species:
                sp.destroy()
            bx.grid.destroy()


def collisions_leg(args, torch, capi, stream, peak):
    """Secondary line item: collision-pairs/s on BASELINE.json configs[1] "C2" -- 2D 256x256 cells,
    64 particles per cell per species, electrons (150 eV) + protons (50 eV), Takizuka-Abe binary
    Coulomb collisions (Clog = 3): e-e, i-i and e-i every step, preceded by prepForScatter (cell
    sort + cell moments of both species, PicSpeciesInterface.cpp:1593-1625).  pairs/step =
    (32 + 32 + 64) per cell.  Device-resident; the C ABI has no per-step host traffic here."""
    deck = decks.deck_c2()
    lo, hi = (0, 0), (deck.ncell[0] - 1, deck.ncell[1] - 1)
    grid = capi.Grid(2, deck.ncell, deck.xmin, deck.dx, deck.nghost, (1, 1), volume_scale=deck.volume_scale)
    rng = np.random.default_rng(deck.seed)
    sps = []
    for sdef in deck.species:
        p = decks.load_species(deck, sdef, lo, hi, rng)
        sp = capi.Species(grid, sdef.mass, sdef.charge, sdef.fnorm_const(deck.units), deck.units.cvac_norm,
                          interp_N=deck.interp_N, interp_J=deck.interp_J, interp_E=deck.interp_E)
        sp.upload(p["x"], p["v"], p["w"], ids=p["id"])
        sps.append(sp)
    dt_sec = deck.dt * deck.units.time
    state = {"step": 0}

    def one_step(prep):
        if prep:
            for sp in sps:
                # a time step begins with updateOldParticlePositions/Velocities (flags only on the device)
                sp.update_old_positions()
                sp.update_old_velocities()
                sp.bin_particles()
                sp.set_moments()
        n = 0
        for (a, b) in ((0, 0), (1, 1), (0, 1)):
            capi.check(capi.load().pgpu_collide_ta(sps[a].h, sps[b].h, 3.0, dt_sec, 1983, state["step"], None))
        state["step"] += 1

    def timed(nsteps, prep):
        capi.check(capi.load().pgpu_synchronize())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(nsteps):
            one_step(prep)
        e1.record(stream)
        capi.check(capi.load().pgpu_synchronize())
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    one_step(True)
    # pair count of one step (counted once, outside the timed region: the count read-back synchronises)
    pairs = 0
    for (a, b) in ((0, 0), (1, 1), (0, 1)):
        pairs += capi.collide_ta(sps[a], sps[b], 3.0, dt_sec, 1983, 10 ** 6)
    nsteps = max(args.steps, 5)
    timed(max(args.warmup, 3), True)
    capi.profile_reset()
    capi.profile_enable(True)
    ms_full = timed(nsteps, True)
    capi.profile_enable(False)
    k_self, n_self = capi.profile_query("collide_ta_self")
    k_inter, n_inter = capi.profile_query("collide_ta_inter")
    prep = {k: round(capi.profile_query(k)[0] / nsteps, 4)
            for k in ("bin_key", "bin_sort", "bin_permute", "bin_starts", "cell_moments")}
    ms_k = timed(nsteps, False)
    n_total = sum(sp.n for sp in sps)
    for sp in sps:
        sp.destroy()
    grid.destroy()
    kern_ms = (k_self + k_inter) / nsteps
    achieved = 96.0 * pairs / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None
    return {"metric": "collision-pairs/s (Takizuka-Abe, intra-cell)", "unit": "collision-pairs/s",
            "value": pairs * nsteps / (ms_full * 1e-3),
            "value_kernels_only": pairs * nsteps / (ms_k * 1e-3),
            "workload": "C2: 2D %dx%d cells, 64 ppc/species, e-e + i-i + e-i TA (Clog 3) per step; `value` includes "
                        "prepForScatter (cell sort + cell moments of both species) every step"
                        % (deck.ncell[0], deck.ncell[1]),
            "particles": n_total, "pairs_per_step": pairs, "steps": nsteps,
            "ms_per_step": ms_full / nsteps, "ms_per_step_kernels_only": ms_k / nsteps,
            "prep_kernel_ms_per_step": prep,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if achieved else None, "bytes_per_unit": 96.0,
                         "kernel": "collide_ta_self + collide_ta_inter (per step 3 staged launches, cells <= 128 "
                                   "particles/species in shared memory, + 3 launches for the larger cells)",
                         "kernel_ms_per_step": kern_ms,
                         "note": "instruction-issue bound, not HBM bound: per pair two Philox4x32-10 blocks, a share of "
                                 "the warp's bitonic sort of the shuffle keys and fp64 sqrt/div/sincospi; ncu "
                                 "(profiles/r02_f_ncu_k_ta.txt) has issue-active 57-60 % at 12-15 % fp64 pipe and "
                                 "<= 30 % DRAM throughput"}}


def c4_leg(args, torch, capi, stream, peak):
    """Secondary line item: BASELINE.json configs[3] "C4" as a periodic stand-in -- 1D, 250 000 cells x 200 ppc x 2
    species = 1e8 particles, CC1 gather/deposit, implicit push + deposit of one nonlinear evaluation per species
    (1D fused kernel pgpu_advance_cc1_1d.cu), and the weighted Coulomb collisions (NANBU, Clog 10: e-e, i-i, e-i).
    Device-resident; algorithmic bytes per advance in 1D = (2*1+7)*8 = 72."""
    deck = decks.deck_c4()
    deck.dt = 0.1
    lo, hi = (0,), (deck.ncell[0] - 1,)
    E, B = decks.analytic_fields(deck, lo, hi, E0=3.0e7, B0=5.0e8)
    grid = capi.Grid(1, deck.ncell, deck.xmin, deck.dx, deck.nghost, (1,), volume_scale=deck.volume_scale)
    grid.set_fields(E, B)
    rng = np.random.default_rng(3)
    lib = capi.load()
    sps = []
    for sdef in deck.species:
        p = decks.load_species(deck, sdef, lo, hi, rng)
        sp = capi.Species(grid, sdef.mass, sdef.charge, sdef.fnorm_const(deck.units), deck.units.cvac_norm,
                          interp_N=deck.interp_N, interp_J=deck.interp_J, interp_E=deck.interp_E, rtol=deck.rtol,
                          iter_max=deck.iter_max)
        sp.upload(p["x"], p["v"], p["w"], ids=p["id"])
        sp.bin_particles()
        sps.append(sp)
        del p
    n = sum(sp.n for sp in sps)

    def timed(fn, reps):
        fn()
        capi.check(lib.pgpu_synchronize())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        capi.check(lib.pgpu_synchronize())
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def adv():
        for sp in sps:
            capi.check(lib.pgpu_advance_particles_iteratively(sp.h, deck.dt, 1, None))
    reps = max(args.steps // 4, 5)
    capi.picard_totals(reset=True)
    capi.profile_reset()
    capi.profile_enable(True)
    ms_adv = timed(adv, reps)
    capi.profile_enable(False)
    k_ms, k_n = capi.profile_query("advance_cc1_1d_fused")
    padv, papp, _ = capi.picard_totals(reset=True)
    for sp in sps:
        sp.bin_particles()
        sp.set_moments()
    grid.debye_length(sps)
    dt_sec = deck.dt * deck.units.time
    state = {"k": 0}

    def col():
        for (a, b) in ((0, 0), (1, 1), (0, 1)):
            capi.check(lib.pgpu_collide_coulomb(sps[a].h, sps[b].h, capi.C.byref(capi.CoulombParams(10.0, 1, 0, 11, 1)),
                                                dt_sec, 1983, state["k"], None))
        state["k"] += 1
    capi.profile_reset()
    capi.profile_enable(True)
    ms_col = timed(col, 3)
    capi.profile_enable(False)
    kc_ms = (capi.profile_query("collide_coulomb_intra")[0] + capi.profile_query("collide_coulomb_inter")[0]) / 4.0
    pairs = sum(capi.collide_coulomb(sps[a], sps[b], 10.0, dt_sec, 1983, 99, angular=1)
                for (a, b) in ((0, 0), (1, 1), (0, 1)))
    # setMassMatrices on the same deck (1D run kernel, pgpu_massmatrix.cu); the orbits are the ones adv() left
    for sp in sps:
        sp.bin_particles()
    adv()
    grid.mass_matrices_init(3)

    def set_mm():
        grid.mass_matrices_zero()
        for sp in sps:
            sp.accumulate_mass_matrices(deck.dt)
        grid.mass_matrices_save_E0()
    ms_mm = timed(set_mm, 3)
    capi.picard_totals(reset=True)
    for sp in sps:
        sp.destroy()
    grid.destroy()
    kern = k_ms / max(k_n, 1)
    achieved = 72.0 * (n / len(sps)) / (kern * 1e-3) / 1e9 if k_n else None
    return {"metric": "particle-advances/s (implicit push + deposit, 1D)", "unit": "particle-advances/s",
            "value": n / (ms_adv * 1e-3), "ms_per_evaluation": ms_adv, "particles": n,
            "workload": "C4 stand-in: 1D periodic, %d cells x 200 ppc x 2 species, CC1, one nonlinear evaluation "
                        "(fused advance+deposit of both species)" % deck.ncell[0],
            "mean_picard_passes": round(papp / max(padv, 1), 3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if achieved else None, "bytes_per_unit": 72.0,
                         "kernel": "advance_cc1_1d_fused", "kernel_ms_per_launch": kern,
                         "units_per_launch": n / len(sps)},
            "coulomb": {"metric": "collision-pairs/s (weighted Coulomb, NANBU)", "value": pairs / (ms_col * 1e-3),
                        "pairs_per_step": pairs, "ms_per_step": ms_col,
                        # a weighted pair reads and writes 3 velocity doubles of both partners and reads
                        # their weights: 112 B
                        "roofline": {"bound": "hbm", "achieved": 112.0 * pairs / (kc_ms * 1e-3) / 1e9 if kc_ms else None,
                                     "peak": peak, "unit": "GB/s",
                                     "frac": 112.0 * pairs / (kc_ms * 1e-3) / 1e9 / peak if kc_ms else None,
                                     "bytes_per_unit": 112.0,
                                     "kernel": "collide_coulomb_intra x2 + collide_coulomb_inter (3 launches per step)",
                                     "kernel_ms_per_step": kc_ms,
                                     "note": "instruction-issue bound: ~1700 warp instructions per pair slot (fp64 sqrt x4, "
                                             "div x5, Nanbu's exp/log inversion, sincos, two Philox blocks); ncu "
                                             "(profiles/r02_f_ncu_k_coulomb.txt): issue-active 49-52 %, fp64 pipe 24-29 %, "
                                             "DRAM 16-20 %, 128 registers"}},
            "mass_matrices": {"metric": "particles/s through setMassMatrices (1D run kernel)", "value": n / (ms_mm * 1e-3),
                              "ms_per_setMassMatrices": ms_mm,
                              "hbm_frac_at_72B": 72.0 * n / (ms_mm * 1e-3) / 1e9 / peak}}


def mass_matrix_cpu_sample(deck, ncell=96):
    """CPU leg of the mass-matrix item: the oracle's cc1_2d_deposit_mass_matrix restatement, one thread, on a bounded
    sample of the same deck (ncell^2 cells x 100 ppc of the electron species, cell-ordered as the loader leaves them,
    xbar half a step from xold)."""
    from oracle import oracle as orc
    import copy
    d = copy.copy(deck)
    d.ncell = (ncell, ncell)
    lo, hi = (0, 0), (ncell - 1, ncell - 1)
    _, B = decks.analytic_fields(d, lo, hi, E0=3.0e7, B0=5.0e8)
    sdef = d.species[0]
    p = decks.load_species(d, sdef, lo, hi, np.random.default_rng(3))
    cn = d.dt * d.units.cvac_norm
    xold = np.ascontiguousarray(p["x"])
    x = np.ascontiguousarray(xold + 0.5 * cn * p["v"][:2])
    v = np.ascontiguousarray(p["v"])
    geom = orc.make_geom(2, d.xmin, d.xmax, d.dx, d.nghost)
    Bf = [orc.Fab(l, h, a) for (l, h, a) in B]
    nc, sigma = orc.mm_alloc(2, orc.CC1, d.nghost, lo, hi)
    J0 = [orc.fab_for(lo, hi, d.nghost, st) for st in orc.E_STAG[2]]
    t0 = time.perf_counter()
    rc = orc.deposit_mass_matrices(geom, orc.CC1, x, xold, v, v, p["w"], sdef.charge / d.volume_scale,
                                   sdef.fnorm_const(d.units) * cn / 2.0, cn, Bf, J0, sigma)
    sec = time.perf_counter() - t0
    n = int(p["w"].size)
    return {"value": n / sec, "unit": "particles/s", "cores": 1, "kind": "port", "rc": int(rc), "seconds": sec,
            "sample": "%d x %d cells x 100 ppc = %d particles of one species, oracle (g++ -O2 -ffp-contract=off), "
                      "one thread" % (ncell, ncell, n)}


def mass_matrix_leg(args, torch, capi, stream, peak):
    """Secondary line item: PicSpeciesInterface::setMassMatrices + computeJfromMassMatrices on the C3 box
    (2D 512x512, 2 species x 100 ppc, CC1, 3 ghost layers -> 231 sigma components = 0.5 GB).  Unit: one particle
    through accumulateMassMatrices = 272 products f*weight_J*weight_E summed into J0 and the nine sigmas.  Algorithmic
    bytes per particle: read xbar[2], xold[2], ubar[3], uold[3], w = 88 B (the sigma arrays stay L2 resident per
    sweep row and cost 1 GB of DRAM traffic per 2.6e7 particles)."""
    deck = decks.deck_c3(ncell=args.ncell, ppc=args.ppc, dt=args.dt, iter_max=args.iter_max)
    lo, hi = (0, 0), (deck.ncell[0] - 1, deck.ncell[1] - 1)
    E, B = decks.analytic_fields(deck, lo, hi, E0=3.0e7, B0=5.0e8)
    grid = capi.Grid(2, deck.ncell, deck.xmin, deck.dx, deck.nghost, (1, 1), volume_scale=deck.volume_scale)
    grid.set_fields(E, B)
    rng = np.random.default_rng(3)
    lib = capi.load()
    sps = []
    for sdef in deck.species:
        p = decks.load_species(deck, sdef, lo, hi, rng)
        sp = capi.Species(grid, sdef.mass, sdef.charge, sdef.fnorm_const(deck.units), deck.units.cvac_norm,
                          interp_N=deck.interp_N, interp_J=deck.interp_J, interp_E=deck.interp_E, rtol=deck.rtol,
                          iter_max=deck.iter_max)
        sp.upload(p["x"], p["v"], p["w"], ids=p["id"])
        sp.bin_particles()
        sps.append(sp)
        del p
    n = sum(sp.n for sp in sps)
    for sp in sps:      # converged orbits (xbar, ubar) of one implicit evaluation
        capi.check(lib.pgpu_advance_particles_iteratively(sp.h, deck.dt, 1, None))
    nc = grid.mass_matrices_init(3)
    ncomp = int(sum(int(a) * int(b) for a, b in nc))

    def timed(fn, reps):
        fn()
        capi.check(lib.pgpu_synchronize())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        capi.check(lib.pgpu_synchronize())
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def set_mm():
        grid.mass_matrices_zero()
        for sp in sps:
            sp.accumulate_mass_matrices(deck.dt)
        grid.mass_matrices_save_E0()

    reps = max(args.steps // 8, 3)
    set_mm()                      # first launches (module load, local-memory reservation) stay out of the kernel timers
    capi.check(lib.pgpu_synchronize())
    capi.profile_reset()
    capi.profile_enable(True)
    ms_set = timed(set_mm, reps)
    capi.profile_enable(False)
    k_ms, k_n = capi.profile_query("mass_matrix_run")
    d_ms, d_n = capi.profile_query("mass_matrix_deferred")
    ms_J = timed(grid.compute_J_from_mass_matrices, reps)
    capi.picard_totals(reset=True)      # raises on a crossing / bounds error of the deposit kernels
    for sp in sps:
        sp.destroy()
    grid.destroy()
    kern = k_ms / max(k_n, 1)
    achieved = 88.0 * (n / len(sps)) / (kern * 1e-3) / 1e9 if k_n else None
    cpu = None
    if not args.no_cpu_baseline:
        cpu = mass_matrix_cpu_sample(deck)
    return {"cpu_baseline": cpu,
            "metric": "particles/s through setMassMatrices (accumulateMassMatrices of both species)",
            "unit": "particles/s", "value": n / (ms_set * 1e-3), "ms_per_setMassMatrices": ms_set,
            "ms_per_computeJfromMassMatrices": ms_J, "particles": n, "sigma_components": ncomp,
            "workload": "C3 box: 2D %dx%d cells, 2 species x %d ppc, CC1, %d ghost layers; zero + accumulate both "
                        "species + save E0, then J = J0 + sigma (E - E0)" % (deck.ncell[0], deck.ncell[1],
                                                                              args.ppc * args.ppc, deck.nghost),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if achieved else None, "bytes_per_unit": 88.0,
                         "kernel": "mass_matrix_run (k_mm_cc1_2d_run)", "kernel_ms_per_launch": kern,
                         "deferred_kernel_ms_per_launch": d_ms / max(d_n, 1), "units_per_launch": n / len(sps),
                         "note": "bound by the fp64 reductions into L2 (272 per run of ~25 particles), not by HBM; see DESIGN.md 4.2"}}


def workload_name(args, world):
    ppc = args.ppc2[0] * args.ppc2[1]
    if args.workload == "c5":
        return ("C5 shard: BASELINE configs[4] (2D 2048x2048 cells x 256 ppc over 8 GPUs, sixteen 512^2 boxes) at %d GPU(s): "
                "%d boxes of %dx%d cells per GPU, 2 species x %d ppc, CC1 gather/deposit, Picard particle loop (rtol 1e-12, "
                "iter_max %d), %d nonlinear evaluations per step, ghost-J add-exchange + particle migration between all boxes"
                % (world, args.boxes_per_gpu, args.ncell, args.ncell, ppc, args.iter_max, args.n_outer))
    return ("C3: 2D implicit energy-conserving PIC, %dx%d cells per GPU box, 2 species x %d ppc, CC1 gather/deposit, Picard "
            "particle loop (rtol 1e-12, iter_max %d), %d nonlinear evaluations per step"
            % (args.ncell, args.ncell, ppc, args.iter_max, args.n_outer))


def c5_shard_leg(args, rank, local, stream, region, capi):
    """The per-GPU shard of C5 on ONE GPU (two 512^2 boxes x 2 species x 128 ppc = 1.34e8 particles, periodic, with the
    box-to-box ghost-J exchange and migration): the like-for-like base of the weak-scaling numbers `--gpus N` reports for
    N > 1, which run this shard on every GPU."""
    import copy
    a = copy.copy(args)
    a.workload, a.boxes_per_gpu, a.ppc2 = "c5", 2, (16, 8)
    a.steps, a.warmup = max(4, min(args.steps, 8)), 3
    eng = Engine(a, rank, 1, local, stream=stream)
    region(a.warmup, False, eng=eng)
    ms, _ = region(a.steps, False, profile=True, eng=eng)
    k_ms, k_n = capi.profile_query("advance_cc1_fused")
    adv, app, unconv = capi.picard_totals(reset=True)
    n = eng.n_particles
    units = float(n) * a.n_outer * a.steps
    per_launch = n / len(eng.species)
    e2e = None
    if not args.no_e2e:      # the same shard through host field / J buffers: the base of the e2e scaling numbers
        region(1, True, eng=eng)
        ms_e, _ = region(a.steps, True, eng=eng)
        e2e = {"value": units / (ms_e * 1e-3), "unit": "particle-advances/s", "ms_per_step": ms_e / a.steps,
               "h2d_bytes_per_step": int(eng.h2d), "d2h_bytes_per_step": int(eng.d2h)}
    out = {"metric": "particle-advances/s (implicit push + deposit)", "value": units / (ms * 1e-3), "e2e": e2e,
           "unit": "particle-advances/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
           "workload": workload_name(a, 1), "particles_per_gpu": n, "mean_picard_passes": round(app / max(adv, 1), 3),
           "kernel_ms_per_launch": k_ms / max(k_n, 1), "units_per_launch": per_launch,
           "roofline_frac_kernel": (BYTES_PER_ADVANCE_2D * per_launch / (k_ms / max(k_n, 1) * 1e-3) / 1e9
                                    / region.peak) if k_n else None,
           "section_ms_per_step": getattr(region, "sections", None)}
    eng.destroy()
    return out


def variants_leg(args, rank, local, stream, region, capi):
    """The C3 box under other particle-Picard workloads, so that the headline fraction cannot be read as tuned to one
    k (mean passes per advance): iter_max_particles = 0 (BASELINE configs[2] lists {0, 21}: one pass, k = 1), fields
    150x stronger with 64x shorter wavelengths (8, 4 and 2.7 cells: k ~ 4, and four times as many particles cross a
    dual-cell face and take the deferred kernel; measured on the way: amplitude x8 -> k 2.54, x300 -> 3.2, x100 with the
    short waves -> 3.9), and the explicit leap-frog step (PIC_EM_EXPLICIT: gather + Boris + move + deposit + second half in
    ONE kernel, pgpu_explicit_step).  Device-timed like `value`, fewer steps."""
    import copy
    out = {}
    nsteps, nwarm = max(4, min(args.steps, 6)), 3
    for label, over in (("iter_max_0", {"iter_max": 0}), ("hard_fields", {"field_scale": 150.0 * args.field_scale, "field_kmul": 64})):
        a = copy.copy(args)
        for k, v in over.items():
            setattr(a, k, v)
        a.steps, a.warmup = nsteps, nwarm
        eng = Engine(a, rank, 1, local, stream=stream)
        region(a.warmup, False, eng=eng)
        ms, _ = region(a.steps, False, profile=True, eng=eng)
        k_ms, k_n = capi.profile_query("advance_cc1_fused")
        d_ms, _ = capi.profile_query("advance_deferred")
        m_ms, _ = capi.profile_query("advance_multiseg")
        adv, app, unconv = capi.picard_totals(reset=True)
        n = eng.n_particles
        per_launch = n / len(eng.species)
        out[label] = {"value": float(n) * a.n_outer * a.steps / (ms * 1e-3), "unit": "particle-advances/s",
                      "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
                      "mean_picard_passes": round(app / max(adv, 1), 3), "unconverged_particles": int(unconv),
                      "kernel_ms_per_launch": k_ms / max(k_n, 1), "multiseg_ms_per_step": m_ms / a.steps,
                      "deferred_ms_per_step": d_ms / a.steps,
                      "roofline_frac_kernel": (BYTES_PER_ADVANCE_2D * per_launch / (k_ms / max(k_n, 1) * 1e-3) / 1e9
                                               / region.peak) if k_n else None,
                      "overrides": over}
        eng.destroy()
    # explicit leap-frog: one advance = one particle through one fused step
    a = copy.copy(args)
    eng = Engine(a, rank, 1, local, stream=stream)
    torch = eng.torch
    bc = (1, 1)

    def estep():
        for bx in eng.boxes:
            bx.grid.current_zero()
            for sp in bx.species:
                sp.update_old_positions()
                sp.update_old_velocities()
                sp.explicit_step(eng.deck.dt, bc, bc, True)
                bx.grid.current_add(sp)
            bx.grid.current_finalize()
        eng.step_no += 1
        if a.sort_every > 0 and eng.step_no % a.sort_every == 0:
            for sp in eng.species:
                eng._sort(sp)

    for _ in range(nwarm):
        estep()
    eng.sync()
    capi.profile_reset()
    capi.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(nsteps):
        estep()
    e1.record(stream)
    eng.sync()
    torch.cuda.synchronize()
    capi.profile_enable(False)
    ms = e0.elapsed_time(e1)
    k_ms, k_n = capi.profile_query("explicit_step_cc1")
    d_ms, _ = capi.profile_query("explicit_step_deferred")
    n = eng.n_particles
    per_launch = n / len(eng.species)
    out["explicit_leapfrog"] = {
        "value": float(n) * nsteps / (ms * 1e-3), "unit": "particle-advances/s (explicit: gather + Boris + move + deposit)",
        "steps": nsteps, "warmup": nwarm, "ms_per_step": ms / nsteps,
        "kernel": "explicit_step_cc1 (the CC1 tile kernel in its one-pass explicit mode; particles that cross a dual-cell "
                  "face go to k_explicit_step)",
        "kernel_ms_per_launch": k_ms / max(k_n, 1), "deferred_ms_per_step": d_ms / nsteps,
        "roofline_frac_kernel": (BYTES_PER_ADVANCE_2D * per_launch / (k_ms / max(k_n, 1) * 1e-3) / 1e9 / region.peak)
                                if k_n else None,
        "bytes_per_unit": BYTES_PER_ADVANCE_2D}
    eng.destroy()
    return out


def bind_to_gpu_cpus(device):
    """Pin this rank (and the pinned host buffers it first-touches afterwards) to the CPUs NVML reports as local to its GPU:
    with eight ranks on one host the field uploads / J downloads of the e2e arm otherwise cross the socket interconnect."""
    if os.environ.get("PGPU_BENCH_NO_AFFINITY"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[device]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else device
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        bind_to_gpu_cpus.original = os.sched_getaffinity(0)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        return {"cpus_before": before, "cpus_bound": len(after), "first": after[0], "last": after[-1]}
    except Exception as e:      # no NVML, container without the sysfs topology, ...: run unbound
        return {"error": repr(e)[:120]}


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torchrun (one rank per GPU)" % args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_cpus(local)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: the image exports NCCL_DEBUG=VERSION, whose only effect is a
        # "NCCL version ..." banner on stdout (WARN prints it too); INFO etc. set by a user are left alone
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from picnic_b200 import capi
    capi.load()
    capi.init(local)
    # run the library on a torch stream so that torch.cuda.Event brackets its kernels and the
    # NCCL exchanges order against them without host synchronisation
    stream = torch.cuda.Stream(device=local)
    capi.check(capi.load().pgpu_set_stream(stream.cuda_stream))
    eng = Engine(args, rank, world, local, stream=stream)

    def barrier():
        capi.check(capi.load().pgpu_synchronize())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def region(nsteps, host_io, profile=False, eng=None):
        eng = eng or region.eng
        barrier()
        if profile:
            capi.profile_reset()
            capi.profile_enable(True)
            capi.picard_totals(reset=True)
        launches0 = capi.load().pgpu_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if profile:
            eng.sections_begin()
        for _ in range(nsteps):
            eng.step(host_io)
        e1.record(stream)
        barrier()
        if profile:
            capi.profile_enable(False)
            dev_s, host_s = eng.sections_end()
            region.sections = {k: round(v / nsteps, 4) for k, v in dev_s.items()}
            region.host_sections = {k: round(v / nsteps, 4) for k, v in host_s.items()}
            if world > 1:      # slowest rank per section
                names = sorted(dev_s)
                t = torch.tensor([dev_s[k] / nsteps for k in names], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                region.sections_max = {k: round(float(v), 4) for k, v in zip(names, t.tolist())}
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, capi.load().pgpu_launch_count() - launches0

    region.eng = eng
    # warm-up (>= 3 steps), then the timed region.  The particle arrays (5 GB per GPU) are far
    # larger than the 126 MB L2, so every pass streams from HBM; no explicit flush is needed.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()           # nvidia-smi needs ~0.1 s to deliver its first sample: start before the warm-up
    region(max(args.warmup, 3), False)
    t_host0 = time.perf_counter()
    ms, launches = region(args.steps, False, profile=True)
    if rank == 0:
        sampler.window(t_host0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None
    prof = {}
    for name in ("advance_cc1_fused", "advance_multiseg", "advance_deferred", "advance_deposit_fused", "advance", "deposit_current", "bin_", "finish_step", "second_half", "fold_periodic",
                 "current_add", "current_scale", "bc_periodic", "halo_", "mig_", "build_tables", "tile_boxes",
                 "bin_key", "bin_sort", "bin_permute", "bin_starts"):
        prof[name] = capi.profile_query(name)
    adv, app, unconv = capi.picard_totals(reset=True)
    k_mean = app / max(adv, 1)
    # share of the particles the tile kernel handed to the generic kernel in the last evaluation (diagnostic read-back)
    deferred_frac = sum(sp.deferred_count() for sp in eng.species) / max(eng.n_particles, 1)

    # weak scaling: every rank loads the same number of particles and migration conserves the sum
    n_total = eng.n_particles * world
    units = float(n_total) * args.n_outer * args.steps
    value = units / (ms * 1e-3)

    e2e = None
    if not args.no_e2e:
        region(1, True)
        ms_e, _ = region(args.steps, True)
        e2e = {"value": units / (ms_e * 1e-3), "unit": "particle-advances/s",
               "h2d_bytes_per_step": int(eng.h2d), "d2h_bytes_per_step": int(eng.d2h),
               "ms_per_step": ms_e / args.steps}

    # migration must conserve the particles of the whole domain
    n_now = sum(sp.n for sp in eng.species)
    if world > 1:
        t = torch.tensor([n_now], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        n_now = int(t.item())
    assert n_now == n_total, "particles lost in migration: %d != %d" % (n_now, n_total)

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        region.peak = peak
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        k_name = "advance_cc1_fused"
        k_ms, k_n = prof[k_name]
        if k_n == 0:
            k_name = "advance_deposit_fused"
            k_ms, k_n = prof[k_name]
        per_launch_units = eng.n_particles / len(eng.species)
        # DRAM bytes of one launch of the headline kernel from the committed `ncu --set full` capture
        # (dram__bytes_read.sum + dram__bytes_write.sum); only quoted for the launch size it was taken at
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["advance_cc1_fused"]
            if int(tr["particles_per_launch"]) == int(per_launch_units):
                traffic = float(tr["dram_bytes_per_launch"])
        except Exception:
            pass
        achieved = (BYTES_PER_ADVANCE_2D * per_launch_units) / (k_ms / max(k_n, 1) * 1e-3) / 1e9 if k_n else None
        out = {
            "metric": "particle-advances/s (implicit push + deposit)", "value": value, "unit": "particle-advances/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, world),
                       "particles_per_gpu": eng.n_particles, "particles_total": n_total,
                       "boxes": "%dx%d" % eng.layout, "boxes_per_gpu": args.boxes_per_gpu, "dt": args.dt,
                       "n_outer": args.n_outer, "sort_every": args.sort_every, "cpu_affinity_rank0": numa,
                       "exchange": (None if world == 1 else
                                    {"ghost_J": "peer-memory kernels over NVLink (CUDA IPC inboxes)" if args.halo == "peer"
                                                else "pack + NCCL send/recv + unpack-add",
                                     "migration": "peer-memory inboxes, device-side counts, one host wait per step"
                                                  if args.migration == "peer" else "count all-gather + NCCL send/recv",
                                     "ghost_J_bytes_per_evaluation": eng.halo_bytes,
                                     "migrated_particles_rank0": int(eng.migrated)}),
                       "mean_picard_passes": round(k_mean, 3), "unconverged_particles": int(unconv),
                       "deferred_fraction_rank0": round(deferred_frac, 5),
                       "l2_policy": "inputs (%.1f GB particle SoA per GPU) exceed the 126 MB L2; no flush needed"
                                    % (eng.n_particles * 96 / 1e9)},
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "kernel": k_name + " (fused gather + Boris + particle-Picard + deposit)",
                         "bytes_per_unit": BYTES_PER_ADVANCE_2D, "units_per_launch": per_launch_units,
                         "kernel_ms_per_launch": k_ms / max(k_n, 1), "kernel_share_of_step": k_ms / ms,
                         "peak_source": peak_src},
            "kernel_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in prof.items() if v[1]},
            "section_ms_per_step": getattr(region, "sections", None),   # rank 0, CUDA events between the phases of a step
            "section_ms_per_step_slowest_rank": getattr(region, "sections_max", None),
            "host_issue_ms_per_step": getattr(region, "host_sections", None),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args)
    eng.destroy()
    if rank == 0 and world == 1 and args.workload == "c3" and not args.no_c5_shard:
        out["c5_shard"] = c5_shard_leg(args, rank, local, stream, region, capi)
    if rank == 0 and world == 1 and args.workload == "c3" and not args.no_variants:
        out["variants"] = variants_leg(args, rank, local, stream, region, capi)
    if rank == 0 and world == 1 and not args.no_collisions:
        out["collisions"] = collisions_leg(args, torch, capi, stream, out["roofline"]["peak"])
    if rank == 0 and world == 1 and not args.no_c4:
        out["c4_1d"] = c4_leg(args, torch, capi, stream, out["roofline"]["peak"])
    if rank == 0 and world == 1 and not args.no_mass_matrix:
        out["mass_matrices"] = mass_matrix_leg(args, torch, capi, stream, out["roofline"]["peak"])
    capi.finalize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


# ------------------------------------------------------------------------------------------
# the CPU arm: the oracle (restatement of the reference algorithm) on the host cores
# ------------------------------------------------------------------------------------------
def cpu_sample_problem(args, nthreads):
    """A strip of the C3 box on one thread (the layout comparison below)."""
    from oracle import oracle as orc
    deck, _ = make_deck(args, 1)
    target = args.cpu_sample or 1000000 * max(nthreads, 1)          # particles per species
    rows = max(1, min(args.ncell, int(round(target / (args.ncell * args.ppc ** 2)))))
    E0, B0 = field_amplitudes(deck)
    lo, hi = (0, 0), (args.ncell - 1, args.ncell - 1)
    E, B = decks.analytic_fields(deck, lo, hi, E0=E0, B0=B0)
    geom = orc.make_geom(2, deck.xmin, deck.xmax, deck.dx, deck.nghost)
    Ef = [orc.Fab(l, h, a) for (l, h, a) in E]
    Bf = [orc.Fab(l, h, a) for (l, h, a) in B]
    rng = np.random.default_rng(deck.seed)
    row0 = args.ncell // 2 - rows // 2
    parts = [decks.load_species(deck, s, (0, row0), (args.ncell - 1, row0 + rows - 1), rng) for s in deck.species]
    return orc, deck, geom, Ef, Bf, parts, rows


def cpu_arm(args, steps, warmup, seconds):
    """The reference's CPU path (oracle port) as R box-owning workers (oracle/cpu_boxes.py): one square box per host
    thread, the same species / ppc / dx / dt / field amplitudes / Picard tolerances / evaluations per step as the GPU
    arm, ghost-J add-exchange after every deposit and particle migration once per step through shared memory.  The
    box edge is sized from a one-thread probe so that warmup + steps take about `seconds`."""
    from oracle import cpu_boxes
    nthreads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def deck_fn(ncell):
        d = decks.deck_c3(ncell=args.ncell, ppc=args.ppc, dt=args.dt, iter_max=args.iter_max)
        d.species = decks.electron_proton(tuple(args.ppc2))
        d.ncell = tuple(ncell)
        return d

    def amps(deck):
        return field_amplitudes(deck, args.field_scale)

    ppc_tot = 2 * args.ppc2[0] * args.ppc2[1]
    if args.cpu_sample:
        bn = max(8, int(round((args.cpu_sample * 2.0 / ppc_tot / max(nthreads, 1)) ** 0.5)))
    else:
        rate1 = cpu_boxes.probe_rate(deck_fn, amps, args.n_outer, EPS_OUTER)
        per_worker = rate1 * seconds / ((steps + warmup) * args.n_outer)      # particles a worker can own
        bn = int(max(8, min(args.ncell, (per_worker / ppc_tot) ** 0.5)))
    r = cpu_boxes.run(deck_fn, amps, nthreads, steps, warmup, args.n_outer, EPS_OUTER, bn)
    res = {"value": r["units"] / r["seconds"], "unit": "particle-advances/s", "cores": r["workers"], "kind": "port",
           "sample": "%s boxes of %d^2 cells (periodic), 2 species x %d ppc = %d particles, one box-owning worker "
                     "thread per box with ghost-J add-exchange per evaluation and migration per step (%d moved), "
                     "%d step(s) + %d warm-up x %d evaluations, oracle (g++ -O2 -ffp-contract=off)"
                     % (r["boxes"], r["box_cells"], ppc_tot // 2, r["particles"], r["migrated"], steps, warmup,
                        args.n_outer),
           "mean_picard_passes": round(r["mean_picard_passes"], 3),
           "host_threads_visible": nthreads, "seconds": r["seconds"]}
    return res, r["seconds"] / max(steps, 1)


def cpu_layout_sample(args, n_sample=400000):
    """One host thread, one nonlinear evaluation of one species on a strip of the C3 box, twice: the SoA oracle (what
    `cpu_baseline.value` runs on every core) and the same arithmetic on 176-byte particle objects in a doubly linked
    list with one kernel call per particle (oracle_aos.cpp) -- the reference's memory behaviour; "scattered"
    places consecutive particles' nodes at random heap addresses, as after many steps of list transfers."""
    orc, deck, geom, Ef, Bf, parts, rows = cpu_sample_problem(args, 1)
    lo, hi = (0, 0), (args.ncell - 1, args.ncell - 1)
    p, sdef = parts[0], deck.species[0]
    n = min(n_sample, p["w"].size)
    x = np.ascontiguousarray(p["x"][:, :n]); v = np.ascontiguousarray(p["v"][:, :n]); w = np.ascontiguousarray(p["w"][:n])
    fn = sdef.fnorm_const(deck.units)
    out = {}
    for name in ("soa", "aos_list", "aos_list_scattered"):
        J = [orc.fab_for(lo, hi, deck.nghost, st) for st in orc.E_STAG[2]]
        if name == "soa":
            xa, va = x.copy(), v.copy()
            t0 = time.perf_counter()
            rc, _, _, _ = orc.advance_particles_iteratively(geom, deck.interp_E, xa, x, va, v, Ef, Bf, fn, deck.cnorm_dt,
                                                            deck.rtol, deck.iter_max)
            orc.deposit_current(geom, deck.interp_J, xa, x, va, w, deck.cnorm_dt, J)
            sec = time.perf_counter() - t0
        else:
            aos = orc.AosList(2, x, x, v, v, w, scattered=name.endswith("scattered"))
            t0 = time.perf_counter()
            rc, _, _ = aos.advance_deposit(geom, deck.interp_E, Ef, Bf, fn, deck.cnorm_dt, deck.rtol, deck.iter_max, J)
            sec = time.perf_counter() - t0
            aos.destroy()
        out[name] = {"value": n / sec, "seconds": sec, "rc": int(rc)}
    out["unit"] = "particle-advances/s on one thread"
    out["sample"] = "%d particles of one species, one evaluation" % n
    return out


def cpu_baseline(args, steps=2):
    orig = getattr(bind_to_gpu_cpus, "original", None)
    if orig:                      # the CPU arm runs on every core the process started with
        os.sched_setaffinity(0, orig)
    res, _ = cpu_arm(args, steps=steps, warmup=1, seconds=15.0)
    try:
        res["layout"] = cpu_layout_sample(args)
    except Exception as e:      # the layout comparison is an extra; never lose the baseline over it
        res["layout"] = {"error": repr(e)}
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the same --steps / --warmup as the GPU arm; the sample (box edge) shrinks so that the run takes ~90 s
    res, sec_per_step = cpu_arm(args, steps=args.steps, warmup=args.warmup, seconds=90.0)
    out = {"impl": "reference", "metric": "particle-advances/s (implicit push + deposit)", "value": res["value"],
           "unit": "particle-advances/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": "C3 (bounded sample): " + res["sample"], "dt": args.dt, "n_outer": args.n_outer,
                      "mean_picard_passes": res["mean_picard_passes"]},
           "cpu_baseline": res,
           "e2e": {"value": res["value"], "unit": "particle-advances/s", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "note": "reference = CPU oracle port of the reference algorithm; the reference binary needs "
                   "Chombo + gfortran + MPI + HDF5 and cannot be built (DESIGN.md)"}
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
