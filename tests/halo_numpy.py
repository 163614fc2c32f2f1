"""CPU stand-ins for the device side of picnic_b200/halo.py (tests only): the same pack /
unpack-add / mark / pack-leavers / append contract as the C ABI, on numpy arrays, so that the
index-box arithmetic and the message pattern can be exercised with gloo (or a mailbox) here."""
import numpy as np
import torch

from picnic_b200 import halo


class NumpyGridBackend:
    def __init__(self, layout, rank):
        self.device = "cpu"
        self.arr = []
        for stag in halo.STAG_J[layout.D]:
            lo, hi = layout.array_bounds(rank, stag)
            shape = tuple(h - l + 1 for l, h in zip(lo, hi))
            self.arr.append((lo, hi, np.zeros(shape, order="F")))

    def new_buffer(self, count):
        return torch.empty(count, dtype=torch.float64)

    def _view(self, comp, lo, hi):
        alo, _, a = self.arr[comp]
        return a[tuple(slice(l - al, h - al + 1) for l, h, al in zip(lo, hi, alo))]

    def pack(self, comp, lo, hi, buf):
        buf.copy_(torch.from_numpy(np.ascontiguousarray(self._view(comp, lo, hi).ravel(order="F"))))

    def unpack_add(self, comp, lo, hi, buf):
        v = self._view(comp, lo, hi)
        v += buf.numpy().reshape(v.shape, order="F")

    def sync(self):
        pass


class NumpySpeciesBackend:
    """Particles of one box; ownership rule of pgpu_species_mark_leavers."""

    def __init__(self, layout, rank, x, xold, v, vold, w, ids, xmin, dx):
        self.layout, self.rank, self.device = layout, rank, "cpu"
        self.D = layout.D
        self.nw = 2 * self.D + 8
        self.p = dict(x=x.copy(), xold=xold.copy(), v=v.copy(), vold=vold.copy(), w=w.copy(), id=ids.copy())
        self.xmin, self.dx = np.asarray(xmin, float), np.asarray(dx, float)
        self.codes = None

    @property
    def n(self):
        return self.p["w"].size

    def owner_codes(self):
        lay = self.layout
        my = lay.coords(self.rank)
        code = np.zeros(self.n, dtype=np.int64)
        mul = 1
        lost = np.zeros(self.n, dtype=bool)
        for d in range(self.D):
            b = np.floor((self.p["x"][d] - self.xmin[d]) / (self.dx[d] * lay.nbox[d])).astype(np.int64)
            diff = b - my[d]
            if lay.periodic[d]:
                diff = np.where(diff > 1, diff - lay.nb[d], diff)
                diff = np.where(diff < -1, diff + lay.nb[d], diff)
            lost |= (b < 0) | (b >= lay.nb[d]) | (np.abs(diff) > 1)
            code += (diff + 1) * mul
            mul *= 3
        if self.D == 1:
            code += 3
        return np.where(lost, 9, code)

    def mark_leavers(self):
        self.codes = self.owner_codes()
        return np.bincount(np.where(self.codes == 4, 10, self.codes), minlength=11)[:10].astype(np.int64)

    def new_buffer(self, nrec):
        return torch.empty(max(nrec, 1) * self.nw, dtype=torch.float64)

    def _records(self, idx):
        p = self.p
        cols = [p["x"][d][idx] for d in range(self.D)] + [p["xold"][d][idx] for d in range(self.D)]
        cols += [p["v"][c][idx] for c in range(3)] + [p["vold"][c][idx] for c in range(3)]
        cols += [p["w"][idx], p["id"][idx].view(np.float64)]
        return np.stack(cols, axis=1)

    def pack_leavers(self, buf):
        order = np.argsort(self.codes, kind="stable")
        order = order[(self.codes[order] != 4) & (self.codes[order] < 9)]
        rec = self._records(order)
        buf[:rec.size].copy_(torch.from_numpy(np.ascontiguousarray(rec).ravel()))
        keep = self.codes == 4
        for k, a in self.p.items():
            self.p[k] = a[..., keep].copy()

    def append(self, nrec, buf):
        rec = buf[:nrec * self.nw].numpy().reshape(nrec, self.nw)
        D, p = self.D, self.p
        p["x"] = np.concatenate([p["x"], rec[:, 0:D].T], axis=1)
        p["xold"] = np.concatenate([p["xold"], rec[:, D:2 * D].T], axis=1)
        p["v"] = np.concatenate([p["v"], rec[:, 2 * D:2 * D + 3].T], axis=1)
        p["vold"] = np.concatenate([p["vold"], rec[:, 2 * D + 3:2 * D + 6].T], axis=1)
        p["w"] = np.concatenate([p["w"], rec[:, 2 * D + 6]])
        p["id"] = np.concatenate([p["id"], np.ascontiguousarray(rec[:, 2 * D + 7]).view(np.uint64)])

    def sync(self):
        pass
