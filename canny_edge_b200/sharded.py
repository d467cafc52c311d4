"""Sharding of the Canny hot path across GPUs: one process per GPU, torch.distributed as plumbing.

Two ways the path shards (the reference itself is single-GPU, single-frame, src/main.cpp:111):

* frames   — independent units; rank r owns frames [r*n, (r+1)*n).  No data-path collective.
             (`frame_slice`; bench.py uses it for BASELINE configs[2].)
* row bands of ONE image — rank r owns global rows [r*H/G, (r+1)*H/G).  Two exchange steps:
     1. halo rows: window/2 + 2 input rows from each neighbour (`exchange_halos`: batched isend/irecv —
        NCCL P2P over NVLink on GPUs, gloo on CPU tensors in the tests);
     2. hysteresis label merge: every band exports its first/last row as (label, flags) records
        (`b200_band_boundary_export`), ONE all-gather (`gather_records`), then every rank unions the records
        that touch across a boundary and finalises its own band (`b200_band_finalize`).
  Both steps now live in C (`b200_bands_*`, csrc/bands_mgpu.cu: copy-engine halo pulls over NVLink peer mappings + sparse
  boundary records read from peer memory, or NCCL send/recv + all-gather); `BandPipeline` is a thin caller of that handle and
  torch.distributed only carries the 128-byte NCCL unique id at creation.  `TorchBandPipeline` is the round-1 form (the two
  exchanges as torch.distributed calls around b200_band_front / _boundary_export / _finalize), kept as a cross-check and because
  its exchange helpers are device-agnostic (the CPU tests run them over gloo).

`canny_bands_virtual` runs G bands of one image as an in-process group on ONE GPU through the same C pipeline
(`b200_bands_create_group`: peers are plain pointers), so halo pulls, flags and the cross-band merge are testable without a
multi-GPU box.

All compute is in libcanny_b200.so (sm_100a CUDA); nothing here computes pixels on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from ._lib import check, load
from .api import Context

RECORD_BYTES = 8  # b200_band_record: int32 label, int32 flags


# ---------------------------------------------------------------------------------------------------
# geometry
# ---------------------------------------------------------------------------------------------------
def frame_slice(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of frames owned by `rank`: (first, count). Remainders go to the low ranks."""
    base, rem = divmod(n_frames, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


@dataclass(frozen=True)
class BandGeometry:
    rank: int
    world: int
    height: int       # global image height
    width: int
    row0: int         # first global row owned
    rows: int         # rows owned
    halo: int         # rows needed beyond an interior band edge (window/2 + 2)
    halo_above: int   # rows actually present above (0 at the image top)
    halo_below: int

    @property
    def buffer_rows(self) -> int:
        return self.halo_above + self.rows + self.halo_below


def band_geometry(height: int, width: int, rank: int, world: int, sigma: float) -> BandGeometry:
    """Row band of `rank`. Every band must be at least `halo` rows tall so a halo comes from ONE neighbour."""
    halo = int(load().b200_band_halo_rows(C.c_float(sigma)))
    base, rem = divmod(height, world)
    row0 = rank * base + min(rank, rem)
    rows = base + (1 if rank < rem else 0)
    if world > 1 and base < max(halo, 2):
        raise ValueError(f"bands of {base} rows are shorter than the {halo}-row halo; use fewer ranks")
    above = min(halo, row0)
    below = min(halo, height - (row0 + rows))
    return BandGeometry(rank, world, height, width, row0, rows, halo, above, below)


# ---------------------------------------------------------------------------------------------------
# the two exchange steps (device-agnostic torch code: NCCL on GPUs, gloo on CPU tensors)
# ---------------------------------------------------------------------------------------------------
def exchange_halos(band, geo: BandGeometry, group=None):
    """band: uint8 tensor (geo.rows, W) owned by this rank.  Returns a (geo.buffer_rows, W) tensor
    [halo from rank-1 | band | halo from rank+1] on the same device.  One batched isend/irecv round."""
    import torch
    import torch.distributed as dist

    W = geo.width
    buf = torch.empty((geo.buffer_rows, W), dtype=torch.uint8, device=band.device)
    buf[geo.halo_above:geo.halo_above + geo.rows].copy_(band)
    if geo.world == 1:
        return buf
    ops = []
    up, down = geo.rank - 1, geo.rank + 1
    # what the neighbours need from me is THEIR halo size, which equals geo.halo for interior edges
    send_up = band[:geo.halo].contiguous() if up >= 0 else None
    send_down = band[geo.rows - geo.halo:].contiguous() if down < geo.world else None
    recv_up = buf[:geo.halo_above] if up >= 0 else None
    recv_down = buf[geo.halo_above + geo.rows:] if down < geo.world else None
    if up >= 0:
        ops.append(dist.P2POp(dist.isend, send_up, up, group))
        ops.append(dist.P2POp(dist.irecv, recv_up, up, group))
    if down < geo.world:
        ops.append(dist.P2POp(dist.isend, send_down, down, group))
        ops.append(dist.P2POp(dist.irecv, recv_down, down, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return buf


def gather_records(records, world: int, group=None):
    """records: uint8 tensor of this band's boundary records (n_records*8 bytes).  Returns all bands'
    records concatenated in band order (world*n_records*8 bytes) — the hysteresis path's one collective."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return records
    out = torch.empty((world * records.numel(),), dtype=torch.uint8, device=records.device)
    dist.all_gather_into_tensor(out, records.contiguous(), group=group)
    return out


# ---------------------------------------------------------------------------------------------------
# band pipeline on one rank
# ---------------------------------------------------------------------------------------------------
class TorchBandPipeline:
    """(Round-1 form: exchanges through torch.distributed.)  Canny of one row band of a larger image on this rank's GPU.

    The pipeline owns ONE persistent device buffer [halo above | band | halo below] (`buffer`), so a step neither allocates nor
    copies the band: callers that produce their rows on the device write them straight into `band_view()`; `run(band)` with an
    external tensor copies it in first."""

    def __init__(self, ctx: Context, height: int, width: int, rank: int, world: int, sigma: float, min_val: int,
                 max_val: int, group=None):
        self.ctx, self.group = ctx, group
        self.sigma, self.lo, self.hi = float(sigma), int(min_val), int(max_val)
        self.geo = band_geometry(height, width, rank, world, sigma)
        self.lib = load()
        self.n_records = int(self.lib.b200_band_record_count(width))
        self.buffer = None      # (buffer_rows, W) uint8
        self.records = None     # this band's boundary records (bytes)
        self.all_records = None
        self.timings = None     # set to a dict to collect per-stage CUDA-event times (ms) of the next run()

    # ---- persistent device state -------------------------------------------------------------------------------
    def _ensure(self, device):
        import torch

        g = self.geo
        if self.buffer is None or self.buffer.device != device:
            self.buffer = torch.empty((g.buffer_rows, g.width), dtype=torch.uint8, device=device)
            self.records = torch.empty((self.n_records * RECORD_BYTES,), dtype=torch.uint8, device=device)
            self.all_records = torch.empty((g.world * self.n_records * RECORD_BYTES,), dtype=torch.uint8, device=device)

    def band_view(self, device=None):
        """The (rows, W) window of the persistent buffer that holds this rank's own rows."""
        import torch

        self._ensure(torch.device("cuda", self.ctx.device) if device is None else device)
        g = self.geo
        return self.buffer[g.halo_above:g.halo_above + g.rows]

    # ---- the two exchange steps, in place ----------------------------------------------------------------------
    def exchange_halos_inplace(self):
        """Fills the halo rows of the persistent buffer from the neighbours' edge rows (one batched NCCL send/recv round)."""
        import torch.distributed as dist

        g = self.geo
        if g.world == 1:
            return
        buf, a = self.buffer, g.halo_above
        ops = []
        if g.rank > 0:
            ops.append(dist.P2POp(dist.isend, buf[a:a + g.halo], g.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, buf[:a], g.rank - 1, self.group))
        if g.rank + 1 < g.world:
            ops.append(dist.P2POp(dist.isend, buf[a + g.rows - g.halo:a + g.rows], g.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, buf[a + g.rows:], g.rank + 1, self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def front(self, buf, edges) -> None:
        g = self.geo
        check(self.lib.b200_band_front(self.ctx.handle, buf.data_ptr(), g.halo_above, g.halo_below, g.rows, g.row0, g.height,
                                       g.width, C.c_float(self.sigma), self.lo, self.hi, edges.data_ptr()))

    def export(self, records) -> None:
        check(self.lib.b200_band_boundary_export(self.ctx.handle, self.geo.rows, self.geo.width, records.data_ptr()))

    def finalize(self, all_records, edges) -> None:
        g = self.geo
        check(self.lib.b200_band_finalize(self.ctx.handle, all_records.data_ptr(), g.world, g.rank, g.rows, g.width, edges.data_ptr()))

    def run(self, band=None, edges=None):
        """band: (rows, W) uint8 CUDA tensor with this rank's rows, or None when they were written into band_view().
        Returns the (rows, W) uint8 0/255 edge band.  Work is issued on torch's current stream (the context is pointed at it), so
        NCCL ordering is the stream's ordering."""
        import torch
        import torch.distributed as dist

        g = self.geo
        dev = band.device if band is not None else torch.device("cuda", self.ctx.device)
        self._ensure(dev)
        self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        if edges is None:
            edges = torch.empty((g.rows, g.width), dtype=torch.uint8, device=dev)
        marks = []

        def mark(name):
            if self.timings is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        if band is not None and band.data_ptr() != self.band_view().data_ptr():
            self.band_view().copy_(band)
        mark("copy_in")
        self.exchange_halos_inplace()
        mark("halo_exchange")
        self.front(self.buffer, edges)
        mark("front+label")
        self.export(self.records)
        mark("export")
        if g.world > 1:
            dist.all_gather_into_tensor(self.all_records, self.records, group=self.group)
            allrec = self.all_records
        else:
            allrec = self.records
        mark("all_gather")
        self.finalize(allrec, edges)
        mark("finalize")
        if self.timings is not None:
            torch.cuda.synchronize()
            for (_, e0), (name, e1) in zip(marks, marks[1:]):
                self.timings[name] = self.timings.get(name, 0.0) + e0.elapsed_time(e1)
        return edges


STAGES = ("signal+interior", "halo_wait", "edges+label", "export", "record_exchange", "finalize")
TRANSPORTS = {0: "single band", 1: "p2p (copy-engine halo pulls + sparse records over NVLink peer mappings)", 2: "nccl (send/recv + all-gather)"}


class BandPipeline:
    """One rank's band of a larger image through the C pipeline (b200_bands_*).  The handle owns the persistent
    [halo | band | halo] buffer: write the rank's rows into `band_view()` and call `run()`.

    world > 1: torch.distributed must be initialised; it only broadcasts the NCCL unique id the C side builds its communicator from."""

    def __init__(self, ctx: Context, height: int, width: int, rank: int, world: int, sigma: float, min_val: int,
                 max_val: int, group=None):
        import torch

        self.ctx, self.lib = ctx, load()
        self.geo = band_geometry(height, width, rank, world, sigma)
        self.handle = C.c_void_p()
        uid = None
        if world > 1:
            import torch.distributed as dist

            raw = (C.c_ubyte * 128)()
            if rank == 0:
                check(self.lib.b200_bands_unique_id(raw))
            t = torch.tensor(list(raw), dtype=torch.uint8)
            if dist.get_backend(group) == "nccl":
                t = t.cuda(ctx.device)
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            uid = (C.c_ubyte * 128)(*t.cpu().tolist())
        check(self.lib.b200_bands_create(ctx.handle, None, uid, rank, world, height, width, C.c_float(sigma), int(min_val),
                                         int(max_val), C.byref(self.handle)))
        r0, rows, halo, tr = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(self.lib.b200_bands_info(self.handle, C.byref(r0), C.byref(rows), C.byref(halo), C.byref(tr)))
        assert (r0.value, rows.value, halo.value) == (self.geo.row0, self.geo.rows, self.geo.halo)
        self.transport = tr.value
        self.timings = None

    def band_view(self, device=None):
        """(rows, W) uint8 CUDA tensor aliasing the band's own rows inside the handle's buffer."""
        ptr = C.c_void_p()
        check(self.lib.b200_bands_input(self.handle, C.byref(ptr)))
        return _tensor_from_ptr(ptr.value, (self.geo.rows, self.geo.width), self.ctx.device)

    def run(self, band=None, edges=None):
        import torch

        g = self.geo
        self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        if edges is None:
            edges = torch.empty((g.rows, g.width), dtype=torch.uint8, device=torch.device("cuda", self.ctx.device))
        if band is not None:
            self.band_view().copy_(band)
        check(self.lib.b200_bands_set_timing(self.handle, 1 if self.timings is not None else 0))
        check(self.lib.b200_bands_run(self.handle, edges.data_ptr()))
        if self.timings is not None:
            ms = (C.c_float * 6)()
            check(self.lib.b200_bands_stage_ms(self.handle, ms))
            for name, v in zip(STAGES, ms):
                self.timings[name] = self.timings.get(name, 0.0) + float(v)
        return edges

    def check(self):
        check(self.lib.b200_bands_check(self.handle))

    def close(self):
        if self.handle:
            check(self.lib.b200_bands_destroy(self.handle))
            self.handle = C.c_void_p()


def _tensor_from_ptr(ptr: int, shape, device: int):
    """uint8 CUDA tensor over existing device memory (no copy, no ownership) via __cuda_array_interface__."""
    import torch

    class _Mem:
        pass

    m = _Mem()
    m.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}
    return torch.as_tensor(m, device=torch.device("cuda", device))


def canny_bands_virtual(img: np.ndarray, n_bands: int, sigma: float, min_val: int, max_val: int, device: int = 0,
                        steps: int = 1) -> np.ndarray:
    """All bands of one image as an in-process group on ONE GPU (one context per band) through b200_bands_run_group: the same halo
    pulls, ready flags, split front launches, sparse record exchange and cross-band merge as the multi-GPU run, with plain
    pointers in place of IPC mappings.  Returns the (H, W) uint8 0/255 map."""
    import torch

    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    dev = torch.device("cuda", device)
    lib = load()
    ctxs: List[Context] = [Context(device) for _ in range(n_bands)]
    handles = (C.c_void_p * n_bands)()
    try:
        check(lib.b200_bands_create_group((C.c_void_p * n_bands)(*[c.handle for c in ctxs]), n_bands, H, W, C.c_float(sigma),
                                          int(min_val), int(max_val), handles))
        d_img = torch.from_numpy(img).to(dev)
        edges = []
        for b in range(n_bands):
            r0, rows = C.c_int(), C.c_int()
            check(lib.b200_bands_info(handles[b], C.byref(r0), C.byref(rows), None, None))
            ptr = C.c_void_p()
            check(lib.b200_bands_input(handles[b], C.byref(ptr)))
            _tensor_from_ptr(ptr.value, (rows.value, W), device).copy_(d_img[r0.value:r0.value + rows.value])
            edges.append(torch.empty((rows.value, W), dtype=torch.uint8, device=dev))
        torch.cuda.synchronize(dev)   # the copies ran on torch's stream; the bands work on their contexts' streams
        e_ptrs = (C.c_void_p * n_bands)(*[e.data_ptr() for e in edges])
        for _ in range(steps):
            check(lib.b200_bands_run_group(handles, n_bands, e_ptrs))
        for b in range(n_bands):
            check(lib.b200_bands_check(handles[b]))
        return torch.cat(edges).cpu().numpy()
    finally:
        for h in handles:
            if h:
                lib.b200_bands_destroy(h)
        for c in ctxs:
            c.close()


def canny_bands_virtual_torch(img: np.ndarray, n_bands: int, sigma: float, min_val: int, max_val: int, device: int = 0) -> np.ndarray:
    """All bands of one image on ONE GPU, sequentially, through the same band kernels (one context per band,
    because a context keeps its band's label state between front/export/finalize).  The two exchanges become
    local slicing and concatenation.  Returns the (H, W) uint8 0/255 map."""
    import torch

    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    dev = torch.device("cuda", device)
    d_img = torch.from_numpy(img).to(dev)
    ctxs: List[Context] = [Context(device) for _ in range(n_bands)]
    try:
        pipes = [TorchBandPipeline(ctxs[b], H, W, b, n_bands, sigma, min_val, max_val) for b in range(n_bands)]
        edges, recs = [], []
        for p in pipes:
            g = p.geo
            buf = d_img[g.row0 - g.halo_above:g.row0 + g.rows + g.halo_below].contiguous()
            e = torch.empty((g.rows, W), dtype=torch.uint8, device=dev)
            p.front(buf, e)
            r = torch.empty((p.n_records * RECORD_BYTES,), dtype=torch.uint8, device=dev)
            p.export(r)
            ctxs[g.rank].synchronize()
            edges.append(e)
            recs.append(r)
        all_records = torch.cat(recs)
        torch.cuda.synchronize(dev)   # torch.cat ran on torch's stream; the contexts finalise on their own streams
        for p, e in zip(pipes, edges):
            p.finalize(all_records, e)
            p.ctx.synchronize()
        return torch.cat(edges).cpu().numpy()
    finally:
        for c in ctxs:
            c.close()
