"""Host-side partitioning of the denoise workload over ranks (one process per GPU).

The reference is single-GPU; both partitionings below are net-new:
  * independent frame sequences, one context + stream per GPU, no data-path collective
    (BASELINE.json configs[4]);
  * one large frame split into contiguous row bands (BASELINE.json configs[3]): rows are contiguous in
    memory (pitch = W texels, reference include/extended_math.h:66-68), so a band plus its halo is one
    contiguous block.
torch.distributed is only the plumbing (rendezvous, max-over-ranks timing).
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist

# vertical reach of the passes below the temporal one (DESIGN.md "Row bands"):
#   variance 7x7 -> 3 rows; a-trous level l -> 2 * 2^l rows (+1 row for the 3x3 variance prefilter)
ATROUS_LEVEL_HALO = [2, 4, 8, 16, 32]
VARIANCE_HALO = 3


def assign_sequences(num_sequences: int, world_size: int):
    """sequence j -> rank j mod world_size; returns one list of sequence ids per rank."""
    if world_size < 1 or num_sequences < 0:
        raise ValueError("world_size >= 1 and num_sequences >= 0 required")
    return [list(range(r, num_sequences, world_size)) for r in range(world_size)]


@dataclass(frozen=True)
class Band:
    rank: int
    row0: int      # first owned row
    rows: int      # owned rows
    halo_top: int  # rows needed above row0 for a halo-recompute frame (clipped at the image border)
    halo_bot: int


def frame_halo(levels: int) -> int:
    """Rows of INPUT a band needs beyond its own rows so that every pass of a frame can be evaluated
    without mid-frame exchange (sum of the per-pass reaches; 62 + 1 + 3 = 66 for 5 levels)."""
    if not 0 <= levels <= len(ATROUS_LEVEL_HALO):
        raise ValueError("levels out of range")
    return sum(ATROUS_LEVEL_HALO[:levels]) + (1 if levels else 0) + VARIANCE_HALO


def row_bands(height: int, world_size: int, levels: int = 5):
    """Contiguous, near-equal row bands covering [0, height) exactly once."""
    if world_size < 1 or height < world_size:
        raise ValueError("need at least one row per rank")
    halo = frame_halo(levels)
    base, extra = divmod(height, world_size)
    bands, row0 = [], 0
    for r in range(world_size):
        rows = base + (1 if r < extra else 0)
        bands.append(Band(r, row0, rows, min(halo, row0), min(halo, height - (row0 + rows))))
        row0 += rows
    return bands


MOTION_MARGIN = 14  # rows: max |motion_y| + 1 (bilinear) + 1 (3x3 search) the banded mode tolerates


def banded_halo(levels: int = 5, motion_margin: int = MOTION_MARGIN) -> int:
    """Halo rows of a band context: filter reach + the rows its halo pixels reproject from."""
    return frame_halo(levels) + motion_margin


def neighbour_swap(rank, send_up, recv_up, send_dn, recv_dn):
    """The exchange step of the halo-recompute band scheme: rank r sends `send_up` to r - 1 and `send_dn` to r + 1 and
    receives their counterparts (None = no neighbour on that side: the first / last band).  Point-to-point only, one
    batch, no collective; works on any backend's tensors (NCCL on the GPU box, gloo in the CPU tests)."""
    ops = []
    if send_up is not None:
        ops += [dist.P2POp(dist.isend, send_up, rank - 1), dist.P2POp(dist.irecv, recv_up, rank - 1)]
    if send_dn is not None:
        ops += [dist.P2POp(dist.isend, send_dn, rank + 1), dist.P2POp(dist.irecv, recv_dn, rank + 1)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class BandedSvgf:
    """One rank's share of a row-banded frame: an ordinary SvgfContext over the band plus its halo
    (no mid-frame exchange: halo rows are recomputed), and the per-frame swap of history rows with the
    two neighbours.  `exchange` moves packed history buffers: torch.distributed send/recv between
    processes (one per GPU), or plain copies when several bands live in one process (tests)."""

    def __init__(self, width, full_height, band: Band, halo: int, device=0):
        from .api import SvgfContext
        import torch as _t
        self.band, self.W, self.H = band, width, full_height
        self.top = min(halo, band.row0)
        self.bot = min(halo, full_height - (band.row0 + band.rows))
        self.ext_row0 = band.row0 - self.top
        self.ext_rows = self.top + band.rows + self.bot
        if (self.top and band.rows < halo) or (self.bot and band.rows < halo):
            raise ValueError("band shorter than the halo its neighbours need")
        self.halo = halo
        self.ctx = SvgfContext(width, self.ext_rows, device)
        dev = _t.device("cuda", device)
        nb = self.ctx.history_bytes(halo)
        self.send_up = _t.empty(nb, dtype=_t.uint8, device=dev) if self.top else None
        self.send_dn = _t.empty(nb, dtype=_t.uint8, device=dev) if self.bot else None
        self.recv_up = _t.empty(nb, dtype=_t.uint8, device=dev) if self.top else None
        self.recv_dn = _t.empty(nb, dtype=_t.uint8, device=dev) if self.bot else None

    def slice_rows(self, plane):
        """Rows of a full-frame (H, W, C) plane this band's context consumes."""
        return plane[self.ext_row0:self.ext_row0 + self.ext_rows]

    def owned(self, ext_plane):
        """Owned rows of a band-local (ext_rows, W, C) plane."""
        return ext_plane[self.top:self.top + self.band.rows]

    def pack(self):
        """After a frame: my boundary rows that the neighbours hold as halo."""
        if self.top:   # the upper neighbour's bottom halo = my first `halo` owned rows
            self.ctx.history_pack(self.top, self.halo, self.send_up)
        if self.bot:   # the lower neighbour's top halo = my last `halo` owned rows
            self.ctx.history_pack(self.top + self.band.rows - self.halo, self.halo, self.send_dn)

    def unpack(self):
        if self.top:
            self.ctx.history_unpack(0, self.top, self.recv_up)
        if self.bot:
            self.ctx.history_unpack(self.top + self.band.rows, self.bot, self.recv_dn)

    def exchange_distributed(self):
        """Neighbour point-to-point swap over torch.distributed (NCCL over NVLink); no collective."""
        self.pack()
        neighbour_swap(self.band.rank, self.send_up, self.recv_up, self.send_dn, self.recv_dn)
        self.unpack()


class P2PLink:
    """NVLink peer-to-peer exchange for one BandedSvgf across processes (one per GPU).  Set-up (once): every
    rank allocates double-buffered receive buffers for both sides plus two flag words through the C ABI
    (rmd_p2p_alloc), publishes their CUDA IPC handles with one all_gather_object, and maps its neighbours'.
    Per frame: pack my boundary rows DIRECTLY into the neighbour's receive buffer, bump the neighbour's flag, wait on
    my own flags, unpack.  Everything is stream-ordered on the device; torch.distributed is not touched again."""

    def __init__(self, band: "BandedSvgf"):
        import ctypes
        from . import _lib
        self.b, self.lib, self.ct = band, _lib.load(), ctypes
        self.nbytes = band.ctx.history_bytes(band.halo)
        self.frame = 0
        # my receive side: [from_up parity0, from_up parity1, from_dn parity0, from_dn parity1] + 2 flags
        self.recv = self._alloc(4 * self.nbytes)
        self.flags = self._alloc(16)
        handles = (self._export(self.recv), self._export(self.flags))
        world = dist.get_world_size()
        table = [None] * world
        dist.all_gather_object(table, handles)
        r = band.band.rank
        self.up = self._open(table[r - 1]) if band.top else None
        self.dn = self._open(table[r + 1]) if band.bot else None
        dist.barrier()

    def _alloc(self, n):
        p = self.ct.c_void_p()
        rc = self.lib.rmd_p2p_alloc(self.ct.byref(p), n)
        if rc:
            raise RuntimeError(f"rmd_p2p_alloc -> {rc}")
        return p.value

    def _export(self, ptr):
        h = self.ct.create_string_buffer(64)
        rc = self.lib.rmd_p2p_export(self.ct.c_void_p(ptr), h)
        if rc:
            raise RuntimeError(f"rmd_p2p_export -> {rc}")
        return bytes(h.raw)

    def _open(self, handles):
        out = []
        for h in handles:
            p = self.ct.c_void_p()
            rc = self.lib.rmd_p2p_open(self.ct.create_string_buffer(h, 64), self.ct.byref(p))
            if rc:
                raise RuntimeError(f"rmd_p2p_open -> {rc}")
            out.append(p.value)
        return out  # [recv base, flags base] of the neighbour

    def exchange(self):
        b, nb = self.b, self.nbytes
        par = self.frame & 1
        self.frame += 1
        s = ctypes_stream()
        if b.top:   # my first owned rows -> the upper neighbour's "from below" buffer, then its flag[1]
            b.ctx.history_pack(b.top, b.halo, self.up[0] + (2 + par) * nb)
            self.lib.rmd_p2p_signal(self.ct.c_void_p(self.up[1] + 8), self.frame, s)
        if b.bot:   # my last owned rows -> the lower neighbour's "from above" buffer, then its flag[0]
            b.ctx.history_pack(b.top + b.band.rows - b.halo, b.halo, self.dn[0] + par * nb)
            self.lib.rmd_p2p_signal(self.ct.c_void_p(self.dn[1]), self.frame, s)
        from .api import RmdError
        if b.top:
            rc = self.lib.rmd_p2p_wait(self.ct.c_void_p(self.flags), self.frame, s)
            if rc:
                raise RmdError(rc)   # RMD_E_TIMEOUT: an earlier wait gave up, the halo history is not the neighbour's
            b.ctx.history_unpack(0, b.top, self.recv + par * nb)
        if b.bot:
            rc = self.lib.rmd_p2p_wait(self.ct.c_void_p(self.flags + 8), self.frame, s)
            if rc:
                raise RmdError(rc)
            b.ctx.history_unpack(b.top + b.band.rows, b.bot, self.recv + (2 + par) * nb)

    def timeouts(self):
        return self.lib.rmd_p2p_timeouts()

    def close(self):
        """Unmaps the neighbours' buffers and frees this rank's (after a barrier: nobody writes into them any more)."""
        torch.cuda.synchronize()
        for nb in (self.up, self.dn):
            if nb:
                for p in nb:
                    self.lib.rmd_p2p_close(self.ct.c_void_p(p))
        self.up = self.dn = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.barrier()
        for p in (self.recv, self.flags):
            if p:
                self.lib.rmd_p2p_free(self.ct.c_void_p(p))
        self.recv = self.flags = None


BAND_HALO = 40  # RMD_BAND_HALO
BAND_MAX_MOTION_Y = 13  # RMD_BAND_MAX_MOTION_Y: beyond it a band's reprojection differs from the single-GPU frame


class BandedSvgfV2:
    """One rank's band with per-level halo exchange (include/rmd_b200.h "Row bands with per-level halo exchange"):
    the context covers own rows + 40 halo rows per interior side; the a-trous levels produce only the own rows and
    push the boundary rows the neighbour's next level reads into the neighbour's receive buffer (NVLink peer stores
    between processes, plain device copies when several bands share one GPU in tests)."""

    def __init__(self, width, full_height, band: Band, device=0):
        import ctypes
        from . import _lib
        from .api import SvgfContext, RmdError
        self.ct, self.lib, self.band, self.W, self.H = ctypes, _lib.load(), band, width, full_height
        self.top = 0 if band.row0 == 0 else BAND_HALO
        self.bot = 0 if band.row0 + band.rows == full_height else BAND_HALO
        if band.rows < BAND_HALO:
            raise ValueError("band shorter than the halo")
        self.ext_row0 = band.row0 - self.top
        self.ext_rows = self.top + band.rows + self.bot
        self.ctx = SvgfContext(width, self.ext_rows, device)
        rc = self.lib.rmd_svgf_band_configure(self.ctx._h, self.top, band.rows)
        if rc:
            raise RmdError(rc)
        self.recv = self._alloc(self.lib.rmd_svgf_band_recv_bytes(self.ctx._h))
        self.flags = self._alloc(16)
        self.link = _lib.RmdBandLink()
        self.link.recv, self.link.flags = self.recv, self.flags
        self._mapped = []

    def _alloc(self, n):
        p = self.ct.c_void_p()
        rc = self.lib.rmd_p2p_alloc(self.ct.byref(p), n)
        if rc:
            raise RuntimeError(f"rmd_p2p_alloc -> {rc}")
        return p.value

    def connect_local(self, up, down):
        """Neighbours in the same process (single-GPU emulation)."""
        if up is not None:
            self.link.peer_recv[0], self.link.peer_flag[0] = up.recv, up.flags + 8   # its "from below" word
        if down is not None:
            self.link.peer_recv[1], self.link.peer_flag[1] = down.recv, down.flags   # its "from above" word

    def connect_ipc(self):
        """Neighbours in other processes (one per GPU): publish CUDA IPC handles once, map the neighbours'."""
        def export(ptr):
            h = self.ct.create_string_buffer(64)
            rc = self.lib.rmd_p2p_export(self.ct.c_void_p(ptr), h)
            if rc:
                raise RuntimeError(f"rmd_p2p_export -> {rc}")
            return bytes(h.raw)

        def open_(h):
            p = self.ct.c_void_p()
            rc = self.lib.rmd_p2p_open(self.ct.create_string_buffer(h, 64), self.ct.byref(p))
            if rc:
                raise RuntimeError(f"rmd_p2p_open -> {rc}")
            self._mapped.append(p.value)
            return p.value
        table = [None] * dist.get_world_size()
        dist.all_gather_object(table, (export(self.recv), export(self.flags)))
        r = self.band.rank
        if self.top:
            self.link.peer_recv[0], self.link.peer_flag[0] = open_(table[r - 1][0]), open_(table[r - 1][1]) + 8
        if self.bot:
            self.link.peer_recv[1], self.link.peer_flag[1] = open_(table[r + 1][0]), open_(table[r + 1][1])
        dist.barrier()

    def slice_rows(self, plane):
        return plane[self.ext_row0:self.ext_row0 + self.ext_rows]

    def owned(self, ext_plane):
        return ext_plane[self.top:self.top + self.band.rows]

    def stage(self, s, color, albedo, guide, motion, out, params, svgf=None, stream=None, out_rgba8=None):
        """One stage of a band frame (tests interleave the stages of several bands that share one GPU)."""
        from . import _lib
        from .api import RmdError, _ptr, _stream_ptr
        f = _lib.RmdSvgfFrame(self.W, self.ext_rows, _ptr(color), _ptr(albedo), _ptr(guide), _ptr(motion), _ptr(out),
                              _ptr(out_rgba8))
        sp = svgf.c() if svgf is not None else None
        rc = self.lib.rmd_svgf_band_stage(self.ctx._h, self.ct.byref(f), self.ct.byref(params.c()),
                                          self.ct.byref(sp) if sp is not None else None, self.ct.byref(self.link), s,
                                          _stream_ptr(stream))
        if rc:
            raise RmdError(rc)

    def frame(self, color, albedo, guide, motion, out, params, svgf=None, stream=None, out_rgba8=None):
        """All stages of one band frame in one call (rmd_svgf_band_frame).  Raises RmdError(RMD_E_TIMEOUT) once a
        neighbour's rows have failed to arrive."""
        from . import _lib
        from .api import RmdError, _ptr, _stream_ptr
        f = _lib.RmdSvgfFrame(self.W, self.ext_rows, _ptr(color), _ptr(albedo), _ptr(guide), _ptr(motion), _ptr(out),
                              _ptr(out_rgba8))
        sp = svgf.c() if svgf is not None else None
        rc = self.lib.rmd_svgf_band_frame(self.ctx._h, self.ct.byref(f), self.ct.byref(params.c()),
                                          self.ct.byref(sp) if sp is not None else None, self.ct.byref(self.link),
                                          _stream_ptr(stream))
        if rc:
            raise RmdError(rc)

    def timeouts(self):
        """Flag waits that gave up so far (host-visible word, no device synchronisation)."""
        return int(self.lib.rmd_svgf_band_timeouts(self.ctx._h))

    def launches_per_frame(self):
        return int(self.lib.rmd_svgf_band_launch_count(self.ctx._h))

    def close(self):
        """Unmaps the neighbours' buffers, then (after a barrier when running under torch.distributed, so that no
        neighbour still writes into them) frees this rank's receive buffer and flag words and the context."""
        if getattr(self, "ctx", None) is None:
            return
        torch.cuda.synchronize()
        for p in self._mapped:
            self.lib.rmd_p2p_close(self.ct.c_void_p(p))
        self._mapped = []
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.barrier()
        for p in (self.recv, self.flags):
            self.lib.rmd_p2p_free(self.ct.c_void_p(p))
        self.recv = self.flags = None
        self.ctx.close()
        self.ctx = None


def frame_in_process_v2(bands, band_planes, outs, params):
    """All bands in one process on one GPU: stage s of every band before stage s+1, so that every flag wait finds
    its signal already enqueued on the device."""
    for s in range(params.depth + 1):
        for b, pl, o in zip(bands, band_planes, outs):
            b.stage(s, *pl, o, params)


def ctypes_stream():
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def exchange_in_process(bands):
    """All bands in one process (single-GPU emulation of the N-rank path): same pack/unpack, copies instead of sends."""
    for b in bands:
        b.pack()
    for i, b in enumerate(bands):
        if b.top:
            b.recv_up.copy_(bands[i - 1].send_dn)
        if b.bot:
            b.recv_dn.copy_(bands[i + 1].send_up)
    for b in bands:
        b.unpack()


def max_over_ranks(value: float, device=None) -> float:
    """Multi-GPU timings are reported as the max over ranks (never wall clock of one rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_mpixels_per_s(pixels_this_rank: float, seconds_this_rank: float, device=None) -> float:
    """Whole-job throughput: all ranks' pixels over the slowest rank's time."""
    return sum_over_ranks(pixels_this_rank, device) / max_over_ranks(seconds_this_rank, device) / 1e6
