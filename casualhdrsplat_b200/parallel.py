"""Data-parallel training step of the formation path.

Work shards by training frames: Gaussians, CRF parameters and spline knots are replicated on every
GPU; the batch of B frames is split B/G per GPU with all n virtual poses of a frame on one GPU, so
the blur average / CRF epilogue needs no communication.  The one exchange step is a sum all-reduce
of the flat gradient buffer

    [ means 3N | quats 4N | scales 3N | opacities N | colors 3N | knots 7K | crf P | exposure B | frame_times B ]

over NCCL (NVLink 5 / NVSwitch).  One process per GPU; ``torch.distributed`` is used for
rendezvous only (or as the collective backend when asked), the data path can go through the C ABI
(``chs_allreduce_grads``).
"""
from __future__ import annotations

import ctypes
from typing import Callable, Dict, Optional, Sequence

import torch

from . import _lib
from .api import BufferPool, _stream, backward_stages, forward_stages

_SORT = {"key64": _lib.CHS_SORT_KEY64, "presort": _lib.CHS_SORT_DEPTH_PRESORT}


def shard_frames(n_frames: int, rank: int, world: int) -> range:
    """Contiguous block of frame indices owned by ``rank`` (blocks differ by at most one frame)."""
    if not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} for world {world}")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


class GradLayout:
    """Offsets of every section of the flat gradient buffer (in floats)."""

    def __init__(self, n_gauss: int, n_knots: int, n_crf: int, n_frames: int):
        self.N, self.K, self.P, self.B = n_gauss, n_knots, n_crf, n_frames
        self.off_knots = 14 * n_gauss
        self.off_crf = self.off_knots + 7 * n_knots
        self.off_exposure = self.off_crf + n_crf
        self.off_frame_times = self.off_exposure + n_frames
        self.total = self.off_frame_times + n_frames

    def views(self, buf: torch.Tensor) -> Dict[str, torch.Tensor]:
        N = self.N
        return {
            "means": buf[0:3 * N].view(N, 3), "quats": buf[3 * N:7 * N].view(N, 4), "scales": buf[7 * N:10 * N].view(N, 3),
            "opacities": buf[10 * N:11 * N], "colors": buf[11 * N:14 * N].view(N, 3),
            "knots": buf[self.off_knots:self.off_crf].view(self.K, 7), "crf_params": buf[self.off_crf:self.off_exposure],
            "exposure_times": buf[self.off_exposure:self.off_frame_times], "frame_times": buf[self.off_frame_times:self.total],
        }


class ChsComm:
    """NCCL communicator owned by libchs (C ABI). The 128-byte unique id travels over torch.distributed."""

    def __init__(self, rank: int, world: int, device: torch.device):
        import torch.distributed as dist

        L = _lib.lib()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_char * 128)()
            _lib.check(L.chs_comm_unique_id(buf), "chs_comm_unique_id")
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.to(device)
        dist.broadcast(uid, src=0)
        raw = bytes(uid.cpu().tolist())
        self._handle = ctypes.c_void_p()
        self._raw = ctypes.create_string_buffer(raw, 128)
        _lib.check(L.chs_comm_init(self._raw, rank, world, ctypes.byref(self._handle)), "chs_comm_init")
        self.rank, self.world = rank, world

    def allreduce_(self, buf: torch.Tensor) -> None:
        _lib.check(_lib.lib().chs_allreduce_grads(self._handle, _lib.ptr(buf), buf.numel(), _stream()), "chs_allreduce_grads")

    def close(self) -> None:
        if self._handle:
            _lib.lib().chs_comm_destroy(self._handle)
            self._handle = ctypes.c_void_p()


class NvlsComm:
    """Hand-written one-shot collectives through the NVSwitch multicast mapping (NVLS), no NCCL in the data path.

    The flat gradient buffer itself lives in symmetric memory (torch.distributed._symmetric_memory owns the allocation
    and the rendezvous — plumbing), so K9 / the accumulations write straight into it; ``allreduce_`` is then
    barrier -> ``chs_nvls_allreduce`` (multimem.ld_reduce + multimem.st on this rank's slice) -> barrier.
    Every rank ends with bit-identical sums.  ``symmetric(name, n)`` hands out further symmetric buffers (e.g. the replicated
    parameters) and ``broadcast_slice_`` fans this rank's slice of one out to every rank (``chs_nvls_broadcast``).
    Raises if the fabric has no multicast support.
    """

    def __init__(self, rank: int, world: int, device: torch.device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self._symm, self._dist = symm, dist
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world, self.device = rank, world, device
        self._bufs = {}  # name -> (tensor, handle)

    def symmetric(self, name: str, n_floats: int) -> torch.Tensor:
        """A symmetric [n_floats] fp32 buffer with a multicast mapping (allocated and rendezvoused once per name)."""
        cur = self._bufs.get(name)
        if cur is None or cur[0].numel() < n_floats:
            padded = (n_floats + 1023) // 1024 * 1024
            buf = self._symm.empty(padded, dtype=torch.float32, device=self.device)
            hdl = self._symm.rendezvous(buf, self.group.group_name)
            if not hdl.multicast_ptr:
                raise RuntimeError("NvlsComm: this fabric/driver exposes no multicast (NVLS) mapping; use ChsComm (NCCL)")
            self._bufs[name] = cur = (buf, hdl)
        return cur[0][:n_floats]

    def flat_buffer(self, n_floats: int) -> torch.Tensor:
        """The symmetric gradient buffer (reused every step)."""
        return self.symmetric("grads", n_floats)

    def _find(self, buf: torch.Tensor):
        for b, h in self._bufs.values():
            if b.data_ptr() == buf.data_ptr():
                return b, h
        raise RuntimeError("NvlsComm: the buffer must come from flat_buffer() / symmetric()")

    def allreduce_(self, buf: torch.Tensor, begin: int = 0, count: Optional[int] = None, channel: int = 0) -> None:
        """Sum floats [begin, begin + count) of the symmetric buffer over all ranks (default: all of ``buf``)."""
        base, hdl = self._find(buf)
        count = buf.numel() - begin if count is None else count
        if begin % 4:
            raise RuntimeError("NvlsComm.allreduce_: begin must be a multiple of 4 floats")
        hdl.barrier(channel=channel)  # every rank's partial sums are in its symmetric buffer
        _lib.check(_lib.lib().chs_nvls_allreduce(ctypes.c_void_p(hdl.multicast_ptr + 4 * begin), count, self.rank, self.world, _stream()),
                   "chs_nvls_allreduce")
        hdl.barrier(channel=channel + 1)  # every slice has been broadcast

    def broadcast_slice_(self, buf: torch.Tensor, begin: int, count: int, channel: int = 2) -> None:
        """Every rank calls this with ITS slice [begin, begin + count): afterwards all ranks hold all slices."""
        base, hdl = self._find(buf)
        _lib.check(_lib.lib().chs_nvls_broadcast(ctypes.c_void_p(hdl.multicast_ptr), _lib.ptr(base), begin, count, _stream()),
                   "chs_nvls_broadcast")
        hdl.barrier(channel=channel)  # every rank's slice has landed everywhere

    def close(self) -> None:
        self._bufs = {}


class TorchComm:
    """Collective through torch.distributed (NCCL on GPUs; gloo in the CPU tests of the sharding logic)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self._dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def allreduce_(self, buf: torch.Tensor) -> None:
        self._dist.all_reduce(buf, op=self._dist.ReduceOp.SUM, group=self.group)

    def close(self) -> None:
        pass


def formation_step(params: Dict[str, torch.Tensor], spline_meta: dict, width: int, height: int, n_virtual: int, crf_kind: int,
                   frame_ids: Sequence[int], upstream: Callable[[Sequence[int], torch.Tensor], torch.Tensor], *,
                   micro_batch: int = 1, sort_mode: str = "presort", comm=None, background=None, out: Optional[torch.Tensor] = None,
                   stats: Optional[dict] = None, crf_before_average: bool = False, tight_bounds: bool = False,
                   pose_fused: bool = False, tuning: Optional[dict] = None, state: Optional["StepState"] = None):
    """One fwd+bwd training step over this rank's frames, then the gradient all-reduce.

    params: CUDA fp32 tensors means [N,3], quats [N,4], scales [N,3], opacities [N], colors [N,3], knots [K,7],
            frame_times [B], exposure_times [B], Ks [B,3,3], crf_params [3,3Hd+1] (or absent for the identity CRF);
            B is the GLOBAL batch.  spline_meta: knot_t0, knot_dt, kind.
    tight_bounds: opacity-aware tile bounds (``rasterize``): same gradients, shorter tile lists.
    pose_fused / tuning: as in ``rasterize`` (chs_config.pose_fused, chs_config.tune_*).
    state: a ``StepState`` kept by the caller across steps.  With it the step never synchronises with the host (after the
            first step has learnt every frame's intersection count M, the buffers are sized from the previous step's M plus a
            margin and K2's count stays on the device) and every stage buffer comes from a pool instead of the allocator.
            Call ``state.verify()`` before trusting a step's gradients: it raises if a frame outgrew its buffers.
    frame_ids: the frames this rank renders (``shard_frames``).  ``upstream(ids, ldr[len(ids),H,W,3])`` returns the
            gradient of the loss w.r.t. those LDR frames (same shape).
    Returns (GradLayout, flat gradient buffer summed over all ranks).
    """
    means, quats, scales = params["means"], params["quats"], params["scales"]
    opacities, colors, knots = params["opacities"], params["colors"], params["knots"]
    crf_params = params.get("crf_params") if crf_kind != _lib.CHS_CRF_IDENTITY else None
    N, K = means.shape[0], knots.shape[0]
    B_total = params["frame_times"].shape[0]
    n_crf = crf_params.numel() if crf_params is not None else 0
    layout = GradLayout(N, K, n_crf, B_total)
    dev = means.device
    if out is None and hasattr(comm, "flat_buffer"):
        out = comm.flat_buffer(layout.total)
    flat = out if out is not None else torch.empty(layout.total, dtype=torch.float32, device=dev)
    flat[layout.off_knots:].zero_()
    v = layout.views(flat)
    first = True
    ids_all = list(frame_ids)
    n_isect_total, m_g_total = 0, 0
    capturing = torch.cuda.is_current_stream_capturing()
    if capturing and state is None:
        raise RuntimeError("formation_step: CUDA-graph capture needs a warmed-up StepState (see GraphedStep)")
    for s in range(0, len(ids_all), micro_batch):
        ids = ids_all[s:s + micro_batch]
        idx = state.index(tuple(ids), dev) if state is not None else torch.as_tensor(ids, device=dev)
        ft, ex, Ks = params["frame_times"][idx].contiguous(), params["exposure_times"][idx].contiguous(), params["Ks"][idx].contiguous()
        cfg = _lib.make_config(N, len(ids), n_virtual, width, height, crf_kind=crf_kind,
                               crf_hidden=_lib.crf_size(crf_kind, crf_params),
                               sort_mode=_SORT[sort_mode], background=background, crf_before_average=crf_before_average,
                               tight_bounds=tight_bounds, pose_fused=pose_fused, tuning=tuning)
        spline = (knots, float(spline_meta["knot_t0"]), float(spline_meta["knot_dt"]), ft, int(spline_meta["kind"]))
        cap = state.capacity(tuple(ids)) if state is not None else None
        if capturing and cap is None:
            raise RuntimeError("formation_step: capture before the StepState has learnt the intersection counts")
        # (the step that learns M sizes its buffers exactly, per frame batch: those do not go into the pool)
        pool = state.pool if (state is not None and cap is not None) else None
        st = forward_stages(means, quats, scales, opacities, colors, None, Ks, ex, crf_params, cfg, spline, isect_capacity=cap, pool=pool,
                            pinned_name=f"n_isect{tuple(ids)}" if capturing else None)
        if capturing:
            state.graph_counts.append((tuple(ids), st._n_pinned, cap))  # checked by GraphedStep.verify after a replay
        elif state is not None:
            state.track(tuple(ids), st)
        v_ldr = upstream(ids, st.ldr).contiguous()
        # the first micro-batch writes K9's output straight into the flat buffer (which may be symmetric memory)
        g = backward_stages(st, means, quats, scales, ex, crf_params, v_ldr, None, grads_out=flat[:14 * N] if first else None, pool=pool)
        if first:
            first = False
        else:
            flat[:14 * N].add_(g["grads_flat"])
        v["knots"].add_(g["v_knots"])
        if crf_params is not None:
            v["crf_params"].add_(g["v_crf"].reshape(-1))
        v["exposure_times"][idx] = g["v_exposure"]
        v["frame_times"][idx] = g["v_frame_times"]
        if state is None or cap is None:
            n_isect_total += st.n_isect  # (in the sync-free mode reading M would wait for the GPU: StepState reports it one step late)
        if stats is not None and stats.get("count_pairs"):
            m_g_total += int((st.tiles_touched > 0).sum())
    if first:
        flat[:14 * N].zero_()
    if comm is not None and comm.world > 1:
        if stats is not None and "events" in stats:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            comm.allreduce_(flat)
            e1.record()
            stats["events"].append((e0, e1))
        else:
            comm.allreduce_(flat)
    if stats is not None:
        stats["n_isect"] = n_isect_total if (state is None or n_isect_total > 0) else state.last_total()
        if stats.get("count_pairs"):
            stats["m_g"] = m_g_total
    return layout, flat


class ShardedHostParams:
    """Host-resident parameters of a frame-sharded job, replicated on the GPUs without every rank uploading everything.

    All parameter tensors live in ONE flat fp32 device buffer (views per tensor).  Each step rank r copies only slice r (1/G of
    the buffer) from its pinned host copy and the slices are exchanged over NVLink: ``chs_nvls_broadcast`` (multimem.st through
    the NVSwitch) when the communicator is an ``NvlsComm``, else ``all_gather_into_tensor``.  Host traffic per step is the
    parameter bytes once in total instead of once per rank.  The reduced gradient buffer is read back the same way: rank r
    downloads slice r (``grad_slice``).
    """

    def __init__(self, host: Dict[str, torch.Tensor], device: torch.device, rank: int, world: int, comm=None, n_slots: int = 2, group=None):
        self.names = list(host)
        self.shapes = {k: tuple(v.shape) for k, v in host.items()}
        sizes = [int(v.numel()) for v in host.values()]
        self.offsets, off = {}, 0
        for k, n in zip(self.names, sizes):
            self.offsets[k] = off
            off += (n + 3) // 4 * 4  # every tensor starts 16-byte aligned
        self.rank, self.world, self.comm, self.group = rank, world, comm, group
        per = (off + world - 1) // world
        self.per = (per + 3) // 4 * 4
        self.total = self.per * world
        self.host_flat = torch.zeros(self.total, dtype=torch.float32)
        if torch.cuda.is_available():
            self.host_flat = self.host_flat.pin_memory()
        for k, v in host.items():
            self.host_flat[self.offsets[k]:self.offsets[k] + v.numel()].copy_(v.reshape(-1))
        self.nvls = isinstance(comm, NvlsComm) and world > 1
        self.dev_flat = []
        for s in range(n_slots):
            if self.nvls:
                self.dev_flat.append(comm.symmetric(f"params{s}", self.total))
            else:
                self.dev_flat.append(torch.empty(self.total, dtype=torch.float32, device=device))
        self.h2d_bytes = self.per * 4 if world > 1 else off * 4

    def views(self, slot: int) -> Dict[str, torch.Tensor]:
        import math

        f = self.dev_flat[slot]
        return {k: f[self.offsets[k]:self.offsets[k] + math.prod(self.shapes[k])].view(self.shapes[k]) for k in self.names}

    def upload_(self, slot: int) -> None:
        """Stream-ordered on the current stream: H2D of this rank's slice, then the slice exchange."""
        f = self.dev_flat[slot]
        b = self.rank * self.per
        if self.world == 1:
            f.copy_(self.host_flat, non_blocking=True)
            return
        f[b:b + self.per].copy_(self.host_flat[b:b + self.per], non_blocking=True)
        if self.nvls:
            self.comm.broadcast_slice_(f, b, self.per, channel=2 + slot)
        else:
            import torch.distributed as dist

            dist.all_gather_into_tensor(f, f[b:b + self.per].clone(), group=self.group)

    def grad_slice(self, n_floats: int):
        """[begin, end) of the flat gradient buffer this rank reads back to its host."""
        per = ((n_floats + self.world - 1) // self.world + 3) // 4 * 4
        b = min(self.rank * per, n_floats)
        return b, min(b + per, n_floats)


class StepState:
    """What ``formation_step`` keeps between steps to run without host synchronisation: the stage-buffer pool, every frame
    batch's intersection capacity (the previous M plus ``margin``) and the forward states whose M has not been checked yet."""

    def __init__(self, margin: float = 0.03, slack: int = 65536):
        self.pool = BufferPool()
        self.margin, self.slack = margin, slack
        self._cap = {}       # frame ids -> capacity for the next step (None until M is known)
        self._last = {}      # frame ids -> last resolved M
        self._pending = []   # (frame ids, forward state) of steps not verified yet
        self.overflows = []  # (frame ids, M, capacity) of frames that outgrew their buffers
        self._idx = {}       # frame ids -> device index tensor (made once: a capture cannot copy from pageable memory)
        self.graph_counts = []  # (frame ids, pinned M, capacity) of a captured step

    def index(self, ids, device):
        t = self._idx.get(ids)
        if t is None:
            t = self._idx[ids] = torch.as_tensor(list(ids), device=device)
        return t

    def capacity(self, ids):
        if not torch.cuda.is_current_stream_capturing():  # (a capture may not query events; GraphedStep resolves beforehand)
            self._resolve(keep_last=True)
        if self._cap.get(ids) is None:
            return None
        # one capacity for all frame batches (rounded up to 1 Mi entries), so that the pooled buffers have one shape
        return (max(c for c in self._cap.values() if c) + (1 << 20) - 1) >> 20 << 20

    def track(self, ids, st) -> None:
        self._pending.append((ids, st))

    def _resolve(self, keep_last: bool) -> None:
        # states of EARLIER steps only: their K2 finished long ago, so reading M does not stall the stream being fed now
        done, keep = self._pending, []
        if keep_last:
            seen = set()
            for item in reversed(self._pending):  # the newest state of each frame batch belongs to the step in flight
                if item[0] not in seen and item[1]._n_isect is None and not item[1]._n_event.query():
                    keep.append(item)
                    seen.add(item[0])
            done = [it for it in self._pending if it not in keep]
        for ids, st in done:
            m = st.n_isect
            self._last[ids] = m
            if st.overflowed:
                self.overflows.append((ids, m, st.isect_capacity))
            want = int(m * (1.0 + self.margin)) + self.slack
            self._cap[ids] = max(want, self._cap.get(ids) or 0) if not st.overflowed else int(m * (1.0 + 2 * self.margin)) + self.slack
        self._pending = keep

    def verify(self) -> None:
        """Wait for every tracked step's intersection count and raise if one exceeded its capacity (that step's lists were
        truncated, so its gradients must be recomputed; the capacity has been raised for the retry)."""
        self._resolve(keep_last=False)
        if self.overflows:
            o, self.overflows = self.overflows, []
            raise RuntimeError(f"formation_step: intersection buffers overflowed for {o}; re-run the step")

    def last_total(self) -> int:
        return int(sum(self._last.values()))


class GraphedStep:
    """``formation_step`` of one rank captured in a CUDA graph and replayed (single process; the all-reduce stays outside).

    The sync-free step launches ~70 kernels per frame from Python; for small scenes (BASELINE configs[1]: 100k Gaussians, 800 x 800,
    4 poses) the host cannot feed the GPU fast enough.  Construction runs ``warmup`` eager steps (they learn every frame's
    intersection count and fill the buffer pool), then captures one more step; ``replay()`` re-runs it with whatever the
    parameter tensors hold NOW (same storage, same shapes).  ``upstream`` must be capturable (device work only, no host reads).
    ``verify()`` waits for the replay and raises if a frame needed more intersections than the captured capacity.
    """

    def __init__(self, params, spline_meta, width, height, n_virtual, crf_kind, frame_ids, upstream, warmup: int = 2, **kw):
        if kw.get("comm") is not None:
            raise RuntimeError("GraphedStep captures the local step only; all-reduce the returned buffer after replay()")
        self.state = kw.pop("state", None) or StepState()
        args = (params, spline_meta, width, height, n_virtual, crf_kind, frame_ids, upstream)
        for _ in range(max(warmup, 2)):
            formation_step(*args, state=self.state, **kw)
        self.state.verify()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # one more eager step on the capture stream: every lazy initialisation happens outside the graph
            formation_step(*args, state=self.state, **kw)
        torch.cuda.current_stream().wait_stream(side)
        self.state.verify()
        self.state.graph_counts = []
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.layout, self.flat = formation_step(*args, state=self.state, **kw)

    def replay(self):
        self.graph.replay()
        return self.layout, self.flat

    def verify(self) -> None:
        torch.cuda.current_stream().synchronize()
        bad = [(ids, int(p.item()), cap) for ids, p, cap in self.state.graph_counts if int(p.item()) > cap]
        if bad:
            raise RuntimeError(f"GraphedStep: intersection buffers overflowed for {bad}; rebuild the graph")
