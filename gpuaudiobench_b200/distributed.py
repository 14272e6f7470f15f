"""Track sharding and the stereo mix-bus reduce — the only multi-GPU logic the path has.

Tracks are independent units (own input, IR, history / delay line), so GPU g of G owns the
contiguous range [g*Tg/G, (g+1)*Tg/G) and no data-path collective is needed except the sum of the
per-GPU stereo mix bus float[2][B].  One process per GPU; torch.distributed
supplies the plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_tracks(total_tracks, world_size, rank):
    """Contiguous track range [t0, t1) of `rank`; ranges tile [0, total_tracks) exactly."""
    t0 = total_tracks * rank // world_size
    t1 = total_tracks * (rank + 1) // world_size
    return t0, t1


def default_mix_gains(total_tracks, t0, t1):
    """Constant-power pan from the GLOBAL track index, bus scale 1/sqrt(Tg) (same formula as the
    engine's default, csrc/engine.cu set_default_gains)."""
    t = torch.arange(t0, t1, dtype=torch.float64)
    theta = (t + 0.5) / total_tracks * (torch.pi / 2)
    scale = 1.0 / (total_tracks ** 0.5)
    return torch.stack([torch.cos(theta) * scale, torch.sin(theta) * scale], dim=1).to(torch.float32)


def reduce_mix_bus(mix, group=None, all_ranks=True):
    """Sum the per-GPU bus partials [2][B] in place.  all_ranks: all-reduce (every rank gets the
    bus), else reduce to rank 0.  4 KiB at B = 512: latency-bound, enqueued on the current stream."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return mix
    if all_ranks:
        dist.all_reduce(mix, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(mix, dst=0, op=dist.ReduceOp.SUM, group=group)
    return mix


def stitch_sample_major(column_tiles, total_tracks):
    """Host stitch of per-rank [B][Tg] column tiles (each rank wrote only its own columns)."""
    out = torch.zeros_like(column_tiles[0])
    for tile in column_tiles:
        out += tile
    assert out.shape[1] == total_tracks
    return out


class EngineBusGroup:
    """One engine per rank, each owning a contiguous track range of the same job: after construction the
    mix bus every engine's process call delivers is already the sum over ALL ranks — the exchange runs
    inside the engine's last convolution kernel over NVLink peer memory (b200conv_attach_bus,
    csrc/bus_tree.cuh).  torch.distributed._symmetric_memory supplies the peer mappings.  If symmetric
    memory cannot be set up (gloo in the CPU tests, no P2P) the engine stays stand-alone and `reduce()`
    falls back to the NCCL / gloo all-reduce of the local bus; `.kind` says which."""

    def __init__(self, engine, bus, group=None, force_nccl=False):
        self.engine, self.bus, self.group = engine, bus, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.kind = "none (1 rank)" if self.world == 1 else "NCCL all-reduce after the engine's launch"
        self.in_kernel = False
        if self.world == 1 or force_nccl or not bus.is_cuda:
            return
        try:
            import torch.distributed._symmetric_memory as symm_mem

            from . import engine as eng_mod
            nbytes = eng_mod.bus_buffer_bytes(self.world, bus.numel())
            self.buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=bus.device)
            self.buf.zero_()
            grp = group if group is not None else dist.group.WORLD
            self.hdl = symm_mem.rendezvous(self.buf, grp.group_name if hasattr(grp, "group_name") else grp)
            torch.cuda.synchronize(bus.device)
            dist.barrier(group)  # every rank's buffer is zeroed before the first push
            engine.attach_bus([int(p) for p in self.hdl.buffer_ptrs], self.rank, self.world)
            self.in_kernel = True
            self.kind = ("in-kernel: the engine's last convolution kernel pushes the bus over NVLink symmetric memory "
                         "and sums in rank order (b200conv_attach_bus)")
        except Exception as exc:  # pragma: no cover - depends on the box
            self.kind = f"NCCL all-reduce after the engine's launch (symmetric memory unavailable: {type(exc).__name__}: {exc})"

    def reduce(self):
        """Call after engine.process on the same stream: a no-op when the exchange ran inside the kernel."""
        if self.world > 1 and not self.in_kernel:
            dist.all_reduce(self.bus, op=dist.ReduceOp.SUM, group=self.group)
        return self.bus

    def check(self):
        if self.in_kernel:
            self.engine.bus_status()

    def close(self):
        if self.in_kernel:
            torch.cuda.synchronize(self.bus.device)
            dist.barrier(self.group)  # nobody unmaps a buffer a peer may still push into
            self.engine.attach_bus(None, 0, 1)
            self.in_kernel = False


class BusAllReduce:
    """All-reduce of the stereo bus [2][B] across the ranks of `group`.

    (The engines do this inside their last kernel — see EngineBusGroup; this class reduces a bus the
    caller owns, and is what the stand-alone kernel's tests drive.)
    Preferred path: the engine's own one-shot kernel (csrc/bus_allreduce.cu) over a symmetric
    peer-mapped buffer obtained from torch.distributed._symmetric_memory — P2P stores + flags over
    NVLink, one launch, fixed summation order.  If symmetric memory cannot be set up (e.g. gloo in
    the CPU tests, no P2P), it falls back to the NCCL/gloo all-reduce and says so in `.kind`.
    """

    def __init__(self, bus, group=None, force_nccl=False):
        import ctypes as C

        self.bus = bus
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.kind = "none (1 rank)" if self.world == 1 else "NCCL all-reduce"
        self.epoch = 0
        self._C = C
        if self.world == 1 or force_nccl or not bus.is_cuda:
            return
        try:
            import torch.distributed._symmetric_memory as symm_mem

            from . import engine
            lib = engine.load_library()
            n = bus.numel()
            nbytes = lib.b200conv_bus_buffer_bytes(self.world, n)
            self.buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=bus.device)
            self.buf.zero_()
            grp = group if group is not None else dist.group.WORLD
            self.hdl = symm_mem.rendezvous(self.buf, grp.group_name if hasattr(grp, "group_name") else grp)
            ptrs = [int(p) for p in self.hdl.buffer_ptrs]
            self.ptrs = (C.c_uint64 * self.world)(*ptrs)
            self.err = torch.zeros(1, dtype=torch.int32, device=bus.device)
            self.lib = lib
            torch.cuda.synchronize(bus.device)
            dist.barrier(group)  # every rank's buffer is zeroed before the first push
            self.kind = "own one-shot P2P kernel over NVLink symmetric memory (b200conv_bus_allreduce)"
        except Exception as exc:  # pragma: no cover - depends on the box
            self.kind = f"NCCL all-reduce (symmetric memory unavailable: {type(exc).__name__}: {exc})"

    def __call__(self, stream=None):
        """Reduce self.bus in place on the current (or given) CUDA stream."""
        if self.world == 1:
            return self.bus
        if not self.kind.startswith("own"):
            dist.all_reduce(self.bus, op=dist.ReduceOp.SUM, group=self.group)
            return self.bus
        C = self._C
        self.epoch += 1
        st = stream if stream is not None else torch.cuda.current_stream(self.bus.device).cuda_stream
        rc = self.lib.b200conv_bus_allreduce(C.c_void_p(self.bus.data_ptr()), C.c_void_p(self.bus.data_ptr()), self.ptrs,
                                             self.rank, self.world, self.bus.numel(), self.epoch,
                                             C.c_void_p(self.err.data_ptr()), C.c_void_p(st))
        if rc != 0:
            raise RuntimeError(f"b200conv_bus_allreduce failed: {rc}")
        return self.bus

    def check(self):
        if self.world > 1 and self.kind.startswith("own") and int(self.err.item()) != 0:
            raise RuntimeError("bus all-reduce: a peer did not signal within the spin bound")
