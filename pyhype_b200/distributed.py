"""Multi-GPU plumbing: one process per GPU.

The shipped transport lives in the C layer (``pyh_comm_init``, pyhype_b200/csrc/pyh_comm.cuh: NCCL over NVLink bound
with dlopen; the strip exchange and the dt all-reduce are part of the CUDA graph of one time step).  This module holds
the block -> rank rule, the helper that distributes the NCCL id, and a HOST-DRIVEN transport over ``torch.distributed``
(``HaloExchanger`` / ``advance``) that drives any engine with the C ABI's pack / unpack contract -- it is what the CPU
arm of bench.py and the gloo tests use with the oracle-backed stand-in engine.

Replaces the reference's mpi4py layer (pyhype/blocks/ghost.py:169-241 Isend/Irecv per ghost
strip, pyhype/blocks/base.py:454-465 Waitall, pyhype/solvers/base.py:128-131 gather+bcast of dt):

* blocks are dealt to ranks with the reference's contiguous rule
  (``Blocks.distribute_blocks_to_processes``, pyhype/blocks/base.py:473-513);
* per RK stage every rank packs the edge strips its remote neighbours need into one device
  buffer (``pyh_pack_halo``), exchanges them with grouped send/recv, and unpacks the received
  strips into its ghost frames (``pyh_unpack_halo``); same-rank edges never leave the device;
* the global CFL step is ``all_reduce(MIN)`` on one fp64 device scalar (min is exact, so the
  result does not depend on the rank count).

The sharded result is bit-identical to the single-GPU one: every residual reads only the
block's own cells and its ghost frame, which holds the same values either way.
"""
from __future__ import annotations

import os

OPPOSITE = {"E": "W", "W": "E", "N": "S", "S": "N"}


def distribute_blocks(num_blocks: int, num_processes: int) -> dict:
    """{block_num: rank}; the first ``num_blocks % num_processes`` ranks get one extra block
    (pyhype/blocks/base.py:473-513)."""
    owner = {}
    counter = 0
    full = num_blocks % num_processes
    base = num_blocks // num_processes
    for r in range(num_processes):
        n = base + 1 if r < full else base
        for g in range(counter, counter + n):
            owner[g] = r
        counter += n
    return owner


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process per GPU)."""
    return (
        int(os.environ.get("RANK", "0")),
        int(os.environ.get("WORLD_SIZE", "1")),
        int(os.environ.get("LOCAL_RANK", "0")),
    )


def init_nccl(device_index):
    """``init_process_group("nccl")`` with a high-priority NCCL stream: the strip exchange runs while
    the stage kernel owns every register file, so its small transfer kernels must win the first free
    thread-block slot instead of queueing behind thousands of pending stage thread blocks."""
    import torch
    import torch.distributed as dist

    opts = None
    try:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
    except Exception:  # older / differently built torch: plain group
        opts = None
    kw = dict(device_id=torch.device("cuda", device_index))
    if opts is not None:
        kw["pg_options"] = opts
    dist.init_process_group("nccl", **kw)


def share_unique_id(make_id, rank, world):
    """Hand rank 0's 128-byte NCCL id to every rank.  The side channel is ``torch.distributed`` (already initialised by the
    caller, or a CPU-only gloo group created here from the torchrun environment) -- control plane only: the strips and the
    dt reduction travel through the NCCL communicator the C layer owns (``pyh_comm_init``)."""
    import torch.distributed as dist

    if not dist.is_initialized():
        dist.init_process_group("gloo", rank=rank, world_size=world)
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def exchange_plan(slots, owner, rank):
    """Order the point-to-point messages of one stage.

    ``slots``: this rank's halo slots from ``Engine.halo_slots()``.  Returns two lists of
    (peer, key, offset, length): sends and receives, each sorted by (peer, key) where the key
    (source block, source side) names a message uniquely on both ends, so the k-th send of
    rank a to rank b meets the k-th receive b posts for a.
    """
    sides = ("E", "W", "N", "S")
    sends, recvs = [], []
    for s in slots:
        peer = owner[s["nbr"]]
        if peer == rank:
            raise ValueError("halo slot whose neighbour is local")
        sends.append((peer, (s["gid"], sides.index(s["side"])), s["offset"], s["length"]))
        recvs.append((peer, (s["nbr"], sides.index(OPPOSITE[s["side"]])), s["offset"], s["length"]))
    sends.sort(key=lambda m: (m[0], m[1]))
    recvs.sort(key=lambda m: (m[0], m[1]))
    return sends, recvs


class HaloExchanger:
    """Owns the send/recv device buffers of one engine and runs the per-stage exchange."""

    def __init__(self, engine, owner, rank, group=None, backend_device=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.engine = engine
        self.group = group
        slots, ndoubles = engine.halo_slots()
        self.sends, self.recvs = exchange_plan(slots, owner, rank)
        dev = backend_device if backend_device is not None else torch.device("cuda", engine.device)
        self.sendbuf = torch.zeros(max(ndoubles, 1), dtype=torch.float64, device=dev)
        self.recvbuf = torch.zeros(max(ndoubles, 1), dtype=torch.float64, device=dev)
        self.empty = ndoubles == 0
        self.dt = torch.zeros(1, dtype=torch.float64, device=dev)
        self.comm = None
        self.pending = None
        self.overlap = False
        # opt-in (PYH_HALO_OVERLAP=1): measured on 2 x B200 the blocking exchange costs ~35 us per stage while the
        # dispatch-ordered launch it needs runs 0.2 ms slower there (profiles/r01k_halo_overlap.md), so it does not pay yet
        if dev.type == "cuda" and not self.empty and os.environ.get("PYH_HALO_OVERLAP", "0") == "1":
            capable, n_remote = engine.overlap_info()
            # thread blocks that may wait for the exchange must never be able to fill the device alone
            self.overlap = capable and n_remote <= 2 * torch.cuda.get_device_properties(dev).multi_processor_count

    def exchange(self):
        """pack -> grouped isend/irecv -> unpack (all on the current torch stream)."""
        if self.empty:
            return
        self.wait()
        dist = self.dist
        self.engine.pack_halo(self.sendbuf.data_ptr())
        ops = []
        for peer, _k, off, ln in self.recvs:
            ops.append(dist.P2POp(dist.irecv, self.recvbuf[off:off + ln], peer, group=self.group))
        for peer, _k, off, ln in self.sends:
            ops.append(dist.P2POp(dist.isend, self.sendbuf[off:off + ln], peer, group=self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        self.engine.unpack_halo(self.recvbuf.data_ptr())

    # -- overlapped variant (CUDA only) ------------------------------------------------------------------
    def exchange_async(self):
        """pack on the compute stream, then send/recv + unpack on a communication stream; returns
        (and remembers in ``self.pending``) the event after which the remote ghost frames are valid.
        The next ``advance`` hands it to ``Engine.stage_overlapped`` so that only the thread blocks
        next to remote edges wait for it."""
        if self.empty:
            return None
        torch, dist = self.torch, self.dist
        if self.comm is None:
            self.comm = torch.cuda.Stream(device=self.sendbuf.device, priority=-1)
            self._ev_packed = torch.cuda.Event()
            self._ev_ready = torch.cuda.Event()
        main = torch.cuda.current_stream(self.sendbuf.device)
        self.engine.pack_halo(self.sendbuf.data_ptr())
        self._ev_packed.record(main)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self._ev_packed)
            ops = []
            for peer, _k, off, ln in self.recvs:
                ops.append(dist.P2POp(dist.irecv, self.recvbuf[off:off + ln], peer, group=self.group))
            for peer, _k, off, ln in self.sends:
                ops.append(dist.P2POp(dist.isend, self.sendbuf[off:off + ln], peer, group=self.group))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            self.engine.unpack_halo_on(self.recvbuf.data_ptr(), self.comm.cuda_stream)
            self._ev_ready.record(self.comm)
        self.pending = self._ev_ready
        return self.pending

    def wait(self):
        """Make the compute stream wait for an outstanding asynchronous exchange (before anything
        other than ``stage_overlapped`` reads remote ghost cells)."""
        if self.pending is not None:
            self.torch.cuda.current_stream(self.sendbuf.device).wait_event(self.pending)
            self.pending = None

    def global_dt(self):
        """CFL * global min, left on the device (solvers/base.py:126-131)."""
        self.engine.local_dt(self.dt.data_ptr())
        self.dist.all_reduce(self.dt, op=self.dist.ReduceOp.MIN, group=self.group)
        return self.dt


def advance(engine, halo, num_stages, dt=None, dt_dev_ptr=None):
    """One time step of ExplicitRungeKutta.integrate (explicit_runge_kutta.py:63-80) on a sharded
    domain: per stage the fused stage kernel, then the remote strip exchange, then the local ghost
    copies + BC functors (Blocks.apply_boundary_condition, blocks/base.py:448-471)."""
    if dt_dev_ptr is not None:
        engine.step_begin_dev(dt_dev_ptr)
    else:
        engine.step_begin(dt)
    overlap = halo is not None and getattr(halo, "overlap", False)
    for s in range(num_stages):
        if overlap:
            # remote strips of this stage's input may still be in flight: only the thread blocks that
            # read them wait; the exchange of the stage's output then runs behind the next stage
            if halo.pending is not None:
                engine.stage_overlapped(s)
            else:
                engine.stage(s)
            halo.exchange_async()
        else:
            engine.stage(s)
            if halo is not None:
                halo.exchange()
        engine.apply_bc()
