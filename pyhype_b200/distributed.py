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

    def exchange(self):
        """pack -> grouped isend/irecv -> unpack (all on the current torch stream)."""
        if self.empty:
            return
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
    for s in range(num_stages):
        engine.stage(s)
        if halo is not None:
            halo.exchange()
        engine.apply_bc()
