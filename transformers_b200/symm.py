"""Peer-mapped (symmetric) buffers for the tensor-parallel collectives we run ourselves over NVLink (csrc/peer.cu).

``PeerWorkspace`` owns, per TP group, two pairs of buffers that every rank of the group has mapped into its address space
(``torch.distributed._symmetric_memory``: CUDA VMM allocations whose handles are exchanged once at rendezvous; torch is
used for the allocation / mapping / device-side barrier only, the data movement is our kernels and copy-engine copies):

  * ``partial[k]``  -- where a rank's rowwise GEMM writes its partial sums [T, H]; peers pull the rows they own;
  * ``shard[k]``    -- where a rank publishes its token shard [T/N, H]; peers copy it out for their all-gather.

Protocol (one device-side barrier per collective, no per-message flags): op j uses buffer pair j % 2.
    write own buffer (op j)  ->  barrier_j  ->  read peers' buffers (op j)
A rank can only overwrite pair j % 2 again in op j + 2, i.e. after barrier_{j+1}; every peer issued its op-j reads before
arriving at barrier_{j+1} (stream order), so the write-after-read hazard is covered by the same barriers.

Validated on NVLink in round 2 (NCCL-free parity at 2 and 8 GPUs with tests/cuda/tp_check.py, profiles/r02_call2.log and
profiles/r02_call3_tp8.log) and the default transport of ``bench.py`` for N > 1; the CPU suite exercises the same host logic
through ``tests/_fake_ops.py`` and a gloo stand-in for the peer mapping (tests/test_tp_gloo.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class PeerWorkspace:
    def __init__(self, group, device: torch.device | None = None, scatter_epilogue: bool = False):
        # scatter_epilogue: the rowwise GEMM stores every finished tile straight into the owner rank's slot (SCATTER mode of
        # the CTA-pair GEMM) and the reduction reads local memory only; otherwise the GEMM writes locally and peers pull
        self.scatter_epilogue = scatter_epilogue
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._elems = 0
        self._partial = [None, None]   # (local tensor, symmetric-memory handle)
        self._shard = [None, None]
        self._j_partial = 0
        self._j_shard = 0
        self.copy_streams = []

    # ------------------------------------------------------------------------------------------------ allocation
    def _alloc(self, elems: int):
        import torch.distributed._symmetric_memory as sm

        t = sm.empty(elems, dtype=torch.bfloat16, device=self.device)
        return t, sm.rendezvous(t, self.group)

    def reserve(self, elems: int) -> None:
        """Collective: (re)allocate the four buffers for at least ``elems`` bf16 elements.  Sizes are a function of the
        batch shape and the model only, so every rank reaches the same decision at the same call."""
        if elems <= self._elems:
            return
        elems = (elems + 1023) // 1024 * 1024
        self._partial = [self._alloc(elems), self._alloc(elems)]
        self._shard = [self._alloc((elems + self.world - 1) // self.world), self._alloc((elems + self.world - 1) // self.world)]
        self._elems = elems
        if not self.copy_streams:
            # several side streams: the pulls from different peers run on different copy engines in parallel (one stream
            # serialises them: 7 x 16.7 MB at single-engine speed per all-gather at N = 8)
            self.copy_streams = [torch.cuda.Stream(device=self.device) for _ in range(min(4, max(1, self.world - 1)))]

    # ---------------------------------------------------------------------------------------- reduce-scatter side
    def next_partial(self, rows: int, cols: int) -> torch.Tensor:
        """This rank's partial-sum buffer for the next reduce-scatter, as a [rows, cols] view (GEMM output operand)."""
        self.reserve(rows * cols)
        self._j_partial += 1
        t, _ = self._partial[self._j_partial % 2]
        return t[: rows * cols].view(rows, cols)

    def next_staging(self, rows: int, cols: int):
        """Scatter-epilogue layout of the same buffer: ``world`` slots of [rows, cols]; slot s of rank r receives the partial
        sums rank s computed for r's rows.  Returns (where this rank's GEMM stores block r for every r, this rank's slots)."""
        self.reserve(rows * cols * self.world)
        self._j_partial += 1
        _, hdl = self._partial[self._j_partial % 2]
        slot_bytes = rows * cols * 2
        dest = [int(hdl.buffer_ptrs[r]) + self.rank * slot_bytes for r in range(self.world)]
        mine = [int(hdl.buffer_ptrs[self.rank]) + s * slot_bytes for s in range(self.world)]
        return dest, mine

    def publish_partial(self) -> None:
        """Device-side barrier on the current stream: everything written into the current partial buffer before this
        point is visible to the peers' pulls issued after it."""
        self._partial[self._j_partial % 2][1].barrier(channel=0)

    def partial_ptrs(self) -> list[int]:
        return [int(p) for p in self._partial[self._j_partial % 2][1].buffer_ptrs]

    # ------------------------------------------------------------------------------------------- all-gather side
    def publish_shard(self, local2: torch.Tensor) -> None:
        """Copy this rank's [T/N, K] shard into its peer-visible slot and barrier."""
        self.reserve(local2.numel() * self.world)
        self._j_shard += 1
        t, hdl = self._shard[self._j_shard % 2]
        t[: local2.numel()].view_as(local2).copy_(local2)
        hdl.barrier(channel=1)

    def peer_shard(self, src: int, rows: int, cols: int) -> torch.Tensor:
        """[rows, cols] view of rank ``src``'s published shard (peer memory; valid until the op after next)."""
        return self._shard[self._j_shard % 2][1].get_buffer(src, (rows, cols), torch.bfloat16)

    def copy_context(self, i: int = 0):
        """Side stream ``i`` (round-robin over the copy streams) for a copy-engine pull; it first waits for everything issued
        so far on the current stream (the barrier included)."""
        st = self.copy_streams[i % len(self.copy_streams)]
        st.wait_stream(torch.cuda.current_stream())
        return torch.cuda.stream(st)

    def join_copies(self) -> None:
        """The current stream waits for the pulls issued inside ``copy_context``."""
        cur = torch.cuda.current_stream()
        for st in self.copy_streams:
            cur.wait_stream(st)
