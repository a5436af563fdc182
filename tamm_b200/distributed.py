"""
Multi-GPU layer: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch; gloo on CPU for the tests).

The path shards by PROFILE: every (profile, frequency, azimuth) lane is independent through the whole
pipeline and the only cross-lane operation, the azimuth mean in Pm/dP (layerstack.py:262,266), stays inside a profile.
So the compute path has no collective at all; the one collective offered here is the final all-gather of results
(`all_gather_profiles`), which callers use when every rank (or rank 0) needs the whole ensemble.
"""
import os

import numpy as np

from .ensemble import shard_range


def world():
    """(rank, world_size, local_rank) from the torchrun environment (single process: 0, 1, 0)."""
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def init(backend=None):
    """Initialise torch.distributed from the environment if WORLD_SIZE > 1.  Returns (rank, world_size)."""
    import torch
    import torch.distributed as dist
    rank, ws, local = world()
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            # NCCL's stream gets HIGH priority: the stack kernel keeps every SM full with one CTA each, and a collective
            # enqueued behind it (PipelinedGather) only starts moving data when its CTAs win the SMs that free up at the next
            # wave boundary -- at default priority the stack kernel's own queued CTAs take them and the "overlapped"
            # gather runs after the kernels instead of under them (measured at N = 2: 107.0 ms vs 107.5 ms back to back).
            opts = None
            if os.environ.get('TAMM_NCCL_HIGH_PRIORITY', '1') != '0':
                try:
                    opts = dist.ProcessGroupNCCL.Options()
                    opts.is_high_priority_stream = True
                except Exception:
                    opts = None
            if opts is not None:
                dist.init_process_group('nccl', device_id=torch.device('cuda', local), pg_options=opts)
            else:
                dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, ws


def shard_sizes(P, ws):
    return [shard_range(P, r, ws)[1] - shard_range(P, r, ws)[0] for r in range(ws)]


def all_gather_profiles(local, P_total, group=None):
    """All-gather a tensor whose leading axis is this rank's profile shard into the full (P_total, ...) tensor.
    Works for uneven shards (pads to the largest shard).  `local` may live on the GPU (NCCL) or the CPU (gloo)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    ws = dist.get_world_size(group)
    sizes = shard_sizes(P_total, ws)
    assert local.shape[0] == sizes[dist.get_rank(group)], (local.shape, sizes)
    m = max(sizes)
    cplx = local.is_complex()
    x = torch.view_as_real(local) if cplx else local        # NCCL moves raw doubles
    if x.shape[0] < m:
        pad = torch.zeros((m - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], dim=0)
    x = x.contiguous()
    out = torch.empty((ws * m,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)
    parts = [out[r * m: r * m + sizes[r]] for r in range(ws)]
    full = torch.cat(parts, dim=0) if len(set(sizes)) > 1 else out
    return torch.view_as_complex(full.contiguous()) if cplx else full


def _round_sizes(P_total, B, ws):
    """Block sizes of the rounds: B may be one size (uniform rounds) or a sequence of per-round sizes; every rank
    contributes a block of B[r] profiles to round r, so sum(B) * ws == P_total."""
    sizes = [int(B)] * (P_total // (ws * int(B))) if np.ndim(B) == 0 else [int(b) for b in B]
    if not sizes or min(sizes) < 1 or sum(sizes) * ws != P_total:
        raise ValueError('gather rounds %s x world %d do not tile P_total = %d' % (B, ws, P_total))
    return sizes


def gather_schedule(P_local, inner=3):
    """Block sizes for a pipelined gather of a shard of P_local profiles: a SMALL first block (nothing can be sent while the
    first block is being computed), a small last block (its transfer is the exposed tail), `inner` equal blocks between."""
    edge = max(1, P_local // 16)
    if P_local < 4 * edge + inner:
        return [P_local]
    rest = P_local - 2 * edge
    mid = [rest // inner + (1 if i < rest % inner else 0) for i in range(inner)]
    return [edge] + mid + [edge]


def block_cyclic_profiles(P_total, B, rank, ws):
    """Profile indices held by `rank` when blocks of B profiles are dealt round-robin over the ranks: round b gives block
    b*ws + rank to `rank`.  With this assignment the per-round all-gathers of `PipelinedGather` land in natural profile
    order.  Requires P_total % (ws * B) == 0."""
    sizes = _round_sizes(P_total, B, ws)
    out, base = [], 0
    for b in sizes:
        out.append(np.arange(base + rank * b, base + (rank + 1) * b))
        base += ws * b
    return np.concatenate(out)


class PipelinedGather:
    """The final all-gather of SURVEY.md 8(e), cut into one collective per profile block and overlapped with the kernels
    of the next block: `submit(b, block)` enqueues the all-gather of this rank's block of round b on the communication
    stream (it starts when the producing kernels on the current stream have finished and does not block the current
    stream), `wait()` makes the current stream wait for everything submitted.  `full` is the gathered tensor
    (P_total, ...): round b of rank r is profiles (b*ws + r)*B ... (see `block_cyclic_profiles`)."""

    def __init__(self, P_total, B, tail_shape, dtype, device, group=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.ws = dist.get_world_size(group) if dist.is_initialized() else 1
        self.sizes = _round_sizes(P_total, B, self.ws)        # B: one block size or a list of per-round sizes
        self.rounds = len(self.sizes)
        self.full = torch.empty((P_total,) + tuple(tail_shape), dtype=dtype, device=device)
        flat = torch.view_as_real(self.full) if self.full.is_complex() else self.full     # NCCL moves raw doubles
        self._rounds, base = [], 0
        for b in self.sizes:
            self._rounds.append(flat[base:base + self.ws * b])
            base += self.ws * b
        self._pending = []

    def submit(self, b, block):
        import torch
        import torch.distributed as dist
        assert 0 <= b < self.rounds and block.shape[0] == self.sizes[b]
        x = torch.view_as_real(block) if block.is_complex() else block
        if self.ws == 1:
            self._rounds[b].copy_(x)
            return
        self._pending.append(dist.all_gather_into_tensor(self._rounds[b], x.contiguous(), group=self.group, async_op=True))

    def wait(self):
        for w in self._pending:
            w.wait()
        self._pending = []
        return self.full


class PeerGather:
    """The final all-gather of north_star without NCCL: the gathered array lives in SYMMETRIC MEMORY
    (torch.distributed._symmetric_memory: every rank maps every rank's buffer over NVLink / NVSwitch) and each rank
    writes its finished block straight into slot (round b, rank) of every rank's array.  `submit(b, block)` enqueues that
    write on a high-priority side stream which waits for the producing kernels, so it runs under the kernels of the next
    block; `wait()` joins the side stream(s) and runs a device-side barrier over the group, after which `full` holds the
    whole ensemble on every rank.  Layout and arguments as PipelinedGather (block-cyclic rounds).  Transports:
      'mc'  one store per 16 bytes to the group's MULTICAST address (tamm_peer_push, include/tamm_b200.h): the NVSwitch
            replicates it into all replicas, the sender's link carries the block once -- a few CTAs for about a
            millisecond per block, against the ~32 SMs an NCCL all-gather kernel holds for the whole transfer;
      'uc'  the same kernel with one store per destination rank (no multicast support);
      'ce'  copy engines: one cudaMemcpyAsync per destination rank on per-peer streams (no SM at all, ~450 GB/s).
    mode='auto' picks 'mc' when the symmetric allocation has a multicast mapping, else 'uc'.
    Reuse: `full` is written by the OTHER ranks; after `wait()` it is complete and stable until any rank submits again, so a
    job that gathers repeatedly must finish reading `full` on every rank (e.g. a `dist.barrier()` or the next `wait()` of a
    second PeerGather used alternately) before the next round of submits."""

    def __init__(self, P_total, B, tail_shape, dtype, device, group=None, mode='auto', ctas=8):
        import ctypes
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.ws, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if dtype not in (torch.complex128, torch.float64):
            raise ValueError('PeerGather moves float64 / complex128 arrays')
        self.sizes = _round_sizes(P_total, B, self.ws)        # B: one block size or a list of per-round sizes
        self.rounds, self.ctas = len(self.sizes), int(ctas)
        self._base = [0]
        for b in self.sizes:
            self._base.append(self._base[-1] + self.ws * b)   # first profile of each round
        self._tail = tuple(tail_shape) + ((2,) if dtype == torch.complex128 else ())
        shape = (P_total,) + self._tail
        self._flat = symm_mem.empty(shape, dtype=torch.float64, device=device)
        self._hdl = symm_mem.rendezvous(self._flat, self.group)
        self.full = torch.view_as_complex(self._flat) if dtype == torch.complex128 else self._flat
        mc = int(getattr(self._hdl, 'multicast_ptr', 0) or 0)
        if mode == 'auto':
            mode = 'mc' if mc else 'uc'
        if mode == 'mc' and not mc:
            raise RuntimeError('PeerGather: this symmetric allocation has no multicast mapping')
        self.mode = mode
        self._row_bytes = int(np.prod(self._tail)) * 8
        off = int(getattr(self._hdl, 'offset', 0) or 0)
        if mode == 'mc':
            self._dst_base = [mc + off]
        elif mode == 'uc':
            self._dst_base = [int(p) + off for p in self._hdl.buffer_ptrs]
        else:
            self._peers = [self._hdl.get_buffer(r, shape, torch.float64) for r in range(self.ws)]
        nstreams = self.ws if mode == 'ce' else 1
        self._streams = [torch.cuda.Stream(device, priority=-1) for _ in range(nstreams)]
        self._lib, self._ctypes, self._torch = _lib.load(), ctypes, torch
        self._check = _lib.check

    def submit(self, b, block):
        torch, ct = self._torch, self._ctypes
        assert 0 <= b < self.rounds and block.shape[0] == self.sizes[b]
        x = torch.view_as_real(block) if block.is_complex() else block
        nbytes = self.sizes[b] * self._row_bytes
        assert x.is_contiguous() and x.numel() * 8 == nbytes
        ready = torch.cuda.Event()
        ready.record()
        first = self._base[b] + self.rank * self.sizes[b]                       # first profile of (round b, this rank)
        if self.mode == 'ce':
            for k in range(self.ws):
                r = (self.rank + k) % self.ws      # own copy first, then the peers in ring order (spreads the link load)
                st = self._streams[r]
                st.wait_event(ready)
                with torch.cuda.stream(st):
                    self._peers[r][first:first + self.sizes[b]].copy_(x, non_blocking=True)
                x.record_stream(st)
            return
        st = self._streams[0]
        st.wait_event(ready)
        dst = (ct.c_ulonglong * len(self._dst_base))(*[p + first * self._row_bytes for p in self._dst_base])
        rc = self._lib.tamm_peer_push(x.data_ptr(), nbytes, dst, len(self._dst_base), 1 if self.mode == 'mc' else 0,
                                      self.ctas, ct.c_void_p(st.cuda_stream))
        self._check(rc, self._lib)
        x.record_stream(st)

    def wait(self):
        torch = self._torch
        cur = torch.cuda.current_stream()
        for st in self._streams:
            cur.wait_stream(st)
        self._hdl.barrier()        # every rank's writes into THIS rank's array were enqueued before and are complete after
        return self.full

    def close(self):
        """Collective: call on every rank before dropping the object.  Freeing symmetric memory unmaps this rank's buffer from
        its peers and from the multicast group, so no rank may still have a store to it in flight: finish this rank's own work,
        then meet the others (host-side barrier), and only then release the views and the handle."""
        import torch.distributed as dist
        self._torch.cuda.synchronize()
        dist.barrier(self.group)
        self.full = self._flat = self._hdl = None
        self._peers = None
        self._dst_base = []


def _via_host(t, group):
    """gloo moves host memory only: CUDA tensors are staged through the CPU for that backend (world-size-2 tests on ONE
    GPU; NCCL refuses two ranks on one device)."""
    import torch.distributed as dist
    return t.is_cuda and dist.get_backend(group) == 'gloo'


def all_reduce_sum(t, group=None):
    """In-place sum over the ranks of `group` (no-op without a process group)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    if _via_host(t, group):
        h = t.cpu()
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=group)
        t.copy_(h)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_reduce_azimuth_sums(sums, group=None):
    """Azimuth sharding of a single stack (configs[1], [3]): the per-depth sums of |E| over the local azimuths
    (tamm_returns_partial) are summed over the ranks before the dB step (layerstack.py:262) -- (N-1) x 2 doubles per
    (profile, frequency), the only exchange on the compute path of any configuration."""
    return all_reduce_sum(sums, group)


def all_gather_axis(local, axis, n_total, group=None):
    """All-gather tensors that are shards of one axis (balanced contiguous split, `shard_range`) into the full tensor on
    every rank.  Uneven shards are padded to the largest one for the collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    ws = dist.get_world_size(group)
    sizes = shard_sizes(n_total, ws)
    assert local.shape[axis] == sizes[dist.get_rank(group)], (tuple(local.shape), axis, sizes)
    m = max(sizes)
    x = local.movedim(axis, 0)
    cplx = x.is_complex()
    x = torch.view_as_real(x.contiguous()) if cplx else x.contiguous()
    if x.shape[0] < m:
        x = torch.cat([x, torch.zeros((m - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)], dim=0)
    dev = x.device
    if _via_host(x, group):
        x = x.cpu()
    out = torch.empty((ws * m,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    full = torch.cat([out[r * m: r * m + sizes[r]] for r in range(ws)], dim=0) if len(set(sizes)) > 1 else out
    full = full.to(dev)
    if cplx:
        full = torch.view_as_complex(full.contiguous())
    return full.movedim(0, axis)


class AxisShard:
    """How ONE stack solve (LayerStack.get_returns) is split over the ranks of a process group:
      axis='azimuth'   : rank r evaluates its contiguous share of the azimuth grid; the per-depth sums of |E| over azimuth
                         are all-reduced (tiny) before the dB step -- configs[1], configs[3];
      axis='frequency' : rank r evaluates its share of the frequency vector; no exchange at all -- configs[2].
    With gather=True (default) every rank ends up with the full result (all-gather along the sharded axis: the final
    NCCL all-gather of north_star); with gather=False each rank returns its own shard."""

    def __init__(self, axis, group=None, gather=True):
        import torch.distributed as dist
        if axis not in ('azimuth', 'frequency'):
            raise ValueError('AxisShard: axis must be "azimuth" or "frequency"')
        self.axis, self.group, self.gather = axis, group, gather
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1

    def range(self, n):
        return shard_range(n, self.rank, self.world)


def gather_bytes(P_total, N, Q, F, what='jones'):
    """Bytes each rank RECEIVES in the final all-gather (SURVEY.md 8e budget check)."""
    per_eval = {'jones': 64, 'fields': 32, 'power': 16}[what]
    return int(per_eval) * int(P_total) * int(N - 1) * int(Q) * int(F)
