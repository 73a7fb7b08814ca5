"""Sharding of the hot path over the GPUs of one node.

The path shards without any exchange inside a kernel:
  * batch / frame-pair sharding: every op is independent per batch item -> no collective at all;
  * query sharding against a REPLICATED reference cloud: rows of kNN / ball query / three-NN /
    knn_points are independent per query, so each rank searches a contiguous slice of the queries and
    ONE all_gather assembles the index (and, if asked, distance / neighbour) outputs.  Payloads are
    small (65 536 x 8 B = 512 KiB of indices per call at C5), i.e. latency-bound over NVLink/NVSwitch.

Works with any torch.distributed backend: NCCL on the GPU box (one process per GPU), gloo in the CPU
tests (where `op` is a CPU stand-in; this module itself never computes anything).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """contiguous, balanced split of range(n): the first n % world shards get one extra item."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def batch_shard(tensors, world=None, rank=None):
    """slice every tensor along dim 0 for this rank (frame-pair / batch sharding, no communication)."""
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(tensors[0].shape[0], world, rank)
    return [t[lo:hi] for t in tensors]


def all_gather_rows(local, n_total, dim=1, group=None):
    """all_gather of ragged shards along `dim` (shards differ by at most one row): pad to the largest
    shard, one all_gather, trim.  Returns the full tensor with `n_total` rows on every rank."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    pad = longest - local.shape[dim]
    if pad:
        shape = list(local.shape); shape[dim] = pad
        local = torch.cat([local, local.new_zeros(shape)], dim=dim)
    bufs = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(bufs, local.contiguous(), group=group)
    return torch.cat([b.narrow(dim, 0, hi - lo) for b, (lo, hi) in zip(bufs, sizes)], dim=dim)


def query_sharded(op, queries, group=None):
    """run `op(query_slice)` on this rank's contiguous slice of `queries` [B,S,...] and all_gather the
    result(s) along dim 1.  `op` closes over the replicated reference cloud, e.g.
        idx = query_sharded(lambda q: knn_point(16, refs, q), queries)
    `op` may return one tensor or a tuple of tensors with the query dimension at dim 1."""
    world = dist.get_world_size(group); rank = dist.get_rank(group)
    S = queries.shape[1]
    lo, hi = shard_bounds(S, world, rank)
    out = op(queries[:, lo:hi].contiguous())
    if isinstance(out, tuple):
        return tuple(all_gather_rows(o, S, 1, group) if o is not None else None for o in out)
    return all_gather_rows(out, S, 1, group)


# ------------------------------------------------------------------------------------------------
# C5: PolyPCI.rebuild (K=1 nearest neighbour + its coordinates) query-sharded over the ranks
# ------------------------------------------------------------------------------------------------
class PeerSlab:
    """A [S_total, B, 4] fp32 buffer on every rank, allocated as torch symmetric memory and rendezvoused over the group:
    every rank holds the base pointers of all its peers' copies, so a kernel can store its shard of records straight
    into every peer over NVLink (`b200pc_rebuild_pack`) -- the all-gather needs no collective launch, only a barrier."""

    def __init__(self, s_total, batch, device, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.buf = symm.empty((int(s_total), int(batch), 4), dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def barrier(self):
        self.hdl.barrier()


def rebuild_sharded(refs, queries, mode="nccl", slab=None, group=None, pack=None):
    """`PolyPCI.rebuild` (PolyPCI/Models/Models_V1.py:102-114) with the queries sharded over the ranks and the refs
    replicated: every rank searches S/world queries and the (index, neighbour) records are assembled on every rank.
      mode "nccl": ONE all_gather_into_tensor of the [S/world, B, 4] record slabs (16 bytes per query);
      mode "peer": the producing kernel stores its slab into every rank's symmetric buffer (`slab`, a PeerSlab),
                   followed by one symmetric-memory barrier -- no collective kernel at all.
    S must divide evenly (C5: 65 536 / 8).  Returns the assembled records [S,B,4] (ops.unpack_rebuild splits them).
    `pack(refs, query_slice, s_offset, peer_ptrs)` defaults to ops.rebuild_pack (the CUDA entry); the gloo tests on CPU pass
    a stand-in, this module itself never computes anything."""
    if pack is None:
        from . import ops
        pack = lambda r, q, s_offset, peer_ptrs: ops.rebuild_pack(r, q, s_offset=s_offset, peer_ptrs=peer_ptrs)
    world = dist.get_world_size(group); rank = dist.get_rank(group)
    B, S, _ = queries.shape
    if S % world:
        raise ValueError("rebuild_sharded: %d queries do not divide over %d ranks (use query_sharded for ragged shards)" % (S, world))
    per = S // world
    mine = queries[:, rank * per:(rank + 1) * per].contiguous()
    if mode == "peer":
        slab.barrier()                                   # every peer is done reading the previous contents
        pack(refs, mine, rank * per, slab.ptrs)
        slab.barrier()                                   # every peer's stores have landed
        return slab.buf
    local = pack(refs, mine, 0, ())
    out = torch.empty(S, B, 4, dtype=torch.float32, device=queries.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out
