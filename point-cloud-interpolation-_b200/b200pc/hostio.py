"""Host-buffer entry points: pinned host tensors in, pinned host tensors out.

The reference's loaders hand the models host tensors (`.cuda(non_blocking=True)` in test.py:56-62 of the
upstream copy) and read results back with `.cpu()`.  For the index-producing searches the read-back is the
expensive part (int64 indices: 16.8 MB for C2 against a 0.72 ms kernel), so these helpers can split the batch into
chunks and double-buffer: while chunk i's indices travel device->host on one stream, chunk i+1 is searched on
the other.  Results are identical to the plain calls (every op is independent per batch item).

A chunk must still fill the GPU, though: the search kernel needs ~131 000 queries (148 SMs x 28 warps x 32) in flight,
and C2 (8 x 16 384) is exactly one such wave -- measured on B200, C2 end to end: 1 chunk 1.13 ms, 2 chunks 1.39 ms,
4 chunks 1.58 ms (tools/hostio_chunks.py).  `chunks="auto"` (the default) therefore only splits batches that hold
several waves.
"""
import torch

from . import pointnet2_utils as P

_STREAMS = {}
_WAVE_QUERIES = 148 * 28 * 32      # queries one resident wave of the search kernel holds (B200, 14 warps x 2 CTAs per SM)


def _streams(device, n):
    key = (device.index, n)
    if key not in _STREAMS:
        _STREAMS[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
    return _STREAMS[key]


def plan_chunks(B, S, chunks="auto"):
    """how many batch chunks a host-buffer search is split into: "auto" = one chunk per resident wave of queries."""
    if chunks == "auto":
        chunks = (B * S) // _WAVE_QUERIES
    return max(1, min(int(chunks), max(B, 1)))


def _pipelined(op, host_inputs, host_out, device, chunks):
    device = torch.device(device)
    B = host_inputs[0].shape[0]
    chunks = plan_chunks(B, host_inputs[-1].shape[1], chunks)         # host_inputs[-1] = the queries [B,S,3]
    bounds = [(i * B // chunks, (i + 1) * B // chunks) for i in range(chunks)]
    streams = _streams(device, 2)
    cur = torch.cuda.current_stream(device)
    for s in streams:
        s.wait_stream(cur)
    for i, (lo, hi) in enumerate(bounds):
        s = streams[i % 2]
        with torch.cuda.stream(s):
            dev_in = [h[lo:hi].to(device, non_blocking=True) for h in host_inputs]
            host_out[lo:hi].copy_(op(*dev_in), non_blocking=True)
    for s in streams:
        cur.wait_stream(s)
    return host_out


def knn_point_host(nsample, xyz, new_xyz, out=None, device="cuda", chunks="auto"):
    """knn_point on pinned HOST tensors xyz [B,N,3], new_xyz [B,S,3] -> pinned host int64 [B,S,nsample]
    (int32 when `out` is an int32 tensor: half the bytes over PCIe, same values).
    The copy into `out` is asynchronous: synchronise the device's current stream before reading it."""
    if out is None:
        out = torch.empty(xyz.shape[0], new_xyz.shape[1], nsample, dtype=torch.int64).pin_memory()
    if out.dtype == torch.int32:
        from . import ops
        return _pipelined(lambda r, q: ops.knn_search_i32(r, q, nsample, ops.FORM_REF_NORM_FIRST), [xyz, new_xyz], out, device, chunks)
    return _pipelined(lambda r, q: P.knn_point(nsample, r, q), [xyz, new_xyz], out, device, chunks)


class KnnHostPipeline:
    """Back-to-back host-buffer searches with the copies of consecutive CALLS overlapped: call i's read-back (the
    expensive part: indices over PCIe) runs on one stream while call i+1's upload and search run on the other.
    Every call still does its own H2D of the inputs and D2H of its result; `submit` returns immediately and
    `finish` joins both streams into the caller's current stream.  Buffers are per slot (two calls in flight)."""

    def __init__(self, nsample, device="cuda", index_dtype=torch.int64):
        self.k = int(nsample)
        self.device = torch.device(device)
        self.streams = _streams(self.device, 2)
        self.dtype = index_dtype
        self.calls = 0

    def submit(self, xyz, new_xyz, out):
        from . import ops
        s = self.streams[self.calls % 2]
        if self.calls < 2:
            s.wait_stream(torch.cuda.current_stream(self.device))
        self.calls += 1
        with torch.cuda.stream(s):
            r = xyz.to(self.device, non_blocking=True); q = new_xyz.to(self.device, non_blocking=True)
            if self.dtype == torch.int32:
                idx = ops.knn_search_i32(r, q, self.k, ops.FORM_REF_NORM_FIRST)
            else:
                idx = P.knn_point(self.k, r, q)
            out.copy_(idx, non_blocking=True)
        return out

    def finish(self):
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)
        self.calls = 0


def query_ball_point_host(radius, nsample, xyz, new_xyz, out=None, device="cuda", chunks="auto"):
    if out is None:
        out = torch.empty(xyz.shape[0], new_xyz.shape[1], nsample, dtype=torch.int64).pin_memory()
    return _pipelined(lambda r, q: P.query_ball_point(radius, nsample, r, q), [xyz, new_xyz], out, device, chunks)
