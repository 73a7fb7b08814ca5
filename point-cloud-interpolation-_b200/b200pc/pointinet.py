"""PointINet forward (FlowNet3D scene flow x2 + warp + points fusion) on the b200pc kernels.

This is the CALLER of the hot path that BASELINE.json's first metric is quoted on ("interp
frames/s, PointINet, 16 384 points"); it exists so that the end-to-end number can be measured on a
box that has no copy of the reference.  It mirrors the upstream model the reference vendors
(PointINet20230424/models/models.py:9-125, layers.py:23-216 and :335-416): same sub-module and
parameter names (`flow.set_conv1.conv.0.weight`, `fusion.conv.0.weight`, ...) so a reference
state_dict loads unchanged, same tensor layouts ([B,C,N] between layers), same CPU-RNG draws
(FPS start indices, two randperm per fused frame) so torch.manual_seed reproduces the reference.

Every geometric step goes through `backend`, whose default is the CUDA library
(b200pc.pointnet2_utils + b200pc.pytorch3d_shim).  The MLPs (1x1 Conv + BatchNorm + ReLU, max over
neighbours, softmax) are stock torch/cuDNN: they are outside the hot path.
tests/ swap in a CPU backend to check this glue against the real upstream model bit for bit.
"""
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


class RngTape:
    """The CPU-RNG draws of one forward, in the reference's order: one torch.randint per FPS call
    (Utils/Pointnet2Utils.py:76) and two torch.randperm per fused frame (PointINet20230424/models/
    layers.py:402-403).  mode "eager": draw and upload on the spot (what the reference does).
    mode "record": same, but remember the sequence and keep every draw in a slice of one flat device
    tensor.  mode "replay": hand out those slices without touching the RNG -- this is what runs inside a
    CUDA-graph capture; `refill()` then redraws the whole sequence on the CPU generator (same calls, same
    order, so torch.manual_seed still reproduces the reference) and uploads it with ONE copy per frame."""

    def __init__(self, device):
        self.device, self.mode = device, "eager"
        self.specs, self.views, self.pos = [], [], 0
        self.flat_dev = self.flat_host = None

    def _draw(self, spec):
        kind, a, b = spec
        return torch.randint(0, a, (b,), dtype=torch.long) if kind == "randint" else torch.randperm(a)[:b]

    def _next(self, spec):
        if self.mode == "eager":
            return self._draw(spec).to(self.device)
        if self.mode == "record":
            self.specs.append(spec)
            v = self._draw(spec).to(self.device)
            self.views.append(v)
            return v
        v = self.views[self.pos]; self.pos += 1           # replay
        assert self.specs[self.pos - 1] == spec, "RNG tape out of sync with the recorded forward"
        return v

    def randint(self, high, count):
        return self._next(("randint", int(high), int(count)))

    def randperm(self, n, keep):
        return self._next(("randperm", int(n), int(keep)))

    def finish_recording(self):
        total = sum(s[2] for s in self.specs)
        self.flat_dev = torch.empty(total, dtype=torch.long, device=self.device)
        # two pinned staging buffers: frame i+1 is drawn into one while frame i's upload may still be reading the other
        self.flat_host = [torch.empty(total, dtype=torch.long).pin_memory() for _ in range(2)]
        self.uploaded = [None, None]          # event recorded after the upload out of each staging buffer
        self.turn = 0
        off, views = 0, []
        for s, old in zip(self.specs, self.views):
            v = self.flat_dev[off:off + s[2]]; v.copy_(old); views.append(v); off += s[2]
        self.views, self.mode = views, "replay"

    def refill(self):
        """redraw the whole sequence for the next frame on the CPU generator and upload it with one copy"""
        self.pos = 0
        buf = self.flat_host[self.turn]
        if self.uploaded[self.turn] is not None:
            self.uploaded[self.turn].synchronize()      # the upload that last read this staging buffer has finished
        off = 0
        for s in self.specs:
            buf[off:off + s[2]] = self._draw(s); off += s[2]
        self.flat_dev.copy_(buf, non_blocking=True)
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream(self.device))
        self.uploaded[self.turn] = ev
        self.turn ^= 1


def cuda_backend(tape=None):
    from . import ops, pointnet2_utils as P, pytorch3d_shim as S
    be = types.SimpleNamespace(
        index_points=P.index_points, query_ball_point=P.query_ball_point, knn_point=P.knn_point,
        three_nn_weights=P.three_nn_weights, three_interpolate=P.three_interpolate, group_points=P.group_points,
        knn_points=S.knn_points, knn_gather=S.knn_gather, fusion_group=P.fusion_group, tape=tape)
    if tape is None:
        be.farthest_point_sample = P.farthest_point_sample
        be.sample_points = P.sample_points
        be.randperm = lambda n, keep, device: torch.randperm(n)[:keep].to(device)
    else:
        be.farthest_point_sample = lambda xyz, npoint: ops.fps(xyz, npoint, tape.randint(xyz.shape[1], xyz.shape[0]))
        be.sample_points = lambda xyz, npoint: ops.fps(xyz, npoint, tape.randint(xyz.shape[1], xyz.shape[0]), want_xyz=True)
        be.randperm = lambda n, keep, device: tape.randperm(n, keep)
    return be


_SIDE_STREAMS = {}


def _concurrently(first, second, ref_tensor, slot):
    """run first() on the current stream and second() on a side stream forked BEFORE first() was enqueued, then join.
    Python call order (hence the order of CPU-RNG draws) stays first -> second, exactly like the sequential reference;
    on the device the two branches overlap (they are independent: e.g. the farthest-point sampling of frame 1 and of
    frame 2, each of which occupies one 8-SM cluster for ~0.6 ms).  Under CUDA-graph capture the fork/join becomes
    two parallel branches of the graph.  CPU tensors and eager (non-captured) calls: plain sequential calls."""
    if not ref_tensor.is_cuda or not torch.cuda.is_current_stream_capturing():
        return first(), second()      # eager calls are bound by Python dispatch: forking only adds host work there
    dev = ref_tensor.device
    cur = torch.cuda.current_stream(dev)
    key = (dev.index, cur.cuda_stream, slot)
    side = _SIDE_STREAMS.get(key)
    if side is None:
        side = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    fork = torch.cuda.Event()
    fork.record(cur)
    a = first()
    side.wait_event(fork)
    with torch.cuda.stream(side):
        b = second()
    cur.wait_stream(side)
    return a, b      # the caller keeps both results alive until it returns, so no block is recycled across streams early


def _pointwise_mlp(channels):
    """[Conv2d 1x1, BatchNorm2d(eps=1e-3), ReLU] * len -- the block every FlowNet3D layer uses."""
    mods = []
    for cin, cout in zip(channels[:-1], channels[1:]):
        mods += [nn.Conv2d(cin, cout, 1, bias=True), nn.BatchNorm2d(cout, eps=0.001), nn.ReLU()]
    return nn.Sequential(*mods)


def _rows(t):
    """[B,C,N] -> contiguous point-major [B,N,C]."""
    return t.transpose(1, 2).contiguous()


def _group(be, refs_cf, centres_cf, feats_cf, nsample, radius=None, refs_rows=None, centres_rows=None):
    """neighbourhoods of `centres` in `refs` -> [B, 3+D, nsample, S]: relative xyz then features.
    radius=None selects kNN (flow embedding / up-conv), else the ball query (set conv).  Callers that already hold
    the point-major copies pass them in (the reference re-permutes the same tensors in every layer)."""
    refs = _rows(refs_cf) if refs_rows is None else refs_rows
    centres = _rows(centres_cf) if centres_rows is None else centres_rows
    feats = _rows(feats_cf)
    B, S, _ = centres.shape
    if radius is None:
        idx = be.knn_point(nsample, refs, centres)
    else:
        idx = be.query_ball_point(radius, nsample, refs, centres)
    fused = getattr(be, "group_points", None)
    if fused is not None:                                                 # one kernel instead of 2 gathers, sub, cat, permute
        return fused(refs, centres, feats, idx)
    rel = be.index_points(refs, idx) - centres.view(B, S, 1, 3)
    out = torch.cat([rel, be.index_points(feats, idx)], dim=-1)          # [B,S,ns,3+D]
    return out.permute(0, 3, 2, 1).contiguous()


class SetConv(nn.Module):
    def __init__(self, be, npoint, radius, nsample, cin, couts):
        super().__init__()
        self.be, self.npoint, self.radius, self.nsample = be, npoint, radius, nsample
        self.conv = _pointwise_mlp([cin + 3, *couts])

    def forward(self, xyz, feats):
        rows = _rows(xyz)
        sample = getattr(self.be, "sample_points", None)
        if sample is not None:
            _, centre_rows = sample(rows, self.npoint)                                                 # FPS + gather behind one call
        else:
            centre_rows = self.be.index_points(rows, self.be.farthest_point_sample(rows, self.npoint))     # [B,S,3]
        centres = _rows(centre_rows)
        g = _group(self.be, xyz, centres, feats, self.nsample, self.radius, refs_rows=rows, centres_rows=centre_rows)
        return centres, self.conv(g).max(dim=2)[0]


class FlowEmbedding(nn.Module):
    def __init__(self, be, nsample, cin, couts):
        super().__init__()
        self.be, self.nsample = be, nsample
        self.conv = _pointwise_mlp([2 * cin + 3, *couts])

    def forward(self, xyz1, xyz2, feats1, feats2):
        g = _group(self.be, xyz2, xyz1, feats2, self.nsample)
        g = torch.cat([g, feats1.unsqueeze(2).expand(-1, -1, self.nsample, -1)], dim=1)
        return self.conv(g).max(dim=2)[0]


class SetUpConv(nn.Module):
    def __init__(self, be, nsample, cin1, cin2, couts1, couts2):
        super().__init__()
        self.be, self.nsample = be, nsample
        self.conv1 = _pointwise_mlp([cin1 + 3, *couts1])
        first = cin1 + cin2 + 3 if len(couts1) == 0 else couts1[-1] + cin2
        self.conv2 = _pointwise_mlp([first, *couts2])

    def forward(self, xyz1, xyz2, feats1, feats2):
        g = self.conv1(_group(self.be, xyz1, xyz2, feats1, self.nsample)).max(dim=2)[0]
        g = torch.cat([g, feats2], dim=1).unsqueeze(3)
        return self.conv2(g).squeeze(3)


class FeaturePropagation(nn.Module):
    def __init__(self, be, cin1, cin2, couts):
        super().__init__()
        self.be = be
        self.conv = _pointwise_mlp([cin1 + cin2, *couts])

    def forward(self, xyz_sparse, xyz_dense, feats_sparse, feats_dense):
        _, idx, w = self.be.three_nn_weights(_rows(xyz_dense), _rows(xyz_sparse), 0)
        up = self.be.three_interpolate(_rows(feats_sparse), idx, w).transpose(1, 2).contiguous()
        return self.conv(torch.cat([up, feats_dense], dim=1).unsqueeze(3)).squeeze(3)


class FlowNet3D(nn.Module):
    def __init__(self, be=None):
        super().__init__()
        be = be or cuda_backend()
        self.set_conv1 = SetConv(be, 1024, 0.5, 16, 3, [32, 32, 64])
        self.set_conv2 = SetConv(be, 256, 1.0, 16, 64, [64, 64, 128])
        self.flow_embedding = FlowEmbedding(be, 64, 128, [128, 128, 128])
        self.set_conv3 = SetConv(be, 64, 2.0, 8, 128, [128, 128, 256])
        self.set_conv4 = SetConv(be, 16, 4.0, 8, 256, [256, 256, 512])
        self.set_upconv1 = SetUpConv(be, 8, 512, 256, [], [256, 256])
        self.set_upconv2 = SetUpConv(be, 8, 256, 256, [128, 128, 256], [256])
        self.set_upconv3 = SetUpConv(be, 8, 256, 64, [128, 128, 256], [256])
        self.fp = FeaturePropagation(be, 256, 3, [256, 256])
        self.classifier = nn.Sequential(nn.Conv1d(256, 128, 1, bias=True), nn.BatchNorm1d(128, eps=0.001),
                                        nn.ReLU(), nn.Conv1d(128, 3, 1, bias=True))

    def forward(self, xyz1, xyz2, feats1, feats2):
        """[B,3,N] x4 -> flow [B,3,N] from frame 1 to frame 2."""
        def cloud1():
            pa, fa = self.set_conv1(xyz1, feats1)
            return (pa, fa) + self.set_conv2(pa, fa)

        def cloud2():
            pa, fa = self.set_conv1(xyz2, feats2)
            return (pa, fa) + self.set_conv2(pa, fa)

        (p1a, f1a, p1b, f1b), (p2a, f2a, p2b, f2b) = _concurrently(cloud1, cloud2, xyz1, "clouds")
        emb = self.flow_embedding(p1b, p2b, f1b, f2b)
        p1c, f1c = self.set_conv3(p1b, emb)
        p1d, f1d = self.set_conv4(p1c, f1c)
        u3 = self.set_upconv1(p1d, p1c, f1d, f1c)
        u2 = self.set_upconv2(p1c, p1b, u3, torch.cat([f1b, emb], dim=1))
        u1 = self.set_upconv3(p1b, p1a, u2, f1a)
        return self.classifier(self.fp(p1a, xyz1, u1, feats1))


def _conv_relu_stack(seq):
    """the (Conv2d 1x1, ReLU) pairs of a BN-folded point-wise MLP, or None if the Sequential holds anything else"""
    mods = list(seq.children())
    if len(mods) % 2:
        return None
    pairs = []
    for conv, act in zip(mods[0::2], mods[1::2]):
        if not (isinstance(conv, nn.Conv2d) and conv.kernel_size == (1, 1) and conv.bias is not None and isinstance(act, nn.ReLU)):
            return None
        pairs.append(conv)
    return pairs


class PointsFusion(nn.Module):
    def __init__(self, be, cin, couts):
        super().__init__()
        self.be = be
        self.conv = _pointwise_mlp([cin, *couts])
        self.batched = True          # False: always the reference's per-item loop (tests compare the two)

    def _score(self, feat):
        """max over the channels of the point-wise MLP (upstream layers.py:415-416): [B,4,N,2k] -> [B,N,2k].
        Inference with folded BatchNorm on the device: the three 1x1 convolutions run as GEMMs over channels-last rows with
        the bias + ReLU in the GEMM epilogue -- torch's conv2d adds the bias and applies the ReLU as two more passes over
        the [1,128,16384,32] activations (0.5 ms of a 3.2 ms frame, profiles/r02_pointinet_graph.txt).  Same arithmetic,
        different summation order inside the GEMM: results agree to fp32 rounding."""
        pairs = None if (self.training or torch.is_grad_enabled() or not feat.is_cuda) else _conv_relu_stack(self.conv)
        if pairs is None:
            return self.conv(feat).max(dim=1)[0]
        B, C, N, K = feat.shape
        x = feat.permute(0, 2, 3, 1).reshape(B * N * K, C)
        # same precision policy as the convolutions these GEMMs replace (torch.backends.cudnn.allow_tf32)
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32
        try:
            for conv in pairs:
                x = torch._addmm_activation(conv.bias, x, conv.weight.view(conv.out_channels, conv.in_channels).t())
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
        from . import ops
        return ops.channel_max(x).view(B, N, K)

    def _neighbours(self, query_cf, ref_cf, ref_feat_cf, k):
        q, r = _rows(query_cf), _rows(ref_cf)
        fused = getattr(self.be, "fusion_group", None)
        if fused is not None and k >= 1 and not (torch.is_grad_enabled() and (q.requires_grad or r.requires_grad)):
            # one C call (direct-form search + one kernel) instead of knn_points, knn_gather, sub, norm, cat and three
            # permute/contiguous copies (upstream layers.py:346-368); falls through to those for differentiable coordinates
            feat, nn, extra, _ = fused(q, r, k, _rows(ref_feat_cf))
            return feat, nn, extra
        res = self.be.knn_points(q, r, K=k, return_nn=True)
        resi = res.knn - q.unsqueeze(2)                                   # [B,N,k,3]
        feat = torch.cat([resi, resi.norm(dim=-1, keepdim=True)], dim=-1)  # + distance channel
        extra = self.be.knn_gather(_rows(ref_feat_cf), res.idx)            # [B,N,k,C]
        cf = lambda x: x.permute(0, 3, 1, 2).contiguous()
        return cf(feat), cf(res.knn), cf(extra)

    def forward(self, xyz1, xyz2, feats1, feats2, k, t, t_host=None):
        B, _, N = xyz1.shape
        if t_host is None:
            t_host = t.detach().reshape(B).to("cpu", torch.float32)        # device->host sync, as in the reference
        rp = getattr(self.be, "randperm", None) or (lambda n, keep, device: torch.randperm(n)[:keep].to(device))
        if self.batched and B > 1 and all(float(t_host[i]) == float(t_host[0]) for i in range(B)):
            # one time stamp for the whole batch -> every item has the same (n1, n2, k1, k2), so the
            # per-item loop of the reference (upstream layers.py:389-411) collapses into two batched searches.  The
            # random subsets are still drawn item by item in the reference's order; each item's result is unchanged.
            n2 = int(N * t_host[0]); n1 = N - n2
            k2 = int(k * t_host[0]); k1 = k - k2
            s1, s2 = [], []
            for i in range(B):
                s1.append(rp(N, n1, xyz1.device)); s2.append(rp(N, n2, xyz1.device))
            pick = lambda x, sel: x.gather(2, torch.stack(sel).unsqueeze(1).expand(-1, 3, -1))
            mixed = torch.cat((pick(xyz1, s1), pick(xyz2, s2)), dim=-1)
            (f1, g1, e1), (f2, g2, e2) = _concurrently(lambda: self._neighbours(mixed, xyz1, feats1, k1),
                                                     lambda: self._neighbours(mixed, xyz2, feats2, k2), mixed, "fusion")
            feat, grouped, extra = torch.cat((f1, f2), dim=-1), torch.cat((g1, g2), dim=-1), torch.cat((e1, e2), dim=-1)
            w = F.softmax(self._score(feat), dim=-1)
            return (w.unsqueeze(1) * torch.cat([grouped, extra], dim=1)).sum(dim=-1)
        fa, ga, ea = [], [], []
        for i in range(B):
            n2 = int(N * t_host[i]); n1 = N - n2
            k2 = int(k * t_host[i]); k1 = k - k2
            sel1 = rp(N, n1, xyz1.device)                                  # CPU generator, like the reference
            sel2 = rp(N, n2, xyz1.device)
            a, b = xyz1[i:i + 1], xyz2[i:i + 1]
            mixed = torch.cat((a[:, :, sel1], b[:, :, sel2]), dim=-1)
            (f1, g1, e1), (f2, g2, e2) = _concurrently(lambda: self._neighbours(mixed, a, feats1[i:i + 1], k1),
                                                     lambda: self._neighbours(mixed, b, feats2[i:i + 1], k2), mixed, "fusion")
            fa.append(torch.cat((f1, f2), dim=-1)); ga.append(torch.cat((g1, g2), dim=-1)); ea.append(torch.cat((e1, e2), dim=-1))
        feat, grouped, extra = torch.cat(fa, 0), torch.cat(ga, 0), torch.cat(ea, 0)
        w = F.softmax(self._score(feat), dim=-1)                           # [B,N,2k]
        return (w.unsqueeze(1) * torch.cat([grouped, extra], dim=1)).sum(dim=-1)


class PointINet(nn.Module):
    def __init__(self, freeze=1, backend=None):
        super().__init__()
        be = backend or cuda_backend()
        self.flow = FlowNet3D(be)
        if freeze == 1:
            for p in self.parameters():
                p.requires_grad = False
        self.fusion = PointsFusion(be, 4, [64, 64, 128])

    def forward(self, points1, points2, features1, features2, t, t_host=None):
        """points [B,3+C,N] (xyz + extra channels), features [B,3,N] (zeros for LiDAR), t [B] in (0,1)
        -> fused frame [B,3+C,N] at time t.  t_host: the same values on the CPU (avoids the device->host
        read of t; required under CUDA-graph capture)."""
        extra1, extra2 = points1[:, 3:].contiguous(), points2[:, 3:].contiguous()
        xyz1, xyz2 = points1[:, :3].contiguous(), points2[:, :3].contiguous()
        with torch.no_grad():
            fwd, bwd = _concurrently(lambda: self.flow(xyz1, xyz2, features1, features2),
                                     lambda: self.flow(xyz2, xyz1, features2, features1), xyz1, "flows")
        tt = t.view(-1, 1, 1)
        return self.fusion(xyz1 + fwd * tt, xyz2 + bwd * (1 - tt), extra1, extra2, 32, tt, t_host)

    def fold_batchnorm_(self):
        """inference only: fold every eval-mode BatchNorm into the 1x1 convolution in front of it
        (71 fewer launches per frame).  Changes results at the 1e-6 level."""
        assert not self.training, "fold_batchnorm_ is for eval mode"
        for m in list(self.modules()):
            if isinstance(m, nn.Sequential):
                _fold_sequential_(m)
        return self


def _fold_sequential_(seq):
    mods = list(seq.children())
    out, i = [], 0
    while i < len(mods):
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        if isinstance(m, (nn.Conv1d, nn.Conv2d)) and isinstance(nxt, (nn.BatchNorm1d, nn.BatchNorm2d)):
            scale = nxt.weight / torch.sqrt(nxt.running_var + nxt.eps)
            with torch.no_grad():
                m.weight.mul_(scale.view(-1, *([1] * (m.weight.dim() - 1))))
                m.bias.copy_((m.bias - nxt.running_mean) * scale + nxt.bias)
            out.append(m); i += 2
        else:
            out.append(m); i += 1
    for k in list(seq._modules.keys()):
        del seq._modules[k]
    for j, m in enumerate(out):
        seq.add_module(str(j), m)


class GraphedPointINet:
    """PointINet forward captured ONCE as a CUDA graph for a fixed (batch, points, t): per frame the host
    only redraws the RNG tape, copies the inputs into the static buffers and replays the graph -- the
    ~650 kernel launches and their Python dispatch disappear from the frame time."""

    def __init__(self, state_dict=None, batch=1, npoints=16384, extra=1, t=0.5, device="cuda", fold_bn=True, freeze=1):
        self.device = torch.device(device)
        self.tape = RngTape(self.device)
        self.net = PointINet(freeze=freeze, backend=cuda_backend(self.tape)).eval()
        if state_dict is not None:
            self.net.load_state_dict(state_dict)
        self.net.to(self.device)
        if fold_bn:
            self.net.fold_batchnorm_()
        self.t_host = torch.full((batch,), float(t), dtype=torch.float32)
        self.static_in = [torch.zeros(batch, 3 + extra, npoints, device=self.device), torch.zeros(batch, 3 + extra, npoints, device=self.device),
                          torch.zeros(batch, 3, npoints, device=self.device), torch.zeros(batch, 3, npoints, device=self.device),
                          self.t_host.to(self.device)]
        self.graph = None
        self.static_out = None

    def _forward(self):
        with torch.no_grad():
            return self.net(*self.static_in, t_host=self.t_host)

    def capture(self, points1, points2, features1, features2):
        for dst, src in zip(self.static_in[:4], (points1, points2, features1, features2)):
            dst.copy_(src)
        self.tape.mode = "record"
        rng_state = torch.get_rng_state()                 # capturing must not consume the caller's CPU-RNG stream: the first
        self._forward()                                   # eager warm-up: records the RNG sequence, sizes every workspace
        torch.set_rng_state(rng_state)                    # replayed frame then makes the draws the reference's first forward would
        self.tape.finish_recording()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.tape.pos = 0
            self._forward()                               # warm-up on the capture stream
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        self.tape.pos = 0
        with torch.cuda.graph(self.graph):
            self.static_out = self._forward()
        return self

    def __call__(self, points1, points2, features1, features2):
        """inputs: CUDA or pinned-host tensors of the captured shapes -> fused frame (static output tensor)."""
        if self.graph is None:
            self.capture(points1, points2, features1, features2)
        self.tape.refill()
        for dst, src in zip(self.static_in[:4], (points1, points2, features1, features2)):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
