"""Host side of the training step on hand-written kernels (config 5: cfgs/refine_car.yaml).

Mirrors what PointNetDet does in train() mode (/root/reference/models/det_base.py:334-375 up to the head logits;
layer factories models/common.py:38-63) as a table of ``fcn_train_layer`` records over dense position-major
tensors, executed by csrc/train.cu:

    grouping (fcn_group_rows, all T*K rows incl. the reference's back-fill duplicates - they weight the batch
    statistics)  ->  per scale conv1..3 (+BN batch stats) -> mask + max over K -> FCN -> deconvs -> heads

The losses stay PyTorch ops on the (B*T2, 2+39) logits (~320 rows); their autograd gives dlogits,
from which ``backward`` runs the hand-written backward kernels.  Parameters live in ONE flat fp32 buffer (every
``nn.Parameter`` is a view), gradients in ONE flat bucket written in the parameter layout by the kernels:
one ``all_reduce`` per step and one fused Adam launch (``fcn_adam_step``).
BatchNorm statistics are per rank (the reference's DataParallel semantics: per-replica BN, no SyncBN).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import torch

from . import _lib
from .synth import fcn_layer_table


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _stream():
    return torch.cuda.current_stream().cuda_stream


class FlatParams:
    """All parameters of a module as views of one flat fp32 buffer (+ a flat gradient bucket of the same layout)."""

    def __init__(self, module: torch.nn.Module):
        ps = [p for p in module.parameters()]
        assert all(p.dtype == torch.float32 and p.is_cuda for p in ps), "fp32 CUDA parameters expected"
        n = sum(p.numel() for p in ps)
        dev = ps[0].device
        self.param = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.offsets: Dict[int, int] = {}
        off = 0
        for p in ps:
            k = p.numel()
            self.param[off: off + k].copy_(p.data.reshape(-1))
            p.data = self.param[off: off + k].view(p.shape)
            self.offsets[id(p)] = off
            off += k
        self.numel = n

    def gptr(self, p) -> int:
        return self.grad.data_ptr() + 4 * self.offsets[id(p)]

    def grad_view(self, p) -> torch.Tensor:
        off = self.offsets[id(p)]
        return self.grad[off: off + p.numel()].view(p.shape)

    def expose_grads(self, module):
        """``p.grad`` = view of the flat bucket (what an optimizer or a test reads after TrainStep.forward_backward)."""
        for p in module.parameters():
            p.grad = self.grad_view(p)


class LogitsFn(torch.autograd.Function):
    """autograd bridge for the reference's own training loop (``losses, _ = model(data); loss.backward();
    optimizer.step()``, train/train_net_det.py:121-128): forward = the kernel forward up to the head logits,
    backward = the kernel backward; parameter gradients are returned to autograd as copies of the flat bucket."""

    @staticmethod
    def forward(ctx, eng, pc, one_hot, ncent, *rest):
        centers, params = rest[:ncent], rest[ncent:]
        cls, reg = eng.forward(pc, list(centers), one_hot)
        ctx.eng, ctx.params, ctx.ncent = eng, params, ncent
        return cls.clone(), reg.clone()

    @staticmethod
    def backward(ctx, dcls, dreg):
        eng = ctx.eng
        eng.flat.grad.zero_()
        eng.backward(dcls.contiguous(), dreg.contiguous())
        grads = tuple(eng.flat.grad_view(p).clone() for p in ctx.params)
        return (None, None, None, None) + (None,) * ctx.ncent + grads


class TrainEngine:
    """Kernel tables + workspaces of one PointNetDet for one input shape (B, N, T...)."""

    def __init__(self, model, B: int, N: int, T: Sequence[int]):
        self.model = model
        arch = model.ARCH
        S = arch.num_scales
        assert len(T) == S
        self.B, self.N, self.T = int(B), int(N), tuple(int(t) for t in T)
        self.dev = next(model.parameters()).device
        self.V = model.num_vec
        self.flat = getattr(model, "_flat_params", None)
        if self.flat is None:
            self.flat = model._flat_params = FlatParams(model)
        dev, f32 = self.dev, torch.float32
        K = arch.nsample
        self.K = K
        # ---- grouping workspace (all T*K rows per section: unique_rows = 0)
        self.rows = [torch.empty((B * T[s] * K[s], 4), dtype=f32, device=dev) for s in range(S)]
        self.cnt = [torch.empty((B, T[s]), dtype=torch.int32, device=dev) for s in range(S)]
        self.tiles = [torch.empty((B * ((T[s] * K[s] + 63) // 64) + 1, 4), dtype=torch.int32, device=dev) for s in range(S)]
        self.idx32 = [torch.empty((B, T[s], K[s]), dtype=torch.int32, device=dev) for s in range(S)]
        self.ntiles = torch.zeros(_lib.MAX_SCALES, dtype=torch.int32, device=dev)
        # ---- layer list: (name, ctypes layer, out key); activations keyed by name
        self.layers: List[_lib.TrainLayer] = []
        self.names: List[str] = []
        self.dx_mask: List[int] = []
        self.act: Dict[str, dict] = {}
        self._zero_fwd: List[torch.Tensor] = []     # fp64 Sum(y), Sum(y^2)
        self._zero_bwd: List[torch.Tensor] = []     # fp64 Sum(dz), Sum(dz*xh) and all dA / dfeat buffers
        self._keep = []

        def new_act(name, rows, cols, T_pos, up=1, cout=None, bn=None, relu=1, grad=True):
            a = dict(rows=rows, cols=cols, T=T_pos, up=up, cout=cout or cols, relu=relu)
            a["Y"] = torch.empty((rows, cols), dtype=f32, device=dev)
            a["dA"] = torch.zeros((rows, cols), dtype=f32, device=dev) if grad else None
            if grad:
                self._zero_bwd.append(a["dA"])
            if bn is not None:
                a["bn"] = bn
                a["sums"] = torch.zeros(2 * a["cout"], dtype=torch.float64, device=dev)
                a["dsums"] = torch.zeros(2 * a["cout"], dtype=torch.float64, device=dev)
                self._zero_fwd.append(a["sums"])
                self._zero_bwd.append(a["dsums"])
            self.act[name] = a
            return a

        def src_of(name, c0=0):
            a = self.act[name]
            s = _lib.TrainSrc()
            s.raw, s.grad = _ptr(a["Y"]), _ptr(a.get("dA"))
            bn = a.get("bn")
            if bn is not None:
                s.sums, s.gamma, s.beta = _ptr(a["sums"]), _ptr(bn.weight), _ptr(bn.bias)
                s.count = float(a["rows"] * a["up"])
                s.Cstat, s.coff, s.relu = a["cout"], 0, a["relu"]
            else:
                s.sums = None
                s.count, s.Cstat, s.coff, s.relu = 1.0, 1, 0, 0
            s.ld, s.T, s.up, s.cup, s.c0 = a["cols"], a["T"], a["up"], a["cout"], c0
            return s

        def add_layer(name, out, segs, conv, bn, N_cols, cout, up, s_co, s_j, T_out, bias=None, dx=True):
            L = _lib.TrainLayer()
            L.B, L.T_out, L.n_seg, L.N, L.Cout, L.up = B, T_out, len(segs), N_cols, cout, up
            L.s_co, L.s_j, L.has_bn, L.relu = s_co, s_j, (1 if bn is not None else 0), (1 if bn is not None else 0)
            mask = 0
            for i, (srcname, Cc, tap, stride, s_ci, w_off) in enumerate(segs):
                g = L.seg[i]
                g.src = src_of(srcname)
                g.C, g.tap, g.stride, g.s_ci, g.w_off = Cc, tap, stride, s_ci, w_off
                if dx and self.act[srcname].get("dA") is not None:
                    mask |= 1 << i
            a = self.act[out]
            L.W, L.dW = _ptr(conv.weight), self.flat.gptr(conv.weight)
            if bias is not None:
                L.bias, L.dbias = _ptr(bias), self.flat.gptr(bias)
            L.Y, L.dA = _ptr(a["Y"]), _ptr(a["dA"])
            if bn is not None:
                L.sums, L.dsums = _ptr(a["sums"]), _ptr(a["dsums"])
                L.gamma, L.beta = _ptr(bn.weight), _ptr(bn.bias)
                L.dgamma, L.dbeta = self.flat.gptr(bn.weight), self.flat.gptr(bn.bias)
                L.run_mean, L.run_var = _ptr(bn.running_mean), _ptr(bn.running_var)
            else:
                a["dsums"] = torch.zeros(2 * N_cols, dtype=torch.float64, device=dev)
                self._zero_bwd.append(a["dsums"])
                L.dsums = _ptr(a["dsums"])
            self.layers.append(L)
            self.names.append(name)
            self.dx_mask.append(mask)

        # ---- PointNet scales
        self.pools = []
        fn = model.feat_net
        for s in range(S):
            c1, c2, c3 = arch.mlps[s]
            pm = getattr(fn, "pointnet%d" % (s + 1))
            TK = T[s] * K[s]
            rname = "rows%d" % (s + 1)
            self.act[rname] = dict(rows=B * TK, cols=4, T=TK, up=1, cout=4, relu=0, Y=self.rows[s], dA=None)
            prev, pc = rname, 3
            for j, co in enumerate((c1, c2, c3)):
                blk = getattr(pm, "conv%d" % (j + 1))
                nm = "pn%d_%d" % (s + 1, j + 1)
                new_act(nm, B * TK, co, TK, bn=blk[1])
                add_layer(nm, nm, [(prev, pc, 0, 1, 1, 0)], blk[0], blk[1], co, co, 1, pc, 0, TK, dx=(j > 0))
                prev, pc = nm, co
            # pooled feature (+ one-hot columns): identity source for the FCN
            fname = "feat%d" % (s + 1)
            fa = dict(rows=B * T[s], cols=c3 + self.V, T=T[s], up=1, cout=c3 + self.V, relu=0)
            fa["Y"] = torch.zeros((B * T[s], c3 + self.V), dtype=f32, device=dev)
            fa["dA"] = torch.zeros((B * T[s], c3 + self.V), dtype=f32, device=dev)
            self._zero_bwd.append(fa["dA"])
            self.act[fname] = fa
            P = _lib.TrainPool()
            last = self.act[prev]
            P.B, P.T, P.K, P.C, P.V, P.ld_feat = B, T[s], K[s], c3, self.V, c3 + self.V
            P.Y, P.cnt, P.sums = _ptr(last["Y"]), _ptr(self.cnt[s]), _ptr(last["sums"])
            P.gamma, P.beta = _ptr(last["bn"].weight), _ptr(last["bn"].bias)
            argmax = torch.zeros((B * T[s], c3), dtype=torch.int32, device=dev)
            self._keep.append(argmax)
            P.feat, P.argmax, P.dfeat, P.dA = _ptr(fa["Y"]), _ptr(argmax), _ptr(fa["dA"]), _ptr(last["dA"])
            self.pools.append(P)
        self.n_pn_layers = len(self.layers)
        # ---- ConvFeatNet
        cn = model.conv_net
        c3s = [m[2] for m in arch.mlps]
        widths = (128, 256, 512, 512)[: S - 1]
        Tl = list(T)
        for i in range(1, S):
            assert Tl[i] == (Tl[i - 1] + 1) // 2, "section counts are not a /2 pyramid"

        def conv3(name, src, ci, co, stride, T_in):
            blk = getattr(cn, name)
            T_out = T_in if stride == 1 else (T_in + 1) // 2
            new_act(name, B * T_out, co, T_out, bn=blk[1])
            segs = [(src, ci, j - 1, stride, 3, j) for j in range(3)]
            add_layer(name, name, segs, blk[0], blk[1], co, co, 1, ci * 3, 0, T_out)
            return T_out

        conv3("block1_conv1", "feat1", c3s[0] + self.V, arch.block1_out, 1, Tl[0])
        prev, pc = "block1_conv1", arch.block1_out
        for i in range(2, S + 1):
            w = widths[i - 2]
            conv3("block%d_conv1" % i, prev, pc, w, 2, Tl[i - 2])
            conv3("block%d_conv2" % i, "block%d_conv1" % i, w, w, 1, Tl[i - 1])
            nm = "block%d_merge" % i
            blk = getattr(cn, nm)
            cb = c3s[i - 1] + self.V
            new_act(nm, B * Tl[i - 1], w, Tl[i - 1], bn=blk[1])
            add_layer(nm, nm, [("block%d_conv2" % i, w, 0, 1, 1, 0), ("feat%d" % i, cb, 0, 1, 1, w)], blk[0], blk[1],
                      w, w, 1, w + cb, 0, Tl[i - 1])
            prev, pc = nm, w
        for i in range(2, S + 1):
            nm = "block%d_deconv" % i
            blk = getattr(cn, nm)
            k = 2 ** (i - 2)
            ci = widths[i - 2]
            new_act(nm, B * Tl[i - 1], k * 256, Tl[i - 1], up=k, cout=256, bn=blk[1])
            add_layer(nm, nm, [("block%d_merge" % i, ci, 0, 1, 256 * k, 0)], blk[0], blk[1], k * 256, 256, k, k, 1,
                      Tl[i - 1])
        # ---- heads on the cropped concat of the three (four) transposed-conv outputs (det_base.py:222-224,367-368)
        T2 = Tl[1]
        cin = 256 * (S - 1)
        hsegs = [("block%d_deconv" % i, 256, 0, 1, 1, 256 * (i - 2)) for i in range(2, S + 1)]
        for nm, conv in (("cls_out", model.cls_out), ("reg_out", model.reg_out)):
            co = conv.weight.shape[0]
            new_act(nm, B * T2, co, T2, bn=None)
            add_layer(nm, nm, hsegs, conv, None, co, co, 1, cin, 0, T2, bias=conv.bias)
        self.T2 = T2
        # device copy of the layer table for the end-of-step kernel
        arr = (_lib.TrainLayer * len(self.layers))(*self.layers)
        self._table_host = arr
        self.table_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self._bn_counters = [b for n, b in model.named_buffers() if n.endswith("num_batches_tracked")]
        # static inputs (the launch sequences are captured into CUDA graphs)
        self.in_pc = torch.zeros((B, 3, N), dtype=f32, device=dev)
        self.in_centers = [torch.zeros((B, 3, T[s]), dtype=f32, device=dev) for s in range(S)]
        self.in_onehot = torch.zeros((B, max(self.V, 1)), dtype=f32, device=dev)
        for P in self.pools:
            P.one_hot = _ptr(self.in_onehot) if self.V > 0 else None
        self.group_args = self._group_args()
        self._graphs = {}
        self.use_graph = True
        lib = _lib.load()
        need = max([int(lib.fcn_train_workspace_floats(C.byref(L))) for L in self.layers] + [1])
        self.workspace = torch.empty(need, dtype=f32, device=dev)   # K-split partial sums of the skinny FCN layers

    # ------------------------------------------------------------------
    def _group_args(self):
        g = _lib.GroupArgs()
        arch = self.model.ARCH
        S = arch.num_scales
        g.num_scales, g.B, g.N, g.num_vec, g.tile_rows, g.unique_rows = S, self.B, self.N, 0, 64, 0
        dists = self.model.feat_net.dists
        for s in range(S):
            g.T[s], g.K[s], g.dis_z[s] = self.T[s], self.K[s], float(dists[s])
            g.c3[s], g.ld_feat[s] = arch.mlps[s][2], 0
            g.row_cap[s] = self.T[s] * self.K[s]
            g.tile_cap[s] = self.tiles[s].shape[0]
            g.rows[s], g.cnt[s], g.feat[s], g.tiles[s] = _ptr(self.rows[s]), _ptr(self.cnt[s]), None, _ptr(self.tiles[s])
            g.idx_scratch[s] = _ptr(self.idx32[s])
            g.feat_pitch[s] = 0
        g.ntiles = _ptr(self.ntiles)
        g.force_scan = 0
        g.pc, g.one_hot = _ptr(self.in_pc), None
        for s in range(S):
            g.centers[s] = _ptr(self.in_centers[s])
        return g

    # ---- launch sequences (all pointers are static: the step is CUDA-graph capturable)
    def _seq_forward(self):
        S = len(self.T)
        st = _stream()
        torch._foreach_zero_(self._zero_fwd)
        _lib.call("fcn_group_rows", C.byref(self.group_args), st)
        li = 0
        for s in range(S):
            for _ in range(3):
                _lib.call("fcn_train_forward", C.byref(self.layers[li]), _ptr(self.workspace), self.workspace.numel(), st)
                li += 1
            _lib.call("fcn_train_pool", C.byref(self.pools[s]), 0, st)
        while li < len(self.layers):
            _lib.call("fcn_train_forward", C.byref(self.layers[li]), _ptr(self.workspace), self.workspace.numel(), st)
            li += 1

    def _seq_backward(self, stage, update_running):
        st = _stream()
        S = len(self.T)
        if stage in ("all", "fcn"):
            # FCN + heads in reverse creation order (a topological order of the backward graph: every consumer of
            # a tensor was created after it)
            for li in range(len(self.layers) - 1, self.n_pn_layers - 1, -1):
                _lib.call("fcn_train_backward", C.byref(self.layers[li]), self.dx_mask[li], st)
            # BN bookkeeping of the FCN layers: their dgamma / dbeta belong to the first gradient bucket
            nf = len(self.layers) - self.n_pn_layers
            _lib.call("fcn_train_finalize", _ptr(self.table_dev) + self.n_pn_layers * C.sizeof(_lib.TrainLayer), nf,
                      1 if update_running else 0, st)
        if stage in ("all", "pointnet"):
            for s in range(S - 1, -1, -1):
                _lib.call("fcn_train_pool", C.byref(self.pools[s]), 1, st)
                for j in (2, 1, 0):
                    li = 3 * s + j
                    _lib.call("fcn_train_backward", C.byref(self.layers[li]), self.dx_mask[li], st)
            _lib.call("fcn_train_finalize", _ptr(self.table_dev), self.n_pn_layers, 1 if update_running else 0, st)
            if update_running and self._bn_counters:
                torch._foreach_add_(self._bn_counters, 1)

    def _graph(self, key, fn):
        """Capture `fn` once (after an eager warm-up run, whose effects are the first execution) and replay it."""
        g = self._graphs.get(key)
        if g is None:
            fn()                                   # eager: this call's work
            try:
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    fn()
                self._graphs[key] = g
            except Exception as e:                 # stay functional without graphs
                import sys
                sys.stderr.write("TrainEngine: CUDA-graph capture of %r failed (%r); launching eagerly\n" % (key, e))
                self._graphs[key] = False
                torch.cuda.synchronize()
            return
        if g is False:
            fn()
        else:
            g.replay()

    @torch.no_grad()
    def forward(self, pc, centers, one_hot):
        """-> (cls logits (B*T2, 2), reg logits (B*T2, out)) as views of the engine's buffers."""
        assert tuple(pc.shape) == (self.B, 3, self.N) and pc.dtype == torch.float32
        self.in_pc.copy_(pc)
        for s, c in enumerate(centers):
            assert tuple(c.shape) == (self.B, 3, self.T[s])
            self.in_centers[s].copy_(c)
        if self.V > 0:
            self.in_onehot.copy_(one_hot)
        if self.use_graph:
            self._graph("fwd", self._seq_forward)
        else:
            self._seq_forward()
        return self.act["cls_out"]["Y"], self.act["reg_out"]["Y"]

    @torch.no_grad()
    def backward(self, dcls, dreg, update_running=True, stage="all"):
        """dlogits -> gradients of every parameter, ACCUMULATED into the flat bucket ``self.flat.grad``.
        stage = "fcn" (heads + ConvFeatNet: everything the feat_net gradients depend on), "pointnet" (the rest +
        BN bookkeeping) or "all"."""
        if stage in ("all", "fcn"):
            torch._foreach_zero_(self._zero_bwd)
            self.act["cls_out"]["dA"].copy_(dcls.reshape(self.act["cls_out"]["dA"].shape))
            self.act["reg_out"]["dA"].copy_(dreg.reshape(self.act["reg_out"]["dA"].shape))
        if self.use_graph:
            self._graph(("bwd", stage, bool(update_running)), lambda: self._seq_backward(stage, update_running))
        else:
            self._seq_backward(stage, update_running)

    def kernel_launches_per_step(self):
        n_fwd = len(self.layers) + len(self.pools) + 2
        n_bwd = sum(1 + L.n_seg + bin(m).count("1") for L, m in zip(self.layers, self.dx_mask)) + len(self.pools) + 1
        return n_fwd + n_bwd + 1      # + fused Adam


class TrainStep:
    """One optimizer step of PointNetDet on the hand-written kernels: forward -> losses (PyTorch ops on the logits)
    -> backward kernels -> ONE all-reduce of the flat gradient bucket (when torch.distributed is initialised and
    world > 1) -> fused Adam.  Mirrors train/train_net_det.py:121-128 (zero_grad / backward / step)."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, comm_stream=None):
        self.model = model
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.engines: Dict[tuple, TrainEngine] = {}
        self.step_count = 0
        self.flat = None
        self.m = self.v = None
        self._comm = None
        self._split = 0
        self._loss_graphs = {}
        # losses + d/dlogits: "kernel" = fused hand-written kernels (csrc/loss.cu), "graph" = the PyTorch ops as one
        # CUDA graph (train_path.LossGraph), "eager" = the reference-shaped PyTorch ops
        self.loss_impl = "kernel"

    def engine(self, B, N, T) -> TrainEngine:
        key = (int(B), int(N), tuple(int(t) for t in T))
        e = self.engines.get(key)
        if e is None:
            e = self.engines[key] = TrainEngine(self.model, *key)
            if self.flat is None:
                self.flat = e.flat
                self.m = torch.zeros_like(self.flat.param)
                self.v = torch.zeros_like(self.flat.param)
        return e

    def _losses(self, eng, cls, reg, center_ref2, data):
        """-> (losses, metrics, dcls, dreg)."""
        from .train_path import FusedLoss, LossGraph, losses_from_logits
        if self.loss_impl in ("kernel", "graph"):
            key = (self.loss_impl, eng.B, eng.T2)
            lg = self._loss_graphs.get(key)
            if lg is None:
                cls_ = FusedLoss if self.loss_impl == "kernel" else LossGraph
                lg = self._loss_graphs[key] = cls_(self.model, eng.B, eng.T2, reg.shape[1], cls.device)
            return lg.run(cls, reg, center_ref2, data)
        cls_l = cls.detach().clone().requires_grad_(True)
        reg_l = reg.detach().clone().requires_grad_(True)
        losses, metrics = losses_from_logits(self.model, cls_l, reg_l, center_ref2, data)
        losses["total_loss"].backward()
        return losses, metrics, cls_l.grad, reg_l.grad

    def forward_backward(self, data, update_running=True):
        """-> (losses, metrics); gradients (this rank's) are left in the flat bucket."""
        from .train_path import losses_from_logits
        model = self.model
        S = model.ARCH.num_scales
        pc = data["point_cloud"][:, :3, :].contiguous()
        centers = [data["center_ref%d" % (i + 1)].contiguous() for i in range(S)]
        eng = self.engine(pc.shape[0], pc.shape[2], [c.shape[2] for c in centers])
        self.flat.grad.zero_()
        cls, reg = eng.forward(pc, centers, data.get("one_hot"))
        losses, metrics, dcls, dreg = self._losses(eng, cls, reg, centers[1], data)
        eng.backward(dcls, dreg, update_running=update_running)
        self.flat.expose_grads(model)
        return losses, metrics

    def step(self, data):
        """forward + losses + backward + gradient all-reduce + Adam.  With world > 1 the flat bucket is reduced in
        TWO pieces on a communication stream: the FCN + heads part (92 % of the parameters, complete after the
        first ~15 % of the backward) overlaps the PointNet backward; the PointNet part follows."""
        import torch.distributed as dist
        from .train_path import losses_from_logits
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        model = self.model
        S = model.ARCH.num_scales
        pc = data["point_cloud"][:, :3, :].contiguous()
        centers = [data["center_ref%d" % (i + 1)].contiguous() for i in range(S)]
        eng = self.engine(pc.shape[0], pc.shape[2], [c.shape[2] for c in centers])
        self.flat.grad.zero_()
        cls, reg = eng.forward(pc, centers, data.get("one_hot"))
        losses, metrics, dcls, dreg = self._losses(eng, cls, reg, centers[1], data)
        if world > 1:
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=self.flat.grad.device)
                first = next(model.conv_net.parameters())
                self._split = self.flat.offsets[id(first)]        # [0, split): feat_net, [split, n): conv_net + heads
            cur = torch.cuda.current_stream()
            eng.backward(dcls, dreg, stage="fcn")
            ev = torch.cuda.Event()
            ev.record(cur)
            self._comm.wait_event(ev)
            with torch.cuda.stream(self._comm):
                dist.all_reduce(self.flat.grad[self._split:])
            eng.backward(None, None, stage="pointnet")
            ev2 = torch.cuda.Event()
            ev2.record(cur)
            self._comm.wait_event(ev2)
            with torch.cuda.stream(self._comm):
                dist.all_reduce(self.flat.grad[: self._split])
            cur.wait_stream(self._comm)
        else:
            eng.backward(dcls, dreg)
        self.step_count += 1
        self.model.refresh()       # parameters change behind autograd's version counters: eval pack must rebuild
        _lib.call("fcn_adam_step", _ptr(self.flat.param), _ptr(self.flat.grad), _ptr(self.m), _ptr(self.v),
                  self.flat.numel, float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                  float(self.wd), int(self.step_count), 1.0 / world, _stream())
        return losses, metrics
