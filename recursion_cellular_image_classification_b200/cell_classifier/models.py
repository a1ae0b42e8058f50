"""Model surface of the hot path (mirrors reference cell_classifier/models.py).

`DenseNet121` is the north star's trunk: torchvision's densenet121 with the reference's 6-channel stem
(models.py:17-27), executed by the native librxb executor (csrc/densenet.cu) — tcgen05 implicit-GEMM
convolutions, fused BatchNorm/ReLU/concat, native backward and SGD.  Parameters live in ONE flat fp32
tensor in torchvision's named_parameters() order; `state_dict()` / `load_state_dict()` speak torchvision's
names so checkpoints interchange with `torchvision.models.densenet121`.

`TwoSitesNN` keeps the reference's constructor and call signature (models.py:8-12, 41):
x[B, G, 6, H, W] -> [B, nb_classes].  The reference splits the G images of a sample into thirds — the sample's own
sites, the negative control's, the positive control's (models.py:46-49) — averages each third in feature space and
concatenates the three means into its MLP (:50-55).  DenseNet-121 has a single-image trunk with ONE linear classifier,
so only the first third (the sample's own sites) can reach it: `sample_group(G)` says how many leading images that
is, their features are averaged (= averaging their logits, the head being linear), and the control thirds are
neither decoded, copied nor run (dataloader.raw_item(controls=False)).  train(), evaluate() and test() all use this
one rule.  `DummyClassifier` is the reference's fake backend (models.py:60-68).
"""
import ctypes
import math
from collections import OrderedDict

import torch

from .. import _lib, ops
from .._lib import Dn121Config, Rn50Config, check, load, ptr, stream_ptr

_BLOCKS = (6, 12, 24, 16)


def densenet121_param_specs(nb_classes=1108):
    """[(torchvision name, shape)] in named_parameters() order, and the BN buffer list in module order."""
    specs, bufs = [], []

    def bn(prefix, c):
        specs.append((prefix + ".weight", (c,)))
        specs.append((prefix + ".bias", (c,)))
        bufs.append((prefix + ".running_mean", (c,)))
        bufs.append((prefix + ".running_var", (c,)))

    specs.append(("features.conv0.weight", (64, 6, 7, 7)))
    bn("features.norm0", 64)
    c = 64
    for b, n_layers in enumerate(_BLOCKS, 1):
        for i in range(1, n_layers + 1):
            p = "features.denseblock%d.denselayer%d" % (b, i)
            bn(p + ".norm1", c)
            specs.append((p + ".conv1.weight", (128, c, 1, 1)))
            bn(p + ".norm2", 128)
            specs.append((p + ".conv2.weight", (32, 128, 3, 3)))
            c += 32
        if b < 4:
            p = "features.transition%d" % b
            bn(p + ".norm", c)
            specs.append((p + ".conv.weight", (c // 2, c, 1, 1)))
            c //= 2
    bn("features.norm5", c)
    specs.append(("classifier.weight", (nb_classes, c)))
    specs.append(("classifier.bias", (nb_classes,)))
    return specs, bufs


def sample_group(G):
    """How many of the G images of an item are the sample's own sites: the first third when the item carries the
    reference's image / negative-control / positive-control thirds (models.py:45-49: shape = int(G/3)), else all."""
    return G // 3 if G >= 3 and G % 3 == 0 else G


def to_s2d32(x_nchw):
    """f32/bf16 [B,6,H,W] (already normalised) -> bf16 [B,H/2,W/2,32], the stem conv's input layout
    (channel = (y&1)*16 + (x&1)*8 + c).  Compatibility path only: the fused loader writes this directly."""
    B, C, H, W = x_nchw.shape
    x = torch.zeros(B, H, W, 8, dtype=torch.bfloat16, device=x_nchw.device)
    x[..., :C] = x_nchw.permute(0, 2, 3, 1)
    x = x.view(B, H // 2, 2, W // 2, 2, 8).permute(0, 1, 3, 2, 4, 5).reshape(B, H // 2, W // 2, 32)
    return x.contiguous()


class DenseNet121(torch.nn.Module):
    wants_controls = False         # a single-image trunk: test() feeds it the sample's own sites only

    def __init__(self, nb_classes=1108, device=None, bn_eps=1e-5, bn_momentum=0.1, seed=None):
        super().__init__()
        if device is None:       # buffers live where compute will run; without a GPU only the host-side surface works
            from ..parallel import default_device
            device = default_device()
        self.nb_classes = nb_classes
        self.bn_eps, self.bn_momentum = bn_eps, bn_momentum
        self.specs, self.buf_specs = densenet121_param_specs(nb_classes)
        n = sum(math.prod(s) for _, s in self.specs)
        nb = sum(math.prod(s) for _, s in self.buf_specs)
        dev = torch.device(device)
        self.flat = torch.nn.Parameter(torch.zeros(n, dtype=torch.float32, device=dev))
        self.flat.grad = torch.zeros_like(self.flat)
        self.register_buffer("momentum_buf", torch.zeros(n, dtype=torch.float32, device=dev))
        self.register_buffer("bn_buffers", torch.zeros(nb, dtype=torch.float32, device=dev))
        self._plans = {}
        self.graph_launches = 0
        self._views = OrderedDict()
        off = 0
        for name, shape in self.specs:
            k = math.prod(shape)
            self._views[name] = (off, k, shape)
            off += k
        self._bviews = OrderedDict()
        off = 0
        for name, shape in self.buf_specs:
            k = math.prod(shape)
            self._bviews[name] = (off, k, shape)
            off += k
        self.reset_parameters(seed)

    # ---------------------------------------------------------------- parameters
    def view(self, name):
        off, k, shape = self._views[name]
        return self.flat.data[off:off + k].view(shape)

    def grad_view(self, name):
        off, k, shape = self._views[name]
        return self.flat.grad[off:off + k].view(shape)

    def buffer_view(self, name):
        off, k, shape = self._bviews[name]
        return self.bn_buffers[off:off + k].view(shape)

    def reset_parameters(self, seed=None):
        """torchvision densenet init: kaiming_normal conv, BN weight 1 / bias 0, classifier bias 0; the stem is
        the channel-mean of a 3-channel kaiming stem replicated 6x (reference models.py:24-26)."""
        g = torch.Generator().manual_seed(0 if seed is None else seed)
        with torch.no_grad():
            for name, shape in self.specs:
                v = self.view(name)
                if name == "features.conv0.weight":
                    w3 = torch.empty(64, 3, 7, 7)
                    torch.nn.init.kaiming_normal_(w3, generator=g)
                    v.copy_(w3.mean(1, keepdim=True).expand(64, 6, 7, 7))
                elif name.endswith("conv.weight") or "conv1.weight" in name or "conv2.weight" in name:
                    w = torch.empty(shape)
                    torch.nn.init.kaiming_normal_(w, generator=g)
                    v.copy_(w)
                elif name == "classifier.weight":
                    w = torch.empty(shape)
                    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5), generator=g)
                    v.copy_(w)
                elif name.endswith(".weight"):
                    v.fill_(1.0)
                else:
                    v.zero_()
            for name, _ in self.buf_specs:
                self.buffer_view(name).fill_(1.0 if name.endswith("running_var") else 0.0)
        self._weights_dirty = True

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        """torchvision densenet121 names (clones of the flat buffers' views).  Honours nn.Module's destination / prefix
        arguments, so a wrapper's state_dict() (torch.nn.DataParallel, main.py:94 and train.py:96) yields
        `module.`-prefixed keys like the reference's checkpoints."""
        if args:                                                     # legacy positional form
            destination, prefix, keep_vars = (list(args) + [prefix, keep_vars])[:3] if len(args) < 3 else args[:3]
        sd = OrderedDict() if destination is None else destination
        for name in self._views:
            sd[prefix + name] = self.view(name).clone()
        for name in self._bviews:
            sd[prefix + name] = self.buffer_view(name).clone()
        return sd

    def _copy_from(self, sd, prefix, strict, missing):
        with torch.no_grad():
            for table, getter in ((self._views, self.view), (self._bviews, self.buffer_view)):
                for name in table:
                    key = prefix + name
                    if key not in sd and not prefix and "module." + name in sd:
                        key = "module." + name                       # DataParallel-prefixed checkpoint (train.py:96)
                    if key in sd:
                        v = sd[key]
                        if name == "features.conv0.weight" and v.dim() == 4 and v.shape[1] == 3:
                            # a stock 3-channel (ImageNet) stem: the reference's surgery, models.py:24-26
                            v = torch.stack([torch.mean(v, 1)] * 6, dim=1)
                        if name.startswith("classifier.") and tuple(v.shape) != tuple(getter(name).shape) and not strict:
                            continue                                 # another label set (ImageNet's 1000): keep our head
                        getter(name).copy_(v)
                    elif strict:
                        missing.append(key)
        self._weights_dirty = True

    def load_state_dict(self, sd, strict=True):
        missing = []
        self._copy_from(sd, "", strict, missing)
        if missing:
            raise KeyError(missing[0])

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # reached when a WRAPPER loads (main.py:147: DataParallel(model).load_state_dict(torch.load(...))): the keys
        # carry the wrapper's prefix and the torchvision names, not this module's flat buffers
        self._copy_from(state_dict, prefix, strict, missing_keys)

    # ---------------------------------------------------------------- plans
    def _plan(self, B, H, W, training):
        key = (B, H, W, bool(training))
        if key in self._plans:
            return self._plans[key]
        _lib.require_gpu()
        lib = load()
        # one live plan per mode: a plan owns a full workspace (12.8 GB at B=128, 512x512), so the odd last batch of
        # an epoch or another evaluation batch size replaces the previous plan of that mode instead of piling up
        for old in [k for k in self._plans if k[3] == bool(training)]:
            lib.rxb_dn121_destroy(self._plans.pop(old)["handle"])
        cfg = Dn121Config(B, H, W, self.nb_classes, self.bn_eps, self.bn_momentum)
        assert lib.rxb_dn121_param_count(ctypes.byref(cfg)) == self.flat.numel()
        assert lib.rxb_dn121_buffer_count(ctypes.byref(cfg)) == self.bn_buffers.numel()
        nbytes = lib.rxb_dn121_workspace_bytes(ctypes.byref(cfg), 1 if training else 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.flat.device)
        handle = ctypes.c_void_p()
        check(lib.rxb_dn121_create(ctypes.byref(cfg), ptr(self.flat.data), ptr(self.flat.grad), ptr(self.momentum_buf),
                                   ptr(self.bn_buffers), ptr(ws), nbytes, 1 if training else 0, ctypes.byref(handle)))
        plan = {"handle": handle, "ws": ws, "cfg": cfg, "synced": False, "graphs": {}, "static": None}
        self._plans[key] = plan
        return plan

    def static_buffers(self, B, H, W):
        """The training plan's fixed-address step buffers — (xs bf16 [B,H/2,W/2,32], target int64 [B], loss f32 [1]).
        Steps whose inputs live here can be replayed from CUDA graphs (`train_step(..., graph=True)`): have the fused
        loader write `xs` directly (ops.load_norm_aug(..., out=xs)) and copy the labels into `target`."""
        plan = self._plan(B, H, W, True)
        if plan["static"] is None:
            dev = self.flat.device
            plan["static"] = (torch.empty(B, H // 2, W // 2, 32, dtype=torch.bfloat16, device=dev),
                              torch.zeros(B, dtype=torch.int64, device=dev),
                              torch.zeros(1, dtype=torch.float32, device=dev))
        return plan["static"]

    def _sync(self, plan):
        if self._weights_dirty:
            for p in self._plans.values():
                p["synced"] = False
            self._weights_dirty = False
        if not plan["synced"]:
            check(load().rxb_dn121_sync_weights(plan["handle"], stream_ptr()))
            plan["synced"] = True

    def __del__(self):
        try:
            lib = load()
            for p in self._plans.values():
                lib.rxb_dn121_destroy(p["handle"])
        except Exception:
            pass

    # ---------------------------------------------------------------- compute
    def _as_s2d(self, x):
        if x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[-1] == 32:
            return x.contiguous()
        return to_s2d32(x.to(self.flat.device))

    def forward(self, x):
        """x: bf16 S2D32 [B,H/2,W/2,32] (fused-loader output) or float [B,6,H,W] / [B,G,6,H,W] (the reference's item
        layout: only the sample's own sites, the first third of G, are run and averaged — see the module docstring)."""
        groups = None
        if x.dim() == 5:
            groups = sample_group(x.shape[1])
            x = x[:, :groups].reshape(-1, *x.shape[2:])
        xs = self._as_s2d(x)
        B, H, W = xs.shape[0], xs.shape[1] * 2, xs.shape[2] * 2
        plan = self._plan(B, H, W, False)
        self._sync(plan)
        logits = torch.empty(B, self.nb_classes, dtype=torch.float32, device=xs.device)
        check(load().rxb_dn121_forward(plan["handle"], ptr(xs), ptr(logits), 1 if self.training else 0, stream_ptr()))
        if groups is not None:
            logits = logits.view(-1, groups, self.nb_classes).mean(1)
        return logits

    def train_step(self, xs, target, global_batch=None, phase=-1, loss_out=None, graph=False):
        """forward + CrossEntropy(mean over global_batch) + backward into self.flat.grad.  Returns the
        device scalar holding this rank's share of the loss.
        graph=True replays the phase from a CUDA graph: the executor enqueues ~340 kernels per step with no host
        synchronisation, so a phase is captured once per (buffers, phase, global batch) — on its second use, the first
        one runs eagerly and loads every kernel — and replayed afterwards (about 4 % of the step at batch 128).  The
        inputs must then sit at fixed addresses: use `static_buffers()`; other tensors are copied into them."""
        xs = self._as_s2d(xs)
        B, H, W = xs.shape[0], xs.shape[1] * 2, xs.shape[2] * 2
        plan = self._plan(B, H, W, True)
        self._sync(plan)
        if not graph:
            if loss_out is None:
                loss_out = torch.empty(1, dtype=torch.float32, device=xs.device)
            check(load().rxb_dn121_train_step(plan["handle"], ptr(xs), ptr(target.contiguous()), global_batch or B,
                                              ptr(loss_out), phase, stream_ptr()))
            return loss_out
        sx, sy, sl = self.static_buffers(B, H, W)
        if xs.data_ptr() != sx.data_ptr() and phase in (-1, 0):
            sx.copy_(xs)
        if target.data_ptr() != sy.data_ptr() and phase in (-1, 0):
            sy.copy_(target)
        key = (phase, global_batch or B)
        entry = plan["graphs"].get(key)
        lib = load()
        if entry is None:                                  # first use: eager (doubles as the warm-up a capture needs)
            check(lib.rxb_dn121_train_step(plan["handle"], ptr(sx), ptr(sy), global_batch or B, ptr(sl), phase, stream_ptr()))
            plan["graphs"][key] = "warm"
        else:
            if entry == "warm":
                torch.cuda.synchronize(self.flat.device)
                g = torch.cuda.CUDAGraph()
                n0 = lib.rxb_launch_count()
                with torch.cuda.graph(g):
                    check(lib.rxb_dn121_train_step(plan["handle"], ptr(sx), ptr(sy), global_batch or B, ptr(sl), phase,
                                                   stream_ptr()))
                entry = plan["graphs"][key] = (g, lib.rxb_launch_count() - n0)
            entry[0].replay()
            self.graph_launches += entry[1]               # kernels replayed (rxb_launch_count only sees enqueues)
        if loss_out is not None and loss_out.data_ptr() != sl.data_ptr() and phase in (-1, 0):
            loss_out.copy_(sl)
            return loss_out
        return sl

    def phase_grad_range(self, B, H, W, phase):
        plan = self._plan(B, H, W, True)
        b, e = ctypes.c_int64(), ctypes.c_int64()
        check(load().rxb_dn121_phase_grad_range(plan["handle"], phase, ctypes.byref(b), ctypes.byref(e)))
        return b.value, e.value

    def head_range(self):
        """[begin, end) of the classifier (the last two tensors) in the flat buffers — the part that stays trainable
        while a pretrained trunk is frozen (train.py:46-58: children named 'mlp' or 'classifier')."""
        off, _, _ = self._views["classifier.weight"]
        return off, self.flat.numel()

    def sgd_step(self, B, H, W, lr, momentum=0.9, weight_decay=3e-5, nesterov=True, grad_scale=1.0, head_only=False):
        """torch.optim.SGD semantics (main.py:89-93) on the flat buffers, then refresh the bf16 operands.
        head_only: update the classifier alone, as if every other parameter had requires_grad=False."""
        if head_only:
            b, e = self.head_range()
            ops.sgd_step_(self.flat.data[b:e], self.flat.grad[b:e], self.momentum_buf[b:e], lr, momentum, weight_decay,
                          nesterov, grad_scale)
            self._weights_dirty = True
            return
        plan = self._plan(B, H, W, True)
        check(load().rxb_dn121_sgd(plan["handle"], lr, momentum, weight_decay, 1 if nesterov else 0, grad_scale,
                                   stream_ptr()))
        for p in self._plans.values():
            p["synced"] = p is plan


def _pretrained_densenet121_state():
    """ImageNet weights for the trunk, or None: a state_dict file named by RXB_PRETRAINED_DENSENET121 (torchvision
    densenet121 key names), else torchvision's own download/cache (needs network or a warm cache)."""
    import os
    path = os.environ.get("RXB_PRETRAINED_DENSENET121")
    if path:
        return torch.load(path, map_location="cpu")
    try:
        import torchvision
        return torchvision.models.densenet121(weights="IMAGENET1K_V1").state_dict()
    except Exception:
        return None


_RN_LAYERS, _RN_WIDTHS = (3, 4, 6, 3), (64, 128, 256, 512)


def two_sites_resnet50_param_specs(nb_classes=1108, size_features=1024):
    """[(name, shape)] of the reference TwoSitesNN (models.py:8-39) in named_parameters() order, and its BatchNorm
    buffers in module order (a Bottleneck registers bn1, bn2, bn3 and then downsample; running_mean / running_var only,
    the int64 num_batches_tracked entries are host bookkeeping)."""
    specs, bufs = [], []

    def bn_p(prefix, c):
        specs.append((prefix + ".weight", (c,)))
        specs.append((prefix + ".bias", (c,)))

    def bn_b(prefix, c):
        bufs.append((prefix + ".running_mean", (c,)))
        bufs.append((prefix + ".running_var", (c,)))

    specs.append(("base_nn.conv1.weight", (64, 6, 7, 7)))
    bn_p("base_nn.bn1", 64)
    bn_b("base_nn.bn1", 64)
    cin = 64
    for l, (n_blocks, w) in enumerate(zip(_RN_LAYERS, _RN_WIDTHS), 1):
        for i in range(n_blocks):
            p = "base_nn.layer%d.%d" % (l, i)
            specs.append((p + ".conv1.weight", (w, cin, 1, 1)))
            bn_p(p + ".bn1", w)
            specs.append((p + ".conv2.weight", (w, w, 3, 3)))
            bn_p(p + ".bn2", w)
            specs.append((p + ".conv3.weight", (4 * w, w, 1, 1)))
            bn_p(p + ".bn3", 4 * w)
            bn_b(p + ".bn1", w)
            bn_b(p + ".bn2", w)
            bn_b(p + ".bn3", 4 * w)
            if i == 0:
                specs.append((p + ".downsample.0.weight", (4 * w, cin, 1, 1)))
                bn_p(p + ".downsample.1", 4 * w)
                bn_b(p + ".downsample.1", 4 * w)
            cin = 4 * w
    bn_p("mlp.0", 3 * cin)
    bn_b("mlp.0", 3 * cin)
    specs.append(("mlp.2.weight", (size_features, 3 * cin)))
    specs.append(("mlp.2.bias", (size_features,)))
    bn_p("mlp.4", size_features)
    bn_b("mlp.4", size_features)
    specs.append(("mlp.6.weight", (nb_classes, size_features)))
    specs.append(("mlp.6.bias", (nb_classes,)))
    return specs, bufs


def _reference_two_sites_init(pretrained, nb_classes, size_features, dropout, seed):
    """Initial values exactly as the reference constructor produces them (models.py:14-39): torchvision resnet50, the
    6-channel stem from the channel mean of its 3-channel kernel, then the MLP — host-side torch modules used for
    their initialisers only (and, with pretrained=True, for torchvision's ImageNet weights when they can be had)."""
    import torchvision
    if seed is not None:
        torch.manual_seed(seed)
    base, loaded = None, False
    if pretrained:
        try:
            base = torchvision.models.resnet50(weights="IMAGENET1K_V1")
            loaded = True
        except Exception:
            import warnings
            warnings.warn("TwoSitesNN(pretrained=True, trunk='resnet50'): ImageNet resnet50 weights are not available "
                          "(no network / cache) - continuing from random initialisation")
    if base is None:
        base = torchvision.models.resnet50(weights=None)
    kernel = base.conv1.weight.detach()
    torch.nn.Conv2d(6, 64, kernel_size=7, stride=2, padding=3, bias=False)              # models.py:18-23 draws its own
    stem = torch.stack([torch.mean(kernel, 1)] * 6, dim=1)                              # init first; then :24-26
    n_feat = 3 * base.fc.in_features
    mlp = torch.nn.Sequential(torch.nn.BatchNorm1d(n_feat), torch.nn.Dropout(dropout), torch.nn.Linear(n_feat, size_features),
                              torch.nn.ReLU(), torch.nn.BatchNorm1d(size_features), torch.nn.Dropout(dropout),
                              torch.nn.Linear(size_features, nb_classes))
    sd = {"base_nn." + k: v for k, v in base.state_dict().items() if not k.startswith("fc.")}
    sd["base_nn.conv1.weight"] = stem
    sd.update({"mlp." + k: v for k, v in mlp.state_dict().items()})
    return sd, loaded


class TwoSitesResNet50(torch.nn.Module):
    """The reference's real model (models.py:7-57) on the device, evaluation mode: ResNet-50 trunk, feature means of
    the image / negative-control / positive-control thirds of an item concatenated, BatchNorm1d/Dropout/Linear MLP
    head — executed by librxb's rxb_rn50 executor (csrc/resnet.cu, tcgen05 implicit-GEMM convolutions).  state_dict()
    / load_state_dict() speak the reference's own names (`base_nn.*`, `mlp.*`, optionally `module.`-prefixed), so a
    checkpoint written by the reference's train() loads unchanged.  Training runs natively too: `train_step()` is the
    reference's step (BatchNorm batch statistics, Dropout, CrossEntropy, full backward) and `sgd_step()` its optimizer,
    with `head_only=True` for the two-epoch freeze of a pretrained trunk (train.py:46-67); cell_classifier.train.train()
    drives them.  forward() itself is evaluation only: in training mode it raises (there is no autograd graph to hand
    to an external trainer, and no PyTorch fallback)."""

    wants_controls = True          # test() must hand it the full reference item (image + control thirds)

    def __init__(self, pretrained=False, nb_classes=1108, size_features=1024, dropout=0.3, device=None, seed=None,
                 bn_eps=1e-5):
        super().__init__()
        if device is None:
            from ..parallel import default_device
            device = default_device()
        self.nb_classes, self.size_features, self.dropout, self.bn_eps = nb_classes, size_features, dropout, bn_eps
        self.specs, self.buf_specs = two_sites_resnet50_param_specs(nb_classes, size_features)
        n = sum(math.prod(sh) for _, sh in self.specs)
        nb = sum(math.prod(sh) for _, sh in self.buf_specs)
        dev = torch.device(device)
        self.flat = torch.nn.Parameter(torch.zeros(n, dtype=torch.float32, device=dev))
        self.flat.grad = torch.zeros_like(self.flat)
        self.register_buffer("momentum_buf", torch.zeros(n, dtype=torch.float32, device=dev))
        self.register_buffer("bn_buffers", torch.zeros(nb, dtype=torch.float32, device=dev))
        self.bn_momentum = 0.1
        self._views, self._bviews, off = OrderedDict(), OrderedDict(), 0
        for name, shape in self.specs:
            k = math.prod(shape)
            self._views[name] = (off, k, shape)
            off += k
        off = 0
        for name, shape in self.buf_specs:
            k = math.prod(shape)
            self._bviews[name] = (off, k, shape)
            off += k
        self._plans = {}
        self._weights_dirty = True
        sd, self.pretrained_loaded = _reference_two_sites_init(pretrained, nb_classes, size_features, dropout, seed)
        self.load_state_dict(sd)
        self.eval()

    def view(self, name):
        off, k, shape = self._views[name]
        return self.flat.data[off:off + k].view(shape)

    def grad_view(self, name):
        off, k, shape = self._views[name]
        return self.flat.grad[off:off + k].view(shape)

    def buffer_view(self, name):
        off, k, shape = self._bviews[name]
        return self.bn_buffers[off:off + k].view(shape)

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        if args:
            destination, prefix, keep_vars = (list(args) + [prefix, keep_vars])[:3] if len(args) < 3 else args[:3]
        sd = OrderedDict() if destination is None else destination
        for name in self._views:
            sd[prefix + name] = self.view(name).clone()
        for name in self._bviews:
            sd[prefix + name] = self.buffer_view(name).clone()
        return sd

    def _copy_from(self, sd, prefix, strict, missing):
        with torch.no_grad():
            for table, getter in ((self._views, self.view), (self._bviews, self.buffer_view)):
                for name in table:
                    key = prefix + name
                    if key not in sd and not prefix and "module." + name in sd:
                        key = "module." + name                       # DataParallel-prefixed checkpoint (main.py:147)
                    if key in sd and tuple(sd[key].shape) == tuple(getter(name).shape):
                        getter(name).copy_(sd[key])
                    elif strict:
                        missing.append(key)
        self._weights_dirty = True

    def load_state_dict(self, sd, strict=True):
        missing = []
        self._copy_from(sd, "", strict, missing)
        if missing:
            raise KeyError(missing[0])

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        self._copy_from(state_dict, prefix, strict, missing_keys)

    def _plan(self, B, G, H, W, training=False):
        key = (B, G, H, W, bool(training))
        if key in self._plans:
            return self._plans[key]
        _lib.require_gpu()
        lib = load()
        for old in [k for k in self._plans if k[4] == bool(training)]:          # one live plan per mode
            lib.rxb_rn50_destroy(self._plans.pop(old)["handle"])
        cfg = Rn50Config(B, G, H, W, self.nb_classes, self.size_features, self.bn_eps, self.bn_momentum)
        assert lib.rxb_rn50_param_count(ctypes.byref(cfg)) == self.flat.numel()
        assert lib.rxb_rn50_buffer_count(ctypes.byref(cfg)) == self.bn_buffers.numel()
        nbytes = lib.rxb_rn50_workspace_bytes(ctypes.byref(cfg), 1 if training else 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.flat.device)
        handle = ctypes.c_void_p()
        check(lib.rxb_rn50_create(ctypes.byref(cfg), ptr(self.flat.data), ptr(self.flat.grad), ptr(self.momentum_buf),
                                  ptr(self.bn_buffers), ptr(ws), nbytes, 1 if training else 0, ctypes.byref(handle)))
        plan = {"handle": handle, "ws": ws, "synced": False}
        self._plans[key] = plan
        return plan

    def _sync(self, plan):
        if self._weights_dirty:
            for p in self._plans.values():
                p["synced"] = False
            self._weights_dirty = False
        if not plan["synced"]:
            check(load().rxb_rn50_sync_weights(plan["handle"], stream_ptr()))
            plan["synced"] = True

    def __del__(self):
        try:
            lib = load()
            for p in self._plans.values():
                lib.rxb_rn50_destroy(p["handle"])
        except Exception:
            pass

    def _as_s2d(self, x, G):
        if x.dim() == 5:
            G = x.shape[1]
            x = to_s2d32(x.reshape(-1, *x.shape[2:]).to(self.flat.device))
        elif G is None:
            raise _lib.RxbError("TwoSitesResNet50: a loader-layout batch needs G (images per sample)")
        return x.contiguous(), G

    def forward(self, x, G=None):
        """x: float [B,G,6,H,W] (the reference's item layout, models.py:42) or the fused loader's bf16 S2D32
        [B*G,H/2,W/2,32] together with G.  Returns logits f32 [B, nb_classes].  Evaluation mode only."""
        if self.training:
            raise _lib.RxbError("TwoSitesResNet50.forward runs in evaluation mode only (call .eval()); training goes "
                                "through train_step() / sgd_step() - there is no autograd graph and no PyTorch fallback")
        x, G = self._as_s2d(x, G)
        B, H, W = x.shape[0] // G, x.shape[1] * 2, x.shape[2] * 2
        plan = self._plan(B, G, H, W, False)
        self._sync(plan)
        logits = torch.empty(B, self.nb_classes, dtype=torch.float32, device=x.device)
        check(load().rxb_rn50_forward(plan["handle"], ptr(x), ptr(logits), stream_ptr()))
        return logits

    def dropout_masks(self, B, generator=None):
        """The two Dropout(p) masks of the head (models.py:33,37) as multipliers: 0 with probability p, else 1/(1-p)."""
        dev, p = self.flat.device, self.dropout
        keep = 1.0 - p
        draw = lambda f: (torch.rand(B, f, device=dev, generator=generator) < keep).float() / keep
        return draw(3 * 2048), draw(self.size_features)

    def train_step(self, x, target, G=None, masks=None, global_batch=None, loss_out=None, logits_out=None,
                   feat_out=None):
        """The reference's train step natively (train.py:37,44): forward with batch statistics and Dropout, mean
        CrossEntropy over global_batch samples, backward into self.flat.grad.  `masks` = dropout_masks(B) unless
        given (explicit masks make the step reproducible).  Returns the device scalar with this rank's loss share."""
        x, G = self._as_s2d(x, G)
        B, H, W = x.shape[0] // G, x.shape[1] * 2, x.shape[2] * 2
        plan = self._plan(B, G, H, W, True)
        self._sync(plan)
        m0, m1 = masks if masks is not None else self.dropout_masks(B)
        if loss_out is None:
            loss_out = torch.empty(1, dtype=torch.float32, device=x.device)
        check(load().rxb_rn50_train_step(plan["handle"], ptr(x), ptr(target.contiguous()), ptr(m0.contiguous()),
                                         ptr(m1.contiguous()), global_batch or B, ptr(loss_out), ptr(logits_out),
                                         ptr(feat_out), stream_ptr()))
        return loss_out

    def head_range(self):
        """[begin, end) of the mlp.* parameters in the flat buffers (what stays trainable while a pretrained trunk is
        frozen, train.py:46-58)."""
        off, _, _ = self._views["mlp.0.weight"]
        return off, self.flat.numel()

    def sgd_step(self, B, G, H, W, lr, momentum=0.9, weight_decay=3e-5, nesterov=True, grad_scale=1.0, head_only=False):
        plan = self._plan(B, G, H, W, True)
        check(load().rxb_rn50_sgd(plan["handle"], lr, momentum, weight_decay, 1 if nesterov else 0, grad_scale,
                                  1 if head_only else 0, stream_ptr()))
        for p in self._plans.values():
            p["synced"] = p is plan


class TwoSitesNN(DenseNet121):
    """Reference constructor signature (models.py:8-12).  trunk='densenet121' (default, the north star's trunk; also
    RXB_TRUNK=densenet121) or trunk='resnet50' (RXB_TRUNK=resnet50): the reference's own model, evaluation mode —
    then the object returned is a TwoSitesResNet50.
    pretrained=True (main.py:43 sets it whenever CUDA is available) loads torchvision's ImageNet densenet121 through
    the reference's stem surgery (3-channel stem -> channel mean replicated 6x, models.py:24-26; the 1000-class head is
    dropped); when the weights cannot be found (no network, no RXB_PRETRAINED_DENSENET121 file) it warns and keeps the
    random initialisation instead of failing, so an unchanged main.py still runs."""

    def __new__(cls, pretrained=False, nb_classes=1108, size_features=1024, dropout=0.3, device=None, trunk=None):
        import os
        trunk = trunk or os.environ.get("RXB_TRUNK", "densenet121")
        if trunk == "resnet50":
            return TwoSitesResNet50(pretrained=pretrained, nb_classes=nb_classes, size_features=size_features,
                                    dropout=dropout, device=device)
        if trunk != "densenet121":
            raise ValueError("trunk must be 'densenet121' or 'resnet50'")
        return super().__new__(cls)

    def __init__(self, pretrained=False, nb_classes=1108, size_features=1024, dropout=0.3, device=None, trunk=None):
        super().__init__(nb_classes=nb_classes, device=device)
        self.pretrained_loaded = False
        if pretrained:
            sd = _pretrained_densenet121_state()
            if sd is None:
                import warnings
                warnings.warn("TwoSitesNN(pretrained=True): ImageNet densenet121 weights are not available (no network "
                              "and RXB_PRETRAINED_DENSENET121 is not set) - continuing from random initialisation")
            else:
                self.load_state_dict(sd, strict=False)
                self.pretrained_loaded = True


class DummyClassifier():
    """models.py:60-68: uniform random 'logits' in [-1, 1)."""

    def __init__(self, nb_classes):
        self.nb_classes = nb_classes

    def __call__(self, x):
        bs = x.shape[0]
        output = torch.zeros((bs, self.nb_classes))
        output = output.random_(-10000, 10000) / 10000
        return output
