"""Drop-in replacement of the reference's ``lib/models/pose_mobilenet.py``.

``get_pose_net(cfg, is_train=False, cfg_arch=None)`` returns an ``nn.Module`` whose
parameter/buffer names, shapes and creation order equal the reference's
(reference lib/models/pose_mobilenet.py:21-176, lib/models/layers/layers.py:18-24,
90-133), so ``load_state_dict(strict=True)`` (valid.py:157), ``network_to_half``
(lib/fp16_utils/fp16util.py:87-91), ``copy.deepcopy`` and seeded initialisation behave
identically.  What differs is ``forward``: in eval mode on a CUDA tensor it runs the
hand-written sm_100a kernels through the C ABI (litepose_b200.engine); the module
tree itself is only executed for CPU tensors (the parameter-count/summary call of
valid.py:147-150) and in training mode (autograd), never for CUDA inference.
"""
import os

import torch
import torch.nn as nn


def _make_divisible(v, divisor, min_value=None):
    floor = divisor if min_value is None else min_value
    out = max(floor, (int(v + divisor / 2) // divisor) * divisor)
    return out + divisor if out < 0.9 * v else out


def _conv_bn(cin, cout, k, stride, groups, act):
    layers = [nn.Conv2d(cin, cout, k, stride, k // 2, groups=groups, bias=False), nn.BatchNorm2d(cout)]
    if act is not None:
        layers.append(act(inplace=True))
    return layers


class convbnrelu(nn.Sequential):
    """conv k x k + BN + ReLU6 (children 0,1,2 like the reference block)."""

    def __init__(self, inp, oup, ker=3, stride=1, groups=1):
        super().__init__(*_conv_bn(inp, oup, ker, stride, groups, nn.ReLU6))


class InvBottleneck(nn.Module):
    """1x1 expand -> k x k depthwise -> 1x1 project (+ identity when shapes allow)."""

    def __init__(self, inplanes, planes, stride=1, ker=3, exp=6):
        super().__init__()
        mid = _make_divisible(round(inplanes * exp), 8)
        self.inv = nn.Sequential(*_conv_bn(inplanes, mid, 1, 1, 1, nn.ReLU6))
        self.depth_conv = nn.Sequential(*_conv_bn(mid, mid, ker, stride, mid, nn.ReLU6))
        self.point_conv = nn.Sequential(*_conv_bn(mid, planes, 1, 1, 1, None))
        self.stride = stride
        self.use_residual_connection = stride == 1 and inplanes == planes

    def forward(self, x):
        y = self.point_conv(self.depth_conv(self.inv(x)))
        return y + x if self.use_residual_connection else y


class SepConv2d(nn.Module):
    """k x k depthwise + BN + ReLU, then a bias-free 1x1."""

    def __init__(self, inp, oup, ker=3, stride=1):
        super().__init__()
        self.conv = nn.Sequential(*(_conv_bn(inp, inp, ker, stride, inp, nn.ReLU)
                                    + [nn.Conv2d(inp, oup, 1, 1, 0, bias=False)]))

    def forward(self, x):
        return self.conv(x)


class _EngineCache(object):
    """Per-device compiled state; deliberately not copied by deepcopy / pickling.  One object is shared by a module
    and its nn.DataParallel replicas (replicate() copies the module __dict__ shallowly), one engine per device, so
    lookups / builds take the lock (DataParallel runs the replicas on threads, reference valid.py:165)."""

    def __init__(self):
        import threading
        self.engines = {}
        self.lock = threading.RLock()
        self.master_sig = None          # signature of the master module at the last DataParallel replication
        self.keys = None                # its state_dict keys (replicas resolve their tensors by attribute path)

    def clear(self):
        with self.lock:
            self.engines.clear()

    def __deepcopy__(self, memo):
        return _EngineCache()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()


class LitePose(nn.Module):
    def __init__(self, cfg, width_mult=1.0, round_nearest=8, cfg_arch=None):
        super().__init__()
        self.cfg_arch = cfg_arch
        c = _make_divisible(cfg_arch['input_channel'] * width_mult, round_nearest)
        self.first = nn.Sequential(convbnrelu(3, 32, ker=3, stride=2),
                                   convbnrelu(32, 32, ker=3, stride=1, groups=32),
                                   nn.Conv2d(32, c, 1, 1, 0, bias=False),
                                   nn.BatchNorm2d(c))
        self.channel = [c]
        stages = []
        for st in cfg_arch['backbone_setting']:
            cout = _make_divisible(st['channel'] * width_mult, round_nearest)
            blocks = []
            for b in range(st['num_blocks']):
                t, k = st['block_setting'][b]
                blocks.append(InvBottleneck(c, cout, st['stride'] if b == 0 else 1, ker=k, exp=t))
                c = cout
            stages.append(nn.Sequential(*blocks))
            self.channel.append(cout)
        self.stage = nn.ModuleList(stages)
        extra = cfg.MODEL.EXTRA
        self.filters = cfg_arch['deconv_setting']
        self.inplanes = self.channel[-1]
        self.num_deconv_layers = extra.NUM_DECONV_LAYERS
        refined, raw, bnrelu = [], [], []
        for i in range(self.num_deconv_layers):
            k = extra.NUM_DECONV_KERNELS[i]
            if k != 4:
                raise ValueError("LitePose fusion deconv uses 4x4 stride-2 kernels (got %d)" % k)
            planes = self.filters[i]
            refined.append(nn.ConvTranspose2d(self.inplanes, planes, 4, 2, 1, 0, bias=False))
            raw.append(nn.ConvTranspose2d(self.channel[-i - 2], planes, 4, 2, 1, 0, bias=False))
            bnrelu.append(nn.Sequential(nn.BatchNorm2d(planes), nn.ReLU(inplace=True)))
            self.inplanes = planes
        self.deconv_refined, self.deconv_raw = nn.ModuleList(refined), nn.ModuleList(raw)
        self.deconv_bnrelu = nn.ModuleList(bnrelu)
        dim_tag = cfg.MODEL.NUM_JOINTS if cfg.MODEL.TAG_PER_JOINT else 1
        f_ref, f_raw, self.final_channel = [], [], []
        for i in range(1, self.num_deconv_layers):
            oup = (cfg.MODEL.NUM_JOINTS if cfg.LOSS.WITH_HEATMAPS_LOSS[i - 1] else 0) + \
                  (dim_tag if cfg.LOSS.WITH_AE_LOSS[i - 1] else 0)
            f_ref.append(SepConv2d(self.filters[i], oup, ker=5))
            f_raw.append(SepConv2d(self.channel[-i - 3], oup, ker=5))
            self.final_channel.append(oup)
        self.final_refined, self.final_raw = nn.ModuleList(f_ref), nn.ModuleList(f_raw)
        self.loss_config = cfg.LOSS
        self._lp_cache = _EngineCache()

    # -- engine management -------------------------------------------------
    def lp_invalidate(self):
        """Drop the packed-weight cache.  Not needed after the usual mutations: .to()/.half(), load_state_dict,
        train()/eval() switches, re-assignment, and any in-place update of a parameter or buffer itself (optimizer
        step, BN statistics, ``p.copy_`` / ``p.mul_`` under no_grad) are detected.  NOT detected: writes through a
        detached alias (``p.data.copy_(...)``, ``p.detach().mul_()``: the alias carries its own version counter) -
        call this after those."""
        self._lp_cache.clear()

    def _apply(self, fn, *args, **kwargs):      # .cuda() / .half() / .float() / .to()
        self._lp_cache.clear()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._lp_cache.clear()
        return super().load_state_dict(*args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):   # when loaded as a child (network_to_half wrapper)
        self._lp_cache.clear()
        return super()._load_from_state_dict(*args, **kwargs)

    def train(self, mode=True):
        # training steps run the nn.Module tree and update parameters / BN statistics in place: the folded weights
        # of a previous eval phase are stale afterwards
        self._lp_cache.clear()
        return super().train(mode)

    def _lp_signature(self):
        """Hash over (identity, data pointer, autograd version counter) of every parameter and buffer: changes with
        every in-place update of the tensor itself (optimizer.step, running statistics, copy_/mul_ under no_grad) and with
        every re-assignment; writes through ``.data`` / ``.detach()`` aliases are invisible to it (lp_invalidate)."""
        return hash(tuple((id(t), t.data_ptr(), t._version)
                          for group in (self.parameters(), self.buffers()) for t in group))

    def _replicate_for_data_parallel(self):
        # nn.DataParallel replicates the master on every forward: publish the master's signature so that the replicas
        # (whose tensors are fresh broadcasts without history) can tell a stale per-device engine from a current one
        self._lp_cache.master_sig = self._lp_signature()
        self._lp_cache.keys = list(self.state_dict().keys())
        return super()._replicate_for_data_parallel()

    def _lp_state(self):
        """state_dict of this module; DataParallel replicas hold their (broadcast) tensors as plain attributes, not as
        registered parameters, so their entries are resolved by attribute path from the master's key list."""
        if not getattr(self, "_is_replica", False):
            return self.state_dict()
        out = {}
        for k in self._lp_cache.keys:
            obj = self
            for part in k.split("."):
                obj = getattr(obj, part)
            out[k] = obj
        return out

    def lp_engine(self, device=None):
        from litepose_b200.engine import LitePoseEngine
        dev = torch.device(device) if device is not None else next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("the litepose_b200 engine needs the module on a CUDA device")
        key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
        # DataParallel replicas receive freshly broadcast tensors on every forward (version 0, new identity): they use
        # the signature the master published when it was replicated
        replica = getattr(self, "_is_replica", False)
        sig = self._lp_cache.master_sig if replica else self._lp_signature()
        with self._lp_cache.lock:
            hit = self._lp_cache.engines.get(key)
            if hit is not None and hit[1] == sig:
                return hit[0]
            eng = LitePoseEngine(self._lp_state(), self.cfg_arch, torch.device("cuda", key[1]))
            self._lp_cache.engines[key] = (eng, sig)
            return eng

    # -- forward -------------------------------------------------------------
    def _forward_modules(self, x):
        x = self.first(x)
        feats = [x]
        for st in self.stage:
            feats.append(st(feats[-1]))
        outs = []
        refined, raw = feats[-1], feats[-2]
        for i in range(self.num_deconv_layers):
            refined = self.deconv_bnrelu[i](self.deconv_refined[i](refined) + self.deconv_raw[i](raw))
            raw = feats[-i - 3]
            if i > 0:
                outs.append(self.final_refined[i - 1](refined) + self.final_raw[i - 1](raw))
        return outs

    def forward(self, x):
        if x.is_cuda and not self.training:
            # CUDA inference: hand-written sm_100a kernels only (raises if the library is missing)
            half_in = x.dtype == torch.float16
            return self.lp_engine(x.device).run(x, flip=False, out_fp32=not half_in, clone=True)
        return self._forward_modules(x)


def get_pose_net(cfg, is_train=False, cfg_arch=None):
    model = LitePose(cfg, cfg_arch=cfg_arch)
    if is_train and cfg.MODEL.INIT_WEIGHTS:
        print(cfg.MODEL.PRETRAINED)
        if os.path.isfile(cfg.MODEL.PRETRAINED):
            print("load pre-train model")
            state = torch.load(cfg.MODEL.PRETRAINED, map_location=torch.device('cpu'))
            keep = {k: v for k, v in state.items() if 'deconv' not in k and 'final' not in k}
            try:
                model.load_state_dict(keep, strict=False)
            except Exception:
                print("Error load!")
    return model
