"""Drop-in ``ViTFaceAntiSpoofing`` (reference: /root/reference/train_advanced.py:187-204, dups
test.py:71-89, simple/train.py:185-202) whose forward/backward run on libvitk's sm_100a kernels.

Boundary contract:
  * same constructor protocol: ``ViTFaceAntiSpoofing(config)`` with ``config.model_name ==
    "vit_base_patch16_224"``, ``.pretrained``, ``.num_classes``, ``.dropout``
  * same attributes: ``.vit`` (``.num_features == 768``) and ``.classifier`` (indexable, parameters at 0, 2, 5)
  * identical ``state_dict`` keys / shapes (156 fp32 tensors, no buffers), strict ``load_state_dict``
  * ``forward(x: float[B,3,224,224]) -> Tensor[B, num_classes]`` differentiable through ``torch.autograd``

The torch sub-modules below are parameter containers only: their ``forward`` is never called.  All
parameters are views into ONE flat fp32 buffer (state_dict order, layout from ``vitk_param_layout``);
gradients come back as views into one flat fp32 gradient buffer, so the fused optimizer and the bucketed
data-parallel all-reduce work on contiguous memory without copies.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L


class _Attn(nn.Module):
    def __init__(self):
        super().__init__()
        self.qkv = nn.Linear(L.DIM, 3 * L.DIM, bias=True)
        self.proj = nn.Linear(L.DIM, L.DIM)


class _Mlp(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(L.DIM, L.MLP)
        self.fc2 = nn.Linear(L.MLP, L.DIM)


class _Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.norm1 = nn.LayerNorm(L.DIM, eps=1e-6)
        self.attn = _Attn()
        self.norm2 = nn.LayerNorm(L.DIM, eps=1e-6)
        self.mlp = _Mlp()


class _PatchEmbed(nn.Module):
    def __init__(self):
        super().__init__()
        self.proj = nn.Conv2d(3, L.DIM, kernel_size=L.PATCH, stride=L.PATCH)


class _ViTParams(nn.Module):
    """Parameter container with timm's ``vit_base_patch16_224`` (num_classes=0) names."""

    def __init__(self, depth: int):
        super().__init__()
        self.num_features = L.DIM
        self.embed_dim = L.DIM
        self.cls_token = nn.Parameter(torch.zeros(1, 1, L.DIM))
        self.pos_embed = nn.Parameter(torch.zeros(1, L.NTOK, L.DIM))
        self.patch_embed = _PatchEmbed()
        self.blocks = nn.Sequential(*[_Block() for _ in range(depth)])
        self.norm = nn.LayerNorm(L.DIM, eps=1e-6)

    def forward(self, x):  # pragma: no cover - the encoder only runs fused with the head
        raise RuntimeError("call the owning ViTFaceAntiSpoofing module; the encoder runs inside libvitk")


def _timm_style_init(model: nn.Module):
    """timm VisionTransformer init: trunc_normal(.02) Linear weights, zero biases, LN ones/zeros."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("cls_token"):
                nn.init.normal_(p, std=1e-6)
            elif name.endswith("pos_embed"):
                nn.init.trunc_normal_(p, std=0.02)
            elif name.startswith("vit.") and p.dim() == 2:
                nn.init.trunc_normal_(p, std=0.02)
            elif name.startswith("vit.") and p.dim() == 1 and "norm" not in name:
                nn.init.zeros_(p)


class _ViTPADFunction(torch.autograd.Function):
    """autograd.Function around the whole encoder + head (one C-ABI call forward, depth+2 backward stages)."""

    @staticmethod
    def forward(ctx, owner, images, *params):
        logits, gen = owner._run_forward(images, training=True)
        ctx.owner = owner
        ctx.gen = gen
        ctx.batch = images.shape[0]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        owner = ctx.owner
        grads = owner._run_backward(dlogits, ctx.batch, ctx.gen)
        return (None, None, *grads)


class ViTFaceAntiSpoofing(nn.Module):
    def __init__(self, config=None, *, dropout: Optional[float] = None, num_classes: Optional[int] = None,
                 depth: int = 12, precision: str = "bf16", engine: int = L.ENGINE_AUTO):
        super().__init__()
        if config is not None:
            name = getattr(config, "model_name", "vit_base_patch16_224")
            if name != "vit_base_patch16_224":
                raise ValueError(f"only vit_base_patch16_224 is implemented (got {name})")
            if getattr(config, "pretrained", False):
                # the reference downloads ImageNet weights through timm; there is no network here and no
                # silent substitute: load a checkpoint with load_state_dict instead
                import warnings
                warnings.warn("pretrained=True ignored: no network; use load_state_dict() for real weights")
            dropout = getattr(config, "dropout", 0.1) if dropout is None else dropout
            num_classes = getattr(config, "num_classes", 2) if num_classes is None else num_classes
        dropout = 0.1 if dropout is None else float(dropout)
        num_classes = 2 if num_classes is None else int(num_classes)
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.depth = int(depth)
        self.num_classes = num_classes
        self.dropout_p = dropout
        self.precision = precision
        self.engine = engine
        self.vit = _ViTParams(self.depth)
        embed_dim = self.vit.num_features
        self.classifier = nn.Sequential(
            nn.LayerNorm(embed_dim), nn.Dropout(dropout), nn.Linear(embed_dim, L.HEAD_HIDDEN), nn.GELU(),
            nn.Dropout(dropout), nn.Linear(L.HEAD_HIDDEN, num_classes))
        _timm_style_init(self)
        self._total, self._offsets, self._sizes = None, None, None
        self._flat = None          # fp32 master parameters
        self._flat16 = None        # bf16 shadow
        self._flat_grad = None
        self._flat_grad_alt = None
        self._zero_stream = None   # side stream that clears the flat gradient buffer under the forward pass
        self._prezero = None       # (forward generation, event): the buffer is (being) cleared for that forward's backward
        self._shadow_version = -1
        self._ws = {}              # (batch, training, frozen) -> workspace tensor
        self._gen = 0
        self._bucket_hook = None   # set by DataParallel: fn(stage, lo, hi) after each backward stage
        self._finish_hook = None   # set by DataParallel: fn() before gradients are handed to autograd
        self._pending_clip = None  # (sumsq tensor, max_norm) left by clip_grad_norm_ for FusedAdam
        self._sm_budget = 0        # set by DataParallel while a collective is in flight (vitk_model.sm_budget)
        self._nvls = None          # set by DataParallel(mode="nvls"): the fused reduce + Adam + all-gather step (dp.NvlsStep)
        self._masters_sync = None  # NVLS: waits for the background fp32-master copy (state_dict / flat_params call it)
        self.register_state_dict_pre_hook(lambda mod, prefix, keep_vars: mod._masters_sync() if mod._masters_sync else None)
        self._flags = 0            # vitk_model.flags (L.FLAG_*)
        self.last_masks = None
        for p in self.parameters():
            p._vitk_owner = weakref.ref(self)

    # ------------------------------------------------------------------ flat storage management
    def _param_list(self):
        """list(self.parameters()), cached: walking the module tree costs ~0.25 ms and the step needs the list ~10 times.
        The cache is validated by identity (every Parameter still sits in the same slot of the same sub-module, every
        sub-module in the same slot of its parent: ~40 us), so replacing a parameter or a sub-module rebuilds it."""
        cache = self.__dict__.get("_plist_cache")
        if cache is not None:
            plist, pslots, mslots = cache
            if all(m._parameters.get(k) is p for (m, k), p in zip(pslots, plist)) and \
                    all(parent._modules.get(k) is child for parent, k, child in mslots):
                return plist
        plist, pslots, mslots = [], [], []
        seen = set()

        def walk(mod):
            for k, p in mod._parameters.items():
                if p is not None and id(p) not in seen:
                    seen.add(id(p))
                    plist.append(p)
                    pslots.append((mod, k))
            for k, child in mod._modules.items():
                if child is not None:
                    mslots.append((mod, k, child))
                    walk(child)

        walk(self)
        ref = list(self.parameters())
        assert len(ref) == len(plist) and all(a is b for a, b in zip(ref, plist)), "parameter order differs from nn.Module.parameters()"
        self.__dict__["_plist_cache"] = (plist, pslots, mslots)
        return plist

    def _layout(self):
        if self._total is None:
            self._total, self._offsets, self._sizes = L.param_layout(self.depth, self.num_classes)
            plist = self._param_list()
            assert len(plist) == len(self._offsets), "state_dict contract broken"
            for p, n in zip(plist, self._sizes):
                assert p.numel() == n, "parameter shape does not match vitk_param_layout"
        return self._total, self._offsets, self._sizes

    def _flat_ok(self, plist) -> bool:
        if self._flat is None:
            return False
        base = self._flat.data_ptr()
        for p, off in zip(plist, self._offsets):
            if p.data_ptr() != base + 4 * off:
                return False
        return True

    def _ensure_flat(self):
        total, offs, sizes = self._layout()
        plist = self._param_list()
        dev = plist[0].device
        if dev.type != "cuda":
            raise RuntimeError("ViTFaceAntiSpoofing (vitk) runs on CUDA only: call .to('cuda') first; there is no CPU fallback")
        if not self._flat_ok(plist):
            flat = torch.zeros(total, dtype=torch.float32, device=dev)
            for p, off, n in zip(plist, offs, sizes):
                if p.dtype != torch.float32:
                    raise RuntimeError("master parameters must stay fp32")
                flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
                p._vitk_owner = weakref.ref(self)
            self._flat = flat
            self._flat16 = None
            self._flat_grad = None
            self._flat_grad_alt = None
            self._shadow_version = -1
            self._ws = {}
        return plist

    def _params_version(self, plist) -> int:
        # version counters of the Parameters plus the flat buffer's own (writes through flat_params(), dist.broadcast)
        return sum(p._version for p in plist) + (self._flat._version if self._flat is not None else 0)

    def adopt_flat_buffers(self, flat: torch.Tensor, flat16: Optional[torch.Tensor], grad: torch.Tensor):
        """Move the flat fp32 masters / bf16 shadow / gradient storage into caller-provided buffers of (at least) the layout's
        size -- used by DataParallel's NVLS mode, whose buffers are symmetric memory mapped behind a multicast address.
        Current parameter values are copied; Parameters (and existing .grad views) are re-pointed."""
        plist = self._ensure_flat()
        total = self._total
        assert flat.numel() >= total and flat.dtype == torch.float32 and grad.numel() >= total and grad.dtype == torch.float32
        flat[:total].copy_(self._flat)
        for p, off, n in zip(plist, self._offsets, self._sizes):
            p.data = flat[off:off + n].view(p.shape)
            p.grad = None
        self._flat = flat[:total]
        self._flat_grad = grad[:total]
        self._flat_grad_alt = None
        self._flat16 = flat16[:total] if flat16 is not None else None
        self._shadow_version = -1
        self.__dict__.pop("_plist_cache", None)

    def invalidate_shadow(self):
        """Force the bf16 weight shadow to be rebuilt before the next forward.  Call after writing the fp32 masters in a
        way that bypasses the Parameters' version counters (``p.data.copy_()``, raw writes through ``flat_params()`` on
        another stream, external collectives)."""
        self._shadow_version = -1

    def _ensure_shadow(self, plist):
        if self.precision != "bf16":
            return
        v = self._params_version(plist)
        if self._flat16 is None:
            self._flat16 = torch.empty(self._total, dtype=torch.bfloat16, device=self._flat.device)
            self._shadow_version = -1
        if v != self._shadow_version:
            L.call("vitk_cast_f32_to_bf16", L.ptr(self._flat), L.ptr(self._flat16), self._total, L.stream_ptr())
            self._shadow_version = v

    def mark_shadow_fresh(self):
        """Called by FusedAdam, whose kernel rewrites the bf16 shadow together with the fp32 masters."""
        self._shadow_version = self._params_version(self._param_list())

    def flat_params(self):
        self._ensure_flat()
        if self._masters_sync is not None:
            self._masters_sync()       # NVLS data parallel: the background copy of the other ranks' fp32 master slices
        return self._flat

    def flat_params16(self):
        plist = self._ensure_flat()
        self._ensure_shadow(plist)
        return self._flat16

    def flat_grads(self):
        self._ensure_flat()
        if self._flat_grad is None:
            self._flat_grad = torch.zeros(self._total, dtype=torch.float32, device=self._flat.device)
        return self._flat_grad

    def param_ranges(self):
        """[(name, offset, size)] in flat order."""
        _, offs, sizes = self._layout()
        return [(n, o, s) for (n, _), o, s in zip(self.named_parameters(), offs, sizes)]

    def stage_ranges(self):
        """Flat [lo, hi) element range whose gradients are final after backward stage s."""
        total, offs, _ = self._layout()
        d = self.depth
        first_block, per_block = 4, 12
        ranges = [(offs[first_block + per_block * d], total)]               # stage 0: vit.norm + classifier
        for i in range(d - 1, -1, -1):                                     # stage 1 + (d-1-i): block i
            lo = offs[first_block + per_block * i]
            hi = offs[first_block + per_block * (i + 1)]
            ranges.append((lo, hi))
        ranges.append((0, offs[first_block]))                              # stage d+1: cls, pos, patch_embed
        return ranges

    @property
    def backbone_frozen(self) -> bool:
        return not any(p.requires_grad for p in self.vit.parameters())

    # ------------------------------------------------------------------ execution
    def _workspace(self, batch: int, training: bool, frozen: bool):
        key = (batch, training, frozen, self.precision)
        ws = self._ws.get(key)
        if ws is None:
            prec = L.PREC_BF16 if self.precision == "bf16" else L.PREC_FP32
            nbytes = L.load().vitk_workspace_bytes(batch, self.depth, prec, (1 if training else 0) | (2 if frozen else 0))
            ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self._flat.device)
            self._ws[key] = ws
        return ws

    def _model_struct(self, batch, training, frozen, ws, images=None, logits=None, dlogits=None, grads=None, masks=None):
        m = L.VitkModel()
        m.batch, m.depth, m.num_classes = batch, self.depth, self.num_classes
        m.precision = L.PREC_BF16 if self.precision == "bf16" else L.PREC_FP32
        m.training = 1 if training else 0
        m.engine = self.engine
        m.params = L.ptr(self._flat)
        m.params16 = L.ptr(self._flat16) if self.precision == "bf16" else None
        m.grads = L.ptr(grads)
        base = ws.data_ptr()
        m.workspace = (base + 255) // 256 * 256
        if images is not None and images.dtype == torch.uint8:
            # uint8 HWC pixels: ToTensor + Normalize happen inside the patch loader
            m.images = None
            m.images_u8 = L.ptr(images)
            m.norm_mean = (C.c_float * 3)(*self.pixel_mean)
            m.norm_std = (C.c_float * 3)(*self.pixel_std)
        else:
            m.images = L.ptr(images)
            m.images_u8 = None
        m.logits = L.ptr(logits)
        m.dlogits = L.ptr(dlogits)
        m.mask1 = L.ptr(masks[0]) if masks else None
        m.mask2 = L.ptr(masks[1]) if masks else None
        m.frozen_backbone = 1 if frozen else 0
        m.sm_budget = int(self._sm_budget)
        m.flags = int(self._flags)
        return m

    # ImageNet statistics of the reference's transforms.Normalize (train_advanced.py:175, 181; test.py:162)
    pixel_mean = (0.485, 0.456, 0.406)
    pixel_std = (0.229, 0.224, 0.225)

    def _prep_images(self, x):
        if not x.is_cuda:
            raise RuntimeError("vitk forward needs CUDA tensors (no CPU fallback)")
        if x.dtype == torch.uint8:
            # raw resized pixels, HWC as PIL / the decoder delivers them: [B,224,224,3]
            if x.dim() != 4 or tuple(x.shape[1:]) != (L.IMG, L.IMG, 3):
                raise ValueError(f"expected uint8 images [B,{L.IMG},{L.IMG},3] (HWC), got {tuple(x.shape)}")
            return x.detach().contiguous()
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, L.IMG, L.IMG):
            raise ValueError(f"expected images [B,3,{L.IMG},{L.IMG}], got {tuple(x.shape)}")
        return x.detach().to(torch.float32).contiguous()

    def _run_forward(self, images, training: bool):
        # `training` here means "keep activations for backward"; dropout follows nn.Module.training
        plist = self._ensure_flat()
        self._ensure_shadow(plist)
        x = self._prep_images(images)
        B = x.shape[0]
        frozen = training and self.backbone_frozen
        ws = self._workspace(B, training, frozen)
        logits = torch.empty(B, self.num_classes, dtype=torch.float32, device=x.device)
        masks = None
        if self.training and self.dropout_p > 0.0:
            keep = 1.0 - self.dropout_p
            masks = (torch.empty(B, L.DIM, device=x.device).bernoulli_(keep).div_(keep),
                     torch.empty(B, L.HEAD_HIDDEN, device=x.device).bernoulli_(keep).div_(keep))
        m = self._model_struct(B, training, frozen, ws, images=x, logits=logits, masks=masks)
        if training:
            self._start_prezero(plist)
        with L.nvtx_range(f"vitk/forward bs={B} {'train' if training else 'eval'}"):
            L.call("vitk_model_fwd", C.byref(m), L.stream_ptr())
        self._gen += 1
        if training:
            self._saved = (x, masks, frozen, ws, self._gen)
            self.last_masks = masks
        return logits, self._gen

    def _start_prezero(self, plist):
        """The weight-gradient GEMMs accumulate (split-K partial tiles, red.global.add), so the flat gradient buffer must be
        zero when backward starts: 345 MB of stores, ~0.1 ms at the head of the backward pass.  The forward GEMMs leave most
        of the HBM write bandwidth idle, so the clear is started here on a side stream and backward only waits for its
        event.  Skipped (backward clears the buffer itself, as before) while any .grad still aliases the buffer --
        gradient accumulation, or zero_grad() called after the forward."""
        if os.environ.get("VITK_NO_PREZERO") == "1":
            return
        g = self.flat_grads()
        lo_ptr, hi_ptr = g.data_ptr(), g.data_ptr() + 4 * self._total
        if any(p.grad is not None and lo_ptr <= p.grad.data_ptr() < hi_ptr for p in plist):
            self._prezero = None
            return
        if self._zero_stream is None:
            self._zero_stream = torch.cuda.Stream(device=g.device)
        self._zero_stream.wait_stream(torch.cuda.current_stream())   # after the previous step's readers (Adam, all-reduce)
        with torch.cuda.stream(self._zero_stream):
            g.zero_()
            ev = torch.cuda.Event()
            ev.record()
        self._prezero = (self._gen + 1, ev)

    def _run_backward(self, dlogits, batch, gen):
        x, masks, frozen, ws, saved_gen = self._saved
        if saved_gen != gen:
            raise RuntimeError("activations of this forward were overwritten by a later forward of the same module")
        plist = self._param_list()
        # gradient accumulation: if .grad tensors still alias the flat buffer, accumulate through a second buffer
        g = self.flat_grads()
        lo_ptr, hi_ptr = g.data_ptr(), g.data_ptr() + 4 * self._total
        aliased = any(p.grad is not None and lo_ptr <= p.grad.data_ptr() < hi_ptr for p in plist)
        if aliased:
            if self._flat_grad_alt is None:
                self._flat_grad_alt = torch.empty_like(g)
            g = self._flat_grad_alt
        pz, self._prezero = self._prezero, None
        if not aliased and pz is not None and pz[0] == gen:
            torch.cuda.current_stream().wait_event(pz[1])      # cleared on the side stream while the forward ran
        else:
            g.zero_()
        self._bwd_serial = getattr(self, "_bwd_serial", 0) + 1   # optim.py: gradients were (re)produced
        dl = dlogits.detach().to(torch.float32).contiguous()
        m = self._model_struct(batch, True, frozen, ws, images=x, dlogits=dl, grads=g, masks=masks)
        st = L.stream_ptr()
        ranges = self.stage_ranges()
        n_stages = 1 if frozen else self.depth + 2
        for s in range(n_stages):
            with L.nvtx_range(f"vitk/backward stage {s}"):
                L.call("vitk_model_bwd_stage", C.byref(m), s, st)
            # data parallel: every backward reduces the gradient IT produced -- also the micro-step gradients of a
            # gradient-accumulation loop and the backward after zero_grad(set_to_none=False), which land in the second
            # buffer (`aliased`) before autograd adds them to .grad.  The all-reduce is linear, so reducing each micro-step
            # and accumulating equals torch DDP's result.
            if self._bucket_hook is not None:
                self._bucket_hook(s, ranges[s][0], ranges[s][1], g)
                m.sm_budget = int(self._sm_budget)   # a collective now in flight owns some SMs (dp.py)
        if self._finish_hook is not None:
            self._finish_hook(g)
        outs = []
        for p, off, n in zip(plist, self._offsets, self._sizes):
            outs.append(g[off:off + n].view(p.shape) if p.requires_grad else None)
        return outs

    def forward(self, x):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            plist = self._ensure_flat()
            return _ViTPADFunction.apply(self, x, *plist)
        logits, _ = self._run_forward(x, training=False)
        return logits

