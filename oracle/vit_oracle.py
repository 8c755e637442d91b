"""CPU oracle for the ViT-B/16 PAD hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product package (``vit_spoof_detection_pda_b200``) must never
import anything from ``oracle/``.

What it restates (all citations relative to /root/reference):

* ``OracleViTFaceAntiSpoofing``   -> train_advanced.py:187-204 (dups test.py:71-89)
* ``OracleFocalLoss``             -> train_advanced.py:90-107 (scalar alpha) plus the
                                     per-class-alpha generalisation north_star asks for
                                     (class weights: train_advanced.py:521-529)
* ``oracle_train_step``           -> train_advanced.py:322-346 in its CPU form (autocast and
                                     GradScaler self-disable on a CPU host -> pure fp32)
* ``oracle_eval_step``            -> train_advanced.py:379-394, test.py:205-218
* ``OracleViTEncoder``            -> the third-party dependency ``timm`` (un-pinned in
                                     requirements.txt:4, not vendored, not installable here):
                                     ``timm.create_model("vit_base_patch16_224", num_classes=0)``
                                     (call site train_advanced.py:190).  Its published algorithm:
                                     Conv2d(3,768,k16,s16) patch embedding, CLS token prepended,
                                     learned pos-embed added, 12 pre-norm blocks
                                     [LN(eps 1e-6) -> qkv Linear(768,2304) -> 12-head SDPA
                                     (scale 64^-0.5) -> proj Linear -> +res -> LN -> fc1
                                     Linear(768,3072) -> exact-erf GELU -> fc2 -> +res], final
                                     LN(eps 1e-6), CLS-token pooling, head = Identity.

Parity pinning status (see DESIGN.md "Oracle"):
  - focal loss, head/wrapper and the train/eval step are pinned against the REAL reference
    classes executed in the build container (oracle/make_golden.py imports
    /root/reference/train_advanced.py with a stub ``timm`` whose ``create_model`` returns
    ``OracleViTEncoder``) -> tests/golden/*.pt.
  - the encoder arithmetic itself ("timm") is absent from /root/reference and from this image;
    it is cross-checked against torchvision.models.vit_b_16 (same published architecture) through
    a key remap -> tests/golden/encoder_xcheck.json.  The reference holds no golden vectors,
    KATs or tests for this path, so for the encoder: "parity unpinned"
    by the reference's own fixtures; pinned only by the torchvision cross-check.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

IMG = 224
PATCH = 16
GRID = IMG // PATCH          # 14
N_PATCH = GRID * GRID        # 196
N_TOK = N_PATCH + 1          # 197
DIM = 768
HEADS = 12
HEAD_DIM = 64
MLP = 3072
DEPTH = 12
HEAD_HIDDEN = 512


class _Attention(nn.Module):
    """timm ``Attention`` (qkv_bias=True, qk_norm=False, no dropout)."""

    def __init__(self, dim: int = DIM, heads: int = HEADS):
        super().__init__()
        self.heads = heads
        self.head_dim = dim // heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        # written out (not F.sdpa) so the arithmetic is explicit: softmax(q k^T * scale) v
        attn = (q * self.scale) @ k.transpose(-2, -1)
        attn = attn.softmax(dim=-1)
        x = attn @ v
        x = x.transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class _Mlp(nn.Module):
    def __init__(self, dim: int = DIM, hidden: int = MLP):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()  # exact erf
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim: int = DIM, heads: int = HEADS, hidden: int = MLP):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, hidden)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class _PatchEmbed(nn.Module):
    def __init__(self, dim: int = DIM):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=PATCH, stride=PATCH, bias=True)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)  # [B,196,768]


class OracleViTEncoder(nn.Module):
    """timm ``vit_base_patch16_224`` with ``num_classes=0`` (returns CLS feature [B,768])."""

    def __init__(self, depth: int = DEPTH):
        super().__init__()
        self.num_features = DIM
        self.patch_embed = _PatchEmbed()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, DIM))
        self.pos_embed = nn.Parameter(torch.zeros(1, N_TOK, DIM))
        self.blocks = nn.Sequential(*[_Block() for _ in range(depth)])
        self.norm = nn.LayerNorm(DIM, eps=1e-6)

    def forward_tokens(self, x):
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        x = x + self.pos_embed
        x = self.blocks(x)
        return self.norm(x)

    def forward(self, x):
        return self.forward_tokens(x)[:, 0]


class OracleViTFaceAntiSpoofing(nn.Module):
    """Restates train_advanced.py:187-204 with the encoder above standing in for timm."""

    def __init__(self, dropout: float = 0.0, num_classes: int = 2, depth: int = DEPTH):
        super().__init__()
        self.vit = OracleViTEncoder(depth)
        self.classifier = nn.Sequential(
            nn.LayerNorm(DIM),             # eps 1e-5 (nn default), train_advanced.py:194
            nn.Dropout(dropout),
            nn.Linear(DIM, HEAD_HIDDEN),
            nn.GELU(),
            nn.Dropout(dropout),
            nn.Linear(HEAD_HIDDEN, num_classes),
        )

    def forward(self, x):
        return self.classifier(self.vit(x))


class OracleFocalLoss(nn.Module):
    """train_advanced.py:90-107; ``alpha`` may also be a per-class sequence (alpha[target])."""

    def __init__(self, alpha=0.25, gamma: float = 2.0, reduction: str = "mean"):
        super().__init__()
        self.alpha = alpha
        self.gamma = gamma
        self.reduction = reduction

    def forward(self, inputs, targets):
        ce = F.cross_entropy(inputs.float(), targets, reduction="none")
        pt = torch.exp(-ce)
        if isinstance(self.alpha, (float, int)):
            a = float(self.alpha)
        else:
            a = torch.as_tensor(self.alpha, dtype=ce.dtype, device=ce.device)[targets]
        fl = a * (1 - pt) ** self.gamma * ce
        if self.reduction == "mean":
            return fl.mean()
        if self.reduction == "sum":
            return fl.sum()
        return fl


def class_weights_from_counts(n_live: int, n_spoof: int):
    """train_advanced.py:521-529: [total/(2*spoof), total/(2*live)] (index 0 = spoof, 1 = live)."""
    total = n_live + n_spoof
    return [total / (2 * n_spoof), total / (2 * n_live)]


# --------------------------------------------------------------------------------------
# deterministic init shared by oracle, tests and bench (weights are COPIED into the product
# module through load_state_dict, so the distribution is immaterial; biases / LN affine are
# perturbed so those paths are exercised)
# --------------------------------------------------------------------------------------
def seeded_init_(model: nn.Module, seed: int = 42, perturb: bool = True) -> nn.Module:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("cls_token"):
                p.copy_(torch.randn(p.shape, generator=g) * (0.02 if perturb else 1e-6))
            elif name.endswith("pos_embed"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            elif p.dim() >= 2:  # Linear / conv weights
                std = 0.02
                p.copy_(torch.randn(p.shape, generator=g).clamp_(-2, 2) * std)
            elif "norm" in name or name.startswith("classifier.0"):
                if name.endswith("weight"):
                    p.copy_(1.0 + (0.1 * torch.randn(p.shape, generator=g) if perturb else 0.0))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g) if perturb else torch.zeros(p.shape))
            else:  # biases
                p.copy_(0.02 * torch.randn(p.shape, generator=g) if perturb else torch.zeros(p.shape))
    return model


def synthetic_batch(batch: int, seed: int = 42):
    """BASELINE config inputs: images ~ N(0,1) fp32 [B,3,224,224]; labels in {0,1} int64."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, IMG, IMG, generator=g)
    labels = torch.randint(0, 2, (batch,), generator=g)
    return images, labels


# --------------------------------------------------------------------------------------
# the step the reference runs (CPU form: fp32, scaler disabled)
# --------------------------------------------------------------------------------------
def oracle_train_step(model, criterion, optimizer, images, labels, max_grad_norm: Optional[float] = 1.0,
                      scheduler=None):
    """train_advanced.py:322-346 (gradient_accumulation_steps == 1, mixed precision off)."""
    model.train()
    outputs = model(images)
    loss = criterion(outputs, labels)
    loss.backward()
    gnorm = None
    if max_grad_norm is not None:
        gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    if scheduler is not None:
        scheduler.step()
    with torch.no_grad():
        preds = torch.argmax(outputs, dim=1)
        acc = (preds == labels).float().mean()
    return loss.item(), acc.item(), (None if gnorm is None else float(gnorm))


@torch.no_grad()
def oracle_eval_step(model, images):
    """train_advanced.py:384-394 / test.py:209-217: logits, softmax, argmax, P(live)=probs[:,1]."""
    model.eval()
    outputs = model(images)
    probs = F.softmax(outputs, dim=1)
    preds = torch.argmax(outputs, dim=1)
    return outputs, probs[:, 1], preds


def make_optimizer(params, kind: str = "adam", lr: float = 1e-5, weight_decay: float = 1e-4):
    """'adam'  -> README.md:140-147 / north_star (Adam + L2 weight decay 1e-4)
       'adamw' -> train_advanced.py:592-597 (AdamW lr 3e-4 wd 0.05 betas (0.9,0.999))."""
    if kind == "adam":
        return torch.optim.Adam(params, lr=lr, weight_decay=weight_decay, betas=(0.9, 0.999), eps=1e-8)
    if kind == "adamw":
        return torch.optim.AdamW(params, lr=lr, weight_decay=weight_decay, betas=(0.9, 0.999), eps=1e-8)
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# torchvision cross-check remap
# --------------------------------------------------------------------------------------
def encoder_state_to_torchvision(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Map OracleViTEncoder keys -> torchvision.models.vit_b_16 keys (heads removed)."""
    out = {}
    for k, v in sd.items():
        if k == "cls_token":
            out["class_token"] = v
        elif k == "pos_embed":
            out["encoder.pos_embedding"] = v
        elif k.startswith("patch_embed.proj."):
            out["conv_proj." + k.split(".")[-1]] = v
        elif k.startswith("norm."):
            out["encoder.ln." + k.split(".")[-1]] = v
        elif k.startswith("blocks."):
            _, i, rest = k.split(".", 2)
            pre = f"encoder.layers.encoder_layer_{i}."
            rest = (rest.replace("norm1.", "ln_1.").replace("norm2.", "ln_2.")
                        .replace("attn.qkv.weight", "self_attention.in_proj_weight")
                        .replace("attn.qkv.bias", "self_attention.in_proj_bias")
                        .replace("attn.proj.", "self_attention.out_proj.")
                        .replace("mlp.fc1.", "mlp.0.").replace("mlp.fc2.", "mlp.3."))
            out[pre + rest] = v
        else:
            raise KeyError(k)
    return out


def expected_state_dict_spec(depth: int = DEPTH, num_classes: int = 2):
    """The 156-tensor state_dict contract as an ordered [(name, shape)]."""
    spec = [("vit.cls_token", (1, 1, DIM)), ("vit.pos_embed", (1, N_TOK, DIM)),
            ("vit.patch_embed.proj.weight", (DIM, 3, PATCH, PATCH)), ("vit.patch_embed.proj.bias", (DIM,))]
    for i in range(depth):
        p = f"vit.blocks.{i}."
        spec += [(p + "norm1.weight", (DIM,)), (p + "norm1.bias", (DIM,)),
                 (p + "attn.qkv.weight", (3 * DIM, DIM)), (p + "attn.qkv.bias", (3 * DIM,)),
                 (p + "attn.proj.weight", (DIM, DIM)), (p + "attn.proj.bias", (DIM,)),
                 (p + "norm2.weight", (DIM,)), (p + "norm2.bias", (DIM,)),
                 (p + "mlp.fc1.weight", (MLP, DIM)), (p + "mlp.fc1.bias", (MLP,)),
                 (p + "mlp.fc2.weight", (DIM, MLP)), (p + "mlp.fc2.bias", (DIM,))]
    spec += [("vit.norm.weight", (DIM,)), ("vit.norm.bias", (DIM,)),
             ("classifier.0.weight", (DIM,)), ("classifier.0.bias", (DIM,)),
             ("classifier.2.weight", (HEAD_HIDDEN, DIM)), ("classifier.2.bias", (HEAD_HIDDEN,)),
             ("classifier.5.weight", (num_classes, HEAD_HIDDEN)), ("classifier.5.bias", (num_classes,))]
    return spec


TRAIN_FLOP_PER_IMG = 105.150e9
FWD_FLOP_PER_IMG = 35.127e9
