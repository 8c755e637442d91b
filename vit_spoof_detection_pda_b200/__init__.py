"""vit_spoof_detection_pda_b200 -- B200-native ViT-B/16 PAD hot path (libvitk: hand-written sm_100a CUDA).

Public API mirrors the reference's own objects for this path (/root/reference/train_advanced.py):
``ViTFaceAntiSpoofing`` (:187-204), ``FocalLoss`` (:90-107), and fused stand-ins for
``torch.optim.AdamW`` (:592-597) and ``torch.nn.utils.clip_grad_norm_`` (:334).
"""
from . import _lib
from .checkpoint import extract_model_state_dict, load_checkpoint, save_checkpoint
from .dp import DataParallel, DevicePrefetcher, HostScalars
from .loss import FocalLoss, eval_postprocess
from .metrics import ThresholdSweep, confusion_counts, find_optimal_threshold
from .module import ViTFaceAntiSpoofing
from .optim import FusedAdam, FusedGradScaler, GraphedTrainStep, clip_grad_norm_

__all__ = ["ViTFaceAntiSpoofing", "FocalLoss", "FusedAdam", "FusedGradScaler", "clip_grad_norm_", "GraphedTrainStep", "DataParallel", "DevicePrefetcher", "HostScalars", "eval_postprocess", "ThresholdSweep",
           "find_optimal_threshold", "confusion_counts", "save_checkpoint", "load_checkpoint", "extract_model_state_dict", "_lib"]
