"""Checkpoint container of the reference, read and written unchanged.

``save_checkpoint`` mirrors /root/reference/train_advanced.py:475-489 (same signature, same dict keys: epoch,
model_state_dict, optimizer_state_dict, scheduler_state_dict, scaler_state_dict, metrics, config) and ``load_checkpoint``
mirrors /root/reference/test.py:167-188 (same signature, returns ``(model, checkpoint)``, FileNotFoundError when the file
is missing).  The loader also accepts what /root/reference/testing_set_analysis_src/evaluate_all_models.py:293-298
tolerates for the published ``best_model_run_eif1jakb.pth``: the weights under ``model_state_dict``, under
``state_dict``, or a bare state dict -- always loaded strictly here, because this module has exactly the reference's 156
keys (a silent ``strict=False`` mismatch is how a wrong architecture goes unnoticed).  ``FusedAdam.state_dict()`` and
``FusedGradScaler.state_dict()`` use torch's layouts, so the optimizer / scaler entries interchange with
``torch.optim.AdamW`` / ``torch.amp.GradScaler`` (tests/test_checkpoint_interop.py).
"""
from __future__ import annotations

import logging
import os
from pathlib import Path

import torch

logger = logging.getLogger(__name__)


def save_checkpoint(model, optimizer, scheduler, scaler, epoch, metrics, config, filename):
    checkpoint = {
        "epoch": epoch,
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": optimizer.state_dict(),
        "scheduler_state_dict": scheduler.state_dict(),
        "scaler_state_dict": scaler.state_dict(),
        "metrics": metrics,
        "config": dict(vars(config)),
    }
    save_path = Path(config.save_dir) / filename
    save_path.parent.mkdir(parents=True, exist_ok=True)
    torch.save(checkpoint, save_path)
    logger.info("Checkpoint saved: %s", save_path)
    return save_path


def extract_model_state_dict(checkpoint):
    """The weights of a reference checkpoint: ``model_state_dict`` (train_advanced.py:478), ``state_dict``, or the object
    itself when it already is a state dict (evaluate_all_models.py:295-298)."""
    if isinstance(checkpoint, dict):
        for key in ("model_state_dict", "state_dict"):
            if key in checkpoint and isinstance(checkpoint[key], dict):
                return checkpoint[key]
    return checkpoint


def load_checkpoint(checkpoint_path, model, device):
    """Load model from checkpoint (test.py:167-188)."""
    logger.info("Loading checkpoint from %s", checkpoint_path)
    if not os.path.exists(checkpoint_path):
        raise FileNotFoundError(f"Checkpoint not found at {checkpoint_path}")
    checkpoint = torch.load(checkpoint_path, map_location=device, weights_only=False)
    model.load_state_dict(extract_model_state_dict(checkpoint))
    if isinstance(checkpoint, dict):
        logger.info("Checkpoint loaded successfully (epoch %s)", checkpoint.get("epoch", "unknown"))
        if checkpoint.get("metrics"):
            logger.info("  - Training metrics: %s", checkpoint["metrics"])
    return model, checkpoint
