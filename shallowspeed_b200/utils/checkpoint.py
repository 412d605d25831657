"""Per-stage checkpoint / resume.  The reference package has none (only the side script
pickles a state_dict, scripts/DDP_PyTorch_MNIST.py:157); SURVEY.md section 5 lists it as an
auxiliary subsystem.  Format: ``stage{S}of{P}.pt`` holding the flat weight arena plus
the layer dims, so a checkpoint can only be loaded into the same DP x PP layout's stage."""
from __future__ import annotations

from pathlib import Path

import torch


def _path(directory, stage, n_stages):
    return Path(directory) / f"stage{stage}of{n_stages}.pt"


def save_stage(model, directory, stage, n_stages, step=0, hparams=None):
    """``hparams`` (lr, global batch, seed mode, schedule ...) travel with the weights so that a resume can refuse to
    continue a run under different settings."""
    Path(directory).mkdir(parents=True, exist_ok=True)
    torch.save({"blocks": model.arena.blocks, "weights": model.arena.weights.detach().cpu().clone(),
                "sizes": model.sizes, "step": step, "hparams": dict(hparams or {})}, _path(directory, stage, n_stages))


def load_stage(model, directory, stage, n_stages, expect_hparams=None):
    """Load this stage's weights; returns the global step the checkpoint was taken at.  Every key of ``expect_hparams``
    that the checkpoint also recorded must match (a resumed run continues the SAME run)."""
    ck = torch.load(_path(directory, stage, n_stages), map_location="cpu")
    assert [tuple(b) for b in ck["blocks"]] == [tuple(b) for b in model.arena.blocks], "checkpoint/model layout mismatch"
    saved = ck.get("hparams", {}) or {}
    for k, v in (expect_hparams or {}).items():
        if k in saved and saved[k] != v:
            raise ValueError(f"checkpoint was written with {k}={saved[k]!r}, this run uses {k}={v!r}")
    model.arena.weights.copy_(ck["weights"].to(model.arena.weights.device))
    return int(ck.get("step", 0))
