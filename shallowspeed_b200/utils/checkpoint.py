"""Per-stage checkpoint / resume.  The reference package has none (only the side script
pickles a state_dict, scripts/DDP_PyTorch_MNIST.py:157); SURVEY.md section 5 lists it as an
auxiliary subsystem.  Format: ``stage{S}of{P}.pt`` holding the flat weight arena plus
the layer dims, so a checkpoint can only be loaded into the same DP x PP layout's stage."""
from __future__ import annotations

from pathlib import Path

import torch


def _path(directory, stage, n_stages):
    return Path(directory) / f"stage{stage}of{n_stages}.pt"


def save_stage(model, directory, stage, n_stages, step=0):
    Path(directory).mkdir(parents=True, exist_ok=True)
    torch.save({"blocks": model.arena.blocks, "weights": model.arena.weights.detach().cpu().clone(),
                "sizes": model.sizes, "step": step}, _path(directory, stage, n_stages))


def load_stage(model, directory, stage, n_stages):
    ck = torch.load(_path(directory, stage, n_stages), map_location="cpu")
    assert [tuple(b) for b in ck["blocks"]] == [tuple(b) for b in model.arena.blocks], "checkpoint/model layout mismatch"
    model.arena.weights.copy_(ck["weights"].to(model.arena.weights.device))
    return ck.get("step", 0)
