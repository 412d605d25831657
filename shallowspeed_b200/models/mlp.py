"""MLP model family + pipeline-stage partitioner.

Parity with the reference's ``MLP(sizes, stage_idx, n_stages, batch_size)``
(``shallowspeed/layers.py:236-270``): ``len(sizes) % n_stages == 0``; stage *s* owns
``sizes[s*k : s*k+k+1]`` (k = len(sizes)//n_stages), i.e. k Linears, the last stage
k-1 Linears + Softmax + MSELoss; the final Linear carries no ReLU only when it sits on
the last stage; ``in_dim`` / ``out_dim`` are exposed for buffer allocation.

All Linears of a stage share one ``ParamArena`` (flat weights + flat grads).
"""
from __future__ import annotations

from typing import Optional, Sequence

from .layers import Linear, MSELoss, ParamArena, Sequential, Softmax

DEFAULT_LAYER_SIZES = [784, 128, 127, 126, 125, 124, 123, 10]  # reference train.py:98


def stage_sizes(sizes: Sequence[int], stage_idx: int, n_stages: int):
    """The slice of ``sizes`` owned by ``stage_idx`` (reference layers.py:242-250)."""
    assert len(sizes) % n_stages == 0, (
        f"len(sizes)={len(sizes)} must be divisible by the number of pipeline stages {n_stages}"
    )
    assert 0 <= stage_idx < n_stages
    k = len(sizes) // n_stages
    return list(sizes[stage_idx * k : min(len(sizes), k * stage_idx + k + 1)])


def stage_layer_specs(sizes: Sequence[int], stage_idx: int, n_stages: int):
    """[(in, out, relu, global_layer_index)] for the Linears of a stage."""
    local = stage_sizes(sizes, stage_idx, n_stages)
    is_last = stage_idx == n_stages - 1
    k = len(sizes) // n_stages
    specs = []
    for i in range(len(local) - 1):
        relu = not (i == len(local) - 2 and is_last)
        specs.append((local[i], local[i + 1], relu, stage_idx * k + i))
    return specs


def mlp_sizes(hidden: int, n_layers: int, in_dim: int = 784, out_dim: int = 10):
    """``n_layers`` Linears: in_dim -> hidden x (n_layers-1) -> out_dim."""
    assert n_layers >= 1
    return [in_dim] + [hidden] * (n_layers - 1) + [out_dim]


class MLP(Sequential):
    def __init__(self, sizes: Sequence[int], stage_idx: int, n_stages: int, batch_size: int,
                 device="cpu", seed_mode: str = "shape", verbose: bool = False):
        """
        :param batch_size: the GLOBAL batch size - the loss gradient is scaled by
            1/batch_size so that summing over micro-batches and DP replicas reproduces
            sequential training (reference layers.py:237-241).
        :param seed_mode: "shape" = reference-identical shape-seeded init;
            "index" = additionally mixes the global layer index into the seed.
        """
        assert seed_mode in ("shape", "index")
        self.sizes = list(sizes)
        self.stage_idx, self.n_stages, self.batch_size = stage_idx, n_stages, batch_size
        specs = stage_layer_specs(sizes, stage_idx, n_stages)
        local = stage_sizes(sizes, stage_idx, n_stages)
        self.is_last_stage = stage_idx == n_stages - 1
        self.arena = ParamArena([(o, i) for i, o, _, _ in specs], device=device)
        layers = [
            Linear(i, o, activation="relu" if relu else None, arena=self.arena, block_index=bi,
                   layer_index=gidx if seed_mode == "index" else None)
            for bi, (i, o, relu, gidx) in enumerate(specs)
        ]
        self.linears = list(layers)
        if self.is_last_stage:
            layers.append(Softmax())
            layers.append(MSELoss(batch_size=batch_size))
        super().__init__(layers)
        if verbose:
            print(layers)
        self.in_dim = local[0]
        self.out_dim = local[-1]  # softmax & loss keep the width

    def to(self, device):
        self.arena.to(device)
        return self

    @property
    def loss_layer(self) -> Optional[MSELoss]:
        return self.layers[-1] if self.is_last_stage else None
