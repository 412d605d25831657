"""Pipeline schedules: naive, GPipe, PipeDream-flush (1F1B) and inference.

Parity: ``Schedule`` ABC, ``NaiveParallelSchedule``, ``GPipeSchedule``,
``InferenceSchedule`` reproduce the per-stage instruction streams of the reference
(``shallowspeed/pipe.py:141-294``; streams listed in SURVEY.md section 2.2).
``PipeDreamSchedule`` is a *working* PipeDream-flush / 1F1B schedule - the reference only
ships a stub that raises (pipe.py:297-299).

Design: a schedule only decides the ORDER OF COMPUTE EVENTS per stage - a list of ticks,
each tick a list of ``("F", mubatch)`` / ``("B", mubatch)`` - and the slot every
micro-batch lives in.  One shared lowering (``Schedule.steps``) inserts loads and
stage-boundary communication around the compute events, so all schedules agree on the
comm protocol and a single validator (``parallel.validate``) can prove them
deadlock-free under rendezvous semantics.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

from .instructions import (BackwardGradAcc, BackwardGradAllReduce, Forward, LoadMuBatchInput,
                           LoadMuBatchTarget, OptimizerStep, RecvActivations, RecvOutputGrad,
                           SendActivations, SendInputGrad, ZeroGrad)


class Schedule(ABC):
    training = True

    def __init__(self, num_micro_batches: int, num_stages: int, stage_id: int):
        assert num_micro_batches >= 1 and num_stages >= 1
        assert 0 <= stage_id < num_stages
        self.num_stages = num_stages
        self.stage_id = stage_id
        self.num_micro_batches = num_micro_batches

    # -- what a concrete schedule defines ----------------------------------------
    @abstractmethod
    def compute_ticks(self):
        """List of ticks; each tick is a list of ("F"|"B", mubatch_id)."""

    def slot(self, mubatch_id: int) -> int:
        """Buffer slot a micro-batch occupies while in flight on this stage."""
        return 0

    @property
    def num_slots(self) -> int:
        return 1

    @property
    def num_buffers(self) -> int:
        """Stage-boundary buffers (always even: one input + one output per slot), the
        contract of the reference's ``Schedule.num_buffers`` (pipe.py:159-166)."""
        return 2 * self.num_slots

    # -- helpers (same names as the reference) -------------------------------------
    @property
    def is_first_stage(self):
        return self.stage_id == 0

    @property
    def is_last_stage(self):
        return self.stage_id == self.num_stages - 1

    def is_first_mubatch(self, mubatch_id):
        return mubatch_id == 0

    def is_last_mubatch(self, mubatch_id):
        return mubatch_id == self.num_micro_batches - 1

    def is_valid_stage_id(self, stage_id):
        return 0 <= stage_id < self.num_stages

    # -- shared lowering -----------------------------------------------------------
    def _final_backward(self):
        last = None
        for tick in self.compute_ticks():
            for kind, mu in tick:
                if kind == "B":
                    last = mu
        return last

    def _lower_forward(self, mu):
        b = self.slot(mu)
        cmds = [LoadMuBatchInput(buffer_id=b, mubatch_id=mu) if self.is_first_stage
                else RecvActivations(buffer_id=b)]
        cmds.append(Forward(buffer_id=b, mubatch_id=mu))
        if not self.is_last_stage:
            cmds.append(SendActivations(buffer_id=b))
        return cmds

    def _lower_backward(self, mu, final):
        b = self.slot(mu)
        cmds = [LoadMuBatchTarget(buffer_id=b, mubatch_id=mu) if self.is_last_stage
                else RecvOutputGrad(buffer_id=b)]
        cls = BackwardGradAllReduce if final else BackwardGradAcc
        cmds.append(cls(buffer_id=b, mubatch_id=mu))
        if not self.is_first_stage:
            cmds.append(SendInputGrad(buffer_id=b))
        return cmds

    def steps(self):
        """Generator over ticks (lists of instructions) that process one batch."""
        if self.training:
            yield [ZeroGrad()]
        final = self._final_backward()
        for tick in self.compute_ticks():
            cmds = []
            for kind, mu in tick:
                if kind == "F":
                    cmds += self._lower_forward(mu)
                else:
                    cmds += self._lower_backward(mu, final == mu)
            yield cmds
        if self.training:
            yield [OptimizerStep()]


class NaiveParallelSchedule(Schedule):
    """No interleaving: one micro-batch at a time runs FWD through all stages and then
    BWD back (reference pipe.py:184-222).  Only one stage is busy at any time."""

    def compute_ticks(self):
        return [[("F", mu), ("B", mu)] for mu in range(self.num_micro_batches)]


class GPipeSchedule(Schedule):
    """All forwards, then all backwards in reverse order (reference pipe.py:225-272;
    arXiv 1811.06965).  Every micro-batch is stashed, so slots = num_micro_batches."""

    def compute_ticks(self):
        M = self.num_micro_batches
        return [[("F", mu)] for mu in range(M)] + [[("B", mu)] for mu in reversed(range(M))]

    def slot(self, mubatch_id):
        return mubatch_id

    @property
    def num_slots(self):
        return self.num_micro_batches


class PipeDreamSchedule(Schedule):
    """PipeDream-flush / 1F1B (arXiv 2006.09503, section 3.2).

    Stage s first runs ``w = min(S-1-s, M)`` warm-up forwards, then alternates one
    forward with one backward (steady state), then drains the remaining backwards.  At
    most ``w + 1`` micro-batches are in flight on stage s, so the activation stash is
    bounded by the pipeline depth instead of by M (GPipe).  The batch still ends with a
    flush + one optimizer step, hence the update is identical to GPipe / sequential
    training up to fp32 summation order.  Backwards run in order 0..M-1, so the DP
    all-reduce is fused into the backward of micro-batch M-1.
    """

    @property
    def warmup(self):
        return min(self.num_stages - 1 - self.stage_id, self.num_micro_batches)

    def compute_ticks(self):
        M, w = self.num_micro_batches, self.warmup
        ticks = [[("F", mu)] for mu in range(w)]
        for i in range(M - w):
            ticks.append([("F", w + i), ("B", i)])
        ticks += [[("B", mu)] for mu in range(M - w, M)]
        return ticks

    @property
    def num_slots(self):
        return min(self.warmup + 1, self.num_micro_batches)

    def slot(self, mubatch_id):
        return mubatch_id % self.num_slots


PipeDreamFlushSchedule = PipeDreamSchedule


class InferenceSchedule(Schedule):
    """Forward-only stream used for validation (reference pipe.py:275-294)."""

    training = False

    def compute_ticks(self):
        return [[("F", mu)] for mu in range(self.num_micro_batches)]


SCHEDULE_NAME_TO_CLS = {
    "naive": NaiveParallelSchedule,
    "gpipe": GPipeSchedule,
    "pipedream": PipeDreamSchedule,
    "pipedream-flush": PipeDreamSchedule,
    "1f1b": PipeDreamSchedule,
    "inference": InferenceSchedule,
}
