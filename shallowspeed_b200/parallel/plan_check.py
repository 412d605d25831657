"""Self-check of a lowered step plan (``PipeEngine.plan_text()``): the scheduler-side race check SURVEY.md section 5 asks
for.  The plan is a list of ops on streams with record / wait edges between them; a step is only well-formed if

  * every wait refers to an event that an EARLIER op of the plan records (otherwise the wait is a no-op on a stale
    event and the dependency it was meant to express does not exist - a silent race);
  * no event is recorded twice (a second record would retarget earlier waits under graph capture);
  * every side stream is forked from work that is already ordered after the step's begin (its first op is a wait) and is
    joined back: its last op is a record that stream 0 eventually waits for (otherwise the step could "finish" on the main
    stream while side-stream work is still running - and graph capture would reject the plan anyway).
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import List


@dataclass
class PlanOp:
    index: int
    name: str
    stream: int
    event: int = -1
    layer: int = -1
    mu: int = -1


class PlanError(AssertionError):
    pass


_FIELD = re.compile(r"(stream|event|layer|mu)=(-?\d+)")


def parse_plan(text: str) -> List[PlanOp]:
    ops = []
    for line in text.splitlines():
        line = line.strip()
        if not line:
            continue
        head = line.split("[", 1)[0].split()
        op = PlanOp(index=int(head[0]), name=head[1], stream=-1)
        for key, val in _FIELD.findall(line.split("[", 1)[0]):
            setattr(op, key, int(val))
        if op.stream < 0:
            raise PlanError(f"op without a stream: {line!r}")
        ops.append(op)
    return ops


def check_plan(text: str) -> dict:
    """Raises PlanError on a malformed plan; returns summary statistics otherwise."""
    ops = parse_plan(text)
    recorded = {}          # event -> (op index, stream)
    waited = {}            # event -> list of waiting streams
    first_op, last_op = {}, {}
    for op in ops:
        first_op.setdefault(op.stream, op)
        last_op[op.stream] = op
        if op.name == "record_event":
            if op.event in recorded:
                raise PlanError(f"event {op.event} recorded twice (ops {recorded[op.event][0]} and {op.index})")
            recorded[op.event] = (op.index, op.stream)
        elif op.name == "wait_event":
            if op.event not in recorded:
                raise PlanError(f"op {op.index}: stream {op.stream} waits for event {op.event} before anything records it")
            if recorded[op.event][1] == op.stream:
                raise PlanError(f"op {op.index}: stream {op.stream} waits for its own event {op.event}")
            waited.setdefault(op.event, []).append(op.stream)
    side = sorted(s for s in first_op if s != 0)
    for s in side:
        if first_op[s].name != "wait_event":
            raise PlanError(f"stream {s} starts with {first_op[s].name} (op {first_op[s].index}) without being forked from ordered work")
        if last_op[s].name != "record_event":
            raise PlanError(f"stream {s} ends with {last_op[s].name} (op {last_op[s].index}) and is never joined")
        if 0 not in waited.get(last_op[s].event, []):
            raise PlanError(f"stream {s}: its final event {last_op[s].event} is not waited for by the main stream")
    kernels = [o for o in ops if o.name not in ("record_event", "wait_event")]
    return {"ops": len(ops), "kernels_and_copies": len(kernels), "streams": 1 + len(side), "events": len(recorded),
            "cross_stream_edges": sum(len(v) for v in waited.values())}
