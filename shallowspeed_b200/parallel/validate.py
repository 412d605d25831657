"""Static validation of pipeline schedules: happens-before, buffer liveness, deadlock.

The reference's own test file asks for exactly this ("define a happens-before
predicate", ``tests/test_schedules.py:4-10``) and ships no multi-stage check at all.
Because our runtime issues sends/receives asynchronously on side streams and the fused
DP kernel spins on peer flags, a schedule bug is a hang on 8 GPUs - so every schedule
is proven safe on the CPU first.

Model
-----
* Each stage executes its flattened instruction stream in order.
* Maximal runs of consecutive communication instructions form one *group*
  (``ncclGroupStart/End`` on the GPU, ``batch_isend_irecv`` on the CPU path).
* Rendezvous semantics (the conservative model of NCCL p2p): the n-th send a->b
  completes only once b has *posted* its n-th receive from a, and vice versa; a stage
  cannot move past a group until every operation in it has completed.
* Data flow: Forward(mu) needs its input (load or receive) in the slot; Backward(mu)
  needs Forward(mu) done, its output gradient in the slot, and must run exactly once;
  a slot may only be re-filled after the backward of the previous occupant.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

from .instructions import (BackwardGradAcc, BackwardGradAllReduce, COMM_INSTRUCTIONS, Forward,
                           LoadMuBatchInput, LoadMuBatchTarget, OptimizerStep, RecvActivations,
                           RecvOutputGrad, SendActivations, SendInputGrad, ZeroGrad, flatten)


class ScheduleError(AssertionError):
    pass


@dataclass
class _Item:
    kind: str  # "compute" | "group"
    instrs: list
    index: int = 0


def _itemize(stream):
    """Coalesce consecutive comm instructions into groups."""
    items, run = [], []
    for ins in stream:
        if isinstance(ins, COMM_INSTRUCTIONS):
            run.append(ins)
        else:
            if run:
                items.append(_Item("group", run))
                run = []
            items.append(_Item("compute", [ins]))
    if run:
        items.append(_Item("group", run))
    for i, it in enumerate(items):
        it.index = i
    return items


def comm_groups(stream):
    """Public helper: the comm groups of a flattened stream (used by the executors)."""
    return [it.instrs for it in _itemize(stream) if it.kind == "group"]


def _peer(stage, ins):
    if isinstance(ins, (SendActivations, RecvOutputGrad)):
        return stage + 1
    return stage - 1


@dataclass
class Trace:
    """Result of a successful simulation: one legal total order of (stage, instr) plus the
    happens-before PARTIAL order every legal execution must respect.

    The partial order is a DAG over items (one item = one compute instruction or one comm group):
      * program order:   (s, i) -> (s, i+1)
      * rendezvous:      a group cannot complete before its peer has POSTED the matching operation,
                         i.e. before the peer finished the item preceding it:
                         (peer, J-1) -> (s, I)   and   (s, I-1) -> (peer, J)
      * data:            a receive completes only after the matching send was posted (same edges).
    ``happens_before(a, b)`` is reachability in that DAG, so it holds in EVERY execution, not just in
    the simulated linearisation."""

    order: list = field(default_factory=list)
    items: list = field(default_factory=list)          # items[s] = list of _Item
    edges: dict = field(default_factory=dict)          # (s, i) -> set of (s', i')

    def position(self, stage, predicate):
        for i, (s, ins) in enumerate(self.order):
            if s == stage and predicate(ins):
                return i
        raise KeyError("instruction not found in trace")

    def _locate(self, stage, predicate):
        for it in self.items[stage]:
            for pos, ins in enumerate(it.instrs):
                if predicate(ins):
                    return it.index, pos
        raise KeyError("instruction not found in schedule")

    def _reachable(self, src, dst):
        if src == dst:
            return True
        seen, stack = {src}, [src]
        while stack:
            n = stack.pop()
            for m in self.edges.get(n, ()):
                if m == dst:
                    return True
                if m not in seen:
                    seen.add(m)
                    stack.append(m)
        return False

    def happens_before(self, a, b):
        """a, b = (stage, predicate).  True iff a precedes b in every legal execution."""
        (sa, pa), (sb, pb) = a, b
        ia, posa = self._locate(sa, pa)
        ib, posb = self._locate(sb, pb)
        if sa == sb and ia == ib:
            return posa < posb
        if not self.items:
            return self.position(*a) < self.position(*b)
        # "a's item completes before b's item completes".  A compute item's only incoming edge is program
        # order, so for compute b this is the same as "a completes before b starts".
        return self._reachable((sa, ia), (sb, ib))

    def timeline(self, cost=None):
        """Earliest-finish (critical path) times of every item under a cost model.

        ``cost(stage, instr) -> float`` gives the duration of a compute instruction (default: Forward 1,
        Backward 2, everything else 0); comm groups take no time of their own but inherit the rendezvous
        dependencies.  Returns ``{(stage, item_index): (start, finish)}`` - the schedule an ideal executor with
        infinitely fast links would achieve, which is what the bubble fractions quoted for GPipe / 1F1B
        assume."""
        if cost is None:
            def cost(_s, ins):
                if isinstance(ins, Forward):
                    return 1.0
                if isinstance(ins, (BackwardGradAcc, BackwardGradAllReduce)):
                    return 2.0
                return 0.0
        preds = {}
        for u, vs in self.edges.items():
            for v in vs:
                preds.setdefault(v, []).append(u)
        finish, out = {}, {}
        # the simulated total order is a topological order of the item DAG
        seen = []
        for s, ins in self.order:
            for it in self.items[s]:
                if it.instrs[0] is ins:
                    seen.append((s, it.index))
        pending = list(seen)
        guard = 0
        while pending:
            guard += 1
            if guard > 4 * len(seen) * len(seen) + 16:
                raise ScheduleError("timeline: dependency cycle")
            node = pending.pop(0)
            ps = preds.get(node, [])
            if any(p not in finish for p in ps):
                pending.append(node)
                continue
            start = max([finish[p] for p in ps], default=0.0)
            it = self.items[node[0]][node[1]]
            dur = sum(cost(node[0], i) for i in it.instrs) if it.kind == "compute" else 0.0
            finish[node] = start + dur
            out[node] = (start, start + dur)
        return out

    def makespan(self, cost=None):
        return max(f for _, f in self.timeline(cost).values())

    def bubble_fraction(self, cost=None):
        """1 - (busy time of the busiest stage) / makespan under the cost model."""
        tl = self.timeline(cost)
        span = max(f for _, f in tl.values())
        busy = {}
        for (s, _i), (a, b) in tl.items():
            busy[s] = busy.get(s, 0.0) + (b - a)
        return 1.0 - max(busy.values()) / span if span > 0 else 0.0

    def concurrent(self, a, b):
        """neither a before b nor b before a: the two may overlap in time."""
        return not self.happens_before(a, b) and not self.happens_before(b, a)


def simulate(schedules: Sequence, check_dataflow: bool = True) -> Trace:
    """Run all stages of one pipeline against each other.  Raises ScheduleError on
    deadlock, unmatched messages or a data-flow violation; returns the trace."""
    S = len(schedules)
    for s, sc in enumerate(schedules):
        if sc.stage_id != s or sc.num_stages != S:
            raise ScheduleError("schedules must be given in stage order for one pipeline")
    streams = [flatten(list(sc.steps())) for sc in schedules]
    items = [_itemize(st) for st in streams]
    ptr = [0] * S

    # message sequence numbers: n-th send a->b pairs with n-th recv at b from a
    send_seq = [dict() for _ in range(S)]   # stage -> {(item_idx, pos): (peer, n)}
    recv_seq = [dict() for _ in range(S)]
    send_cnt, recv_cnt = {}, {}
    send_at, recv_at = {}, {}               # (src, dst, n) -> item index where posted
    for s in range(S):
        for it in items[s]:
            if it.kind != "group":
                continue
            for ins in it.instrs:
                p = _peer(s, ins)
                if not (0 <= p < S):
                    raise ScheduleError(f"stage {s}: {ins} addresses stage {p} outside the pipeline")
                if isinstance(ins, (SendActivations, SendInputGrad)):
                    n = send_cnt.get((s, p), 0)
                    send_cnt[(s, p)] = n + 1
                    send_at[(s, p, n)] = it.index
                else:
                    n = recv_cnt.get((p, s), 0)
                    recv_cnt[(p, s)] = n + 1
                    recv_at[(p, s, n)] = it.index
    for key, n in send_cnt.items():
        if recv_cnt.get(key, 0) != n:
            raise ScheduleError(f"{n} sends {key[0]}->{key[1]} but {recv_cnt.get(key, 0)} receives")
    for key, n in recv_cnt.items():
        if send_cnt.get(key, 0) != n:
            raise ScheduleError(f"{n} receives {key[0]}->{key[1]} but {send_cnt.get(key, 0)} sends")

    def group_ready(s, it):
        seen_s, seen_r = {}, {}
        for ins in it.instrs:
            p = _peer(s, ins)
            if isinstance(ins, (SendActivations, SendInputGrad)):
                # which n is this?  count sends to p in earlier items + earlier in this group
                base = sum(1 for k, idx in send_at.items() if k[0] == s and k[1] == p and idx < it.index)
                n = base + seen_s.get(p, 0)
                seen_s[p] = seen_s.get(p, 0) + 1
                if ptr[p] < recv_at[(s, p, n)]:
                    return False
            else:
                base = sum(1 for k, idx in recv_at.items() if k[0] == p and k[1] == s and idx < it.index)
                n = base + seen_r.get(p, 0)
                seen_r[p] = seen_r.get(p, 0) + 1
                if ptr[p] < send_at[(p, s, n)]:
                    return False
        return True

    trace = Trace()
    trace.items = items
    edges = {}

    def add_edge(u, v):
        if u[1] >= 0:
            edges.setdefault(u, set()).add(v)

    for st in range(S):
        for it in items[st]:
            if it.index + 1 < len(items[st]):
                add_edge((st, it.index), (st, it.index + 1))
    for (src, dst, nmsg), i_send in send_at.items():
        j_recv = recv_at[(src, dst, nmsg)]
        add_edge((dst, j_recv - 1), (src, i_send))      # the send completes only once the receive is posted
        add_edge((src, i_send - 1), (dst, j_recv))      # the receive completes only once the send is posted
    trace.edges = edges
    total = sum(len(x) for x in items)
    done = 0
    while done < total:
        progressed = False
        for s in range(S):
            while ptr[s] < len(items[s]):
                it = items[s][ptr[s]]
                if it.kind == "group" and not group_ready(s, it):
                    break
                for ins in it.instrs:
                    trace.order.append((s, ins))
                ptr[s] += 1
                done += 1
                progressed = True
        if not progressed:
            where = {s: (items[s][ptr[s]].instrs if ptr[s] < len(items[s]) else "done") for s in range(S)}
            raise ScheduleError(f"deadlock under rendezvous semantics; blocked at {where}")

    if check_dataflow:
        for s, sc in enumerate(schedules):
            _check_stage_dataflow(sc, streams[s])
        _check_message_payloads(schedules, streams)
    return trace


def _check_stage_dataflow(sc, stream):
    M, s = sc.num_micro_batches, sc.stage_id
    first, last = sc.is_first_stage, sc.is_last_stage
    in_slot, gout_slot, out_slot, gin_slot = {}, {}, {}, {}   # slot -> mubatch (or tag)
    live = {}            # slot -> mubatch whose stash is alive (forward done, backward pending)
    fwd_done, bwd_done = set(), set()
    n_zero = n_opt = 0
    pending_recv_act = []  # order in which activations arrive = mubatch order of the previous stage's sends
    allreduce_seen = None
    for pos, ins in enumerate(stream):
        if isinstance(ins, ZeroGrad):
            if pos != 0:
                raise ScheduleError(f"stage {s}: ZeroGrad must be the first instruction")
            n_zero += 1
        elif isinstance(ins, OptimizerStep):
            if pos != len(stream) - 1:
                raise ScheduleError(f"stage {s}: OptimizerStep must be the last instruction")
            n_opt += 1
        elif isinstance(ins, LoadMuBatchInput):
            if not first:
                raise ScheduleError(f"stage {s}: only the first stage loads inputs")
            if ins.buffer_id in live:
                raise ScheduleError(f"stage {s}: slot {ins.buffer_id} overwritten while mubatch {live[ins.buffer_id]} is in flight")
            in_slot[ins.buffer_id] = ins.mubatch_id
        elif isinstance(ins, RecvActivations):
            if first:
                raise ScheduleError(f"stage {s}: first stage cannot receive activations")
            if ins.buffer_id in live:
                raise ScheduleError(f"stage {s}: slot {ins.buffer_id} overwritten while mubatch {live[ins.buffer_id]} is in flight")
            in_slot[ins.buffer_id] = "recv"
        elif isinstance(ins, Forward):
            if ins.buffer_id not in in_slot:
                raise ScheduleError(f"stage {s}: Forward({ins.mubatch_id}) without input in slot {ins.buffer_id}")
            if first and in_slot[ins.buffer_id] != ins.mubatch_id:
                raise ScheduleError(f"stage {s}: Forward({ins.mubatch_id}) reads input of mubatch {in_slot[ins.buffer_id]}")
            if ins.mubatch_id in fwd_done:
                raise ScheduleError(f"stage {s}: Forward({ins.mubatch_id}) issued twice")
            if sc.slot(ins.mubatch_id) != ins.buffer_id:
                raise ScheduleError(f"stage {s}: Forward({ins.mubatch_id}) uses slot {ins.buffer_id}, schedule says {sc.slot(ins.mubatch_id)}")
            fwd_done.add(ins.mubatch_id)
            del in_slot[ins.buffer_id]
            if sc.training:
                live[ins.buffer_id] = ins.mubatch_id
            out_slot[ins.buffer_id] = ins.mubatch_id
        elif isinstance(ins, SendActivations):
            if last:
                raise ScheduleError(f"stage {s}: last stage cannot send activations")
            if ins.buffer_id not in out_slot:
                raise ScheduleError(f"stage {s}: SendActivations from empty slot {ins.buffer_id}")
            del out_slot[ins.buffer_id]
        elif isinstance(ins, LoadMuBatchTarget):
            if not last:
                raise ScheduleError(f"stage {s}: only the last stage loads targets")
            gout_slot[ins.buffer_id] = ins.mubatch_id
        elif isinstance(ins, RecvOutputGrad):
            if last:
                raise ScheduleError(f"stage {s}: last stage cannot receive output grads")
            gout_slot[ins.buffer_id] = "recv"
        elif isinstance(ins, (BackwardGradAcc, BackwardGradAllReduce)):
            mu, b = ins.mubatch_id, ins.buffer_id
            if mu not in fwd_done:
                raise ScheduleError(f"stage {s}: Backward({mu}) before Forward({mu})")
            if mu in bwd_done:
                raise ScheduleError(f"stage {s}: Backward({mu}) issued twice")
            if live.get(b) != mu:
                raise ScheduleError(f"stage {s}: Backward({mu}) but slot {b} stashes {live.get(b)}")
            if b not in gout_slot:
                raise ScheduleError(f"stage {s}: Backward({mu}) without output grad / target in slot {b}")
            if last and gout_slot[b] != mu:
                raise ScheduleError(f"stage {s}: Backward({mu}) uses target of mubatch {gout_slot[b]}")
            if allreduce_seen is not None:
                raise ScheduleError(f"stage {s}: Backward({mu}) after the all-reduce backward")
            if isinstance(ins, BackwardGradAllReduce):
                allreduce_seen = mu
            bwd_done.add(mu)
            del gout_slot[b]
            del live[b]
            gin_slot[b] = mu
        elif isinstance(ins, SendInputGrad):
            if first:
                raise ScheduleError(f"stage {s}: first stage cannot send input grads")
            if ins.buffer_id not in gin_slot:
                raise ScheduleError(f"stage {s}: SendInputGrad from empty slot {ins.buffer_id}")
            del gin_slot[ins.buffer_id]
    if fwd_done != set(range(M)):
        raise ScheduleError(f"stage {s}: forwards {sorted(fwd_done)} != all {M} micro-batches")
    if sc.training:
        if bwd_done != set(range(M)):
            raise ScheduleError(f"stage {s}: backwards {sorted(bwd_done)} != all {M} micro-batches")
        if allreduce_seen is None:
            raise ScheduleError(f"stage {s}: no BackwardGradAllReduce in the step")
        if n_zero != 1 or n_opt != 1:
            raise ScheduleError(f"stage {s}: need exactly one ZeroGrad and one OptimizerStep")
        if live:
            raise ScheduleError(f"stage {s}: slots still alive at the end: {live}")
    if len(set(sc.slot(m) for m in range(M))) > sc.num_slots:
        raise ScheduleError(f"stage {s}: uses more slots than num_slots={sc.num_slots}")


def _mubatch_order(stream, send_cls, compute_cls):
    """mubatch ids in the order their payloads are sent (send follows its compute)."""
    order, last = [], {}
    for ins in stream:
        if isinstance(ins, compute_cls):
            last[ins.buffer_id] = ins.mubatch_id
        elif isinstance(ins, send_cls):
            order.append(last[ins.buffer_id])
    return order


def _recv_consumer_order(stream, recv_cls, compute_cls):
    """mubatch ids in the order received payloads are consumed."""
    order, waiting = [], {}
    idx = 0
    for ins in stream:
        if isinstance(ins, recv_cls):
            waiting[ins.buffer_id] = idx
            order.append(None)
            idx += 1
        elif isinstance(ins, compute_cls) and ins.buffer_id in waiting:
            order[waiting.pop(ins.buffer_id)] = ins.mubatch_id
    return order


def _check_message_payloads(schedules, streams):
    """The n-th activation sent by stage s must be consumed as the same micro-batch by
    stage s+1 (and likewise for gradients going back)."""
    S = len(schedules)
    bwd = (BackwardGradAcc, BackwardGradAllReduce)
    for s in range(S - 1):
        sent = _mubatch_order(streams[s], SendActivations, Forward)
        used = _recv_consumer_order(streams[s + 1], RecvActivations, Forward)
        if sent != used:
            raise ScheduleError(f"activation order mismatch {s}->{s+1}: sent {sent}, consumed {used}")
        if schedules[s].training:
            sent = _mubatch_order(streams[s + 1], SendInputGrad, bwd)
            used = _recv_consumer_order(streams[s], RecvOutputGrad, bwd)
            if sent != used:
                raise ScheduleError(f"gradient order mismatch {s+1}->{s}: sent {sent}, consumed {used}")


def validate(schedule_cls, num_micro_batches: int, num_stages: int) -> Trace:
    """Validate a schedule class for a pipeline shape; returns the trace."""
    scheds = [schedule_cls(num_micro_batches, num_stages, s) for s in range(num_stages)]
    return simulate(scheds)


def max_in_flight(schedule) -> int:
    """Peak number of stashed micro-batches (forward done, backward pending)."""
    peak = cur = 0
    for ins in flatten(list(schedule.steps())):
        if isinstance(ins, Forward):
            cur += 1
            peak = max(peak, cur)
        elif isinstance(ins, (BackwardGradAcc, BackwardGradAllReduce)):
            cur -= 1
    return peak


# ------------------------------------------------------------------------------------------------------------
# Stream-level model of the PEER-MEMORY pipeline transport (csrc/runtime/pipe_engine.cpp, pp_ctx_ branch of
# plan_per_mubatch): the native executor lowers a stage's instruction stream to ops on CUDA streams; with the one-sided
# transport a comm group becomes, on ONE comm stream, [wait for the producing compute event, push] per send followed by
# one flag wait per receive.  Pushes additionally wait for the consumer's credit of the previous step.  This simulation
# replays that lowering for all stages over several steps and reports a deadlock if the stream programs can block each
# other (e.g. a flag wait queued in front of a push the neighbour is waiting for).
# ------------------------------------------------------------------------------------------------------------
def simulate_one_sided(schedules: Sequence, n_steps: int = 3, n_mu_streams: int = 4, sends_first: bool = True,
                       layout: str = "comm_stream") -> dict:
    """``sends_first=False`` models the WRONG lowering (flag waits queued in front of the pushes of the same group) and
    exists so the tests can show that this model does detect the resulting cross-stage deadlock.

    ``layout`` selects how the executor places the transport ops:
      * ``"comm_stream"``  - every push / flag wait on the stage's one communication stream (big boundary tiles);
      * ``"mubatch"``      - small tiles (the default lowering since round 2): the push sits on the stream of the micro-batch
        that produced the tile, right behind its compute op, the flag wait on the stream of the micro-batch that consumes
        it, right in front of its compute op.  The FOLDED lowering (wait and push inside the chain kernel launch) is the
        same program with [wait, compute, push] fused into one launch, so this layout models it as well."""
    assert layout in ("comm_stream", "mubatch")
    S = len(schedules)
    progs = []          # per stage: {stream name: [op, ...]}, op = (kind, payload)
    for s, sc in enumerate(schedules):
        stream = flatten(list(sc.steps()))
        streams = {"comm": []}
        ev_in, ev_gout, ev_fwd, ev_bwd = {}, {}, {}, {}
        n_ev = 0
        for it in _itemize(stream):
            if it.kind == "group" and layout == "mubatch":
                mu_of = _resolve_group_mubatches(stream, it)
                for ins in it.instrs:
                    mu = mu_of[id(ins)]
                    ops = streams.setdefault(f"c{mu % n_mu_streams}", [])
                    if isinstance(ins, (SendActivations, SendInputGrad)):
                        is_act = isinstance(ins, SendActivations)
                        ops.append(("push", ("act" if is_act else "dz", s + 1 if is_act else s - 1, mu)))
                    else:
                        ops.append(("wait_flag", ("act" if isinstance(ins, RecvActivations) else "dz", s, mu)))
                continue
            if it.kind == "group":
                recvs = []
                sends = [i for i in it.instrs if isinstance(i, (SendActivations, SendInputGrad))]
                rest = [i for i in it.instrs if not isinstance(i, (SendActivations, SendInputGrad))]
                mu_of = _resolve_group_mubatches(stream, it)
                def emit_waits():
                    for ins in rest:
                        mu = mu_of[id(ins)]
                        is_act = isinstance(ins, RecvActivations)
                        streams["comm"].append(("wait_flag", ("act" if is_act else "dz", s, mu)))
                        recvs.append((mu, is_act))

                if not sends_first:
                    emit_waits()
                for ins in sends:
                    mu = mu_of[id(ins)]
                    is_act = isinstance(ins, SendActivations)
                    dep = (ev_fwd if is_act else ev_bwd).get(mu)
                    if dep is not None:
                        streams["comm"].append(("wait_event", dep))
                    streams["comm"].append(("push", ("act" if is_act else "dz", s + 1 if is_act else s - 1, mu)))
                if sends_first:
                    emit_waits()
                if recvs:
                    ev = (s, n_ev); n_ev += 1
                    streams["comm"].append(("record", ev))
                    for mu, is_act in recvs:
                        (ev_in if is_act else ev_gout)[mu] = ev
                continue
            ins = it.instrs[0]
            if isinstance(ins, (Forward, BackwardGradAcc, BackwardGradAllReduce)):
                mu = ins.mubatch_id
                name = f"c{mu % n_mu_streams}"
                ops = streams.setdefault(name, [])
                is_f = isinstance(ins, Forward)
                dep = (ev_in if is_f else ev_gout).pop(mu, None)
                if dep is not None:
                    ops.append(("wait_event", dep))
                ops.append(("compute", ("F" if is_f else "B", mu)))
                ev = (s, n_ev); n_ev += 1
                ops.append(("record", ev))
                (ev_fwd if is_f else ev_bwd)[mu] = ev
        progs.append(streams)

    arrived = {}                          # (kind, dst_stage, mu) -> last step whose data landed
    credit = {}                           # (kind, producer_stage) -> last step whose slots the consumer released
    finished_steps = [0] * S
    stats = {"pushes": 0, "waits": 0}
    for step in range(1, n_steps + 1):
        pcs = [{name: 0 for name in progs[s]} for s in range(S)]
        events = set()
        active = [True] * S               # stages still inside this step
        while any(active):
            progressed = False
            for s in range(S):
                if not active[s]:
                    continue
                for name, ops in progs[s].items():
                    while pcs[s][name] < len(ops):
                        kind, arg = ops[pcs[s][name]]
                        if kind == "wait_event" and (step, arg) not in events:
                            break
                        if kind == "wait_flag" and arrived.get(arg, 0) < step:
                            break
                        if kind == "push":
                            k, dst, mu = arg
                            if not (0 <= dst < S):
                                raise ScheduleError(f"stage {s} pushes to stage {dst}")
                            if credit.get((k, s), 0) < step - 1:
                                break                      # the consumer still owns the slots of the previous step
                            arrived[(k, dst, mu)] = step
                            stats["pushes"] += 1
                        if kind == "wait_flag":
                            stats["waits"] += 1
                        if kind == "record":
                            events.add((step, arg))
                        pcs[s][name] += 1
                        progressed = True
                if all(pcs[s][n] == len(o) for n, o in progs[s].items()):
                    # step finished on every stream: hand the receive slots back to the producers
                    active[s] = False
                    finished_steps[s] = step
                    if s > 0:
                        credit[("act", s - 1)] = step
                    if s + 1 < S:
                        credit[("dz", s + 1)] = step
                    progressed = True
            if not progressed:
                where = {s: {n: (o[pcs[s][n]] if pcs[s][n] < len(o) else "done") for n, o in progs[s].items()}
                         for s in range(S) if active[s]}
                raise ScheduleError(f"deadlock in the one-sided transport model at step {step}: {where}")
    stats["steps"] = n_steps
    return stats


def _resolve_group_mubatches(stream, item):
    """micro-batch carried by every comm instruction of ``item`` (same rule as PipeEngine::build: a receive belongs to
    the next compute instruction on its buffer, a send to the previous one)."""
    pos = {id(ins): i for i, ins in enumerate(stream)}
    out = {}
    for ins in item.instrs:
        i = pos[id(ins)]
        if isinstance(ins, (RecvActivations, RecvOutputGrad)):
            want = Forward if isinstance(ins, RecvActivations) else (BackwardGradAcc, BackwardGradAllReduce)
            rng = range(i + 1, len(stream))
        else:
            want = Forward if isinstance(ins, SendActivations) else (BackwardGradAcc, BackwardGradAllReduce)
            rng = range(i - 1, -1, -1)
        for j in rng:
            c = stream[j]
            if isinstance(c, want) and c.buffer_id == ins.buffer_id:
                out[id(ins)] = c.mubatch_id
                break
        else:
            raise ScheduleError(f"cannot resolve the micro-batch of {ins}")
    return out
