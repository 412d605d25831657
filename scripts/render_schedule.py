#!/usr/bin/env python
"""Render a pipeline schedule as an SVG Gantt chart (our take on the reference's pebble-graph animation,
.github/assets/PP_pebble_graph.gif): one row per stage, one box per Forward/Backward, placed at the
critical-path times of the validated happens-before DAG (parallel/validate.py:Trace.timeline).

    python scripts/render_schedule.py --schedule pipedream --pp 4 --n-mubatches 8 -o .github/assets/1f1b_pp4_m8.svg
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from shallowspeed_b200.parallel.instructions import BackwardGradAcc, BackwardGradAllReduce, Forward  # noqa: E402
from shallowspeed_b200.parallel.schedules import SCHEDULE_NAME_TO_CLS  # noqa: E402
from shallowspeed_b200.parallel.validate import max_in_flight, validate  # noqa: E402

PALETTE = ["#4e79a7", "#f28e2b", "#59a14f", "#e15759", "#b07aa1", "#76b7b2", "#edc948", "#ff9da7",
           "#9c755f", "#bab0ac", "#1f77b4", "#ff7f0e", "#2ca02c", "#d62728", "#9467bd", "#8c564b"]


def render(name, S, M, fwd=1.0, bwd=2.0, unit=28, row=34):
    cls = SCHEDULE_NAME_TO_CLS[name]
    tr = validate(cls, M, S)

    def cost(_s, ins):
        return fwd if isinstance(ins, Forward) else bwd if isinstance(ins, (BackwardGradAcc, BackwardGradAllReduce)) else 0.0

    tl = tr.timeline(cost)
    span = max(f for _, f in tl.values())
    W, H = int(span * unit) + 150, S * row + 70
    out = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" font-family="monospace" font-size="11">',
           f'<rect width="{W}" height="{H}" fill="white"/>',
           f'<text x="8" y="16" font-size="13">{name}  pp={S}  micro-batches={M}  makespan={span:g}  '
           f'bubble={tr.bubble_fraction(cost):.3f}  max in-flight (stage 0)={max_in_flight(cls(M, S, 0))}</text>']
    for s in range(S):
        y = 30 + s * row
        out.append(f'<text x="8" y="{y + 18}">stage {s}</text>')
        out.append(f'<rect x="70" y="{y}" width="{int(span * unit)}" height="{row - 6}" fill="#f2f2f2"/>')
        for it in tr.items[s]:
            ins = it.instrs[0]
            if it.kind != "compute" or not isinstance(ins, (Forward, BackwardGradAcc, BackwardGradAllReduce)):
                continue
            a, b = tl[(s, it.index)]
            col = PALETTE[ins.mubatch_id % len(PALETTE)]
            is_f = isinstance(ins, Forward)
            label = ("F" if is_f else "B") + str(ins.mubatch_id) + ("*" if isinstance(ins, BackwardGradAllReduce) else "")
            out.append(f'<rect x="{70 + a * unit:.1f}" y="{y}" width="{(b - a) * unit - 1:.1f}" height="{row - 6}" '
                       f'fill="{col}" fill-opacity="{1.0 if is_f else 0.55}" stroke="#333" stroke-width="0.5"/>')
            out.append(f'<text x="{70 + a * unit + 3:.1f}" y="{y + 18}" fill="black">{label}</text>')
    out.append(f'<text x="8" y="{H - 10}" fill="#555">F = forward, B = backward (B* also runs the data-parallel reduction); '
               f'time unit = one forward</text>')
    out.append("</svg>")
    return "\n".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--schedule", default="gpipe", choices=sorted(SCHEDULE_NAME_TO_CLS))
    ap.add_argument("--pp", type=int, default=4)
    ap.add_argument("--n-mubatches", type=int, default=8)
    ap.add_argument("-o", "--out", default="-")
    a = ap.parse_args()
    svg = render(a.schedule, a.pp, a.n_mubatches)
    if a.out == "-":
        print(svg)
    else:
        with open(a.out, "w") as f:
            f.write(svg)


if __name__ == "__main__":
    main()
