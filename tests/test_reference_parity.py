"""Direct parity checks against the reference's own code (mounted read-only at /root/reference,
imported with the mpi4py shim on the path): instruction streams of every schedule it implements, the
model partitioner, and the dataset sharding arithmetic."""
import contextlib
import io
import os
import sys

import pytest

REF = "/root/reference"
SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "mpi_shim")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "shallowspeed")), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, SHIM)
    sys.path.insert(0, REF)
    try:
        import shallowspeed.layers as rlayers
        import shallowspeed.pipe as rpipe
    finally:
        sys.path.remove(REF)
        sys.path.remove(SHIM)
    return rpipe, rlayers


def _stream(sched):
    return [[(type(i).__name__, getattr(i, "mubatch_id", None)) for i in tick] for tick in sched.steps()]


@pytest.mark.parametrize("M", [1, 2, 4, 7])
@pytest.mark.parametrize("S", [1, 2, 4, 8])
def test_instruction_streams_equal_the_reference(ref, M, S):
    rpipe, _ = ref
    from shallowspeed_b200 import pipe

    for ours, theirs in ((pipe.NaiveParallelSchedule, rpipe.NaiveParallelSchedule),
                         (pipe.GPipeSchedule, rpipe.GPipeSchedule), (pipe.InferenceSchedule, rpipe.InferenceSchedule)):
        for stage in range(S):
            a, b = _stream(ours(M, S, stage)), _stream(theirs(M, S, stage))
            assert a == b, (ours.__name__, M, S, stage)


def test_reference_pipedream_is_a_stub_ours_is_not(ref):
    rpipe, _ = ref
    from shallowspeed_b200 import pipe
    from shallowspeed_b200.parallel.validate import validate

    with pytest.raises(NotImplementedError):
        rpipe.PipeDreamSchedule()
    validate(pipe.PipeDreamSchedule, 8, 4)


@pytest.mark.parametrize("pp", [1, 2, 4, 8])
def test_partitioner_matches_reference(ref, pp):
    _, rlayers = ref
    from shallowspeed_b200.layers import MLP

    sizes = [784, 128, 127, 126, 125, 124, 123, 10]
    for stage in range(pp):
        with contextlib.redirect_stdout(io.StringIO()):
            r = rlayers.MLP(sizes, stage, pp, 128)
        o = MLP(sizes, stage, pp, 128)
        assert (o.in_dim, o.out_dim) == (r.in_dim, r.out_dim)
        assert [type(l).__name__ for l in o.layers] == [type(l).__name__ for l in r.layers]
        for lo, lr in zip(o.layers, r.layers):
            if type(lo).__name__ == "Linear":
                assert (lo.activation is None) == (lr.activation is None)
                assert tuple(lo._params["W"].data.shape) == tuple(lr._params["W"].data.shape)
                assert tuple(lo._params["b"].data.shape) == tuple(lr._params["b"].data.shape)
