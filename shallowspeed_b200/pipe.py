"""``shallowspeed_b200.pipe`` - same module name/surface as the reference's
``shallowspeed/pipe.py`` (instruction IR, schedules, DP hooks, Worker); implementation
lives in ``parallel``."""
from .parallel.instructions import *  # noqa: F401,F403
from .parallel.instructions import PipeInstr  # noqa: F401
from .parallel.schedules import (GPipeSchedule, InferenceSchedule, NaiveParallelSchedule,  # noqa: F401
                                 PipeDreamFlushSchedule, PipeDreamSchedule, SCHEDULE_NAME_TO_CLS, Schedule)
from .parallel.worker import Worker, backprop_allreduce_gradient, backprop_block_for_comms  # noqa: F401
