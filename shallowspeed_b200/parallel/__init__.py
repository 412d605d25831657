from .comm import Comm, ProcessGrid, SelfComm, ThreadComm, ThreadFabric, TorchComm, make_torch_comms
from .instructions import *  # noqa: F401,F403
from .schedules import (GPipeSchedule, InferenceSchedule, NaiveParallelSchedule, PipeDreamFlushSchedule,
                        PipeDreamSchedule, SCHEDULE_NAME_TO_CLS, Schedule)
from .validate import ScheduleError, comm_groups, max_in_flight, simulate, validate
from .worker import Worker, backprop_allreduce_gradient, backprop_block_for_comms
