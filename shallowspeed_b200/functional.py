"""``shallowspeed_b200.functional`` - same module name/surface as the reference's
``shallowspeed/functional.py``; implementation lives in ``ops.functional``."""
from .ops.functional import *  # noqa: F401,F403
from .ops.functional import (linear, linear_grad, loss_head_backward, mse_loss, mse_loss_grad, relu,
                             relu_grad, softmax, softmax_grad)
