"""Helpers that the in-scope layers reference (normflows/utils/nn.py, utils/optim.py)."""
from .nn import ActNorm, ConstScaleLayer, ClampExp
from .optim import clear_grad, set_requires_grad, update_lipschitz
