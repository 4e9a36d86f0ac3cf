from .mlp import MLP
from .resnet import ResidualNet, ResidualBlock
from .made import MADE, MaskedLinear, MaskedResidualBlock
