from .mlp import MLP
from .resnet import ResidualNet, ResidualBlock
from .made import MADE, MaskedLinear, MaskedResidualBlock
from .cnn import ConvNet2d
from .lipschitz import LipschitzMLP, InducedNormLinear, Swish
