from .base import BaseDistribution, DiagGaussian, ClassCondDiagGaussian, ConditionalDiagGaussian
from .target import TwoMoons
