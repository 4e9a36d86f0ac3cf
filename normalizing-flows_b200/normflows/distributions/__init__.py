from .base import BaseDistribution, DiagGaussian, ClassCondDiagGaussian
from .target import TwoMoons
