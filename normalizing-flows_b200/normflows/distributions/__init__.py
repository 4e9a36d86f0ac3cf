from .base import BaseDistribution, DiagGaussian
from .target import TwoMoons
