"""odf — host side of the B200-native FALKON hot path (see include/odf.h for the C ABI)."""
from .falkon import Falkon, FalkonOptions, GaussianKernel, InCoreFalkon, fit_flops, sweep_flops  # noqa: F401
