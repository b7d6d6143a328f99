"""
tad_dftd3_b200 -- B200-native DFT-D3 dispersion hot path.

Drop-in for `tad_dftd3` (dftd3/tad-dftd3 v0.5.0) on CUDA devices:

    import tad_dftd3_b200 as d3
    energy = d3.dftd3(numbers, positions, param)      # atom-resolved, Hartree

Same signature, padded-batch convention, `Reference` data holder and damping
parameter dicts as the reference; the body runs as hand-written sm_100a CUDA
kernels behind the C ABI in `include/d3b200.h`, with analytic first and second
derivatives wired into torch autograd / torch.func.
"""
from . import damping, data, defaults, disp, exception, model, ncoord, periodic, reference, typing
from .disp import dftd3, hessian
from .periodic import dftd3_periodic
from .host import HostPipeline, bind_host_thread_to_gpu, dftd3_host
from .reference import Reference

from .__version__ import __version__

__all__ = [
    "dftd3", "hessian", "dftd3_periodic", "periodic", "dftd3_host", "HostPipeline", "bind_host_thread_to_gpu", "damping", "data", "defaults", "disp", "exception",
    "model", "ncoord", "reference", "typing",
    "Reference", "__version__",
]
