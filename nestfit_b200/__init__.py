"""nestfit_b200 -- B200-native likelihood hot path of NestFit.

Same public names as ``nestfit`` for the path it replaces (reference
nestfit/__init__.py:8-62); everything numerical runs in hand-written sm_100a
CUDA kernels behind the C ABI of ``include/nestfit_b200.h``.
"""
from .core import (  # noqa: F401
    Distribution, Prior, ConstantPrior, DuplicatePrior, OrderedPrior, SpacedPrior, CenSepPrior,
    ResolvedCenSepPrior, ResolvedPlacementPrior, PriorTransformer, Spectrum, Runner,
)
from .pixels import PixelBlock  # noqa: F401
from .models.ammonia import amm_predict, AmmoniaSpectrum, AmmoniaRunner  # noqa: F401
from .models.gaussian import gauss_predict, GaussianRunner  # noqa: F401
from .models.diazenylium import nnhp_predict, DiazenyliumSpectrum, DiazenyliumRunner  # noqa: F401
from .prior_constructors import get_irdc_priors, get_synth_priors  # noqa: F401
from .sampler import NestedSamplingBatch, Dumper, run_multinest  # noqa: F401
from .store import HdfStore, MemGroup  # noqa: F401
from . import postprocess  # noqa: F401
from .main import (  # noqa: F401
    NoiseMap, NoiseMapUniform, DataCube, CubeStack, CubeFitter, get_multiproc_indices, get_block_indices,
)
