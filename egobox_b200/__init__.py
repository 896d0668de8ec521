"""egobox_b200 -- B200-native (sm_100a) kriging hot path of egobox-gp.

Host-side mirror of the reference interface for this path; all arithmetic runs
in hand-written CUDA kernels behind the C ABI of ``include/egobox_gpu.h``."""
from ._lib import GpuError, device_count, load as load_library  # noqa: F401
from .context import (GpContext, SQUARED_EXPONENTIAL, ABSOLUTE_EXPONENTIAL, MATERN32, MATERN52,  # noqa: F401
                      CONSTANT, LINEAR, QUADRATIC, DEFAULT_NUGGET)
from .gp import (GaussianProcess, GpParams, Kriging, ThetaTuning, GpError, LinalgError,  # noqa: F401
                 LikelihoodComputationError, InvalidValueError,
                 SquaredExponentialCorr, AbsoluteExponentialCorr, Matern32Corr, Matern52Corr,
                 ConstantMean, LinearMean, QuadraticMean)
from .gpx import Gpx, GpMix, RegressionSpec, CorrelationSpec, Recombination  # noqa: F401
from .sgp import (SparseGaussianProcess, SgpParams, SparseKriging, SgpContext, ParamTuning, Inducings,  # noqa: F401
                  SparseMethod, SparseGpx, SparseGpMix)
from .moe import GaussianMixture, GpMixture, GpMixtureParams, fit_gmm, find_best_number_of_clusters  # noqa: F401
from . import metrics  # noqa: F401
from .mixture import ExpertMixture, recombine_smooth, hard_clusters  # noqa: F401,E402
