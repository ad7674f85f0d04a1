"""deepcgp_b200 -- B200-native (sm_100a) drop-in for the conv-GP doubly-stochastic forward pass of
kekeblom/DeepCGP (conv_gp/views.py, layers.py, conditionals.py, kernels.py + the DS-DGP loop around them).

Python here is only the reference-facing surface; compute lives in lib/libdcgp.so (hand-written CUDA).
"""
import os

from . import _lib
from ._lib import ALGO_SIMT, ALGO_TC, NotPositiveDefiniteError  # noqa: F401


def default_algo():
    """Product path = tensor cores (tcgen05).  DCGP_ALGO=simt selects the fp32 CUDA-core validation path."""
    return ALGO_SIMT if os.environ.get("DCGP_ALGO", "tc").lower() == "simt" else ALGO_TC


from .views import FullView, View  # noqa: E402,F401
from .kernels import RBF, ConvKernel, AdditivePatchKernel, PatchInducingFeatures, Kuu, Kuf  # noqa: E402,F401
from .conditionals import conditional  # noqa: E402,F401
from .layers import Conv2dMean, ConvLayer, Layer, MultiOutputConvKernel, SVGP_Layer, Zero  # noqa: E402,F401
from .likelihoods import BroadcastingLikelihood, MultiClass  # noqa: E402,F401
from .dgp import DGP_Base  # noqa: E402,F401
from .grad import Adam, ElboGradient, NatGrad, TrainStep  # noqa: E402,F401
from .models import ModelBuilder, save_model_parameters  # noqa: E402,F401
from .experiment import Experiment, accuracy, exponential_decay, train_steps  # noqa: E402,F401
