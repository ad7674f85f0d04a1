"""gpflow.likelihoods.MultiClass (RobustMax, eps=1e-3) and DS/utils.py:54-121 BroadcastingLikelihood --
only the variational expectation the ELBO needs (DS/dgp.py:83-90)."""
import torch

from . import _lib


class MultiClass(object):
    def __init__(self, num_classes, epsilon=1e-3):
        self.num_classes = int(num_classes)
        self.epsilon = float(epsilon)

    def variational_expectations(self, Fmu, Fvar, Y, S=1, out_sum=None):
        """Fmu, Fvar [S*N, K] float32; Y [N] integer labels -> [S*N] float64 (and the device-side sum)."""
        Fmu, Fvar = _lib.f32(Fmu), _lib.f32(Fvar, Fmu.device)
        SN, K = Fmu.shape
        N = SN // S
        Y = torch.as_tensor(Y, device=Fmu.device).reshape(-1).to(torch.int32).contiguous()
        assert Y.numel() == N and K == self.num_classes
        ve = torch.empty((SN,), dtype=torch.float64, device=Fmu.device)
        total = out_sum if out_sum is not None else torch.empty(1, dtype=torch.float64, device=Fmu.device)
        _lib.check(_lib.lib.dcgp_multiclass_varexp(_lib.ptr(Fmu), _lib.ptr(Fvar), _lib.ptr(Y), S, N, K, self.epsilon,
                                                   _lib.ptr(ve), _lib.ptr(total), _lib.stream()))
        return ve, total


class BroadcastingLikelihood(object):
    """DS/utils.py:54-93: flatten the S dimension around the wrapped likelihood."""

    def __init__(self, likelihood):
        self.likelihood = likelihood

    def variational_expectations(self, Fmu, Fvar, Y):
        S, N, K = Fmu.shape
        ve, _ = self.likelihood.variational_expectations(Fmu.reshape(S * N, K), Fvar.reshape(S * N, K), Y, S=S)
        return ve.reshape(S, N, 1)
