"""gpflow.likelihoods.MultiClass (RobustMax, eps=1e-3) and DS/utils.py:54-121 BroadcastingLikelihood: the variational
expectation the ELBO needs (DS/dgp.py:83-90) and the prediction path (predict_mean_and_var / predict_density,
DS/dgp.py:116-126)."""
import torch

from . import _lib


class MultiClass(object):
    def __init__(self, num_classes, epsilon=1e-3):
        self.num_classes = int(num_classes)
        self.epsilon = float(epsilon)

    def _labels(self, Y, device):
        """Labels as int32 on the device.  Host-side labels are range-checked here (tf.one_hot would silently produce an
        all-zero row); labels already on the device are clamped by the kernels instead of being read back."""
        if not (isinstance(Y, torch.Tensor) and Y.is_cuda):
            Yh = torch.as_tensor(Y).reshape(-1)
            if Yh.numel() and (int(Yh.min()) < 0 or int(Yh.max()) >= self.num_classes):
                raise ValueError("MultiClass: labels must lie in [0, %d)" % self.num_classes)
        return torch.as_tensor(Y, device=device).reshape(-1).to(torch.int32).contiguous()

    def variational_expectations(self, Fmu, Fvar, Y, S=1, out_sum=None):
        """Fmu, Fvar [S*N, K] float32; Y [N] integer labels -> [S*N] float64 (and the device-side sum)."""
        Fmu, Fvar = _lib.f32(Fmu), _lib.f32(Fvar, Fmu.device)
        SN, K = Fmu.shape
        N = SN // S
        Y = self._labels(Y, Fmu.device)
        assert Y.numel() == N and K == self.num_classes
        ve = torch.empty((SN,), dtype=torch.float64, device=Fmu.device)
        total = out_sum if out_sum is not None else torch.empty(1, dtype=torch.float64, device=Fmu.device)
        _lib.check(_lib.lib.dcgp_multiclass_varexp(_lib.ptr(Fmu), _lib.ptr(Fvar), _lib.ptr(Y), S, N, K, self.epsilon,
                                                   _lib.ptr(ve), _lib.ptr(total), _lib.stream()))
        return ve, total

    def _predict(self, Fmu, Fvar, Y=None, S=1):
        Fmu, Fvar = _lib.f32(Fmu), _lib.f32(Fvar, Fmu.device)
        SN, K = Fmu.shape
        N = SN // S
        assert K == self.num_classes
        dev = Fmu.device
        pmean = torch.empty((SN, K), dtype=torch.float64, device=dev)
        pvar = torch.empty((SN, K), dtype=torch.float64, device=dev)
        logd = None
        if Y is not None:
            Y = self._labels(Y, dev)
            assert Y.numel() == N
            logd = torch.empty((SN,), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib.dcgp_multiclass_predict(_lib.ptr(Fmu), _lib.ptr(Fvar), _lib.ptr(Y), S, N, K, self.epsilon,
                                                    _lib.ptr(pmean), _lib.ptr(pvar), _lib.ptr(logd), _lib.stream()))
        return pmean, pvar, logd

    def predict_mean_and_var(self, Fmu, Fvar):
        """GPflow MultiClass.predict_mean_and_var: class probabilities ps [n, K] and ps - ps^2."""
        pmean, pvar, _ = self._predict(Fmu, Fvar)
        return pmean, pvar

    def predict_density(self, Fmu, Fvar, Y, S=1):
        """GPflow MultiClass.predict_density: log p(y | Fmu, Fvar), [n]; Y holds the N labels shared by the S samples."""
        return self._predict(Fmu, Fvar, Y, S)[2]


class BroadcastingLikelihood(object):
    """DS/utils.py:54-93: flatten the S dimension around the wrapped likelihood."""

    def __init__(self, likelihood):
        self.likelihood = likelihood

    def variational_expectations(self, Fmu, Fvar, Y):
        S, N, K = Fmu.shape
        ve, _ = self.likelihood.variational_expectations(Fmu.reshape(S * N, K), Fvar.reshape(S * N, K), Y, S=S)
        return ve.reshape(S, N, 1)

    def predict_mean_and_var(self, Fmu, Fvar):
        """DS/utils.py:107-112 -> ([S,N,K], [S,N,K])"""
        S, N, K = Fmu.shape
        m, v = self.likelihood.predict_mean_and_var(Fmu.reshape(S * N, K), Fvar.reshape(S * N, K))
        return m.reshape(S, N, K), v.reshape(S, N, K)

    def predict_density(self, Fmu, Fvar, Y):
        """DS/utils.py:114-121 -> [S,N,1]"""
        S, N, K = Fmu.shape
        d = self.likelihood.predict_density(Fmu.reshape(S * N, K), Fvar.reshape(S * N, K), Y, S=S)
        return d.reshape(S, N, 1)
