"""ELBO gradient and optimiser step: the reference gets gradients from TensorFlow autodiff inside
GPflow's AdamOptimizer (conv_gp/experiment.py:84-108); here the minibatch-sized backward runs in libdcgp.so
(dcgp_layer_backward: split-fp16 tcgen05 GEMMs per layer + elementwise kernels) and the small minibatch-independent
chain rule (dS_r, dalpha, KL -> Z, kernel hyper-parameters, q_mu, q_sqrt) is a closed form in dense algebra on the device:
its R-batched M^3 products on the library's own tensor-core GEMM (dcgp_bgemm_nt), the single-matrix float64 products as
torch ops (cuBLAS; O(M^3), independent of the batch -- the one place on the step where library kernels are used).
"""
import math
import os

import numpy as np
import torch

from . import _lib
from .dist import allreduce_sum_, shard_range, world
from .kernels import JITTER
from .layers import Conv2dMean, ConvLayer


def _rbf(Z, variance, lengthscale):
    Zs = Z / lengthscale
    n = (Zs * Zs).sum(1)
    d = n[:, None] + n[None, :] - 2.0 * Zs @ Zs.T
    return variance * torch.exp(-0.5 * d)


def softplus_inv(x):
    """GPflow transforms.positive: x = log(1 + exp(u)) + 1e-6"""
    y = x - 1e-6
    return y + math.log(-math.expm1(-y))


class LayerBackward(object):
    """Per-layer buffers + the M-only chain rule."""

    NATIVE = True                     # M-only chain rule in libdcgp.so (dcgp_layer_chain_rule); False: the torch restatement
    BATCHED_DTYPE = torch.float32     # (torch restatement) dtype of the R-batched M^3 products (torch.float64 = all-double)
    TC_BATCHED = True                 # (torch restatement) batched products on the library's tcgen05 GEMM (dcgp_bgemm_nt)

    def __init__(self, layer):
        self.layer = layer
        d = layer._desc()
        self.M, self.R = layer.num_inducing, layer._R
        self.Mp = (self.M + 63) // 64 * 64
        self.Jp = (self.R + 1) * self.Mp + 64
        dev = layer.device
        self.L = layer._view.patch_length
        self.gQB = torch.zeros((self.Jp, self.Mp), dtype=torch.float64, device=dev)
        self.gZ = torch.zeros((self.M, self.L), dtype=torch.float64, device=dev)
        self.gscal = torch.zeros(4, dtype=torch.float64, device=dev)
        self.gw = torch.zeros(layer._view.patch_count, dtype=torch.float64, device=dev)
        self.ws = _lib.Workspace()
        self._offs = None

    def t_sized(self, X, n_rep, g_mean, g_var, need_gX, phases=3):
        """dcgp_layer_backward_phases -> gX (or None); fills gQB, gZ, gscal, gw.  phases=1 runs only what gX needs;
        t_sized_rest() then queues the parameter-only remainder with the same arguments."""
        layer = self.layer
        d = layer._desc()
        X = _lib.f32(X, layer.device)
        n_rows = X.shape[0]
        gX = torch.empty_like(X) if need_gX else None
        ws = self.ws.get("bwd", _lib.lib.dcgp_backward_workspace_bytes(d, n_rows, n_rep), layer.device)
        aws = layer._ws.get("apply", 0, layer.device)
        w = layer._patch_weights()
        self._keep = (X, g_mean, g_var, w, gX, ws, aws, d, n_rows, n_rep)
        self._call(phases)
        mf = getattr(layer, "mean_function", None)
        if gX is not None and isinstance(mf, Conv2dMean):      # the fixed mean function passes g_mean straight to its taps
            v = layer._view
            gX += mf.backward(g_mean, int(v.input_size[0]), int(v.input_size[1]))
        return gX

    def t_sized_rest(self):
        self._call(2)

    def _call(self, phases):
        layer = self.layer
        X, g_mean, g_var, w, gX, ws, aws, d, n_rows, n_rep = self._keep
        _lib.check(_lib.lib.dcgp_layer_backward_phases(
            d, _lib.ptr(layer._prep), _lib.ptr(aws), _lib.ptr(layer._keep[0]), _lib.ptr(w), _lib.ptr(X), n_rows, n_rep,
            _lib.ptr(g_mean), _lib.ptr(g_var), _lib.ptr(gX), _lib.ptr(self.gQB), _lib.ptr(self.gZ), _lib.ptr(self.gscal),
            _lib.ptr(self.gw) if layer._kind == _lib.LAYER_SVGP_CONV else None, _lib.ptr(ws), ws.numel(), phases,
            _lib.stream()))

    @staticmethod
    def _rbf_parts(Z, var, ls):
        Zs = Z / ls
        n = (Zs * Zs).sum(1)
        D = n[:, None] + n[None, :] - 2.0 * Zs @ Zs.T
        return var * torch.exp(-0.5 * D), D

    @staticmethod
    def _rbf_chain(G, Kn, D, Z, var, ls, need_Z=True):
        """Gradients of sum(G * K(Z)) for K = var * exp(-D/2): d/dvar, d/dls and (optionally) d/dZ."""
        H = G * Kn
        gvar = H.sum() / var
        gls = (H * D).sum() / ls
        if not need_Z:
            return gvar, gls, None
        Hs = H + H.T
        gZ = -(Hs.sum(1, keepdim=True) * Z - Hs @ Z) / (ls * ls)
        return gvar, gls, gZ

    # ---- native chain rule (dcgp_layer_chain_rule): the product path
    def _chain(self, parts, kl_weight, hyp):
        layer = self.layer
        dev = layer.device
        d = layer._desc()
        M, R, L = self.M, self.R, self.L
        if getattr(self, "_chain_out", None) is None:
            self._chain_out = (torch.zeros((M, L), dtype=torch.float64, device=dev), torch.zeros(2, dtype=torch.float64, device=dev),
                               torch.zeros((M, R), dtype=torch.float64, device=dev),
                               torch.zeros((R, M, M), dtype=torch.float64, device=dev))
        gZ, ghyp, gqmu, gqsqrt = self._chain_out
        ws = self.ws.get("chain", _lib.lib.dcgp_chain_rule_workspace_bytes(d), dev)
        prep_ws = layer._ws.get("prep", 0, dev)
        Z = _lib.f64(layer.feature.Z, dev)
        Zp = layer._Z_prior()
        self._chain_keep = (Z, Zp, hyp)
        _lib.check(_lib.lib.dcgp_layer_chain_rule(
            d, _lib.ptr(layer._prep), _lib.ptr(prep_ws), _lib.ptr(Z), _lib.ptr(Zp), _lib.ptr(_lib.f64(layer.q_mu, dev)),
            _lib.ptr(_lib.f64(layer.q_sqrt, dev)), _lib.ptr(hyp), _lib.ptr(self.gQB), _lib.ptr(self.gZ), _lib.ptr(self.gscal),
            float(kl_weight), int(parts), _lib.ptr(gZ), _lib.ptr(ghyp), _lib.ptr(gqmu), _lib.ptr(gqsqrt), _lib.ptr(ws), ws.numel(),
            _lib.stream()))

    def m_only_static(self, kl_weight=1.0, hyp=None):
        """The part of the chain rule that depends only on the parameters and on this step's dcgp_layer_prepare (the KL
        gradient): everything that can be computed BEFORE the layer's dS / dalpha exist.  TrainStep queues it on the layer's
        side stream during the forward pass.  `hyp`: device tensor [variance, lengthscale] (else the host's values)."""
        if not self.NATIVE:
            return self.m_only_static_torch(kl_weight, hyp)
        self._chain(1, kl_weight, hyp)
        return True

    def m_only(self, kl_weight=1.0, hyp=None, static=None):
        """Chain rule through the minibatch-independent operands (csrc/dcgp_chain.cu has the algebra): returns
        d ELBO / d{Z, variance, lengthscale, q_mu, q_sqrt (lower), patch_weights}.  `kl_weight` = 1/world_size so that summing
        over ranks counts the KL once.  `hyp` (optional): device tensor [variance, lengthscale] to use instead of the host
        floats -- keeps the chain free of host values so that it can be captured in a CUDA graph (TrainStep).  `static`: the
        token returned by m_only_static() for the same parameters (both parts run here when absent)."""
        if not self.NATIVE:
            return self.m_only_torch(kl_weight, hyp, static)
        self._chain(2 if static else 3, kl_weight, hyp)
        gZ, ghyp, gqmu, gqsqrt = self._chain_out
        out = {"Z": gZ, "variance": ghyp[0], "lengthscale": ghyp[1], "q_mu": gqmu, "q_sqrt": gqsqrt}
        if self.layer._kind == _lib.LAYER_SVGP_CONV:
            out["patch_weights"] = self.gw.clone()
        return out

    def _views(self):
        """Views of this step's float64 factors inside the workspace of the layer's last dcgp_layer_prepare -- Kuu^-1 [M,M],
        Lm [M,M] (lower), Lm^-1 [M,M] (lower), the prior's Lp^-1 [M,M] -- and of the float32 tensor-core products it left in
        the `prep` buffer: C_r = Lm^-1 L_r (L_r when whitened) and S_r = C_r C_r^T, each [R,M,M]."""
        import ctypes as C
        layer, M, R = self.layer, self.M, self.R
        if self._offs is None:
            ok, ol, op, olm, ld = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_int()
            _lib.check(_lib.lib.dcgp_prepare_workspace_layout2(layer._desc(), C.byref(ok), C.byref(ol), C.byref(op), C.byref(olm),
                                                               C.byref(ld)))
            oc, os_, ldb = C.c_size_t(), C.c_size_t(), C.c_int()
            _lib.check(_lib.lib.dcgp_prepare_layout2(layer._desc(), C.byref(oc), C.byref(os_), C.byref(ldb)))
            self._offs = (ok.value, ol.value, op.value, olm.value, ld.value, oc.value, os_.value, ldb.value)
        ok, ol, op, olm, ld, oc, os_, ldb = self._offs
        ws = layer._ws.get("prep", 0, layer.device)
        f64 = lambda off, n: ws[off:off + n * n * 8].view(torch.float64).view(n, n)
        f32 = lambda off: layer._prep[off:off + R * ldb * ldb * 4].view(torch.float32).view(R, ldb, ldb)[:, :M, :M]
        return dict(Kinv=f64(ok, M), Li=f64(ol, ld)[:M, :M], Lpinv=f64(op, ld)[:M, :M], Lm=f64(olm, M), C=f32(oc), S=f32(os_))

    # The R-batched M^3 products of the chain rule: on the library's own tcgen05 GEMM (default), else through cuBLAS in shapes
    # it handles well (the [M,R*M] x [R*M,M] form of sum_r A_r B_r^T gets 32x32 tiles and 4 TFLOP/s: bmm + reduction instead)
    def _use_tc(self):
        return (self.TC_BATCHED and self.BATCHED_DTYPE == torch.float32 and self.layer._algo() == _lib.ALGO_TC
                and self.M % 4 == 0)        # (the GEMM's vectorised stores need a row length that is a multiple of 4)

    def _gemm_nt(self, A, B):
        """C[r] = A[r] @ B[r]^T (float32; a 2-D operand is shared by every r) on the tensor cores via dcgp_bgemm_nt."""
        A, B = A.contiguous(), B.contiguous()
        batch = B.shape[0] if B.dim() == 3 else A.shape[0]
        m, k = A.shape[-2], A.shape[-1]
        n = B.shape[-2]
        Cm = torch.empty((batch, m, n), dtype=torch.float32, device=A.device)
        nbytes = _lib.lib.dcgp_bgemm_workspace_bytes(batch, m, n, k)
        ws = self.ws.get("bgemm", nbytes, A.device)
        _lib.check(_lib.lib.dcgp_bgemm_nt(_lib.ptr(A), _lib.ptr(B), _lib.ptr(Cm), batch, m, n, k,
                                          m * k if A.dim() == 3 else 0, n * k if B.dim() == 3 else 0,
                                          _lib.ptr(ws), ws.numel(), _lib.stream()))
        return Cm

    def _left(self, Am, B3):
        """Am @ B_r for every r."""
        if self._use_tc():
            return self._gemm_nt(Am, B3.transpose(1, 2))
        R_, M_, N_ = B3.shape          # one [M,M] x [M,R*M] product (cuBLAS gets 32x32 tiles for the broadcast-batched form)
        return (Am @ B3.permute(1, 0, 2).reshape(M_, R_ * N_)).reshape(M_, R_, N_).permute(1, 0, 2)

    def _bmm(self, A3, B3):
        """A_r @ B_r for every r."""
        if self._use_tc():
            return self._gemm_nt(A3, B3.transpose(1, 2))
        return torch.bmm(A3, B3)

    def _bsum(self, A3, B3):
        """sum_r A_r @ B_r^T"""
        if self._use_tc():
            return self._gemm_nt(A3, B3).sum(0)
        return torch.bmm(A3, B3.transpose(1, 2)).sum(0)

    @torch.no_grad()
    def m_only_static_torch(self, kl_weight=1.0, hyp=None):
        """(torch restatement of the native chain rule, kept as its cross-check: tests/test_gpu_backward_pieces.py)
        The part of the chain rule that depends only on the parameters and on this step's dcgp_layer_prepare -- Kuu and its
        distance matrix, the KL gradient, the operand casts -- i.e. everything that can be computed BEFORE the layer's
        dS / dalpha exist.  TrainStep runs it on the layer's side stream during the forward pass, which takes ~40 % of the
        chain off the serial tail of the step."""
        layer = self.layer
        M, R = self.M, self.R
        white = layer.white
        Z = layer.feature.Z.to(torch.float64)
        if hyp is None:
            var, ls = float(layer._base_kernel.variance), float(layer._base_kernel.lengthscales)
        else:
            var, ls = hyp[0], hyp[1]
        q_mu, Lq = layer.q_mu, torch.tril(layer.q_sqrt)
        Kn, D = self._rbf_parts(Z, var, ls)
        vw = self._views()
        Kinv, Lpinv, Li = vw["Kinv"], vw["Lpinv"], vw["Li"]
        Lm = torch.tril(vw["Lm"])
        conv = isinstance(layer, ConvLayer)
        bt = self.BATCHED_DTYPE
        # The R-batched M^3 products run in BATCHED_DTYPE (float32 by default: their inputs -- dS from the split-fp16 GEMMs,
        # C_r / S_r from the forward -- carry 22-24 bits anyway); everything single-matrix stays float64.
        Lqb = Lq.to(bt)
        inv_diag = torch.diag_embed(1.0 / torch.diagonal(Lq, dim1=1, dim2=2))
        gvar_p = gls_p = 0.0
        GU_kl = None
        if white:       # KL = 1/2 [ |q_mu|^2 - MR - sum log diag(L_r)^2 + sum |L_r|^2 ]
            g_qmu_kl = -kl_weight * q_mu
            gLq_kl = -kl_weight * (Lq - inv_diag)
        else:
            if conv:    # prior = Kuu at the initial Z (a constant) with the live hyper-parameters (layers.py:149-150)
                Zp = layer.Z_prior.to(torch.float64)
                Kpn, Dp = self._rbf_parts(Zp, var, ls)
                Kpinv = Lpinv.T @ Lpinv
            else:
                Kpinv = Kinv
            Cb = self._left(Kpinv.to(bt), Lqb)
            a = Kpinv @ q_mu
            dKL_dKp = 0.5 * (-(a @ a.T) - self._bsum(Cb, Cb).to(torch.float64) + R * Kpinv)
            g_qmu_kl = -kl_weight * a
            gLq_kl = -kl_weight * (Cb.to(torch.float64) - inv_diag)
            if conv:
                gvar_p, gls_p, _ = self._rbf_chain(-kl_weight * dKL_dKp, Kpn, Dp, Zp, var, ls, need_Z=False)
            else:
                GU_kl = -kl_weight * dKL_dKp
        alpha = q_mu if white else Li @ q_mu
        return dict(Z=Z, var=var, ls=ls, q_mu=q_mu, Kn=Kn, D=D, Li=Li, LiT=Li.T.contiguous(), LiTb=Li.T.to(bt).contiguous(), Lm=Lm,
                    alpha=alpha, Cb=vw["C"].to(bt), Sb=vw["S"].to(bt), g_qmu_kl=g_qmu_kl, gLq_kl=gLq_kl, GU_kl=GU_kl, gvar_p=gvar_p,
                    gls_p=gls_p)

    @torch.no_grad()
    def m_only_torch(self, kl_weight=1.0, hyp=None, static=None):
        """(torch restatement of the native chain rule, kept as its cross-check)
        Chain rule through the minibatch-independent operands, in the order of the forward (conditionals.py:29-58):
             Lm = chol(Kuu), Li = Lm^-1, a = Li k;  C_r = Li L_r (L_r when whitened), S_r = C_r C_r^T, alpha = Li q_mu (q_mu)
           mean_r = alpha_r^T a,  var_r = knn - |a|^2 + a^T S_r a.
        Inputs (dcgp_layer_backward): dS_r = sum_t s_r a a^T, dalpha = sum_t a g_mean^T, and the direct paths gZ, gscal, gw.
          H  = sum_t da_t a_t^T = alpha dalpha^T + sum_r 2 (S_r - I) dS_r                    (d/dLi through a = Li k is H Lm^T)
          W  = H + sum_r 2 dS_r S_r + dalpha alpha^T    (non-whitened: C_r and alpha move with Li as well),  d/dLi = W Lm^T
          dLm = tril(-Li^T W),  dKuu = 1/2 Li^T (P + P^T) Li  with P = Phi(Lm^T dLm)          (Cholesky backward; Phi: tril, diag/2)
          d/dL_r = tril(Li^T 2 dS_r C_r)  (tril(2 dS_r C_r) whitened),  d/dq_mu = Li^T dalpha  (dalpha whitened)
        Returns d ELBO / d{Z, variance, lengthscale, q_mu, q_sqrt (lower), patch_weights}.  `kl_weight` = 1/world_size so that
        summing over ranks counts the KL once.  `hyp` (optional): device tensor [variance, lengthscale] to use instead of the
        host floats -- keeps the whole chain free of host values so that it can be captured in a CUDA graph (TrainStep).
        `static` = the result of m_only_static() for the same parameters (computed here when absent)."""
        layer = self.layer
        M, R, Mp = self.M, self.R, self.Mp
        white = layer.white
        st = static if static is not None else self.m_only_static_torch(kl_weight, hyp)
        bt = self.BATCHED_DTYPE
        gS = self.gQB[Mp:(R + 1) * Mp].reshape(R, Mp, Mp)[:, :M, :M]
        galpha = self.gQB[(R + 1) * Mp:(R + 1) * Mp + R, :M].T              # [M, R]
        gSb = gS.to(bt)
        Tm = self._bmm(st["Sb"], gSb)                                       # S_r dS_r
        Tsum = Tm.sum(0).to(torch.float64)
        alpha, Li, LiT, Lm = st["alpha"], st["Li"], st["LiT"], st["Lm"]
        H = alpha @ galpha.T + 2.0 * (Tsum - gS.sum(0))
        GC = 2.0 * self._bmm(gSb, st["Cb"])                                 # d/dC_r
        if white:
            W = H
            gLq = GC.to(torch.float64) + st["gLq_kl"]
            g_qmu = galpha + st["g_qmu_kl"]
        else:
            W = H + 2.0 * Tsum.T + galpha @ alpha.T
            gLq = self._left(st["LiTb"], GC).to(torch.float64) + st["gLq_kl"]
            g_qmu = LiT @ galpha + st["g_qmu_kl"]
        P = torch.tril(Lm.T @ torch.tril(-(LiT @ W)))
        P = P - 0.5 * torch.diag_embed(torch.diagonal(P))
        GU = 0.5 * (LiT @ (P + P.T) @ Li)                                   # d/dKuu
        if st["GU_kl"] is not None:
            GU = GU + st["GU_kl"]
        gvar, gls, gZ = self._rbf_chain(GU, st["Kn"], st["D"], st["Z"], st["var"], st["ls"])
        out = {"Z": gZ + self.gZ, "variance": gvar + st["gvar_p"] + self.gscal[0],
               "lengthscale": gls + st["gls_p"] + self.gscal[1], "q_mu": g_qmu, "q_sqrt": torch.tril(gLq)}
        if layer._kind == _lib.LAYER_SVGP_CONV:
            out["patch_weights"] = self.gw.clone()
        return out


class ElboGradient(object):
    """elbo, grads = ElboGradient(model)(X, Y, zs): forward ELBO + its gradient w.r.t. every layer's parameters."""

    def __init__(self, model):
        self.model = model
        self.bwd = [LayerBackward(l) for l in model.layers]
        self._side = None

    def _forward(self, X, Y, zs, n_global, defer_sum=False):
        """Forward ELBO (layer outputs kept) and the gradient of the data term w.r.t. the last layer's mean / var."""
        model = self.model
        X = _lib.f32(X, model.device)
        N, S = X.shape[0], model.num_samples
        if zs is None:      # rank-count-invariant draws: indexed by the GLOBAL position of this rank's images
            rank, wsize = world()
            n_g = int(n_global or N)
            n0 = shard_range(n_g, rank, wsize)[0] if n_g != N else 0
            zs = model.draw_zs(N, n_g, n0)
        zs = [_lib.f32(z, model.device) for z in zs]
        elbo = model._build_likelihood(X, Y, zs=zs, n_global=n_global, keep=True, defer_sum=defer_sum)
        Fs, Fmeans, Fvars = model._fwd
        K = Fmeans[-1].shape[2]
        coef = float(model.num_data) / float(n_global or N) / S
        Yd = model.likelihood.likelihood._labels(Y, model.device)
        Fm, Fv = Fmeans[-1].reshape(S * N, K).contiguous(), Fvars[-1].reshape(S * N, K).contiguous()
        g_mean, g_var = torch.empty_like(Fm), torch.empty_like(Fv)
        lik = model.likelihood.likelihood
        _lib.check(_lib.lib.dcgp_multiclass_varexp_grad(_lib.ptr(Fm), _lib.ptr(Fv), _lib.ptr(Yd), S, N, K, lik.epsilon, coef,
                                                        _lib.ptr(g_mean), _lib.ptr(g_var), _lib.stream()))
        if self._side is None:
            # chain streams outrank the main stream (their kernels are small and latency-critical); the first layer's chain
            # -- the one the next forward pass waits for first -- outranks the others
            # (CUDA clamps priorities to the supported range; lower = more urgent)
            self._side = [torch.cuda.Stream(device=model.device, priority=-2 if i == 0 else -1) for i in range(len(model.layers))]
        return X, zs, elbo, g_mean, g_var

    def _sample_backward(self, i, gX, zs):
        """DS/utils.py:41 backward through layer i-1's reparameterised sample -> (g_mean, g_var) for layer i-1."""
        Fvars = self.model._fwd[2]
        n = gX.numel()
        g_mean, g_var = torch.empty_like(gX), torch.empty_like(gX)
        zprev = zs[i - 1].reshape(-1).contiguous()
        vprev = Fvars[i - 1].reshape(-1).contiguous()
        _lib.check(_lib.lib.dcgp_sample_backward(_lib.ptr(gX), _lib.ptr(zprev), _lib.ptr(vprev), n, JITTER,
                                                 _lib.ptr(g_mean), _lib.ptr(g_var), _lib.stream()))
        return g_mean, g_var

    def __call__(self, X, Y, zs=None, n_global=None):
        model = self.model
        X, zs, elbo, g_mean, g_var = self._forward(X, Y, zs, n_global)
        Fs, Fmeans, Fvars = model._fwd
        N, S = X.shape[0], model.num_samples
        rank, wsize = world()
        grads = [None] * len(model.layers)
        main = torch.cuda.current_stream(model.device)
        # (1) the minibatch-sized backward of every layer, top to bottom, queued back to back on the main stream
        done = [None] * len(model.layers)
        for i in range(len(model.layers) - 1, -1, -1):
            first = (i == 0)
            Xin = X if first else Fs[i - 1].reshape(S * N, -1)
            gX = self.bwd[i].t_sized(Xin, S if first else 1, g_mean, g_var, need_gX=not first)
            done[i] = torch.cuda.Event()
            done[i].record(main)
            if not first:
                g_mean, g_var = self._sample_backward(i, gX, zs)
        # (2) the M-only chain rule of each layer on its own side stream: it only needs that layer's dQ / dbeta and
        #     overlaps with the minibatch-sized work still running below it
        for i in range(len(model.layers) - 1, -1, -1):
            self._side[i].wait_event(done[i])
            with torch.cuda.stream(self._side[i]):
                grads[i] = self.bwd[i].m_only(kl_weight=1.0 / wsize)
                for t in grads[i].values():
                    if isinstance(t, torch.Tensor):
                        t.record_stream(main)      # consumed by Adam on the main stream
        for side in self._side:
            main.wait_stream(side)
        return elbo, grads


class Adam(object):
    """tf.train.AdamOptimizer on GPflow's unconstrained variables (experiment.py:97-99): Z, softplus^-1(variance),
    softplus^-1(lengthscale), q_mu, the lower triangle of q_sqrt, patch_weights -- all layers in ONE flat float64 vector
    (one fused kernel; one all-reduce of the flat gradient when image-sharded)."""

    NAMES = ("Z", "variance", "lengthscale", "q_mu", "q_sqrt", "patch_weights")

    def __init__(self, model, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, frozen=(), sgd=False):
        """frozen: parameter names excluded from the update (`param.set_trainable(False)`, experiment.py:91-94: the
        variational parameters when the natural-gradient optimiser owns them).  sgd: plain gradient ascent with step lr
        (gpflow.train.GradientDescentOptimizer, experiment.py:101-104) instead of the Adam rule."""
        self.model, self.lr, self.b1, self.b2, self.eps = model, lr, beta1, beta2, eps
        self.frozen, self.sgd = tuple(frozen), bool(sgd)
        self.step_no = 0
        dev = model.device
        self.slots = []
        off = 0
        for li, layer in enumerate(model.layers):
            shapes = {"Z": tuple(layer.feature.Z.shape), "variance": (1,), "lengthscale": (1,), "q_mu": tuple(layer.q_mu.shape),
                      "q_sqrt": tuple(layer.q_sqrt.shape)}
            if layer._kind == _lib.LAYER_SVGP_CONV:
                shapes["patch_weights"] = (layer._view.patch_count,)
            for name in self.NAMES:
                if name in shapes:
                    n = int(np.prod(shapes[name]))
                    self.slots.append((li, name, off, n, shapes[name]))
                    off += n
        self.n = off
        self.flat = torch.zeros(off, dtype=torch.float64, device=dev)
        self.grad = torch.zeros(off, dtype=torch.float64, device=dev)
        self.m = torch.zeros(off, dtype=torch.float64, device=dev)
        self.v = torch.zeros(off, dtype=torch.float64, device=dev)
        self._pull()

    def _view(self, buf, slot):
        _, _, off, n, shape = slot
        return buf[off:off + n].view(shape)

    def _pull(self):
        for slot in self.slots:
            li, name = slot[0], slot[1]
            layer = self.model.layers[li]
            dst = self._view(self.flat, slot)
            if name == "Z":
                dst.copy_(layer.feature.Z)
            elif name == "variance":
                dst.fill_(softplus_inv(float(layer._base_kernel.variance)))
            elif name == "lengthscale":
                dst.fill_(softplus_inv(float(layer._base_kernel.lengthscales)))
            elif name == "q_mu":
                dst.copy_(layer.q_mu)
            elif name == "q_sqrt":
                dst.copy_(layer.q_sqrt)
            else:
                dst.copy_(_lib.f64(layer.kern.patch_weights, self.flat.device))

    def _push(self):
        hyp = []
        for slot in self.slots:
            li, name = slot[0], slot[1]
            layer = self.model.layers[li]
            src = self._view(self.flat, slot)
            if name == "Z":
                layer.feature.Z = src
            elif name in ("variance", "lengthscale"):
                hyp.append((layer, name, src))
            elif name == "q_mu":
                layer.q_mu = src
            elif name == "q_sqrt":
                layer.q_sqrt = src
            else:
                layer.kern.patch_weights = src
            layer._fresh = False            # any prepare() queued before this point is for the old parameters
        # kernel hyper-parameters travel by value in dcgp_layer_desc: ONE host read-back per step for all layers
        vals = torch.nn.functional.softplus(torch.cat([s for _, _, s in hyp])) + 1e-6
        vals = vals.cpu().tolist()
        for (layer, name, _), v in zip(hyp, vals):
            if name == "variance":
                layer._base_kernel.variance = v
            else:
                layer._base_kernel.lengthscales = v

    def _store_grads(self, grads, layer_index=None):
        for slot in self.slots:
            li, name = slot[0], slot[1]
            if layer_index is not None and li != layer_index:
                continue
            g = grads[li][name]
            dst = self._view(self.grad, slot)
            if name in self.frozen:
                dst.zero_()
                continue
            if name in ("variance", "lengthscale"):
                u = self._view(self.flat, slot)
                dst.copy_((g * torch.sigmoid(u)).reshape(1))          # d softplus(u) / du
            elif g.data_ptr() != dst.data_ptr():                      # (TrainStep lets the chain rule write straight into self.grad)
                dst.copy_(g.reshape(dst.shape))

    # ---- per-layer form used by TrainStep (same update rule; the layers' slices of the flat vector are disjoint)
    def layer_range(self, li):
        offs = [(s[2], s[2] + s[3]) for s in self.slots if s[0] == li]
        return min(o[0] for o in offs), max(o[1] for o in offs)

    def hyp_slice(self, li):
        """The layer's unconstrained (variance, lengthscale) pair inside the flat vector."""
        hyp = [s for s in self.slots if s[0] == li and s[1] in ("variance", "lengthscale")]
        assert hyp[0][1] == "variance" and hyp[1][2] == hyp[0][2] + 1
        return self.flat[hyp[0][2]:hyp[0][2] + 2]

    def _allreduce_slice(self, lo, hi, extra=None):
        """Sum self.grad[lo:hi] over the ranks (current stream).  The exchange is a single float32 bucket (the
        flat gradient travels as fp32; the optimiser state stays float64); `extra` (a 1-element float64 tensor, e.g. the
        ELBO's data term) rides in the bucket's last slot and is reduced in place with it."""
        _, wsize = world()
        if wsize == 1:
            return
        key = (lo, hi)
        if not hasattr(self, "_buckets"):
            self._buckets = {}
        b = self._buckets.get(key)
        if b is None:
            b = self._buckets[key] = torch.zeros(hi - lo + 1, dtype=torch.float32, device=self.grad.device)
        b[:hi - lo].copy_(self.grad[lo:hi])
        if extra is not None:
            b[hi - lo:].copy_(extra)
        allreduce_sum_(b)
        self.grad[lo:hi].copy_(b[:hi - lo])
        if extra is not None:
            extra.copy_(b[hi - lo:])

    def step_layer(self, li, grads, step_no, extra=None):
        """Adam update of layer li's slice on the CURRENT stream; returns (pinned host tensor, event) of the layer's
        constrained (variance, lengthscale) -- they travel by value in dcgp_layer_desc, so the host needs them back.
        grads=None: the layer's slice of self.grad has already been filled (captured graph)."""
        if grads is not None:
            self._store_grads(grads, li)
        lo, hi = self.layer_range(li)
        self._allreduce_slice(lo, hi, extra)
        _lib.check(_lib.lib.dcgp_adam(_lib.ptr(self.flat[lo:hi]), _lib.ptr(self.grad[lo:hi]), _lib.ptr(self.m[lo:hi]),
                                      _lib.ptr(self.v[lo:hi]), hi - lo, self.lr, self.b1, self.b2, self.eps, step_no, 1,
                                      _lib.stream()))
        vals = torch.nn.functional.softplus(self.hyp_slice(li)) + 1e-6
        if not hasattr(self, "_host_hyp"):
            self._host_hyp = {}
        if li not in self._host_hyp:
            self._host_hyp[li] = torch.empty(2, dtype=torch.float64).pin_memory()
        host = self._host_hyp[li]
        host.copy_(vals, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.flat.device))
        self._dev_hyp = getattr(self, "_dev_hyp", {})
        self._dev_hyp[li] = vals                     # the same pair on the device (layers.prepare(hyp=...))
        return host, ev

    def bind(self):
        """Point the layers' parameters at their slices of the flat vector (idempotent)."""
        self._push()

    def step(self, grads):
        """grads: list (per layer) of dicts of d ELBO / d constrained parameter (this rank's share)."""
        self._store_grads(grads)
        self._allreduce_slice(0, self.n)                                # the single exchange of the image-sharded step
        self.step_no += 1
        if self.sgd:
            self.flat.add_(self.grad, alpha=self.lr)                    # ascent on the ELBO
        else:
            _lib.check(_lib.lib.dcgp_adam(_lib.ptr(self.flat), _lib.ptr(self.grad), _lib.ptr(self.m), _lib.ptr(self.v), self.n,
                                          self.lr, self.b1, self.b2, self.eps, self.step_no, 1, _lib.stream()))
        self._push()


class NatGrad(object):
    """gpflow.train.NatGradOptimizer(gamma) on every layer's (q_mu, q_sqrt) (experiment.py:88-99), default xi = natural
    parameters: with L = -ELBO, theta = (S^-1 mu, -1/2 S^-1), eta = (mu, S + mu mu^T),
        theta <- theta - gamma dL/deta,      dL/deta_2 = dL/dS =: G,   dL/deta_1 = dL/dmu - 2 G mu,
    where dL/dS follows from dL/dq_sqrt by the Cholesky backward rule (S = L L^T).  The step leaves q(u) Gaussian only while
    -2 theta_2 stays positive definite; otherwise the Cholesky of the new covariance fails -- the reference's
    tf.errors.InvalidArgumentError (experiment.py:45-49) -- and NotPositiveDefiniteError is raised with the parameters
    untouched, for the caller's gamma back-off.  M-only float64 algebra (R x M x M), identical on every rank."""

    def __init__(self, model):
        self.model = model

    @staticmethod
    def _dS_from_dL(L, gL):
        """Gradient w.r.t. S = L L^T from the (lower-triangular) gradient w.r.t. L, batched over r."""
        P = torch.tril(L.transpose(1, 2) @ torch.tril(gL))
        P = P - 0.5 * torch.diag_embed(torch.diagonal(P, dim1=1, dim2=2))
        eye = torch.eye(L.shape[-1], dtype=L.dtype, device=L.device).expand_as(L)
        Linv = torch.linalg.solve_triangular(L, eye, upper=False)
        return 0.5 * Linv.transpose(1, 2) @ (P + P.transpose(1, 2)) @ Linv

    @torch.no_grad()
    def step(self, grads, gamma):
        """grads: per-layer dicts of d ELBO / d(q_mu [M,R], q_sqrt [R,M,M]) (all ranks' sum).  All layers are updated, or none."""
        new = []
        for layer, g in zip(self.model.layers, grads):
            mu = layer.q_mu.T.unsqueeze(2)                                 # [R, M, 1]
            L = torch.tril(layer.q_sqrt)
            gq_mu, gq_sqrt = allreduce_sum_(g["q_mu"].clone()), allreduce_sum_(g["q_sqrt"].clone())   # image-sharded: sum over ranks
            g_mu = -gq_mu.T.unsqueeze(2)                                   # L = -ELBO
            G = self._dS_from_dL(L, -torch.tril(gq_sqrt))
            eye = torch.eye(L.shape[-1], dtype=L.dtype, device=L.device).expand_as(L)
            Linv = torch.linalg.solve_triangular(L, eye, upper=False)
            Sinv = Linv.transpose(1, 2) @ Linv
            th1 = Sinv @ mu - gamma * (g_mu - 2.0 * G @ mu)
            prec = Sinv + 2.0 * gamma * G                                  # -2 theta_2'
            prec = 0.5 * (prec + prec.transpose(1, 2))
            Lp, info = torch.linalg.cholesky_ex(prec)
            if int(info.abs().max().item()) != 0 or not bool(torch.isfinite(Lp).all()):
                raise _lib.NotPositiveDefiniteError("natural-gradient step with gamma = %g leaves the set of valid covariances "
                                                    "(cf. experiment.py:45-49)" % gamma)
            Lpinv = torch.linalg.solve_triangular(Lp, eye, upper=False)
            S = Lpinv.transpose(1, 2) @ Lpinv
            Ln, info2 = torch.linalg.cholesky_ex(0.5 * (S + S.transpose(1, 2)))
            if int(info2.abs().max().item()) != 0:
                raise _lib.NotPositiveDefiniteError("natural-gradient step: new covariance is not positive definite")
            new.append(((S @ th1).squeeze(2).T.contiguous(), Ln.contiguous()))
        for layer, (q_mu, q_sqrt) in zip(self.model.layers, new):
            layer.q_mu.copy_(q_mu)
            layer.q_sqrt.copy_(q_sqrt)
            layer._fresh = False


class TrainStep(object):
    """One optimisation step = ElboGradient + Adam with the SAME arithmetic, scheduled so that the minibatch-independent
    tail of the step stops being serial:

      * every layer's backward is split (dcgp_layer_backward_phases): the part the layer below needs (input gradient) runs
        first, all the way down the stack; the parameter-only GEMMs (dQ, dbeta, dZ) of the upper layers are queued after it;
      * as soon as a layer's gradients are complete, its M-only chain rule, its slice of the Adam update (and of the
        gradient all-reduce) and its dcgp_layer_prepare for the NEXT step run on that layer's side stream, concurrently
        with the minibatch-sized kernels still on the main stream;
      * the kernel hyper-parameters travel by value (dcgp_layer_desc), so the host reads each layer's pair back from pinned
        memory lazily -- right before that layer is applied in the next forward pass (layer._pending).

    Parameter values after k calls are those of k x (ElboGradient, Adam.step) (tests/test_gpu_grad.py)."""

    GRAPH_AFTER = 2     # eager calls before a layer's M-only chain is captured in a CUDA graph

    def __init__(self, model, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, use_graphs=True):
        self.model = model
        self.eg = ElboGradient(model)
        self.opt = Adam(model, lr=lr, beta1=beta1, beta2=beta2, eps=eps)
        self.opt.bind()
        # the chain rule's outputs for Z, q_mu, q_sqrt ARE the layers' slices of the optimiser's flat gradient (no copies of
        # the [R, M, M] q_sqrt gradient per step)
        for i in range(len(model.layers)):
            views = {slot[1]: self.opt._view(self.opt.grad, slot) for slot in self.opt.slots if slot[0] == i}
            self.eg.bwd[i]._chain_out = (views["Z"], torch.zeros(2, dtype=torch.float64, device=model.device), views["q_mu"],
                                         views["q_sqrt"])
        self.use_graphs = use_graphs
        self.early_static = os.environ.get("DCGP_EARLY_STATIC", "1") != "0"   # diagnostic switch
        self.early_prepare = os.environ.get("DCGP_EARLY_PREPARE", "1") != "0"  # diagnostic switch (see _chain)
        # SMs the deferred (parameter-only) GEMMs leave to the chains running underneath them (dcgp_set_reserved_sms)
        self.reserve_sms = int(os.environ.get("DCGP_RESERVE_SMS", "0"))
        self._graphs = {"static": {}, "dynamic": {}}
        self._calls = {"static": {}, "dynamic": {}}
        self._static = {}

    def _graphed(self, key, i, fn):
        """Run fn() on the current stream: eagerly for the first GRAPH_AFTER calls, then captured once and replayed as a
        single CUDA-graph launch (fixed shapes and pointers, no host values inside)."""
        layer = self.model.layers[i]
        if not self.use_graphs:
            return fn()
        g = self._graphs[key].get(i)
        if g is not None:
            return g.replay()
        self._calls[key][i] = self._calls[key].get(i, 0) + 1
        if self._calls[key][i] <= self.GRAPH_AFTER:
            return fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=torch.cuda.current_stream(self.model.device), capture_error_mode="thread_local"):
            fn()
        self._graphs[key][i] = g
        g.replay()

    def _m_only_static(self, i, wsize):
        """Parameter-only part of layer i's chain rule (LayerBackward.m_only_static) on the current stream."""
        eg, opt = self.eg, self.opt

        def run():
            hyp = torch.nn.functional.softplus(opt.hyp_slice(i)) + 1e-6       # == the values the host holds
            self._static[i] = eg.bwd[i].m_only_static(kl_weight=1.0 / wsize, hyp=hyp)

        self._graphed("static", i, run)

    def _m_only_to_grad(self, i, wsize):
        """Layer i's M-only chain rule, results stored in its slice of opt.grad (current stream).  ~40 small torch ops with
        fixed shapes and pointers and no host values: replayed as ONE CUDA graph launch (the chain is otherwise bound by
        host launch overhead)."""
        eg, opt = self.eg, self.opt

        def run():
            hyp = torch.nn.functional.softplus(opt.hyp_slice(i)) + 1e-6
            grads = [None] * len(self.model.layers)
            grads[i] = eg.bwd[i].m_only(kl_weight=1.0 / wsize, hyp=hyp, static=self._static.get(i))
            opt._store_grads(grads, i)

        self._graphed("dynamic", i, run)

    def _chain(self, i, wsize, with_elbo=False):
        """Queue layer i's M-only chain rule + Adam slice on its side stream and leave the rest (hyper-parameter read-back,
        prepare for the next step) as the layer's pending hook.  with_elbo: this layer's gradient bucket also carries the
        ELBO's data term (image-sharded runs), and the ELBO scalar is formed behind it on the same stream."""
        model, eg, opt = self.model, self.eg, self.opt
        layer = model.layers[i]
        main = torch.cuda.current_stream(model.device)
        side = eg._side[i]
        done = torch.cuda.Event()
        done.record(main)
        side.wait_event(done)
        with torch.cuda.stream(side):
            if not self.early_static:
                self._m_only_static(i, wsize)
            self._m_only_to_grad(i, wsize)
            host, ev = opt.step_layer(i, None, opt.step_no, extra=model._sum if with_elbo else None)
            if with_elbo:
                model._finish_elbo()
                self._elbo_ready = torch.cuda.Event()
                self._elbo_ready.record(side)
            if self.early_prepare:
                # the next step's prepare right behind the update, hyper-parameters from the device: it no longer waits for
                # the host to notice the update, read the pair back and launch ~30 kernels one by one
                layer.prepare(hyp=opt._dev_hyp[i])
                layer._ready = torch.cuda.Event()
                layer._ready.record(side)

        def finish():
            ev.synchronize()
            v, l = host.tolist()
            layer._base_kernel.variance, layer._base_kernel.lengthscales = v, l
            if not self.early_prepare:
                with torch.cuda.stream(side):
                    layer.prepare()
                    layer._ready = torch.cuda.Event()
                    layer._ready.record(side)
            layer._fresh = True          # consumed (and cleared) by the next forward pass

        layer._pending = finish

    def __call__(self, X, Y, zs=None, n_global=None):
        model, eg, opt = self.model, self.eg, self.opt
        _, wsize = world()
        sharded = wsize > 1 and n_global is not None and int(n_global) != int(X.shape[0])
        # image-sharded: the data term of the ELBO is summed over the ranks inside layer 0's gradient bucket (no separate
        # collective between the forward and the backward pass)
        X, zs, elbo, g_mean, g_var = eg._forward(X, Y, zs, n_global, defer_sum=sharded)
        Fs = model._fwd[0]
        N, S = X.shape[0], model.num_samples
        nl = len(model.layers)
        opt.step_no += 1
        main = torch.cuda.current_stream(model.device)
        # parameter-only part of every layer's chain rule: queued now on the layers' side streams (behind this step's
        # prepare), it overlaps the backward pass instead of sitting in the serial tail
        for i in range(nl if self.early_static else 0):
            eg._side[i].wait_stream(main)       # (the first step's prepare ran on the model's own side streams)
            with torch.cuda.stream(eg._side[i]):
                self._m_only_static(i, wsize)
        # input-gradient path, top to bottom (the first layer has no input gradient: its whole backward runs here)
        for i in range(nl - 1, -1, -1):
            first = (i == 0)
            Xin = X if first else Fs[i - 1].reshape(S * N, -1)
            gX = eg.bwd[i].t_sized(Xin, S if first else 1, g_mean, g_var, need_gX=not first, phases=3 if first else 1)
            if not first:
                g_mean, g_var = eg._sample_backward(i, gX, zs)
        self._chain(0, wsize, with_elbo=sharded)
        # parameter-only remainder of the upper layers, each followed by its own chain
        if self.reserve_sms:
            _lib.lib.dcgp_set_reserved_sms(self.reserve_sms)
        try:
            for i in range(1, nl):
                eg.bwd[i].t_sized_rest()
                self._chain(i, wsize)
        finally:
            if self.reserve_sms:
                _lib.lib.dcgp_set_reserved_sms(0)
        if sharded:
            main.wait_event(self._elbo_ready)      # the returned scalar is complete in main-stream order
        return elbo

    def finish(self):
        """Complete every pending per-layer update (host-visible parameters are then current)."""
        for layer in self.model.layers:
            layer._run_pending()
