"""Training driver -- host-side mirror of conv_gp/experiment.py:13-136 for the default optimiser (`--optimizer Adam`):
learning-rate schedule (:71-73), the optimisation loop of `test_every` iterations (:38-44), the accuracy logger
(conv_gp/utils/log.py:50-68) and the parameter dump (:56-64); `--optimizer SGD` (:101-104) and the NatGrad + Adam hybrid
(:88-99) with its gamma schedule (:74-81) and Cholesky-failure back-off (:38-49).  SURVEY.md 8 rows f2/f4.
"""
import math
import os

import numpy as np

import torch

from . import _lib
from .grad import Adam, ElboGradient, NatGrad, TrainStep
from .models import ModelBuilder, save_model_parameters


def train_steps(flags):
    """arguments.py:4-7: roughly until the learning rate becomes 1e-5."""
    decay_count = math.log(5e-5 / flags.lr, 0.1)
    return math.ceil(flags.lr_decay_steps * decay_count / flags.test_every)


def exponential_decay(lr, global_step, decay_steps, decay_rate=0.1, staircase=True):
    """tf.train.exponential_decay as configured at experiment.py:72-73."""
    p = global_step / float(decay_steps)
    if staircase:
        p = math.floor(p)
    return lr * decay_rate ** p


def accuracy(model, X_test, Y_test, batch_size=32, num_samples=5):
    """utils/log.py:55-68 AccuracyLogger: mean class probability over `num_samples` samples, argmax, batches of 32."""
    X_test = np.asarray(X_test).reshape(len(X_test), -1).astype(np.float32)
    Y_test = np.asarray(Y_test).reshape(-1, 1)
    correct = 0
    for i in range(len(Y_test) // batch_size + 1):
        sl = slice(i * batch_size, (i + 1) * batch_size)
        X, Y = X_test[sl], Y_test[sl]
        if len(X) == 0:
            continue
        mean_samples, _ = model.predict_y(X, num_samples)
        probabilities = mean_samples.mean(dim=0)
        predicted_class = probabilities.argmax(dim=1)[:, None].cpu().numpy()
        correct += int((predicted_class == Y).sum())
    return correct / Y_test.size


class Experiment(object):
    """experiment.py:13-64 with in-memory data: `flags` carries the reference's options (name, log_dir, lr,
    lr_decay_steps, test_every, batch_size, M, ...)."""

    def __init__(self, flags, X_train, Y_train, X_test=None, Y_test=None, device="cuda", seed=0):
        self.optimizer = getattr(flags, "optimizer", "Adam")
        if self.optimizer not in ("Adam", "NatGrad", "SGD"):
            raise ValueError("Not a supported optimizer. Try Adam or NatGrad.")          # experiment.py:110-111
        self.flags = flags
        self.X_train, self.Y_train = np.asarray(X_train), np.asarray(Y_train)
        self.X_test, self.Y_test = X_test, Y_test
        path = self._model_path(flags.load_model) if getattr(flags, "load_model", None) else None
        builder = ModelBuilder(flags, self.X_train, self.Y_train, model_path=path, device=device, seed=seed)
        self.model = builder.build()                                          # _setup_model
        self.global_step = int(builder.global_step or 0)
        self.steps_back = 0                                                   # experiment.py:79 (gamma back-off counter)
        if self.optimizer == "Adam":                                          # _setup_optimizer
            self.step = TrainStep(self.model, lr=flags.lr)
        else:
            # NatGrad: the variational parameters belong to the natural-gradient action and are frozen for Adam (:88-99)
            self.eg = ElboGradient(self.model)
            self.opt = Adam(self.model, lr=flags.lr, frozen=("q_mu", "q_sqrt") if self.optimizer == "NatGrad" else (),
                            sgd=self.optimizer == "SGD")
            self.opt.bind()
            self.natgrad = NatGrad(self.model) if self.optimizer == "NatGrad" else None
        self.entries = []

    def _model_path(self, model_name=None):
        return os.path.join(self.flags.log_dir, (model_name or self.flags.name) + '.npy')

    def learning_rate(self):
        return exponential_decay(self.flags.lr, self.global_step, self.flags.lr_decay_steps)

    def gamma(self):
        """experiment.py:74-81: min((global_step / 100 * 1e-3 + flags.gamma) * 0.2 ** steps_back, 1)."""
        t = self.global_step / 100.0
        return min((t * 1e-3 + getattr(self.flags, "gamma", 1e-3)) * 0.2 ** self.steps_back, 1.0)

    def _batch(self):
        X, Y = self.model._next_batch()
        return np.asarray(X).reshape(len(X), -1).astype(np.float32), Y

    def _checked_gradient(self):
        """One ELBO + gradient evaluation; a failed Cholesky of Kuu or a non-finite ELBO is the reference's
        tf.errors.InvalidArgumentError at session.run."""
        X, Y = self._batch()
        elbo, grads = self.eg(X, Y)
        for layer in self.model.layers:
            _lib.raise_if_not_pd(layer._info)
        if not bool(torch.isfinite(elbo)):
            raise _lib.NotPositiveDefiniteError("non-finite ELBO")
        return elbo, grads

    def _iterations(self, numiter):
        """Loop(self.loop, stop=numiter)(): the loop's actions run one after the other, each on its own minibatch."""
        if self.optimizer == "Adam":
            for _ in range(numiter):
                self.step.opt.lr = self.learning_rate()
                X, Y = self._batch()
                self.last_elbo = self.step(X, Y)
                self.global_step += 1
            self.step.finish()
            for layer in self.model.layers:          # never write a checkpoint from a step that went wrong
                _lib.raise_if_not_pd(layer._info)
            if not bool(torch.isfinite(self.last_elbo)):
                raise FloatingPointError("non-finite ELBO at global step %d" % self.global_step)
            return
        for _ in range(numiter):
            if self.natgrad is not None:
                _, grads = self._checked_gradient()
                self.natgrad.step(grads, self.gamma())
            self.opt.lr = self.learning_rate()
            self.last_elbo, grads = self._checked_gradient()
            self.opt.step(grads)
            self.global_step += 1

    def _optimize(self, retry=0, error=None):
        """experiment.py:38-49: `test_every` iterations; with NatGrad a failed Cholesky shrinks gamma by 0.2 and the loop is
        started again, at most five times."""
        if retry > 5:
            raise error
        try:
            self._iterations(self.flags.test_every)
        except _lib.NotPositiveDefiniteError as exception:
            if self.optimizer != "NatGrad":
                raise
            self.steps_back += 1                                              # step_back_gamma
            self._optimize(retry=retry + 1, error=exception)

    def train_step(self):
        """experiment.py:28-31"""
        self._optimize()
        entry = {"global_step": self.global_step, "lr": self.learning_rate(), "elbo": float(self.last_elbo.item())}
        if self.X_test is not None:
            entry["test_accuracy"] = accuracy(self.model, self.X_test, self.Y_test)
        self.entries.append(entry)
        os.makedirs(self.flags.log_dir, exist_ok=True)
        save_model_parameters(self.model, self._model_path(), self.global_step)
        return entry
