"""CPU: host-side model assembly helpers -- conv_gp/models.py:14-33,200-233, kernels.py:139-164."""
import numpy as np


def test_parse_and_image_hw():
    from deepcgp_b200 import models as Mo
    assert Mo.parse_ints('') == [] and Mo.parse_ints('384,384') == [384, 384]      # models.py:14-18
    assert Mo.image_HW(196) == [14, 14]
    f = Mo.default_parser().parse_args([])
    assert (f.M, f.feature_maps, f.filter_sizes, f.strides, f.batch_size, f.num_samples) == ('384,384', '10', '5,5', '2,1', 32, 10)


def test_identity_conv_is_the_centre_tap_sum():
    """mean_functions.py:6-26 with the filter of :21-25: VALID conv, stride s, 1.0 at the centre tap for every (in, out) pair."""
    from deepcgp_b200 import models as Mo
    rng = np.random.RandomState(0)
    X = rng.standard_normal((7, 9, 9, 3))
    out = Mo.identity_conv(X, 5, 3, 4, 2, rng=np.random.RandomState(1), samples=5)
    idx = np.random.RandomState(1).choice(np.arange(7), size=5)
    assert out.shape == (5, 3, 3, 4)
    filt = np.zeros((5, 5, 3, 4))
    filt[2, 2, :, :] = 1.0
    for k, n in enumerate(idx):
        for oy in range(3):
            for ox in range(3):
                ref = np.einsum("yxc,yxco->o", X[n, oy * 2:oy * 2 + 5, ox * 2:ox * 2 + 5, :], filt)
                np.testing.assert_allclose(out[k, oy, ox], ref, atol=1e-12)


def test_cluster_patches_returns_M_centroids_of_real_patches():
    from deepcgp_b200 import models as Mo
    rng = np.random.RandomState(3)
    X = rng.standard_normal((20, 8, 8, 2))
    Z = Mo.cluster_patches(X, 6, 3, rng=np.random.RandomState(5), samples_per_inducing_point=20)
    Z2 = Mo.cluster_patches(X, 6, 3, rng=np.random.RandomState(5), samples_per_inducing_point=20)
    assert Z.shape == (6, 3 * 3 * 2) and np.isfinite(Z).all()
    np.testing.assert_allclose(Z, Z2)                       # seeded
    assert np.abs(Z).max() <= np.abs(X).max()               # centroids are averages of patches


def test_checkpoint_keys_group_by_layer_and_shift_the_last_layer():
    """models.py:200-233: substring matching on GPflow pathnames; a 2-layer checkpoint loaded into a 3-layer model keeps
    layer 0 and moves its last layer to the model's last layer."""
    from deepcgp_b200 import models as Mo
    ck = {"DGP/layers/0/feature/Z": np.zeros((4, 9)), "DGP/layers/0/q_mu": np.ones((4, 2)),
          "DGP/layers/0/q_sqrt": np.zeros((2, 4, 4)), "DGP/layers/0/base_kernel/variance": np.float64(2.0),
          "DGP/layers/0/base_kernel/lengthscales": np.float64(3.0),
          "DGP/layers/1/feature/Z": np.zeros((5, 18)), "DGP/layers/1/q_mu": np.ones((5, 10)),
          "DGP/layers/1/kern/base_kernel/variance": np.float64(4.0), "DGP/layers/1/kern/patch_weights": np.ones(16),
          "DGP/likelihood/invlink/epsilon": 1e-3, "global_step": 7}
    step, lp = Mo.load_layer_parameters(ck, 2)
    assert step == 7 and sorted(lp) == [0, 1]
    assert lp[0]["base_kernel/variance"] == 2.0 and lp[0]["base_kernel/lengthscales"] == 3.0 and lp[0]["Z"].shape == (4, 9)
    assert lp[1]["base_kernel/variance"] == 4.0 and lp[1]["patch_weights"].shape == (16,)
    step, lp3 = Mo.load_layer_parameters(ck, 3)
    assert sorted(lp3) == [0, 2] and lp3[2]["Z"].shape == (5, 18)


def test_learning_rate_schedule_and_step_count():
    """experiment.py:72-73 (staircase decay x0.1 every lr_decay_steps) and arguments.py:4-7."""
    import argparse
    from deepcgp_b200 import experiment as E
    assert E.exponential_decay(0.01, 0, 100000) == 0.01
    assert E.exponential_decay(0.01, 99999, 100000) == 0.01
    assert abs(E.exponential_decay(0.01, 100000, 100000) - 0.001) < 1e-15
    assert abs(E.exponential_decay(0.01, 250000, 100000) - 0.0001) < 1e-15
    assert abs(E.exponential_decay(0.01, 50000, 100000, staircase=False) - 0.01 * 0.1 ** 0.5) < 1e-15
    f = argparse.Namespace(lr=0.01, lr_decay_steps=100000, test_every=50000)
    assert E.train_steps(f) == 5            # log_0.1(5e-5/0.01) = 2.30 rounds of decay -> ceil(2.30 * 2)
