"""The alignment prior the trainer facade attaches to synthetic stage-1 batches (trainers.beta_binomial_prior_distribution,
a restatement of fastpitch/data_function.py:85-99 without scipy) against scipy.stats.betabinom -- the function the
reference calls -- and against the oracle's copy. CPU only."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("P,M,scale", [(7, 20, 1.0), (24, 64, 1.0), (160, 880, 1.0), (12, 30, 0.5)])
def test_beta_binomial_prior_matches_scipy(P, M, scale):
    from scipy.stats import betabinom

    from oracle import fastpitch as ofp
    from xva_trainer_b200 import trainers

    want = np.array([betabinom(P, scale * i, scale * (M + 1 - i)).pmf(np.arange(P)) for i in range(1, M + 1)])
    got = trainers.beta_binomial_prior_distribution(P, M, scale)
    assert tuple(got.shape) == (M, P) and got.dtype == torch.float32
    np.testing.assert_allclose(got.numpy(), want, rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(ofp.beta_binomial_prior(P, M, scale).numpy(), want, rtol=2e-5, atol=1e-9)
    # rows are the pmf over 0..P (P + 1 outcomes) evaluated at 0..P-1: they sum to 1 minus the mass of outcome P
    tail = np.array([betabinom(P, scale * i, scale * (M + 1 - i)).pmf(P) for i in range(1, M + 1)])
    np.testing.assert_allclose(got.double().sum(1).numpy(), 1.0 - tail, rtol=1e-4, atol=1e-6)


def test_target_delta_table_follows_the_reference():
    """get_target_delta, xva_train.py:588-672: spot values computed by hand from the reference's branches."""
    from xva_trainer_b200.trainers import get_target_delta as d

    assert d(1, 5000) == 2e-5 and d(1, 3000) == 15e-5 and d(1, 1000) == 4e-4 and d(1, 100) == 4e-4 and d(1, 500) == 0
    assert d(2, 5000) == 5e-5 * 1.5 and d(2, 3000) == 1e-4 * 1.5 and d(2, 1000) == 5e-4 * 1.5 and d(2, 499) == 4e-3 * 1.5
    assert d(3, 5000) == 5e-5 * 2.5 and d(3, 1000) == 6e-4 * 2.5 and d(3, 300) == 1e-3 * 2.5 and d(3, 100) == 2e-3 * 2.5
    assert d(4, 5000) == 35e-6 * 3 and d(4, 3000) == 1e-4 * 3 and d(4, 1000) == 25e-5 * 3 and d(4, 300) == 45e-5 * 3
    assert d(4, 100) == 15e-4 * 3
