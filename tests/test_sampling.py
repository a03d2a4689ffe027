"""`sample_mixture` draws from the same distribution as the reference's
MixtureSameFamily(Categorical(probs), Independent(Normal(mean, std), 1)).sample() (PMoE/model/moe.py:127-133): component
frequencies, per-component mean and standard deviation, and the mixture moments torch.distributions reports."""
import torch
import torch.distributions as D


def test_sample_mixture_matches_the_mixture_distribution():
    from pmoe_b200.model.moe import sample_mixture
    torch.manual_seed(0)
    K, B = 4, 400000
    probs = torch.tensor([0.05, 0.5, 0.15, 0.3])
    mean = torch.tensor([[-2.0, 0.5], [0.0, -1.0], [1.5, 2.0], [3.0, 0.0]])
    std = torch.tensor([[0.1, 0.2], [0.3, 0.05], [0.5, 0.4], [0.2, 0.6]])
    P, M, S = probs.expand(B, K).contiguous(), mean.expand(B, K, 2).contiguous(), std.expand(B, K, 2).contiguous()
    x = sample_mixture(P, M, S)
    assert x.shape == (B, 2)
    dist = D.MixtureSameFamily(D.Categorical(probs), D.Independent(D.Normal(mean, std), 1))
    se = dist.variance.sqrt() / B ** 0.5
    assert ((x.mean(0) - dist.mean).abs() < 5 * se).all()
    assert ((x.var(0) - dist.variance).abs() / dist.variance < 2e-2).all()
    # component recovery: the components are well separated in at least one coordinate, so nearest-mean assignment is exact enough
    d = ((x[:, None, :] - mean[None]) / std[None]).pow(2).sum(-1)
    k = d.argmin(1)
    freq = torch.bincount(k, minlength=K).float() / B
    assert ((freq - probs).abs() < 5 * (probs * (1 - probs) / B).sqrt() + 2e-3).all()
    for j in range(K):
        xs = x[k == j]
        assert ((xs.mean(0) - mean[j]).abs() < 0.02).all() and ((xs.std(0) - std[j]).abs() / std[j] < 0.05).all()


def test_sample_mixture_edge_probabilities():
    from pmoe_b200.model.moe import sample_mixture
    torch.manual_seed(1)
    B = 1000
    probs = torch.tensor([0.0, 1.0, 0.0]).expand(B, 3).contiguous()     # routing collapses onto one expert
    mean = torch.tensor([[9.0, 9.0], [1.0, -1.0], [-9.0, -9.0]]).expand(B, 3, 2).contiguous()
    std = torch.full((B, 3, 2), 1e-6)
    x = sample_mixture(probs, mean, std)
    assert (x - torch.tensor([1.0, -1.0])).abs().max() < 1e-4
