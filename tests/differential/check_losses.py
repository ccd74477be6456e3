"""Loss modules vs the reference: values AND gradients with respect to the inputs, on the same random tensors."""
import torch
from torch import nn

import fl4health.losses.contrastive_loss as ref_con
import fl4health.losses.cosine_similarity_loss as ref_cos
import fl4health.losses.deep_mmd_loss as ref_dmmd
import fl4health.losses.perfcl_loss as ref_perfcl
import fl4health.losses.weight_drift_loss as ref_drift
import fl4health.preprocessing.autoencoders.loss as ref_vae
import fl4health_b200.losses.contrastive_loss as my_con
import fl4health_b200.losses.cosine_similarity_loss as my_cos
import fl4health_b200.losses.deep_mmd_loss as my_dmmd
import fl4health_b200.losses.perfcl_loss as my_perfcl
import fl4health_b200.losses.weight_drift_loss as my_drift
import fl4health_b200.preprocessing.autoencoders.loss as my_vae

cpu = torch.device("cpu")
torch.manual_seed(11)
agreed = 0


def value_and_grads(fn, *tensors):
    leaves = [t.clone().requires_grad_(True) for t in tensors]
    out = fn(*leaves)
    total = out if isinstance(out, torch.Tensor) else sum(out)
    total.backward()
    return out, [leaf.grad for leaf in leaves]


def agree(theirs, ours, *tensors, tol=1e-5):
    global agreed
    a, ga = value_and_grads(theirs, *tensors)
    b, gb = value_and_grads(ours, *tensors)
    for x, y in zip(a if isinstance(a, tuple) else (a,), b if isinstance(b, tuple) else (b,)):
        assert torch.allclose(x, y, atol=tol, rtol=tol), (x, y)
    for x, y in zip(ga, gb):
        assert (x is None) == (y is None)
        if x is not None:
            assert torch.allclose(x, y, atol=tol, rtol=tol), (x - y).abs().max()
    agreed += 1


for temperature in (0.5, 0.1):
    for n_neg in (1, 3):
        feats, pos, neg = torch.randn(8, 16), torch.randn(1, 8, 16), torch.randn(n_neg, 8, 16)
        agree(ref_con.MoonContrastiveLoss(cpu, temperature), my_con.MoonContrastiveLoss(cpu, temperature), feats, pos, neg)
    agree(ref_con.NtXentLoss(cpu, temperature), my_con.NtXentLoss(cpu, temperature), torch.randn(6, 12), torch.randn(6, 12))
    agree(ref_perfcl.PerFclLoss(cpu, temperature, 2 * temperature), my_perfcl.PerFclLoss(cpu, temperature, 2 * temperature),
          *[torch.randn(5, 10) for _ in range(5)])
agree(ref_cos.CosineSimilarityLoss(cpu), my_cos.CosineSimilarityLoss(cpu), torch.randn(7, 9), torch.randn(7, 9))
agree(ref_cos.CosineSimilarityLoss(cpu, dim=0), my_cos.CosineSimilarityLoss(cpu, dim=0), torch.randn(7, 9), torch.randn(7, 9))

# weight drift: penalty and gradient on the model parameters
net = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 3))
anchors = [p.detach() + 0.1 * torch.randn_like(p) for p in net.parameters()]
for weight in (0.0, 0.5, 2.0):
    grads = []
    for module in (ref_drift.WeightDriftLoss(cpu), my_drift.WeightDriftLoss(cpu)):
        net.zero_grad()
        value = module(net, anchors, weight)
        value.backward()
        grads.append((value.detach(), [p.grad.clone() for p in net.parameters()]))
    assert torch.allclose(grads[0][0], grads[1][0], atol=1e-6)
    assert all(torch.allclose(x, y, atol=1e-6) for x, y in zip(grads[0][1], grads[1][1]))
    agreed += 1

# VAE loss: reconstruction + KL on packed (logvar, mu, reconstruction) predictions
for latent in (4, 9):
    packed, target = torch.randn(6, 2 * latent + 20), torch.randn(6, 20)
    agree(lambda p, t: ref_vae.VaeLoss(latent, nn.MSELoss())(p, t), lambda p, t: my_vae.VaeLoss(latent, nn.MSELoss())(p, t), packed, target)

# deep MMD: same deep-kernel parameters -> same kernel statistic, same value after kernel training steps
for unbiased, degree in ((True, 1), (False, 2)):
    torch.manual_seed(5)
    theirs = ref_dmmd.DeepMmdLoss(cpu, input_size=10, hidden_size=8, output_size=6, is_unbiased=unbiased, gaussian_degree=degree, optimization_steps=2)
    ours = my_dmmd.DeepMmdLoss(cpu, input_size=10, hidden_size=8, output_size=6, is_unbiased=unbiased, gaussian_degree=degree, optimization_steps=2)
    ours.featurizer.load_state_dict(theirs.featurizer.state_dict())
    for name in ("epsilon_opt", "sigma_q_opt", "sigma_phi_opt"):
        getattr(ours, name).data.copy_(getattr(theirs, name).data)
    x, y = torch.randn(12, 10), torch.randn(12, 10) + 0.5
    theirs.training = ours.training = False
    assert torch.allclose(theirs(x, y).float(), ours(x, y).float(), atol=1e-5), (theirs(x, y), ours(x, y))
    agreed += 1
    theirs.training = ours.training = True  # forward first optimises the kernel for `optimization_steps` steps
    torch.manual_seed(77)  # each kernel-training step shuffles the target sample with the global generator
    a = theirs(x, y)
    torch.manual_seed(77)
    b = ours(x, y)
    assert torch.allclose(a.float(), b.float(), atol=1e-4, rtol=1e-3), (a, b)
    for (name, p), q in zip(theirs.featurizer.named_parameters(), ours.featurizer.parameters()):
        if name.endswith("6.bias"):
            continue  # the output bias cancels in every pairwise distance: its gradient is rounding noise, Adam's step its sign
        assert torch.allclose(p, q, atol=2e-4), (name, (p - q).abs().max())
    agreed += 1
print("configs agree:", agreed)
