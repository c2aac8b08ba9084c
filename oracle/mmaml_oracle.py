"""CPU restatement of the MMAML conv nets (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Functional, on flat ``name -> tensor`` mappings with the reference's own parameter names; runs in the dtype of the
tensors.  ``file:line`` citations are relative to the reference checkout.
Parity status: pinned against golden vectors of the live reference (tests/golden/make_golden_mmaml.py).
"""
import torch
import torch.nn.functional as F


def gated_conv(params, x, embeddings=None):
    """GatedConvModel.forward, networks/gated_conv_net.py:161-212 (use_max_pool=False, condition_type='affine'):
    per layer conv 3x3 s2 p1 (:180-182) -> F.batch_norm(training=True, no affine, :185-189) -> FiLM
    x * (1 + gamma) + beta with (gamma, beta) = split(embedding) (:154-159) -> ReLU (:192-193); mean over the map
    (:201-205); Linear + tanh (:207-210)."""
    for i in range(1, 5):
        x = F.conv2d(x, params[f"features.layer{i}_conv.weight"], params[f"features.layer{i}_conv.bias"], stride=2, padding=1)
        x = F.batch_norm(x, None, None, training=True)
        if embeddings is not None:
            e = embeddings[i - 1]
            gammas, betas = torch.split(e, x.size(1), dim=-1)
            x = x * (gammas.view(1, -1, 1, 1) + 1.0) + betas.view(1, -1, 1, 1)
        x = F.relu(x)
    x = x.view(x.size(0), x.size(1), -1).mean(dim=2)
    return torch.tanh(F.linear(x, params["classifier.fully_connected.weight"], params["classifier.fully_connected.bias"]))


def conv_embedding(params, x, num_conv=4, pooling="avg", n_heads=4):
    """ConvEmbeddingModel.forward, networks/conv_embedding_model.py:99-184 (convolutional, batch_norm,
    avgpool_after_conv, no RNN): conv 3x3 s2 p1 -> affine F.batch_norm(training=True) -> ReLU (:106-118); mean over the
    map (:119-123); relu(linear) transposed to [1, hidden, N] and pooled over the N samples (:147-154); one Linear per
    head (:176-179)."""
    for i in range(1, num_conv + 1):
        x = F.conv2d(x, params[f"conv.conv{i}.weight"], params[f"conv.conv{i}.bias"], stride=2, padding=1)
        x = F.batch_norm(x, None, None, weight=params[f"conv.bn{i}.weight"], bias=params[f"conv.bn{i}.bias"], training=True)
        x = F.relu(x)
    x = x.view(x.size(0), x.size(1), -1).mean(dim=2)
    inputs = F.relu(F.linear(x, params["linear.weight"], params["linear.bias"]).view(1, x.size(0), -1).transpose(1, 2))
    pooled = (F.avg_pool1d if pooling == "avg" else F.max_pool1d)(inputs, x.size(0)).view(1, -1)
    return [F.linear(pooled, params[f"_embeddings.{j}.weight"], params[f"_embeddings.{j}.bias"]) for j in range(n_heads)], pooled
