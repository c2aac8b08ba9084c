"""Parameter holders for the neural-process building blocks.

The reference keeps its weights in ``ImageEncoder`` / ``NPDecoder`` / ``EncoderFC`` / ``AttnLinear``
(networks/models.py:27-203) and a BatchNorm-free ``ResNet(BasicBlock,[1,1,1,1])``
(networks/ResNet.py:36-74,126-215).  The classes below own parameters with the *same names,
shapes, creation order and initialisers* -- so a seeded construction consumes the torch RNG
stream identically and ``state_dict`` round-trips strictly with reference checkpoints -- but they
contain no arithmetic: ``forward`` is implemented by the CUDA engine (``b200np.engine``), which
reads ``.weight`` / ``.bias`` straight from these holders.
"""
import math

import torch
from torch import nn


class _ResidualStage(nn.Module):
    """One stride-2 residual block: conv1 3x3 s2 -> ReLU -> conv2 3x3 s1, plus a 1x1 s2 projection
    (``downsample.0``) on the skip path; no normalisation (networks/ResNet.py:36-74)."""

    def __init__(self, skip_proj, width=64):
        super().__init__()
        self.conv1 = nn.Conv2d(width, width, 3, stride=2, padding=1, bias=True)
        self.conv2 = nn.Conv2d(width, width, 3, stride=1, padding=1, bias=True)
        self.downsample = skip_proj


class _Trunk(nn.Module):
    """``resnet`` sub-module: layer1..layer4 (one stage each, all 64 channels) and the dead
    ``fc`` 512->1000 classifier that the reference never calls but keeps in its state_dict
    (networks/ResNet.py:144-153).  Conv weights are re-drawn kaiming-normal(fan_out) after
    construction, biases keep their default draw (networks/ResNet.py:155-157)."""

    def __init__(self, width=64):
        super().__init__()
        for l in (1, 2, 3, 4):
            # networks/ResNet.py:192-215: the skip projection is built before the block itself.
            proj = nn.Sequential(nn.Conv2d(width, width, 1, stride=2, bias=True))
            setattr(self, f"layer{l}", nn.Sequential(_ResidualStage(proj, width)))
        self.fc = nn.Linear(512, 1000)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def stages(self):
        return [getattr(self, f"layer{l}")[0] for l in (1, 2, 3, 4)]


class ImageEncoder(nn.Module):
    """Weights of the image encoder CNN: stem ``conv1`` 5x5 s2 p2 + ``resnet`` trunk
    (networks/models.py:75-90).  ``aggregate`` selects the pooling applied after layer4
    (models.py:105-113)."""

    def __init__(self, aggregate, task_num, img_channels):
        super().__init__()
        self.img_channels = img_channels
        self.task_num = task_num
        self.aggregate = aggregate
        self.conv1 = nn.Conv2d(img_channels, 64, kernel_size=5, stride=2, padding=2, bias=True)
        self.resnet = _Trunk()


class NPDecoder(nn.Module):
    """Weights of the decoder: a second CNN with its own parameters plus the ``fc_mu`` MLP
    512->256->256->out (networks/models.py:122-145).  ``pr_unc`` (variance head, models.py:146-154)
    is never enabled on the hot path (var=None, models.py:185-190) and is not built."""

    def __init__(self, aggregate, output_dim, task_num, img_channels, img_size, pr_unc=False):
        super().__init__()
        if pr_unc:
            raise NotImplementedError("variance head is outside the hot path (SURVEY.md section 0 #2)")
        self.img_channels = img_channels
        self.task_num = task_num
        self.img_size = img_size
        self.output_dim = output_dim
        self.aggregate = aggregate
        self.conv1 = nn.Conv2d(img_channels, 64, kernel_size=5, stride=2, padding=2, bias=True)
        self.resnet = _Trunk()
        self.fc_mu = nn.Sequential(
            nn.Linear(512, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(),
            nn.Linear(256, output_dim))


class EncoderFC(nn.Module):
    """``layers``: Linear/ReLU chain input_dim -> hidden... -> dim_r (networks/models.py:27-60)."""

    def __init__(self, input_dim, n_hidden_units_r, dim_r):
        super().__init__()
        self.input_dim, self.n_hidden_units_r, self.dim_r = input_dim, n_hidden_units_r, dim_r
        widths = [input_dim] + list(n_hidden_units_r)
        mods = []
        for a, b in zip(widths[:-1], widths[1:]):
            mods += [nn.Linear(a, b), nn.ReLU(inplace=True)]
        mods.append(nn.Linear(widths[-1], dim_r))
        self.layers = nn.Sequential(*mods)


class AttnLinear(nn.Module):
    """Linear whose weight is N(0, 1/in) (networks/models.py:195-203)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.linear = nn.Linear(in_channels, out_channels, bias=True)
        nn.init.normal_(self.linear.weight, std=in_channels ** -0.5)


class FastAttention(nn.Module):
    """Holder of the FAVOR+ ``projection_matrix`` buffer [int(d ln d), d]
    (networks/fast_attention.py:159-179).  The matrix is drawn exactly like the reference does it
    -- QR-orthogonalised d x d gaussian blocks on the CPU, rows rescaled by the norms of a fresh
    gaussian matrix (fast_attention.py:111-146, ``scaling=0``) -- because it consumes the seeded
    RNG stream and is saved in checkpoints.  It is never redrawn (no caller in the reference)."""

    def __init__(self, dim_heads, nb_features=None):
        super().__init__()
        self.dim_heads = dim_heads
        self.nb_features = int(dim_heads * math.log(dim_heads)) if nb_features is None else nb_features
        self.register_buffer("projection_matrix", self._draw(self.nb_features, dim_heads))

    @staticmethod
    def _draw(rows, cols):
        blocks = []
        full, rest = divmod(rows, cols)
        for take in [cols] * full + ([rest] if rest else []):
            q, _ = torch.linalg.qr(torch.randn(cols, cols), mode="reduced")
            blocks.append(q.t()[:take])
        basis = torch.cat(blocks)
        scale = torch.randn(rows, cols).norm(dim=1)
        return scale[:, None] * basis
