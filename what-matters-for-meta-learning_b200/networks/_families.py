"""The two model families behind the six drop-in classes.

* ``ResNetFamilyNP``  -- CNPDistractor / CondNeuralProcess / ANPDistractor / ANP: 64-channel
  residual CNN encoder + decoder CNN (networks/CNPDistractor.py, CondNeuralProcess.py,
  ANPDistractor.py, ANP.py in the reference).
* ``ShapeNet1DFamilyNP`` -- CNPShapeNet1D / ANPShapeNet1D: 3-conv ``encoder_w0`` shared by
  context and target images, no decoder CNN (networks/CNPShapeNet1D.py, ANPShapeNet1D.py).

Constructors create sub-modules in exactly the reference's order (the seeded RNG stream and the
``state_dict`` key order depend on it); ``forward`` hands the tensors to the CUDA engine.
"""
import os

import torch
from torch import nn

from networks._holders import AttnLinear, EncoderFC, FastAttention, ImageEncoder, NPDecoder

N_HEADS = 8


def _attention_block(module, h_dim, dim_heads):
    """_W_k / _W_v / _W_q (8 heads each), output projection _W and the FAVOR+ buffer, in the
    order of networks/ANPDistractor.py:59-72."""
    module._W_k = nn.ModuleList([AttnLinear(h_dim, h_dim) for _ in range(N_HEADS)])
    module._W_v = nn.ModuleList([AttnLinear(h_dim, h_dim) for _ in range(N_HEADS)])
    module._W_q = nn.ModuleList([AttnLinear(h_dim, h_dim) for _ in range(N_HEADS)])
    module._W = AttnLinear(N_HEADS * h_dim, h_dim)
    module.attn = FastAttention(dim_heads=dim_heads)
    module.n_heads = N_HEADS


def _mlp3(in_dim, width=256):
    return nn.Sequential(nn.Linear(in_dim, width), nn.ReLU(), nn.Linear(width, width), nn.ReLU(),
                         nn.Linear(width, width), nn.ReLU())


class _NPBase(nn.Module):
    def _read_common(self, config):
        self.device = config.device
        self.img_size = config.img_size
        self.task_num = config.tasks_per_batch
        self.label_dim = config.input_dim
        self.agg_mode = config.agg_mode
        self.img_agg = config.img_agg
        self.y_dim = config.output_dim

    def enable_cuda_graphs(self, on=True):
        """Replay the forward (and, through autograd, the backward) as CUDA graphs cached per input shape
        (b200np/graphed.py); the call contract is unchanged.  Also switched on by B200NP_GRAPHS=1."""
        object.__setattr__(self, "_graphed", None)
        object.__setattr__(self, "_use_graphs", bool(on))
        return self

    def forward(self, batch_train_images, label_train, batch_test_images, test=False):
        """(ctx_x [T,nc,C,H,W], ctx_y [T,nc,L], tgt_x [T,nt,C,H,W]) -> (mu [T,nt,out], None, 0);
        same contract as networks/ANPDistractor.py:103-135."""
        from b200np import engine
        use = getattr(self, "_use_graphs", None)
        if use is None:
            use = os.environ.get("B200NP_GRAPHS", "0") == "1"
        if use and not torch.cuda.is_current_stream_capturing():
            g = getattr(self, "_graphed", None)
            if g is None:
                from b200np.graphed import GraphedModel
                g = GraphedModel(self)
                object.__setattr__(self, "_graphed", g)   # not a sub-module / buffer: invisible to state_dict
            mu = g(batch_train_images, label_train, batch_test_images)
        else:
            mu = engine.forward(self, batch_train_images, label_train, batch_test_images)
        return mu, None, 0


class ResNetFamilyNP(_NPBase):
    family = "resnet"

    def __init__(self, config, use_transform_y, attention):
        super().__init__()
        self._read_common(config)
        self.img_channels = self.img_size[2] - 1 if config.task == "shapenet_3d" else self.img_size[2]
        self.attention = attention
        if use_transform_y:
            self.dim_w = config.dim_w
        if attention and not use_transform_y:
            self.temperature = config.temperature  # read (and unused) by networks/ANP.py:39
        if not attention and self.agg_mode not in ("mean", "max", "baco"):
            pass  # reference raises lazily in forward (CNPDistractor.py:111-112); so do we
        torch.manual_seed(config.seed)

        self.img_encoder = ImageEncoder(aggregate=self.img_agg, task_num=self.task_num,
                                        img_channels=self.img_channels)
        lab_w = self.label_dim
        if use_transform_y:
            self.transform_y = nn.Linear(self.label_dim, self.dim_w)
            lab_w = self.dim_w
        self.task_encoder = _mlp3(256 + lab_w)
        if not attention and self.agg_mode == "baco":
            self.latent_mu = nn.Linear(256, 256)
            self.latent_var = nn.Linear(256, 256)
        self.mu = nn.Linear(256, 256)
        self.decoder = NPDecoder(aggregate=self.img_agg, output_dim=self.y_dim, task_num=self.task_num,
                                 img_channels=self.img_channels, img_size=self.img_size)
        if attention:
            _attention_block(self, 256, 256)


class ShapeNet1DFamilyNP(_NPBase):
    family = "shapenet1d"

    def __init__(self, config, attention):
        super().__init__()
        self._read_common(config)
        self.img_channels = self.img_size[2]
        self.attention = attention
        self.dim_w = config.dim_w
        self.n_hidden_units_r = config.n_hidden_units_r
        self.dim_r = config.dim_r
        self.dim_z = config.dim_z
        torch.manual_seed(config.seed)

        # networks/CNPShapeNet1D.py:46-56 -- indices 0,2,5,8 carry parameters
        self.encoder_w0 = nn.Sequential(
            nn.Conv2d(self.img_channels, 32, kernel_size=3, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(32, 48, kernel_size=3, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.MaxPool2d((2, 2)),
            nn.Conv2d(48, 64, kernel_size=3, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Flatten(), nn.Linear(4096, self.dim_w))
        self.transform_y = nn.Linear(self.label_dim, self.dim_w // 4)
        self.encoder_r = EncoderFC(input_dim=self.dim_w + self.dim_w // 4,
                                   n_hidden_units_r=self.n_hidden_units_r, dim_r=self.dim_r)
        self.r_to_z = nn.Linear(self.dim_r, self.dim_z)
        self.decoder0 = nn.Sequential(
            nn.Linear(self.dim_w + self.dim_z, 100), nn.ReLU(inplace=True),
            nn.Linear(100, 100), nn.ReLU(inplace=True), nn.Linear(100, self.y_dim), nn.Tanh())
        if attention:
            _attention_block(self, self.dim_w, self.dim_r)  # ANPShapeNet1D.py:74-89
        elif self.agg_mode == "baco":
            self.rs_to_mu = nn.Linear(256, 256)  # CNPShapeNet1D.py:74-76
            self.rs_to_var = nn.Linear(256, 256)
