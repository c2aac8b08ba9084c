"""CPU restatement of the reference's task-batched neural-process hot path.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Functional (no nn.Module): every function
takes a flat ``state_dict``-style mapping ``name -> tensor`` using the reference's own key
names, so weights move between the reference, this oracle and the CUDA
product by ``state_dict()`` alone.  Runs in whatever dtype the tensors carry (fp32 for parity,
fp64 for the oracle's own noise floor).  Gradients come from torch autograd on CPU.

All ``file:line`` citations are relative to the reference checkout (/root/reference).
Parity status: pinned against the live reference and golden vectors (see oracle/__init__.py).
"""
import math

import torch
import torch.nn.functional as F

N_HEADS = 8  # networks/ANPDistractor.py:61, networks/ANP.py:58, networks/ANPShapeNet1D.py:75
FAVOR_EPS = 1e-4  # networks/fast_attention.py:74


# --------------------------------------------------------------------------------------------
# CNN trunk
# --------------------------------------------------------------------------------------------
def basic_block(sd, p, x):
    """relu(conv2(relu(conv1_s2(x))) + downsample_1x1_s2(x)); no BatchNorm.
    networks/ResNet.py:58-74 (forward), :27-35 (bias=True convs), :200-204 (downsample)."""
    h = F.relu(F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], stride=2, padding=1))
    out = F.conv2d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], stride=1, padding=1)
    idn = F.conv2d(x, sd[p + "downsample.0.weight"], sd[p + "downsample.0.bias"], stride=2)
    return F.relu(out + idn)


def cnn_features(sd, p, img, img_agg, want_idx=False):
    """Stem 5x5 s2 p2 + ReLU, four stride-2 BasicBlocks, pooling, NCHW-order flatten.
    networks/models.py:92-113 (ImageEncoder.forward) == :160-180 (NPDecoder.forward).
    Returns [N, F] (F = 256 for every usable ``img_agg``) and optionally the pool argmax."""
    x = F.relu(F.conv2d(img, sd[p + "conv1.weight"], sd[p + "conv1.bias"], stride=2, padding=2))
    for l in (1, 2, 3, 4):
        x = basic_block(sd, f"{p}resnet.layer{l}.0.", x)
    idx = None
    if img_agg in ("max", "baco"):
        # networks/ResNet.py:152 AdaptiveMaxPool2d((2,2)); models.py:107-110
        x, idx = F.adaptive_max_pool2d(x, (2, 2), return_indices=True)
    elif img_agg == "mean":
        x = F.adaptive_avg_pool2d(x, (1, 1))  # networks/ResNet.py:151; models.py:105-106
    elif img_agg == "reshape":
        pass  # models.py:111-112
    x = x.reshape(x.size(0), -1)  # models.py:113 -- NCHW flatten: index c*4 + oh*2 + ow
    return (x, idx) if want_idx else x


def encoder_w0(sd, img, p="encoder_w0."):
    """networks/CNPShapeNet1D.py:46-56 (== ANPShapeNet1D.py:46-56): three 3x3 s2 convs with ReLU,
    a 2x2 max-pool after the second, NCHW flatten (4096) and Linear 4096 -> dim_w."""
    x = F.relu(F.conv2d(img, sd[p + "0.weight"], sd[p + "0.bias"], stride=2, padding=1))
    x = F.relu(F.conv2d(x, sd[p + "2.weight"], sd[p + "2.bias"], stride=2, padding=1))
    x = F.max_pool2d(x, 2)
    x = F.relu(F.conv2d(x, sd[p + "5.weight"], sd[p + "5.bias"], stride=2, padding=1))
    x = x.flatten(1)
    return F.linear(x, sd[p + "8.weight"], sd[p + "8.bias"])


# --------------------------------------------------------------------------------------------
# MLP pieces
# --------------------------------------------------------------------------------------------
def linear(sd, p, x):
    return F.linear(x, sd[p + "weight"], sd[p + "bias"])


def task_encoder(sd, x, p="task_encoder."):
    """3 x (Linear -> ReLU).  networks/CNPDistractor.py:45-52, ANPDistractor.py:48-55."""
    for i in (0, 2, 4):
        x = F.relu(linear(sd, f"{p}{i}.", x))
    return x


def fc_mu(sd, x, p="decoder.fc_mu."):
    """Linear-ReLU-Linear-ReLU-Linear.  networks/models.py:139-145."""
    x = F.relu(linear(sd, p + "0.", x))
    x = F.relu(linear(sd, p + "2.", x))
    return linear(sd, p + "4.", x)


def encoder_fc(sd, x, p="encoder_r.layers."):
    """networks/models.py:27-60 with n_hidden_units_r=[100,100]: Linear-ReLU-Linear-ReLU-Linear."""
    x = F.relu(linear(sd, p + "0.", x))
    x = F.relu(linear(sd, p + "2.", x))
    return linear(sd, p + "4.", x)


def decoder0(sd, x, p="decoder0."):
    """networks/CNPShapeNet1D.py:65-72: Linear-ReLU-Linear-ReLU-Linear-Tanh."""
    x = F.relu(linear(sd, p + "0.", x))
    x = F.relu(linear(sd, p + "2.", x))
    return torch.tanh(linear(sd, p + "4.", x))


def np_decoder(sd, tgt_imgs, sample_features, task_num, img_agg, p="decoder."):
    """NPDecoder.forward, networks/models.py:156-192 (second CNN with own weights, cat, fc_mu)."""
    n_per = sample_features.size(1)
    x = cnn_features(sd, p, tgt_imgs, img_agg).reshape(task_num, n_per, -1)
    return fc_mu(sd, torch.cat([x, sample_features], dim=-1), p + "fc_mu.")


# --------------------------------------------------------------------------------------------
# CNP aggregation
# --------------------------------------------------------------------------------------------
def aggregate(feats, mode, want_idx=False):
    """mean / max over the context dim.  networks/CNPDistractor.py:96-103,
    CNPShapeNet1D.py:115-120, CondNeuralProcess.py:94-101."""
    if mode == "mean":
        return (feats.mean(dim=1), None) if want_idx else feats.mean(dim=1)
    if mode == "max":
        v, i = feats.max(dim=1)
        return (v, i) if want_idx else v
    raise TypeError("agg_mode is not applicable for CNP, choose from ['mean', 'max', 'baco']")


def baco(mu, var):
    """Bayesian context aggregation, networks/CNPDistractor.py:60-75 (prior mean 0, var 1)."""
    sigma_inv = 1.0 / var
    sigma_z = 1.0 / (1.0 + sigma_inv.sum(dim=1))
    mu_z = sigma_z * (sigma_inv * mu).sum(dim=1)
    return mu_z, sigma_z


# --------------------------------------------------------------------------------------------
# FAVOR+ attention
# --------------------------------------------------------------------------------------------
def softmax_kernel(data, proj, is_query, eps=FAVOR_EPS):
    """Positive random features.  networks/fast_attention.py:74-99.
    data [B,H,n,d], proj [M,d] -> [B,H,n,M].  Maxima are taken on ``dd`` before ``diag`` is
    subtracted; the key max is over the WHOLE tensor (:97)."""
    d = data.shape[-1]
    dn = d ** -0.25  # :77
    ratio = proj.shape[0] ** -0.5  # :79
    dd = torch.einsum("bhid,jd->bhij", dn * data, proj.to(data.dtype))  # :81-84
    diag = (data ** 2).sum(dim=-1, keepdim=True) / 2.0 * (dn ** 2)  # :86-89
    if is_query:
        m = dd.max(dim=-1, keepdim=True).values  # :93
    else:
        m = dd.max()  # :97
    return ratio * (torch.exp(dd - diag - m) + eps)


def linear_attention(q, k, v):
    """networks/fast_attention.py:151-156 in the reference's own association order."""
    k_cumsum = k.sum(dim=-2)
    d_inv = 1.0 / torch.einsum("...nd,...d->...n", q, k_cumsum)
    context = torch.einsum("...nd,...ne->...de", k, v)
    return torch.einsum("...de,...nd,...n->...ne", context, q, d_inv)


def linear_attention_reassoc(q, k, v):
    """Same result as ``linear_attention`` with A = q' k'^T first;
    this is the association the CUDA kernel uses, kept here to bound the re-association error."""
    a = torch.einsum("...id,...jd->...ij", q, k)
    return torch.einsum("...ij,...je->...ie", a, v) / a.sum(dim=-1, keepdim=True)


def multihead_attention(sd, k, v, q, want_inter=False):
    """networks/ANPDistractor.py:78-101 (== ANP.py:75-98, ANPShapeNet1D.py:93-116)."""
    ks = torch.stack([linear(sd, f"_W_k.{h}.linear.", k) for h in range(N_HEADS)], dim=1)
    vs = torch.stack([linear(sd, f"_W_v.{h}.linear.", v) for h in range(N_HEADS)], dim=1)
    qs = torch.stack([linear(sd, f"_W_q.{h}.linear.", q) for h in range(N_HEADS)], dim=1)
    proj = sd["attn.projection_matrix"]
    qp = softmax_kernel(qs, proj, True)  # fast_attention.py:199
    kp = softmax_kernel(ks, proj, False)  # fast_attention.py:200
    out = linear_attention(qp, kp, vs)  # fast_attention.py:203
    out = out.permute(0, 2, 3, 1).contiguous()  # ANPDistractor.py:98 -> index e*8 + h
    out = out.view(out.shape[0], out.shape[1], -1)
    rep = linear(sd, "_W.linear.", out)
    if want_inter:
        return rep, {"q_heads": qs, "k_heads": ks, "v_heads": vs, "q_prime": qp, "k_prime": kp}
    return rep


# --------------------------------------------------------------------------------------------
# Models: forward(ctx_x [T,nc,C,H,W], ctx_y [T,nc,L], tgt_x [T,nt,C,H,W]) -> mu [T,nt,out]
# --------------------------------------------------------------------------------------------
def _flat_imgs(x):
    return x.reshape(-1, *x.shape[2:])


def cnp_distractor_forward(sd, cfg, ctx_x, ctx_y, tgt_x, inter=None):
    """networks/CNPDistractor.py:77-124 and networks/CondNeuralProcess.py:77-123 (the latter has
    no transform_y and feeds raw labels)."""
    T, nc, nt = cfg["tasks_per_batch"], ctx_x.shape[1], tgt_x.shape[1]
    if nc:
        lab = linear(sd, "transform_y.", ctx_y) if "transform_y.weight" in sd else ctx_y
        feats, pidx = cnn_features(sd, "img_encoder.", _flat_imgs(ctx_x), cfg["img_agg"], True)
        x_ctx = feats.view(T, -1, feats.size(1))  # models.py:115
        cf = task_encoder(sd, torch.cat([x_ctx, lab], dim=2))
        if cfg["agg_mode"] == "baco":
            m = linear(sd, "latent_mu.", cf)
            var = 1e-5 + F.softplus(linear(sd, "latent_var.", cf))
            r, _ = baco(m, var)
            aidx = None
        else:
            r, aidx = aggregate(cf, cfg["agg_mode"], True)
        mu = linear(sd, "mu.", r)
        sample = mu[:, None, :].repeat(1, nt, 1)
        if inter is not None:
            inter.update(x_ctx=x_ctx, pool_idx=pidx, ctx_feat=cf, agg=r, agg_idx=aidx)
    else:
        sample = torch.zeros(T, nt, 256, dtype=tgt_x.dtype)  # CNPDistractor.py:114
    return np_decoder(sd, _flat_imgs(tgt_x), sample, T, cfg["img_agg"])


def anp_distractor_forward(sd, cfg, ctx_x, ctx_y, tgt_x, inter=None):
    """networks/ANPDistractor.py:103-135 and networks/ANP.py:100-130 (no transform_y)."""
    T, nc, nt = cfg["tasks_per_batch"], ctx_x.shape[1], tgt_x.shape[1]
    if nc:
        lab = linear(sd, "transform_y.", ctx_y) if "transform_y.weight" in sd else ctx_y
        fc, pidx = cnn_features(sd, "img_encoder.", _flat_imgs(ctx_x), cfg["img_agg"], True)
        ft = cnn_features(sd, "img_encoder.", _flat_imgs(tgt_x), cfg["img_agg"])
        x_ctx = fc.view(T, -1, fc.size(1))
        x_tgt = ft.view(T, -1, ft.size(1))
        cf = task_encoder(sd, torch.cat([x_ctx, lab], dim=2))
        rep, att = multihead_attention(sd, x_ctx, cf, x_tgt, want_inter=True)
        sample = linear(sd, "mu.", rep)
        if inter is not None:
            inter.update(x_ctx=x_ctx, x_tgt=x_tgt, pool_idx=pidx, ctx_feat=cf, rep=rep, **att)
    else:
        sample = torch.zeros(T, nt, 256, dtype=tgt_x.dtype)  # ANPDistractor.py:128
    return np_decoder(sd, _flat_imgs(tgt_x), sample, T, cfg["img_agg"])


def cnp_shapenet1d_forward(sd, cfg, ctx_x, ctx_y, tgt_x, inter=None):
    """networks/CNPShapeNet1D.py:95-143."""
    T, nc, nt = cfg["tasks_per_batch"], ctx_x.shape[1], tgt_x.shape[1]
    dim_w, dim_z = cfg["dim_w"], cfg["dim_z"]
    if nc:
        x_ctx = encoder_w0(sd, _flat_imgs(ctx_x)).reshape(T, nc, dim_w)
        lab = linear(sd, "transform_y.", ctx_y)
        rs = encoder_fc(sd, torch.cat([x_ctx, lab], dim=2))
        r, aidx = aggregate(rs, cfg["agg_mode"], True)
        z = linear(sd, "r_to_z.", r)[:, None, :].repeat(1, nt, 1)
        if inter is not None:
            inter.update(x_ctx=x_ctx, rs=rs, agg=r, agg_idx=aidx)
    else:
        z = torch.zeros(T, nt, dim_z, dtype=tgt_x.dtype)
    x_qry = encoder_w0(sd, _flat_imgs(tgt_x)).reshape(T, nt, dim_w)
    return decoder0(sd, torch.cat([x_qry, z], dim=-1))


def anp_shapenet1d_forward(sd, cfg, ctx_x, ctx_y, tgt_x, inter=None):
    """networks/ANPShapeNet1D.py:118-161 (target images are encoded first, :130-131)."""
    T, nc, nt = cfg["tasks_per_batch"], ctx_x.shape[1], tgt_x.shape[1]
    dim_w, dim_z = cfg["dim_w"], cfg["dim_z"]
    x_qry = encoder_w0(sd, _flat_imgs(tgt_x)).reshape(T, nt, dim_w)
    if nc:
        x_ctx = encoder_w0(sd, _flat_imgs(ctx_x)).reshape(T, nc, dim_w)
        lab = linear(sd, "transform_y.", ctx_y)
        rs = encoder_fc(sd, torch.cat([x_ctx, lab], dim=2))
        if cfg["agg_mode"] != "attention":
            raise TypeError("agg_mode is not applicable for CNP, choose from ['attention']")
        r, att = multihead_attention(sd, x_ctx, rs, x_qry, want_inter=True)
        z = linear(sd, "r_to_z.", r)
        if inter is not None:
            inter.update(x_ctx=x_ctx, x_tgt=x_qry, rs=rs, rep=r, **att)
    else:
        z = torch.zeros(T, nt, dim_z, dtype=tgt_x.dtype)
    return decoder0(sd, torch.cat([x_qry, z], dim=-1))


FORWARD = {
    "CNPDistractor": cnp_distractor_forward,
    "CondNeuralProcess": cnp_distractor_forward,
    "ANPDistractor": anp_distractor_forward,
    "ANP": anp_distractor_forward,
    "CNPShapeNet1D": cnp_shapenet1d_forward,
    "ANPShapeNet1D": anp_shapenet1d_forward,
}


# --------------------------------------------------------------------------------------------
# Losses  (trainer/losses.py:32-80)
# --------------------------------------------------------------------------------------------
def quaternion_loss(q_gt, q_pr):
    """trainer/losses.py:50-57."""
    q_pr = q_pr / torch.sqrt((q_pr ** 2).sum(dim=-1, keepdim=True))
    pos = (q_gt - q_pr).abs().sum(dim=-1)
    neg = (-q_gt - q_pr).abs().sum(dim=-1)
    return torch.minimum(pos, neg).mean()


def degree_loss(q_gt, q_pr):
    """trainer/losses.py:63-76 (evaluation only, no gradient needed)."""
    gt = torch.rad2deg(q_gt[..., -1])
    pr_cos, pr_sin = q_pr[..., 0], q_pr[..., 1]
    deg = torch.acos(pr_cos)
    deg = torch.where(pr_sin < 0, -deg + 2 * math.pi, deg)
    deg = torch.rad2deg(deg)
    err = torch.stack(((gt - deg).abs(), (gt + 360.0 - deg).abs(), (gt - (deg + 360.0)).abs()), -1)
    return err.min(dim=-1).values.mean()


def calc_loss(task, mu, y, test=False):
    """LossFunc.calc_loss with loss_type == "mse" (the only implemented type), losses.py:32-48."""
    if task == "distractor":
        return torch.sqrt(((y - mu) ** 2).sum(dim=-1)).mean()  # :35-36
    if task == "shapenet_3d":
        return quaternion_loss(y, mu)
    if task == "shapenet_1d":
        if test:
            return degree_loss(y, mu)
        return ((y[..., :2] - mu) ** 2).sum(dim=-1).mean()  # :59-61
    raise ValueError(task)


# --------------------------------------------------------------------------------------------
# Whole meta-train step (the unit BASELINE.json's metric counts): trainer/model_trainer.py:59-93
# --------------------------------------------------------------------------------------------
def params_with_grad(sd):
    """Leaf fp tensors that are Parameters in the reference (buffers excluded)."""
    return {k: v for k, v in sd.items() if k != "attn.projection_matrix"}


class OracleTrainer:
    """zero_grad -> forward -> loss -> backward -> Adam, on CPU.  Used as the `port` CPU baseline
    and as the multi-step checker.  torch.optim.Adam defaults as train.py:52-56 (lr from cfg)."""

    def __init__(self, method, cfg, state_dict, lr=1e-4, dtype=torch.float32):
        self.method, self.cfg = method, dict(cfg)
        self.sd = {k: v.detach().clone().to(dtype) for k, v in state_dict.items()}
        self.params = params_with_grad(self.sd)
        for v in self.params.values():
            v.requires_grad_(True)
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr)

    def forward_loss(self, ctx_x, ctx_y, tgt_x, tgt_y, inter=None):
        mu = FORWARD[self.method](self.sd, self.cfg, ctx_x, ctx_y, tgt_x, inter)
        return mu, calc_loss(self.cfg["task"], mu, tgt_y)

    def step(self, ctx_x, ctx_y, tgt_x, tgt_y):
        self.opt.zero_grad(set_to_none=True)
        mu, loss = self.forward_loss(ctx_x, ctx_y, tgt_x, tgt_y)
        loss.backward()
        self.opt.step()
        return float(loss.detach())

    def grads(self):
        return {k: (None if v.grad is None else v.grad.detach().clone())
                for k, v in self.params.items()}


# --------------------------------------------------------------------------------------------
# Closed-form FAVOR+ backward -- checked against autograd in the tests and
# used as the line-by-line specification for the CUDA backward kernel.
# --------------------------------------------------------------------------------------------
def favor_attention_fwd_bwd(xq, xk, v, proj, d_out, eps=FAVOR_EPS):
    """xq [B,H,nt,d], xk [B,H,nc,d], v [B,H,nc,e], proj [M,d], d_out [B,H,nt,e].
    Returns out and (dxq, dxk, dv).  Row-argmax: first index; global argmax: even split on ties."""
    d = xq.shape[-1]
    M = proj.shape[0]
    c = d ** -0.25
    rho = M ** -0.5
    P = proj.to(xq.dtype)
    U = c * xq @ P.t()
    W = c * xk @ P.t()
    s = c * c * (xq ** 2).sum(-1, keepdim=True) / 2
    t = c * c * (xk ** 2).sum(-1, keepdim=True) / 2
    m, am = U.max(dim=-1, keepdim=True)
    g = W.max()
    Qp = rho * (torch.exp(U - s - m) + eps)
    Kp = rho * (torch.exp(W - t - g) + eps)
    A = Qp @ Kp.transpose(-1, -2)
    D = A.sum(-1, keepdim=True)
    out = (A @ v) / D
    # backward
    dA = (d_out @ v.transpose(-1, -2) - (d_out * out).sum(-1, keepdim=True)) / D
    dv = (A / D).transpose(-1, -2) @ d_out
    dQp = dA @ Kp
    dKp = dA.transpose(-1, -2) @ Qp
    dU = dQp * (Qp - rho * eps)
    ds = -dU.sum(-1, keepdim=True)
    dU = dU.scatter_add(-1, am, ds)
    dW = dKp * (Kp - rho * eps)
    dt = -dW.sum(-1, keepdim=True)
    dg = -dW.sum()
    tie = (W == g).to(W.dtype)
    dW = dW + dg * tie / tie.sum()
    dxq = c * (dU @ P) + c * c * ds * xq
    dxk = c * (dW @ P) + c * c * dt * xk
    return out, (dxq, dxk, dv)
