"""CUDA graphs behind the unchanged ``nn.Module`` contract.

The reference's trainer drives the model as ``zero_grad -> .to(device) -> model(...) -> calc_loss ->
losses.backward() -> optimizer.step() -> losses.item()`` (trainer/model_trainer.py:59-93) with its own
``torch.optim.Adam``.  Launched eagerly, that costs ~15 ms of Python / ctypes / autograd enqueue time per step
against ~8 ms of GPU work.  ``GraphedModel`` keeps the call sequence and replaces what happens inside it:

* ``model(ctx_x, ctx_y, tgt_x)`` copies the inputs into static buffers and replays ONE graph holding every
  forward kernel; the returned ``mu`` carries a single autograd node;
* ``losses.backward()`` reaches that node, which copies ``d loss / d mu`` into a static buffer and replays ONE
  graph holding every backward kernel; the parameter gradients come back as views of one flat buffer (a fresh
  copy per step, so ``p.grad`` tensors of different steps never alias);
* under ``torch.no_grad()`` (validation, ``evaluation.py``) a forward-only graph is used.

Graphs are cached per (grad mode, input shapes): the reference draws ``shot ~ U{1..max_ctx_num}`` per batch
(dataset/shapenet_distractor.py:197), so at most ``max_ctx_num`` train shapes occur.  All graphs share one memory
pool -- they never run concurrently.  Parameters are read through their storage (optimizers update them in
place), so any optimizer works; ``.to()`` / re-allocation of parameters after the first call invalidates the
cache (checked by data pointer).  ``mu`` is a static buffer: it is overwritten by the next call with the same
shapes, which is how the reference's loops use it.

Enable with ``model.enable_cuda_graphs()`` or ``B200NP_GRAPHS=1`` (``b200_run.py`` sets it).
"""
import torch
from torch.autograd import Function


class _Replay(Function):
    @staticmethod
    def forward(ctx, ent, *params):
        ent["fwd"].replay()
        ctx.ent = ent
        return ent["mu"].detach()

    @staticmethod
    def backward(ctx, dmu):
        ent = ctx.ent
        ent["dmu"].copy_(dmu)
        ent["bwd"].replay()
        flat = ent["gflat"].clone()        # fresh storage per step: gradients handed to autograd never alias
        grads = []
        for off, n, shape, used in ent["gviews"]:
            grads.append(flat[off:off + n].view(shape) if used else None)
        return (None,) + tuple(grads)


class GraphedModel:
    def __init__(self, model, warmup=2):
        self.model, self.warmup = model, warmup
        self.cache = {}
        self.pool = None
        self.side = None
        self._param_sig = None

    def _params(self):
        return [p for p in self.model.parameters() if p.requires_grad]

    def _check_params(self):
        sig = tuple(p.data_ptr() for p in self.model.parameters())
        if sig != self._param_sig:
            self.cache.clear()             # parameters moved (.to(), load with assign=True): graphs hold stale pointers
            self._param_sig = sig

    def _build(self, key, inputs, train):
        from torch.nn.utils.stateless import _reparametrize_module

        from . import engine, ops
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
            self.side = torch.cuda.Stream(priority=engine.MAIN_PRIORITY if engine.USE_PRIORITIES else 0)
        model = self.model
        static = [torch.empty_like(t) for t in inputs]
        for s, t in zip(static, inputs):
            s.copy_(t)
        ent = {"static": static}
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        params = [p for _, p in named]
        # Warm-up and capture see the parameters through fresh leaf ALIASES (same storage): their gradient
        # accumulators are created on the capturing streams and die with this build.  The model's own parameters may
        # carry accumulators from an earlier iteration (kept alive by a retained loss tensor) that live on the
        # caller's stream; autograd would then pull that stream into the capture and invalidate it.  It also leaves
        # the caller's `.grad` fields alone.
        alias = {n: p.detach().requires_grad_(True) for n, p in named}
        leaves = [alias[n] for n, _ in named]

        def fwd_bwd(dmu):
            with _reparametrize_module(model, alias):
                mu = engine.forward(model, *static)
            if train:
                mu.backward(dmu if dmu is not None else torch.ones_like(mu))
            return mu
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side), torch.set_grad_enabled(train):
            for _ in range(self.warmup):
                fwd_bwd(None)
                for a in leaves:
                    a.grad = None
        cur.wait_stream(self.side)
        torch.cuda.synchronize()
        fwd = torch.cuda.CUDAGraph()
        with torch.set_grad_enabled(train):
            with torch.cuda.graph(fwd, stream=self.side, pool=self.pool):
                with _reparametrize_module(model, alias):
                    mu = engine.forward(model, *static)
        ent["fwd"], ent["mu"] = fwd, mu
        if train:
            with torch.cuda.stream(self.side):
                dmu = torch.zeros_like(mu)
            torch.cuda.synchronize()
            bwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(bwd, stream=self.side, pool=self.pool):
                # .backward(), not autograd.grad(inputs=...): the engine joins the streams of executed accumulator
                # nodes (decoder CNN / weight-gradient companions) back into the caller's; captured inputs are not joined
                mu.backward(dmu)
                grads = [a.grad for a in leaves]
                # one flat buffer (16-byte aligned segments), filled by a multi-copy launch per 64 tensors
                pad = lambda k: (k + 3) // 4 * 4
                total = sum(pad(g.numel()) for g in grads if g is not None)
                gflat = torch.empty(max(total, 4), device=mu.device, dtype=torch.float32)
                segs, views, off = [], [], 0
                keep = []
                for p, g in zip(params, grads):
                    if g is None:
                        views.append((0, 0, tuple(p.shape), False))
                        continue
                    g = g.contiguous()
                    keep.append(g)
                    segs.append((g, off, g.numel()))
                    views.append((off, g.numel(), tuple(p.shape), True))
                    off += pad(g.numel())
                ops.multi_copy(gflat, segs)
            ent.update(bwd=bwd, dmu=dmu, gflat=gflat, gviews=views, params=params)
            # drop the captured autograd graph: its saved activations go back to the (graph-private) pool, where the
            # two graphs of this entry keep using them and later captures may share the addresses
            ent["mu"] = mu.detach()
            for a in leaves:
                a.grad = None
            del mu, grads, keep
        torch.cuda.synchronize()
        self.cache[key] = ent
        return ent

    def __call__(self, ctx_x, ctx_y, tgt_x):
        from . import engine
        engine._check_inputs(self.model, ctx_x, ctx_y, tgt_x)
        train = torch.is_grad_enabled()
        self._check_params()
        inputs = (ctx_x, ctx_y, tgt_x)
        key = (train, engine.PRECISION) + tuple(tuple(t.shape) for t in inputs)
        ent = self.cache.get(key)
        if ent is None:
            ent = self._build(key, inputs, train)
        for s, t in zip(ent["static"], inputs):
            s.copy_(t, non_blocking=True)
        if not train:
            ent["fwd"].replay()
            return ent["mu"]
        return _Replay.apply(ent, *ent["params"])
