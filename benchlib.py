"""Workloads of bench.py: one entry per BASELINE.json config, each knowing how to build OUR drop-in
model, a seeded synthetic batch, the training (or inference) step through the public module API,
and the same step through the CPU oracle (the reference algorithm's port; used only by the
``cpu_baseline`` / ``gpu_eager_baseline`` / ``--impl reference`` legs, never by the measured path).

    cfg1a  others/realformer.py   State_Transfer(300,35,74,96,50,50,50,6,2,2), B=32 x P=6 windows
    cfg1b  cmu-mosei/run.py       Concat_Trans(96,20,100,200,6,1,1), B=32 (two towers)
    cfg2   others/realformer.py   6 x Attention_Block(512, 8) chain, B=64, L=128     (headline)
    cfg3   rencecps/run.py        Concat_Linear(2304), B=128 (faithful); composite = chain at
                                  (128, 256, 512) -> cls/max/mean pooling -> Concat_Linear(1536)
    cfg4   Ren-MME/run.py         Base_model(), GLOBAL batch 256 sharded over the ranks (strong scaling)
    cfg5   robot_demo.py          4 x Multi_class(192,25,100,100,6,2,2) ensemble, eval, B=1 / 32, p50
"""
from __future__ import annotations

import time
from typing import Callable, Dict, List, Optional

import torch

import mmemo_b200
from mmemo_b200 import ops, synth


def _seeded(model: torch.nn.Module, seed: int = 1) -> Dict[str, torch.Tensor]:
    """state_dict with the zero-initialised ReZero gates a, b, c drawn from U(-0.5, 0.5) (at their
    init every attention / FFN gradient is exactly zero, SURVEY §8a note 6)."""
    sd = synth.randomize_gates({k: v.detach().clone() for k, v in model.state_dict().items()}, seed)
    model.load_state_dict(sd)
    return sd


class Workload:
    """name, samples per step per rank, how to build / call.  ``inputs`` are the float tensors the
    step consumes (copied H2D every e2e step); ``loss(model, dev_batch)`` returns a 0-dim tensor."""
    name = ""
    cfg = ""
    metric = "train samples/s (fwd+bwd)"
    mode = "train"
    description = ""
    samples_per_item = 1

    def __init__(self, batch: int):
        self.batch = batch

    def model(self) -> torch.nn.Module:
        raise NotImplementedError

    def host_batch(self, seed: int) -> Dict[str, torch.Tensor]:
        raise NotImplementedError

    def loss(self, model, b) -> torch.Tensor:
        raise NotImplementedError

    def oracle_loss(self, O, sd, b) -> torch.Tensor:
        raise NotImplementedError


def _flat(b) -> List[torch.Tensor]:
    out = []
    for v in (b.values() if isinstance(b, dict) else b):
        if torch.is_tensor(v):
            out.append(v)
        elif isinstance(v, (list, tuple)):
            out.extend(_flat(v))
    return out


class Cfg1a(Workload):
    name, cfg = "cfg1a_realformer_state_transfer", "configs[0] (others/realformer.py defaults)"
    description = ("State_Transfer(300,35,74,d=96,L=50/50/50,6 heads,2 layers,ffn 2), B=32 x P=6 "
                   "windows, loss=(circle*wmask).mean()")
    KW = dict(l_dim=300, v_dim=35, a_dim=74, dim=96, l_len=50, v_len=50, a_len=50, n_heads=6,
              n_layers=2, ffn=2)

    def model(self):
        return mmemo_b200.realformer.State_Transfer(**self.KW)

    def host_batch(self, seed):
        return synth.realformer_batch(seed=seed, B=self.batch, P=6)

    def loss(self, m, b):
        out = m(b["l"], b["v"], b["a"], b["l_mask"], b["v_mask"], b["a_mask"])
        return (ops.circle_loss_op(out, b["label"]) * b["wmask"]).mean()

    def oracle_loss(self, O, sd, b):
        out = O.realformer_state_transfer(sd, b["l"], b["v"], b["a"], b["l_mask"], b["v_mask"],
                                          b["a_mask"], 6, 2)
        return O.window_masked_loss(out, b["label"], b["wmask"])


class Cfg1b(Workload):
    name, cfg = "cfg1b_mosei_concat_trans", "configs[0] (cmu-mosei/run.py native lengths)"
    description = ("Concat_Trans(d=96,L=20/100/200,6 heads,1 layer), dims 300/35/74, B=32, two "
                   "towers + 7x7x7 bilinear head, loss=circle.mean()")
    KW = dict(dim=96, l_len=20, v_len=100, a_len=200, n_heads=6, n_layers=1, ffn=1)

    def model(self):
        return mmemo_b200.cmu_mosei.Concat_Trans(**self.KW)

    def host_batch(self, seed):
        return synth.mosei_batch(seed=seed, B=self.batch)

    def loss(self, m, b):
        out = m(b["l"], b["v"], b["a"], b["l_mask"], b["v_mask"], b["a_mask"])
        return ops.circle_loss_op(out, b["label"]).mean()

    def oracle_loss(self, O, sd, b):
        out = O.mosei_concat_trans(sd, b["l"], b["v"], b["a"], b["l_mask"], b["v_mask"], b["a_mask"],
                                   6, 1)
        return O.multi_circle_loss(out, b["label"]).mean()


class Cfg3Faithful(Workload):
    name, cfg = "cfg3_rencecps_concat_linear", "configs[2] (rencecps/run.py as written)"
    description = "Concat_Linear(2304) on pooled BERT features (B,2,2304), B=128, loss=circle.mean()"

    def model(self):
        return mmemo_b200.rencecps.Concat_Linear(2304)

    def host_batch(self, seed):
        return synth.rencecps_batch(seed=seed, B=self.batch)

    def loss(self, m, b):
        return ops.circle_loss_op(m(b["feat"]), b["label"]).mean()

    def oracle_loss(self, O, sd, b):
        return O.multi_circle_loss(O.rencecps_concat_linear(sd, b["feat"]), b["label"]).mean()


class CompositeText(torch.nn.Module):
    """BASELINE configs[2] as worded ("seq 256, batch 128, RealFormer encoder"): a synthetic
    composite of reference classes — chain of others/realformer.py Attention_Blocks at
    (B, 256, 512) -> [cls | max | mean] pooling as in rencecps/run.py:103-109 (flatten_array) ->
    Concat_Linear(3*512) on (previous, current) sentence features."""

    def __init__(self, dim=512, n_heads=8, n_layers=6):
        super().__init__()
        self.encoder = mmemo_b200.ResidualEncoder(dim, n_heads, n_layers, 2)
        self.head = mmemo_b200.rencecps.Concat_Linear(3 * dim)

    def forward(self, x, mask):
        """x (B, 2, L, d): index 0 = previous sentence, 1 = current; mask (B, 2, L)."""
        B, two, L, d = x.shape
        h = self.encoder(x.reshape(B * two, L, d), mask.reshape(B * two, L))
        pooled = ops.pool([h], 1)                       # float32 [mean | max] over the positions
        feat = torch.cat([h[:, 0].float(), pooled[:, d:], pooled[:, :d]], 1)
        return self.head(feat.view(B, two, 3 * d))


def composite_oracle(O, sd, x, mask, n_heads=8, n_layers=6):
    B, two, L, d = x.shape
    pres = [f"encoder.blocks.{i}." for i in range(n_layers)]
    h = O.encoder_chain(sd, pres, x.reshape(B * two, L, d), mask.reshape(B * two, L), n_heads)[0]
    h = h.float() if h.dtype == torch.bfloat16 else h
    feat = torch.cat([h[:, 0], h.max(1)[0], h.mean(1)], 1).view(B, two, 3 * d)
    hp = {k[len("head."):]: v for k, v in sd.items() if k.startswith("head.")}
    if feat.dtype != hp["trans"].dtype:
        hp = {k: v.to(feat.dtype) for k, v in hp.items()}
    return O.rencecps_concat_linear(hp, feat)


class Cfg3Composite(Workload):
    name, cfg = "cfg3_composite_text_encoder", "configs[2] (as worded: seq 256, batch 128)"
    description = ("6 x Attention_Block(512, 8) chain at L=256 on (previous, current) sentences -> "
                   "cls|max|mean pool -> Concat_Linear(1536); `batch` counts sentence PAIRS: 64 pairs "
                   "= the 128 sequences of 256 tokens BASELINE names; samples/s counts sequences")
    samples_per_item = 2

    def __init__(self, batch: int, L: int = 256):
        super().__init__(batch)
        self.L = L

    def model(self):
        return CompositeText()

    def host_batch(self, seed):
        g = synth.gen(seed)
        return {"x": synth.feats(g, self.batch, 2, self.L, 512),
                "mask": synth.prefix_mask(g, (self.batch, 2), self.L),
                "label": synth.labels(g, (self.batch,), 9)}

    def loss(self, m, b):
        return ops.circle_loss_op(m(b["x"], b["mask"]), b["label"]).mean()

    def oracle_loss(self, O, sd, b):
        return O.multi_circle_loss(composite_oracle(O, sd, b["x"], b["mask"]), b["label"]).mean()


class Cfg4(Workload):
    name, cfg = "cfg4_renmme_base_model", "configs[3] (Ren-MME/run.py defaults, global batch 256)"
    description = ("Base_model(d=128,L=40/76/275,8 heads,1 layer), dims 768/640/205, R-Drop pairs, "
                   "loss=multi_loss + symmetric sigmoid-KL; training dropout 0.1 as in the reference "
                   "(Ren-MME/run.py:36; grouped counter-based masks, new on every step) - the CPU / "
                   "eager baselines run the same model without dropout")
    dropout = 0.1

    def model(self):
        mmemo_b200.ren_mme.DROP = self.dropout
        return mmemo_b200.ren_mme.Base_model()

    def host_batch(self, seed):
        b = synth.renmme_batch(seed=seed, B=self.batch)
        out = {f"x{i}": t for i, t in enumerate(b["inputs"])}
        out["label"] = b["label"]
        return out

    def loss(self, m, b):
        logits = m(*[b[f"x{i}"] for i in range(12)])
        return ops.circle_loss_op(logits, b["label"]).mean() + ops.rdrop_kl_op(logits)

    def oracle_loss(self, O, sd, b):
        logits = O.renmme_base_model(sd, *[b[f"x{i}"] for i in range(12)])
        return O.multi_loss(logits, b["label"]) + O.rdrop_kl(logits)


class Cfg5(Workload):
    name, cfg = "cfg5_robot_demo_ensemble", "configs[4] (robot_demo.py native lengths 25/100/100)"
    metric, mode = "p50 request latency (ms), 4-model ensemble forward", "infer"
    description = ("4 x Multi_class(d=192,L=25/100/100,6 heads,2 layers,ffn 2), eval, no_grad; one "
                   "request = pinned host inputs -> H2D -> ensemble -> D2H of the (B,7) prediction")
    KW = dict(dim=192, l_len=25, v_len=100, a_len=100, n_heads=6, n_layers=2, ffn=2)
    NAMES = ("l", "v_256", "v_512", "v_1024", "a", "l_mask", "v_mask", "a_mask")

    def models(self):
        mmemo_b200.robot_demo.DROP = 0.0
        ms, sds = [], []
        for i in range(4):
            torch.manual_seed(i)
            m = mmemo_b200.robot_demo.Multi_class(**self.KW)
            sds.append(_seeded(m, seed=10 + i))
            ms.append(m.eval())
        return ms, sds

    def host_batch(self, seed):
        return synth.robot_batch(seed=seed, B=self.batch)

    def oracle_pred(self, O, sds, b):
        preds = [O.robot_multi_class(sd, *[b[k] for k in self.NAMES], 6, 2) for sd in sds]
        p = preds[0]
        for q in preds[1:]:
            p = p + q
        return p / len(preds)


# ------------------------------------------------------------------------------------------------
# generic measurement of one training workload on the current device
# ------------------------------------------------------------------------------------------------
def _to(b, dev, dtype=None):
    out = {}
    for k, v in b.items():
        t = v.to(dev)
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        out[k] = t
    return out


class GraphedStep:
    """zero_grad -> forward -> loss -> backward of a workload, captured once into a CUDA graph
    over static device inputs (eager fallback if capture fails)."""

    def __init__(self, wl: Workload, dev, seed: int, reducer_factory: Optional[Callable] = None,
                 use_graph: bool = True):
        self.wl, self.dev = wl, dev
        torch.manual_seed(0)
        self.model = wl.model()
        self.sd = _seeded(self.model)
        self.model = self.model.to(dev).train()
        self.host = {k: (v.pin_memory() if v.is_floating_point() or v.dtype == torch.int64 else v)
                     for k, v in wl.host_batch(seed).items()}
        self.static = _to(self.host, dev)
        self.loss_dev = torch.zeros((), device=dev)
        self.host_loss = torch.zeros((), pin_memory=True)
        self.reducer = reducer_factory(self.model) if reducer_factory else None
        self.graph = None
        self.launches = 0
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if use_graph:
            try:
                ops.clear_shadow_cache()
                self.model.zero_grad(set_to_none=True)
                n0 = ops.launch_count
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step()
                self.launches = ops.launch_count - n0
                self.graph = g
            except Exception as ex:  # pragma: no cover
                import sys
                print(f"[bench] {wl.name}: CUDA-graph capture failed ({type(ex).__name__}: {ex}); "
                      "eager mode", file=sys.stderr)
                self.graph = None
                torch.cuda.synchronize()
        if self.graph is None:
            n0 = ops.launch_count
            self._step()
            self.launches = ops.launch_count - n0

    def _step(self):
        self.model.zero_grad(set_to_none=True)
        if getattr(self.wl, "dropout", 0.0) > 0:
            ops.advance_dropout_step(self.dev)     # captured: every graph replay draws new masks
        loss = self.wl.loss(self.model, self.static)
        if self.reducer is not None:
            self.reducer.backward(loss)
        else:
            loss.backward()
        self.loss_dev.copy_(loss.detach())

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()

    def h2d_bytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.host.values())

    def run_e2e(self) -> float:
        """One step from pinned host inputs to the loss on the host (synchronous per step)."""
        for k, v in self.host.items():
            self.static[k].copy_(v, non_blocking=True)
        self.run()
        self.host_loss.copy_(self.loss_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.host_loss)

    def run_e2e_pipelined(self, n: int) -> float:
        """n steps driven like a real input pipeline: the H2D copy of step i+1 (pinned host ->
        one of two device staging sets, on a copy stream) overlaps the compute of step i; the
        compute stream copies the staged batch into the graph's static inputs, replays, and sends
        the loss to the host, where it is read one step later.  Every step's inputs cross PCIe and
        every step's loss reaches the host inside the timed region."""
        if not hasattr(self, "_pipe"):
            self._pipe = dict(
                copy=torch.cuda.Stream(),
                stage=[{k: torch.empty_like(v) for k, v in self.static.items()} for _ in range(2)],
                h2d=[torch.cuda.Event() for _ in range(2)], free=[torch.cuda.Event() for _ in range(2)],
                done=[torch.cuda.Event() for _ in range(2)],
                loss=[torch.zeros((), pin_memory=True) for _ in range(2)])
        P = self._pipe
        cur = torch.cuda.current_stream()

        def issue(i):
            s_ = i % 2
            with torch.cuda.stream(P["copy"]):
                P["copy"].wait_event(P["free"][s_])         # compute no longer reads this stage
                for k, v in self.host.items():
                    P["stage"][s_][k].copy_(v, non_blocking=True)
                P["h2d"][s_].record(P["copy"])

        issue(0)
        for i in range(n):
            s_ = i % 2
            if i + 1 < n:
                issue(i + 1)
            cur.wait_event(P["h2d"][s_])
            torch._foreach_copy_(list(self.static.values()), list(P["stage"][s_].values()))
            P["free"][s_].record(cur)
            self.run()
            P["loss"][s_].copy_(self.loss_dev, non_blocking=True)
            P["done"][s_].record(cur)
            if i >= 1:
                P["done"][1 - s_].synchronize()
                float(P["loss"][1 - s_])
        P["done"][(n - 1) % 2].synchronize()
        return float(P["loss"][(n - 1) % 2])


def time_events(fn: Callable, steps: int, warmup: int, barrier: Callable) -> float:
    """ms for `steps` calls of fn, CUDA events on the current stream, barrier+sync on both sides."""
    for _ in range(warmup):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def time_wall(fn: Callable, steps: int, warmup: int) -> List[float]:
    ts = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        fn()
        if i >= warmup:
            ts.append(time.perf_counter() - t)
    return ts


def oracle_train_step(O, wl: Workload, device, dtype=torch.float32, seed: int = 1234):
    """fwd+bwd of the workload through the oracle (reference algorithm in plain torch) on `device`
    — the CPU baseline (device='cpu') and the torch-eager-on-B200 baseline (device='cuda')."""
    torch.manual_seed(0)
    model = wl.model()
    sd = _seeded(model)
    sd = {k: v.to(device=device, dtype=dtype if v.is_floating_point() else v.dtype)
          .requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    b = _to(wl.host_batch(seed), device, dtype)

    def step():
        for v in sd.values():
            v.grad = None
        loss = wl.oracle_loss(O, sd, b).float()
        loss.backward()
        return loss
    return step


WORKLOADS = {"cfg1a": Cfg1a, "cfg1b": Cfg1b, "cfg3": Cfg3Faithful, "cfg3c": Cfg3Composite,
             "cfg4": Cfg4, "cfg5": Cfg5}
