"""Not a test: prints per-tensor errors for the full-size configs (run on the GPU box)."""
import sys
import torch
sys.path.insert(0, ".")
import mmemo_b200
from mmemo_b200 import ops, synth
from oracle import mmemo_oracle as O
from tests import cases
from tests.cases import rel_err
from tests.test_gpu_models import Loss, to_dev, _model_and_state

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
if which == "cfg2":
    dim, H, nl = 512, 8, int(sys.argv[2]) if len(sys.argv) > 2 else 2
    m, sd = _model_and_state(lambda: cases._RefChain(mmemo_b200.realformer.Attention_Block, dim, H, nl))
    b = synth.encoder_batch(seed=1234, B=int(sys.argv[3]) if len(sys.argv) > 3 else 8, L=128, d=dim)
    pres = [f"blocks.{i}." for i in range(nl)]
    ref_out, ref_loss, ref_grads, ref_ig = cases.run_with_grads(
        lambda s, bb: O.encoder_chain(s, pres, bb["x"], bb["mask"], H)[0], sd, b, cases._sq_mean, O, ["x"])
    c = cases.CASES["encoder_chain"]
    out, loss, grads, ig = cases.run_module_with_grads(m, c, to_dev(b), Loss)
    print("out", rel_err(out, ref_out), "dx", rel_err(ig["x"], ref_ig["x"]))
    for k, v in ref_grads.items():
        print(f"{k:32s} {rel_err(grads[k], v):.3e}  |ref|max {v.abs().max().item():.3e}")
else:
    kw = dict(l_dim=300, v_dim=35, a_dim=74, dim=96, l_len=50, v_len=50, a_len=50, n_heads=6, n_layers=2, ffn=2)
    m, sd = _model_and_state(lambda: mmemo_b200.realformer.State_Transfer(**kw))
    b = synth.realformer_batch(seed=1234, B=32, P=6, empty_windows=(which == "cfg1a"))
    c = cases.CASES["realformer_state_transfer"]
    ref_logits, ref_loss, ref_grads, _ = cases.run_with_grads(
        lambda s, bb: O.realformer_state_transfer(s, bb["l"], bb["v"], bb["a"], bb["l_mask"], bb["v_mask"], bb["a_mask"], 6, 2),
        sd, b, c.loss, O, [])
    logits, loss, grads, _ = cases.run_module_with_grads(m, cases.Case(**{**c.__dict__, "grad_inputs": []}), to_dev(b), Loss)
    live = b["wmask"].bool()
    print("logits live", rel_err(logits.cpu()[live], ref_logits[live]), "all", rel_err(logits, ref_logits), "loss", loss.item(), ref_loss.item())
    errs = sorted(((rel_err(grads[k], v), k) for k, v in ref_grads.items()), reverse=True)
    for e, k in errs[:25]:
        print(f"{k:60s} {e:.3e} |ref|max {ref_grads[k].abs().max().item():.3e}")
