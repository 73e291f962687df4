"""debug: per-tensor gradient errors of the seg model vs the fp64 oracle (tests/test_gpu_backward.py)"""
import sys, torch
sys.path.insert(0, ".")
from tests.test_gpu_backward import _prepared
from oracle import harness
from samble_b200.testing import synthetic_clouds
which = sys.argv[1] if len(sys.argv) > 1 else "seg"
train = len(sys.argv) > 2 and sys.argv[2] == "train"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
m, sd, cfg = _prepared(which, B, 256, (128, 64), seed=4, train=train)
x, cat = synthetic_clouds(B, 256, 6)
rep = harness.gradient_parity(m, sd, cfg, x, cat, which=which)
print("logits", rep["logits_close_frac"], rep["logits_close_frac_fp64"], "scale", rep["model_grad_scale"], "lrelu calls", rep["lrelu_calls"], "elements", rep["lrelu_elements"], "flips", rep["lrelu_flips_vs_fp32"], rep["lrelu_flips_vs_fp64"], "unconsumed", rep["lrelu_masks_unconsumed"])
for n, e in list(rep["params"].items()) + [("<input>", rep["input"])]:
    flag = "  <<<" if e["rel"] > 1e-4 and e["native"] > 8 * e["oracle32"] else ""
    print(f"{n:55s} rel {e['rel']:.1e} ref32 {e['rel_oracle32']:.1e} own {e['own_scale']:.1e}{flag}")
