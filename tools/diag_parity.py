"""Teacher-forced parity report (oracle/harness.py) for a few model sizes and kernel modes -> gpurun_out/diag_parity.json."""
import json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import harness
from samble_b200 import _lib as L, models
from samble_b200.config import cls_config, seg_config
from samble_b200.testing import fill_state_dict_, synthetic_clouds

cases = [("seg", 2, 256, (128, 64), 1, 2), ("seg", 2, 2048, (1024, 512), 5, 6), ("cls", 4, 1024, (512, 256), 7, 8)]
if len(sys.argv) > 1 and sys.argv[1] == "full":
    cases += [("seg", 16, 2048, (1024, 512), 3, 5), ("cls", 32, 1024, (512, 256), 3, 5)]
out = {}
for which, B, N, M, wseed, xseed in cases:
    cfg = (seg_config if which == "seg" else cls_config)(M=M)
    m = (models.ShapeNetModel if which == "seg" else models.ModelNetModel)(cfg)
    sd = fill_state_dict_(m.state_dict(), seed=wseed, sharpen=4.0)
    m.load_state_dict(sd)
    m = m.eval().cuda()
    x, cat = synthetic_clouds(B, N, xseed)
    xc, catc = synthetic_clouds(B, N, xseed + 100)       # calibration batch != evaluation batch (no cut sits exactly on a z)
    with torch.no_grad():
        m(xc.cuda(), catc.cuda()) if which == "seg" else m(xc.cuda())
    models.freeze_boundaries(m)
    from samble_b200 import blocks
    for mode_name, knn_mode, ds_exact in (("default", 0, True), ("3xtf32-ds", 0, False)):
        L.lib().samble_set_knn_mode(knn_mode)
        blocks.DS_EXACT = ds_exact
        t = time.time()
        rep = harness.forward_parity(m, sd, cfg, x, cat, which=which)
        L.lib().samble_set_knn_mode(0)
        blocks.DS_EXACT = True
        key = f"{which}_B{B}_N{N}_{mode_name}"
        out[key] = rep
        print(key, f"({time.time() - t:.1f}s):", harness.brief(rep), flush=True)
        try:
            harness.assert_report(rep)
            print("   PASS")
        except AssertionError as e:
            print("   FAIL:", str(e)[:600])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/diag_parity.json", "w"), indent=1, default=str)
