"""Diagnostic: error of the C4 res5 shared_head routes (tcgen05 3xTF32 for all convs / cuDNN fp32 3x3 / cuDNN fp32 all) against fp64 on the CPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import copy, torch
from fgn_b200.roi_head import make_c4_shared_head
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
dev = torch.device("cuda:0")
head = make_c4_shared_head(1024, 512, 3).eval()
x = torch.randn(24, 1024, 7, 7)
with torch.no_grad():
    ref64 = copy.deepcopy(head).double()(x.double())
    ref32 = head(x)
    hd = copy.deepcopy(head).to(dev)
    xd = x.to(dev).contiguous(memory_format=torch.channels_last)
    out = {}
    out["tcgen05_all"] = hd(xd).cpu()
    for m in hd: m.tc_3x3 = False
    out["tcgen05_1x1_cudnn_3x3"] = hd(xd).cpu()
    for m in hd: m.tc_1x1 = False
    out["cudnn_all"] = hd(xd).cpu()
    out["mkl_fp32"] = ref32
scale = float(ref64.abs().max())
for k, v in out.items():
    e = (v.double() - ref64).abs()
    print(json.dumps({"route": k, "max_abs_err_vs_fp64": float(e.max()), "mean_abs_err": float(e.mean()), "mean_signed_err": float((v.double() - ref64).mean()),
                      "out_absmax": scale, "rel_to_scale": float(e.max()) / scale}))
