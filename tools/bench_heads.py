"""Post-RoI heads: the library's tcgen05 route against the same torch modules on cuDNN.
Prints one JSON line per case (CUDA events, 5 warm-up + 20 timed passes, allow_tf32 as given)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgn_b200 import FCNMaskHead
from fgn_b200.roi_head import make_c4_shared_head

dev = torch.device("cuda:0")


def timed(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    prec = "tf32" if tf32 else "fp32"
    for r in (300, 1000):
        head = make_c4_shared_head(1024, 512, 3).to(dev).eval()
        x = torch.randn(r, 7, 7, 1024, device=dev).permute(0, 3, 1, 2)
        with torch.no_grad():
            for m in head:
                m.tc_1x1, m.tc_3x3 = True, True              # (True: the library's 3x3 under strict fp32 too)
            a = head(x)
            t_tc = timed(lambda: head(x))
            for m in head:
                m.tc_3x3 = False
            t_13 = timed(lambda: head(x))
            for m in head:
                m.tc_1x1 = False
            b = head(x)
            t_dnn = timed(lambda: head(x))
        flop = r * 49 * 3 * (2 * 1024 * 512 * 2 + 512 * 512 * 9 * 2)
        print(json.dumps({"case": "res5 shared_head (3 bottlenecks 1024-512-1024 @7x7)", "R": r, "precision": prec,
                          "us_tcgen05_all": round(t_tc, 1), "us_tcgen05_1x1_cudnn_3x3": round(t_13, 1), "us_cudnn": round(t_dnn, 1),
                          "tflops_tcgen05": round(flop / t_tc / 1e6, 1), "tflops_cudnn": round(flop / t_dnn / 1e6, 1),
                          "max_abs_diff": float((a - b).abs().max()), "out_absmax": float(b.abs().max())}), flush=True)
    for r, cin in ((100, 256), (100, 1024), (1600, 256)):
        mh = FCNMaskHead(num_convs=4, in_channels=cin, conv_out_channels=256, num_classes=1, class_agnostic=True).to(dev).eval()
        x = torch.randn(r, 14, 14, cin, device=dev).permute(0, 3, 1, 2)
        with torch.no_grad():
            a = mh(x)
            t_tc = timed(lambda: mh(x))
            mh.tc = False
            mh_cl = mh.to(memory_format=torch.channels_last)
            b = mh_cl(x)
            t_dnn = timed(lambda: mh_cl(x))
        flop = r * 196 * (cin * 256 * 9 * 2 + 3 * 256 * 256 * 9 * 2 + 256 * 1024 * 2) + r * 784 * 256 * 2
        print(json.dumps({"case": "FCNMaskHead (4x conv3x3 + deconv2x2 + logits) @14x14", "R": r, "Cin": cin, "precision": prec,
                          "us_tcgen05": round(t_tc, 1), "us_cudnn": round(t_dnn, 1), "tflops_tcgen05": round(flop / t_tc / 1e6, 1),
                          "tflops_cudnn": round(flop / t_dnn / 1e6, 1), "max_abs_diff": float((a - b).abs().max()),
                          "out_absmax": float(b.abs().max())}), flush=True)
