"""GPU diagnostic: device time of the classifier (nmn.py:75-83) forward + backward and of its individual products on
pnmn_gemm_split, next to the library (cuBLAS / cuDNN, IEEE fp32 and TF32) nn.Sequential."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnmn_clevr_b200.nmn import NeuralModuleNetwork, _gemm
from probnmn_clevr_b200.synthetic import make_nmn_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary

vocab = Vocabulary.clevr()
m = NeuralModuleNetwork(vocab); m.load_state_dict(make_nmn_state_dict(vocab, 0)); m = m.cuda()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B in (123, 256):
    final = torch.randn(B, 128, 14, 14, device="cuda").relu_()
    def run(native):
        m.zero_grad(set_to_none=True)
        f = final.clone().requires_grad_(True)
        out = m._classifier_split(f) if native else m.classifier(f)
        out.square().sum().backward()
    torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
    t_native, t_ieee = timeit(lambda: run(True)), timeit(lambda: run(False))
    torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
    t_tf32 = timeit(lambda: run(False))
    torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
    print(f"B={B}: classifier fwd+bwd  native {t_native:.3f} ms | library IEEE fp32 {t_ieee:.3f} ms | library TF32 {t_tf32:.3f} ms")
    M1 = B * 196
    x1 = torch.randn(M1, 128, device="cuda"); w1 = torch.randn(1024, 128, device="cuda"); g1 = torch.randn(M1, 1024, device="cuda")
    x2 = torch.randn(B, 50176, device="cuda"); w2 = torch.randn(1024, 50176, device="cuda"); g2 = torch.randn(B, 1024, device="cuda")
    prods = {
        "conv fwd   [B*196,128]x[1024,128]^T": lambda: _gemm(x1, 128, 1, w1, 128, 1, M1, 1024, 128),
        "conv dgrad [B*196,1024]x[1024,128] ": lambda: _gemm(g1, 1024, 1, w1, 1, 128, M1, 128, 1024),
        "conv wgrad [1024,B*196]x[B*196,128]": lambda: _gemm(g1, 1, 1024, x1, 1, 128, 1024, 128, M1),
        "fc1 fwd    [B,50176]x[1024,50176]^T": lambda: _gemm(x2, 50176, 1, w2, 50176, 1, B, 1024, 50176),
        "fc1 dgrad  [B,1024]x[1024,50176]   ": lambda: _gemm(g2, 1024, 1, w2, 1, 50176, B, 50176, 1024),
        "fc1 wgrad  [1024,B]x[B,50176]      ": lambda: _gemm(g2, 1, 1024, x2, 1, 50176, 1024, 50176, B),
    }
    for name, fn in prods.items():
        print(f"   {name}: {timeit(fn) * 1e3:7.1f} us")
