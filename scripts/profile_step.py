"""GPU diagnostic: where a bench step spends its time (host vs device, which kernels)."""
import ctypes, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from probnmn_clevr_b200 import _lib as L
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler, make_answers, make_features, make_nmn_state_dict
from probnmn_clevr_b200.vocabulary import Vocabulary

B = int(os.environ.get("B", 256))
vocab = Vocabulary.clevr()
m = NeuralModuleNetwork(vocab); m.load_state_dict(make_nmn_state_dict(vocab, 0)); m = m.cuda().train()
feats = make_features(B, 0).cuda(); progs = ProgramSampler(vocab, seed=0).sample(B, 40).cuda(); ans = make_answers(B, 0).cuda()

def step():
    m.zero_grad(set_to_none=True)
    out = m(feats, progs, ans)
    out["loss"].mean().backward()

for _ in range(3): step()
torch.cuda.synchronize()
# host-side cost of the program compiler
ph = progs.cpu().contiguous()
lib = L.lib()
t0 = time.perf_counter()
for _ in range(10):
    plan = lib.pnmn_plan_create(m._model_handle, ctypes.cast(ph.data_ptr(), ctypes.POINTER(ctypes.c_int64)), B, 40, 1)
    lib.pnmn_plan_destroy(plan)
print(f"plan_create+destroy host ms: {(time.perf_counter()-t0)*100:.3f}")
t0 = time.perf_counter()
for _ in range(5): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"wall per step (launch) {(t1-t0)*200:.2f} ms, incl. drain {(t2-t0)*200:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
