"""Diagnostic: C++ time of pnmn_plan_create on the main thread vs a helper thread while the main thread is idle / runs
Python / launches CUDA kernels (why is the look-ahead compile slower than the inline one?)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from concurrent.futures import ThreadPoolExecutor
from probnmn_clevr_b200 import _lib as L
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler
from probnmn_clevr_b200.vocabulary import Vocabulary
vocab = Vocabulary.clevr()
m = NeuralModuleNetwork(vocab).cuda(); m._ensure_flat()
progs = ProgramSampler(vocab, seed=0).sample(256, 40).contiguous()
lib = L.lib(); h = m._model_handle
ms = (ctypes.c_double * 4)()
def one():
    plan = lib.pnmn_plan_create(h, ctypes.cast(progs.data_ptr(), ctypes.POINTER(ctypes.c_int64)), 256, 40, 1)
    return plan
def ctime(n):
    lib.pnmn_debug_host_times(ms); return ms[0] / n
for _ in range(5): lib.pnmn_plan_destroy(one())
lib.pnmn_debug_host_times(ms)
for _ in range(20): lib.pnmn_plan_destroy(one())
print("main thread                      : %.2f ms" % ctime(20))
pool = ThreadPoolExecutor(max_workers=1)
for _ in range(20): lib.pnmn_plan_destroy(pool.submit(one).result())
print("helper, main blocked in result() : %.2f ms" % ctime(20))
x = torch.randn(256, 256, device="cuda")
for _ in range(20):
    f = pool.submit(one)
    while not f.done(): y = x @ x
    lib.pnmn_plan_destroy(f.result())
torch.cuda.synchronize()
print("helper, main launching kernels   : %.2f ms" % ctime(20))
for _ in range(20):
    f = pool.submit(one)
    k = 0
    while not f.done(): k += 1
    lib.pnmn_plan_destroy(f.result())
print("helper, main spinning in Python  : %.2f ms" % ctime(20))
for _ in range(20):
    f = pool.submit(one)
    while not f.done(): time.sleep(0.0002)
    lib.pnmn_plan_destroy(f.result())
print("helper, main sleeping            : %.2f ms" % ctime(20))
