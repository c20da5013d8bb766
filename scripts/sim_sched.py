"""Offline schedule analysis of the persistent executor (no GPU needed).

Rebuilds the bench workload's plan on the host (pnmn_plan_create runs without a device), reads the task dependency
lists (pnmn_debug_plan_meta) and replays the list with the in-order blocking fetch exec.cu uses.  Task durations come
from a GPU trace (scripts/trace_exec.py -> gpurun_out/trace.npy): when the trace belongs to the SAME plan (same task
count) the measured per-task times are used, otherwise the mean time per task class (type, samples, M tiles, MMAs,
flags) of the trace, which lets a trace recorded before a scheduling change predict the effect of that change.

  python scripts/sim_sched.py            # critical path, replay under three fetch policies
  CHAIN=1 python scripts/sim_sched.py    # also the chain of tasks that finishes last
  PNMN_NOSTRANDS=1 python scripts/sim_sched.py   # one chain per sample (the pre-strand scheduler)
"""
import ctypes, heapq, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probnmn_clevr_b200 import _lib as L
from probnmn_clevr_b200.nmn import NeuralModuleNetwork
from probnmn_clevr_b200.synthetic import ProgramSampler
from probnmn_clevr_b200.vocabulary import Vocabulary

B = int(os.environ.get("B", 256))
SLOTS = int(os.environ.get("SLOTS", 296))
VIS = float(os.environ.get("VIS", 1.0))  # us between a producer's publish and a polling consumer seeing it


def plan_metas():
    vocab = Vocabulary.clevr()
    m = NeuralModuleNetwork(vocab)
    m._ensure_flat()
    progs = ProgramSampler(vocab, seed=0).sample(B, 40)
    plan = m._compile(progs, True, None)
    lib = L.lib()
    lib.pnmn_debug_plan_meta.restype = ctypes.c_int64
    lib.pnmn_debug_plan_meta.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
    out = []
    for p in (0, 1):
        n = lib.pnmn_debug_plan_meta(plan, p, None, 0)
        buf = np.zeros((n, 16), dtype=np.int32)
        lib.pnmn_debug_plan_meta(plan, p, buf.ctypes.data, n)
        out.append(buf)
    return out


def simulate(deps, dur, policy, slots=SLOTS, vis=VIS, prio=None):
    """event-driven replay; deps[i] = list of producer ids (< i), dur[i] = body + publish time (us)"""
    n = len(dur)
    if policy == "inorder":
        fin = np.zeros(n)
        free = [0.0] * slots
        heapq.heapify(free)
        for i in range(n):
            t = heapq.heappop(free)
            ready = max([fin[j] + vis for j in deps[i]], default=0.0)
            fin[i] = max(t, ready) + dur[i]
            heapq.heappush(free, fin[i])
        return fin.max(), fin
    succ = [[] for _ in range(n)]
    for i, d in enumerate(deps):
        for j in d:
            succ[j].append(i)
    missing = np.array([len(d) for d in deps])
    key = (lambda i: i) if prio is None else (lambda i: (-prio[i], i))
    readyq = [(key(i), i) for i in range(n) if missing[i] == 0]
    heapq.heapify(readyq)
    events = []  # (time, kind, id): kind 0 = task becomes visible-ready, 1 = slot free
    free_slots, now, done, fin = slots, 0.0, 0, np.zeros(n)
    while done < n:
        while free_slots and readyq:
            _, i = heapq.heappop(readyq)
            free_slots -= 1
            fin[i] = now + dur[i]
            heapq.heappush(events, (fin[i], 1, i))
        now, kind, i = heapq.heappop(events)
        if kind == 1:
            free_slots += 1
            done += 1
            for s in succ[i]:
                missing[s] -= 1
                if missing[s] == 0:
                    heapq.heappush(events, (now + vis, 0, s))
        else:
            heapq.heappush(readyq, (key(i), i))
    return fin.max(), fin


def task_class(t):
    """trace row -> class key"""
    typ = int(t[5] & 0xff)
    if typ == 0:
        return (0, int((t[5] >> 8) & 0xff), int((t[5] >> 16) & 0xff), int(t[6]), int(t[7]))
    return (1, int(t[7]), 0, 0, 0)


def meta_class(row):
    if row[0] == 0:
        return (0, int(row[12]), int(row[13]), int(row[14]), int(row[15]))
    return (1, int(row[12]), 0, 0, 0)


def main():
    tr = np.load(os.environ.get("TRACE", "gpurun_out/trace.npy"))
    metas = plan_metas()
    n_f, n_b = len(metas[0]), len(metas[1])
    same_plan = len(tr) == n_f + n_b
    print(f"plan: {n_f} forward + {n_b} backward tasks; trace rows {len(tr)} -> {'measured per-task times' if same_plan else 'class-mean times of the trace'}")
    # class-mean durations from the trace (both passes pooled; late-phase tasks run faster than early ones, so this is a mean)
    ok_all, pub_all = tr[:, 1] / 1e3, tr[:, 3] / 1e3
    cls = {}
    for i in range(len(tr)):
        cls.setdefault(task_class(tr[i]), []).append(pub_all[i] - ok_all[i])
    cls_mean = {k: float(np.mean(v)) for k, v in cls.items()}
    conv_fallback = float(np.mean([d for k, v in cls.items() if k[0] == 0 for d in v]))
    off = 0
    for name, meta in zip(("forward", "backward"), metas):
        n = len(meta)
        deps = [[int(d) for d in row[2:2 + row[1]]] for row in meta]
        if same_plan:
            t = tr[off:off + n]
            t0 = t[:, 0].min()
            ok, pub = (t[:, 1] - t0) / 1e3, (t[:, 3] - t0) / 1e3
            dur = pub - ok
            print(f"== {name}: {n} tasks, measured span {pub.max():.0f} us")
        else:
            t = None
            dur = np.array([cls_mean.get(meta_class(r), conv_fallback if r[0] == 0 else 9.0) for r in meta])
            print(f"== {name}: {n} tasks")
        off += n
        cp = np.zeros(n)
        depth = np.zeros(n, dtype=int)
        for i in range(n):
            if deps[i]:
                cp[i] = max(cp[k] for k in deps[i]) + VIS
                depth[i] = max(depth[k] for k in deps[i]) + 1
            cp[i] += dur[i]
        rem = dur.copy()
        for i in range(n - 1, -1, -1):
            for j in deps[i]:
                rem[j] = max(rem[j], dur[j] + VIS + rem[i])
        print(f"   sum(dur)/{SLOTS} slots = {dur.sum() / SLOTS:.0f} us; critical path (+{VIS} us visibility per stage) = {cp.max():.0f} us, depth {depth.max() + 1} stages")
        for pol, pr in (("inorder", None), ("ready", None), ("ready-cp", rem)):
            print(f"   replay {pol:9s} ({SLOTS} slots): {simulate(deps, dur, pol, SLOTS, VIS, pr)[0]:.0f} us")
        for scale in (0.8, 0.6):
            print(f"   replay inorder with bodies x{scale}: {simulate(deps, dur * scale, 'inorder')[0]:.0f} us")
        w = np.bincount(depth)
        print("   tasks per stage:", [int(x) for x in w[:12]], "...", [int(x) for x in w[-6:]])
        if os.environ.get("CHAIN") and same_plan:
            last = int(np.argmax(pub))
            chain = [last]
            while deps[chain[-1]]:
                chain.append(max(deps[chain[-1]], key=lambda k: pub[k]))
            chain.reverse()
            print(f"   chain that finishes last ({len(chain)} stages)")
            prev = 0.0
            for i in chain:
                typ = t[i, 5] & 0xff
                desc = (f"conv ns={(t[i,5]>>8)&0xff} nmt={(t[i,5]>>16)&0xff} mmas={t[i,6]} flags={int(t[i,7]):#05x}" if typ == 0 else f"elt op={t[i,7]}")
                print(f"     task {i:6d} {desc:45s} producers done {prev:7.1f} | fetched {(t[i,0]-t0)/1e3:7.1f} | start {ok[i]:7.1f} (+{ok[i]-prev:5.1f}) | run {pub[i]-ok[i]:5.1f} | published {pub[i]:7.1f}")
                prev = pub[i]


if __name__ == "__main__":
    main()
